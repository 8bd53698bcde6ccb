"""Pins the oracle against the REFERENCE's own arithmetic: tests/golden/reference_arith.npz
was produced by the reference headers (jet.h, rotation.h, autodiff.h, corrector.h,
loss_function_cuda.h, snavely_reprojection_error.h) compiled from /root/reference
(tests/golden/make_golden.py, oracle/ref_arith.cc).  CPU only.

Tolerance: a few ulp.  The restatement follows the reference operation by operation;
what remains is libm's std::hypot(x, y, z) vs the restated call and summation order."""
import os

import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import problems as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_arith.npz"))


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_snavely_matches_reference():
    for i in range(G["snavely_cam"].shape[0]):
        ok, r, J = O.cost_evaluate(P.SNAVELY, G["snavely_obs"][i],
                                   [G["snavely_cam"][i], G["snavely_pt"][i]])
        assert ok
        assert _rel(r, G["snavely_res"][i]) <= 1e-13
        assert _rel(J[0].ravel(), G["snavely_jcam"][i]) <= 1e-13
        assert _rel(J[1].ravel(), G["snavely_jpt"][i]) <= 1e-13


def test_snavely_quaternions_matches_reference():
    for i in range(G["quat_cam"].shape[0]):
        ok, r, J = O.cost_evaluate(P.SNAVELY_QUAT, G["snavely_obs"][i],
                                   [G["quat_cam"][i], G["snavely_pt"][i]])
        assert ok
        assert _rel(r, G["quat_res"][i]) <= 1e-13
        assert _rel(J[0].ravel(), G["quat_jcam"][i]) <= 1e-13
        assert _rel(J[1].ravel(), G["quat_jpt"][i]) <= 1e-13


def test_losses_match_reference_exactly():
    for kind, a, b, s, r0, r1, r2 in G["loss_cases"]:
        rho = O.loss_evaluate(int(kind), a, b, s)
        assert rho.tolist() == [r0, r1, r2], (kind, a, b, s)


def test_corrector_matches_reference():
    for cin, cout in zip(G["corrector_in"], G["corrector_out"]):
        s, rho, r, J = cin[0], cin[1:4], cin[4:6], cin[6:]
        rr, JJ = O.corrector(s, rho, r, J)
        assert np.array_equal(rr, cout[:2])
        assert np.array_equal(JJ, cout[2:])


def test_angle_axis_rotate_point_matches_reference():
    for aa, pt, rot in zip(G["aa"], G["aa_pt"], G["aa_rot"]):
        assert _rel(O.angle_axis_rotate_point(aa, pt), rot) <= 4e-16


def test_quaternion_to_angle_axis_jet_matches_reference():
    for q, v, j in zip(G["q2aa_q"], G["q2aa_value"], G["q2aa_jac"]):
        ov, oj = O.quaternion_to_angle_axis_jet(q)
        assert _rel(ov, v) <= 1e-15 and _rel(oj, j) <= 1e-14


def test_jet_operations_match_reference():
    """Inputs of internal/ceres/jet_cuda_test.cu.cc:108-110 (x = 2.3, y = 1.7) and two
    more points; 40 Jet operations each."""
    for xy, want in zip(G["jet_xy"], G["jet_battery"]):
        got = O.jet_battery(xy)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-15, atol=0)
