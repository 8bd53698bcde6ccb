"""Kernel branches and host hooks no BASELINE configuration reaches:

* the Corrector's alpha branch (corrector.cc:105-130: rho'' > 0), with a user-defined convex
  loss - every library loss has rho'' <= 0;
* the one documented semantic difference of the masked-lane Jets (ceres/jet.h header): the
  derivative of a CONSTANT sub-expression is exactly zero, where the reference's dense Jets
  compute 0 * inf = NaN and its evaluator then rejects the evaluation;
* ceres::EvaluationCallback (program_evaluator_cuda.h:116-121);
* a second engine on the same device and repeated create / destroy (launch configuration is
  per launch, not cached per process).
"""
import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B, problems as P

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)


def _no_radial_problem(loss_kind, a, n_cam=7, n_pts=160, n_obs=800, seed=9):
    """BAL-shaped problem on SnavelyReprojectionErrorNoRadialDistortion<2,7,3>."""
    base = P.bal_problem(n_cam, n_pts, n_obs, seed=seed)
    cams = base.pb_values[3 * n_pts:].reshape(n_cam, 9)[:, :7]
    return P.ProblemSpec(
        pb_size=np.concatenate([np.full(n_pts, 3, np.int32), np.full(n_cam, 7, np.int32)]),
        pb_values=np.concatenate([base.pb_values[:3 * n_pts], cams.ravel()]),
        rb_type=np.full(n_obs, P.SNAVELY_NO_RADIAL, np.int32), rb_pb=base.rb_pb, fdata=base.fdata,
        rb_loss_kind=np.full(n_obs, loss_kind, np.int32), rb_loss_a=np.full(n_obs, a),
        rb_loss_b=np.zeros(n_obs), num_eliminate_blocks=n_pts)


@pytest.mark.parametrize("fmt", [0, 1])
def test_corrector_alpha_branch_with_a_convex_loss(fmt):
    spec = _no_radial_problem(P.LOSS_CONVEX_TEST, 0.05)
    cp = B.CudaProblem(spec, jacobian_format=fmt)
    op = O.OracleProblem(spec, jacobian_format=fmt)
    x = op.initial_state()
    ok, c, r, g, j = cp.evaluate(x)
    ok_o, c_o, r_o, g_o, j_o = op.evaluate(x)
    assert ok and ok_o
    # the branch really ran: corrected residuals differ from sqrt(rho') r (alpha != 0)
    _, _, r_raw, _, _ = op.evaluate(x, apply_loss_function=False)
    s = (r_raw.reshape(-1, 2) ** 2).sum(axis=1)
    plain = r_raw.reshape(-1, 2) * np.sqrt(1.0 + 0.1 * s)[:, None]
    assert _rel(r_o.reshape(-1, 2), plain) > 1e-3
    assert abs(c - c_o) <= 1e-10 * abs(c_o)
    assert _rel(r, r_o) <= 1e-12
    assert _rel(j[:op.num_jacobian_values], j_o[:op.num_jacobian_values]) <= 1e-12
    assert _rel(g, g_o) <= 1e-10
    cp.close()


def test_derivative_of_a_constant_subexpression_is_zero():
    # r = x + sqrt(T(c)).  c > 0: identical to the reference.  c == 0: d sqrt(c) = (1 / (2
    # sqrt(0))) * 0 is NaN in the reference's dense Jets, so ITS evaluator rejects the
    # evaluation; here a constant carries no derivative lanes at all, its derivative is the
    # exact zero it is mathematically, and the evaluation is valid with dr/dx = 1.
    for c, reference_accepts in ((4.0, True), (0.0, False)):
        b = P.ProblemBuilder()
        x0 = b.add_parameter_block([1.5])
        b.add_residual_block(P.SQRT_OF_CONSTANT, [x0], [c])
        spec = b.build()
        cp = B.CudaProblem(spec, jacobian_format=0)
        op = O.OracleProblem(spec, jacobian_format=0)
        ok, cost, r, g, j = cp.evaluate(np.array([1.5]))
        ok_o, *_ = op.evaluate(np.array([1.5]))
        assert ok_o == reference_accepts
        assert ok
        assert r[0] == 1.5 + np.sqrt(c) and j[0] == 1.0 and g[0] == r[0]
        cp.close()
    # a non-finite value that does reach an output is still rejected
    b = P.ProblemBuilder()
    x0 = b.add_parameter_block([1.5])
    b.add_residual_block(P.SQRT_OF_CONSTANT, [x0], [-1.0])  # sqrt(-1) = NaN residual
    cp = B.CudaProblem(b.build(), jacobian_format=0)
    ok, *_ = cp.evaluate(np.array([1.5]))
    assert not ok
    cp.close()


def test_evaluation_callback_sees_every_evaluation_point():
    spec = P.bal_problem(6, 120, 500, seed=3)
    cp = B.CudaProblem(spec, jacobian_format=0, evaluation_callback=True)
    x = cp.initial_state()
    before = cp.callback_info()["calls"]
    x1 = x + 0.001
    ok, *_ = cp.evaluate(x1)
    info = cp.callback_info()
    assert ok and info["calls"] == before + 1
    assert info["evaluate_jacobians"] and info["new_evaluation_point"]
    # the user's parameter blocks held the evaluation point when the callback ran
    assert abs(info["user_value_sum"] - float(x1.sum())) <= 1e-9 * abs(float(x1.sum()))
    ok, *_ = cp.evaluate(x1, gradient=False, jacobian=False, new_evaluation_point=False)
    info = cp.callback_info()
    assert ok and info["calls"] == before + 2
    assert not info["evaluate_jacobians"] and not info["new_evaluation_point"]
    cp.close()


def test_engines_come_and_go_on_one_device():
    spec = P.bal_problem(6, 200, 900, seed=5)
    op = O.OracleProblem(spec, jacobian_format=0)
    x = op.initial_state()
    _, c_o, r_o, g_o, j_o = op.evaluate(x)
    first = B.CudaProblem(spec, jacobian_format=0)
    for _ in range(3):
        second = B.CudaProblem(spec, jacobian_format=1)
        ok, c, *_ = second.evaluate(x)
        assert ok and abs(c - c_o) <= 1e-10 * abs(c_o)
        second.close()
    ok, c, r, g, j = first.evaluate(x)
    assert ok and abs(c - c_o) <= 1e-10 * abs(c_o) and _rel(r, r_o) <= 1e-12
    first.close()
