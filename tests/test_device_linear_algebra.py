"""The device-resident Jacobian: products, column norms,
column scaling and the conjugate-gradient solve on the normal equations run on the values
the evaluation kernel left in HBM.  Checked against the oracle's dense Jacobian with numpy /
scipy on the same seeded problems, both value layouts, manifolds and constant blocks."""
import numpy as np
import pytest
import torch

import oracle_py as O
from ceres_b200 import binding as B, lm, problems as P

gpu = pytest.mark.gpu

RTOL = 1e-11  # sums of a few dozen products of values that agree to 1e-12


def _problems():
    bal = P.bal_problem(7, 90, 400, seed=11)
    sub = P.bal_problem(6, 60, 260, seed=12, subset_manifold=True, constant_cameras=1)
    pose = P.pose_graph_problem(40, 120, seed=13)
    return {"bal": bal, "bal_subset_constant": sub, "pose_graph": pose}


def _dense(spec, fmt):
    op = O.OracleProblem(spec, jacobian_format=fmt)
    ok, cost, r, g, values = op.evaluate(op.initial_state())
    assert ok
    return op, op.dense_jacobian(values), r


def _rel(a, b):
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)


@gpu
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("name", ["bal", "bal_subset_constant", "pose_graph"])
def test_products_norms_and_scaling(name, fmt):
    spec = _problems()[name]
    op, J, r = _dense(spec, fmt)
    cp = B.CudaProblem(spec, jacobian_format=fmt)
    ok, *_ = cp.evaluate(op.initial_state())
    assert ok
    rng = np.random.default_rng(5)
    x = rng.normal(size=J.shape[1])
    w = rng.normal(size=J.shape[0])
    assert _rel(cp.jacobian_multiply(x), J @ x) <= RTOL
    assert _rel(cp.jacobian_multiply(w, transpose=True), J.T @ w) <= RTOL
    assert _rel(cp.jacobian_squared_column_norm(), (J * J).sum(axis=0)) <= RTOL
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    cp.jacobian_scale_columns(scale)
    assert _rel(cp.jacobian_multiply(x), (J * scale) @ x) <= RTOL
    # the next evaluation rewrites every value: the scaling does not accumulate
    cp.evaluate(op.initial_state())
    assert _rel(cp.jacobian_multiply(x), J @ x) <= RTOL


@gpu
@pytest.mark.parametrize("fmt", [0, 1])
def test_cgnr_solves_the_damped_normal_equations(fmt):
    spec = _problems()["bal"]
    op, J, r = _dense(spec, fmt)
    cp = B.CudaProblem(spec, jacobian_format=fmt)
    cp.evaluate(op.initial_state())
    d2 = (J * J).sum(axis=0) / 1e4
    exact = np.linalg.solve(J.T @ J + np.diag(d2), J.T @ r)
    y, s = cp.cgnr_solve(d2, max_iterations=2000, r_tolerance=1e-13, q_tolerance=-1.0)
    assert s["termination"] == 0, s
    assert _rel(y, exact) <= 1e-7
    assert abs(s["gradient_norm"] - np.linalg.norm(J.T @ r)) <= 1e-10 * s["gradient_norm"]
    Jy = J @ y
    assert abs(s["jy_dot_b"] - Jy @ r) <= 1e-9 * abs(Jy @ r)
    assert abs(s["jy_squared_norm"] - Jy @ Jy) <= 1e-9 * (Jy @ Jy)
    # the reference's stopping rule (relative decrease of the quadratic model) stops early
    y2, s2 = cp.cgnr_solve(d2, max_iterations=2000, r_tolerance=-1.0, q_tolerance=0.1)
    assert s2["termination"] == 0 and 0 < s2["iterations"] < s["iterations"]
    # iteration limit is reported
    _, s3 = cp.cgnr_solve(d2, max_iterations=2, r_tolerance=1e-13, q_tolerance=-1.0)
    assert s3["termination"] == 1 and s3["iterations"] == 2


@gpu
def test_linear_algebra_needs_a_device_jacobian():
    spec = _problems()["bal"]
    cp = B.CudaProblem(spec)
    with pytest.raises(RuntimeError, match="no Jacobian on the device"):
        cp.jacobian_squared_column_norm()
    cp.evaluate(jacobian=False)
    with pytest.raises(RuntimeError, match="no Jacobian on the device"):
        cp.jacobian_multiply(np.zeros(cp.num_effective_parameters))


@gpu
@pytest.mark.parametrize("subset_manifold", [False, True])
def test_device_resident_solve_reaches_the_exact_solve_cost(subset_manifold):
    """ceres::Solve with CGNR + CUDA_SPARSE keeps the Jacobian in HBM for the whole solve
    (reference: cgnr_solver.cc CudaCgnrSolver).  Inexact steps, so the final cost is compared
    loosely with the exact-step LM on the oracle."""
    spec = P.bal_problem(12, 600, 2600, seed=31)
    rng = np.random.default_rng(31)
    spec.pb_values[:] += rng.normal(0, 0.02, spec.pb_values.size) * (np.abs(spec.pb_values) < 50)
    if subset_manifold:
        ncam = spec.meta["num_cameras"]
        spec.pb_manifold_kind[-ncam:] = P.MANIFOLD_SUBSET
        spec.pb_manifold_param[-ncam:] = 0b1000001
    out = B.solve(spec, B.CGNR, max_num_iterations=40, cuda_sparse=True)
    assert out["usable"], out["message"]
    ref = lm.solve(O.OracleProblem(spec, jacobian_format=1), max_num_iterations=40)
    assert abs(out["initial_cost"] - ref["initial_cost"]) <= 1e-10 * ref["initial_cost"]
    assert out["final_cost"] <= 1.02 * ref["cost"]
    host = B.solve(spec, B.CGNR, max_num_iterations=40, cuda_sparse=False)
    assert abs(out["final_cost"] - host["final_cost"]) <= 0.02 * host["final_cost"]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bal", "bal_subset", "pose_graph"])
def test_trust_region_loop_on_the_device_matches_the_host_loop(kind, monkeypatch):
    """With CGNR + CUDA_SPARSE the whole iteration runs in HBM (cb200_engine_trust_region_step:
    LM diagonal, conjugate gradients, Program::Plus with the manifolds, norms).  The host loop
    around the same device Jacobian (CB200_HOST_TRUST_REGION=1: Plus, diagonals and norms on
    the host, state and step over PCIe) must take the same steps."""
    if kind == "pose_graph":
        spec = P.pose_graph_problem(120, 420, seed=8)     # EigenQuaternion x R^3 Plus, pose 0 fixed
    else:
        spec = P.bal_problem(12, 600, 2600, seed=33, subset_manifold=kind == "bal_subset")
        rng = np.random.default_rng(33)
        spec.pb_values[:] += rng.normal(0, 0.02, spec.pb_values.size) * (np.abs(spec.pb_values) < 50)
    # near-exact steps: with the default eta = 0.1 conjugate gradients stop on the model
    # decrease after a handful of iterations, the stopping iteration depends on the last bits
    # of the (atomic) sums and the two trajectories drift apart (measured up to 6e-5 on the
    # final cost, run to run)
    monkeypatch.setenv("CB200_DRIVER_ETA", "1e-9")
    dev = B.solve(spec, B.CGNR, max_num_iterations=12, cuda_sparse=True)
    monkeypatch.setenv("CB200_HOST_TRUST_REGION", "1")
    host = B.solve(spec, B.CGNR, max_num_iterations=12, cuda_sparse=True)
    assert dev["usable"] and host["usable"], (dev["message"], host["message"])
    assert dev["iterations"] == host["iterations"]
    assert dev["successful_steps"] == host["successful_steps"]
    assert abs(dev["initial_cost"] - host["initial_cost"]) <= 1e-12 * host["initial_cost"]
    assert abs(dev["final_cost"] - host["final_cost"]) <= 1e-5 * host["final_cost"]
    assert np.max(np.abs(dev["x"] - host["x"])) <= 1e-3 * np.max(np.abs(host["x"]))
    assert dev["final_cost"] < 0.9 * dev["initial_cost"]


@pytest.mark.gpu
@pytest.mark.parametrize("with_ordering", [False, True])
def test_resident_solve_runs_on_the_per_type_kernels(with_ordering, monkeypatch, capfd):
    """CGNR + CUDA_SPARSE keeps the Jacobian in HBM in a two-region layout (the ordering's
    first group, or - without an ordering - the parameter blocks that only appear in one
    argument slot), so that the evaluation kernel's bulk stores and the per-type linear
    algebra kernel apply.  The engine reports every fall-back to its table-walk kernels under
    CB200_VERBOSE; a bundle adjustment problem must not produce one."""
    spec = P.bal_problem(9, 400, 1800, seed=21)
    ordering = None
    if with_ordering:
        ordering = np.array([0 if s == 3 else 1 for s in spec.pb_size], dtype=np.int32)
    monkeypatch.setenv("CB200_VERBOSE", "1")
    out = B.solve(spec, B.CGNR, max_num_iterations=5, cuda_sparse=True, ordering=ordering)
    captured = capfd.readouterr()
    assert out["usable"], out["message"]
    assert out["final_cost"] < 0.5 * out["initial_cost"]
    assert "table-walk" not in captured.err, captured.err
    assert "declined" not in captured.err, captured.err
    # the same solve on the table-walk kernels takes the same steps
    monkeypatch.setenv("CB200_GENERIC_NORMAL_PRODUCT", "1")
    monkeypatch.setenv("CB200_DRIVER_ETA", "1e-9")
    generic = B.solve(spec, B.CGNR, max_num_iterations=5, cuda_sparse=True, ordering=ordering)
    monkeypatch.delenv("CB200_GENERIC_NORMAL_PRODUCT")
    typed = B.solve(spec, B.CGNR, max_num_iterations=5, cuda_sparse=True, ordering=ordering)
    assert generic["iterations"] == typed["iterations"]
    assert abs(generic["final_cost"] - typed["final_cost"]) <= 1e-6 * typed["final_cost"]


def test_abi_exports_the_linear_algebra_entry_points():
    lib = B.abi()
    for name in ("cb200_engine_jacobian_multiply", "cb200_engine_jacobian_squared_column_norm",
                 "cb200_engine_jacobian_scale_columns", "cb200_engine_cgnr_solve"):
        assert hasattr(lib, name)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_device_resident_solve_fails_loudly_without_a_gpu():
    spec = P.bal_problem(2, 5, 10, seed=1)
    out = B.solve(spec, B.CGNR, max_num_iterations=3, cuda_sparse=True)
    assert not out["usable"] and out["message"]
