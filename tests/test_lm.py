"""A full Levenberg-Marquardt solve driven by the evaluator (ceres-solver-cuda_b200/lm.py):
CPU test with the oracle as evaluator (host logic), GPU test with the CUDA evaluator, both
converging to the same final cost."""
import numpy as np
import torch
import pytest

import oracle_py as O
from ceres_b200 import binding as B, lm, problems as P


def _problem(seed=31):
    spec = P.bal_problem(12, 600, 2600, seed=seed)
    # start away from the optimum so LM has work to do
    rng = np.random.default_rng(seed)
    spec.pb_values[:] += rng.normal(0, 0.02, spec.pb_values.size) * (np.abs(spec.pb_values) < 50)
    return spec


def test_lm_with_oracle_evaluator_converges():
    spec = _problem()
    op = O.OracleProblem(spec, jacobian_format=1)
    out = lm.solve(op, max_num_iterations=15)
    costs = [h["cost"] for h in out["iterations"] if h["accepted"]]
    assert out["cost"] < 0.25 * out["initial_cost"]
    assert all(b <= a for a, b in zip(costs, costs[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("subset_manifold", [False, True])
def test_lm_cuda_evaluator_reaches_the_same_cost(subset_manifold):
    spec = _problem()
    if subset_manifold:
        ncam = spec.meta["num_cameras"]
        spec.pb_manifold_kind[-ncam:] = P.MANIFOLD_SUBSET
        spec.pb_manifold_param[-ncam:] = 0b1000001   # hold rx and focal fixed
    op = O.OracleProblem(spec, jacobian_format=1)
    cp = B.CudaProblem(spec, jacobian_format=1)
    ref = lm.solve(op, max_num_iterations=15)
    out = lm.solve(cp, max_num_iterations=15)
    assert out["cost"] < 0.25 * out["initial_cost"]
    assert len(out["iterations"]) == len(ref["iterations"])
    # measured: 2e-13 (cost) and 6e-12 (state) after 15 iterations
    assert abs(out["cost"] - ref["cost"]) <= 1e-10 * ref["cost"]
    assert np.max(np.abs(out["x"] - ref["x"])) <= 1e-6 * np.max(np.abs(ref["x"]))


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["ITERATIVE_SCHUR", "CGNR"])
def test_cxx_solve_api(solver):
    """ceres::Solve(options, ProblemCUDA*, summary) — the reference's entry point
    (problem_cuda.h:490-502) — with a linear-solver ordering (points before cameras) as
    examples/bundle_adjuster.cu.cc sets it.  The stand-in linear solver is inexact CG, so
    the final cost is compared loosely with the exact-solve LM of lm.py."""
    spec = _problem()
    npts = spec.meta["num_points"]
    ordering = np.r_[np.zeros(npts, np.int32), np.ones(spec.num_pb - npts, np.int32)]
    out = B.solve(spec, getattr(B, solver), max_num_iterations=30,
                  ordering=ordering if solver == "ITERATIVE_SCHUR" else None)
    assert out["usable"], out["message"]
    ref = lm.solve(O.OracleProblem(spec, jacobian_format=1), max_num_iterations=30)
    assert out["final_cost"] < 0.2 * out["initial_cost"]
    assert abs(out["initial_cost"] - ref["initial_cost"]) <= 1e-10 * ref["initial_cost"]
    assert out["final_cost"] <= 1.05 * ref["cost"]
    # the solution is written back into the user's parameter blocks
    op = O.OracleProblem(P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=out["x"], rb_type=spec.rb_type, rb_pb=spec.rb_pb,
        fdata=spec.fdata, rb_loss_kind=spec.rb_loss_kind, rb_loss_a=spec.rb_loss_a,
        rb_loss_b=spec.rb_loss_b, num_eliminate_blocks=spec.num_eliminate_blocks))
    assert abs(op.evaluate(residuals=False, gradient=False, jacobian=False)[1] -
               out["final_cost"]) <= 1e-9 * out["final_cost"]


def test_cxx_options_reject_line_search():
    """solver.cc:702-708: the CUDA evaluator supports TRUST_REGION minimizers only."""
    spec = P.bal_problem(2, 5, 10, seed=1)
    ok, msg = B.options_is_valid(spec, 0)   # LINE_SEARCH
    assert not ok and "TRUST_REGION" in msg
    ok, msg = B.options_is_valid(spec, 1)
    assert ok and msg == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cxx_solve_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device ceres::Solve reports FAILURE and leaves the
    parameters untouched."""
    spec = P.bal_problem(2, 5, 10, seed=1)
    out = B.solve(spec, B.CGNR, max_num_iterations=3)
    assert not out["usable"] and out["message"]
    assert np.array_equal(out["x"], spec.pb_values)
