"""The user-level entry points around the path: Problem::Evaluate (problem_impl.cc:599-760),
run here on the CUDA evaluator with a compressed-row Jacobian, and the host-side
Problem::EvaluateResidualBlock (problem_impl.cc:762-810) plus the query functions.  Checked
against the oracle on the same seeded problems."""
import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B, problems as P


def _spec(loss="huber"):
    return P.bal_problem(6, 40, 170, seed=21, subset_manifold=True, constant_cameras=1, loss=loss)


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b))) / max(float(np.max(np.abs(b))), 1e-300)


@pytest.mark.parametrize("apply_loss", [True, False])
def test_evaluate_residual_block_matches_the_oracle(apply_loss):
    spec = _spec()
    op = O.OracleProblem(spec, jacobian_format=1, reduce=False)
    ok, cost, r, g, values = op.evaluate(op.initial_state(), apply_loss_function=apply_loss)
    J = op.dense_jacobian(values)
    cp = B.CudaProblem(spec, with_device=False)
    tangent = [cp.tangent_size(i) for i in range(spec.num_pb)]
    col = np.concatenate([[0], np.cumsum(tangent)])
    total = 0.0
    for rb in range(spec.num_rb):
        ok, c, res, jac, ids = cp.evaluate_residual_block(rb, apply_loss_function=apply_loss)
        assert ok
        assert ids.tolist() == spec.rb_pb[2 * rb:2 * rb + 2].tolist()   # query functions
        total += c
        assert _rel(res, r[2 * rb:2 * rb + 2]) <= 1e-12 or np.max(np.abs(r[2 * rb:2 * rb + 2])) < 1e-9
        for j, pb in enumerate(ids):
            if spec.pb_constant[pb]:
                assert jac[j] is None
                continue
            block = J[2 * rb:2 * rb + 2, col[pb]:col[pb + 1]]
            assert np.max(np.abs(jac[j] - block)) <= 1e-12 * max(np.max(np.abs(block)), 1e-300)
    assert abs(total - cost) <= 1e-12 * cost
    # a Jacobian for a constant block is an error (problem_impl.cc:777-782) - the driver
    # passes null for those, so the cost-only form must still work on such a block
    ok, c, res, jac, ids = cp.evaluate_residual_block(0, jacobians=False)
    assert ok and all(j is None for j in jac)


@pytest.mark.gpu
@pytest.mark.parametrize("apply_loss", [True, False])
def test_problem_evaluate_with_every_block(apply_loss):
    """Default options: every parameter block in creation order; constant blocks keep their
    (zero) columns, as the reference's unreduced program does."""
    spec = _spec()
    op = O.OracleProblem(spec, jacobian_format=1, reduce=False)
    ok_o, c_o, r_o, g_o, v_o = op.evaluate(op.initial_state(), apply_loss_function=apply_loss)
    cp = B.CudaProblem(spec, with_device=False)
    ok, c, r, g, (rows, cols, vals), (nr, nc) = cp.problem_evaluate(apply_loss_function=apply_loss)
    assert ok and ok_o
    assert (nr, nc) == (op.num_residuals, op.num_effective_parameters)
    assert np.array_equal(rows, op.ints("crs_rows")) and np.array_equal(cols, op.ints("crs_cols")[:vals.size])
    assert _rel(vals, v_o[:vals.size]) <= 1e-12
    assert _rel(r, r_o) <= 1e-12 and _rel(g, g_o) <= 1e-10 and abs(c - c_o) <= 1e-10 * c_o
    # the constant camera's gradient entries are exactly zero
    tangent = [cp.tangent_size(i) for i in range(spec.num_pb)]
    col = np.concatenate([[0], np.cumsum(tangent)])
    constant = int(np.flatnonzero(spec.pb_constant)[0])
    assert np.all(g[col[constant]:col[constant + 1]] == 0.0)


@pytest.mark.gpu
def test_problem_evaluate_with_a_subset_of_blocks():
    """parameter_blocks = the points only: the cameras are held constant for the call and get
    no columns; residual_blocks = every other residual block."""
    spec = P.bal_problem(5, 30, 120, seed=22)
    npts = spec.meta["num_points"]
    keep = np.arange(0, spec.num_rb, 2, dtype=np.int32)
    sub = P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[keep],
        rb_pb=spec.rb_pb.reshape(-1, 2)[keep].ravel(), fdata=spec.fdata.reshape(-1, 2)[keep].ravel(),
        pb_constant=np.r_[np.zeros(npts, np.uint8), np.ones(spec.num_pb - npts, np.uint8)],
        rb_loss_kind=spec.rb_loss_kind[keep], rb_loss_a=spec.rb_loss_a[keep],
        rb_loss_b=spec.rb_loss_b[keep])
    op = O.OracleProblem(sub, jacobian_format=1, reduce=True)
    ok_o, c_o, r_o, g_o, v_o = op.evaluate(op.initial_state())
    # every point is still observed, so the reduced program keeps all of them
    assert op.num_effective_parameters == 3 * npts
    cp = B.CudaProblem(spec, with_device=False)
    ok, c, r, g, (rows, cols, vals), (nr, nc) = cp.problem_evaluate(
        parameter_blocks=np.arange(npts), residual_blocks=keep)
    assert ok and (nr, nc) == (op.num_residuals, 3 * npts)
    assert np.array_equal(rows, op.ints("crs_rows")) and np.array_equal(cols, op.ints("crs_cols")[:vals.size])
    assert _rel(vals, v_o[:vals.size]) <= 1e-12 and _rel(r, r_o) <= 1e-12
    assert _rel(g, g_o) <= 1e-10 and abs(c - c_o) <= 1e-10 * c_o
    # the temporary constness is undone: a full evaluation sees the cameras again
    ok, c2, r2, g2, _, (nr2, nc2) = cp.problem_evaluate()
    assert ok and nc2 == 3 * npts + 9 * (spec.num_pb - npts)


@pytest.mark.gpu
def test_problem_evaluate_cost_only_and_nothing():
    spec = _spec()
    op = O.OracleProblem(spec, jacobian_format=1, reduce=False)
    ok_o, c_o, *_ = op.evaluate(op.initial_state())
    cp = B.CudaProblem(spec, with_device=False)
    ok, c, r, g, _, _ = cp.problem_evaluate()
    assert ok and abs(c - c_o) <= 1e-10 * c_o


def test_problem_evaluate_fails_loudly_without_a_gpu():
    """No CPU fallback behind Problem::Evaluate either: without a device it returns false
    (and the blocks it held constant for the call are variable again)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    spec = P.bal_problem(5, 30, 120, seed=22)
    cp = B.CudaProblem(spec, with_device=False)
    ok, *_ = cp.problem_evaluate(parameter_blocks=np.arange(spec.meta["num_points"]))
    assert not ok
    # the cameras are not left constant: a Jacobian for one can still be requested
    ok, c, res, jac, ids = cp.evaluate_residual_block(0)
    assert ok and all(j is not None for j in jac)
