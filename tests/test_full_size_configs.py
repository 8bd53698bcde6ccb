"""BASELINE.json configs 2, 4 and 5 at their full sizes (beyond what the oracle evaluates in
seconds), checked like the headline configuration in test_full_size_properties.py through
size-independent properties:

* a prefix of the problem (the first residual blocks, every parameter block kept in place)
  evaluated by the oracle reproduces the corresponding residuals and Jacobian values of the
  full evaluation to 1e-12 (north_star tolerance for residuals / Jacobian);
* linearity: the gradient the evaluation kernel accumulates equals J' r computed by the
  independent device product kernel from the Jacobian and residuals it wrote (1e-10);
* cost = sum of rho(|r_raw|^2) / 2 recomputed on the host from the uncorrected residuals;
* two evaluations give bit-identical residuals / Jacobian / cost.

config 2: synthetic BAL 1778 x 993923 (5,001,946 blocks), Huber, BlockSparse
config 4: synthetic BAL 13682 x 4456117 + SubsetManifold(9, {0}) on every camera,
          CompressedRowSparseMatrix Jacobian (manifold projection + compressed-row scatter)
config 5: pose graph, 2.5 M poses / 10 M edges, RelativePoseError<6,7,7>,
          ProductManifold<EigenQuaternion, Euclidean<3>>, pose 0 constant, BlockSparse
"""
import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B, problems as P

pytestmark = pytest.mark.gpu

PREFIX = 150_000


def _prefix(spec, n):
    _, sizes, flen = P.COST_TYPES[int(spec.rb_type[0])]
    return P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[:n],
        rb_pb=spec.rb_pb[:len(sizes) * n], fdata=spec.fdata[:flen * n],
        pb_constant=spec.pb_constant, pb_manifold_kind=spec.pb_manifold_kind,
        pb_manifold_param=spec.pb_manifold_param, rb_loss_kind=spec.rb_loss_kind[:n],
        rb_loss_a=spec.rb_loss_a[:n], rb_loss_b=spec.rb_loss_b[:n],
        num_eliminate_blocks=spec.num_eliminate_blocks)


def _rel(a, b):
    return float(np.max(np.abs(a - b))) / float(np.max(np.abs(b)))


def _evaluate(spec, fmt):
    cp = B.CudaProblem(spec, jacobian_format=fmt)
    x = cp.initial_state()
    ok, cost, r, g, j = cp.evaluate(x)
    assert ok
    return cp, x, cost, r.copy(), g.copy(), j


def _common_properties(cp, x, cost, r, g, j, kres):
    # gradient = J' r through the independent device product kernel
    jt_r = cp.jacobian_multiply(r, transpose=True)
    assert float(np.max(np.abs(jt_r - g))) <= 1e-10 * float(np.max(np.abs(g)))
    # reproducibility
    head = min(j.size, 40_000_000)
    j_first = j[:head].copy()
    ok, cost2, r2, g2, j2 = cp.evaluate(x)
    assert ok and cost2 == cost
    assert np.array_equal(r2, r) and np.array_equal(j2[:head], j_first)
    assert float(np.max(np.abs(g2 - g))) <= 1e-12 * float(np.max(np.abs(g)))


def test_config2_bal_m_block_sparse():
    spec = P.bal_shape("M")
    cp, x, cost, r, g, j = _evaluate(spec, 0)
    try:
        n = PREFIX
        op = O.OracleProblem(_prefix(spec, n), jacobian_format=0, reduce=False)
        ok, c_o, r_o, g_o, j_o = op.evaluate(op.initial_state())
        assert ok
        assert _rel(r[:2 * n], r_o) <= 1e-12
        # block-sparse values: the E cells (points, 2x3) of the first n blocks come first
        assert _rel(j[:6 * n], j_o[:6 * n]) <= 1e-12
        f_full, f_pre = 6 * spec.num_rb, 6 * n
        assert _rel(j[f_full:f_full + 18 * n], j_o[f_pre:f_pre + 18 * n]) <= 1e-12
        # cost = sum of the Huber losses of the uncorrected residuals
        ok, cost_raw, r_raw, _, _ = cp.evaluate(x, gradient=False, jacobian=False,
                                                apply_loss_function=False)
        s = (r_raw.reshape(-1, 2) ** 2).sum(axis=1)
        rho = np.where(s > 1.0, 2.0 * np.sqrt(s) - 1.0, s)
        assert abs(cost - 0.5 * rho.sum()) <= 1e-10 * cost
        _common_properties(cp, x, cost, r, g, j, 2)
    finally:
        cp.close()


def test_config4_bal_l_subset_manifold_compressed_row():
    spec = P.bal_shape("L", subset_manifold=True)
    cp, x, cost, r, g, j = _evaluate(spec, 1)
    try:
        assert cp.num_effective_parameters == 3 * 4456117 + 8 * 13682
        n = PREFIX
        op = O.OracleProblem(_prefix(spec, n), jacobian_format=1, reduce=False)
        ok, c_o, r_o, g_o, j_o = op.evaluate(op.initial_state())
        assert ok
        assert _rel(r[:2 * n], r_o) <= 1e-12
        # compressed rows: 3 point + 8 camera columns per row, rows in residual order
        assert _rel(j[:22 * n], j_o[:22 * n]) <= 1e-12
        ok, cost_raw, r_raw, _, _ = cp.evaluate(x, gradient=False, jacobian=False,
                                                apply_loss_function=False)
        s = (r_raw.reshape(-1, 2) ** 2).sum(axis=1)
        rho = np.where(s > 1.0, 2.0 * np.sqrt(s) - 1.0, s)
        assert abs(cost - 0.5 * rho.sum()) <= 1e-10 * cost
        _common_properties(cp, x, cost, r, g, j, 2)
    finally:
        cp.close()


def test_config5_pose_graph_ten_million_edges():
    spec = P.pose_graph_problem(2_500_000, 10_000_000, seed=5)
    cp, x, cost, r, g, j = _evaluate(spec, 0)
    try:
        assert cp.num_residual_blocks == 10_000_000
        assert cp.num_effective_parameters == 6 * (2_500_000 - 1)  # pose 0 constant, 7 -> 6
        n = PREFIX
        op = O.OracleProblem(_prefix(spec, n), jacobian_format=0, reduce=False)
        ok, c_o, r_o, g_o, j_o = op.evaluate(op.initial_state())
        assert ok
        assert _rel(r[:6 * n], r_o) <= 1e-12
        # block-sparse values without eliminate blocks: cells in residual-block order
        nv = op.num_jacobian_values
        assert _rel(j[:nv], j_o[:nv]) <= 1e-12
        # no loss: cost = |r|^2 / 2
        assert abs(cost - 0.5 * float(r @ r)) <= 1e-10 * cost
        _common_properties(cp, x, cost, r, g, j, 6)
    finally:
        cp.close()
