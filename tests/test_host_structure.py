"""CPU tests of the product's host layer: the C-ABI library loads and exports every
symbol include/ceres_b200.h declares; the reduced program, the BlockSparseMatrix /
CompressedRowSparseMatrix structure and the per-residual layouts are bit-exact
against the oracle (north_star: "Jacobian row/column structure and value offsets
bit-exact")."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B
from ceres_b200 import problems as P

STRUCT = ["residual_layout", "jacobian_per_residual_layout", "jacobian_per_residual_offsets",
          "program_rbs", "program_pbs", "constant_pbs"]
BSM = ["col_block_size", "col_block_pos", "row_block_size", "row_block_pos", "row_cells_start",
       "cell_block_id", "cell_position", "jacobian_layout_storage"]
CRS = ["crs_rows", "crs_cols"]


def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(B.ROOT, "include", "ceres_b200.h")).read()
    declared = set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"cb200_launch_fn"}
    assert declared == set(B.ABI_SYMBOLS), declared ^ set(B.ABI_SYMBOLS)
    lib = B.abi()
    for name in B.ABI_SYMBOLS:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.cb200_version()


def test_engine_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = B.abi()
    eng = ctypes.c_void_p()
    rc = lib.cb200_engine_create(0, ctypes.byref(eng))
    assert rc != 0 and not eng.value
    spec = P.bal_problem(3, 10, 30, seed=1)
    with pytest.raises(RuntimeError, match="no usable CUDA device|no CPU fallback"):
        B.CudaProblem(spec, with_device=True)
    cp = B.CudaProblem(spec, with_device=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cp.evaluate()


def _compare_structure(spec, fmt, **kw):
    op = O.OracleProblem(spec, jacobian_format=fmt, **kw)
    cp = B.CudaProblem(spec, jacobian_format=fmt, with_device=False, **kw)
    for a in ("num_parameters", "num_effective_parameters", "num_residuals",
              "num_residual_blocks", "num_parameter_blocks", "num_jacobian_values",
              "values_size", "num_constant_parameters"):
        assert getattr(op, a) == getattr(cp, a), a
    for name in STRUCT + (BSM if fmt == 0 else CRS):
        assert np.array_equal(op.ints(name), cp.ints(name)), name
    assert np.array_equal(op.pb_table(), cp.pb_table())
    assert np.array_equal(op.initial_state(), cp.initial_state())
    return op, cp


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("bulk", [False, True])
def test_bal_structure(fmt, bulk):
    spec = P.bal_problem(7, 60, 200, seed=4)
    op = O.OracleProblem(spec, jacobian_format=fmt)
    cp = B.CudaProblem(spec, jacobian_format=fmt, with_device=False, bulk=bulk)
    for name in STRUCT + (BSM if fmt == 0 else CRS):
        assert np.array_equal(op.ints(name), cp.ints(name)), name


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("nelim", [0, 13, 60])
def test_bal_structure_eliminate_blocks_and_reorder(fmt, nelim):
    spec = P.bal_problem(7, 60, 200, seed=5, constant_cameras=2, subset_manifold=True)
    # shuffle residual blocks so the Schur reordering has work to do
    rng = np.random.default_rng(0)
    perm = rng.permutation(spec.num_rb)
    spec = P.ProblemSpec(pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[perm],
                         rb_pb=spec.rb_pb.reshape(-1, 2)[perm].ravel(),
                         fdata=spec.fdata.reshape(-1, 2)[perm].ravel(),
                         pb_constant=spec.pb_constant, pb_manifold_kind=spec.pb_manifold_kind,
                         pb_manifold_param=spec.pb_manifold_param,
                         rb_loss_kind=spec.rb_loss_kind[perm], rb_loss_a=spec.rb_loss_a[perm],
                         rb_loss_b=spec.rb_loss_b[perm], num_eliminate_blocks=nelim)
    _compare_structure(spec, fmt, schur_reorder=True)


@pytest.mark.parametrize("fmt", [0, 1])
def test_fork_fixture_structure(fmt):
    op, cp = _compare_structure(P.evaluator_cuda_test_problem(), fmt)
    assert cp.num_residual_blocks == 5 and cp.num_residuals == 11
    assert abs(cp.fixed_cost - op.fixed_cost) <= 1e-15


@pytest.mark.parametrize("fmt", [0, 1])
def test_pose_graph_structure(fmt):
    _compare_structure(P.pose_graph_problem(40, 100, seed=3), fmt)


def test_all_constant_residual_block_goes_to_fixed_cost():
    b = P.ProblemBuilder()
    c = b.add_parameter_block(np.r_[0.01, 0.02, -0.01, 0.1, 0.2, -8.0, 500.0, 1e-7, 1e-13])
    p = b.add_parameter_block([0.3, -0.2, 0.5], constant=True)
    q = b.add_parameter_block([0.1, 0.4, -0.3], constant=True)
    b.add_residual_block(P.SNAVELY, [c, p], [3.0, -2.0], (P.LOSS_HUBER, 1.0))
    b.add_residual_block(P.POINT_DISPLACEMENT, [q], [0.5, 0.5, 0.5])
    spec = b.build()
    op, cp = _compare_structure(spec, 0)
    assert cp.num_residual_blocks == 1
    assert op.fixed_cost > 0 and abs(cp.fixed_cost - op.fixed_cost) <= 1e-15 * op.fixed_cost


def test_plus_matches_oracle():
    spec = P.pose_graph_problem(20, 40, seed=8)
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec, with_device=False)
    x = op.initial_state()
    d = np.random.default_rng(1).normal(0, 0.1, op.num_effective_parameters)
    assert np.allclose(op.plus(x, d), cp.plus(x, d), rtol=0, atol=1e-15)
    spec = P.bal_problem(5, 30, 90, seed=2, subset_manifold=True)
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec, with_device=False)
    x = op.initial_state()
    d = np.random.default_rng(2).normal(0, 0.1, op.num_effective_parameters)
    assert np.array_equal(op.plus(x, d), cp.plus(x, d))


def test_plus_on_several_threads_is_identical():
    """Program::Plus partitions the parameter blocks over the host threads
    (program.cc:121-150); the result does not depend on the thread count."""
    spec = P.pose_graph_problem(12000, 30000, seed=7)   # quaternion x R^3 manifolds
    cp = B.CudaProblem(spec, with_device=False)
    rng = np.random.default_rng(0)
    state = cp.initial_state()
    delta = rng.normal(0, 0.05, cp.num_effective_parameters)
    one = cp.plus(state, delta)
    many = cp.plus(state, delta, num_threads=8)
    assert np.array_equal(one, many)
    op = O.OracleProblem(spec)
    assert np.max(np.abs(op.plus(state, delta) - one)) <= 1e-15 * np.max(np.abs(one))


def random_mixed_problem(seed):
    rng = np.random.default_rng(100 + seed)
    b = P.ProblemBuilder()
    by_size = {}
    for size in (1, 2, 3, 4):
        for _ in range(int(rng.integers(3, 9)) + (10 if size == 1 else 0)):
            manifold = (P.MANIFOLD_NONE, 0)
            if size >= 2 and rng.random() < 0.3:
                mask = int(rng.integers(1, 2 ** size - 1))       # at least one free coordinate
                manifold = (P.MANIFOLD_SUBSET, mask)
            pb = b.add_parameter_block(rng.normal(size=size), manifold=manifold,
                                       constant=bool(rng.random() < 0.25))
            by_size.setdefault(size, []).append(pb)
    kinds = [P.AFFINE_1_3_234, P.AFFINE_1_3_432, P.AFFINE_1_2_23, P.AFFINE_2_3_24, P.AFFINE_3_4_34,
             P.BINARY_SCALAR, P.POINT_DISPLACEMENT, P.TEN_PARAMETER, P.PARAMETER_SENSITIVE]
    for _ in range(int(rng.integers(20, 60))):
        kind = kinds[int(rng.integers(len(kinds)))]
        nres, sizes, flen = P.COST_TYPES[kind]
        chosen, ok = [], True
        for s in sizes:
            free = [pb for pb in by_size[s] if pb not in chosen]
            if not free:
                ok = False
                break
            chosen.append(free[int(rng.integers(len(free)))])
        if not ok:
            continue
        b.add_residual_block(kind, chosen, fdata=rng.normal(size=flen))
    return b.build(num_eliminate_blocks=int(rng.integers(0, 6)))


@pytest.mark.parametrize("seed", range(12))
def test_random_mixed_problems_structure(seed):
    """Random problems mixing residual-block types with 1 to 10 arguments, constant blocks,
    subset manifolds and unused blocks: program, offsets and both Jacobian layouts equal the
    oracle's bit for bit, for the reduced program and for the unreduced one (whose constant
    blocks keep their columns, as Problem::Evaluate needs)."""
    spec = random_mixed_problem(seed)
    for fmt in (0, 1):
        for reduce in (True, False):
            _compare_structure(spec, fmt, reduce=reduce)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_random_mixed_problems_values(seed):
    """The same random problems evaluated on the device, reduced and unreduced, both layouts."""
    spec = random_mixed_problem(seed)
    for fmt in (0, 1):
        for reduce in (True, False):
            op = O.OracleProblem(spec, jacobian_format=fmt, reduce=reduce)
            cp = B.CudaProblem(spec, jacobian_format=fmt, reduce=reduce)
            state = op.initial_state()
            ok_o, c_o, r_o, g_o, j_o = op.evaluate(state)
            ok, c, r, g, j = cp.evaluate(state)
            assert ok == ok_o
            if not ok_o:
                continue
            scale = lambda v: max(float(np.max(np.abs(v))) if v.size else 0.0, 1e-300)
            assert abs(c - c_o) <= 1e-10 * max(abs(c_o), 1e-300)
            assert np.max(np.abs(r - r_o), initial=0.0) <= 1e-12 * scale(r_o)
            assert np.max(np.abs(g - g_o), initial=0.0) <= 1e-10 * scale(g_o)
            n = op.num_jacobian_values
            assert np.max(np.abs(j[:n] - j_o[:n]), initial=0.0) <= 1e-12 * scale(j_o[:n])


def test_argument_slot_ordering_finds_the_points():
    """ceres::Solve keeps the Jacobian of CGNR + CUDA_SPARSE in HBM in a two-region layout.
    Without a linear_solver_ordering the first region is the set of parameter blocks that only
    appear in one argument slot (internal::ArgumentSlotOrdering): the points of a bundle
    adjustment problem, constant blocks aside; a pose graph, whose poses appear in both slots,
    gets no ordering and keeps the row-by-row layout."""
    spec = P.bal_problem(9, 300, 1400, seed=3, constant_cameras=1)
    found, groups = B.argument_slot_ordering(spec)
    assert found
    sizes = np.asarray(spec.pb_size)
    assert np.all(groups[sizes == 3] == 0) and np.all(groups[sizes == 9] == 1)

    pose = P.pose_graph_problem(60, 200, seed=2)
    found, groups = B.argument_slot_ordering(pose)
    assert not found and np.all(groups == -1)

    # several residual-block types: the same rule restated in numpy
    mixed = P.evaluator_cuda_test_problem()
    found, groups = B.argument_slot_ordering(mixed)
    slots, at = {}, 0
    common = min(len(P.COST_TYPES[int(t)][1]) for t in mixed.rb_type)
    for t in mixed.rb_type:
        nb = len(P.COST_TYPES[int(t)][1])
        for j, i in enumerate(mixed.rb_pb[at:at + nb]):
            slots[int(i)] = slots.get(int(i), 0) | (1 << j)
        at += nb
    variable = [i for i in range(mixed.num_pb) if not mixed.pb_constant[i]]
    counts = [sum(1 for i in variable if slots.get(i, 0) == (1 << j)) for j in range(common)]
    best = int(np.argmax(counts))
    assert found == (2 * counts[best] >= len(variable))
    if found:
        expect = np.array([0 if slots.get(i, 0) == (1 << best) else 1 for i in range(mixed.num_pb)])
        assert np.array_equal(groups, expect)
