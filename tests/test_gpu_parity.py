"""GPU parity tests: the CUDA path (through the public C++ API and the C ABI) against
the oracle on the same seeded inputs.  Tolerances are north_star's: residual and
Jacobian values 1e-12 relative, cost and gradient 1e-10 relative; structure bit-exact
(tests/test_host_structure.py)."""
import itertools

import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B
from ceres_b200 import problems as P

pytestmark = pytest.mark.gpu

RTOL_VALUES = 1e-12
RTOL_SUMS = 1e-10


def _close(a, b, rtol, what):
    """max|a - b| <= rtol * max|b|: relative error in the infinity norm of the whole
    vector, the norm-wise notion the fork's own GPU-vs-CPU test uses
    (evaluator_cuda_test.cu.cc:425-440, isApprox).  An element-wise bound is not
    meaningful here: a residual is predicted - observed with both ~1e2..1e3 px, so two
    correct implementations differ by ~1e-16 * 1e3 in absolute terms whatever the size
    of the residual itself."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    if b.size == 0:
        return
    scale = max(float(np.max(np.abs(b))), 1e-300)
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= rtol, f"{what}: max error / max magnitude = {err:.3e} > {rtol:.0e}"


def _check(spec, fmt=0, state=None, **kw):
    op = O.OracleProblem(spec, jacobian_format=fmt, **kw)
    cp = B.CudaProblem(spec, jacobian_format=fmt, **kw)
    if state is None:
        state = op.initial_state()
    ok_o, c_o, r_o, g_o, j_o = op.evaluate(state)
    cp.jacobian_values[:] = -1.0  # every value must be overwritten (no memset in the engine)
    ok_c, c_c, r_c, g_c, j_c = cp.evaluate(state)
    assert ok_o and ok_c
    _close(c_c, c_o, RTOL_SUMS, "cost")
    _close(r_c, r_o, RTOL_VALUES, "residuals")
    _close(j_c[:op.num_jacobian_values], j_o[:op.num_jacobian_values], RTOL_VALUES, "jacobian")
    _close(g_c, g_o, RTOL_SUMS, "gradient")
    return op, cp


@pytest.mark.parametrize("fmt", [0, 1])
def test_fork_fixture(fmt):
    """internal/ceres/evaluator_cuda_test.cu.cc:280-459 (the fork's own GPU-vs-CPU test,
    tolerance 1e-13 there)."""
    op, cp = _check(P.evaluator_cuda_test_problem(), fmt)
    ok, c_c, r_c, g_c, j_c = cp.evaluate()
    ok, c_o, r_o, g_o, j_o = op.evaluate()
    assert abs(c_c - c_o) <= 1e-13 * max(1.0, abs(c_o))
    n = op.num_jacobian_values
    assert np.linalg.norm(r_c - r_o) <= 1e-13 * np.linalg.norm(r_o)
    assert np.linalg.norm(g_c - g_o) <= 1e-13 * np.linalg.norm(g_o)
    assert np.linalg.norm(j_c[:n] - j_o[:n]) <= 1e-13 * np.linalg.norm(j_o[:n])


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("loss", ["none", "huber", "cauchy"])
def test_bal_small(fmt, loss):
    _check(P.bal_problem(16, 300, 1200, seed=1, loss=loss), fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_bal_subset_manifold_and_constant_cameras(fmt):
    _check(P.bal_problem(12, 200, 800, seed=2, subset_manifold=True, constant_cameras=3), fmt)


def test_bal_s_shape_all_output_combinations():
    """BAL "S" shape (16 x 22106, 83718 blocks): every combination of requested
    outputs, as evaluator_test.cc:127-220 CheckAllEvaluationCombinations does."""
    spec = P.bal_shape("S")
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec)
    x = op.initial_state()
    ok, c_o, r_o, g_o, j_o = op.evaluate(x, num_threads=8)
    for wr, wg, wj in itertools.product([False, True], repeat=3):
        cp.jacobian_values[:] = -1.0
        ok, c, r, g, j = cp.evaluate(x, residuals=wr, gradient=wg, jacobian=wj)
        assert ok
        _close(c, c_o, RTOL_SUMS, "cost")
        if wr:
            _close(r, r_o, RTOL_VALUES, "residuals")
        if wg:
            _close(g, g_o, RTOL_SUMS, "gradient")
        if wj:
            _close(j[:op.num_jacobian_values], j_o, RTOL_VALUES, "jacobian")


def test_apply_loss_function_false():
    spec = P.bal_problem(8, 100, 400, seed=3)
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec)
    x = op.initial_state()
    _, c_o, r_o, g_o, j_o = op.evaluate(x, apply_loss_function=False)
    _, c_c, r_c, g_c, j_c = cp.evaluate(x, apply_loss_function=False)
    _close(c_c, c_o, RTOL_SUMS, "cost")
    _close(r_c, r_o, RTOL_VALUES, "residuals")
    _close(j_c[:op.num_jacobian_values], j_o, RTOL_VALUES, "jacobian")
    _close(g_c, g_o, RTOL_SUMS, "gradient")


@pytest.mark.parametrize("kind,a,b", [(P.LOSS_TRIVIAL, 0, 0), (P.LOSS_SCALED_HUBER, 1.0, 0.7),
                                      (P.LOSS_SCALED_CAUCHY, 2.0, 1.3),
                                      (P.LOSS_SCALED_TRIVIAL, 0, 0.4)])
def test_other_losses(kind, a, b):
    spec = P.bal_problem(6, 80, 300, seed=4)
    spec.rb_loss_kind[:] = kind
    spec.rb_loss_a[:] = a
    spec.rb_loss_b[:] = b
    _check(spec)


def test_mixed_losses_in_one_type():
    """Several loss objects of one type: exercises the loss table + per-block index."""
    spec = P.bal_problem(6, 80, 300, seed=5)
    spec.rb_loss_a[::3] = 0.5
    spec.rb_loss_a[1::3] = 2.0
    _check(spec)


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("loss", ["none", "cauchy"])
def test_pose_graph(fmt, loss):
    _check(P.pose_graph_problem(300, 900, seed=5, loss=loss), fmt)


def test_schur_reordered_program_uses_positions():
    """SURVEY.md hazard 1: after LexicographicallyOrderResidualBlocks the program
    position differs from the insertion index; values must land in position order."""
    spec = P.bal_problem(9, 120, 500, seed=6)
    perm = np.random.default_rng(1).permutation(spec.num_rb)
    spec = P.ProblemSpec(pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[perm],
                         rb_pb=spec.rb_pb.reshape(-1, 2)[perm].ravel(),
                         fdata=spec.fdata.reshape(-1, 2)[perm].ravel(),
                         rb_loss_kind=spec.rb_loss_kind[perm], rb_loss_a=spec.rb_loss_a[perm],
                         rb_loss_b=spec.rb_loss_b[perm],
                         num_eliminate_blocks=spec.num_eliminate_blocks)
    for fmt in (0, 1):
        _check(spec, fmt, schur_reorder=True)


# ---- the reference's CPU evaluator tables through the CUDA path (evaluator_test.cc:227-560)
def _xyz(b):
    return {n: b.add_parameter_block(np.zeros(s)) for n, s in (("x", 2), ("y", 3), ("z", 4))}


@pytest.mark.parametrize("fmt,nelim", [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0)])
def test_evaluator_test_tables(fmt, nelim):
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.set_manifold(v["y"], P.MANIFOLD_SUBSET, 0b001)
    b.set_manifold(v["z"], P.MANIFOLD_SUBSET, 0b0010)
    b.add_residual_block(P.AFFINE_1_2_23, [v["x"], v["y"]])
    b.add_residual_block(P.AFFINE_2_3_24, [v["x"], v["z"]])
    b.add_residual_block(P.AFFINE_3_4_34, [v["y"], v["z"]])
    cp = B.CudaProblem(b.build(), jacobian_format=fmt, reduce=False, num_eliminate_blocks=nelim)
    jac = np.array([1, 2, 2, 3, 0, 0, 0] * 2 + [2, 4, 0, 0, 2, 6, 8] * 3 +
                   [0, 0, 6, 9, 3, 9, 12] * 4, float).reshape(9, 7)
    for wr, wg, wj in itertools.product([False, True], repeat=3):
        ok, c, r, g, j = cp.evaluate(np.zeros(9), residuals=wr, gradient=wg, jacobian=wj)
        assert ok and c == (1 + 4 + 1 + 4 + 9 + 1 + 4 + 9 + 16) / 2.0
        if wr:
            assert r.tolist() == [1.0, 2.0, 1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 4.0]
        if wg:
            assert g.tolist() == [15.0, 30.0, 66.0, 99.0, 42.0, 126.0, 168.0]
        if wj:
            assert np.array_equal(cp.dense_jacobian(), jac)


def test_constant_parameter_table():
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.add_residual_block(P.AFFINE_1_2_23, [v["x"], v["y"]])
    b.add_residual_block(P.AFFINE_2_3_24, [v["x"], v["z"]])
    b.add_residual_block(P.AFFINE_3_4_34, [v["y"], v["z"]])
    b.set_constant(v["z"])
    cp = B.CudaProblem(b.build(), reduce=True)
    ok, c, r, g, j = cp.evaluate(np.zeros(5))
    assert ok and c == 24.5 and g.tolist() == [15.0, 30.0, 33.0, 66.0, 99.0]
    jac = np.array([1, 2, 1, 2, 3] * 2 + [2, 4, 0, 0, 0] * 3 + [0, 0, 3, 6, 9] * 4, float)
    assert np.array_equal(cp.dense_jacobian(), jac.reshape(9, 5))


def test_failing_functor_returns_false():
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.add_residual_block(P.AFFINE_FAIL, [v["x"], v["y"], v["z"]])
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, *_ = cp.evaluate(np.zeros(9), residuals=False, gradient=False, jacobian=False)
    assert not ok


def test_unwritten_residual_is_rejected():
    """autodiff_cost_function_cuda_test.cu.cc:230-293 + residual_block.cc:110-129: a
    functor that leaves an output unwritten yields kImpossibleValue, which the CPU
    evaluator (and this engine, unlike the fork) rejects."""
    b = P.ProblemBuilder()
    x = b.add_parameter_block([1.0])
    b.add_residual_block(P.ONLY_FILLS_ONE, [x])
    spec = b.build()
    assert not O.OracleProblem(spec, reduce=False).evaluate()[0]
    assert not B.CudaProblem(spec, reduce=False).evaluate()[0]


def test_non_finite_residual_is_rejected():
    spec = P.bal_problem(4, 30, 100, seed=7)
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec)
    x = op.initial_state()
    x[5] = np.inf
    assert not op.evaluate(x)[0]
    assert not cp.evaluate(x)[0]


def test_autodiff_known_answers():
    """autodiff_cost_function_cuda_test.cu.cc:102-116,205-222 through the engine."""
    b = P.ProblemBuilder()
    x = b.add_parameter_block([1.0, 2.0])
    y = b.add_parameter_block([3.0, 4.0])
    b.add_residual_block(P.BINARY_SCALAR, [x, y], [1.0])
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, c, r, g, j = cp.evaluate()
    assert ok and r.tolist() == [10.0] and cp.dense_jacobian().tolist() == [[3.0, 4.0, 1.0, 2.0]]
    b = P.ProblemBuilder()
    xs = [b.add_parameter_block([float(i)]) for i in range(10)]
    b.add_residual_block(P.TEN_PARAMETER, xs)
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, c, r, g, j = cp.evaluate()
    assert ok and r.tolist() == [45.0] and cp.dense_jacobian().tolist() == [[1.0] * 10]


def test_state_changes_are_respected():
    """evaluator_test.cc:598-650."""
    b = P.ProblemBuilder()
    x = b.add_parameter_block([1.0, 1.0])
    b.add_residual_block(P.PARAMETER_SENSITIVE, [x])
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, c, r, g, j = cp.evaluate(np.array([1.0, 1.0]))
    assert c == 1.0 and r.tolist() == [1.0, 1.0] and cp.dense_jacobian().tolist() == [[2, 0], [0, 2]]
    ok, c, r, g, j = cp.evaluate(np.array([2.0, 3.0]))
    assert c == 48.5 and r.tolist() == [4.0, 9.0] and cp.dense_jacobian().tolist() == [[4, 0], [0, 6]]


def test_single_gpu_shards_reassemble():
    """Multi-GPU sharding emulated on one device: world_size ranks evaluated one after
    the other, each writing its own slices; the union must equal the unsharded
    result and cost/gradient must add up (no NCCL involved)."""
    spec = P.bal_problem(10, 400, 1500, seed=8)
    for fmt in (0, 1):
        op = O.OracleProblem(spec, jacobian_format=fmt)
        x = op.initial_state()
        _, c_o, r_o, g_o, j_o = op.evaluate(x)
        W = 3
        r = np.full(op.num_residuals, np.nan)
        jv = np.full(op.values_size, np.nan)
        g = np.zeros(op.num_effective_parameters)
        c = 0.0
        covered = 0
        for k in range(W):
            cp = B.CudaProblem(spec, jacobian_format=fmt, rank=k, world_size=W)
            ok, ck, rk, gk, jk = cp.evaluate(x, out_residuals=r)
            assert ok
            info = cp.shard_info()
            for (gb, ln, lb) in info["segments"]:
                jv[gb:gb + ln] = jk[gb:gb + ln]
                covered += ln
            c += ck
            g += gk
            cp.close()
        assert covered == op.num_jacobian_values
        _close(c, c_o, RTOL_SUMS, "cost")
        _close(r, r_o, RTOL_VALUES, "residuals")
        _close(jv[:op.num_jacobian_values], j_o[:op.num_jacobian_values], RTOL_VALUES, "jacobian")
        _close(g, g_o, RTOL_SUMS, "gradient")


def test_device_resident_evaluation_matches():
    spec = P.bal_problem(8, 200, 700, seed=9)
    cp = B.CudaProblem(spec)
    ok, c, r, g, j = cp.evaluate()
    ok2, c2 = cp.evaluate_device()
    assert ok and ok2 and c2 == c
    t = cp.timing()
    assert t["launches"] >= 2 and t["kernel_ms"] > 0


# ---- the CUDA path against the REFERENCE's own arithmetic (tests/golden/reference_arith.npz,
# produced from /root/reference headers by tests/golden/make_golden.py)
import os as _os
_G = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "reference_arith.npz"))


def _golden_problem(cost_type, cams, pts, obs):
    b = P.ProblemBuilder()
    for cam, pt, ob in zip(cams, pts, obs):
        c = b.add_parameter_block(cam)
        p = b.add_parameter_block(pt)
        b.add_residual_block(cost_type, [c, p], ob)
    return b.build()


@pytest.mark.parametrize("name,cost_type,csize", [("snavely", P.SNAVELY, 9), ("quat", P.SNAVELY_QUAT, 10)])
def test_golden_reprojection_errors(name, cost_type, csize):
    cams = _G["snavely_cam"] if name == "snavely" else _G["quat_cam"]
    res_ref = _G[name + "_res"]
    jc_ref, jp_ref = _G[name + "_jcam"], _G[name + "_jpt"]
    spec = _golden_problem(cost_type, cams, _G["snavely_pt"], _G["snavely_obs"])
    cp = B.CudaProblem(spec, reduce=False)
    ok, c, r, g, j = cp.evaluate()
    assert ok
    n = cams.shape[0]
    r = r.reshape(n, 2)
    # BlockSparseMatrix, num_eliminate_blocks = 0: cells in residual-block order,
    # camera cell then point cell
    cell = j[:n * 2 * (csize + 3)].reshape(n, 2 * (csize + 3))
    jc, jp = cell[:, :2 * csize], cell[:, 2 * csize:]
    for i in range(n):
        scale_r = np.max(np.abs(_G["snavely_obs"][i])) + np.max(np.abs(res_ref[i]))
        assert np.max(np.abs(r[i] - res_ref[i])) <= 1e-12 * scale_r
        assert np.max(np.abs(jc[i] - jc_ref[i])) <= 1e-12 * np.max(np.abs(jc_ref[i]))
        assert np.max(np.abs(jp[i] - jp_ref[i])) <= 1e-12 * np.max(np.abs(jp_ref[i]))


_GP = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "reference_pose.npz"))


def test_golden_relative_pose_error():
    """RelativePoseError<6,7,7> on the device against the reference's own header
    (internal/ceres/autodiff_benchmarks/relative_pose_error.h:46-92 over the Eigen shim)."""
    n = _GP["rp_pose_i"].shape[0]
    b = P.ProblemBuilder()
    for i in range(n):
        pi = b.add_parameter_block(_GP["rp_pose_i"][i])
        pj = b.add_parameter_block(_GP["rp_pose_j"][i])
        b.add_residual_block(P.RELATIVE_POSE, [pi, pj], _GP["rp_meas"][i])
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, c, r, g, j = cp.evaluate()
    assert ok
    r = r.reshape(n, 6)
    cell = j[:n * 84].reshape(n, 84)   # block sparse: cell of pose i, then of pose j (6 x 7 each)
    for i in range(n):
        scale = max(float(np.max(np.abs(_GP["rp_res"][i]))), 1.0)
        assert np.max(np.abs(r[i] - _GP["rp_res"][i])) <= 1e-12 * scale
        for got, want in ((cell[i, :42], _GP["rp_jac_i"][i]), (cell[i, 42:], _GP["rp_jac_j"][i])):
            assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


def test_golden_pose_graph_3d_error_term():
    """PoseGraph3dErrorTerm<6,3,4,3,4> on the device against the reference's own header
    (examples/slam/pose_graph_3d/pose_graph_3d_error_term.h:71-124 over the Eigen shim)."""
    n = _GP["pg_p_a"].shape[0]
    b = P.ProblemBuilder()
    for i in range(n):
        ids = [b.add_parameter_block(_GP[k][i]) for k in ("pg_p_a", "pg_q_a", "pg_p_b", "pg_q_b")]
        b.add_residual_block(P.POSE_GRAPH_3D, ids, _GP["pg_data"][i])
    cp = B.CudaProblem(b.build(), reduce=False)
    ok, c, r, g, j = cp.evaluate()
    assert ok
    r = r.reshape(n, 6)
    cell = j[:n * 84].reshape(n, 84)   # cells 6 x 3, 6 x 4, 6 x 3, 6 x 4 in block order
    for i in range(n):
        assert np.max(np.abs(r[i] - _GP["pg_res"][i])) <= 1e-12 * np.max(np.abs(_GP["pg_res"][i]))
        at = 0
        for name, width in (("pg_j0", 18), ("pg_j1", 24), ("pg_j2", 18), ("pg_j3", 24)):
            want = _GP[name][i]
            assert np.max(np.abs(cell[i, at:at + width] - want)) <= 1e-12 * np.max(np.abs(want))
            at += width


def test_golden_jet_operations_on_device():
    """Every Jet operation evaluated by the device Jet (ceres/jet.h of this repo) against
    the reference's jet.h; tolerance 1e-13 as in internal/ceres/jet_cuda_test.cu.cc:72-98."""
    for xy, want in zip(_G["jet_xy"], _G["jet_battery"]):
        b = P.ProblemBuilder()
        x = b.add_parameter_block(xy)
        b.add_residual_block(P.JET_BATTERY, [x])
        cp = B.CudaProblem(b.build(), reduce=False)
        ok, c, r, g, j = cp.evaluate()
        assert ok
        want = want.reshape(40, 3)
        J = cp.dense_jacobian()
        assert np.allclose(r, want[:, 0], rtol=1e-13, atol=1e-300)
        assert np.allclose(J, want[:, 1:], rtol=1e-13, atol=1e-15)


def test_generic_manifold_path_matches_device_manifold_path():
    """A manifold the kernel does not know (plus-Jacobian from the host pool,
    CB200_MANIFOLD_GENERIC) must give the same Jacobian as the device-side SubsetManifold."""
    spec = P.bal_problem(10, 150, 600, seed=12, subset_manifold=True)
    opaque = P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type, rb_pb=spec.rb_pb,
        fdata=spec.fdata, pb_constant=spec.pb_constant,
        pb_manifold_kind=np.where(spec.pb_manifold_kind == P.MANIFOLD_SUBSET,
                                  P.MANIFOLD_OPAQUE_SUBSET, spec.pb_manifold_kind),
        pb_manifold_param=spec.pb_manifold_param, rb_loss_kind=spec.rb_loss_kind,
        rb_loss_a=spec.rb_loss_a, rb_loss_b=spec.rb_loss_b,
        num_eliminate_blocks=spec.num_eliminate_blocks)
    for fmt in (0, 1):
        _check(opaque, fmt)
        _check(spec, fmt)


def test_subset_manifold_with_several_fixed_coordinates():
    spec = P.bal_problem(10, 150, 600, seed=13, subset_manifold=True)
    spec.pb_manifold_param[spec.pb_manifold_kind == P.MANIFOLD_SUBSET] = 0b101000010
    for fmt in (0, 1):
        _check(spec, fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_pose_graph_3d_error_term(fmt):
    """<6, 3, 4, 3, 4> with EigenQuaternionManifold on the rotations (two derivative
    passes, four parameter blocks, 344-byte functor kept in global memory)."""
    rng = np.random.default_rng(21)
    n = 60
    b = P.ProblemBuilder()
    ps, qs = [], []
    for i in range(n):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        ps.append(b.add_parameter_block(rng.normal(size=3)))
        qs.append(b.add_parameter_block(q, manifold=(P.MANIFOLD_EIGEN_QUATERNION, 0)))
    b.set_constant(ps[0]); b.set_constant(qs[0])
    for e in range(200):
        i, j = rng.choice(n, 2, replace=False)
        mq = rng.normal(size=4); mq /= np.linalg.norm(mq)
        S = np.eye(6) + 0.1 * rng.normal(size=(6, 6))
        d = np.concatenate([rng.normal(size=3), mq, S.ravel()])
        b.add_residual_block(P.POSE_GRAPH_3D, [ps[i], qs[i], ps[j], qs[j]], d)
    _check(b.build(), fmt)


@pytest.mark.parametrize("nobs", [4, 31, 33, 127, 129, 385])
def test_ragged_sizes(nobs):
    """Block counts around the warp (32) and CTA (128) tile sizes: the tail of the last tile
    is masked, not skipped or duplicated."""
    spec = P.bal_problem(40, max(2, nobs // 3), nobs, seed=nobs)
    assert spec.num_rb == nobs
    _check(spec, fmt=0)
    _check(spec, fmt=1)


def test_every_parameter_block_constant():
    """RemoveFixedBlocks leaves an empty program (program.cc:324-430): nothing to launch, the
    cost is the fixed cost and the outputs are empty."""
    spec = P.bal_problem(6, 10, 40, seed=9)
    spec.pb_constant[:] = 1
    op = O.OracleProblem(spec)
    cp = B.CudaProblem(spec)
    assert cp.num_effective_parameters == 0 and cp.num_residuals == 0
    assert cp.num_residuals == op.num_residuals
    ok, cost, r, g, j = cp.evaluate()
    assert ok and cost == 0.0 and r.size == 0 and g.size == 0
    assert abs(cp.fixed_cost - op.fixed_cost) <= RTOL_SUMS * op.fixed_cost


def test_one_residual_block():
    spec = P.bal_problem(2, 2, 4, seed=3)
    keep = slice(0, 1)
    one = P.ProblemSpec(pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[keep],
                        rb_pb=spec.rb_pb[:2], fdata=spec.fdata[:2],
                        rb_loss_kind=spec.rb_loss_kind[keep], rb_loss_a=spec.rb_loss_a[keep],
                        rb_loss_b=spec.rb_loss_b[keep])
    _check(one, fmt=0)
    _check(one, fmt=1)
