"""Pins the oracle (oracle/oracle_eval.cc) against the known-answer tests the
reference holds for the evaluation path.  CPU only."""
import itertools

import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import problems as P


# ---- internal/ceres/autodiff_cost_function_cuda_test.cu.cc:102-116
def test_bilinear_differentiation():
    ok, r, _ = O.cost_evaluate(P.BINARY_SCALAR, [1.0], [[1, 2], [3, 4]], want_jacobians=False)
    assert ok and r[0] == 10.0
    ok, r, J = O.cost_evaluate(P.BINARY_SCALAR, [1.0], [[1, 2], [3, 4]])
    assert ok and r[0] == 10.0
    assert J[0].tolist() == [[3.0, 4.0]] and J[1].tolist() == [[1.0, 2.0]]


# ---- autodiff_cost_function_cuda_test.cu.cc:205-222
def test_many_parameter_autodiff():
    params = [[float(i)] for i in range(10)]
    ok, r, J = O.cost_evaluate(P.TEN_PARAMETER, [], params)
    assert ok and r[0] == 45.0
    assert all(j.tolist() == [[1.0]] for j in J)


# ---- autodiff_cost_function_cuda_test.cu.cc:290-292: Evaluate succeeds but leaves
# kImpossibleValue behind in the unfilled residual / jacobian row.
def test_partially_filled_residual():
    ok, r, J = O.cost_evaluate(P.ONLY_FILLS_ONE, [], [[1.0]])
    assert ok
    assert r[0] == 1.0 and r[1] == 1e302
    assert J[0][0, 0] == 1.0 and J[0][1, 0] == 1e302


# ---- internal/ceres/loss_function_test.cc:46-72 (finite-difference consistency)
@pytest.mark.parametrize("kind,a,b", [
    (P.LOSS_TRIVIAL, 0, 0), (P.LOSS_HUBER, 1.4, 0), (P.LOSS_HUBER, 0.7, 0),
    (P.LOSS_CAUCHY, 0.7, 0), (P.LOSS_CAUCHY, 1.3, 0), (P.LOSS_SCALED_HUBER, 1.0, 0.6),
    (P.LOSS_SCALED_CAUCHY, 0.7, 1.792), (P.LOSS_SCALED_TRIVIAL, 0, 0.3)])
@pytest.mark.parametrize("s", [0.357, 1.792, 0.0 + 1e-3, 0.5, 3.5])
def test_loss_derivatives(kind, a, b, s):
    h = 1e-6
    rho = O.loss_evaluate(kind, a, b, s)
    fwd = O.loss_evaluate(kind, a, b, s + h)
    bwd = O.loss_evaluate(kind, a, b, s - h)
    assert abs((fwd[0] - bwd[0]) / (2 * h) - rho[1]) < 1e-6
    assert abs((fwd[1] - bwd[1]) / (2 * h) - rho[2]) < 1e-6


def test_huber_values():
    # loss_function.cc:52-66 at a = 1: inlier identity, outlier 2 sqrt(s) - 1
    assert O.loss_evaluate(P.LOSS_HUBER, 1.0, 0, 0.25).tolist() == [0.25, 1.0, 0.0]
    rho = O.loss_evaluate(P.LOSS_HUBER, 1.0, 0, 4.0)
    assert rho.tolist() == [3.0, 0.5, -0.0625]


# ---- internal/ceres/corrector_test.cc:56-120 scalar Gauss-Newton identities
def _scalar_corrector_check(sq_norm, rho, x, jac):
    # robustified gradient / GN hessian must match corrected r, J
    r, j = O.corrector(sq_norm, rho, [x], [jac])
    g_res = -2.0 * rho[1] * x * jac                      # derivative of rho(x^2)/... up to sign
    return r[0], j[0]


def test_corrector_scalar_cases():
    # corrector_test.cc:56-82 ScalarCorrection: rho = {., 0.5, 0.25}? values from the test:
    x, jac = np.sqrt(3.0), 10.0
    sq_norm = x * x
    for rho in ([sq_norm, 0.1, -0.01], [sq_norm, 0.5, 0.0], [sq_norm, 0.3, 0.04]):
        r, j = O.corrector(sq_norm, rho, [x], [jac])
        # gradient identity: J^T r == rho' * jac * x  (corrector_test.cc:84-120)
        assert abs(j[0] * r[0] - rho[1] * jac * x) < 1e-10
        # hessian identity when the curvature correction applies (rho'' > 0)
        if rho[2] > 0:
            h = jac * jac * (rho[1] + 2 * rho[2] * sq_norm)
            assert abs(j[0] * j[0] - h) < 1e-9
        else:
            assert abs(j[0] * j[0] - rho[1] * jac * jac) < 1e-10


# ---- corrector_test.cc:149-271 multidimensional identities
def test_corrector_multidimensional():
    rng = np.random.default_rng(7)
    for _ in range(10):
        J = rng.normal(size=(3, 2))
        r = rng.normal(size=3)
        sq = float(r @ r)
        for rho in ([sq, rng.random(), rng.random()], [sq, rng.random(), -rng.random()]):
            rc, Jc = O.corrector(sq, rho, r, J)
            Jc = Jc.reshape(3, 2)
            g_expect = rho[1] * J.T @ r
            assert np.allclose(Jc.T @ rc, g_expect, atol=1e-10)
            if rho[2] > 0:
                H = J.T @ (rho[1] * np.eye(3) + 2 * rho[2] * np.outer(r, r)) @ J
                assert np.allclose(Jc.T @ Jc, H, atol=1e-10)
            else:
                assert np.allclose(Jc.T @ Jc, rho[1] * J.T @ J, atol=1e-10)


def test_corrector_zero_norm():
    # corrector.h:95-99: s == 0 -> plain sqrt(rho') scaling, no division by zero
    r, j = O.corrector(0.0, [0.0, 0.25, 0.1], [0.0, 0.0], [1.0, 2.0, 3.0, 4.0])
    assert r.tolist() == [0.0, 0.0] and j.tolist() == [0.5, 1.0, 1.5, 2.0]


# ---- internal/ceres/rotation_test.cc:1707-1810 AngleAxisRotatePoint vs rotation matrix
def _rot_matrix(aa):
    th = np.linalg.norm(aa)
    if th == 0:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_angle_axis_rotate_point():
    rng = np.random.default_rng(3)
    for _ in range(200):
        aa = rng.normal(size=3)
        aa *= rng.uniform(0, np.pi) / np.linalg.norm(aa)
        pt = rng.uniform(-1, 1, 3)
        assert np.allclose(O.angle_axis_rotate_point(aa, pt), _rot_matrix(aa) @ pt, atol=1e-10)
    # near-zero rotations (rotation_test.cc NearZeroAngleAxisRotatePoint...)
    for _ in range(100):
        aa = rng.normal(size=3)
        aa *= 1e-10 / np.linalg.norm(aa)
        pt = rng.uniform(-1, 1, 3)
        assert np.allclose(O.angle_axis_rotate_point(aa, pt), _rot_matrix(aa) @ pt, atol=1e-10)
    assert np.array_equal(O.angle_axis_rotate_point(np.zeros(3), [1.0, 2, 3]), [1.0, 2, 3])


# ---- internal/ceres/manifold_test.cc:129-160 (Subset), :500-680 (quaternions)
def test_subset_manifold_plus_jacobian():
    J = O.manifold_plus_jacobian(P.MANIFOLD_SUBSET, 0b0010, np.zeros(4))
    assert J.tolist() == [[1, 0, 0], [0, 0, 0], [0, 1, 0], [0, 0, 1]]
    x = np.array([1.0, 2, 3, 4])
    assert O.manifold_plus(P.MANIFOLD_SUBSET, 0b0010, x, [0.5, 0.25, 0.125]).tolist() == \
        [1.5, 2.0, 3.25, 4.125]


@pytest.mark.parametrize("kind,w", [(P.MANIFOLD_QUATERNION, 0), (P.MANIFOLD_EIGEN_QUATERNION, 3)])
def test_quaternion_manifold_plus_jacobian_is_derivative_of_plus(kind, w):
    rng = np.random.default_rng(11)
    for _ in range(20):
        x = rng.normal(size=4); x /= np.linalg.norm(x)
        J = O.manifold_plus_jacobian(kind, 0, x)
        num = np.zeros((4, 3)); h = 1e-6
        for c in range(3):
            d = np.zeros(3); d[c] = h
            num[:, c] = (O.manifold_plus(kind, 0, x, d) - O.manifold_plus(kind, 0, x, -d)) / (2 * h)
        assert np.allclose(J, num, atol=1e-8)
        # Plus keeps unit norm; zero delta is the identity (manifold.cc:28-37)
        assert abs(np.linalg.norm(O.manifold_plus(kind, 0, x, [0.3, -0.2, 0.1])) - 1) < 1e-12
        assert np.array_equal(O.manifold_plus(kind, 0, x, np.zeros(3)), x)


def test_product_manifold_block_diagonal():
    x = np.array([0.5, 0.5, 0.5, 0.5, 1, 2, 3, 4, 5, 6.0])
    J = O.manifold_plus_jacobian(P.MANIFOLD_QUATERNION_X_EUCLIDEAN, 0, x)
    assert J.shape == (10, 9)
    assert np.array_equal(J[:4, :3], O.manifold_plus_jacobian(P.MANIFOLD_QUATERNION, 0, x[:4]))
    assert np.array_equal(J[4:, 3:], np.eye(6)) and not J[:4, 3:].any() and not J[4:, :3].any()


# ---- internal/ceres/evaluator_test.cc:227-560 expected tables (exact equality)
def _xyz(b, order=("x", "y", "z")):
    sizes = {"x": 2, "y": 3, "z": 4}
    return {n: b.add_parameter_block(np.zeros(sizes[n])) for n in order}


def _check_all_combinations(spec, expected, fmt, nelim, reduce=False):
    rows, cols, cost, res, grad, jac = expected
    op = O.OracleProblem(spec, jacobian_format=fmt, reduce=reduce, num_eliminate_blocks=nelim)
    assert op.num_residuals == rows and op.num_effective_parameters == cols
    for wr, wg, wj in itertools.product([False, True], repeat=3):
        ok, c, r, g, j = op.evaluate(np.zeros(op.num_parameters), residuals=wr, gradient=wg,
                                     jacobian=wj)
        assert ok and c == cost
        if wr:
            assert r.tolist() == res
        if wg:
            assert g.tolist() == grad
        if wj:
            dense = op.dense_jacobian(j)
            assert np.array_equal(dense, np.array(jac, float).reshape(rows, cols))


_FORMATS = [(0, n) for n in range(5)] + [(1, 0)]


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_single_residual_problem(fmt, nelim):
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.add_residual_block(P.AFFINE_1_3_234, [v["x"], v["y"], v["z"]])
    expected = (3, 9, 7.0, [1.0, 2.0, 3.0], [6.0, 12.0, 6.0, 12.0, 18.0, 6.0, 12.0, 18.0, 24.0],
                [1, 2, 1, 2, 3, 1, 2, 3, 4] * 3)
    _check_all_combinations(b.build(), expected, fmt, nelim)


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_single_residual_problem_permuted(fmt, nelim):
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.add_residual_block(P.AFFINE_1_3_432, [v["z"], v["y"], v["x"]])
    expected = (3, 9, 7.0, [1.0, 2.0, 3.0], [6.0, 12.0, 6.0, 12.0, 18.0, 6.0, 12.0, 18.0, 24.0],
                [1, 2, 1, 2, 3, 1, 2, 3, 4] * 3)
    _check_all_combinations(b.build(), expected, fmt, nelim)


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_single_residual_problem_nuisance(fmt, nelim):
    b = P.ProblemBuilder()
    b.add_parameter_block(np.zeros(2))          # a
    x = b.add_parameter_block(np.zeros(2))
    b.add_parameter_block(np.zeros(1))          # b
    y = b.add_parameter_block(np.zeros(3))
    b.add_parameter_block(np.zeros(1))          # c
    z = b.add_parameter_block(np.zeros(4))
    b.add_parameter_block(np.zeros(3))          # d
    b.add_residual_block(P.AFFINE_1_3_234, [x, y, z])
    row = [0, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 0, 0]
    expected = (3, 16, 7.0, [1.0, 2.0, 3.0],
                [0.0, 0.0, 6.0, 12.0, 0.0, 6.0, 12.0, 18.0, 0.0, 6.0, 12.0, 18.0, 24.0, 0.0, 0.0, 0.0],
                row * 3)
    _check_all_combinations(b.build(), expected, fmt, min(nelim, 4))


def _multi(b, v):
    b.add_residual_block(P.AFFINE_1_2_23, [v["x"], v["y"]])
    b.add_residual_block(P.AFFINE_2_3_24, [v["x"], v["z"]])
    b.add_residual_block(P.AFFINE_3_4_34, [v["y"], v["z"]])


_MULTI_RES = [1.0, 2.0, 1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 4.0]
_MULTI_COST = (1 + 4 + 1 + 4 + 9 + 1 + 4 + 9 + 16) / 2.0


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_multiple_residual_problem(fmt, nelim):
    b = P.ProblemBuilder()
    _multi(b, _xyz(b))
    jac = ([1, 2, 1, 2, 3, 0, 0, 0, 0] * 2 + [2, 4, 0, 0, 0, 2, 4, 6, 8] * 3 +
           [0, 0, 3, 6, 9, 3, 6, 9, 12] * 4)
    expected = (9, 9, _MULTI_COST, _MULTI_RES,
                [15.0, 30.0, 33.0, 66.0, 99.0, 42.0, 84.0, 126.0, 168.0], jac)
    _check_all_combinations(b.build(), expected, fmt, min(nelim, 3))


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_multiple_residuals_with_manifolds(fmt, nelim):
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.set_manifold(v["y"], P.MANIFOLD_SUBSET, 0b001)   # fix y's first dimension
    b.set_manifold(v["z"], P.MANIFOLD_SUBSET, 0b0010)  # fix z's second dimension
    _multi(b, v)
    jac = ([1, 2, 2, 3, 0, 0, 0] * 2 + [2, 4, 0, 0, 2, 6, 8] * 3 + [0, 0, 6, 9, 3, 9, 12] * 4)
    expected = (9, 7, _MULTI_COST, _MULTI_RES, [15.0, 30.0, 66.0, 99.0, 42.0, 126.0, 168.0], jac)
    _check_all_combinations(b.build(), expected, fmt, min(nelim, 3))


@pytest.mark.parametrize("fmt,nelim", _FORMATS)
def test_multiple_residual_problem_with_constant_parameters(fmt, nelim):
    b = P.ProblemBuilder()
    v = _xyz(b)
    _multi(b, v)
    b.set_constant(v["z"])
    jac = ([1, 2, 1, 2, 3] * 2 + [2, 4, 0, 0, 0] * 3 + [0, 0, 3, 6, 9] * 4)
    expected = (9, 5, _MULTI_COST, _MULTI_RES, [15.0, 30.0, 33.0, 66.0, 99.0], jac)
    _check_all_combinations(b.build(), expected, fmt, min(nelim, 2), reduce=True)


def test_evaluator_aborts_for_failing_residuals():
    b = P.ProblemBuilder()
    v = _xyz(b)
    b.add_residual_block(P.AFFINE_FAIL, [v["x"], v["y"], v["z"]])
    op = O.OracleProblem(b.build(), reduce=False)
    ok, *_ = op.evaluate(np.zeros(9), residuals=False, gradient=False, jacobian=False)
    assert not ok


# ---- evaluator_test.cc:598-650 EvaluatorRespectsParameterChanges
def test_evaluator_respects_parameter_changes():
    b = P.ProblemBuilder()
    x = b.add_parameter_block([1.0, 1.0])
    b.add_residual_block(P.PARAMETER_SENSITIVE, [x])
    op = O.OracleProblem(b.build(), reduce=False)
    ok, c, r, g, j = op.evaluate(np.array([1.0, 1.0]))
    assert ok and c == 1.0 and r.tolist() == [1.0, 1.0]
    assert op.dense_jacobian(j).tolist() == [[2.0, 0.0], [0.0, 2.0]]
    ok, c, r, g, j = op.evaluate(np.array([2.0, 3.0]))
    assert c == 48.5 and r.tolist() == [4.0, 9.0]
    assert op.dense_jacobian(j).tolist() == [[4.0, 0.0], [0.0, 6.0]]


# ---- structure of the fork's fixture (evaluator_cuda_test.cu.cc:280-316)
@pytest.mark.parametrize("fmt", [0, 1])
def test_fork_fixture_structure(fmt):
    spec = P.evaluator_cuda_test_problem()
    op = O.OracleProblem(spec, jacobian_format=fmt, reduce=True)
    # camera2, point2 constant; RB6 (all-constant) folded into fixed_cost
    assert op.num_residual_blocks == 5 and op.num_residuals == 11
    assert op.num_parameter_blocks == 3 and op.num_parameters == 10 + 3 + 7
    assert op.num_effective_parameters == 9 + 3 + 7
    assert op.fixed_cost == 0.0
    ok, c, r, g, j = op.evaluate()
    assert ok and np.isfinite(c) and np.isfinite(r).all() and np.isfinite(g).all()
    # gradient == J^T r on the dense matrix (program_evaluator.h:241-256)
    J = op.dense_jacobian(j)
    assert np.allclose(J.T @ r, g, rtol=1e-12, atol=1e-9)


def test_schur_reorder_groups_by_e_block():
    spec = P.bal_problem(5, 40, 120, seed=9)
    op = O.OracleProblem(spec, schur_reorder=True)
    rbs = op.ints("program_rbs")
    pts = spec.rb_pb.reshape(-1, 2)[rbs, 1]
    assert np.all(np.diff(pts) >= 0)
    # buckets are filled back to front (reorder_program.cc:296-312)
    first = np.flatnonzero(pts == pts[0])
    assert np.all(np.diff(rbs[first]) < 0)


def test_convex_test_loss_and_the_corrector_alpha_branch():
    """The test-only loss rho(s) = s + a s^2 (rho'' = 2a > 0) is the way into the Corrector's
    alpha branch (corrector.cc:105-130); the corrected residual and Jacobian must satisfy the
    Gauss-Newton identities of Triggs et al. that corrector_test.cc:56-271 checks:
    J'J (corrected) = rho' J'J + 2 rho'' J'r r'J and J'r (corrected) = rho' J'r."""
    a = 0.07
    rng = np.random.default_rng(5)
    for _ in range(20):
        r = rng.normal(0, 1.5, 3)
        J = rng.normal(0, 2.0, (3, 4))
        s = float(r @ r)
        rho = O.loss_evaluate(P.LOSS_CONVEX_TEST, a, 0.0, s)
        assert np.allclose(rho, [s + a * s * s, 1 + 2 * a * s, 2 * a], rtol=1e-15)
        rc, Jc = O.corrector(s, rho, r, J.ravel())
        Jc = Jc.reshape(3, 4)
        g_want = rho[1] * J.T @ r
        H_want = rho[1] * J.T @ J + 2 * rho[2] * np.outer(J.T @ r, J.T @ r)
        assert np.allclose(Jc.T @ rc, g_want, rtol=1e-12, atol=1e-12)
        assert np.allclose(Jc.T @ Jc, H_want, rtol=1e-12, atol=1e-12)


def test_sqrt_of_a_constant_zero_is_rejected_like_the_reference():
    """jet.h:617-621: sqrt(Jet(0)) has derivative lanes (1 / (2 sqrt(0))) * 0 = NaN, so the
    reference (and the oracle that restates it) rejects r = x + sqrt(T(0))."""
    ok, r, J = O.cost_evaluate(P.SQRT_OF_CONSTANT, [4.0], [[1.5]])
    assert ok and r[0] == 3.5 and J[0].ravel()[0] == 1.0
    ok, r, J = O.cost_evaluate(P.SQRT_OF_CONSTANT, [0.0], [[1.5]])
    assert not np.isfinite(J[0]).all()
