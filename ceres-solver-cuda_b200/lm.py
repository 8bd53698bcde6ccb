"""Minimal Levenberg-Marquardt loop around an Evaluator, to show the evaluation engine
driving a full solve (north_star: "a full Levenberg-Marquardt solve converging to the
same final cost").

Only the trust-region bookkeeping is restated here
(internal/ceres/levenberg_marquardt_strategy.cc:68-165: diagonal = clamp(diag(J'J)),
D = sqrt(diagonal / radius), radius update on accept / reject;
internal/ceres/trust_region_minimizer.cc:259-275 Jacobi scaling, :780-830 step
evaluation with a cost-only Evaluate, min_relative_decrease = 1e-3, function_tolerance
= 1e-6).  The linear solve is out of scope for this repository (it stays on the
reference's solvers); a CPU sparse Cholesky-free solve from scipy stands in for it.

The `problem` argument is anything with the Evaluator-shaped methods of
binding.CudaProblem / oracle_py.OracleProblem built with the CompressedRow Jacobian:
evaluate(), plus(), ints("crs_rows"/"crs_cols"), num_* attributes.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _jacobian_matrix(problem, values):
    rows = problem.ints("crs_rows")
    cols = problem.ints("crs_cols")
    nnz = problem.num_jacobian_values
    return sp.csr_matrix((values[:nnz].copy(), cols[:nnz], rows),
                         shape=(problem.num_residuals, problem.num_effective_parameters))


def solve(problem, x0=None, max_num_iterations=25, function_tolerance=1e-6,
          gradient_tolerance=1e-10, initial_radius=1e4, verbose=False):
    """Returns dict(x, cost, initial_cost, iterations=[...], evaluations)."""
    assert problem.jacobian_format == 1, "build the problem with the CompressedRow Jacobian"
    x = problem.initial_state() if x0 is None else np.array(x0, dtype=float)
    radius, decrease_factor = float(initial_radius), 2.0
    min_diagonal, max_diagonal, max_radius = 1e-6, 1e32, 1e16
    evaluations = {"jacobian": 0, "cost_only": 0}

    ok, cost, r, g, jv = problem.evaluate(x)
    evaluations["jacobian"] += 1
    if not ok:
        raise RuntimeError("initial evaluation failed")
    J = _jacobian_matrix(problem, jv)
    # Jacobi scaling, computed once (trust_region_minimizer.cc:259-275)
    scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).ravel()))
    history = [dict(iteration=0, cost=cost, radius=radius, accepted=True)]
    initial_cost = cost
    reuse_diagonal = False
    diagonal = None
    for it in range(1, max_num_iterations + 1):
        Js = J @ sp.diags(scale)
        if not reuse_diagonal:
            diagonal = np.clip(np.asarray(Js.multiply(Js).sum(axis=0)).ravel(),
                               min_diagonal, max_diagonal)
        D2 = diagonal / radius
        # (Js' Js + D^2) y = Js' r ; step = -y
        A = (Js.T @ Js + sp.diags(D2)).tocsc()
        y = spla.spsolve(A, Js.T @ r)
        step_s = -y
        model_residuals = Js @ step_s
        model_cost_change = -float(model_residuals @ (r + 0.5 * model_residuals))
        delta = step_s * scale
        x_plus = problem.plus(x, delta)
        ok, new_cost, *_ = problem.evaluate(x_plus, residuals=False, gradient=False, jacobian=False)
        evaluations["cost_only"] += 1
        relative_decrease = (cost - new_cost) / model_cost_change if (ok and model_cost_change > 0) else -1.0
        accepted = ok and relative_decrease > 1e-3
        if verbose:
            print(f"{it:3d} cost {cost:.10e} -> {new_cost:.10e} rho {relative_decrease:+.3f} "
                  f"radius {radius:.2e} {'ok' if accepted else 'rejected'}")
        if accepted:
            cost_change = cost - new_cost
            x = x_plus
            ok, cost, r, g, jv = problem.evaluate(x)
            evaluations["jacobian"] += 1
            J = _jacobian_matrix(problem, jv)
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * relative_decrease - 1.0) ** 3))
            decrease_factor, reuse_diagonal = 2.0, False
            history.append(dict(iteration=it, cost=cost, radius=radius, accepted=True))
            if abs(cost_change) <= function_tolerance * cost:
                break
            if np.max(np.abs(g)) <= gradient_tolerance:
                break
        else:
            radius /= decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = True
            history.append(dict(iteration=it, cost=cost, radius=radius, accepted=False))
            if radius < 1e-32:
                break
    return dict(x=x, cost=cost, initial_cost=initial_cost, iterations=history,
                evaluations=evaluations)
