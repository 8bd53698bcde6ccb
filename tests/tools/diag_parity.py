"""Diagnostic: error statistics of the CUDA path vs the oracle on one problem."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import ceres_b200
from ceres_b200 import binding as B, problems as P
import oracle_py as O

spec = P.bal_problem(16, 2000, 8000, seed=1)
op = O.OracleProblem(spec); cp = B.CudaProblem(spec)
x = op.initial_state()
ok, c, r, g, j = cp.evaluate(x)
_, c1, r1, g1, j1 = op.evaluate(x, num_threads=1)
_, c8, r8, g8, j8 = op.evaluate(x, num_threads=8)
def rel(a, b): return np.max(np.abs(a-b))/np.max(np.abs(b))
print("cost gpu vs cpu1", abs(c-c1)/abs(c1), "cpu8 vs cpu1", abs(c8-c1)/abs(c1))
print("res  inf-rel", rel(r, r1), "max|r|", np.max(np.abs(r1)))
print("jac  inf-rel", rel(j, j1), "max|j|", np.max(np.abs(j1)))
print("grad inf-rel gpu", rel(g, g1), "cpu8 vs cpu1", rel(g8, g1), "max|g|", np.max(np.abs(g1)))
k = np.argmax(np.abs(g-g1)); print("worst grad idx", k, g[k], g1[k], "n_eff", g.size)
# gradient terms scale: sum |J|^T |r|
J = op.dense_jacobian(j1) if op.num_residuals*op.num_effective_parameters < 4e8 else None
if J is not None:
    absg = np.abs(J).T @ np.abs(r1)
    print("max |g-g1| / sum|terms| :", np.max(np.abs(g-g1)/np.maximum(absg,1e-300)))
    print("cpu8: max |g8-g1| / sum|terms| :", np.max(np.abs(g8-g1)/np.maximum(absg,1e-300)))
    print("at worst idx: sum|terms|", absg[k])
print(cp.timing())
