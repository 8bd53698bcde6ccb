import sys; sys.path[:0]=['.','oracle']
import ceres_b200, numpy as np
from ceres_b200 import lm, problems as P, binding as B
import oracle_py as O
for sub in (False, True):
    spec = P.bal_problem(12, 600, 2600, seed=31)
    rng = np.random.default_rng(31)
    spec.pb_values[:] += rng.normal(0, 0.02, spec.pb_values.size) * (np.abs(spec.pb_values) < 50)
    if sub:
        ncam = spec.meta["num_cameras"]
        spec.pb_manifold_kind[-ncam:] = P.MANIFOLD_SUBSET
        spec.pb_manifold_param[-ncam:] = 0b1000001
    op = O.OracleProblem(spec, jacobian_format=1); cp = B.CudaProblem(spec, jacobian_format=1)
    ref = lm.solve(op, max_num_iterations=15); out = lm.solve(cp, max_num_iterations=15)
    print(sub, len(ref["iterations"]), len(out["iterations"]), ref["cost"], out["cost"], abs(out["cost"]-ref["cost"])/ref["cost"], np.max(np.abs(out["x"]-ref["x"]))/np.max(np.abs(ref["x"])))
    for a,b in zip(ref["iterations"], out["iterations"]): print("   ", a["cost"], b["cost"], abs(a["cost"]-b["cost"])/a["cost"])
