"""Times the device-resident evaluation for each combination of requested outputs."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import ceres_b200
from ceres_b200 import binding as B, problems as P
shape = sys.argv[1] if len(sys.argv) > 1 else "M"
spec = P.bal_shape(shape)
cp = B.CudaProblem(spec)
cp.evaluate()
for (r, g, j) in [(1,1,1),(1,0,1),(1,1,0),(0,0,1),(0,1,0),(1,0,0),(0,0,0)]:
    ts = []
    for _ in range(8):
        cp.evaluate_device(residuals=bool(r), gradient=bool(g), jacobian=bool(j))
        ts.append(cp.timing()["kernel_ms"])
    print(f"r={r} g={g} j={j}: kernel {np.median(ts[2:]):.3f} ms  -> {spec.num_rb/np.median(ts[2:])/1e6:.2f} G RB/s")
