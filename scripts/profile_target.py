"""Small fixed program for ncu: builds a BAL problem and runs a few device-resident
evaluations (all outputs).  Usage: profile_target.py [shape] [scale] [n_eval]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import ceres_b200
from ceres_b200 import binding as B, problems as P
shape = sys.argv[1] if len(sys.argv) > 1 else "M"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cp = B.CudaProblem(P.bal_shape(shape, scale=scale))
cp.evaluate()
for _ in range(n):
    ok, cost = cp.evaluate_device()
print("ok", ok, cost, cp.timing())
