"""Device-time measurements of the secondary BASELINE.json configs (4: BAL +
SubsetManifold(9,{0}) + CompressedRowSparseMatrix; 5: pose graph <6,7,7> with
EigenQuaternion x R^3 manifold), with a parity check against the oracle on a prefix.
Usage: bench_configs.py [scale]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import ceres_b200
from ceres_b200 import binding as B, problems as P
import oracle_py as O

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def run(name, spec, fmt):
    t0 = time.time()
    cp = B.CudaProblem(spec, jacobian_format=fmt)
    setup = time.time() - t0
    ok, c, r, g, j = cp.evaluate()
    assert ok
    ts, td = [], []
    for _ in range(8):
        cp.evaluate_device()
        t = cp.timing(); ts.append(t["kernel_ms"]); td.append(t["device_ms"])
    k, d = float(np.median(ts[2:])), float(np.median(td[2:]))
    print(f"CONFIG {name}: blocks {cp.num_residual_blocks} nnz {cp.num_jacobian_values} "
          f"kernel {k:.3f} ms device {d:.3f} ms -> {cp.num_residual_blocks / k / 1e6:.2f} G blocks/s "
          f"({cp.num_jacobian_values * 8 / k / 1e6:.0f} GB/s of Jacobian) setup {setup:.1f}s cost {c:.6e}",
          flush=True)
    return cp, (c, r, g, j)


def check(spec, fmt, cp, got, n_check=400_000):
    # parity on the whole problem if small, else cost only against a threaded oracle run
    if spec.num_rb <= n_check:
        op = O.OracleProblem(spec, jacobian_format=fmt)
        ok, c, r, g, j = op.evaluate(num_threads=os.cpu_count())
        def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        print(f"   parity: cost {abs(got[0]-c)/abs(c):.2e} residuals {rel(got[1], r):.2e} "
              f"jacobian {rel(got[3][:op.num_jacobian_values], j[:op.num_jacobian_values]):.2e} "
              f"gradient {rel(got[2], g):.2e}", flush=True)


# config 4: L + SubsetManifold(9, {0}) on every camera + CRS
nc, npts, nobs, seed = P.BAL_SHAPES["L"]
spec4 = P.bal_problem(max(2, int(nc * scale)), max(4, int(npts * scale)), int(nobs * scale),
                      seed=seed, subset_manifold=True)
cp, got = run("4 (BAL L x %.3g, SubsetManifold(9,{0}), CompressedRow)" % scale, spec4, 1)
check(spec4, 1, cp, got)
cp.close()
# same with BlockSparse for comparison
cp, got = run("4b (same, BlockSparse)", spec4, 0)
cp.close(); del spec4

# config 5: pose graph, ~10 M edges at scale 1
n_pose, n_edge = max(10, int(2_500_000 * scale)), max(20, int(10_000_000 * scale))
spec5 = P.pose_graph_problem(n_pose, n_edge, seed=5)
cp, got = run("5 (pose graph %d poses, RelativePoseError<6,7,7>, EigenQuaternion x R3)" % n_pose, spec5, 0)
check(spec5, 0, cp, got)
cp.close()
cp, got = run("5b (same, CompressedRow)", spec5, 1)
cp.close()
