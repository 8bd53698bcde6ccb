"""Generates tests/golden/reference_arith.npz from the REFERENCE's own arithmetic
(oracle/_ref/libref_arith.so = /root/reference headers compiled over oracle/eigen_shim,
see oracle/ref_arith.cc and oracle/Makefile).  Run in the build container, where
/root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The .npz travels with the repo; tests compare the oracle (CPU) and the CUDA path (GPU)
against it without needing /root/reference."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_arith.so"))


def p(a):
    return a.ctypes.data_as(C.c_void_p)


rng = np.random.default_rng(20261018)
out = {}

# ---- SnavelyReprojectionError<2,9,3> and ...WithQuaternions<2,10,3>
n = 256
cams = np.empty((n, 9))
cams[:, 0:3] = rng.normal(0, 0.3, (n, 3))
cams[:4, 0:3] = 0.0                       # exact zero rotation: the Taylor branch
cams[4:8, 0:3] *= 1e-9                    # tiny rotations
cams[:, 3:5] = rng.normal(0, 0.5, (n, 2))
cams[:, 5] = -8 + rng.normal(0, 0.5, n)
cams[:, 6] = rng.uniform(400, 1200, n)
cams[:, 7] = rng.normal(0, 1e-7, n)
cams[:, 8] = rng.normal(0, 1e-13, n)
pts = rng.normal(0, 1.0, (n, 3))
obs = rng.normal(0, 300.0, (n, 2))
res = np.zeros((n, 2)); jc = np.zeros((n, 18)); jp = np.zeros((n, 6))
for i in range(n):
    assert L.ref_snavely(p(cams[i]), p(pts[i]), p(obs[i]), p(res[i]), p(jc[i]), p(jp[i])) == 1
out.update(snavely_cam=cams, snavely_pt=pts, snavely_obs=obs, snavely_res=res,
           snavely_jcam=jc, snavely_jpt=jp)

qc = np.empty((n, 10))
q = rng.normal(0, 1, (n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
q[: n // 2] *= rng.uniform(0.5, 2.0, (n // 2, 1))   # non-unit quaternions too
qc[:, 0:4] = q
qc[:, 4:10] = cams[:, 3:9]
res = np.zeros((n, 2)); jc = np.zeros((n, 20)); jp = np.zeros((n, 6))
for i in range(n):
    assert L.ref_snavely_quaternions(p(qc[i]), p(pts[i]), p(obs[i]), p(res[i]), p(jc[i]),
                                     p(jp[i])) == 1
out.update(quat_cam=qc, quat_res=res, quat_jcam=jc, quat_jpt=jp)

# ---- loss functions
L.ref_loss.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
cases = []
for kind, a, b in [(1, 0, 0), (2, 1.0, 0), (2, 0.3, 0), (3, 1.0, 0), (3, 2.5, 0), (4, 1.0, 0.7),
                   (5, 0.8, 1.9), (6, 0, 0.4)]:
    for s in [0.0, 1e-12, 0.05, 0.25, 0.99, 1.0, 1.01, 3.7, 250.0, 1e8]:
        rho = np.zeros(3)
        L.ref_loss(kind, a, b, s, p(rho))
        cases.append([kind, a, b, s, *rho])
out["loss_cases"] = np.array(cases)

# ---- corrector
L.ref_corrector.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
cr_in, cr_out = [], []
for _ in range(64):
    r = rng.normal(0, 2, 2); J = rng.normal(0, 50, (2, 9)); s = float(r @ r)
    rho = np.array([s, rng.uniform(0.05, 1.0), rng.normal(0, 0.2)])
    if _ % 8 == 0:
        rho[2] = 0.0
    rr, JJ = r.copy(), J.copy()
    L.ref_corrector(s, p(rho), 2, 9, p(rr), p(JJ))
    cr_in.append(np.concatenate([[s], rho, r, J.ravel()]))
    cr_out.append(np.concatenate([rr, JJ.ravel()]))
out["corrector_in"] = np.array(cr_in); out["corrector_out"] = np.array(cr_out)

# ---- rotations on doubles and QuaternionToAngleAxis on Jets
aa = rng.normal(0, 1, (128, 3)); aa[:4] = 0; aa[4:8] *= 1e-10
pt = rng.normal(0, 1, (128, 3)); rot = np.zeros((128, 3))
for i in range(128):
    L.ref_angle_axis_rotate_point(p(aa[i]), p(pt[i]), p(rot[i]))
out.update(aa=aa, aa_pt=pt, aa_rot=rot)
qq = rng.normal(0, 1, (128, 4)); qq[:, 0] = np.abs(qq[:, 0]) * np.where(np.arange(128) % 2, 1, -1)
qq[:4, 1:] = 0.0
qv = np.zeros((128, 3)); qj = np.zeros((128, 12))
for i in range(128):
    L.ref_quaternion_to_angle_axis_jet(p(qq[i]), p(qv[i]), p(qj[i]))
out.update(q2aa_q=qq, q2aa_value=qv, q2aa_jac=qj)

# ---- Jet operation battery (jet_cuda_test.cu.cc:108-110 uses x = 2.3, y = 1.7)
xy = np.array([[2.3, 1.7], [0.4, 2.9], [1.1, 1.1]])
bat = []
for row in xy:
    buf = np.zeros(3 * 64)
    k = L.ref_jet_battery(p(row), p(buf))
    bat.append(buf[:3 * k])
out["jet_xy"] = xy; out["jet_battery"] = np.array(bat)

np.savez_compressed(os.path.join(HERE, "reference_arith.npz"), **out)
print("wrote", os.path.join(HERE, "reference_arith.npz"),
      {k: v.shape for k, v in out.items()})

# ---- pose-graph functors (second file, so reference_arith.npz stays byte-identical):
# RelativePoseError<6,7,7> (internal/ceres/autodiff_benchmarks/relative_pose_error.h:46-92) and
# PoseGraph3dErrorTerm<6,3,4,3,4> (examples/slam/pose_graph_3d/pose_graph_3d_error_term.h:71-124)
# compiled unmodified over oracle/eigen_shim/Eigen/{Core,Geometry}.
pose = {}
n = 192


def unit_quats(k, spread=1.0):
    q = rng.normal(0, 1, (k, 4))
    q[:, 3] = np.abs(q[:, 3]) + spread          # (x, y, z, w), w > 0
    return q / np.linalg.norm(q, axis=1, keepdims=True)


pi = np.concatenate([unit_quats(n), rng.normal(0, 3, (n, 3))], axis=1)
pj = np.concatenate([unit_quats(n), rng.normal(0, 3, (n, 3))], axis=1)
pj[:8] = pi[:8]                                  # identical poses: relative rotation = identity
pi[8:16, :4] *= rng.uniform(0.9, 1.1, (8, 1))    # slightly non-unit state quaternions
meas = np.concatenate([unit_quats(n, 3.0), rng.normal(0, 1, (n, 3))], axis=1)
meas[:8, :4] = [0, 0, 0, 1]                      # ... and identity measurement: theta = 0 branch
res = np.zeros((n, 6)); ji = np.zeros((n, 42)); jj = np.zeros((n, 42))
for i in range(n):
    assert L.ref_relative_pose(p(pi[i]), p(pj[i]), p(meas[i]), p(res[i]), p(ji[i]), p(jj[i])) == 1
pose.update(rp_pose_i=pi, rp_pose_j=pj, rp_meas=meas, rp_res=res, rp_jac_i=ji, rp_jac_j=jj)

pa, pb = rng.normal(0, 3, (n, 3)), rng.normal(0, 3, (n, 3))
qa, qb = unit_quats(n), unit_quats(n)
data = np.zeros((n, 43))
data[:, 0:3] = rng.normal(0, 1, (n, 3))
data[:, 3:7] = unit_quats(n, 2.0)
for i in range(n):
    a = rng.normal(0, 1, (6, 6))
    data[i, 7:] = np.linalg.cholesky(a @ a.T + 6 * np.eye(6)).T.ravel()   # a sqrt-information
res = np.zeros((n, 6))
j0 = np.zeros((n, 18)); j1 = np.zeros((n, 24)); j2 = np.zeros((n, 18)); j3 = np.zeros((n, 24))
for i in range(n):
    assert L.ref_pose_graph_3d(p(pa[i]), p(qa[i]), p(pb[i]), p(qb[i]), p(data[i]), p(res[i]),
                               p(j0[i]), p(j1[i]), p(j2[i]), p(j3[i])) == 1
pose.update(pg_p_a=pa, pg_q_a=qa, pg_p_b=pb, pg_q_b=qb, pg_data=data, pg_res=res,
            pg_j0=j0, pg_j1=j1, pg_j2=j2, pg_j3=j3)
np.savez_compressed(os.path.join(HERE, "reference_pose.npz"), **pose)
print("wrote", os.path.join(HERE, "reference_pose.npz"), {k: v.shape for k, v in pose.items()})
