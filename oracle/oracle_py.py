"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liboracle.so (the CPU restatement of the reference's
ProgramEvaluator path, see oracle_eval.cc).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_FAST = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_eval.cc", "oracle_jet.h", "oracle_functors.h")]
    if force or not os.path.exists(so) or any(
            os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_fast():
    """The same source at -O3 -march=native (vectorised, FMA contraction allowed) for
    TIMING the CPU baseline only; compiled on the machine it runs on.  Parity checks
    always use the strict build (liboracle.so, -ffp-contract=off)."""
    import tempfile
    so = os.path.join(tempfile.gettempdir(), f"liboracle_fast_{os.getuid()}_{os.getpid()}.so")
    subprocess.check_call(["g++", "-std=c++17", "-O3", "-march=native", "-fPIC", "-pthread",
                           "-shared", "-Wno-array-bounds", "-o", so,
                           os.path.join(_HERE, "oracle_eval.cc")])
    return so


def lib(fast=False):
    global _LIB, _FAST
    if fast:
        if _FAST is None:
            _FAST = _declare(C.CDLL(build_fast()))
        return _FAST
    if _LIB is None:
        _LIB = _declare(C.CDLL(build()))
    return _LIB


def _declare(L):
    if True:
        L.oracle_problem_create.restype = C.c_void_p
        L.oracle_problem_create.argtypes = [
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_problem_destroy.argtypes = [C.c_void_p]
        L.oracle_problem_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_problem_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_problem_fixed_cost.restype = C.c_double
        L.oracle_problem_fixed_cost.argtypes = [C.c_void_p]
        L.oracle_problem_initial_state.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_problem_get_ints.restype = C.c_int64
        L.oracle_problem_get_ints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_problem_pb_table.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_problem_evaluate.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p]
        L.oracle_problem_plus.argtypes = [C.c_void_p] * 4
        L.oracle_cost_evaluate.argtypes = [C.c_int] + [C.c_void_p] * 4
        L.oracle_loss_evaluate.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.oracle_corrector.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_manifold_plus_jacobian.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_manifold_plus.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_angle_axis_rotate_point.argtypes = [C.c_void_p] * 3
        L.oracle_cost_type_info.argtypes = [C.c_int, C.c_void_p]
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_INT_ARRAYS = {
    "residual_layout": 0, "jacobian_per_residual_layout": 1, "jacobian_per_residual_offsets": 2,
    "program_rbs": 3, "program_pbs": 4, "col_block_size": 5, "col_block_pos": 6,
    "row_block_size": 7, "row_block_pos": 8, "row_cells_start": 9, "cell_block_id": 10,
    "cell_position": 11, "crs_rows": 12, "crs_cols": 13, "constant_pbs": 14,
    "jacobian_layout_storage": 15,
}


class OracleProblem:
    """Program + evaluator built from a ProblemSpec (see ceres-solver-cuda_b200/problems.py)."""

    def __init__(self, spec, jacobian_format=0, reduce=True, schur_reorder=False,
                 num_eliminate_blocks=None, fast=False):
        L = self.L = lib(fast)
        self.spec = spec
        self.h = L.oracle_problem_create(
            spec.num_pb, _p(spec.pb_size), _p(spec.pb_values), _p(spec.pb_constant),
            _p(spec.pb_manifold_kind), _p(spec.pb_manifold_param), spec.num_rb,
            _p(spec.rb_type), _p(spec.rb_pb), _p(spec.rb_loss_kind), _p(spec.rb_loss_a),
            _p(spec.rb_loss_b), _p(spec.fdata))
        ne = spec.num_eliminate_blocks if num_eliminate_blocks is None else num_eliminate_blocks
        self.jacobian_format = jacobian_format
        self.ok = bool(L.oracle_problem_build(self.h, int(reduce), int(schur_reorder), int(ne),
                                              int(jacobian_format)))
        dims = np.zeros(10, dtype=np.int64)
        L.oracle_problem_dims(self.h, _p(dims))
        (self.num_parameters, self.num_effective_parameters, self.num_residuals,
         self.num_residual_blocks, self.num_parameter_blocks, self.num_jacobian_values,
         self.values_size, self.num_cells, self.offsets_size,
         self.num_constant_parameters) = (int(x) for x in dims)
        self.fixed_cost = float(L.oracle_problem_fixed_cost(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.oracle_problem_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ints(self, name):
        L = self.L
        n = L.oracle_problem_get_ints(self.h, _INT_ARRAYS[name], None)
        out = np.zeros(n, dtype=np.int32)
        L.oracle_problem_get_ints(self.h, _INT_ARRAYS[name], _p(out))
        return out

    def pb_table(self):
        out = np.zeros((self.num_parameter_blocks, 4), dtype=np.int32)
        self.L.oracle_problem_pb_table(self.h, _p(out))
        return out

    def initial_state(self):
        s = np.zeros(self.num_parameters)
        self.L.oracle_problem_initial_state(self.h, _p(s))
        return s

    def evaluate(self, state=None, residuals=True, gradient=True, jacobian=True,
                 apply_loss_function=True, num_threads=1):
        """Returns (ok, cost, residuals, gradient, jacobian_values)."""
        if state is None:
            state = self.initial_state()
        state = np.ascontiguousarray(state, dtype=np.float64)
        cost = np.zeros(1)
        r = np.full(self.num_residuals, np.nan) if residuals else None
        g = np.full(self.num_effective_parameters, np.nan) if gradient else None
        j = np.full(self.values_size, np.nan) if jacobian else None
        ok = self.L.oracle_problem_evaluate(self.h, _p(state), int(apply_loss_function),
                                           int(num_threads), _p(cost), _p(r), _p(g), _p(j))
        return bool(ok), float(cost[0]), r, g, j

    def plus(self, state, delta):
        out = np.zeros(self.num_parameters)
        self.L.oracle_problem_plus(self.h, _p(np.ascontiguousarray(state)),
                                  _p(np.ascontiguousarray(delta)), _p(out))
        return out

    def dense_jacobian(self, values):
        """ToDenseMatrix of the BSM / CRS Jacobian (block_sparse_matrix.cc / crs)."""
        J = np.zeros((self.num_residuals, self.num_effective_parameters))
        if self.jacobian_format == 0:
            rs, rp = self.ints("row_block_size"), self.ints("row_block_pos")
            cs, cp = self.ints("col_block_size"), self.ints("col_block_pos")
            start, bid, pos = (self.ints("row_cells_start"), self.ints("cell_block_id"),
                               self.ints("cell_position"))
            for i in range(rs.size):
                for c in range(start[i], start[i + 1]):
                    b = bid[c]
                    blk = values[pos[c]:pos[c] + rs[i] * cs[b]].reshape(rs[i], cs[b])
                    J[rp[i]:rp[i] + rs[i], cp[b]:cp[b] + cs[b]] += blk
        else:
            rows, cols = self.ints("crs_rows"), self.ints("crs_cols")
            for r in range(self.num_residuals):
                for k in range(rows[r], rows[r + 1]):
                    J[r, cols[k]] += values[k]
        return J


def cost_evaluate(cost_type, fdata, params, want_jacobians=True):
    info = np.zeros(16, dtype=np.int32)
    lib().oracle_cost_type_info(cost_type, _p(info))
    nres, nb = int(info[0]), int(info[1])
    sizes = [int(x) for x in info[3:3 + nb]]
    params = np.ascontiguousarray(np.concatenate([np.asarray(p, float).ravel() for p in params]))
    fdata = np.ascontiguousarray(np.asarray(fdata, float))
    res = np.full(nres, np.nan)
    jac = np.full(nres * sum(sizes), np.nan) if want_jacobians else None
    ok = lib().oracle_cost_evaluate(cost_type, _p(fdata), _p(params), _p(res), _p(jac))
    jl = None
    if want_jacobians:
        jl, o = [], 0
        for s in sizes:
            jl.append(jac[o:o + nres * s].reshape(nres, s))
            o += nres * s
    return bool(ok), res, jl


def loss_evaluate(kind, a, b, s):
    rho = np.zeros(3)
    lib().oracle_loss_evaluate(kind, a, b, s, _p(rho))
    return rho


def corrector(sq_norm, rho, residuals, jacobian=None):
    rho = np.ascontiguousarray(rho, dtype=float)
    r = np.array(residuals, dtype=float)
    j = None if jacobian is None else np.array(jacobian, dtype=float)
    nr = r.size
    nc = 0 if j is None else j.size // nr
    lib().oracle_corrector(float(sq_norm), _p(rho), nr, nc, _p(r), _p(j))
    return r, j


def manifold_plus_jacobian(kind, param, x):
    x = np.ascontiguousarray(x, dtype=float)
    from_kind = {0: x.size, 2: 3, 3: 3}
    t = x.size - bin(param).count("1") if kind == 1 else from_kind.get(kind, x.size - 1)
    J = np.zeros((x.size, t))
    lib().oracle_manifold_plus_jacobian(kind, param, x.size, _p(x), _p(J))
    return J


def manifold_plus(kind, param, x, delta):
    x = np.ascontiguousarray(x, dtype=float)
    delta = np.ascontiguousarray(delta, dtype=float)
    out = np.zeros_like(x)
    lib().oracle_manifold_plus(kind, param, x.size, _p(x), _p(delta), _p(out))
    return out


def quaternion_to_angle_axis_jet(q):
    v, j = np.zeros(3), np.zeros(12)
    lib().oracle_quaternion_to_angle_axis_jet(_p(np.ascontiguousarray(q, float)), _p(v), _p(j))
    return v, j


def jet_battery(xy):
    buf = np.zeros(3 * 64)
    k = lib().oracle_jet_battery(_p(np.ascontiguousarray(xy, float)), _p(buf))
    return buf[:3 * k]


def angle_axis_rotate_point(aa, pt):
    out = np.zeros(3)
    lib().oracle_angle_axis_rotate_point(_p(np.ascontiguousarray(aa, float)),
                                         _p(np.ascontiguousarray(pt, float)), _p(out))
    return out
