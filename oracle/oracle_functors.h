// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_jet.h header).
//
// Rotation helpers and the cost functors of the hot path's workloads, restated
// for the oracle's plain-array Jet.  Each function cites the reference lines it
// follows.  Functors take their per-residual-block constants ("functor data",
// e.g. the observed pixel) as a `const double* d` so that one type serves every
// residual block of that kind.
#ifndef ORACLE_FUNCTORS_H_
#define ORACLE_FUNCTORS_H_

#include <cmath>

#include "oracle_jet.h"

namespace oracle {

// ---------------------------------------------------------------- rotation.h
// include/ceres/rotation.h:753-776 UnitQuaternionRotatePoint
template <typename T>
inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  T uv0 = q[2] * pt[2] - q[3] * pt[1];
  T uv1 = q[3] * pt[0] - q[1] * pt[2];
  T uv2 = q[1] * pt[1] - q[2] * pt[0];
  uv0 += uv0;
  uv1 += uv1;
  uv2 += uv2;
  result[0] = pt[0] + q[0] * uv0;
  result[1] = pt[1] + q[0] * uv1;
  result[2] = pt[2] + q[0] * uv2;
  result[0] += q[2] * uv2 - q[3] * uv1;
  result[1] += q[3] * uv0 - q[1] * uv2;
  result[2] += q[1] * uv1 - q[2] * uv0;
}

// include/ceres/rotation.h:778-797 QuaternionRotatePoint
template <typename T>
inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T scale =
      T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
  UnitQuaternionRotatePoint(unit, pt, result);
}

// include/ceres/rotation.h:830-901 AngleAxisRotatePoint
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta = hypot(angle_axis[0], angle_axis[1], angle_axis[2]);
  if (fpclassify(theta) != FP_ZERO) {
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse,
                    angle_axis[1] * theta_inverse,
                    angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1],
                             w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp =
        (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1],
                             angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// include/ceres/rotation.h:356-406 QuaternionToAngleAxis (wxyz order)
template <typename T>
inline void QuaternionToAngleAxis(const T* quaternion, T* angle_axis) {
  const T& q1 = quaternion[1];
  const T& q2 = quaternion[2];
  const T& q3 = quaternion[3];
  const T sin_theta = hypot(q1, q2, q3);
  if (fpclassify(sin_theta) != FP_ZERO) {
    const T& cos_theta = quaternion[0];
    const T two_theta =
        T(2.0) * ((cos_theta < T(0.0)) ? atan2(-sin_theta, -cos_theta)
                                       : atan2(sin_theta, cos_theta));
    const T k = two_theta / sin_theta;
    angle_axis[0] = q1 * k;
    angle_axis[1] = q2 * k;
    angle_axis[2] = q3 * k;
  } else {
    const T k(2.0);
    angle_axis[0] = q1 * k;
    angle_axis[1] = q2 * k;
    angle_axis[2] = q3 * k;
  }
}

// Eigen::Quaternion product / conjugate / rotate, storage order (x, y, z, w).
// Eigen is absent from /root/reference (un-vendored dependency, pinned
// "Eigen 3.4 or higher", README.md:10); this restates its published Hamilton
// product (Eigen/src/Geometry/Quaternion.h quat_product) and the
// _transformVector formula  v + 2w(u x v) + 2 u x (u x v)  in the form
// uv = u x v; uv += uv; v + w*uv + u x uv.
template <typename T>
inline void EigenQuatProduct(const T a[4], const T b[4], T out[4]) {
  // a = (ax, ay, az, aw)
  out[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  out[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  out[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  out[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
template <typename T>
inline void EigenQuatConjugate(const T a[4], T out[4]) {
  out[0] = -a[0]; out[1] = -a[1]; out[2] = -a[2]; out[3] = a[3];
}
template <typename T>
inline void EigenQuatRotate(const T q[4], const T v[3], T out[3]) {
  T uv[3] = {q[1] * v[2] - q[2] * v[1],
             q[2] * v[0] - q[0] * v[2],
             q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

// ------------------------------------------------------------------ functors
// examples/snavely_reprojection_error.h:52-101  <2, 9, 3>; d = {observed_x, observed_y}
struct SnavelyReprojectionError {
  template <typename T>
  bool operator()(const double* d, const T* const camera, const T* const point,
                  T* residuals) const {
    T p[3];
    AngleAxisRotatePoint(camera, point, p);
    p[0] += camera[3];
    p[1] += camera[4];
    p[2] += camera[5];
    const T xp = -p[0] / p[2];
    const T yp = -p[1] / p[2];
    const T& l1 = camera[7];
    const T& l2 = camera[8];
    const T r2 = xp * xp + yp * yp;
    const T distortion = 1.0 + r2 * (l1 + l2 * r2);
    const T& focal = camera[6];
    const T predicted_x = focal * distortion * xp;
    const T predicted_y = focal * distortion * yp;
    residuals[0] = predicted_x - d[0];
    residuals[1] = predicted_y - d[1];
    return true;
  }
};

// examples/snavely_reprojection_error.h:108-179 and
// internal/ceres/evaluator_cuda_test.cu.cc:171-230  <2, 10, 3>
struct SnavelyReprojectionErrorWithQuaternions {
  template <typename T>
  bool operator()(const double* d, const T* const camera, const T* const point,
                  T* residuals) const {
    T p[3];
    QuaternionRotatePoint(camera, point, p);
    p[0] += camera[4];
    p[1] += camera[5];
    p[2] += camera[6];
    const T xp = -p[0] / p[2];
    const T yp = -p[1] / p[2];
    const T& l1 = camera[8];
    const T& l2 = camera[9];
    const T r2 = xp * xp + yp * yp;
    const T distortion = 1.0 + r2 * (l1 + l2 * r2);
    const T& focal = camera[7];
    const T predicted_x = focal * distortion * xp;
    const T predicted_y = focal * distortion * yp;
    residuals[0] = predicted_x - d[0];
    residuals[1] = predicted_y - d[1];
    return true;
  }
};

// internal/ceres/evaluator_cuda_test.cu.cc:115-164  <2, 7, 3>
struct SnavelyReprojectionErrorNoRadialDistortion {
  template <typename T>
  bool operator()(const double* d, const T* const camera, const T* const point,
                  T* residuals) const {
    T p[3];
    AngleAxisRotatePoint(camera, point, p);
    p[0] += camera[3];
    p[1] += camera[4];
    p[2] += camera[5];
    const T xp = -p[0] / p[2];
    const T yp = -p[1] / p[2];
    const T& focal = camera[6];
    const T predicted_x = focal * xp;
    const T predicted_y = focal * yp;
    residuals[0] = predicted_x - d[0];
    residuals[1] = predicted_y - d[1];
    return true;
  }
};

// internal/ceres/evaluator_cuda_test.cu.cc:83-109  <3, 3>; d = {x, y, z}
struct PointDisplacementError {
  template <typename T>
  bool operator()(const double* d, const T* const point, T* residuals) const {
    residuals[0] = abs(d[0]) - abs(point[0]);
    residuals[1] = abs(d[1]) - abs(point[1]);
    residuals[2] = abs(d[2]) - abs(point[2]);
    return true;
  }
};

// internal/ceres/autodiff_benchmarks/relative_pose_error.h:46-92  <6, 7, 7>
// pose = [q(x,y,z,w), t(3)];  d = {meas_q(x,y,z,w), meas_t(3)}
struct RelativePoseError {
  template <typename T>
  bool operator()(const double* d, const T* const pose_i, const T* const pose_j,
                  T* residuals) const {
    const T* q_w_i = pose_i;
    const T* t_w_i = pose_i + 4;
    const T* q_w_j = pose_j;
    const T* t_w_j = pose_j + 4;
    T q_w_j_conj[4];
    EigenQuatConjugate(q_w_j, q_w_j_conj);
    T est_q_j_i[4];
    EigenQuatProduct(q_w_j_conj, q_w_i, est_q_j_i);
    const T dt[3] = {t_w_i[0] - t_w_j[0], t_w_i[1] - t_w_j[1], t_w_i[2] - t_w_j[2]};
    T est_t_j_i[3];
    EigenQuatRotate(q_w_j_conj, dt, est_t_j_i);
    const T meas_q[4] = {T(d[0]), T(d[1]), T(d[2]), T(d[3])};
    T res_q[4];
    EigenQuatProduct(meas_q, est_q_j_i, res_q);
    T res_t[3];
    EigenQuatRotate(meas_q, est_t_j_i, res_t);
    res_t[0] += d[4];
    res_t[1] += d[5];
    res_t[2] += d[6];
    const T res_q_ceres[4] = {res_q[3], res_q[0], res_q[1], res_q[2]};
    QuaternionToAngleAxis(res_q_ceres, residuals);
    residuals[3] = res_t[0];
    residuals[4] = res_t[1];
    residuals[5] = res_t[2];
    return true;
  }
};

// internal/ceres/autodiff_cost_function_cuda_test.cu.cc:42-54  <1, 2, 2>; d = {a}
struct BinaryScalarCost {
  template <typename T>
  bool operator()(const double* d, const T* const x, const T* const y, T* cost) const {
    cost[0] = x[0] * y[0] + x[1] * y[1] - T(d[0]);
    return true;
  }
};

// internal/ceres/autodiff_cost_function_cuda_test.cu.cc:128-146  <1, 1 x 10>
struct TenParameterCost {
  template <typename T>
  bool operator()(const double*, const T* const x0, const T* const x1, const T* const x2,
                  const T* const x3, const T* const x4, const T* const x5,
                  const T* const x6, const T* const x7, const T* const x8,
                  const T* const x9, T* cost) const {
    cost[0] = *x0 + *x1 + *x2 + *x3 + *x4 + *x5 + *x6 + *x7 + *x8 + *x9;
    return true;
  }
};

// internal/ceres/autodiff_cost_function_cuda_test.cu.cc:230-237  <2, 1>
struct OnlyFillsOneOutputFunctor {
  template <typename T>
  bool operator()(const double*, const T* x, T* output) const {
    output[0] = x[0];
    return true;
  }
};

// <1, 1>  r = x + sqrt(T(c)): with c == 0 the reference's dense Jets give the constant's
// derivative lanes 0 * inf = NaN (jet.h:617-621), so the CPU evaluator rejects the evaluation
// (residual_block.cc:110-129); see tests/test_gpu_robustness.py for the product's behaviour.
struct SqrtOfConstantCost {
  template <typename T>
  bool operator()(const double* data, const T* x, T* residual) const {
    residual[0] = x[0] + sqrt(T(data[0]));
    return true;
  }
};

// Autodiff restatement of internal/ceres/evaluator_test.cc:59-100
// ParameterIgnoringCostFunction<kFactor, kNumResiduals, Ns...>: residual i is
// (i + 1) and d r_i / d x_k[j] = kFactor * (j + 1).  As an autodiff functor it
// is affine in the parameters; the tests evaluate it at the all-zero state so
// the residuals are exactly i + 1 as in the reference tables.
template <int kFactor, int kNumResiduals, bool kSucceeds, int... Ns>
struct AffineTestCost {
  template <typename T>
  bool operator()(const double*, const T* const a, const T* const b, T* residuals) const {
    const T* params[] = {a, b};
    return Impl(params, residuals);
  }
  template <typename T>
  bool operator()(const double*, const T* const a, const T* const b, const T* const c,
                  T* residuals) const {
    const T* params[] = {a, b, c};
    return Impl(params, residuals);
  }
  template <typename T>
  static bool Impl(const T* const* params, T* residuals) {
    constexpr int sizes[] = {Ns...};
    for (int i = 0; i < kNumResiduals; ++i) {
      T r(static_cast<double>(i + 1));
      for (int k = 0; k < static_cast<int>(sizeof...(Ns)); ++k)
        for (int j = 0; j < sizes[k]; ++j)
          r += params[k][j] * static_cast<double>(kFactor * (j + 1));
      residuals[i] = r;
    }
    return kSucceeds;
  }
};

// Autodiff restatement of internal/ceres/evaluator_test.cc:573-596
// ParameterSensitiveCostFunction <2, 2>: r = (x1^2, x2^2).
struct ParameterSensitiveCost {
  template <typename T>
  bool operator()(const double*, const T* const x, T* residuals) const {
    residuals[0] = x[0] * x[0];
    residuals[1] = x[1] * x[1];
    return true;
  }
};

// examples/slam/pose_graph_3d/pose_graph_3d_error_term.h:71-124 <6, 3, 4, 3, 4>
// d = {p_ab(3), q_ab(x,y,z,w), sqrt_information row-major 6x6}
struct PoseGraph3dErrorTerm {
  template <typename T>
  bool operator()(const double* d, const T* const p_a, const T* const q_a,
                  const T* const p_b, const T* const q_b, T* residuals) const {
    T q_a_inverse[4];
    EigenQuatConjugate(q_a, q_a_inverse);
    T q_ab_estimated[4];
    EigenQuatProduct(q_a_inverse, q_b, q_ab_estimated);
    const T dp[3] = {p_b[0] - p_a[0], p_b[1] - p_a[1], p_b[2] - p_a[2]};
    T p_ab_estimated[3];
    EigenQuatRotate(q_a_inverse, dp, p_ab_estimated);
    const T meas_q[4] = {T(d[3]), T(d[4]), T(d[5]), T(d[6])};
    T q_ab_est_conj[4];
    EigenQuatConjugate(q_ab_estimated, q_ab_est_conj);
    T delta_q[4];
    EigenQuatProduct(meas_q, q_ab_est_conj, delta_q);
    T e[6];
    e[0] = p_ab_estimated[0] - d[0];
    e[1] = p_ab_estimated[1] - d[1];
    e[2] = p_ab_estimated[2] - d[2];
    e[3] = T(2.0) * delta_q[0];
    e[4] = T(2.0) * delta_q[1];
    e[5] = T(2.0) * delta_q[2];
    const double* S = d + 7;
    for (int i = 0; i < 6; ++i) {
      T acc = e[0] * S[i * 6 + 0];
      for (int j = 1; j < 6; ++j) acc += e[j] * S[i * 6 + j];
      residuals[i] = acc;
    }
    return true;
  }
};

// The 40 Jet operations of oracle/ref_arith.cc ref_jet_battery as a cost functor
// <40, 2>: residual k = op_k(x, y) (inputs of internal/ceres/jet_cuda_test.cu.cc).
struct JetBatteryCost {
  template <typename T>
  bool operator()(const double*, const T* const p, T* r) const {
    const T& x = p[0];
    const T& y = p[1];
    int k = 0;
    r[k++] = x + y; r[k++] = x - y; r[k++] = x * y; r[k++] = x / y; r[k++] = -x;
    r[k++] = x + 1.5; r[k++] = 1.5 - x; r[k++] = x * 1.5; r[k++] = 1.5 / x; r[k++] = x / 1.5;
    r[k++] = sqrt(x); r[k++] = exp(x); r[k++] = log(x); r[k++] = sin(x); r[k++] = cos(x);
    r[k++] = tan(x); r[k++] = asin(x / 3.0); r[k++] = acos(x / 3.0); r[k++] = atan(x);
    r[k++] = sinh(x); r[k++] = cosh(x); r[k++] = tanh(x); r[k++] = abs(-x); r[k++] = atan2(y, x);
    r[k++] = pow(x, 1.7); r[k++] = pow(x, y); r[k++] = hypot(x, y); r[k++] = hypot(x, y, x * y);
    r[k++] = cbrt(x); r[k++] = exp2(x); r[k++] = log2(x); r[k++] = log10(x); r[k++] = log1p(x);
    r[k++] = expm1(x); r[k++] = fmax(x, y); r[k++] = fmin(x, y); r[k++] = erf(x); r[k++] = erfc(x);
    r[k++] = copysign(x, -y); r[k++] = fma(x, y, x);
    return true;
  }
};

}  // namespace oracle

#endif  // ORACLE_FUNCTORS_H_
