// Pose-graph cost functors without Eigen (BASELINE.json config 5).
//
// RelativePoseError is the 6-residual SE(3) relative-pose error of the reference's
// internal/ceres/autodiff_benchmarks/relative_pose_error.h:46-92 (pose =
// [q(x,y,z,w), t], residual = log of the residual pose in SO(3) x R^3), and
// PoseGraph3dErrorTerm that of examples/slam/pose_graph_3d/pose_graph_3d_error_term.h:71-124.
// The reference versions use Eigen::Quaternion inside the functor, which is neither
// HOST_DEVICE nor available here; these are the same formulas on plain arrays.
#ifndef CERES_B200_EXAMPLES_RELATIVE_POSE_ERROR_H_
#define CERES_B200_EXAMPLES_RELATIVE_POSE_ERROR_H_

#include "ceres/cost_function.h"
#include "ceres/internal/cuda_defs.h"
#include "ceres/rotation.h"

namespace ceres {
namespace examples {

namespace pose_internal {
// Quaternions stored (x, y, z, w) like Eigen::Quaternion.
template <typename T>
HOST_DEVICE CERES_B200_INLINE void Product(const T a[4], const T b[4], T out[4]) {
  out[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  out[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  out[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  out[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
template <typename T>
HOST_DEVICE CERES_B200_INLINE void Conjugate(const T a[4], T out[4]) {
  out[0] = -a[0];
  out[1] = -a[1];
  out[2] = -a[2];
  out[3] = a[3];
}
// v + 2 w (u x v) + 2 u x (u x v)
template <typename T>
HOST_DEVICE CERES_B200_INLINE void Rotate(const T q[4], const T v[3], T out[3]) {
  T uv[3];
  CrossProduct(q, v, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  T uuv[3];
  CrossProduct(q, uv, uuv);
  out[0] = v[0] + q[3] * uv[0] + uuv[0];
  out[1] = v[1] + q[3] * uv[1] + uuv[1];
  out[2] = v[2] + q[3] * uv[2] + uuv[2];
}
}  // namespace pose_internal

struct RelativePoseError {
  // meas = [q_i_j (x, y, z, w), t_i_j]
  HOST_DEVICE explicit RelativePoseError(const double* meas) {
    for (int i = 0; i < 4; ++i) meas_q[i] = meas[i];
    for (int i = 0; i < 3; ++i) meas_t[i] = meas[4 + i];
  }

  template <typename T>
  HOST_DEVICE bool operator()(const T* const pose_i, const T* const pose_j, T* residuals) const {
    using namespace pose_internal;
    const T* q_w_i = pose_i;
    const T* t_w_i = pose_i + 4;
    const T* q_w_j = pose_j;
    const T* t_w_j = pose_j + 4;
    // Estimate of the relative pose from i to j.
    T q_j_w[4];
    Conjugate(q_w_j, q_j_w);
    T est_q_j_i[4];
    Product(q_j_w, q_w_i, est_q_j_i);
    const T dt[3] = {t_w_i[0] - t_w_j[0], t_w_i[1] - t_w_j[1], t_w_i[2] - t_w_j[2]};
    T est_t_j_i[3];
    Rotate(q_j_w, dt, est_t_j_i);
    // Residual pose.
    const T mq[4] = {T(meas_q[0]), T(meas_q[1]), T(meas_q[2]), T(meas_q[3])};
    T res_q[4];
    Product(mq, est_q_j_i, res_q);
    T res_t[3];
    Rotate(mq, est_t_j_i, res_t);
    // Log of the rotation part, Ceres quaternion order (w, x, y, z).
    const T res_q_ceres[4] = {res_q[3], res_q[0], res_q[1], res_q[2]};
    QuaternionToAngleAxis(res_q_ceres, residuals);
    residuals[3] = res_t[0] + meas_t[0];
    residuals[4] = res_t[1] + meas_t[1];
    residuals[5] = res_t[2] + meas_t[2];
    return true;
  }

  double meas_q[4];
  double meas_t[3];
};

struct PoseGraph3dErrorTerm {
  // d = [p_ab(3), q_ab(x, y, z, w), sqrt_information (row-major 6 x 6)]
  HOST_DEVICE explicit PoseGraph3dErrorTerm(const double* d) {
    for (int i = 0; i < 3; ++i) p_ab[i] = d[i];
    for (int i = 0; i < 4; ++i) q_ab[i] = d[3 + i];
    for (int i = 0; i < 36; ++i) sqrt_information[i] = d[7 + i];
  }

  template <typename T>
  HOST_DEVICE bool operator()(const T* const p_a, const T* const q_a, const T* const p_b,
                              const T* const q_b, T* residuals) const {
    using namespace pose_internal;
    T q_a_inverse[4];
    Conjugate(q_a, q_a_inverse);
    T q_ab_estimated[4];
    Product(q_a_inverse, q_b, q_ab_estimated);
    const T dp[3] = {p_b[0] - p_a[0], p_b[1] - p_a[1], p_b[2] - p_a[2]};
    T p_ab_estimated[3];
    Rotate(q_a_inverse, dp, p_ab_estimated);
    const T mq[4] = {T(q_ab[0]), T(q_ab[1]), T(q_ab[2]), T(q_ab[3])};
    T q_ab_estimated_conj[4];
    Conjugate(q_ab_estimated, q_ab_estimated_conj);
    T delta_q[4];
    Product(mq, q_ab_estimated_conj, delta_q);
    T e[6];
    e[0] = p_ab_estimated[0] - p_ab[0];
    e[1] = p_ab_estimated[1] - p_ab[1];
    e[2] = p_ab_estimated[2] - p_ab[2];
    e[3] = T(2.0) * delta_q[0];
    e[4] = T(2.0) * delta_q[1];
    e[5] = T(2.0) * delta_q[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      T acc = e[0] * sqrt_information[i * 6];
#pragma unroll
      for (int j = 1; j < 6; ++j) acc += e[j] * sqrt_information[i * 6 + j];
      residuals[i] = acc;
    }
    return true;
  }

  double p_ab[3];
  double q_ab[4];
  double sqrt_information[36];
};

}  // namespace examples
}  // namespace ceres

#endif  // CERES_B200_EXAMPLES_RELATIVE_POSE_ERROR_H_
