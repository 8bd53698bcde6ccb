// Rotation helpers usable from HOST_DEVICE cost functors.
//
// Same signatures and conventions as the reference's include/ceres/rotation.h
// (quaternions are (w, x, y, z); angle-axis is a 3-vector whose norm is the
// angle): AngleAxisRotatePoint :830-901, QuaternionRotatePoint :778-797,
// UnitQuaternionRotatePoint :753-776, QuaternionProduct :799-813,
// CrossProduct/DotProduct :815-828, AngleAxisToQuaternion :322-353,
// QuaternionToAngleAxis :356-406.  Only the helpers cost functors on the
// evaluation path use are provided.
#ifndef CERES_B200_ROTATION_H_
#define CERES_B200_ROTATION_H_

#include "ceres/internal/cuda_defs.h"
#include "ceres/jet.h"

namespace ceres {

template <typename T>
HOST_DEVICE CERES_B200_INLINE T DotProduct(const T x[3], const T y[3]) {
  return (x[0] * y[0] + x[1] * y[1] + x[2] * y[2]);
}

template <typename T>
HOST_DEVICE CERES_B200_INLINE void CrossProduct(const T x[3], const T y[3], T out[3]) {
  out[0] = x[1] * y[2] - x[2] * y[1];
  out[1] = x[2] * y[0] - x[0] * y[2];
  out[2] = x[0] * y[1] - x[1] * y[0];
}

// zw = z * w (Hamilton product).
template <typename T>
HOST_DEVICE CERES_B200_INLINE void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}

// Rotates pt by the unit quaternion q:  pt + 2 q0 (u x pt) + 2 u x (u x pt), u = q[1..3].
template <typename T>
HOST_DEVICE CERES_B200_INLINE void UnitQuaternionRotatePoint(const T q[4], const T pt[3],
                                                            T result[3]) {
  T uv[3];
  CrossProduct(q + 1, pt, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  result[0] = pt[0] + q[0] * uv[0];
  result[1] = pt[1] + q[0] * uv[1];
  result[2] = pt[2] + q[0] * uv[2];
  result[0] += q[2] * uv[2] - q[3] * uv[1];
  result[1] += q[3] * uv[0] - q[1] * uv[2];
  result[2] += q[1] * uv[1] - q[2] * uv[0];
}

// As above for a quaternion of any non-zero norm.
template <typename T>
HOST_DEVICE CERES_B200_INLINE void QuaternionRotatePoint(const T q[4], const T pt[3],
                                                        T result[3]) {
  const T scale = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
  UnitQuaternionRotatePoint(unit, pt, result);
}

// Rodrigues' formula away from zero; first-order Taylor expansion R = I + [w]x at
// exactly zero angle so that Jets get meaningful derivatives there.
template <typename T>
HOST_DEVICE CERES_B200_INLINE void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3],
                                                       T result[3]) {
  const T theta = hypot(angle_axis[0], angle_axis[1], angle_axis[2]);
  if (fpclassify(theta) != FP_ZERO) {
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse,
                    angle_axis[2] * theta_inverse};
    T w_cross_pt[3];
    CrossProduct(w, pt, w_cross_pt);
    const T tmp = DotProduct(w, pt) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    T w_cross_pt[3];
    CrossProduct(angle_axis, pt, w_cross_pt);
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

template <typename T>
HOST_DEVICE CERES_B200_INLINE void AngleAxisToQuaternion(const T* angle_axis, T* quaternion) {
  const T& a0 = angle_axis[0];
  const T& a1 = angle_axis[1];
  const T& a2 = angle_axis[2];
  const T theta = hypot(a0, a1, a2);
  if (fpclassify(theta) != FP_ZERO) {
    const T half_theta = theta * T(0.5);
    const T k = sin(half_theta) / theta;
    quaternion[0] = cos(half_theta);
    quaternion[1] = a0 * k;
    quaternion[2] = a1 * k;
    quaternion[3] = a2 * k;
  } else {
    const T k(0.5);
    quaternion[0] = T(1.0);
    quaternion[1] = a0 * k;
    quaternion[2] = a1 * k;
    quaternion[3] = a2 * k;
  }
}

template <typename T>
HOST_DEVICE CERES_B200_INLINE void QuaternionToAngleAxis(const T* quaternion, T* angle_axis) {
  const T& q1 = quaternion[1];
  const T& q2 = quaternion[2];
  const T& q3 = quaternion[3];
  const T sin_theta = hypot(q1, q2, q3);
  if (fpclassify(sin_theta) != FP_ZERO) {
    const T& cos_theta = quaternion[0];
    // cos < 0 means the angle 2 theta exceeds pi; use the equivalent 2 theta - 2 pi.
    const T two_theta = T(2.0) * ((cos_theta < T(0.0)) ? atan2(-sin_theta, -cos_theta)
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

}  // namespace ceres

#endif  // CERES_B200_ROTATION_H_
