// Bundle-adjustment cost functors (the workload of BASELINE.json configs 1-4).
//
// Same camera models and residual definition as the reference's
// examples/snavely_reprojection_error.h:52-179: a pinhole camera with two radial
// distortion terms in Bundler's convention (the camera looks down -z), residual =
// predicted - observed pixel.  Written against T only, so the same functor runs on
// the host (doubles) and inside the device kernel (ceres::Jet).
#ifndef CERES_B200_EXAMPLES_SNAVELY_REPROJECTION_ERROR_H_
#define CERES_B200_EXAMPLES_SNAVELY_REPROJECTION_ERROR_H_

#include "ceres/cost_function.h"
#include "ceres/internal/cuda_defs.h"
#include "ceres/rotation.h"

namespace ceres {
namespace examples {

namespace snavely_internal {
// Shared tail: perspective division, radial distortion, focal scaling.
template <typename T>
HOST_DEVICE CERES_B200_INLINE void Project(const T p[3], const T& focal, const T& l1, const T& l2,
                                          double observed_x, double observed_y, T* residuals) {
  const T xp = -p[0] / p[2];
  const T yp = -p[1] / p[2];
  const T r2 = xp * xp + yp * yp;
  const T distortion = 1.0 + r2 * (l1 + l2 * r2);
  const T predicted_x = focal * distortion * xp;
  const T predicted_y = focal * distortion * yp;
  residuals[0] = predicted_x - observed_x;
  residuals[1] = predicted_y - observed_y;
}
}  // namespace snavely_internal

// camera = [angle-axis(3), translation(3), focal, k1, k2]; point = [X, Y, Z].
struct SnavelyReprojectionError {
  HOST_DEVICE SnavelyReprojectionError(double observed_x, double observed_y)
      : observed_x(observed_x), observed_y(observed_y) {}

  template <typename T>
  HOST_DEVICE bool operator()(const T* const camera, const T* const point, T* residuals) const {
    T p[3];
    AngleAxisRotatePoint(camera, point, p);
    p[0] += camera[3];
    p[1] += camera[4];
    p[2] += camera[5];
    snavely_internal::Project(p, camera[6], camera[7], camera[8], observed_x, observed_y,
                              residuals);
    return true;
  }

  static CostFunction* Create(double observed_x, double observed_y) {
    return new AutoDiffCostFunction<SnavelyReprojectionError, 2, 9, 3>(
        new SnavelyReprojectionError(observed_x, observed_y));
  }

  double observed_x;
  double observed_y;
};

// camera = [quaternion(w,x,y,z), translation(3), focal, k1, k2]; the quaternion
// need not be normalised (QuaternionRotatePoint normalises).
struct SnavelyReprojectionErrorWithQuaternions {
  HOST_DEVICE SnavelyReprojectionErrorWithQuaternions(double observed_x, double observed_y)
      : observed_x(observed_x), observed_y(observed_y) {}

  template <typename T>
  HOST_DEVICE bool operator()(const T* const camera, const T* const point, T* residuals) const {
    T p[3];
    QuaternionRotatePoint(camera, point, p);
    p[0] += camera[4];
    p[1] += camera[5];
    p[2] += camera[6];
    snavely_internal::Project(p, camera[7], camera[8], camera[9], observed_x, observed_y,
                              residuals);
    return true;
  }

  static CostFunction* Create(double observed_x, double observed_y) {
    return new AutoDiffCostFunction<SnavelyReprojectionErrorWithQuaternions, 2, 10, 3>(
        new SnavelyReprojectionErrorWithQuaternions(observed_x, observed_y));
  }

  double observed_x;
  double observed_y;
};

}  // namespace examples
}  // namespace ceres

#endif  // CERES_B200_EXAMPLES_SNAVELY_REPROJECTION_ERROR_H_
