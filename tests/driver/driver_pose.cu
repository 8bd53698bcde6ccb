#include "driver.h"
#include "relative_pose_error.h"

namespace driver {
bool AddRunPose3d(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                  const int* pb, const double* fdata, bool bulk, bool* handled);

bool AddRunPose(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                const int* pb, const double* fdata, bool bulk, bool* handled) {
  using namespace ceres::examples;
  *handled = true;
  if (type == 4)
    return AddRunCommonLosses<RelativePoseError, 6, 7, 7>(
        dp, loss_kind, a, b, n, pb, fdata, 7, bulk,
        [](const double* d) { return RelativePoseError(d); });
  return AddRunPose3d(dp, type, loss_kind, a, b, n, pb, fdata, bulk, handled);
}
}  // namespace driver
