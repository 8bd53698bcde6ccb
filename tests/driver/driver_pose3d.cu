// (a translation unit of its own: the kernels of one residual-block type take minutes to
// compile, and make builds the units in parallel)
#include "driver.h"
#include "relative_pose_error.h"

namespace driver {
bool AddRunPose3d(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                  const int* pb, const double* fdata, bool bulk, bool* handled) {
  using namespace ceres::examples;
  *handled = true;
  if (type == 15)
    return AddRunCommonLosses<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>(
        dp, loss_kind, a, b, n, pb, fdata, 43, bulk,
        [](const double* d) { return PoseGraph3dErrorTerm(d); });
  *handled = false;
  return false;
}
}  // namespace driver
