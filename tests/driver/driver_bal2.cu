// (a translation unit of its own: see driver_pose3d.cu)
#include "driver.h"
#include "snavely_reprojection_error.h"
#include "test_functors.h"

namespace driver {
bool AddRunBalVariants(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                       const int* pb, const double* fdata, bool bulk, bool* handled) {
  using namespace ceres::examples;
  *handled = true;
  switch (type) {
    case 1:
      return AddRunCommonLosses<SnavelyReprojectionErrorWithQuaternions, 2, 10, 3>(
          dp, loss_kind, a, b, n, pb, fdata, 2, bulk,
          [](const double* d) { return SnavelyReprojectionErrorWithQuaternions(d[0], d[1]); });
    case 2:
      if (loss_kind == kConvexTest) {
        AddRun<test_functors::SnavelyReprojectionErrorNoRadialDistortion, 2, 7, 3>(
            dp, dp.GetLoss<test_functors::ConvexTestLoss>(loss_kind, a, b, a), n, pb, fdata, 2, bulk,
            [](const double* d) {
              return test_functors::SnavelyReprojectionErrorNoRadialDistortion(d[0], d[1]);
            });
        return true;
      }
      return AddRunCommonLosses<test_functors::SnavelyReprojectionErrorNoRadialDistortion, 2, 7, 3>(
          dp, loss_kind, a, b, n, pb, fdata, 2, bulk, [](const double* d) {
            return test_functors::SnavelyReprojectionErrorNoRadialDistortion(d[0], d[1]);
          });
  }
  *handled = false;
  return false;
}
}  // namespace driver
