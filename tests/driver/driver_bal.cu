#include "driver.h"
#include "snavely_reprojection_error.h"

namespace driver {
bool AddRunBalVariants(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                       const int* pb, const double* fdata, bool bulk, bool* handled);

bool AddRunBal(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
               const int* pb, const double* fdata, bool bulk, bool* handled) {
  using namespace ceres::examples;
  *handled = true;
  if (type == 0)
    return AddRunAllLosses<SnavelyReprojectionError, 2, 9, 3>(
        dp, loss_kind, a, b, n, pb, fdata, 2, bulk,
        [](const double* d) { return SnavelyReprojectionError(d[0], d[1]); });
  return AddRunBalVariants(dp, type, loss_kind, a, b, n, pb, fdata, bulk, handled);
}
}  // namespace driver
