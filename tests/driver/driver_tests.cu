#include "driver.h"
#include "test_functors.h"

namespace driver {
bool AddRunTests(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                 const int* pb, const double* fdata, bool bulk, bool* handled) {
  using namespace test_functors;
  *handled = true;
  switch (type) {
    case 3:
      return AddRunNoLoss<PointDisplacementError, 3, 3>(
          dp, loss_kind, n, pb, fdata, 3, bulk,
          [](const double* d) { return PointDisplacementError(d[0], d[1], d[2]); });
    case 5:
      return AddRunNoLoss<BinaryScalarCost, 1, 2, 2>(
          dp, loss_kind, n, pb, fdata, 1, bulk,
          [](const double* d) { return BinaryScalarCost(d[0]); });
    case 6:
      return AddRunNoLoss<TenParameterCost, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1>(
          dp, loss_kind, n, pb, fdata, 0, bulk, [](const double*) { return TenParameterCost(); });
    case 7:
      return AddRunNoLoss<OnlyFillsOneOutputFunctor, 2, 1>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return OnlyFillsOneOutputFunctor(); });
    case 8:
      return AddRunNoLoss<AffineTestCost<1, 3, true, 2, 3, 4>, 3, 2, 3, 4>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<1, 3, true, 2, 3, 4>(); });
    case 9:
      return AddRunNoLoss<AffineTestCost<1, 3, true, 4, 3, 2>, 3, 4, 3, 2>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<1, 3, true, 4, 3, 2>(); });
    case 10:
      return AddRunNoLoss<AffineTestCost<1, 2, true, 2, 3>, 2, 2, 3>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<1, 2, true, 2, 3>(); });
    case 11:
      return AddRunNoLoss<AffineTestCost<2, 3, true, 2, 4>, 3, 2, 4>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<2, 3, true, 2, 4>(); });
    case 12:
      return AddRunNoLoss<AffineTestCost<3, 4, true, 3, 4>, 4, 3, 4>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<3, 4, true, 3, 4>(); });
    case 13:
      return AddRunNoLoss<AffineTestCost<20, 3, false, 2, 3, 4>, 3, 2, 3, 4>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return AffineTestCost<20, 3, false, 2, 3, 4>(); });
    case 14:
      return AddRunNoLoss<ParameterSensitiveCost, 2, 2>(
          dp, loss_kind, n, pb, fdata, 0, bulk,
          [](const double*) { return ParameterSensitiveCost(); });
    case 17:
      return AddRunNoLoss<SqrtOfConstantCost, 1, 1>(
          dp, loss_kind, n, pb, fdata, 1, bulk,
          [](const double* d) { return SqrtOfConstantCost(d[0]); });
    case 16:
      return AddRunNoLoss<JetBatteryCost, 40, 2>(
          dp, loss_kind, n, pb, fdata, 0, bulk, [](const double*) { return JetBatteryCost(); });
  }
  *handled = false;
  return false;
}
}  // namespace driver
