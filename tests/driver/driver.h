// Test/bench driver: replays a ProblemSpec (ceres-solver-cuda_b200/problems.py)
// through the product's public C++ API (ProblemCUDA::AddResidualBlock<...>,
// SetManifold, SetParameterBlockConstant, Evaluator::Create / Evaluate) and exposes
// the result to Python through a small C interface.  The templated calls are
// split over several translation units so they compile in parallel.
#ifndef TESTS_DRIVER_DRIVER_H_
#define TESTS_DRIVER_DRIVER_H_

#include <array>
#include <map>
#include <memory>
#include <tuple>
#include <utility>
#include <vector>

#include "ceres/problem_cuda.h"

namespace driver {

// Loss kinds / manifold kinds / cost type ids: see problems.py.
enum LossKind { kNone = 0, kTrivial, kHuber, kCauchy, kScaledHuber, kScaledCauchy, kScaledTrivial,
                kConvexTest };

// Problem::Options::evaluation_callback of every driver problem: counts the notifications and
// remembers what it saw (the flags, and the user's parameter values at that moment).
struct RecordingCallback : ceres::EvaluationCallback {
  int calls = 0;
  bool evaluate_jacobians = false, new_evaluation_point = false;
  double user_value_sum = 0.0;
  const std::vector<double>* values = nullptr;
  void PrepareForEvaluation(bool jacobians, bool new_point) override {
    ++calls;
    evaluate_jacobians = jacobians;
    new_evaluation_point = new_point;
    user_value_sum = 0.0;
    if (values) for (double v : *values) user_value_sum += v;
  }
};

struct DriverProblem {
  RecordingCallback callback;      // (before `problem`: its options point at it)
  std::vector<double> values;      // user state of every parameter block
  std::vector<int64_t> pb_offset;  // into values
  std::vector<int> pb_size;
  ceres::ProblemCUDA problem;
  std::map<std::tuple<int, double, double>, void*> losses;
  std::unique_ptr<ceres::internal::Program> full_program, program;
  std::unique_ptr<ceres::internal::Evaluator> evaluator;
  std::unique_ptr<ceres::internal::SparseMatrix> jacobian;
  ceres::internal::JacobianLayout layout;  // layout-only builds (no device)
  double fixed_cost = 0.0;
  std::string error;
  // results of the last drv_problem_evaluate
  std::vector<double> pe_residuals, pe_gradient;
  ceres::CRSMatrix pe_jacobian;
  // with_callback: install the recording EvaluationCallback (it makes every Evaluate copy the
  // state into the user's parameter blocks first, so the benchmarks leave it out)
  explicit DriverProblem(bool with_callback) : problem(MakeOptions(with_callback ? &callback : nullptr)) {
    callback.values = &values;
  }
  static ceres::Problem::Options MakeOptions(RecordingCallback* callback) {
    ceres::Problem::Options o;
    o.evaluation_callback = callback;
    return o;
  }
  double* pb(int id) { return values.data() + pb_offset[id]; }

  template <typename Loss, typename... Args>
  Loss* GetLoss(int kind, double a, double b, Args... args) {
    auto key = std::make_tuple(kind, a, b);
    auto it = losses.find(key);
    if (it != losses.end()) return static_cast<Loss*>(it->second);
    Loss* l = new Loss(args...);
    losses[key] = l;
    return l;
  }
};

template <typename F, int kRes, int... Ns, typename Loss, typename Make, std::size_t... Is>
void AddRunImpl(DriverProblem& dp, Loss* loss, int n, const int* pb, const double* fdata, int flen,
                bool bulk, Make make, std::index_sequence<Is...>) {
  constexpr int kNB = sizeof...(Ns);
  if (bulk) {
    std::vector<F> functors;
    functors.reserve(n);
    std::vector<double*> ptrs(static_cast<size_t>(n) * kNB);
    for (int i = 0; i < n; ++i) {
      functors.push_back(make(fdata + static_cast<size_t>(i) * flen));
      for (int j = 0; j < kNB; ++j)
        ptrs[static_cast<size_t>(i) * kNB + j] = dp.pb(pb[static_cast<size_t>(i) * kNB + j]);
    }
    dp.problem.AddResidualBlocks<F, kRes, Ns...>(n, functors.data(), loss, ptrs.data());
  } else {
    for (int i = 0; i < n; ++i) {
      auto* cost_function = new ceres::AutoDiffCostFunction<F, kRes, Ns...>(
          new F(make(fdata + static_cast<size_t>(i) * flen)));
      const int* ids = pb + static_cast<size_t>(i) * kNB;
      dp.problem.AddResidualBlock<F, kRes, Ns...>(cost_function, loss, dp.pb(ids[Is])...);
    }
  }
}

template <typename F, int kRes, int... Ns, typename Loss, typename Make>
void AddRun(DriverProblem& dp, Loss* loss, int n, const int* pb, const double* fdata, int flen,
            bool bulk, Make make) {
  AddRunImpl<F, kRes, Ns...>(dp, loss, n, pb, fdata, flen, bulk, make,
                             std::make_index_sequence<sizeof...(Ns)>{});
}

// Dispatch on the loss kind for cost types that support every loss.
template <typename F, int kRes, int... Ns, typename Make>
bool AddRunAllLosses(DriverProblem& dp, int loss_kind, double a, double b, int n, const int* pb,
                     const double* fdata, int flen, bool bulk, Make make) {
  using namespace ceres;
  switch (loss_kind) {
    case kNone:
      AddRun<F, kRes, Ns...>(dp, static_cast<TrivialLossCUDA*>(nullptr), n, pb, fdata, flen, bulk,
                             make);
      return true;
    case kTrivial:
      AddRun<F, kRes, Ns...>(dp, dp.GetLoss<TrivialLossCUDA>(loss_kind, a, b), n, pb, fdata, flen,
                             bulk, make);
      return true;
    case kHuber:
      AddRun<F, kRes, Ns...>(dp, dp.GetLoss<HuberLossCUDA>(loss_kind, a, b, a), n, pb, fdata, flen,
                             bulk, make);
      return true;
    case kCauchy:
      AddRun<F, kRes, Ns...>(dp, dp.GetLoss<CauchyLossCUDA>(loss_kind, a, b, a), n, pb, fdata,
                             flen, bulk, make);
      return true;
    case kScaledHuber:
      AddRun<F, kRes, Ns...>(
          dp, dp.GetLoss<ScaledLossCUDA<HuberLossCUDA>>(loss_kind, a, b, HuberLossCUDA(a), b), n,
          pb, fdata, flen, bulk, make);
      return true;
    case kScaledCauchy:
      AddRun<F, kRes, Ns...>(
          dp, dp.GetLoss<ScaledLossCUDA<CauchyLossCUDA>>(loss_kind, a, b, CauchyLossCUDA(a), b), n,
          pb, fdata, flen, bulk, make);
      return true;
    case kScaledTrivial:
      AddRun<F, kRes, Ns...>(
          dp, dp.GetLoss<ScaledLossCUDA<TrivialLossCUDA>>(loss_kind, a, b, TrivialLossCUDA(), b),
          n, pb, fdata, flen, bulk, make);
      return true;
  }
  return false;
}

// null / Huber / Cauchy only.
template <typename F, int kRes, int... Ns, typename Make>
bool AddRunCommonLosses(DriverProblem& dp, int loss_kind, double a, double b, int n, const int* pb,
                        const double* fdata, int flen, bool bulk, Make make) {
  using namespace ceres;
  switch (loss_kind) {
    case kNone:
      AddRun<F, kRes, Ns...>(dp, static_cast<TrivialLossCUDA*>(nullptr), n, pb, fdata, flen, bulk,
                             make);
      return true;
    case kHuber:
      AddRun<F, kRes, Ns...>(dp, dp.GetLoss<HuberLossCUDA>(loss_kind, a, b, a), n, pb, fdata, flen,
                             bulk, make);
      return true;
    case kCauchy:
      AddRun<F, kRes, Ns...>(dp, dp.GetLoss<CauchyLossCUDA>(loss_kind, a, b, a), n, pb, fdata,
                             flen, bulk, make);
      return true;
  }
  return false;
}

template <typename F, int kRes, int... Ns, typename Make>
bool AddRunNoLoss(DriverProblem& dp, int loss_kind, int n, const int* pb, const double* fdata,
                  int flen, bool bulk, Make make) {
  if (loss_kind != kNone) return false;
  AddRun<F, kRes, Ns...>(dp, static_cast<ceres::TrivialLossCUDA*>(nullptr), n, pb, fdata, flen,
                         bulk, make);
  return true;
}

// One function per translation unit; each returns false for types it does not own.
bool AddRunBal(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
               const int* pb, const double* fdata, bool bulk, bool* handled);
bool AddRunPose(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                const int* pb, const double* fdata, bool bulk, bool* handled);
bool AddRunTests(DriverProblem& dp, int type, int loss_kind, double a, double b, int n,
                 const int* pb, const double* fdata, bool bulk, bool* handled);

}  // namespace driver

#endif  // TESTS_DRIVER_DRIVER_H_
