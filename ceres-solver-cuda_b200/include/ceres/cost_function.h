// CostFunction, SizedCostFunction and AutoDiffCostFunction (host side).
//
// ProblemCUDA::AddResidualBlock takes a CostFunction* that must be an
// AutoDiffCostFunction<CostFunctor, kNumResiduals, Ns...> (reference:
// include/ceres/problem_cuda.h:440-453, include/ceres/autodiff_cost_function.h).
// The host Evaluate is used for residual blocks whose parameters are all constant
// (Program::RemoveFixedBlocks evaluates their cost once on the host,
// internal/ceres/program.cc:394-410) and by Problem::Evaluate-style callers; the
// hot path evaluates the same functor on the device.
#ifndef CERES_B200_COST_FUNCTION_H_
#define CERES_B200_COST_FUNCTION_H_

#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "ceres/jet.h"
#include "ceres/types.h"

namespace ceres {

class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  CostFunction(const CostFunction&) = delete;
  void operator=(const CostFunction&) = delete;
  virtual ~CostFunction() {}
  // jacobians[i] is a row-major num_residuals x parameter_block_sizes()[i] array,
  // or null; jacobians itself may be null.
  virtual bool Evaluate(double const* const* parameters, double* residuals,
                        double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  static_assert(kNumResiduals > 0, "the number of residuals must be static and positive");
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...};
  }
};

namespace internal {
template <typename Functor, typename T, std::size_t... Is>
inline bool HostCall(const Functor& f, T* const* p, T* out, std::index_sequence<Is...>) {
  return f(static_cast<const T*>(p[Is])..., out);
}
}  // namespace internal

template <typename CostFunctor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction final : public SizedCostFunction<kNumResiduals, Ns...> {
 public:
  explicit AutoDiffCostFunction(CostFunctor* functor, Ownership ownership = TAKE_OWNERSHIP)
      : functor_(functor), ownership_(ownership) {}
  ~AutoDiffCostFunction() override {
    if (ownership_ == DO_NOT_TAKE_OWNERSHIP) functor_.release();
  }

  bool Evaluate(double const* const* parameters, double* residuals,
                double** jacobians) const override {
    constexpr int kNB = sizeof...(Ns);
    constexpr int kNP = (Ns + ...);
    constexpr int sizes[kNB] = {Ns...};
    if (jacobians == nullptr) {
      double* p[kNB];
      for (int k = 0; k < kNB; ++k) p[k] = const_cast<double*>(parameters[k]);
      return internal::HostCall<CostFunctor, double>(*functor_, p, residuals,
                                                     std::make_index_sequence<kNB>{});
    }
    using JetT = Jet<double, kNP>;
    std::vector<JetT> x(kNP);
    JetT* unpacked[kNB];
    int offset = 0;
    for (int k = 0; k < kNB; ++k) {
      unpacked[k] = x.data() + offset;
      for (int j = 0; j < sizes[k]; ++j) x[offset + j] = JetT(parameters[k][j], offset + j);
      offset += sizes[k];
    }
    JetT out[kNumResiduals];
    for (int i = 0; i < kNumResiduals; ++i) out[i] = JetT::Filled(1e302, 1e302);
    if (!internal::HostCall<CostFunctor, JetT>(*functor_, unpacked, out,
                                               std::make_index_sequence<kNB>{}))
      return false;
    for (int i = 0; i < kNumResiduals; ++i) residuals[i] = out[i].a;
    offset = 0;
    for (int k = 0; k < kNB; ++k) {
      if (jacobians[k] != nullptr)
        for (int i = 0; i < kNumResiduals; ++i)
          for (int j = 0; j < sizes[k]; ++j)
            jacobians[k][i * sizes[k] + j] = out[i].v[offset + j];
      offset += sizes[k];
    }
    return true;
  }

  const CostFunctor& functor() const { return *functor_; }

 private:
  std::unique_ptr<CostFunctor> functor_;
  Ownership ownership_;
};

}  // namespace ceres

#endif  // CERES_B200_COST_FUNCTION_H_
