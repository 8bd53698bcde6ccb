// ProblemCUDA: the user-facing entry point of the evaluation path.
//
// Source compatible with the reference's ceres::ProblemCUDA
// (include/ceres/problem_cuda.h:85-486):
//
//   ceres::ProblemCUDA problem;
//   problem.AddResidualBlock<SnavelyReprojectionError, 2, 9, 3>(
//       cost_function, new ceres::HuberLossCUDA(1.0), camera, point);
//   problem.SetManifold(camera, new ceres::SubsetManifold(9, {0}));
//
// The cost functor type, the number of residuals and the parameter block sizes are
// template arguments because the device kernel is instantiated for them here, in
// the user's translation unit (which therefore has to be compiled by nvcc, as with
// the reference: README.md:18-19).  Everything that is not a template lives in
// libceres_b200.so behind the C ABI of ceres_b200.h.
//
// Beyond the reference: AddResidualBlocks() adds residual blocks in bulk from
// arrays, avoiding per-block heap objects.
#ifndef CERES_B200_PROBLEM_CUDA_H_
#define CERES_B200_PROBLEM_CUDA_H_

#include <array>
#include <cstdio>
#include <memory>
#include <typeindex>
#include <typeinfo>
#include <unordered_map>
#include <vector>

#include "ceres/cost_function.h"
#include "ceres/internal/evaluate_kernel.cuh"
#include "ceres/internal/evaluator.h"
#include "ceres/internal/program.h"
#include "ceres/loss_function_cuda.h"
#include "ceres/manifold.h"
#include "ceres/problem.h"
#include "ceres/solver.h"
#include "ceres/types.h"

namespace ceres {

namespace internal {
template <typename CostFunctor, typename LossFunctionCUDA, int kNumResiduals, int... Ns>
struct ResidualTypeTag {};

template <typename LossFunctionCUDA>
void HostLossThunk(const void* loss, double s, double rho[3]) {
  static_cast<const LossFunctionCUDA*>(loss)->Evaluate(s, rho);
}
template <typename CostFunctor, int kNumResiduals, int... Ns>
bool HostFunctorThunk(const void* functor, double const* const* parameters, double* residuals,
                      double** jacobians) {
  AutoDiffCostFunction<CostFunctor, kNumResiduals, Ns...> cost_function(
      const_cast<CostFunctor*>(static_cast<const CostFunctor*>(functor)),
      DO_NOT_TAKE_OWNERSHIP);
  return cost_function.Evaluate(parameters, residuals, jacobians);
}
}  // namespace internal

class ProblemCUDA {
 public:
  ProblemCUDA()
      : problem_(new Problem),
        registered_cuda_evaluators_(
            new internal::RegisteredCUDAEvaluators(problem_->mutable_impl())) {}
  explicit ProblemCUDA(const Problem::Options& options)
      : problem_(new Problem(options)),
        registered_cuda_evaluators_(
            new internal::RegisteredCUDAEvaluators(problem_->mutable_impl())) {}
  ProblemCUDA(ProblemCUDA&&) = default;
  ProblemCUDA& operator=(ProblemCUDA&&) = default;
  ProblemCUDA(const ProblemCUDA&) = delete;
  ProblemCUDA& operator=(const ProblemCUDA&) = delete;

  // Adds a residual block.  cost_function must be an
  // AutoDiffCostFunction<CostFunctor, kNumResiduals, Ns...>; a null loss_function
  // means the trivial loss.
  template <typename CostFunctor, int kNumResiduals, int... Ns, typename... Ts,
            typename LossFunctionCUDA>
  ResidualBlockId AddResidualBlock(CostFunction* cost_function, LossFunctionCUDA* loss_function,
                                   double* x0, Ts*... xs) {
    if (!loss_function) {
      return AddResidualBlock<CostFunctor, kNumResiduals, Ns...>(cost_function, nullptr, x0,
                                                                  xs...);
    }
    return InternalAddResidualBlock<CostFunctor, kNumResiduals, Ns...>(cost_function,
                                                                       loss_function, x0, xs...);
  }

  template <typename CostFunctor, int kNumResiduals, int... Ns, typename... Ts>
  ResidualBlockId AddResidualBlock(CostFunction* cost_function, std::nullptr_t, double* x0,
                                   Ts*... xs) {
    return InternalAddResidualBlock<CostFunctor, kNumResiduals, Ns...>(
        cost_function, static_cast<TrivialLossCUDA*>(nullptr), x0, xs...);
  }

  // Bulk form: n residual blocks of one type.  functors[i] is copied;
  // parameter_blocks is [n][sizeof...(Ns)] pointers to parameter blocks (added to
  // the problem on first use, like AddResidualBlock does).
  template <typename CostFunctor, int kNumResiduals, int... Ns, typename LossFunctionCUDA>
  void AddResidualBlocks(int n, const CostFunctor* functors, LossFunctionCUDA* loss_function,
                         double* const* parameter_blocks) {
    static TrivialLossCUDA trivial;
    const int type = RegisterType<CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>();
    const void* loss = loss_function ? static_cast<const void*>(loss_function)
                                     : static_cast<const void*>(&trivial);
    TakeLossOwnership(loss_function);
    constexpr int kNB = sizeof...(Ns);
    for (int i = 0; i < n; ++i)
      problem_->mutable_impl()->AddResidualBlock(type, nullptr, functors + i, loss,
                                                 parameter_blocks + static_cast<size_t>(i) * kNB);
  }

  void AddParameterBlock(double* values, int size) { problem_->AddParameterBlock(values, size); }
  void AddParameterBlock(double* values, int size, Manifold* manifold) {
    problem_->AddParameterBlock(values, size, manifold);
  }
  void SetParameterBlockConstant(const double* values) {
    problem_->SetParameterBlockConstant(values);
  }
  void SetParameterBlockVariable(double* values) { problem_->SetParameterBlockVariable(values); }
  bool IsParameterBlockConstant(const double* values) const {
    return problem_->IsParameterBlockConstant(values);
  }
  void SetManifold(double* values, Manifold* manifold) { problem_->SetManifold(values, manifold); }
  const Manifold* GetManifold(const double* values) const { return problem_->GetManifold(values); }
  bool HasManifold(const double* values) const { return problem_->HasManifold(values); }
  void SetParameterLowerBound(double* values, int index, double lower_bound) {
    problem_->SetParameterLowerBound(values, index, lower_bound);
  }
  void SetParameterUpperBound(double* values, int index, double upper_bound) {
    problem_->SetParameterUpperBound(values, index, upper_bound);
  }
  double GetParameterUpperBound(const double* values, int index) const {
    return problem_->GetParameterUpperBound(values, index);
  }
  double GetParameterLowerBound(const double* values, int index) const {
    return problem_->GetParameterLowerBound(values, index);
  }
  int NumParameterBlocks() const { return problem_->NumParameterBlocks(); }
  int NumParameters() const { return problem_->NumParameters(); }
  int NumResidualBlocks() const { return problem_->NumResidualBlocks(); }
  int NumResiduals() const { return problem_->NumResiduals(); }
  int ParameterBlockSize(const double* values) const {
    return problem_->ParameterBlockSize(values);
  }
  int ParameterBlockTangentSize(const double* values) const {
    return problem_->ParameterBlockTangentSize(values);
  }
  bool HasParameterBlock(const double* values) const {
    return problem_->HasParameterBlock(values);
  }
  void GetParameterBlocks(std::vector<double*>* parameter_blocks) const {
    problem_->GetParameterBlocks(parameter_blocks);
  }
  void GetResidualBlocks(std::vector<ResidualBlockId>* residual_blocks) const {
    problem_->GetResidualBlocks(residual_blocks);
  }
  void GetParameterBlocksForResidualBlock(ResidualBlockId residual_block,
                                          std::vector<double*>* parameter_blocks) const {
    problem_->GetParameterBlocksForResidualBlock(residual_block, parameter_blocks);
  }
  const CostFunction* GetCostFunctionForResidualBlock(ResidualBlockId residual_block) const {
    return problem_->GetCostFunctionForResidualBlock(residual_block);
  }
  void GetResidualBlocksForParameterBlock(const double* values,
                                          std::vector<ResidualBlockId>* residual_blocks) const {
    problem_->GetResidualBlocksForParameterBlock(values, residual_blocks);
  }
  // problem_cuda.h:364-396.  Evaluate runs on the CUDA evaluator (the reference forwards to
  // the CPU one); EvaluateResidualBlock is a host evaluation of one block.
  bool Evaluate(const Problem::EvaluateOptions& evaluate_options, double* cost,
                std::vector<double>* residuals, std::vector<double>* gradient,
                CRSMatrix* jacobian) {
    return problem_->Evaluate(evaluate_options, cost, residuals, gradient, jacobian);
  }
  bool EvaluateResidualBlock(ResidualBlockId residual_block_id, bool apply_loss_function,
                             double* cost, double* residuals, double** jacobians) const {
    return problem_->EvaluateResidualBlock(residual_block_id, apply_loss_function, cost,
                                           residuals, jacobians);
  }
  bool EvaluateResidualBlockAssumingParametersUnchanged(ResidualBlockId residual_block_id,
                                                        bool apply_loss_function, double* cost,
                                                        double* residuals,
                                                        double** jacobians) const {
    return problem_->EvaluateResidualBlock(residual_block_id, apply_loss_function, cost,
                                           residuals, jacobians);
  }
  const Problem::Options& options() const { return problem_->options(); }

  Problem* mutable_problem() { return problem_.get(); }
  internal::RegisteredCUDAEvaluators* mutable_registered_cuda_evaluators() {
    return registered_cuda_evaluators_.get();
  }

 private:
  template <typename CostFunctor, typename LossFunctionCUDA, int kNumResiduals, int... Ns>
  int RegisterType() {
    static_assert(kNumResiduals != DYNAMIC,
                  "Can't use the CUDA evaluator if the number of residuals is ceres::DYNAMIC.");
    using Tag = internal::ResidualTypeTag<CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>;
    static const cb200_residual_type desc =
        internal::MakeResidualType<CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>();
    internal::ProblemImpl* impl = problem_->mutable_impl();
    const int type = impl->FindOrAddType(std::type_index(typeid(Tag)), desc);
    internal::ResidualTypeStore& store = impl->types()[type];
    store.host_loss = &internal::HostLossThunk<LossFunctionCUDA>;
    store.host_functor = &internal::HostFunctorThunk<CostFunctor, kNumResiduals, Ns...>;
    return type;
  }

  template <typename LossFunctionCUDA>
  void TakeLossOwnership(LossFunctionCUDA* loss_function) {
    if (loss_function && problem_->options().loss_function_ownership == TAKE_OWNERSHIP) {
      LossFunctionCUDABase* base = loss_function;
      if (loss_function_ptrs_.find(base) == loss_function_ptrs_.end())
        loss_function_ptrs_[base] = std::unique_ptr<LossFunctionCUDABase>(base);
    }
  }

  template <typename CostFunctor, int kNumResiduals, int... Ns, typename... Ts,
            typename LossFunctionCUDA>
  ResidualBlockId InternalAddResidualBlock(CostFunction* cost_function,
                                           LossFunctionCUDA* loss_function, double* x0,
                                           Ts*... xs) {
    static_assert(sizeof...(Ts) + 1 == sizeof...(Ns),
                  "one parameter block pointer per block size is required");
    static TrivialLossCUDA trivial;
    const std::array<double*, sizeof...(Ts) + 1> parameter_blocks{{x0, xs...}};
    auto* autodiff_cost_function =
        dynamic_cast<AutoDiffCostFunction<CostFunctor, kNumResiduals, Ns...>*>(cost_function);
    if (autodiff_cost_function == nullptr) {
      // The reference dereferences the failed cast (problem_cuda.h:443-453); report instead.
      std::fprintf(stderr,
                   "ProblemCUDA::AddResidualBlock: the cost function is not an "
                   "AutoDiffCostFunction<CostFunctor, kNumResiduals, Ns...> of the given "
                   "template arguments; only those can be evaluated with CUDA.\n");
      return nullptr;
    }
    const int type = RegisterType<CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>();
    const void* loss = loss_function ? static_cast<const void*>(loss_function)
                                     : static_cast<const void*>(&trivial);
    TakeLossOwnership(loss_function);
    return problem_->mutable_impl()->AddResidualBlock(
        type, cost_function, &autodiff_cost_function->functor(), loss, parameter_blocks.data());
  }

  std::unique_ptr<Problem> problem_;
  std::unique_ptr<internal::RegisteredCUDAEvaluators> registered_cuda_evaluators_;
  // Loss functions this object owns (problem_cuda.h:481-485).
  std::unordered_map<LossFunctionCUDABase*, std::unique_ptr<LossFunctionCUDABase>>
      loss_function_ptrs_;
};

// Same helper as the reference's (include/ceres/problem_cuda.h:490-502): solve with the
// CUDA evaluator of this problem.
inline void Solve(const Solver::Options& options, ProblemCUDA* problem_cuda,
                  Solver::Summary* summary) {
  Solver solver;
  Solver::Options options_with_cuda = options;
  options_with_cuda.use_cuda_for_evaluator = true;
  options_with_cuda.registered_cuda_evaluators =
      problem_cuda->mutable_registered_cuda_evaluators();
  solver.Solve(options_with_cuda, problem_cuda->mutable_problem(), summary);
}

}  // namespace ceres

#endif  // CERES_B200_PROBLEM_CUDA_H_
