// ceres::Problem — the parameter-block side of the problem description
// (reference: include/ceres/problem.h; only the members ProblemCUDA forwards to,
// include/ceres/problem_cuda.h:166-401).  Residual blocks are added through
// ProblemCUDA::AddResidualBlock<...>, which needs the functor type.
#ifndef CERES_B200_PROBLEM_H_
#define CERES_B200_PROBLEM_H_

#include <memory>
#include <vector>

#include "ceres/crs_matrix.h"
#include "ceres/internal/program.h"
#include "ceres/manifold.h"
#include "ceres/types.h"

namespace ceres {

using ResidualBlockId = internal::ResidualBlock*;

class Problem {
 public:
  using Options = internal::ProblemOptions;

  Problem() : impl_(new internal::ProblemImpl) {}
  explicit Problem(const Options& options) : impl_(new internal::ProblemImpl(options)) {}

  void AddParameterBlock(double* values, int size) { impl_->AddParameterBlock(values, size); }
  void AddParameterBlock(double* values, int size, Manifold* manifold) {
    impl_->AddParameterBlock(values, size, manifold);
  }
  void SetParameterBlockConstant(const double* values) {
    impl_->SetParameterBlockConstant(values);
  }
  void SetParameterBlockVariable(double* values) { impl_->SetParameterBlockVariable(values); }
  bool IsParameterBlockConstant(const double* values) const {
    return impl_->FindParameterBlockOrDie(values, "it can be queried if it is constant")->IsConstant();
  }
  void SetManifold(double* values, Manifold* manifold) { impl_->SetManifold(values, manifold); }
  const Manifold* GetManifold(const double* values) const {
    return impl_->FindParameterBlockOrDie(values, "you can get its manifold")->manifold;
  }
  bool HasManifold(const double* values) const { return GetManifold(values) != nullptr; }
  void SetParameterLowerBound(double* values, int index, double bound) {
    impl_->SetParameterLowerBound(values, index, bound);
  }
  void SetParameterUpperBound(double* values, int index, double bound) {
    impl_->SetParameterUpperBound(values, index, bound);
  }
  double GetParameterLowerBound(const double* values, int index) const {
    return impl_->GetParameterLowerBound(values, index);
  }
  double GetParameterUpperBound(const double* values, int index) const {
    return impl_->GetParameterUpperBound(values, index);
  }
  int NumParameterBlocks() const { return impl_->NumParameterBlocks(); }
  int NumParameters() const { return impl_->NumParameters(); }
  int NumResidualBlocks() const { return impl_->NumResidualBlocks(); }
  int NumResiduals() const { return impl_->NumResiduals(); }
  int ParameterBlockSize(const double* values) const {
    return impl_->FindParameterBlockOrDie(values, "you can get its size")->size;
  }
  int ParameterBlockTangentSize(const double* values) const {
    return impl_->FindParameterBlockOrDie(values, "you can get its tangent size")->TangentSize();
  }
  bool HasParameterBlock(const double* values) const {
    return impl_->FindParameterBlock(values) != nullptr;
  }
  void GetParameterBlocks(std::vector<double*>* parameter_blocks) const {
    parameter_blocks->clear();
    for (const auto& pb : impl_->parameter_blocks()) parameter_blocks->push_back(pb->user_state);
  }
  // ---- queries (include/ceres/problem.h:385-420)
  void GetResidualBlocks(std::vector<ResidualBlockId>* residual_blocks) const {
    residual_blocks->clear();
    for (int i = 0; i < impl_->NumResidualBlocks(); ++i)
      residual_blocks->push_back(internal::ProblemImpl::HandleOf(i));
  }
  void GetParameterBlocksForResidualBlock(ResidualBlockId residual_block,
                                          std::vector<double*>* parameter_blocks) const {
    impl_->GetParameterBlocksForResidualBlock(internal::ProblemImpl::IdOf(residual_block),
                                              parameter_blocks);
  }
  // Null for residual blocks added in bulk (ProblemCUDA::AddResidualBlocks).
  const CostFunction* GetCostFunctionForResidualBlock(ResidualBlockId residual_block) const {
    return impl_->CostFunctionOf(internal::ProblemImpl::IdOf(residual_block));
  }
  void GetResidualBlocksForParameterBlock(const double* values,
                                          std::vector<ResidualBlockId>* residual_blocks) const {
    std::vector<int32_t> ids;
    impl_->GetResidualBlocksForParameterBlock(values, &ids);
    residual_blocks->clear();
    for (int32_t id : ids) residual_blocks->push_back(internal::ProblemImpl::HandleOf(id));
  }

  // include/ceres/problem.h:426-470
  struct EvaluateOptions {
    // Columns of the gradient / Jacobian, in this order; empty = every parameter block in
    // the order they were added.  Blocks left out are held constant for the call.
    std::vector<double*> parameter_blocks;
    // Rows; empty = every residual block in the order they were added.
    std::vector<ResidualBlockId> residual_blocks;
    bool apply_loss_function = true;
    int num_threads = 1;
    int cuda_device = 0;  // extension: the device that evaluates
  };
  // Evaluates the problem at the current values of the user's parameter blocks
  // (internal/ceres/problem_impl.cc:599-760) — here on the CUDA evaluator, with a
  // CompressedRowSparseMatrix Jacobian like the reference.  Any output may be null.
  // Constant parameter blocks keep their columns (all zero), as in the reference.
  bool Evaluate(const EvaluateOptions& options, double* cost, std::vector<double>* residuals,
                std::vector<double>* gradient, CRSMatrix* jacobian);

  // One residual block on the host, at the current user state (problem_impl.cc:762-810):
  // residuals and Jacobians with the manifolds and (optionally) the loss function applied;
  // jacobians[i] is num_residuals x tangent size, row-major, and may be null.
  bool EvaluateResidualBlock(ResidualBlockId residual_block, bool apply_loss_function,
                             double* cost, double* residuals, double** jacobians) const {
    return impl_->EvaluateResidualBlockOnHost(internal::ProblemImpl::IdOf(residual_block),
                                              apply_loss_function, cost, residuals, jacobians);
  }
  bool EvaluateResidualBlockAssumingParametersUnchanged(ResidualBlockId residual_block,
                                                        bool apply_loss_function, double* cost,
                                                        double* residuals,
                                                        double** jacobians) const {
    return EvaluateResidualBlock(residual_block, apply_loss_function, cost, residuals, jacobians);
  }

  const Options& options() const { return impl_->options(); }
  internal::ProblemImpl* mutable_impl() { return impl_.get(); }

 private:
  std::unique_ptr<internal::ProblemImpl> impl_;
};

}  // namespace ceres

#endif  // CERES_B200_PROBLEM_H_
