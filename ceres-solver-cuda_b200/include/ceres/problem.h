// ceres::Problem — the parameter-block side of the problem description
// (reference: include/ceres/problem.h; only the members ProblemCUDA forwards to,
// include/ceres/problem_cuda.h:166-401).  Residual blocks are added through
// ProblemCUDA::AddResidualBlock<...>, which needs the functor type.
#ifndef CERES_B200_PROBLEM_H_
#define CERES_B200_PROBLEM_H_

#include <memory>
#include <vector>

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
    return impl_->FindParameterBlock(values)->IsConstant();
  }
  void SetManifold(double* values, Manifold* manifold) { impl_->SetManifold(values, manifold); }
  const Manifold* GetManifold(const double* values) const {
    return impl_->FindParameterBlock(values)->manifold;
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
    return impl_->FindParameterBlock(values)->size;
  }
  int ParameterBlockTangentSize(const double* values) const {
    return impl_->FindParameterBlock(values)->TangentSize();
  }
  bool HasParameterBlock(const double* values) const {
    return impl_->FindParameterBlock(values) != nullptr;
  }
  void GetParameterBlocks(std::vector<double*>* parameter_blocks) const {
    parameter_blocks->clear();
    for (const auto& pb : impl_->parameter_blocks()) parameter_blocks->push_back(pb->user_state);
  }
  const Options& options() const { return impl_->options(); }
  internal::ProblemImpl* mutable_impl() { return impl_.get(); }

 private:
  std::unique_ptr<internal::ProblemImpl> impl_;
};

}  // namespace ceres

#endif  // CERES_B200_PROBLEM_H_
