// The Evaluator interface the TRUST_REGION minimizer calls, and its CUDA
// implementation.
//
// Same interface as the reference's ceres::internal::Evaluator
// (internal/ceres/evaluator.h:60-150).  Evaluator::Create supplies the dispatch
// that is missing from the reference snapshot ("snapshot gap";
// internal/ceres/evaluator.cc:53-95 vs. evaluator_cuda_test.cu.cc:396-400,451-459):
// Schur-type / CGNR / SPARSE_NORMAL_CHOLESKY solvers get a BlockSparseMatrix
// Jacobian, except sparse_linear_algebra_library_type == CUDA_SPARSE which gets a
// CompressedRowSparseMatrix.  There is one implementation, ProgramEvaluatorCUDA;
// a process without a CUDA device gets a null evaluator and an error string (no
// CPU fallback).
#ifndef CERES_B200_INTERNAL_EVALUATOR_H_
#define CERES_B200_INTERNAL_EVALUATOR_H_

#include "ceres/evaluation_callback.h"
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ceres/internal/program.h"
#include "ceres/internal/sparse_matrix.h"
#include "ceres/types.h"

namespace ceres {
namespace internal {

// Owned by ProblemCUDA and handed to the evaluator through Evaluator::Options,
// like the reference's RegisteredCUDAEvaluators
// (include/ceres/internal/registered_cuda_evaluators.h:64-121).  It names the
// residual-block type registry (launch thunks + per-type tables) of one problem.
class RegisteredCUDAEvaluators {
 public:
  explicit RegisteredCUDAEvaluators(ProblemImpl* problem) : problem_(problem) {}
  ProblemImpl* problem() const { return problem_; }
  int NumRegisteredTypes() const { return static_cast<int>(problem_->types().size()); }

 private:
  ProblemImpl* problem_;
};

struct CallStatistics {
  double time = 0.0;  // seconds
  int calls = 0;
};

// The per-residual-block layouts of the fork's writers
// (block_jacobian_writer.cc:62-160, compressed_row_jacobian_writer.cc:240-300)
// plus the residual layout (program_evaluator_cuda.h:159-170).
struct JacobianLayout {
  int jacobian_format = CB200_JACOBIAN_BLOCK_SPARSE;
  int num_eliminate_blocks = 0;
  std::vector<int32_t> residual_layout;
  std::vector<int32_t> jacobian_per_residual_layout;
  std::vector<int32_t> jacobian_per_residual_offsets;
  // BlockSparseMatrix only: start of each active cell, in argument order
  // (BlockJacobianWriter::jacobian_layout_).
  std::vector<int32_t> cell_positions;
  int64_t num_jacobian_values = 0;
  int num_residuals = 0;
};

// Pure host code; no device needed.
void BuildJacobianLayout(const Program& program, int jacobian_format, int num_eliminate_blocks,
                         JacobianLayout* layout);
std::unique_ptr<SparseMatrix> CreateJacobianFromLayout(const Program& program,
                                                       const JacobianLayout& layout);

class Evaluator {
 public:
  virtual ~Evaluator() {}

  struct Options {
    int num_threads = 1;
    int num_eliminate_blocks = -1;
    LinearSolverType linear_solver_type = DENSE_QR;
    SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = NO_SPARSE;
    bool dynamic_sparsity = false;
    bool use_cuda = true;
    RegisteredCUDAEvaluators* registered_cuda_evaluators = nullptr;
    EvaluationCallback* evaluation_callback = nullptr;  // internal/ceres/evaluator.h:76
    // Extensions (not in the reference): device ordinal and residual-block sharding.
    int device = 0;
    int shard_rank = 0;
    int shard_world_size = 1;
    const void* nccl_unique_id = nullptr;  // 128 bytes, enables the all-reduce
    // CreateJacobian() returns a DeviceResidentJacobian: values stay in HBM and the linear
    // algebra runs there (block-sparse value layout, whatever the solver type).
    bool jacobian_on_device = false;
  };

  static std::unique_ptr<Evaluator> Create(const Options& options, Program* program,
                                           std::string* error);

  virtual std::unique_ptr<SparseMatrix> CreateJacobian() const = 0;

  struct EvaluateOptions {
    bool apply_loss_function = true;
    bool new_evaluation_point = true;
  };

  // residuals, gradient and jacobian may each be null; cost may not.
  virtual bool Evaluate(const EvaluateOptions& evaluate_options, const double* state,
                        double* cost, double* residuals, double* gradient,
                        SparseMatrix* jacobian) = 0;
  bool Evaluate(const double* state, double* cost, double* residuals, double* gradient,
                SparseMatrix* jacobian) {
    return Evaluate(EvaluateOptions(), state, cost, residuals, gradient, jacobian);
  }

  virtual bool Plus(const double* state, const double* delta,
                    double* state_plus_delta) const = 0;
  virtual int NumParameters() const = 0;
  virtual int NumEffectiveParameters() const = 0;
  virtual int NumResiduals() const = 0;
  virtual std::map<std::string, CallStatistics> Statistics() const {
    return std::map<std::string, CallStatistics>();
  }
  // Extension: the C-ABI engine behind this evaluator (device pointers, timing).
  virtual cb200_engine* engine() const { return nullptr; }
  virtual const JacobianLayout* layout() const { return nullptr; }
};

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_EVALUATOR_H_
