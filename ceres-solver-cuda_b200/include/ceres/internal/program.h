// Host-side problem graph and (reduced) program for the evaluation path.
//
// The evaluation engine needs exactly what the reference's Program hands its
// evaluators (internal/ceres/program.h, program.cc:80-103,152-177,306-430;
// parameter_block.h; residual_block.h): parameter blocks with size / tangent size /
// constness / manifold and their index, state_offset and delta_offset; residual
// blocks in program order with their parameter blocks; and the constant blocks
// that were removed from the program but are still read by residual blocks.
//
// Storage is structure-of-arrays per residual-block TYPE (one store per
// <CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>): the reference keeps
// three heap objects per residual block (problem_cuda.h:452-473), which is what
// makes its preprocessor take 47 s on the 29 M-block BAL problem (README.md:186).
#ifndef CERES_B200_INTERNAL_PROGRAM_H_
#define CERES_B200_INTERNAL_PROGRAM_H_

#include "ceres/evaluation_callback.h"
#include <cstdint>
#include <memory>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <vector>

#include "ceres/cost_function.h"
#include "ceres/loss_function_cuda.h"
#include "ceres/manifold.h"
#include "ceres/types.h"
#include "ceres_b200.h"

namespace ceres {
namespace internal {

class ResidualBlock;  // opaque: ResidualBlockId = ResidualBlock*

struct ParameterBlock {
  double* user_state = nullptr;
  int size = 0;
  bool is_set_constant = false;
  Manifold* manifold = nullptr;
  std::unique_ptr<double[]> lower_bounds, upper_bounds;
  int id = -1;  // position in ProblemImpl::parameter_blocks()
  // Program bookkeeping (program.cc:152-177); rewritten by every Program.
  int index = -1;
  int state_offset = -1;
  int delta_offset = -1;

  int TangentSize() const { return manifold ? manifold->TangentSize() : size; }
  // parameter_block.h:120
  bool IsConstant() const { return is_set_constant || TangentSize() == 0; }
};

// All residual blocks of one <CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>.
struct ResidualTypeStore {
  cb200_residual_type desc{};
  std::type_index key = std::type_index(typeid(void));
  std::vector<int32_t> parameter_blocks;  // [n][num_parameter_blocks] ParameterBlock::id
  std::vector<char> functors;             // n * desc.functor_size bytes
  std::vector<const void*> loss_objects;  // one representative host object per distinct loss
  std::vector<char> loss_table;           // their bytes, desc.loss_size each
  std::unordered_map<std::string, int> loss_by_bytes;  // content -> index into loss_table
  std::vector<int32_t> loss_index;        // [n] index into loss_table
  std::vector<CostFunction*> cost_functions;  // [n] (null for bulk-added blocks)
  std::vector<int32_t> residual_block_id;     // [n] global id
  bool has_loss = true;  // false for the nullptr-loss overload: apply_loss is a no-op
  // Host evaluation thunks (type-erased, instantiated with the type).
  void (*host_loss)(const void* loss, double s, double rho[3]) = nullptr;
  bool (*host_functor)(const void* functor, double const* const* parameters, double* residuals,
                       double** jacobians) = nullptr;
  int32_t size() const { return static_cast<int32_t>(residual_block_id.size()); }
};

struct ResidualBlockRef {
  int32_t type;   // index into ProblemImpl::types()
  int32_t local;  // index inside the type store
};

struct ProblemOptions {
  Ownership cost_function_ownership = TAKE_OWNERSHIP;
  Ownership loss_function_ownership = TAKE_OWNERSHIP;
  Ownership manifold_ownership = TAKE_OWNERSHIP;
  bool enable_fast_removal = false;
  bool disable_all_safety_checks = false;
  // Notified before every evaluation (include/ceres/problem.h:166-183); not owned.
  EvaluationCallback* evaluation_callback = nullptr;
};

class ProblemImpl {
 public:
  explicit ProblemImpl(const ProblemOptions& options = ProblemOptions());
  ~ProblemImpl();
  ProblemImpl(const ProblemImpl&) = delete;
  void operator=(const ProblemImpl&) = delete;

  ParameterBlock* AddParameterBlock(double* values, int size, Manifold* manifold = nullptr);
  ParameterBlock* FindParameterBlock(const double* values) const;
  // Aborts with the reference's message when the block was never added.
  ParameterBlock* FindParameterBlockOrDie(const double* values, const char* what) const;
  void SetManifold(double* values, Manifold* manifold);
  void SetParameterBlockConstant(const double* values);
  void SetParameterBlockVariable(double* values);
  void SetParameterLowerBound(double* values, int index, double bound);
  void SetParameterUpperBound(double* values, int index, double bound);
  double GetParameterLowerBound(const double* values, int index) const;
  double GetParameterUpperBound(const double* values, int index) const;

  // Returns the type store for `key`, creating it from `desc` on first use.
  int FindOrAddType(std::type_index key, const cb200_residual_type& desc);
  // Appends one residual block to a type store.  `functor` / `loss` point at host
  // objects whose bytes are copied (desc.functor_size / desc.loss_size).
  ResidualBlock* AddResidualBlock(int type, CostFunction* cost_function, const void* functor,
                                  const void* loss, double* const* parameter_blocks);

  const std::vector<std::unique_ptr<ParameterBlock>>& parameter_blocks() const { return pbs_; }
  const std::vector<ResidualBlockRef>& residual_blocks() const { return rbs_; }
  std::vector<ResidualTypeStore>& types() { return types_; }
  const std::vector<ResidualTypeStore>& types() const { return types_; }
  const ProblemOptions& options() const { return options_; }

  int NumParameterBlocks() const { return static_cast<int>(pbs_.size()); }
  int NumParameters() const;
  int NumResidualBlocks() const { return static_cast<int>(rbs_.size()); }
  int NumResiduals() const;

  static int32_t IdOf(const ResidualBlock* rb) {
    return static_cast<int32_t>(reinterpret_cast<intptr_t>(rb)) - 1;
  }
  static ResidualBlock* HandleOf(int32_t id) {
    return reinterpret_cast<ResidualBlock*>(static_cast<intptr_t>(id) + 1);
  }

  // Evaluates one residual block on the host at the user state (fixed costs of the
  // reduced program, Problem::EvaluateResidualBlock): residual_block.cc:68-204 with the
  // manifold projection and the loss correction; jacobians[j] (may be null) is
  // num_residuals x tangent size of argument j.
  bool EvaluateResidualBlockOnHost(int32_t id, bool apply_loss_function, double* cost,
                                   double* residuals, double** jacobians) const;
  void GetParameterBlocksForResidualBlock(int32_t id, std::vector<double*>* out) const;
  void GetResidualBlocksForParameterBlock(const double* values, std::vector<int32_t>* out) const;
  const CostFunction* CostFunctionOf(int32_t id) const;

 private:
  ProblemOptions options_;
  std::vector<std::unique_ptr<ParameterBlock>> pbs_;
  std::unordered_map<const double*, ParameterBlock*> pb_map_;
  std::vector<ResidualBlockRef> rbs_;
  std::vector<ResidualTypeStore> types_;
  std::unordered_map<std::type_index, int> type_map_;
  std::vector<Manifold*> manifolds_to_delete_;
  std::vector<CostFunction*> cost_functions_to_delete_;
};

// The program an evaluator works on.
class Program {
 public:
  explicit Program(ProblemImpl* problem);
  // Explicit block lists (Problem::Evaluate); parameter blocks that residual blocks use
  // but that are not listed go to constant_parameter_blocks().
  Program(ProblemImpl* problem, std::vector<ParameterBlock*> parameter_blocks,
          std::vector<int32_t> residual_blocks);

  // Program::CreateReducedProgram (program.cc:306-322): drops constant parameter
  // blocks and residual blocks that depend only on them (their cost goes to
  // *fixed_cost), then SetParameterOffsetsAndIndex.
  std::unique_ptr<Program> CreateReducedProgram(std::vector<double*>* removed_parameter_blocks,
                                                double* fixed_cost, std::string* error) const;
  void SetParameterOffsetsAndIndex();
  // ApplyOrdering-style reordering: parameter blocks sorted (stably) by group id.
  void ReorderParameterBlocksByGroup(const std::unordered_map<const double*, int>& group);
  // reorder_program.cc:254-335
  bool LexicographicallyOrderResidualBlocks(int size_of_first_elimination_group);

  void ParameterBlocksToStateVector(double* state) const;
  void StateVectorToParameterBlocks(const double* state) const;  // writes user state
  void ConstantParameterBlocksToStateVector(double* state) const;
  // program.cc:121-150 (a ParallelFor over the parameter blocks)
  bool Plus(const double* state, const double* delta, double* state_plus_delta,
            int num_threads = 1) const;

  int NumParameterBlocks() const { return static_cast<int>(parameter_blocks_.size()); }
  int NumResidualBlocks() const { return static_cast<int>(residual_blocks_.size()); }
  int NumParameters() const;
  int NumEffectiveParameters() const;
  int NumResiduals() const;
  int NumConstantParameters() const;

  const std::vector<ParameterBlock*>& parameter_blocks() const { return parameter_blocks_; }
  const std::vector<ParameterBlock*>& constant_parameter_blocks() const {
    return constant_parameter_blocks_;
  }
  // Global residual block ids in program order; POSITION in this vector is what
  // indexes residual_layout and the Jacobian layouts.
  const std::vector<int32_t>& residual_blocks() const { return residual_blocks_; }
  ProblemImpl* problem() const { return problem_; }

  // A parameter block takes part in the Jacobian iff it is in parameter_blocks().
  bool IsActive(const ParameterBlock* pb) const {
    return pb->index >= 0 && pb->index < NumParameterBlocks() &&
           parameter_blocks_[pb->index] == pb;
  }

 private:
  ProblemImpl* problem_;
  std::vector<ParameterBlock*> parameter_blocks_;
  std::vector<ParameterBlock*> constant_parameter_blocks_;
  std::vector<int32_t> residual_blocks_;
};

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_PROGRAM_H_
