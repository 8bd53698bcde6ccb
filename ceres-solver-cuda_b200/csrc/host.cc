// Host side of the evaluation path: problem graph, reduced program, Jacobian
// layouts and ProgramEvaluatorCUDA.  Thin by design — it prepares arrays and calls
// the C ABI (include/ceres_b200.h); all arithmetic runs in the CUDA kernels.
//
// Reference counterparts: internal/ceres/problem_impl.cc, program.cc,
// reorder_program.cc:254-335, block_jacobian_writer.cc, compressed_row_jacobian_writer.cc,
// program_evaluator_cuda.h:65-183, registered_cuda_evaluators.cc:123-280.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>

#include "ceres/internal/evaluator.h"
#include "ceres/internal/parallel_for.h"
#include "ceres/internal/program.h"
#include "ceres/internal/sparse_matrix.h"
#include "ceres/problem.h"
#include <nvtx3/nvToolsExt.h>

namespace ceres {
namespace internal {

// ------------------------------------------------------------------ ProblemImpl
ProblemImpl::ProblemImpl(const ProblemOptions& options) : options_(options) {}

ProblemImpl::~ProblemImpl() {
  if (options_.manifold_ownership == TAKE_OWNERSHIP) {
    std::sort(manifolds_to_delete_.begin(), manifolds_to_delete_.end());
    manifolds_to_delete_.erase(
        std::unique(manifolds_to_delete_.begin(), manifolds_to_delete_.end()),
        manifolds_to_delete_.end());
    for (Manifold* m : manifolds_to_delete_) delete m;
  }
  if (options_.cost_function_ownership == TAKE_OWNERSHIP) {
    for (auto& t : types_)
      for (CostFunction* c : t.cost_functions)
        if (c) cost_functions_to_delete_.push_back(c);
    std::sort(cost_functions_to_delete_.begin(), cost_functions_to_delete_.end());
    cost_functions_to_delete_.erase(
        std::unique(cost_functions_to_delete_.begin(), cost_functions_to_delete_.end()),
        cost_functions_to_delete_.end());
    for (CostFunction* c : cost_functions_to_delete_) delete c;
  }
}

ParameterBlock* ProblemImpl::FindParameterBlock(const double* values) const {
  auto it = pb_map_.find(values);
  return it == pb_map_.end() ? nullptr : it->second;
}

ParameterBlock* ProblemImpl::AddParameterBlock(double* values, int size, Manifold* manifold) {
  ParameterBlock* pb = FindParameterBlock(values);
  if (pb != nullptr) {
    if (pb->size != size) {
      std::fprintf(stderr,
                   "Tried adding a parameter block with the same double pointer, %p, twice, "
                   "but with different block sizes. Original size was %d but new size is %d\n",
                   static_cast<void*>(values), pb->size, size);
      std::abort();  // the reference CHECK-fails here too (problem_impl.cc:139-147)
    }
    if (manifold != nullptr) SetManifold(values, manifold);
    return pb;
  }
  pbs_.emplace_back(new ParameterBlock);
  pb = pbs_.back().get();
  pb->user_state = values;
  pb->size = size;
  pb->id = static_cast<int>(pbs_.size()) - 1;
  pb_map_[values] = pb;
  if (manifold != nullptr) SetManifold(values, manifold);
  return pb;
}

void ProblemImpl::SetManifold(double* values, Manifold* manifold) {
  ParameterBlock* pb = FindParameterBlock(values);
  if (pb == nullptr) {
    std::fprintf(stderr, "SetManifold: parameter block %p not found\n", static_cast<void*>(values));
    std::abort();
  }
  if (manifold != nullptr && manifold->AmbientSize() != pb->size) {
    std::fprintf(stderr, "SetManifold: ambient size %d != parameter block size %d\n",
                 manifold->AmbientSize(), pb->size);
    std::abort();
  }
  pb->manifold = manifold;
  if (manifold != nullptr) manifolds_to_delete_.push_back(manifold);
}

// The reference LOG(FATAL)s with this message when the block is unknown
// (problem_impl.cc:455-515).
ParameterBlock* ProblemImpl::FindParameterBlockOrDie(const double* values, const char* what) const {
  ParameterBlock* pb = FindParameterBlock(values);
  if (pb == nullptr) {
    std::fprintf(stderr,
                 "Parameter block not found: %p. You must add the parameter block to the "
                 "problem before %s.\n",
                 static_cast<const void*>(values), what);
    std::abort();
  }
  return pb;
}

void ProblemImpl::SetParameterBlockConstant(const double* values) {
  FindParameterBlockOrDie(values, "it can be set constant")->is_set_constant = true;
}
void ProblemImpl::SetParameterBlockVariable(double* values) {
  FindParameterBlockOrDie(values, "it can be set varying")->is_set_constant = false;
}

void ProblemImpl::SetParameterLowerBound(double* values, int index, double bound) {
  ParameterBlock* pb = FindParameterBlockOrDie(values, "you can set a lower bound on one of its components");
  if (!pb->lower_bounds) {
    pb->lower_bounds.reset(new double[pb->size]);
    std::fill_n(pb->lower_bounds.get(), pb->size, -std::numeric_limits<double>::max());
  }
  pb->lower_bounds[index] = bound;
}
void ProblemImpl::SetParameterUpperBound(double* values, int index, double bound) {
  ParameterBlock* pb = FindParameterBlockOrDie(values, "you can set an upper bound on one of its components");
  if (!pb->upper_bounds) {
    pb->upper_bounds.reset(new double[pb->size]);
    std::fill_n(pb->upper_bounds.get(), pb->size, std::numeric_limits<double>::max());
  }
  pb->upper_bounds[index] = bound;
}
double ProblemImpl::GetParameterLowerBound(const double* values, int index) const {
  const ParameterBlock* pb = FindParameterBlockOrDie(values, "you can get its lower bound");
  return pb->lower_bounds ? pb->lower_bounds[index] : -std::numeric_limits<double>::max();
}
double ProblemImpl::GetParameterUpperBound(const double* values, int index) const {
  const ParameterBlock* pb = FindParameterBlockOrDie(values, "you can get its upper bound");
  return pb->upper_bounds ? pb->upper_bounds[index] : std::numeric_limits<double>::max();
}

int ProblemImpl::FindOrAddType(std::type_index key, const cb200_residual_type& desc) {
  auto it = type_map_.find(key);
  if (it != type_map_.end()) return it->second;
  types_.emplace_back();
  types_.back().desc = desc;
  types_.back().key = key;
  const int id = static_cast<int>(types_.size()) - 1;
  type_map_.emplace(key, id);
  return id;
}

ResidualBlock* ProblemImpl::AddResidualBlock(int type, CostFunction* cost_function,
                                             const void* functor, const void* loss,
                                             double* const* parameter_blocks) {
  ResidualTypeStore& t = types_[type];
  const int nb = t.desc.num_parameter_blocks;
  for (int j = 0; j < nb; ++j) {
    // Duplicate parameter blocks in one residual block are an error in Ceres
    // (problem_impl.cc:264-283).
    for (int k = 0; k < j; ++k) {
      if (parameter_blocks[j] == parameter_blocks[k]) {
        std::fprintf(stderr, "Duplicate parameter blocks in a residual block are not allowed.\n");
        std::abort();
      }
    }
    ParameterBlock* pb = AddParameterBlock(parameter_blocks[j], t.desc.parameter_block_sizes[j]);
    t.parameter_blocks.push_back(pb->id);
  }
  const char* f = static_cast<const char*>(functor);
  t.functors.insert(t.functors.end(), f, f + t.desc.functor_size);
  // Deduplicate loss objects by CONTENT: user code typically does
  // `new HuberLossCUDA(1.0)` per residual block (examples/bundle_adjuster.cu.cc:323-324)
  // and the reference then copies one loss object per block to the device
  // (autodiff_residual_block_cuda_evaluator.h:96-133).  Loss functors are plain data
  // evaluated by value on the device, so equal bytes mean the same loss.
  // (Same pointer as the last block is only a shortcut when its bytes still match what was
  // captured: a stack loss may have been modified, a freed one reallocated at the address.)
  int loss_id = -1;
  if (!t.loss_objects.empty() && t.loss_objects.back() == loss &&
      std::memcmp(loss, t.loss_table.data() + (t.loss_objects.size() - 1) * t.desc.loss_size,
                  t.desc.loss_size) == 0) {
    loss_id = static_cast<int>(t.loss_objects.size()) - 1;
  } else {
    const std::string bytes(static_cast<const char*>(loss), t.desc.loss_size);
    auto it = t.loss_by_bytes.find(bytes);
    if (it != t.loss_by_bytes.end()) {
      loss_id = it->second;
    } else {
      loss_id = static_cast<int>(t.loss_objects.size());
      t.loss_objects.push_back(loss);
      t.loss_table.insert(t.loss_table.end(), bytes.begin(), bytes.end());
      t.loss_by_bytes.emplace(bytes, loss_id);
    }
  }
  t.loss_index.push_back(loss_id);
  t.cost_functions.push_back(cost_function);
  const int32_t id = static_cast<int32_t>(rbs_.size());
  t.residual_block_id.push_back(id);
  rbs_.push_back(ResidualBlockRef{type, t.size() - 1});
  return HandleOf(id);
}

int ProblemImpl::NumParameters() const {
  int n = 0;
  for (const auto& pb : pbs_) n += pb->size;
  return n;
}
int ProblemImpl::NumResiduals() const {
  int n = 0;
  for (const auto& t : types_) n += t.size() * t.desc.num_residuals;
  return n;
}

// residual_block.cc:68-204 on the host, for residual blocks outside the program.
bool ProblemImpl::EvaluateResidualBlockOnHost(int32_t id, bool apply_loss_function, double* cost,
                                              double* residuals, double** jacobians) const {
  const ResidualBlockRef ref = rbs_[id];
  const ResidualTypeStore& t = types_[ref.type];
  const int nb = t.desc.num_parameter_blocks, kres = t.desc.num_residuals;
  const double* params[CB200_MAX_PARAMETER_BLOCKS];
  const ParameterBlock* blocks[CB200_MAX_PARAMETER_BLOCKS];
  for (int j = 0; j < nb; ++j) {
    blocks[j] = pbs_[t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j]].get();
    params[j] = blocks[j]->user_state;
    if (jacobians && jacobians[j] && blocks[j]->IsConstant()) {
      // problem_impl.cc:777-782
      std::fprintf(stderr, "Jacobian requested for parameter block : %d. But the parameter block "
                           "is marked constant.\n", j);
      return false;
    }
  }
  std::vector<double> scratch(kres);
  double* r = residuals ? residuals : scratch.data();
  const void* functor = t.functors.data() + static_cast<size_t>(ref.local) * t.desc.functor_size;
  // ambient Jacobians of the requested blocks (residual_block.cc:86-98)
  std::vector<double> ambient_storage;
  double* ambient[CB200_MAX_PARAMETER_BLOCKS] = {};
  if (jacobians) {
    size_t total = 0;
    for (int j = 0; j < nb; ++j)
      if (jacobians[j]) total += static_cast<size_t>(kres) * blocks[j]->size;
    ambient_storage.assign(total, 0.0);
    size_t cursor = 0;
    for (int j = 0; j < nb; ++j)
      if (jacobians[j]) {
        ambient[j] = ambient_storage.data() + cursor;
        cursor += static_cast<size_t>(kres) * blocks[j]->size;
      }
  }
  if (!t.host_functor(functor, params, r, jacobians ? ambient : nullptr)) return false;
  double s = 0.0;
  for (int i = 0; i < kres; ++i) {
    if (!std::isfinite(r[i]) || r[i] == 1e302) return false;
    s += r[i] * r[i];
  }
  // J <- J * PlusJacobian (residual_block.cc:137-155)
  if (jacobians) {
    for (int j = 0; j < nb; ++j) {
      if (!jacobians[j]) continue;
      const int size = blocks[j]->size, tangent = blocks[j]->TangentSize();
      for (size_t i = 0; i < static_cast<size_t>(kres) * size; ++i)
        if (!std::isfinite(ambient[j][i])) return false;
      if (blocks[j]->manifold) {
        std::vector<double> plus(static_cast<size_t>(size) * tangent);
        if (!blocks[j]->manifold->PlusJacobian(params[j], plus.data())) return false;
        for (int row = 0; row < kres; ++row)
          for (int c = 0; c < tangent; ++c) {
            double acc = 0.0;
            for (int k = 0; k < size; ++k) acc += ambient[j][row * size + k] * plus[k * tangent + c];
            jacobians[j][row * tangent + c] = acc;
          }
      } else {
        std::memcpy(jacobians[j], ambient[j], sizeof(double) * kres * size);
      }
    }
  }
  if (!apply_loss_function) {
    *cost = 0.5 * s;
    return true;
  }
  double rho[3];
  t.host_loss(t.loss_table.data() + static_cast<size_t>(t.loss_index[ref.local]) * t.desc.loss_size,
              s, rho);
  *cost = 0.5 * rho[0];
  if (!residuals && !jacobians) return true;
  // Corrector (corrector.cc:88-160)
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling = sqrt_rho1, alpha_sq_norm = 0.0;
  if (!(s == 0.0 || rho[2] <= 0.0)) {
    const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1.0 - alpha);
    alpha_sq_norm = alpha / s;
  }
  if (jacobians) {
    for (int j = 0; j < nb; ++j) {
      if (!jacobians[j]) continue;
      const int tangent = blocks[j]->TangentSize();
      for (int c = 0; c < tangent; ++c) {
        if (alpha_sq_norm == 0.0) {
          for (int row = 0; row < kres; ++row) jacobians[j][row * tangent + c] *= sqrt_rho1;
          continue;
        }
        double r_transpose_j = 0.0;
        for (int row = 0; row < kres; ++row) r_transpose_j += jacobians[j][row * tangent + c] * r[row];
        for (int row = 0; row < kres; ++row)
          jacobians[j][row * tangent + c] =
              sqrt_rho1 * (jacobians[j][row * tangent + c] - alpha_sq_norm * r[row] * r_transpose_j);
      }
    }
  }
  if (residuals)
    for (int i = 0; i < kres; ++i) residuals[i] *= residual_scaling;
  return true;
}

void ProblemImpl::GetParameterBlocksForResidualBlock(int32_t id, std::vector<double*>* out) const {
  const ResidualBlockRef ref = rbs_[id];
  const ResidualTypeStore& t = types_[ref.type];
  const int nb = t.desc.num_parameter_blocks;
  out->clear();
  for (int j = 0; j < nb; ++j)
    out->push_back(pbs_[t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j]]->user_state);
}

void ProblemImpl::GetResidualBlocksForParameterBlock(const double* values,
                                                     std::vector<int32_t>* out) const {
  out->clear();
  const ParameterBlock* pb = FindParameterBlock(values);
  if (!pb) return;
  for (int32_t id = 0; id < static_cast<int32_t>(rbs_.size()); ++id) {
    const ResidualBlockRef ref = rbs_[id];
    const ResidualTypeStore& t = types_[ref.type];
    const int nb = t.desc.num_parameter_blocks;
    for (int j = 0; j < nb; ++j)
      if (t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j] == pb->id) {
        out->push_back(id);
        break;
      }
  }
}

const CostFunction* ProblemImpl::CostFunctionOf(int32_t id) const {
  const ResidualBlockRef ref = rbs_[id];
  return types_[ref.type].cost_functions[ref.local];
}

// ---------------------------------------------------------------------- Program
Program::Program(ProblemImpl* problem) : problem_(problem) {
  for (const auto& pb : problem->parameter_blocks()) parameter_blocks_.push_back(pb.get());
  residual_blocks_.resize(problem->residual_blocks().size());
  std::iota(residual_blocks_.begin(), residual_blocks_.end(), 0);
}

Program::Program(ProblemImpl* problem, std::vector<ParameterBlock*> parameter_blocks,
                 std::vector<int32_t> residual_blocks)
    : problem_(problem),
      parameter_blocks_(std::move(parameter_blocks)),
      residual_blocks_(std::move(residual_blocks)) {
  std::vector<char> listed(problem->parameter_blocks().size(), 0), seen(listed.size(), 0);
  for (const ParameterBlock* pb : parameter_blocks_) listed[pb->id] = 1;
  const auto& rbs = problem->residual_blocks();
  const auto& types = problem->types();
  for (int32_t id : residual_blocks_) {
    const ResidualBlockRef ref = rbs[id];
    const ResidualTypeStore& t = types[ref.type];
    const int nb = t.desc.num_parameter_blocks;
    for (int j = 0; j < nb; ++j) {
      const int pid = t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j];
      if (!listed[pid] && !seen[pid]) {
        seen[pid] = 1;
        constant_parameter_blocks_.push_back(problem->parameter_blocks()[pid].get());
      }
    }
  }
  SetParameterOffsetsAndIndex();
}

void Program::SetParameterOffsetsAndIndex() {
  for (const auto& pb : problem_->parameter_blocks()) pb->index = -1;
  int state_offset = 0, delta_offset = 0;
  for (size_t i = 0; i < parameter_blocks_.size(); ++i) {
    ParameterBlock* pb = parameter_blocks_[i];
    pb->index = static_cast<int>(i);
    pb->state_offset = state_offset;
    pb->delta_offset = delta_offset;
    state_offset += pb->size;
    delta_offset += pb->TangentSize();
  }
  state_offset = 0;
  for (size_t i = 0; i < constant_parameter_blocks_.size(); ++i) {
    ParameterBlock* pb = constant_parameter_blocks_[i];
    pb->index = static_cast<int>(i);
    pb->state_offset = state_offset;
    pb->delta_offset = -1;
    state_offset += pb->size;
  }
}

std::unique_ptr<Program> Program::CreateReducedProgram(
    std::vector<double*>* removed_parameter_blocks, double* fixed_cost,
    std::string* error) const {
  auto reduced = std::make_unique<Program>(*this);
  *fixed_cost = 0.0;
  const auto& rbs = problem_->residual_blocks();
  const auto& types = problem_->types();
  std::vector<char> used(problem_->parameter_blocks().size(), 0);
  std::vector<int32_t> kept;
  kept.reserve(residual_blocks_.size());
  for (int32_t id : residual_blocks_) {
    const ResidualBlockRef ref = rbs[id];
    const ResidualTypeStore& t = types[ref.type];
    const int nb = t.desc.num_parameter_blocks;
    bool all_constant = true;
    for (int j = 0; j < nb; ++j) {
      const int pid = t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j];
      if (!problem_->parameter_blocks()[pid]->IsConstant()) {
        all_constant = false;
        used[pid] = 1;
      }
    }
    if (!all_constant) {
      kept.push_back(id);
      continue;
    }
    double cost = 0.0;
    if (!problem_->EvaluateResidualBlockOnHost(id, true, &cost, nullptr, nullptr)) {
      *error = "Evaluation of the residual " + std::to_string(id) +
               " failed during removal of fixed residual blocks.";
      return nullptr;
    }
    *fixed_cost += cost;
  }
  reduced->residual_blocks_.swap(kept);
  removed_parameter_blocks->clear();
  reduced->constant_parameter_blocks_.clear();
  std::vector<ParameterBlock*> active;
  for (ParameterBlock* pb : parameter_blocks_) {
    if (used[pb->id]) {
      active.push_back(pb);
    } else {
      reduced->constant_parameter_blocks_.push_back(pb);
      removed_parameter_blocks->push_back(pb->user_state);
    }
  }
  reduced->parameter_blocks_.swap(active);
  reduced->SetParameterOffsetsAndIndex();
  return reduced;
}

void Program::ReorderParameterBlocksByGroup(
    const std::unordered_map<const double*, int>& group) {
  auto group_of = [&](const ParameterBlock* pb) {
    auto it = group.find(pb->user_state);
    return it == group.end() ? std::numeric_limits<int>::max() : it->second;
  };
  std::stable_sort(parameter_blocks_.begin(), parameter_blocks_.end(),
                   [&](const ParameterBlock* a, const ParameterBlock* b) {
                     return group_of(a) < group_of(b);
                   });
  SetParameterOffsetsAndIndex();
}

bool Program::LexicographicallyOrderResidualBlocks(int size_of_first_elimination_group) {
  const int E = size_of_first_elimination_group;
  if (E < 1) return false;
  const auto& rbs = problem_->residual_blocks();
  const auto& types = problem_->types();
  const int n = NumResidualBlocks();
  std::vector<int> per_bucket(E + 1, 0), bucket_of(n);
  for (int i = 0; i < n; ++i) {
    const ResidualBlockRef ref = rbs[residual_blocks_[i]];
    const ResidualTypeStore& t = types[ref.type];
    const int nb = t.desc.num_parameter_blocks;
    int position = E;
    for (int j = 0; j < nb; ++j) {
      const ParameterBlock* pb =
          problem_->parameter_blocks()[t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j]]
              .get();
      if (IsActive(pb) && !pb->IsConstant()) position = std::min(position, pb->index);
    }
    bucket_of[i] = position;
    per_bucket[position]++;
  }
  // Buckets are filled back to front, as the reference does, so the order inside
  // an E block is the reverse of the input order.
  std::vector<int> cursor(E + 1);
  std::partial_sum(per_bucket.begin(), per_bucket.end(), cursor.begin());
  std::vector<int32_t> reordered(n, -1);
  for (int i = 0; i < n; ++i) reordered[--cursor[bucket_of[i]]] = residual_blocks_[i];
  residual_blocks_.swap(reordered);
  return true;
}

void Program::ParameterBlocksToStateVector(double* state) const {
  for (const ParameterBlock* pb : parameter_blocks_) {
    std::memcpy(state, pb->user_state, sizeof(double) * pb->size);
    state += pb->size;
  }
}
void Program::StateVectorToParameterBlocks(const double* state) const {
  for (const ParameterBlock* pb : parameter_blocks_) {
    std::memcpy(pb->user_state, state, sizeof(double) * pb->size);
    state += pb->size;
  }
}
void Program::ConstantParameterBlocksToStateVector(double* state) const {
  for (const ParameterBlock* pb : constant_parameter_blocks_) {
    std::memcpy(state, pb->user_state, sizeof(double) * pb->size);
    state += pb->size;
  }
}
bool Program::Plus(const double* state, const double* delta, double* state_plus_delta,
                   int num_threads) const {
  std::atomic<bool> ok{true};
  ParallelFor(num_threads, static_cast<int64_t>(parameter_blocks_.size()),
              [&](int64_t begin, int64_t end, int) {
    for (int64_t k = begin; k < end; ++k) {
      const ParameterBlock* pb = parameter_blocks_[k];
      const double* x = state + pb->state_offset;
      const double* d = delta + pb->delta_offset;
      double* out = state_plus_delta + pb->state_offset;
      if (pb->manifold) {
        if (!pb->manifold->Plus(x, d, out)) ok = false;
      } else {
        for (int i = 0; i < pb->size; ++i) out[i] = x[i] + d[i];
      }
      if (pb->lower_bounds || pb->upper_bounds) {  // parameter_block.h:255-276 projection
        for (int i = 0; i < pb->size; ++i) {
          if (pb->lower_bounds) out[i] = std::max(out[i], pb->lower_bounds[i]);
          if (pb->upper_bounds) out[i] = std::min(out[i], pb->upper_bounds[i]);
        }
      }
    }
  });
  return ok;
}
int Program::NumParameters() const {
  int n = 0;
  for (const ParameterBlock* pb : parameter_blocks_) n += pb->size;
  return n;
}
int Program::NumEffectiveParameters() const {
  int n = 0;
  for (const ParameterBlock* pb : parameter_blocks_) n += pb->TangentSize();
  return n;
}
int Program::NumConstantParameters() const {
  int n = 0;
  for (const ParameterBlock* pb : constant_parameter_blocks_) n += pb->size;
  return n;
}
int Program::NumResiduals() const {
  int n = 0;
  const auto& rbs = problem_->residual_blocks();
  for (int32_t id : residual_blocks_) n += problem_->types()[rbs[id].type].desc.num_residuals;
  return n;
}

// ---------------------------------------------------------------- Jacobian layout
namespace {
struct ActiveBlock {
  int index;     // program index of the parameter block
  int argument;  // position among the ACTIVE arguments of the residual block
  int tangent;
};
// The active parameter blocks of residual block `id`, in argument order.
inline int CollectActive(const Program& program, int32_t id, ActiveBlock* out) {
  const ProblemImpl* problem = program.problem();
  const ResidualBlockRef ref = problem->residual_blocks()[id];
  const ResidualTypeStore& t = problem->types()[ref.type];
  const int nb = t.desc.num_parameter_blocks;
  int n = 0;
  for (int j = 0; j < nb; ++j) {
    const ParameterBlock* pb =
        problem->parameter_blocks()[t.parameter_blocks[static_cast<size_t>(ref.local) * nb + j]]
            .get();
    // a constant block inside the program keeps its columns but gets no cells
    // (block_jacobian_writer.cc:100-103, compressed_row_jacobian_writer.cc:104-107)
    if (program.IsActive(pb) && !pb->IsConstant()) {
      out[n] = ActiveBlock{pb->index, n, pb->TangentSize()};
      ++n;
    }
  }
  return n;
}
inline int NumResidualsOf(const Program& program, int32_t id) {
  const ProblemImpl* problem = program.problem();
  return problem->types()[problem->residual_blocks()[id].type].desc.num_residuals;
}
}  // namespace

void BuildJacobianLayout(const Program& program, int jacobian_format, int num_eliminate_blocks,
                         JacobianLayout* layout) {
  const auto& rbs = program.residual_blocks();
  const int nrb = static_cast<int>(rbs.size());
  layout->jacobian_format = jacobian_format;
  layout->num_eliminate_blocks = num_eliminate_blocks;
  layout->residual_layout.resize(nrb);
  layout->jacobian_per_residual_layout.resize(nrb);
  layout->cell_positions.clear();
  ActiveBlock blocks[CB200_MAX_PARAMETER_BLOCKS];

  // Pass 1: residual positions, size of the E region, size of the offsets table.
  int64_t e_size = 0, total = 0, offsets = 0;
  int residual_pos = 0;
  for (int i = 0; i < nrb; ++i) {
    const int kres = NumResidualsOf(program, rbs[i]);
    layout->residual_layout[i] = residual_pos;
    residual_pos += kres;
    const int na = CollectActive(program, rbs[i], blocks);
    for (int a = 0; a < na; ++a) {
      const int64_t cell = static_cast<int64_t>(kres) * blocks[a].tangent;
      total += cell;
      if (blocks[a].index < num_eliminate_blocks) e_size += cell;
    }
    offsets += static_cast<int64_t>(na) * kres;
  }
  layout->num_residuals = residual_pos;
  layout->num_jacobian_values = total;
  layout->jacobian_per_residual_offsets.assign(offsets, -1);
  if (total > std::numeric_limits<int32_t>::max()) {
    std::fprintf(stderr, "Jacobian has %lld values; the layouts are 32-bit (block_structure.h)\n",
                 static_cast<long long>(total));
    std::abort();
  }

  // Pass 2: positions.
  int32_t* off = layout->jacobian_per_residual_offsets.data();
  int cursor = 0;
  if (jacobian_format == CB200_JACOBIAN_BLOCK_SPARSE) {
    // E cells first in residual-block order, then F cells (block_jacobian_writer.cc:72-150).
    int32_t e_pos = 0, f_pos = static_cast<int32_t>(e_size);
    for (int i = 0; i < nrb; ++i) {
      const int kres = NumResidualsOf(program, rbs[i]);
      layout->jacobian_per_residual_layout[i] = cursor;
      const int na = CollectActive(program, rbs[i], blocks);
      for (int a = 0; a < na; ++a) {
        int32_t& pos = blocks[a].index < num_eliminate_blocks ? e_pos : f_pos;
        layout->cell_positions.push_back(pos);
        for (int k = 0; k < kres; ++k) {
          off[cursor++] = pos;
          pos += blocks[a].tangent;
        }
      }
    }
  } else {
    // Scalar rows; inside a row, blocks in increasing parameter-block index
    // (compressed_row_jacobian_writer.cc:240-300).
    int32_t row_start = 0;
    for (int i = 0; i < nrb; ++i) {
      const int kres = NumResidualsOf(program, rbs[i]);
      layout->jacobian_per_residual_layout[i] = cursor;
      const int na = CollectActive(program, rbs[i], blocks);
      ActiveBlock sorted[CB200_MAX_PARAMETER_BLOCKS];
      std::copy(blocks, blocks + na, sorted);
      std::sort(sorted, sorted + na,
                [](const ActiveBlock& x, const ActiveBlock& y) { return x.index < y.index; });
      int row_width = 0;
      for (int a = 0; a < na; ++a) row_width += sorted[a].tangent;
      int col_pos = 0;
      for (int a = 0; a < na; ++a) {
        for (int r = 0; r < kres; ++r)
          off[cursor + r + kres * sorted[a].argument] = row_start + r * row_width + col_pos;
        col_pos += sorted[a].tangent;
      }
      cursor += na * kres;
      row_start += kres * row_width;
    }
  }
}

std::unique_ptr<SparseMatrix> CreateJacobianFromLayout(const Program& program,
                                                       const JacobianLayout& layout) {
  const auto& rbs = program.residual_blocks();
  const int nrb = static_cast<int>(rbs.size());
  ActiveBlock blocks[CB200_MAX_PARAMETER_BLOCKS];
  if (layout.jacobian_format == CB200_JACOBIAN_BLOCK_SPARSE) {
    // block_jacobian_writer.cc:192-250
    auto* bs = new CompressedRowBlockStructure;
    bs->cols.resize(program.NumParameterBlocks());
    int cursor = 0;
    for (int i = 0; i < program.NumParameterBlocks(); ++i) {
      bs->cols[i].size = program.parameter_blocks()[i]->TangentSize();
      bs->cols[i].position = cursor;
      cursor += bs->cols[i].size;
    }
    bs->rows.resize(nrb);
    bs->row_cell_begin.resize(nrb + 1);
    bs->cells.reserve(layout.cell_positions.size());
    size_t cell_cursor = 0;
    for (int i = 0; i < nrb; ++i) {
      bs->rows[i].size = NumResidualsOf(program, rbs[i]);
      bs->rows[i].position = layout.residual_layout[i];
      bs->row_cell_begin[i] = static_cast<int32_t>(bs->cells.size());
      const int na = CollectActive(program, rbs[i], blocks);
      const size_t first = bs->cells.size();
      for (int a = 0; a < na; ++a) {
        Cell c;
        c.block_id = blocks[a].index;
        c.position = layout.cell_positions[cell_cursor++];
        bs->cells.push_back(c);
      }
      std::sort(bs->cells.begin() + first, bs->cells.end(), [](const Cell& x, const Cell& y) {
        return x.block_id != y.block_id ? x.block_id < y.block_id : x.position < y.position;
      });
    }
    bs->row_cell_begin[nrb] = static_cast<int32_t>(bs->cells.size());
    return std::make_unique<BlockSparseMatrix>(bs, layout.num_jacobian_values);
  }
  // compressed_row_jacobian_writer.cc:93-193
  const int num_cols = program.NumEffectiveParameters();
  auto m = std::make_unique<CompressedRowSparseMatrix>(
      layout.num_residuals, num_cols, layout.num_jacobian_values + num_cols);
  int* rows = m->mutable_rows();
  int* cols = m->mutable_cols();
  rows[0] = 0;
  int row_pos = 0;
  for (int i = 0; i < nrb; ++i) {
    const int kres = NumResidualsOf(program, rbs[i]);
    const int na = CollectActive(program, rbs[i], blocks);
    std::sort(blocks, blocks + na,
              [](const ActiveBlock& x, const ActiveBlock& y) { return x.index < y.index; });
    int width = 0;
    for (int a = 0; a < na; ++a) width += blocks[a].tangent;
    for (int r = 0; r < kres; ++r) rows[row_pos + r + 1] = rows[row_pos + r] + width;
    int col_pos = 0;
    for (int a = 0; a < na; ++a) {
      const int delta = program.parameter_blocks()[blocks[a].index]->delta_offset;
      for (int r = 0; r < kres; ++r)
        for (int c = 0; c < blocks[a].tangent; ++c)
          cols[rows[row_pos + r] + col_pos + c] = delta + c;
      col_pos += blocks[a].tangent;
    }
    row_pos += kres;
  }
  m->set_num_nonzeros(layout.num_jacobian_values);
  // PopulateJacobianRowAndColumnBlockVectors (compressed_row_jacobian_writer.cc:46-70)
  auto& col_blocks = *m->mutable_col_blocks();
  int cursor = 0;
  for (const ParameterBlock* pb : program.parameter_blocks()) {
    Block b; b.size = pb->TangentSize(); b.position = cursor; cursor += b.size;
    col_blocks.push_back(b);
  }
  auto& row_blocks = *m->mutable_row_blocks();
  for (int i = 0; i < nrb; ++i) {
    Block b; b.size = NumResidualsOf(program, rbs[i]); b.position = layout.residual_layout[i];
    row_blocks.push_back(b);
  }
  return m;
}

// ------------------------------------------------------------ ProgramEvaluatorCUDA
namespace {

class ProgramEvaluatorCUDA final : public Evaluator {
 public:
  ProgramEvaluatorCUDA(const Evaluator::Options& options, Program* program, int jacobian_format)
      : options_(options), program_(program) {
    const auto t0 = std::chrono::steady_clock::now();
    BuildJacobianLayout(*program, jacobian_format, std::max(0, options.num_eliminate_blocks),
                        &layout_);
    if (std::getenv("CB200_SETUP_TIMING"))
      std::fprintf(stderr, "setup: %-28s %.3f s\n", "BuildJacobianLayout",
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  ~ProgramEvaluatorCUDA() override {
    if (engine_) cb200_engine_destroy(engine_);
    if (plus_jacobians_) cb200_host_free(plus_jacobians_);
  }

  // RegisteredCUDAEvaluators::Init (registered_cuda_evaluators.cc:226-280).
  bool Init(std::string* error) {
    const bool timing = std::getenv("CB200_SETUP_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      const auto t1 = now();
      std::fprintf(stderr, "setup: %-28s %.3f s\n", what,
                   std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    };
    int rc = cb200_engine_create(options_.device, &engine_);
    if (rc != CB200_OK) {
      *error = "cb200_engine_create failed: no usable CUDA device " +
               std::to_string(options_.device) +
               " (there is no CPU fallback; device -1 only plans the sharding)";
      engine_ = nullptr;
      return false;
    }
    ProblemImpl* problem = program_->problem();
    const auto& active = program_->parameter_blocks();
    const auto& constant = program_->constant_parameter_blocks();
    std::vector<cb200_parameter_block> blocks;
    blocks.reserve(active.size() + constant.size());
    // engine block id of every problem parameter block
    std::vector<int32_t> engine_id(problem->parameter_blocks().size(), -1);
    int plus_pool = 0;
    // Constant blocks that are still part of the program (Problem::Evaluate does not reduce
    // it): they keep their state and gradient slots but are evaluated like the removed ones,
    // from their own (user) state and without derivatives (program.cc:80-88).
    std::vector<const ParameterBlock*> constant_in_program;
    for (const ParameterBlock* pb : active) {
      if (pb->IsConstant()) {
        constant_in_program.push_back(pb);
        continue;
      }
      cb200_parameter_block b;
      b.size = pb->size;
      b.tangent_size = pb->TangentSize();
      b.state_offset = pb->state_offset;
      b.delta_offset = pb->delta_offset;
      b.plus_jacobian_offset = -1;
      b.manifold_kind = CB200_MANIFOLD_NONE;
      b.manifold_param = 0;
      if (pb->manifold) {
        int kind = 0, param = 0;
        if (pb->manifold->DeviceDescription(&kind, &param)) {
          // The kernel applies this manifold itself: nothing to compute or copy per call.
          b.manifold_kind = kind;
          b.manifold_param = param;
        } else {
          b.manifold_kind = CB200_MANIFOLD_GENERIC;
          b.plus_jacobian_offset = plus_pool;
          plus_pool += pb->size * b.tangent_size;
          manifold_blocks_.push_back(pb);
        }
      }
      engine_id[pb->id] = static_cast<int32_t>(blocks.size());
      blocks.push_back(b);
    }
    for (const ParameterBlock* pb : constant) {
      cb200_parameter_block b;
      b.size = pb->size;
      b.tangent_size = pb->TangentSize();
      b.state_offset = pb->state_offset;
      b.delta_offset = -1;
      b.plus_jacobian_offset = -1;
      b.manifold_kind = CB200_MANIFOLD_NONE;
      b.manifold_param = 0;
      engine_id[pb->id] = static_cast<int32_t>(blocks.size());
      blocks.push_back(b);
    }
    int constant_parameters = program_->NumConstantParameters();
    for (const ParameterBlock* pb : constant_in_program) {
      cb200_parameter_block b;
      b.size = pb->size;
      b.tangent_size = pb->TangentSize();
      b.state_offset = constant_parameters;
      b.delta_offset = -1;
      b.plus_jacobian_offset = -1;
      b.manifold_kind = CB200_MANIFOLD_NONE;
      b.manifold_param = 0;
      constant_parameters += pb->size;
      engine_id[pb->id] = static_cast<int32_t>(blocks.size());
      blocks.push_back(b);
    }
    plus_pool_ = plus_pool;
    if (plus_pool > 0)
      plus_jacobians_ = static_cast<double*>(cb200_host_alloc(sizeof(double) * plus_pool));
    std::vector<double> constant_state(constant_parameters + 1);
    program_->ConstantParameterBlocksToStateVector(constant_state.data());
    {
      double* cursor = constant_state.data() + program_->NumConstantParameters();
      for (const ParameterBlock* pb : constant_in_program) {
        std::memcpy(cursor, pb->user_state, sizeof(double) * pb->size);
        cursor += pb->size;
      }
    }
    const int32_t num_engine_active =
        static_cast<int32_t>(active.size() - constant_in_program.size());
    rc = cb200_engine_set_parameter_blocks(
        engine_, num_engine_active,
        static_cast<int32_t>(constant.size() + constant_in_program.size()), blocks.data(),
        program_->NumParameters(), program_->NumEffectiveParameters(), constant_state.data(),
        constant_parameters, plus_pool);
    if (rc != CB200_OK) return Fail(error);

    lap("parameter blocks");
    // Bucket the program's residual blocks by type, keeping their program POSITION.
    const auto& rbs = program_->residual_blocks();
    const auto& refs = problem->residual_blocks();
    auto& types = problem->types();
    std::vector<std::vector<int32_t>> position(types.size()), local(types.size());
    for (int i = 0; i < static_cast<int>(rbs.size()); ++i) {
      const ResidualBlockRef ref = refs[rbs[i]];
      position[ref.type].push_back(i);
      local[ref.type].push_back(ref.local);
    }
    for (size_t ti = 0; ti < types.size(); ++ti) {
      const ResidualTypeStore& t = types[ti];
      const int32_t n = static_cast<int32_t>(position[ti].size());
      if (n == 0) continue;
      const int nb = t.desc.num_parameter_blocks;
      std::vector<int32_t> ids(static_cast<size_t>(n) * nb);
      std::vector<char> functors(static_cast<size_t>(n) * t.desc.functor_size);
      std::vector<int32_t> loss_index(n);
      for (int32_t k = 0; k < n; ++k) {
        const int32_t l = local[ti][k];
        for (int j = 0; j < nb; ++j)
          ids[static_cast<size_t>(k) * nb + j] =
              engine_id[t.parameter_blocks[static_cast<size_t>(l) * nb + j]];
        std::memcpy(functors.data() + static_cast<size_t>(k) * t.desc.functor_size,
                    t.functors.data() + static_cast<size_t>(l) * t.desc.functor_size,
                    t.desc.functor_size);
        loss_index[k] = t.loss_index[l];
      }
      const int32_t num_losses = static_cast<int32_t>(t.loss_objects.size());
      rc = cb200_engine_add_residual_blocks(engine_, &t.desc, n, position[ti].data(), ids.data(),
                                            functors.data(), t.loss_table.data(), num_losses,
                                            num_losses > 1 ? loss_index.data() : nullptr);
      if (rc != CB200_OK) return Fail(error);
    }
    lap("residual blocks -> engine");
    rc = cb200_engine_set_layout(
        engine_, layout_.jacobian_format, static_cast<int32_t>(rbs.size()), layout_.num_residuals,
        layout_.residual_layout.data(), layout_.jacobian_per_residual_layout.data(),
        layout_.jacobian_per_residual_offsets.data(),
        static_cast<int64_t>(layout_.jacobian_per_residual_offsets.size()),
        layout_.num_jacobian_values);
    if (rc != CB200_OK) return Fail(error);
    rc = cb200_engine_set_shard(engine_, options_.shard_rank, options_.shard_world_size);
    if (rc != CB200_OK) return Fail(error);
    lap("set_layout");
    rc = cb200_engine_finalize(engine_);
    lap("engine finalize");
    if (rc != CB200_OK) return Fail(error);
    if (options_.nccl_unique_id && options_.shard_world_size > 1) {
      rc = cb200_engine_comm_init(engine_, options_.nccl_unique_id, options_.shard_rank,
                                  options_.shard_world_size);
      if (rc != CB200_OK) return Fail(error);
    }
    return true;
  }

  std::unique_ptr<SparseMatrix> CreateJacobian() const override {
    if (options_.jacobian_on_device)
      return std::make_unique<DeviceResidentJacobian>(engine_, layout_.num_residuals,
                                                      program_->NumEffectiveParameters(),
                                                      layout_.num_jacobian_values);
    std::unique_ptr<SparseMatrix> m = CreateJacobianFromLayout(*program_, layout_);
    // Page-lock the slices this rank's device writes into.
    int64_t segments[3 * 32];
    const int n = cb200_engine_shard_info(engine_, nullptr, nullptr, nullptr, nullptr, segments, 32);
    for (int i = 0; i < n && i < 32; ++i)
      cb200_host_pin(m->mutable_values() + segments[3 * i], sizeof(double) * segments[3 * i + 1]);
    return m;
  }

  bool Evaluate(const EvaluateOptions& evaluate_options, const double* state, double* cost,
                double* residuals, double* gradient, SparseMatrix* jacobian) override {
    struct Range { Range() { nvtxRangePushA("ProgramEvaluatorCUDA::Evaluate"); } ~Range() { nvtxRangePop(); } } nvtx_range;
    const auto start = std::chrono::steady_clock::now();
    if (options_.device < 0) {
      std::fprintf(stderr, "Evaluate on a planning-only evaluator: no device, no CPU fallback\n");
      return false;
    }
    // Notify the user about the evaluation point (program_evaluator_cuda.h:116-121): the
    // user's parameter blocks are set to `state` first, like
    // Program::StateVectorToParameterBlocks + CopyParameterBlockStateToUserState.
    if (options_.evaluation_callback != nullptr) {
      program_->StateVectorToParameterBlocks(state);
      options_.evaluation_callback->PrepareForEvaluation(
          /*evaluate_jacobians=*/gradient != nullptr || jacobian != nullptr,
          evaluate_options.new_evaluation_point);
    }
    // ParameterBlock::SetState -> UpdatePlusJacobian (parameter_block.h:91-99,312-338):
    // the plus-Jacobians of the blocks with a manifold, at the new state.
    if (jacobian != nullptr || gradient != nullptr) {
      double* cursor = plus_jacobians_;
      for (const ParameterBlock* pb : manifold_blocks_) {
        const int n = pb->size * pb->TangentSize();
        if (!pb->manifold->PlusJacobian(state + pb->state_offset, cursor)) return false;
        for (int i = 0; i < n; ++i)
          if (!std::isfinite(cursor[i])) return false;
        cursor += n;
      }
    }
    uint32_t flags = evaluate_options.apply_loss_function ? CB200_APPLY_LOSS_FUNCTION : 0u;
    // A device-resident Jacobian keeps its values, and the residuals its solver reads, in HBM.
    if (dynamic_cast<DeviceResidentJacobian*>(jacobian) != nullptr)
      flags |= CB200_KEEP_JACOBIAN_ON_DEVICE | CB200_KEEP_RESIDUALS_ON_DEVICE;
    const int rc = cb200_engine_evaluate(engine_, state, plus_jacobians_, flags, cost, residuals,
                                         gradient, jacobian ? jacobian->mutable_values() : nullptr);
    const double seconds =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    CallStatistics& total = statistics_["Evaluator::Total"];
    total.time += seconds;
    total.calls++;
    CallStatistics& kind = statistics_[(gradient == nullptr && jacobian == nullptr)
                                           ? "Evaluator::Residual"
                                           : "Evaluator::Jacobian"];
    kind.time += seconds;
    kind.calls++;
    if (rc < 0) {
      std::fprintf(stderr, "cb200_engine_evaluate: %s\n", cb200_engine_last_error(engine_));
      std::abort();  // CUDA API failure: the reference CHECK-aborts (cuda_buffer.h:61-79)
    }
    return rc == CB200_OK;
  }

  bool Plus(const double* state, const double* delta, double* state_plus_delta) const override {
    return program_->Plus(state, delta, state_plus_delta, options_.num_threads);
  }
  int NumParameters() const override { return program_->NumParameters(); }
  int NumEffectiveParameters() const override { return program_->NumEffectiveParameters(); }
  int NumResiduals() const override { return layout_.num_residuals; }
  std::map<std::string, CallStatistics> Statistics() const override { return statistics_; }
  cb200_engine* engine() const override { return engine_; }
  const JacobianLayout* layout() const override { return &layout_; }

 private:
  bool Fail(std::string* error) {
    *error = std::string("evaluation engine: ") + cb200_engine_last_error(engine_);
    return false;
  }
  Evaluator::Options options_;
  Program* program_;
  JacobianLayout layout_;
  cb200_engine* engine_ = nullptr;
  std::vector<const ParameterBlock*> manifold_blocks_;
  double* plus_jacobians_ = nullptr;
  int plus_pool_ = 0;
  std::map<std::string, CallStatistics> statistics_;
};

}  // namespace

std::unique_ptr<Evaluator> Evaluator::Create(const Evaluator::Options& options, Program* program,
                                             std::string* error) {
  int format;
  switch (options.linear_solver_type) {
    case DENSE_SCHUR:
    case SPARSE_SCHUR:
    case ITERATIVE_SCHUR:
    case CGNR:
      format = (options.sparse_linear_algebra_library_type == CUDA_SPARSE &&
                !options.jacobian_on_device)
                   ? CB200_JACOBIAN_COMPRESSED_ROW
                   : CB200_JACOBIAN_BLOCK_SPARSE;
      break;
    case SPARSE_NORMAL_CHOLESKY:
      if (options.dynamic_sparsity) {
        *error = "The CUDA evaluator stores the Jacobian as BlockSparseMatrix or "
                 "CompressedRowSparseMatrix only (dynamic_sparsity is not supported).";
        return nullptr;
      }
      format = CB200_JACOBIAN_BLOCK_SPARSE;
      break;
    case DENSE_QR:
    case DENSE_NORMAL_CHOLESKY:
      *error = "The CUDA evaluator stores the Jacobian as BlockSparseMatrix or "
               "CompressedRowSparseMatrix only (dense linear solvers are not supported).";
      return nullptr;
    default:
      *error = "Invalid Linear Solver Type. Unable to create evaluator.";
      return nullptr;
  }
  if (!options.use_cuda) {
    *error = "This library only provides the CUDA evaluator (use_cuda must be true).";
    return nullptr;
  }
  auto evaluator = std::make_unique<ProgramEvaluatorCUDA>(options, program, format);
  if (!evaluator->Init(error)) return nullptr;
  return evaluator;
}

}  // namespace internal

// problem_impl.cc:599-760, on the CUDA evaluator.
bool Problem::Evaluate(const EvaluateOptions& options, double* cost, std::vector<double>* residuals,
                       std::vector<double>* gradient, CRSMatrix* jacobian) {
  if (!cost && !residuals && !gradient && !jacobian) return true;
  internal::ProblemImpl* impl = impl_.get();
  std::vector<int32_t> rbs;
  if (options.residual_blocks.empty()) {
    rbs.resize(impl->NumResidualBlocks());
    std::iota(rbs.begin(), rbs.end(), 0);
  } else {
    for (ResidualBlockId id : options.residual_blocks) rbs.push_back(internal::ProblemImpl::IdOf(id));
  }
  std::vector<internal::ParameterBlock*> pbs;
  std::vector<internal::ParameterBlock*> held;  // excluded blocks held constant for the call
  if (options.parameter_blocks.empty()) {
    for (const auto& pb : impl->parameter_blocks()) pbs.push_back(pb.get());
  } else {
    std::vector<char> included(impl->parameter_blocks().size(), 0);
    for (double* values : options.parameter_blocks) {
      internal::ParameterBlock* pb = impl->FindParameterBlock(values);
      if (!pb) {
        std::fprintf(stderr, "No known parameter block for Problem::Evaluate::Options."
                             "parameter_blocks = %p\n", static_cast<void*>(values));
        std::abort();  // the reference LOG(FATAL)s (problem_impl.cc:638-642)
      }
      pbs.push_back(pb);
      included[pb->id] = 1;
    }
    for (const auto& pb : impl->parameter_blocks())
      if (!included[pb->id] && !pb->IsConstant()) {
        held.push_back(pb.get());
        pb->is_set_constant = true;
      }
  }
  bool status = false;
  {
    internal::Program program(impl, pbs, rbs);
    internal::RegisteredCUDAEvaluators registry(impl);
    internal::Evaluator::Options eo;
    eo.linear_solver_type = CGNR;  // with CUDA_SPARSE: "use a compressed-row Jacobian"
    eo.sparse_linear_algebra_library_type = CUDA_SPARSE;
    eo.num_threads = options.num_threads;
    eo.use_cuda = true;
    eo.registered_cuda_evaluators = &registry;
    eo.evaluation_callback = impl->options().evaluation_callback;
    eo.device = options.cuda_device;
    eo.num_eliminate_blocks = 0;
    std::string error;
    std::unique_ptr<internal::Evaluator> evaluator =
        internal::Evaluator::Create(eo, &program, &error);
    if (!evaluator) {
      std::fprintf(stderr, "Problem::Evaluate: %s\n", error.c_str());
    } else {
      if (residuals) residuals->assign(evaluator->NumResiduals(), 0.0);
      if (gradient) gradient->assign(evaluator->NumEffectiveParameters(), 0.0);
      std::unique_ptr<internal::SparseMatrix> tmp_jacobian;
      if (jacobian) tmp_jacobian = evaluator->CreateJacobian();
      std::vector<double> parameters(program.NumParameters() + 1);
      program.ParameterBlocksToStateVector(parameters.data());
      double tmp_cost = 0.0;
      internal::Evaluator::EvaluateOptions evaluate_options;
      evaluate_options.apply_loss_function = options.apply_loss_function;
      status = evaluator->Evaluate(evaluate_options, parameters.data(), &tmp_cost,
                                   residuals && !residuals->empty() ? residuals->data() : nullptr,
                                   gradient && !gradient->empty() ? gradient->data() : nullptr,
                                   tmp_jacobian.get());
      if (status) {
        if (cost) *cost = tmp_cost;
        if (jacobian) {
          // CompressedRowSparseMatrix::ToCRSMatrix
          auto* crs = static_cast<internal::CompressedRowSparseMatrix*>(tmp_jacobian.get());
          jacobian->num_rows = crs->num_rows();
          jacobian->num_cols = crs->num_cols();
          jacobian->rows.assign(crs->rows(), crs->rows() + crs->num_rows() + 1);
          jacobian->cols.assign(crs->cols(), crs->cols() + crs->num_nonzeros());
          jacobian->values.assign(crs->values(), crs->values() + crs->num_nonzeros());
        }
      }
    }
  }
  for (internal::ParameterBlock* pb : held) pb->is_set_constant = false;
  return status;
}

}  // namespace ceres
