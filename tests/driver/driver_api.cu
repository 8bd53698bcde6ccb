// C interface of the test/bench driver (see driver.h).  Python binds it with
// ctypes (ceres-solver-cuda_b200/binding.py).
#include <cstdlib>
#include <cstring>

#include "driver.h"

using driver::DriverProblem;
using ceres::internal::Evaluator;

namespace {
// A SubsetManifold the device is not told about: exercises the generic path
// (Manifold::PlusJacobian on the host every Evaluate + plus-Jacobian pool upload).
class OpaqueSubsetManifold : public ceres::SubsetManifold {
 public:
  using ceres::SubsetManifold::SubsetManifold;
  bool DeviceDescription(int*, int*) const override { return false; }
};

ceres::Manifold* MakeManifold(int kind, int param, int size) {
  switch (kind) {
    case 6: {
      std::vector<int> constant;
      for (int i = 0; i < size; ++i)
        if ((param >> i) & 1) constant.push_back(i);
      return new OpaqueSubsetManifold(size, constant);
    }
    case 1: {  // SubsetManifold, param = bitmask of constant coordinates
      std::vector<int> constant;
      for (int i = 0; i < size; ++i)
        if ((param >> i) & 1) constant.push_back(i);
      return new ceres::SubsetManifold(size, constant);
    }
    case 2: return new ceres::QuaternionManifold;
    case 3: return new ceres::EigenQuaternionManifold;
    case 4:
      switch (size - 4) {
        case 3: return new ceres::ProductManifold<ceres::QuaternionManifold, ceres::EuclideanManifold<3>>;
        case 6: return new ceres::ProductManifold<ceres::QuaternionManifold, ceres::EuclideanManifold<6>>;
      }
      return nullptr;
    case 5:
      switch (size - 4) {
        case 3: return new ceres::ProductManifold<ceres::EigenQuaternionManifold, ceres::EuclideanManifold<3>>;
        case 6: return new ceres::ProductManifold<ceres::EigenQuaternionManifold, ceres::EuclideanManifold<6>>;
      }
      return nullptr;
  }
  return nullptr;
}

const int kNumTypes = 18;
const int kTypeBlocks[kNumTypes] = {2, 2, 2, 1, 2, 2, 10, 1, 3, 3, 2, 2, 2, 3, 1, 4, 1, 1};
const int kTypeFdata[kNumTypes] = {2, 2, 2, 3, 7, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 43, 0, 1};
}  // namespace

extern "C" {

// Builds the ProblemCUDA.  bulk != 0 uses ProblemCUDA::AddResidualBlocks for runs of
// residual blocks of one type/loss; otherwise every block goes through
// AddResidualBlock<...> with its own AutoDiffCostFunction, like user code.
void* drv_create(int num_pb, const int* pb_size, const double* pb_values,
                 const uint8_t* pb_constant, const int* pb_manifold_kind,
                 const int* pb_manifold_param, int num_rb, const int* rb_type, const int* rb_pb,
                 const int* rb_loss_kind, const double* rb_loss_a, const double* rb_loss_b,
                 const double* fdata, int bulk) {
  auto* dp = new DriverProblem((bulk & 2) != 0);  // bit 1: install the EvaluationCallback
  bulk &= 1;
  dp->pb_offset.resize(num_pb);
  dp->pb_size.assign(pb_size, pb_size + num_pb);
  int64_t off = 0;
  for (int i = 0; i < num_pb; ++i) {
    dp->pb_offset[i] = off;
    off += pb_size[i];
  }
  dp->values.assign(pb_values, pb_values + off);
  for (int i = 0; i < num_pb; ++i) dp->problem.AddParameterBlock(dp->pb(i), pb_size[i]);

  int64_t pb_cursor = 0, fd_cursor = 0;
  int i = 0;
  while (i < num_rb) {
    const int type = rb_type[i];
    if (type < 0 || type >= kNumTypes) { dp->error = "unknown cost type"; return dp; }
    int j = i + 1;
    while (j < num_rb && rb_type[j] == type && rb_loss_kind[j] == rb_loss_kind[i] &&
           rb_loss_a[j] == rb_loss_a[i] && rb_loss_b[j] == rb_loss_b[i])
      ++j;
    const int n = j - i;
    bool handled = false, ok = false;
    ok = driver::AddRunBal(*dp, type, rb_loss_kind[i], rb_loss_a[i], rb_loss_b[i], n,
                           rb_pb + pb_cursor, fdata + fd_cursor, bulk != 0, &handled);
    if (!handled)
      ok = driver::AddRunPose(*dp, type, rb_loss_kind[i], rb_loss_a[i], rb_loss_b[i], n,
                              rb_pb + pb_cursor, fdata + fd_cursor, bulk != 0, &handled);
    if (!handled)
      ok = driver::AddRunTests(*dp, type, rb_loss_kind[i], rb_loss_a[i], rb_loss_b[i], n,
                               rb_pb + pb_cursor, fdata + fd_cursor, bulk != 0, &handled);
    if (!handled || !ok) {
      dp->error = "cost type / loss combination not instantiated in the test driver";
      return dp;
    }
    pb_cursor += static_cast<int64_t>(n) * kTypeBlocks[type];
    fd_cursor += static_cast<int64_t>(n) * kTypeFdata[type];
    i = j;
  }
  for (int k = 0; k < num_pb; ++k) {
    if (pb_manifold_kind && pb_manifold_kind[k] != 0) {
      ceres::Manifold* m = MakeManifold(pb_manifold_kind[k], pb_manifold_param[k], pb_size[k]);
      if (!m) { dp->error = "manifold kind not available in the test driver"; return dp; }
      dp->problem.SetManifold(dp->pb(k), m);
    }
    if (pb_constant && pb_constant[k]) dp->problem.SetParameterBlockConstant(dp->pb(k));
  }
  return dp;
}

void drv_destroy(void* h) { delete static_cast<DriverProblem*>(h); }

const char* drv_error(void* h) { return static_cast<DriverProblem*>(h)->error.c_str(); }

// Program + layout (+ evaluator when with_device != 0).  Returns 1 on success.
int drv_build(void* h, int reduce, int schur_reorder, int num_eliminate_blocks,
              int jacobian_format, int with_device, int device, int rank, int world_size,
              const void* nccl_unique_id) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->error.empty()) return 0;
  ceres::internal::ProblemImpl* impl = dp->problem.mutable_problem()->mutable_impl();
  dp->evaluator.reset();
  dp->jacobian.reset();
  dp->full_program = std::make_unique<ceres::internal::Program>(impl);
  dp->fixed_cost = 0.0;
  if (reduce) {
    std::vector<double*> removed;
    dp->program = dp->full_program->CreateReducedProgram(&removed, &dp->fixed_cost, &dp->error);
    if (!dp->program) return 0;
  } else {
    dp->program = std::make_unique<ceres::internal::Program>(*dp->full_program);
    dp->program->SetParameterOffsetsAndIndex();
  }
  if (schur_reorder && num_eliminate_blocks > 0)
    dp->program->LexicographicallyOrderResidualBlocks(num_eliminate_blocks);

  if (!with_device) {
    ceres::internal::BuildJacobianLayout(*dp->program, jacobian_format, num_eliminate_blocks,
                                         &dp->layout);
    dp->jacobian = ceres::internal::CreateJacobianFromLayout(*dp->program, dp->layout);
    return 1;
  }
  Evaluator::Options options;
  options.num_eliminate_blocks = num_eliminate_blocks;
  options.linear_solver_type = ceres::ITERATIVE_SCHUR;
  options.sparse_linear_algebra_library_type =
      jacobian_format == CB200_JACOBIAN_COMPRESSED_ROW ? ceres::CUDA_SPARSE : ceres::NO_SPARSE;
  options.use_cuda = true;
  options.registered_cuda_evaluators = dp->problem.mutable_registered_cuda_evaluators();
  options.device = device;
  options.shard_rank = rank;
  options.shard_world_size = world_size;
  options.nccl_unique_id = nccl_unique_id;
  options.evaluation_callback = dp->problem.mutable_problem()->options().evaluation_callback;
  dp->evaluator = Evaluator::Create(options, dp->program.get(), &dp->error);
  if (!dp->evaluator) return 0;
  dp->layout = *dp->evaluator->layout();
  dp->jacobian = dp->evaluator->CreateJacobian();
  return 1;
}

// dims: [num_parameters, num_effective_parameters, num_residuals, num_residual_blocks,
//        num_parameter_blocks, num_jacobian_values, values_size, num_cells,
//        per_residual_offsets_size, num_constant_parameters]
void drv_dims(void* h, int64_t* dims) {
  auto* dp = static_cast<DriverProblem*>(h);
  const auto& p = *dp->program;
  dims[0] = p.NumParameters();
  dims[1] = p.NumEffectiveParameters();
  dims[2] = dp->layout.num_residuals;
  dims[3] = p.NumResidualBlocks();
  dims[4] = p.NumParameterBlocks();
  dims[5] = dp->layout.num_jacobian_values;
  dims[6] = dp->jacobian ? dp->jacobian->values_size() : 0;
  dims[7] = static_cast<int64_t>(dp->layout.cell_positions.size());
  dims[8] = static_cast<int64_t>(dp->layout.jacobian_per_residual_offsets.size());
  dims[9] = p.NumConstantParameters();
}

double drv_fixed_cost(void* h) { return static_cast<DriverProblem*>(h)->fixed_cost; }

void drv_initial_state(void* h, double* state) {
  static_cast<DriverProblem*>(h)->program->ParameterBlocksToStateVector(state);
}

// Same `which` numbering as oracle_problem_get_ints.
int64_t drv_get_ints(void* h, int which, int* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  std::vector<int> tmp;
  const std::vector<int>* v = &tmp;
  const auto* bsm = dynamic_cast<const ceres::internal::BlockSparseMatrix*>(dp->jacobian.get());
  const auto* crs =
      dynamic_cast<const ceres::internal::CompressedRowSparseMatrix*>(dp->jacobian.get());
  switch (which) {
    case 0: v = &dp->layout.residual_layout; break;
    case 1: v = &dp->layout.jacobian_per_residual_layout; break;
    case 2: v = &dp->layout.jacobian_per_residual_offsets; break;
    case 3: tmp.assign(dp->program->residual_blocks().begin(), dp->program->residual_blocks().end()); break;
    case 4: for (auto* pb : dp->program->parameter_blocks()) tmp.push_back(pb->id); break;
    case 5: if (bsm) for (auto& b : bsm->block_structure()->cols) tmp.push_back(b.size); break;
    case 6: if (bsm) for (auto& b : bsm->block_structure()->cols) tmp.push_back(b.position); break;
    case 7: if (bsm) for (auto& b : bsm->block_structure()->rows) tmp.push_back(b.size); break;
    case 8: if (bsm) for (auto& b : bsm->block_structure()->rows) tmp.push_back(b.position); break;
    case 9: if (bsm) tmp = bsm->block_structure()->row_cell_begin; break;
    case 10: if (bsm) for (auto& c : bsm->block_structure()->cells) tmp.push_back(c.block_id); break;
    case 11: if (bsm) for (auto& c : bsm->block_structure()->cells) tmp.push_back(c.position); break;
    case 12: if (crs) tmp.assign(crs->rows(), crs->rows() + crs->num_rows() + 1); break;
    case 13: if (crs) tmp.assign(crs->cols(), crs->cols() + crs->values_size()); break;
    case 14: for (auto* pb : dp->program->constant_parameter_blocks()) tmp.push_back(pb->id); break;
    case 15: v = &dp->layout.cell_positions; break;
    default: return -1;
  }
  if (out) std::copy(v->begin(), v->end(), out);
  return static_cast<int64_t>(v->size());
}

void drv_pb_table(void* h, int* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  int i = 0;
  for (auto* pb : dp->program->parameter_blocks()) {
    out[4 * i + 0] = pb->index;
    out[4 * i + 1] = pb->state_offset;
    out[4 * i + 2] = pb->delta_offset;
    out[4 * i + 3] = pb->TangentSize();
    ++i;
  }
}

// The pinned values array of the Jacobian created by Evaluator::CreateJacobian().
double* drv_jacobian_values(void* h) {
  auto* dp = static_cast<DriverProblem*>(h);
  return dp->jacobian ? dp->jacobian->mutable_values() : nullptr;
}

// Evaluator::Evaluate.  want_jacobian writes into drv_jacobian_values().
// Returns 1 (true), 0 (false: evaluation failed), -1 (no evaluator).
// out: [calls, evaluate_jacobians, new_evaluation_point, sum of the user's parameter values]
// as the EvaluationCallback saw them at its last notification.
void drv_callback_info(void* h, double* out4) {
  auto* dp = static_cast<DriverProblem*>(h);
  out4[0] = dp->callback.calls;
  out4[1] = dp->callback.evaluate_jacobians;
  out4[2] = dp->callback.new_evaluation_point;
  out4[3] = dp->callback.user_value_sum;
}

// apply_loss_function: bit 0; bit 1 set = EvaluateOptions::new_evaluation_point false
int drv_evaluate(void* h, const double* state, int apply_loss_function, double* cost,
                 double* residuals, double* gradient, int want_jacobian) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -1;
  Evaluator::EvaluateOptions eo;
  eo.apply_loss_function = (apply_loss_function & 1) != 0;
  eo.new_evaluation_point = (apply_loss_function & 2) == 0;
  return dp->evaluator->Evaluate(eo, state, cost, residuals, gradient,
                                 want_jacobian ? dp->jacobian.get() : nullptr)
             ? 1
             : 0;
}

// Device-resident evaluation through the C ABI (inputs already in HBM).
int drv_evaluate_device(void* h, int apply_loss_function, int want_residuals, int want_gradient,
                        int want_jacobian, double* cost) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -1;
  return cb200_engine_evaluate_device(dp->evaluator->engine(), nullptr, nullptr,
                                      apply_loss_function ? CB200_APPLY_LOSS_FUNCTION : 0u,
                                      want_residuals, want_gradient, want_jacobian, cost);
}

// `steps` back-to-back calls of cb200_engine_evaluate_device (each one a complete evaluation:
// launches, cost read back, stream synchronised) with the per-call timings of
// cb200_engine_last_timing collected on the way - bench.py's timed loop without an
// interpreter round trip per step.  Returns the last call's status.
int drv_evaluate_device_steps(void* h, int steps, int apply_loss_function, int want_residuals,
                              int want_gradient, int want_jacobian, double* kernel_ms,
                              double* device_ms, double* cost) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -1;
  int rc = 0;
  for (int k = 0; k < steps; ++k) {
    rc = cb200_engine_evaluate_device(dp->evaluator->engine(), nullptr, nullptr,
                                      apply_loss_function ? CB200_APPLY_LOSS_FUNCTION : 0u,
                                      want_residuals, want_gradient, want_jacobian, cost);
    if (rc < 0) return rc;
    double t[4] = {0, 0, 0, 0};
    cb200_engine_last_timing(dp->evaluator->engine(), t);
    if (kernel_ms) kernel_ms[k] = t[0];
    if (device_ms) device_ms[k] = t[1];
  }
  return rc;
}

// Problem::Evaluate through ProblemCUDA's wrapped Problem.  parameter_blocks / residual_blocks
// are ids in creation order (NULL = all).  dims: [num_residuals, num_gradient, num_rows,
// num_cols, num_nonzeros]; the arrays are then fetched with drv_problem_evaluate_get.
int drv_problem_evaluate(void* h, int apply_loss_function, int num_pb, const int* pb_ids,
                         int num_rb, const int* rb_ids, int device, double* cost, int64_t* dims) {
  auto* dp = static_cast<DriverProblem*>(h);
  ceres::Problem::EvaluateOptions options;
  options.apply_loss_function = apply_loss_function != 0;
  options.cuda_device = device;
  for (int i = 0; i < num_pb; ++i) options.parameter_blocks.push_back(dp->pb(pb_ids[i]));
  std::vector<ceres::ResidualBlockId> all;
  dp->problem.GetResidualBlocks(&all);
  for (int i = 0; i < num_rb; ++i) options.residual_blocks.push_back(all[rb_ids[i]]);
  const bool ok =
      dp->problem.Evaluate(options, cost, &dp->pe_residuals, &dp->pe_gradient, &dp->pe_jacobian);
  dims[0] = static_cast<int64_t>(dp->pe_residuals.size());
  dims[1] = static_cast<int64_t>(dp->pe_gradient.size());
  dims[2] = dp->pe_jacobian.num_rows;
  dims[3] = dp->pe_jacobian.num_cols;
  dims[4] = static_cast<int64_t>(dp->pe_jacobian.values.size());
  return ok ? 1 : 0;
}
// which: 0 residuals, 1 gradient, 2 rows (int32), 3 cols (int32), 4 values
void drv_problem_evaluate_get(void* h, int which, void* dst) {
  auto* dp = static_cast<DriverProblem*>(h);
  switch (which) {
    case 0: std::memcpy(dst, dp->pe_residuals.data(), dp->pe_residuals.size() * 8); break;
    case 1: std::memcpy(dst, dp->pe_gradient.data(), dp->pe_gradient.size() * 8); break;
    case 2: std::memcpy(dst, dp->pe_jacobian.rows.data(), dp->pe_jacobian.rows.size() * 4); break;
    case 3: std::memcpy(dst, dp->pe_jacobian.cols.data(), dp->pe_jacobian.cols.size() * 4); break;
    default: std::memcpy(dst, dp->pe_jacobian.values.data(), dp->pe_jacobian.values.size() * 8);
  }
}

// Problem::EvaluateResidualBlock (host).  jacobians: the blocks of the non-constant
// arguments, each num_residuals x tangent, concatenated in argument order; want_jacobians
// = 0 skips them.  Also exercises the query functions: parameter block ids of the residual
// block are written to pb_ids_out.
int drv_evaluate_residual_block(void* h, int rb, int num_residuals, int apply_loss_function,
                                int want_jacobians, double* cost, double* residuals,
                                double* jacobians, int* pb_ids_out) {
  auto* dp = static_cast<DriverProblem*>(h);
  ceres::Problem* problem = dp->problem.mutable_problem();
  std::vector<ceres::ResidualBlockId> all;
  problem->GetResidualBlocks(&all);
  std::vector<double*> blocks;
  problem->GetParameterBlocksForResidualBlock(all[rb], &blocks);
  std::vector<double*> ptrs(blocks.size(), nullptr);
  double* cursor = jacobians;
  for (size_t j = 0; j < blocks.size(); ++j) {
    for (size_t k = 0; k < dp->pb_offset.size(); ++k)
      if (dp->pb(static_cast<int>(k)) == blocks[j]) pb_ids_out[j] = static_cast<int>(k);
    if (!want_jacobians || problem->IsParameterBlockConstant(blocks[j])) continue;
    ptrs[j] = cursor;
    cursor += num_residuals * problem->ParameterBlockTangentSize(blocks[j]);
  }
  return problem->EvaluateResidualBlock(all[rb], apply_loss_function != 0, cost, residuals,
                                        want_jacobians ? ptrs.data() : nullptr)
             ? 1
             : 0;
}

// Linear algebra on the Jacobian the last evaluation left on the device (C ABI pass-through).
static int LaResult(DriverProblem* dp, int rc) {
  if (rc != CB200_OK) dp->error = cb200_engine_last_error(dp->evaluator->engine());
  return rc;
}
int drv_jacobian_multiply(void* h, int transpose, const double* x, double* y) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -100;
  return LaResult(dp, cb200_engine_jacobian_multiply(dp->evaluator->engine(), transpose, x, y));
}
int drv_jacobian_squared_column_norm(void* h, double* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -100;
  return LaResult(dp, cb200_engine_jacobian_squared_column_norm(dp->evaluator->engine(), out));
}
int drv_jacobian_scale_columns(void* h, const double* scale) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -100;
  return LaResult(dp, cb200_engine_jacobian_scale_columns(dp->evaluator->engine(), scale));
}
// out: [iterations, termination, |J'b|, final residual norm, (Jy).b, |Jy|^2, ms]
int drv_cgnr_solve(void* h, const double* d_squared, int min_iterations, int max_iterations,
                   double r_tolerance, double q_tolerance, double* solution, double* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -100;
  cb200_cgnr_options o{min_iterations, max_iterations, r_tolerance, q_tolerance};
  cb200_cgnr_summary s;
  const int rc = cb200_engine_cgnr_solve(dp->evaluator->engine(), d_squared, &o, solution, &s);
  if (rc != CB200_OK) { dp->error = cb200_engine_last_error(dp->evaluator->engine()); return rc; }
  out[0] = s.num_iterations; out[1] = s.termination; out[2] = s.initial_gradient_norm;
  out[3] = s.final_residual_norm; out[4] = s.jy_dot_b; out[5] = s.jy_squared_norm;
  out[6] = s.solve_ms;
  return rc;
}

int drv_timing(void* h, double* out4) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -1;
  return cb200_engine_last_timing(dp->evaluator->engine(), out4);
}

int drv_shard_info(void* h, int32_t* info4, int64_t* segments, int max_segments) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -1;
  return cb200_engine_shard_info(dp->evaluator->engine(), info4, info4 + 1, info4 + 2, info4 + 3,
                                 segments, max_segments);
}

int drv_exchange_mode(void* h) {
  auto* dp = static_cast<DriverProblem*>(h);
  return dp->evaluator ? cb200_engine_exchange_mode(dp->evaluator->engine()) : -1;
}

int drv_exchange_plan(void* h, int32_t* chunks, int max_chunks, int64_t* exclusive,
                      int32_t* shared_count) {
  auto* dp = static_cast<DriverProblem*>(h);
  if (!dp->evaluator) return -2;
  return cb200_engine_exchange_plan(dp->evaluator->engine(), chunks, max_chunks, exclusive,
                                    shared_count);
}

int drv_plus(void* h, const double* state, const double* delta, double* out) {
  return static_cast<DriverProblem*>(h)->program->Plus(state, delta, out) ? 1 : 0;
}
// Program::Plus on num_threads host threads (program.cc:121-150).
int drv_plus_threads(void* h, const double* state, const double* delta, double* out,
                     int num_threads) {
  return static_cast<DriverProblem*>(h)->program->Plus(state, delta, out, num_threads) ? 1 : 0;
}

// ceres::Solve(options, ProblemCUDA*, summary) on the driver's problem.  ordering: group id
// per parameter block (or NULL).  out: [initial_cost, final_cost, iterations,
// successful_steps, termination_type, num_jacobian_evaluations, num_residual_evaluations].
// The solution is written to the driver's parameter values (drv_user_values).
int drv_solve(void* h, int linear_solver_type, int cuda_sparse, int max_num_iterations,
              const int* ordering, int device, double* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  ceres::Solver::Options options;
  options.linear_solver_type = static_cast<ceres::LinearSolverType>(linear_solver_type);
  if (cuda_sparse) options.sparse_linear_algebra_library_type = ceres::CUDA_SPARSE;
  options.max_num_iterations = max_num_iterations;
  options.cuda_device = device;
  // (tests that compare two trust-region loops ask for near-exact steps: with the default
  // eta = 0.1 the point where conjugate gradients stop depends on the last bits of the sums)
  if (const char* eta = std::getenv("CB200_DRIVER_ETA")) options.eta = std::atof(eta);
  if (ordering) {
    auto* o = new ceres::ParameterBlockOrdering;
    for (size_t k = 0; k < dp->pb_offset.size(); ++k)
      o->AddElementToGroup(dp->pb(static_cast<int>(k)), ordering[k]);
    options.linear_solver_ordering.reset(o);
  }
  ceres::Solver::Summary summary;
  ceres::Solve(options, &dp->problem, &summary);
  dp->error = summary.message;
  out[0] = summary.initial_cost;
  out[1] = summary.final_cost;
  out[2] = static_cast<double>(summary.iterations.size());
  out[3] = summary.num_successful_steps;
  out[4] = static_cast<double>(summary.termination_type);
  out[5] = summary.num_jacobian_evaluations;
  out[6] = summary.num_residual_evaluations;
  out[7] = summary.linear_solver_time_in_seconds;
  return summary.IsSolutionUsable() ? 1 : 0;
}

// Solver::Options::IsValid for a given minimizer type (0 LINE_SEARCH, 1 TRUST_REGION) with
// the CUDA evaluator requested.  Returns 1 when valid; the message is copied to msg.
int drv_options_is_valid(void* h, int minimizer_type, char* msg, int msg_size) {
  auto* dp = static_cast<DriverProblem*>(h);
  ceres::Solver::Options options;
  options.minimizer_type = static_cast<ceres::MinimizerType>(minimizer_type);
  options.use_cuda_for_evaluator = true;
  options.registered_cuda_evaluators = dp->problem.mutable_registered_cuda_evaluators();
  std::string error;
  const bool ok = options.IsValid(&error);
  std::snprintf(msg, msg_size, "%s", error.c_str());
  return ok ? 1 : 0;
}

// internal::ArgumentSlotOrdering (the resident Jacobian's two-region layout when the caller gives
// no ordering): group per parameter block in creation order, -1 when the block got none.
// Returns 1 when an ordering was found.  Host code only.
int drv_argument_slot_ordering(void* h, int* groups) {
  auto* dp = static_cast<DriverProblem*>(h);
  ceres::ParameterBlockOrdering ordering;
  const bool found = ceres::internal::ArgumentSlotOrdering(
      *dp->problem.mutable_problem()->mutable_impl(), &ordering);
  for (size_t k = 0; k < dp->pb_offset.size(); ++k)
    groups[k] = found ? ordering.GroupId(dp->pb(static_cast<int>(k))) : -1;
  return found ? 1 : 0;
}

void drv_user_values(void* h, double* out) {
  auto* dp = static_cast<DriverProblem*>(h);
  std::memcpy(out, dp->values.data(), dp->values.size() * sizeof(double));
}

void drv_dense_jacobian(void* h, double* dense) {
  auto* dp = static_cast<DriverProblem*>(h);
  std::vector<double> d;
  dp->jacobian->ToDenseMatrix(&d);
  std::memcpy(dense, d.data(), d.size() * sizeof(double));
}

}  // extern "C"
