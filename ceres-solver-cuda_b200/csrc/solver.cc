// Minimal TRUST_REGION driver around the CUDA evaluator; see ceres/solver.h for scope.
#include "ceres/solver.h"

#include "ceres/internal/parallel_for.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <nvtx3/nvToolsExt.h>

namespace ceres {
namespace {

double Seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

// The vectors that cross PCIe every iteration (state, gradient, step, LM diagonal) live in
// page-locked memory: pageable copies of 108 MB run at a fraction of the link rate.
class PinnedVector {
 public:
  explicit PinnedVector(size_t n = 0, double value = 0.0) { assign(n, value); }
  ~PinnedVector() { Release(); }
  PinnedVector(const PinnedVector&) = delete;
  PinnedVector& operator=(const PinnedVector& other) {
    if (size_ != other.size_) assign(other.size_, 0.0);
    std::copy(other.begin(), other.end(), begin());
    return *this;
  }
  void assign(size_t n, double value) {
    if (n != size_) {
      Release();
      size_ = n;
      if (n > 0) {
        data_ = static_cast<double*>(cb200_host_alloc(n * sizeof(double)));
        cb200_host_pin(data_, n * sizeof(double));
      }
    }
    std::fill(begin(), end(), value);
  }
  size_t size() const { return size_; }
  double* data() { return data_; }
  const double* data() const { return data_; }
  double& operator[](size_t i) { return data_[i]; }
  const double& operator[](size_t i) const { return data_[i]; }
  double* begin() { return data_; }
  double* end() { return data_ + size_; }
  const double* begin() const { return data_; }
  const double* end() const { return data_ + size_; }

 private:
  void Release() {
    if (data_) cb200_host_free(data_);
    data_ = nullptr;
    size_ = 0;
  }
  double* data_ = nullptr;
  size_t size_ = 0;
};

// Wall-clock buckets of the trust-region loop, printed when CB200_SOLVER_TIMING is set.
struct Buckets {
  double diagonal = 0, scaling = 0, step = 0, plus = 0, bookkeeping = 0;
};

// Element-wise pass over [0, n) on num_threads threads; f(i) must only touch index i.
template <typename F>
void ForEach(int num_threads, int64_t n, F&& f) {
  internal::ParallelFor(num_threads, n, [&](int64_t begin, int64_t end, int) {
    for (int64_t i = begin; i < end; ++i) f(i);
  });
}
// Sum (or maximum) of f(i) over [0, n) with one partial per thread.
template <typename F>
double Reduce(int num_threads, int64_t n, bool maximum, F&& f) {
  std::vector<double> partial(std::max(1, num_threads), 0.0);
  internal::ParallelFor(num_threads, n, [&](int64_t begin, int64_t end, int t) {
    double acc = 0.0;
    for (int64_t i = begin; i < end; ++i) acc = maximum ? std::max(acc, f(i)) : acc + f(i);
    partial[t] = acc;
  });
  double total = 0.0;
  for (double p : partial) total = maximum ? std::max(total, p) : total + p;
  return total;
}

// Solves (J'J + D^2) y = J'r with Jacobi-preconditioned conjugate gradients (CGNR);
// stand-in for the reference's linear solvers.  Returns the iteration count.
int SolveNormalEquations(const internal::SparseMatrix& J, const double* D2, const double* r,
                         int max_iterations, double tolerance, PinnedVector* y) {
  const int n = J.num_cols(), m = J.num_rows();
  std::vector<double> b(n, 0.0), res(n), z(n), p(n), Ap(n), tmp(m), precond(n);
  J.LeftMultiplyAndAccumulate(r, b.data());
  J.SquaredColumnNorm(precond.data());
  for (int i = 0; i < n; ++i) precond[i] = 1.0 / (precond[i] + D2[i]);
  y->assign(n, 0.0);
  res = b;
  double norm_b = 0.0;
  for (int i = 0; i < n; ++i) norm_b += b[i] * b[i];
  norm_b = std::sqrt(norm_b);
  if (norm_b == 0.0) return 0;
  double rho = 0.0;
  int it = 0;
  for (; it < max_iterations; ++it) {
    for (int i = 0; i < n; ++i) z[i] = precond[i] * res[i];
    double rho_new = 0.0;
    for (int i = 0; i < n; ++i) rho_new += res[i] * z[i];
    if (it == 0) {
      p = z;
    } else {
      const double beta = rho_new / rho;
      for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    rho = rho_new;
    std::fill(tmp.begin(), tmp.end(), 0.0);
    J.RightMultiplyAndAccumulate(p.data(), tmp.data());
    for (int i = 0; i < n; ++i) Ap[i] = D2[i] * p[i];
    J.LeftMultiplyAndAccumulate(tmp.data(), Ap.data());
    double pAp = 0.0;
    for (int i = 0; i < n; ++i) pAp += p[i] * Ap[i];
    if (!(pAp > 0.0)) break;
    const double alpha = rho / pAp;
    double norm_res = 0.0;
    for (int i = 0; i < n; ++i) {
      (*y)[i] += alpha * p[i];
      res[i] -= alpha * Ap[i];
      norm_res += res[i] * res[i];
    }
    if (std::sqrt(norm_res) <= tolerance * norm_b) { ++it; break; }
  }
  return it;
}

}  // namespace

bool Solver::Options::IsValid(std::string* error) const {
  // solver.cc:702-708: the CUDA evaluator supports TRUST_REGION only.
  if (use_cuda_for_evaluator && minimizer_type != TRUST_REGION) {
    *error = "Using CUDA for cost function evaluation is currently only supported for "
             "TRUST_REGION minimizers.";
    return false;
  }
  if (minimizer_type != TRUST_REGION) {
    *error = "Only TRUST_REGION minimization is provided by this library.";
    return false;
  }
  if (use_cuda_for_evaluator && registered_cuda_evaluators == nullptr) {
    *error = "use_cuda_for_evaluator requires registered_cuda_evaluators (use "
             "ceres::Solve(options, ProblemCUDA*, summary)).";
    return false;
  }
  if (max_num_iterations < 0) { *error = "max_num_iterations < 0"; return false; }
  return true;
}

std::string Solver::Summary::BriefReport() const {
  char buf[256];
  std::snprintf(buf, sizeof(buf), "Ceres Solver Report: Iterations: %d, Initial cost: %e, "
                "Final cost: %e, Termination: %s",
                num_successful_steps + num_unsuccessful_steps, initial_cost, final_cost,
                termination_type == CONVERGENCE ? "CONVERGENCE"
                : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE");
  return buf;
}

std::string Solver::Summary::FullReport() const {
  char buf[2048];
  std::snprintf(
      buf, sizeof(buf),
      "\nSolver Summary (B200 evaluation engine)\n\n"
      "                                     Original                  Reduced\n"
      "Parameter blocks                 %12d             %12d\n"
      "Parameters                       %12d             %12d\n"
      "Effective parameters             %12d             %12d\n"
      "Residual blocks                  %12d             %12d\n"
      "Residuals                        %12d             %12d\n\n"
      "Cost:\nInitial                          %e\nFinal                            %e\n"
      "Change                           %e\n\n"
      "Minimizer iterations             %12d\nSuccessful steps                 %12d\n"
      "Unsuccessful steps               %12d\n\n"
      "Time (in seconds):\nPreprocessor                     %12.6f\n\n"
      "  Residual only evaluation       %12.6f (%d)\n"
      "  Jacobian & residual evaluation %12.6f (%d)\n"
      "  Linear solver                  %12.6f\n"
      "Minimizer                        %12.6f\n\nTotal                            %12.6f\n\n"
      "Termination:                     %s (%s)\n",
      num_parameter_blocks, num_parameter_blocks_reduced, num_parameters, num_parameters_reduced,
      num_effective_parameters, num_effective_parameters_reduced, num_residual_blocks,
      num_residual_blocks_reduced, num_residuals, num_residuals_reduced, initial_cost, final_cost,
      initial_cost - final_cost, num_successful_steps + num_unsuccessful_steps,
      num_successful_steps, num_unsuccessful_steps, preprocessor_time_in_seconds,
      residual_evaluation_time_in_seconds, num_residual_evaluations,
      jacobian_evaluation_time_in_seconds, num_jacobian_evaluations,
      linear_solver_time_in_seconds, minimizer_time_in_seconds, total_time_in_seconds,
      termination_type == CONVERGENCE ? "CONVERGENCE"
      : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE",
      message.c_str());
  return buf;
}

// The trust-region loop with state, step, gradient and diagonals resident in HBM
// (cb200_engine_trust_region_step and friends, include/ceres_b200.h): per iteration the host
// reads a handful of scalars.  Same decisions as the host loop below
// (trust_region_minimizer.cc step acceptance, levenberg_marquardt_strategy.cc radius update).
// Returns false when the program needs the host loop (bounds, a manifold only the host knows,
// an EvaluationCallback that must see the parameter blocks).
namespace {
bool DeviceLoopPossible(const Solver::Options& options, const internal::Program& program,
                        const internal::ProblemImpl& problem) {
  if (std::getenv("CB200_HOST_TRUST_REGION")) return false;
  if (problem.options().evaluation_callback != nullptr) return false;
  for (const internal::ParameterBlock* pb : program.parameter_blocks()) {
    if (pb->lower_bounds || pb->upper_bounds) return false;
    int kind = 0, param = 0;
    if (pb->manifold && !pb->manifold->DeviceDescription(&kind, &param)) return false;
  }
  return true;
}

void MinimizeOnDevice(const Solver::Options& options, cb200_engine* engine, double fixed_cost,
                      double start, double* x, Solver::Summary* summary) {
  auto fail = [&](const char* what) {
    summary->termination_type = FAILURE;
    summary->message = std::string(what) + ": " + cb200_engine_last_error(engine);
  };
  const uint32_t kLoss = CB200_APPLY_LOSS_FUNCTION;
  double evaluation_seconds[2] = {0.0, 0.0};
  int evaluation_calls[2] = {0, 0};
  auto evaluate = [&](int which, bool jacobian, double* cost) {
    const double t0 = Seconds();
    const int rc = cb200_engine_evaluate_state(
        engine, which, kLoss | CB200_KEEP_JACOBIAN_ON_DEVICE | CB200_KEEP_RESIDUALS_ON_DEVICE,
        jacobian, jacobian, jacobian, cost);
    evaluation_seconds[jacobian] += Seconds() - t0;
    ++evaluation_calls[jacobian];
    return rc;
  };
  double cost = 0.0;
  if (cb200_engine_state_upload(engine, x) != CB200_OK) return fail("cb200_engine_state_upload");
  int rc = evaluate(0, true, &cost);
  if (rc != CB200_OK) {
    summary->message = "Initial residual and Jacobian evaluation failed.";
    if (rc < 0) fail("cb200_engine_evaluate_state");
    return;
  }
  summary->initial_cost = cost + fixed_cost;
  if (options.jacobi_scaling && cb200_engine_jacobi_scale(engine, 1) != CB200_OK)
    return fail("cb200_engine_jacobi_scale");
  double radius = options.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  summary->termination_type = NO_CONVERGENCE;
  summary->message = "Maximum number of iterations reached.";
  if (options.minimizer_progress_to_stdout)
    std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter\n"
                "%4d  %.6e    %.2e   %.2e   %.2e  %.2e  %.2e   %5d\n",
                0, cost + fixed_cost, 0.0, 0.0, 0.0, 0.0, radius, 0);
  for (int it = 1; it <= options.max_num_iterations; ++it) {
    if (Seconds() - start > options.max_solver_time_in_seconds) {
      summary->message = "Maximum solver time reached.";
      break;
    }
    cb200_step_options so{};
    so.radius = radius;
    so.min_lm_diagonal = options.min_lm_diagonal;
    so.max_lm_diagonal = options.max_lm_diagonal;
    so.reuse_diagonal = reuse_diagonal;
    so.cg.min_num_iterations = options.min_linear_solver_iterations;
    so.cg.max_num_iterations = options.max_linear_solver_iterations;
    so.cg.r_tolerance = -1.0;
    so.cg.q_tolerance = options.eta;  // levenberg_marquardt_strategy.cc:97-118
    cb200_step_summary ss;
    const double ls_start = Seconds();
    if (cb200_engine_trust_region_step(engine, &so, &ss) != CB200_OK) {
      fail("cb200_engine_trust_region_step");
      break;
    }
    summary->linear_solver_time_in_seconds += Seconds() - ls_start;
    if (std::getenv("CB200_SOLVER_TIMING"))
      std::fprintf(stderr, "iteration %d: trust-region step %.1f ms wall, conjugate gradients %.1f ms "
                   "on the device (%d iterations)\n", it, (Seconds() - ls_start) * 1e3,
                   ss.cg.solve_ms, ss.cg.num_iterations);
    Solver::IterationSummary is;
    is.iteration = it;
    is.linear_solver_iterations = ss.cg.num_iterations;
    is.step_norm = ss.step_norm;
    double new_cost = 0.0;
    bool ok = ss.cg.termination != 2 && ss.model_cost_change > 0.0 && ss.plus_ok;
    if (ok) {
      rc = evaluate(1, false, &new_cost);
      if (rc < 0) { fail("cb200_engine_evaluate_state"); break; }
      ok = rc == CB200_OK;
    }
    const double relative_decrease = ok ? (cost - new_cost) / ss.model_cost_change : -1.0;
    is.relative_decrease = relative_decrease;
    if (ok && relative_decrease > options.min_relative_decrease) {
      is.step_is_successful = true;
      is.cost_change = cost - new_cost;
      cb200_engine_accept_candidate(engine);
      rc = evaluate(0, true, &cost);
      if (rc != CB200_OK) {
        summary->termination_type = FAILURE;
        summary->message = "Residual and Jacobian evaluation failed.";
        break;
      }
      if (options.jacobi_scaling && cb200_engine_jacobi_scale(engine, 0) != CB200_OK) {
        fail("cb200_engine_jacobi_scale");
        break;
      }
      radius = std::min(options.max_trust_region_radius,
                        radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3)));
      decrease_factor = 2.0;
      reuse_diagonal = false;
      ++summary->num_successful_steps;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      ++summary->num_unsuccessful_steps;
    }
    double gmax = 0.0;
    cb200_engine_gradient_max_norm(engine, &gmax);
    is.cost = cost + fixed_cost;
    is.gradient_max_norm = gmax;
    is.trust_region_radius = radius;
    summary->iterations.push_back(is);
    if (options.minimizer_progress_to_stdout)
      std::printf("%4d  %.6e    %.2e   %.2e   %.2e  %.2e  %.2e   %5d\n", it, is.cost, is.cost_change,
                  gmax, ss.step_norm, relative_decrease, radius, ss.cg.num_iterations);
    if (is.step_is_successful) {
      if (std::fabs(is.cost_change) <= options.function_tolerance * cost) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Function tolerance reached.";
        break;
      }
      if (gmax <= options.gradient_tolerance) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Gradient tolerance reached.";
        break;
      }
      if (ss.step_norm <= options.parameter_tolerance * (ss.state_norm + options.parameter_tolerance)) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Parameter tolerance reached.";
        break;
      }
    } else if (radius < options.min_trust_region_radius) {
      summary->termination_type = CONVERGENCE;
      summary->message = "Minimum trust region radius reached.";
      break;
    }
  }
  cb200_engine_state_download(engine, 0, x);
  summary->final_cost = cost + fixed_cost;
  summary->residual_evaluation_time_in_seconds = evaluation_seconds[0];
  summary->num_residual_evaluations = evaluation_calls[0];
  summary->jacobian_evaluation_time_in_seconds = evaluation_seconds[1];
  summary->num_jacobian_evaluations = evaluation_calls[1];
}
}  // namespace

namespace internal {
// The parameter blocks that only ever appear as argument j of their residual blocks form an
// independent set: every residual block holds exactly one of them (the points of a bundle
// adjustment problem for j = 1).  Puts them in group 0 and everything else in group 1 for the
// argument slot with the most such blocks; false when no slot covers half of the variable
// blocks (a small first group is not worth its own region of the Jacobian).
bool ArgumentSlotOrdering(const ProblemImpl& problem, ParameterBlockOrdering* ordering) {
  const auto& pbs = problem.parameter_blocks();
  std::vector<uint32_t> slots(pbs.size(), 0u);
  int common_slots = CB200_MAX_PARAMETER_BLOCKS;
  for (const ResidualTypeStore& t : problem.types()) {
    const int nb = t.desc.num_parameter_blocks;
    if (t.size() == 0) continue;
    common_slots = std::min(common_slots, nb);
    for (size_t k = 0; k < t.parameter_blocks.size(); ++k)
      slots[t.parameter_blocks[k]] |= 1u << (k % nb);
  }
  int best_slot = -1;
  size_t best_count = 0, variable = 0;
  for (size_t i = 0; i < pbs.size(); ++i) variable += !pbs[i]->IsConstant();
  for (int j = 0; j < common_slots; ++j) {
    size_t count = 0;
    for (size_t i = 0; i < pbs.size(); ++i) count += slots[i] == (1u << j) && !pbs[i]->IsConstant();
    if (count > best_count) { best_count = count; best_slot = j; }
  }
  if (best_slot < 0 || 2 * best_count < variable) return false;
  for (size_t i = 0; i < pbs.size(); ++i)
    ordering->AddElementToGroup(pbs[i]->user_state, slots[i] == (1u << best_slot) ? 0 : 1);
  return true;
}
}  // namespace internal

void Solver::Solve(const Options& options, Problem* problem, Summary* summary) {
  const double start = Seconds();
  *summary = Summary();
  if (!options.IsValid(&summary->message)) {
    summary->termination_type = FAILURE;
    return;
  }
  if (!options.use_cuda_for_evaluator) {
    summary->message = "This library only provides the CUDA evaluator: solve a ProblemCUDA "
                       "with ceres::Solve(options, ProblemCUDA*, summary).";
    return;
  }
  internal::ProblemImpl* impl = problem->mutable_impl();
  summary->num_parameter_blocks = impl->NumParameterBlocks();
  summary->num_parameters = impl->NumParameters();
  summary->num_residual_blocks = impl->NumResidualBlocks();
  summary->num_residuals = impl->NumResiduals();
  for (const auto& pb : impl->parameter_blocks())
    summary->num_effective_parameters += pb->TangentSize();

  // ---- preprocessing (trust_region_preprocessor.cc:373-407)
  internal::Program full(impl);
  std::vector<double*> removed;
  double fixed_cost = 0.0;
  std::unique_ptr<internal::Program> program =
      full.CreateReducedProgram(&removed, &fixed_cost, &summary->message);
  if (!program) return;
  summary->fixed_cost = fixed_cost;
  if (program->NumParameterBlocks() == 0) {
    summary->termination_type = CONVERGENCE;
    summary->message = "Function tolerance reached. No non-constant parameter blocks found.";
    summary->initial_cost = summary->final_cost = fixed_cost;
    return;
  }
  const bool schur = options.linear_solver_type == DENSE_SCHUR ||
                     options.linear_solver_type == SPARSE_SCHUR ||
                     options.linear_solver_type == ITERATIVE_SCHUR;
  // CGNR with CUDA_SPARSE keeps the Jacobian in HBM (below): its layout is private to the
  // device, so a two-group ordering is applied there too.  With the first group's cells in one
  // region and the rest in another, the 32 cells a warp writes per argument are one contiguous
  // run, which is what the bulk stores of the evaluation kernel and the per-type J'(J p) kernel
  // need; in the row-by-row layout of num_eliminate_blocks = 0 both fall back to their
  // per-thread paths (measured on BAL L: evaluation 2.63 instead of 1.83 ms, 11.7 instead of
  // 1.43 ms per conjugate-gradient product; profiles/r2_solve_L_launches_rowwise_layout.csv).
  const bool resident_cgnr = options.linear_solver_type == CGNR &&
                             options.sparse_linear_algebra_library_type == CUDA_SPARSE;
  const ParameterBlockOrdering* ordering = options.linear_solver_ordering.get();
  // No ordering given for the resident layout: the argument-slot independent set (above)
  // becomes the first group.
  ParameterBlockOrdering slot_ordering;
  if (resident_cgnr && !ordering && internal::ArgumentSlotOrdering(*impl, &slot_ordering))
    ordering = &slot_ordering;
  int num_eliminate_blocks = 0;
  if ((schur || resident_cgnr) && ordering) {
    // ApplyOrdering + size of the first elimination group (reorder_program.cc:469-560).
    program->ReorderParameterBlocksByGroup(ordering->element_to_group());
    const int first = ordering->MinGroup();
    for (const internal::ParameterBlock* pb : program->parameter_blocks())
      num_eliminate_blocks += ordering->GroupId(pb->user_state) == first;
    if (num_eliminate_blocks > 0 && num_eliminate_blocks < program->NumParameterBlocks())
      program->LexicographicallyOrderResidualBlocks(num_eliminate_blocks);
    else
      num_eliminate_blocks = 0;
  }
  internal::Evaluator::Options eo;
  eo.num_threads = options.num_threads;
  eo.num_eliminate_blocks = num_eliminate_blocks;
  eo.linear_solver_type = options.linear_solver_type;
  eo.sparse_linear_algebra_library_type = options.sparse_linear_algebra_library_type;
  eo.use_cuda = true;
  eo.registered_cuda_evaluators = options.registered_cuda_evaluators;
  eo.evaluation_callback = impl->options().evaluation_callback;
  eo.device = options.cuda_device;
  // CGNR with CUDA_SPARSE is the reference's device-resident combination
  // (cgnr_solver.cc CudaCgnrSolver): the Jacobian stays in HBM and conjugate gradients on
  // the normal equations run there (cb200_engine_cgnr_solve).
  eo.jacobian_on_device = options.linear_solver_type == CGNR &&
                          options.sparse_linear_algebra_library_type == CUDA_SPARSE;
  std::unique_ptr<internal::Evaluator> evaluator =
      internal::Evaluator::Create(eo, program.get(), &summary->message);
  if (!evaluator) return;
  std::unique_ptr<internal::SparseMatrix> jacobian = evaluator->CreateJacobian();
  summary->num_parameter_blocks_reduced = program->NumParameterBlocks();
  summary->num_parameters_reduced = program->NumParameters();
  summary->num_effective_parameters_reduced = program->NumEffectiveParameters();
  summary->num_residual_blocks_reduced = program->NumResidualBlocks();
  summary->num_residuals_reduced = program->NumResiduals();
  summary->preprocessor_time_in_seconds = Seconds() - start;

  // ---- trust-region loop
  const double minimizer_start = Seconds();
  const int n = program->NumParameters(), ne = program->NumEffectiveParameters();
  const int m = program->NumResiduals();
  const int nt = std::max(1, options.num_threads);
  Buckets buckets;
  PinnedVector x(n);
  program->ParameterBlocksToStateVector(x.data());
  auto* resident = dynamic_cast<internal::DeviceResidentJacobian*>(jacobian.get());
  if (resident && DeviceLoopPossible(options, *program, *impl)) {
    // (the loop on the device needs the state on the host twice - up, and down at the end -
    // and none of the host loop's work vectors: page-locking those cost 0.5 s on BAL L)
    MinimizeOnDevice(options, resident->engine(), fixed_cost, start, x.data(), summary);
    program->StateVectorToParameterBlocks(x.data());  // results back into the user's arrays
    summary->minimizer_time_in_seconds = Seconds() - minimizer_start;
    summary->total_time_in_seconds = Seconds() - start;
    return;
  }
  PinnedVector x_plus(n), gradient(ne), scale(ne, 1.0), diagonal(ne), D2(ne), y(ne);
  std::vector<double> delta(ne);
  // only the host linear solver reads the residuals and the model on the host
  std::vector<double> residuals, model;
  // with a device-resident Jacobian the residuals stay on the device as well
  if (!resident) {
    residuals.resize(m);
    model.resize(m);
  }
  double* const residuals_out = resident ? nullptr : residuals.data();
  double cost = 0.0;
  if (!evaluator->Evaluate(x.data(), &cost, residuals_out, gradient.data(), jacobian.get())) {
    summary->message = "Initial residual and Jacobian evaluation failed.";
    return;
  }
  summary->initial_cost = cost + fixed_cost;
  if (options.jacobi_scaling) {
    jacobian->SquaredColumnNorm(scale.data());
    ForEach(nt, ne, [&](int64_t i) { scale[i] = 1.0 / (1.0 + std::sqrt(scale[i])); });
    jacobian->ScaleColumns(scale.data());
  }
  double radius = options.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  summary->termination_type = NO_CONVERGENCE;
  summary->message = "Maximum number of iterations reached.";
  if (options.minimizer_progress_to_stdout)
    std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter\n"
                "%4d  %.6e    %.2e   %.2e   %.2e  %.2e  %.2e   %5d\n",
                0, cost + fixed_cost, 0.0, 0.0, 0.0, 0.0, radius, 0);
  for (int it = 1; it <= options.max_num_iterations; ++it) {
    if (Seconds() - start > options.max_solver_time_in_seconds) {
      summary->message = "Maximum solver time reached.";
      break;
    }
    double mark = Seconds();
    auto lap = [&](double* bucket) {
      const double now = Seconds();
      *bucket += now - mark;
      mark = now;
    };
    if (!reuse_diagonal) {
      jacobian->SquaredColumnNorm(diagonal.data());
      ForEach(nt, ne, [&](int64_t i) {
        diagonal[i] =
            std::min(std::max(diagonal[i], options.min_lm_diagonal), options.max_lm_diagonal);
      });
    }
    ForEach(nt, ne, [&](int64_t i) { D2[i] = diagonal[i] / radius; });
    lap(&buckets.diagonal);
    const double ls_start = Seconds();
    int ls_iterations = 0;
    double model_cost_change = 0.0, step_norm = 0.0, x_norm = 0.0;
    bool linear_solver_ok = true;
    if (resident) {
      // levenberg_marquardt_strategy.cc:97-118: inexact step, stopped by the relative
      // decrease of the quadratic model (eta)
      cb200_cgnr_options cg_options;
      cg_options.min_num_iterations = options.min_linear_solver_iterations;
      cg_options.max_num_iterations = options.max_linear_solver_iterations;
      cg_options.r_tolerance = -1.0;
      cg_options.q_tolerance = options.eta;
      cb200_cgnr_summary cg;
      y.assign(ne, 0.0);
      if (cb200_engine_cgnr_solve(resident->engine(), D2.data(), &cg_options, y.data(), &cg) !=
          CB200_OK) {
        summary->termination_type = FAILURE;
        summary->message = std::string("cb200_engine_cgnr_solve: ") +
                           cb200_engine_last_error(resident->engine());
        break;
      }
      ls_iterations = cg.num_iterations;
      linear_solver_ok = cg.termination != 2;
      // step = -y: model cost change = (J y).r - |J y|^2 / 2
      model_cost_change = cg.jy_dot_b - 0.5 * cg.jy_squared_norm;
      ForEach(nt, ne, [&](int64_t i) { y[i] = -y[i]; });
    } else {
      ls_iterations = SolveNormalEquations(*jacobian, D2.data(), residuals.data(),
                                           options.max_linear_solver_iterations,
                                           1e-2 * options.eta, &y);
      // step = -y (scaled space); model cost change = -m'(r + m/2), m = J step
      std::fill(model.begin(), model.end(), 0.0);
      for (int i = 0; i < ne; ++i) y[i] = -y[i];
      jacobian->RightMultiplyAndAccumulate(y.data(), model.data());
      for (int i = 0; i < m; ++i) model_cost_change -= model[i] * (residuals[i] + 0.5 * model[i]);
    }
    summary->linear_solver_time_in_seconds += Seconds() - ls_start;
    mark = Seconds();
    ForEach(nt, ne, [&](int64_t i) { delta[i] = y[i] * scale[i]; });
    step_norm = Reduce(nt, ne, false, [&](int64_t i) { return delta[i] * delta[i]; });
    x_norm = Reduce(nt, n, false, [&](int64_t i) { return x[i] * x[i]; });
    step_norm = std::sqrt(step_norm);
    IterationSummary is;
    is.iteration = it;
    is.linear_solver_iterations = ls_iterations;
    is.step_norm = step_norm;
    double new_cost = 0.0;
    lap(&buckets.step);
    bool ok = linear_solver_ok && model_cost_change > 0.0 &&
              evaluator->Plus(x.data(), delta.data(), x_plus.data());
    lap(&buckets.plus);
    ok = ok && evaluator->Evaluate(x_plus.data(), &new_cost, nullptr, nullptr, nullptr);
    mark = Seconds();
    const double relative_decrease = ok ? (cost - new_cost) / model_cost_change : -1.0;
    is.relative_decrease = relative_decrease;
    if (ok && relative_decrease > options.min_relative_decrease) {
      is.step_is_successful = true;
      is.cost_change = cost - new_cost;
      ForEach(nt, n, [&](int64_t i) { x[i] = x_plus[i]; });
      if (!evaluator->Evaluate(x.data(), &cost, residuals_out, gradient.data(),
                               jacobian.get())) {
        summary->termination_type = FAILURE;
        summary->message = "Residual and Jacobian evaluation failed.";
        break;
      }
      mark = Seconds();
      if (options.jacobi_scaling) jacobian->ScaleColumns(scale.data());
      lap(&buckets.scaling);
      radius = std::min(options.max_trust_region_radius,
                        radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3)));
      decrease_factor = 2.0;
      reuse_diagonal = false;
      ++summary->num_successful_steps;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      ++summary->num_unsuccessful_steps;
    }
    mark = Seconds();
    const double gmax =
        Reduce(nt, ne, true, [&](int64_t i) { return std::fabs(gradient[i]); });
    lap(&buckets.bookkeeping);
    is.cost = cost + fixed_cost;
    is.gradient_max_norm = gmax;
    is.trust_region_radius = radius;
    summary->iterations.push_back(is);
    if (options.minimizer_progress_to_stdout)
      std::printf("%4d  %.6e    %.2e   %.2e   %.2e  %.2e  %.2e   %5d\n", it, is.cost, is.cost_change,
                  gmax, step_norm, relative_decrease, radius, ls_iterations);
    if (is.step_is_successful) {
      if (std::fabs(is.cost_change) <= options.function_tolerance * cost) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Function tolerance reached.";
        break;
      }
      if (gmax <= options.gradient_tolerance) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Gradient tolerance reached.";
        break;
      }
      if (step_norm <= options.parameter_tolerance * (std::sqrt(x_norm) + options.parameter_tolerance)) {
        summary->termination_type = CONVERGENCE;
        summary->message = "Parameter tolerance reached.";
        break;
      }
    } else if (radius < options.min_trust_region_radius) {
      summary->termination_type = CONVERGENCE;
      summary->message = "Minimum trust region radius reached.";
      break;
    }
  }
  program->StateVectorToParameterBlocks(x.data());  // results back into the user's arrays
  summary->final_cost = cost + fixed_cost;
  for (const auto& kv : evaluator->Statistics()) {
    if (kv.first == "Evaluator::Residual") {
      summary->residual_evaluation_time_in_seconds = kv.second.time;
      summary->num_residual_evaluations = kv.second.calls;
    } else if (kv.first == "Evaluator::Jacobian") {
      summary->jacobian_evaluation_time_in_seconds = kv.second.time;
      summary->num_jacobian_evaluations = kv.second.calls;
    }
  }
  summary->minimizer_time_in_seconds = Seconds() - minimizer_start;
  if (std::getenv("CB200_SOLVER_TIMING"))
    std::fprintf(stderr,
                 "solver buckets (s): LM diagonal %.3f, Jacobi scaling %.3f, step vectors %.3f, "
                 "Plus %.3f, gradient norm %.3f\n",
                 buckets.diagonal, buckets.scaling, buckets.step, buckets.plus,
                 buckets.bookkeeping);
  summary->total_time_in_seconds = Seconds() - start;
}

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  nvtxRangePushA("ceres::Solve");
  Solver solver;
  solver.Solve(options, problem, summary);
  nvtxRangePop();
}

}  // namespace ceres
