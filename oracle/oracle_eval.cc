// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, link or call anything under oracle/.
//
// CPU restatement of the reference's residual/Jacobian evaluation path
// (the CPU ProgramEvaluator, which the fork's GPU evaluator is tested against).
//
// Pinning status: the arithmetic (Jet / rotation / functors / Corrector / loss)
// is validated against the reference's own headers compiled over an Eigen shim
// (oracle/_ref, see oracle/Makefile and tests/test_oracle_vs_ref.py) and the
// layout/evaluator logic against the known-answer tables of
// internal/ceres/evaluator_test.cc (tests/test_oracle_known_answers.py).
//
// Reference files followed (all paths relative to /root/reference):
//   include/ceres/internal/autodiff.h:318-381 ........ AutoDifferentiate
//   include/ceres/autodiff_cost_function.h ............ Evaluate dispatch
//   include/ceres/loss_function_cuda.h:62-149 ......... Trivial/Huber/Cauchy/Scaled
//   include/ceres/internal/corrector.h:82-213 ......... Corrector
//   internal/ceres/manifold.cc:28-79,199-214 .......... Quaternion / Subset PlusJacobian
//   include/ceres/product_manifold.h:117-124,200-218 .. block-diagonal product
//   internal/ceres/parameter_block.h:120,152-156 ...... IsConstant / TangentSize
//   internal/ceres/program.cc:80-103,152-177,306-430 .. state packing, offsets, RemoveFixedBlocks
//   internal/ceres/reorder_program.cc:254-335 ......... LexicographicallyOrderResidualBlocks
//   internal/ceres/residual_block.cc:68-204 ........... ResidualBlock::Evaluate
//   internal/ceres/program_evaluator.h:134-283,336-347 ProgramEvaluator::Evaluate
//   internal/ceres/block_jacobian_writer.cc:62-150,192-250
//   internal/ceres/compressed_row_jacobian_writer.cc:71-193,195-300
//   internal/ceres/array_utils.cc:40-66 ............... IsArrayValid
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <thread>
#include <utility>
#include <vector>

#include "oracle_functors.h"
#include "oracle_jet.h"

namespace oracle {

// ----------------------------------------------------------------- autodiff
// include/ceres/internal/autodiff.h:318-381.  jacobians == nullptr selects the
// plain-double path of AutoDiffCostFunction::Evaluate
// (include/ceres/autodiff_cost_function.h "if (!jacobians) VariadicEvaluate").
using EvalFn = bool (*)(const double* fdata, const double* const* params,
                        double* residuals, double** jacobians);

template <typename F, typename T, std::size_t... Is>
inline bool CallFunctor(const double* fdata, T* const* p, T* out,
                        std::index_sequence<Is...>) {
  return F()(fdata, static_cast<const T*>(p[Is])..., out);
}

template <typename F, int kRes, int... Ns>
struct AutoDiff {
  static constexpr int kNumBlocks = sizeof...(Ns);
  static constexpr int kNumParams = (Ns + ...);
  static bool Evaluate(const double* fdata, const double* const* params,
                       double* residuals, double** jacobians) {
    constexpr int sizes[kNumBlocks] = {Ns...};
    if (jacobians == nullptr) {
      double* p[kNumBlocks];
      for (int k = 0; k < kNumBlocks; ++k) p[k] = const_cast<double*>(params[k]);
      return CallFunctor<F, double>(fdata, p, residuals,
                                    std::make_index_sequence<kNumBlocks>{});
    }
    using JetT = Jet<kNumParams>;
    JetT x[kNumParams];
    JetT* unpacked[kNumBlocks];
    int offset = 0;
    for (int k = 0; k < kNumBlocks; ++k) {
      unpacked[k] = x + offset;
      // autodiff.h:186-204 Make1stOrderPerturbation
      for (int j = 0; j < sizes[k]; ++j) x[offset + j] = JetT(params[k][j], offset + j);
      offset += sizes[k];
    }
    JetT output[kRes];
    // autodiff.h:358-363: invalidate the outputs.
    for (int i = 0; i < kRes; ++i) {
      output[i].a = kImpossibleValue;
      for (int j = 0; j < kNumParams; ++j) output[i].v[j] = kImpossibleValue;
    }
    if (!CallFunctor<F, JetT>(fdata, unpacked, output,
                              std::make_index_sequence<kNumBlocks>{})) {
      return false;
    }
    // autodiff.h:245-268 Take0thOrderPart / Take1stOrderPart
    for (int i = 0; i < kRes; ++i) residuals[i] = output[i].a;
    offset = 0;
    for (int k = 0; k < kNumBlocks; ++k) {
      if (jacobians[k] != nullptr) {
        for (int i = 0; i < kRes; ++i)
          for (int j = 0; j < sizes[k]; ++j)
            jacobians[k][i * sizes[k] + j] = output[i].v[offset + j];
      }
      offset += sizes[k];
    }
    return true;
  }
};

struct CostTypeInfo {
  int num_residuals;
  int num_blocks;
  int sizes[10];
  int fdata_len;
  EvalFn eval;
};

template <typename F, int kRes, int... Ns>
CostTypeInfo MakeType(int fdata_len) {
  CostTypeInfo t{};
  t.num_residuals = kRes;
  t.num_blocks = sizeof...(Ns);
  const int s[] = {Ns...};
  for (int i = 0; i < t.num_blocks; ++i) t.sizes[i] = s[i];
  t.fdata_len = fdata_len;
  t.eval = &AutoDiff<F, kRes, Ns...>::Evaluate;
  return t;
}

// Cost-type ids: shared numbering with the product's test driver
// (tests/driver/cost_types.h) and the python generators.
static const std::vector<CostTypeInfo>& CostTypes() {
  static const std::vector<CostTypeInfo> types = {
      /* 0 */ MakeType<SnavelyReprojectionError, 2, 9, 3>(2),
      /* 1 */ MakeType<SnavelyReprojectionErrorWithQuaternions, 2, 10, 3>(2),
      /* 2 */ MakeType<SnavelyReprojectionErrorNoRadialDistortion, 2, 7, 3>(2),
      /* 3 */ MakeType<PointDisplacementError, 3, 3>(3),
      /* 4 */ MakeType<RelativePoseError, 6, 7, 7>(7),
      /* 5 */ MakeType<BinaryScalarCost, 1, 2, 2>(1),
      /* 6 */ MakeType<TenParameterCost, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1>(0),
      /* 7 */ MakeType<OnlyFillsOneOutputFunctor, 2, 1>(0),
      /* 8 */ MakeType<AffineTestCost<1, 3, true, 2, 3, 4>, 3, 2, 3, 4>(0),
      /* 9 */ MakeType<AffineTestCost<1, 3, true, 4, 3, 2>, 3, 4, 3, 2>(0),
      /* 10 */ MakeType<AffineTestCost<1, 2, true, 2, 3>, 2, 2, 3>(0),
      /* 11 */ MakeType<AffineTestCost<2, 3, true, 2, 4>, 3, 2, 4>(0),
      /* 12 */ MakeType<AffineTestCost<3, 4, true, 3, 4>, 4, 3, 4>(0),
      /* 13 */ MakeType<AffineTestCost<20, 3, false, 2, 3, 4>, 3, 2, 3, 4>(0),
      /* 14 */ MakeType<ParameterSensitiveCost, 2, 2>(0),
      /* 15 */ MakeType<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>(43),
      /* 16 */ MakeType<JetBatteryCost, 40, 2>(0),
      /* 17 */ MakeType<SqrtOfConstantCost, 1, 1>(1),
  };
  return types;
}

// --------------------------------------------------------------------- loss
// include/ceres/loss_function_cuda.h:62-149 (== internal/ceres/loss_function.cc:44-82)
enum LossKind {
  kLossNone = 0,  // nullptr loss
  kLossTrivial = 1,
  kLossHuber = 2,
  kLossCauchy = 3,
  kLossScaledHuber = 4,   // ScaledLossCUDA<HuberLossCUDA>(Huber(a), b)
  kLossScaledCauchy = 5,  // ScaledLossCUDA<CauchyLossCUDA>(Cauchy(a), b)
  kLossScaledTrivial = 6,  // ScaledLossCUDA<TrivialLossCUDA>(_, b)
  kLossConvexTest = 7      // test loss rho(s) = s + a s^2: rho'' = 2a > 0 (Corrector alpha path)
};

static void HuberEvaluate(double a_, double s, double rho[3]) {
  const double b_ = a_ * a_;
  if (s > b_) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a_ * r - b_;
    rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s;
    rho[1] = 1.0;
    rho[2] = 0.0;
  }
}

static void CauchyEvaluate(double a, double s, double rho[3]) {
  const double b_ = a * a;
  const double c_ = 1 / b_;
  const double sum = 1.0 + s * c_;
  const double inv = 1.0 / sum;
  rho[0] = b_ * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c_ * (inv * inv);
}

static void LossEvaluate(int kind, double a, double b, double s, double rho[3]) {
  switch (kind) {
    case kLossTrivial:
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
      return;
    case kLossHuber:
      HuberEvaluate(a, s, rho);
      return;
    case kLossCauchy:
      CauchyEvaluate(a, s, rho);
      return;
    case kLossScaledHuber:
      HuberEvaluate(a, s, rho);
      rho[0] *= b; rho[1] *= b; rho[2] *= b;
      return;
    case kLossScaledCauchy:
      CauchyEvaluate(a, s, rho);
      rho[0] *= b; rho[1] *= b; rho[2] *= b;
      return;
    case kLossScaledTrivial:
      rho[0] = b * s; rho[1] = b; rho[2] = 0.0;
      return;
    case kLossConvexTest:
      rho[0] = s + a * s * s; rho[1] = 1.0 + 2.0 * a * s; rho[2] = 2.0 * a;
      return;
    default:
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

// ---------------------------------------------------------------- corrector
// include/ceres/internal/corrector.h:82-147 ctor, :159-166, :174-213
struct Corrector {
  double sqrt_rho1_, residual_scaling_, alpha_sq_norm_;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1_ = std::sqrt(rho[1]);
    if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
      residual_scaling_ = sqrt_rho1_;
      alpha_sq_norm_ = 0.0;
      return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling_ = sqrt_rho1_ / (1 - alpha);
    alpha_sq_norm_ = alpha / sq_norm;
  }
  void CorrectResiduals(int num_rows, double* residuals) const {
    for (int i = 0; i < num_rows; ++i) residuals[i] *= residual_scaling_;
  }
  void CorrectJacobian(int num_rows, int num_cols, const double* residuals,
                       double* jacobian) const {
    if (alpha_sq_norm_ == 0.0) {
      for (int i = 0; i < num_rows * num_cols; ++i) jacobian[i] *= sqrt_rho1_;
      return;
    }
    for (int c = 0; c < num_cols; ++c) {
      double r_transpose_j = 0.0;
      for (int r = 0; r < num_rows; ++r)
        r_transpose_j += jacobian[r * num_cols + c] * residuals[r];
      for (int r = 0; r < num_rows; ++r)
        jacobian[r * num_cols + c] =
            sqrt_rho1_ * (jacobian[r * num_cols + c] -
                          alpha_sq_norm_ * residuals[r] * r_transpose_j);
    }
  }
};

// ---------------------------------------------------------------- manifolds
enum ManifoldKind {
  kManifoldNone = 0,
  kManifoldSubset = 1,           // param = bitmask of constant coordinates
  kManifoldQuaternion = 2,       // (w, x, y, z)
  kManifoldEigenQuaternion = 3,  // (x, y, z, w)
  kManifoldQuaternionTimesEuclidean = 4,      // ProductManifold<Quaternion, Euclidean<size-4>>
  kManifoldEigenQuaternionTimesEuclidean = 5,  // ProductManifold<EigenQuaternion, Euclidean<size-4>>
  kManifoldSubsetAlias = 6  // same as kManifoldSubset (the product's tests use it to force its generic path)
};

static int ManifoldTangentSize(int kind, int param, int ambient) {
  switch (kind) {
    case kManifoldNone: return ambient;
    case kManifoldSubset:
    case kManifoldSubsetAlias: return ambient - __builtin_popcount(static_cast<unsigned>(param));
    case kManifoldQuaternion:
    case kManifoldEigenQuaternion: return 3;
    default: return ambient - 1;
  }
}

// internal/ceres/manifold.cc:62-79 QuaternionPlusJacobianImpl<Order>, 4x3 row-major
static void QuaternionPlusJacobian(const double* x, int kW, int kX, int kY, int kZ,
                                   double* J /* 4x3 */) {
  auto at = [&](int r, int c) -> double& { return J[r * 3 + c]; };
  at(kW, 0) = -x[kX]; at(kW, 1) = -x[kY]; at(kW, 2) = -x[kZ];
  at(kX, 0) = x[kW];  at(kX, 1) = x[kZ];  at(kX, 2) = -x[kY];
  at(kY, 0) = -x[kZ]; at(kY, 1) = x[kW];  at(kY, 2) = x[kX];
  at(kZ, 0) = x[kY];  at(kZ, 1) = -x[kX]; at(kZ, 2) = x[kW];
}

// Row-major ambient x tangent.
static void ManifoldPlusJacobian(int kind, int param, int ambient, const double* x,
                                 double* J) {
  const int tangent = ManifoldTangentSize(kind, param, ambient);
  std::fill(J, J + ambient * tangent, 0.0);
  switch (kind) {
    case kManifoldSubsetAlias:
    case kManifoldSubset: {
      // manifold.cc:199-214
      for (int r = 0, c = 0; r < ambient; ++r)
        if (!((param >> r) & 1)) J[r * tangent + c++] = 1.0;
      return;
    }
    case kManifoldQuaternion:
      QuaternionPlusJacobian(x, 0, 1, 2, 3, J);
      return;
    case kManifoldEigenQuaternion:
      QuaternionPlusJacobian(x, 3, 0, 1, 2, J);
      return;
    case kManifoldQuaternionTimesEuclidean:
    case kManifoldEigenQuaternionTimesEuclidean: {
      // product_manifold.h:200-218: block diagonal [4x3 | I].
      double q[12];
      if (kind == kManifoldQuaternionTimesEuclidean) QuaternionPlusJacobian(x, 0, 1, 2, 3, q);
      else QuaternionPlusJacobian(x, 3, 0, 1, 2, q);
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 3; ++c) J[r * tangent + c] = q[r * 3 + c];
      for (int r = 4; r < ambient; ++r) J[r * tangent + (r - 1)] = 1.0;
      return;
    }
    default:
      return;
  }
}

// internal/ceres/manifold.cc:28-58 QuaternionPlusImpl<Order>
static void QuaternionPlus(const double* x, const double* delta, int kW, int kX, int kY,
                           int kZ, double* x_plus_delta) {
  const double norm_delta = std::hypot(delta[0], delta[1], delta[2]);
  if (std::fpclassify(norm_delta) == FP_ZERO) {
    std::copy_n(x, 4, x_plus_delta);
    return;
  }
  const double sin_delta_by_delta = (std::sin(norm_delta) / norm_delta);
  double q_delta[4];
  q_delta[kW] = std::cos(norm_delta);
  q_delta[kX] = sin_delta_by_delta * delta[0];
  q_delta[kY] = sin_delta_by_delta * delta[1];
  q_delta[kZ] = sin_delta_by_delta * delta[2];
  x_plus_delta[kW] = q_delta[kW] * x[kW] - q_delta[kX] * x[kX] -
                     q_delta[kY] * x[kY] - q_delta[kZ] * x[kZ];
  x_plus_delta[kX] = q_delta[kW] * x[kX] + q_delta[kX] * x[kW] +
                     q_delta[kY] * x[kZ] - q_delta[kZ] * x[kY];
  x_plus_delta[kY] = q_delta[kW] * x[kY] - q_delta[kX] * x[kZ] +
                     q_delta[kY] * x[kW] + q_delta[kZ] * x[kX];
  x_plus_delta[kZ] = q_delta[kW] * x[kZ] + q_delta[kX] * x[kY] -
                     q_delta[kY] * x[kX] + q_delta[kZ] * x[kW];
}

static void ManifoldPlus(int kind, int param, int ambient, const double* x,
                         const double* delta, double* out) {
  switch (kind) {
    case kManifoldNone:
      for (int i = 0; i < ambient; ++i) out[i] = x[i] + delta[i];
      return;
    case kManifoldSubsetAlias:
    case kManifoldSubset:
      // manifold.cc:184-197
      for (int i = 0, j = 0; i < ambient; ++i)
        out[i] = ((param >> i) & 1) ? x[i] : x[i] + delta[j++];
      return;
    case kManifoldQuaternion:
      QuaternionPlus(x, delta, 0, 1, 2, 3, out);
      return;
    case kManifoldEigenQuaternion:
      QuaternionPlus(x, delta, 3, 0, 1, 2, out);
      return;
    case kManifoldQuaternionTimesEuclidean:
      QuaternionPlus(x, delta, 0, 1, 2, 3, out);
      for (int i = 4; i < ambient; ++i) out[i] = x[i] + delta[i - 1];
      return;
    case kManifoldEigenQuaternionTimesEuclidean:
      QuaternionPlus(x, delta, 3, 0, 1, 2, out);
      for (int i = 4; i < ambient; ++i) out[i] = x[i] + delta[i - 1];
      return;
  }
}

// internal/ceres/array_utils.cc:40-66
static bool IsArrayValid(int size, const double* x) {
  if (x != nullptr) {
    for (int i = 0; i < size; ++i)
      if (!std::isfinite(x[i]) || (x[i] == kImpossibleValue)) return false;
  }
  return true;
}

// ------------------------------------------------------------------ problem
struct ParameterBlock {
  int size = 0;
  int tangent_size = 0;
  bool set_constant = false;
  int manifold_kind = 0;
  int manifold_param = 0;
  int64_t user_offset = 0;  // into Problem::user_values
  // Program bookkeeping (program.cc:152-177)
  int index = -1;
  int state_offset = -1;
  int delta_offset = -1;
  bool in_constant_list = false;
  // parameter_block.h:120: IsConstant = set constant or zero-dimensional tangent.
  bool IsConstant() const { return set_constant || tangent_size == 0; }
  bool HasManifold() const { return manifold_kind != kManifoldNone; }
};

struct Problem {
  std::vector<ParameterBlock> pbs;
  std::vector<double> user_values;
  // Residual blocks, SoA.
  int num_rb = 0;
  std::vector<int> rb_type;
  std::vector<int> rb_pb_start;  // size num_rb + 1
  std::vector<int> rb_pb;        // parameter block ids
  std::vector<int> rb_loss_kind;
  std::vector<double> rb_loss_a, rb_loss_b;
  std::vector<int64_t> rb_fdata_start;  // size num_rb + 1
  std::vector<double> fdata;

  // ---- program (after Build) ----
  bool built = false;
  int jacobian_format = 0;  // 0 = BlockSparseMatrix, 1 = CompressedRowSparseMatrix
  int num_eliminate_blocks = 0;
  std::vector<int> program_pbs;   // active parameter blocks in program order
  std::vector<int> constant_pbs;  // program.cc constant_parameter_blocks_
  std::vector<int> program_rbs;   // residual block ids in program order
  double fixed_cost = 0.0;
  int num_parameters = 0, num_effective_parameters = 0, num_residuals = 0;
  int num_constant_parameters = 0;
  std::vector<double> constant_state;  // ConstantParameterBlocksToStateVector
  std::vector<int> residual_layout;
  // fork: per-residual layout (block_jacobian_writer.cc:62-150 /
  // compressed_row_jacobian_writer.cc:240-300)
  std::vector<int> jacobian_per_residual_layout, jacobian_per_residual_offsets;
  int num_jacobian_values = 0;
  // BSM: jacobian_layout (cell positions, active blocks in argument order)
  std::vector<int> jacobian_layout_start;    // per program rb, into jacobian_layout_storage
  std::vector<int> jacobian_layout_storage;
  // BSM block structure (block_jacobian_writer.cc:192-250)
  std::vector<int> col_block_size, col_block_pos;
  std::vector<int> row_block_size, row_block_pos;
  std::vector<int> row_cells_start, cell_block_id, cell_position;
  // CRS (compressed_row_jacobian_writer.cc:93-193)
  std::vector<int> crs_rows, crs_cols;
  int max_residuals = 0, max_params = 0, max_blocks = 0, max_scratch = 0;
};

static const CostTypeInfo& TypeOf(const Problem& p, int rb) {
  return CostTypes()[p.rb_type[rb]];
}

// residual_block.cc:68-204.  `params` point at the current state of each block.
// jacobians[j] may be null (constant block / not requested).  `plus_jacobians[j]`
// is null when the block has no manifold.  scratch must hold
// NumScratchDoublesForEvaluate doubles.
static bool EvaluateResidualBlock(const Problem& p, int rb, bool apply_loss_function,
                                  const double* const* params,
                                  const double* const* plus_jacobians,
                                  const int* tangent_sizes, double* cost,
                                  double* residuals, double** jacobians,
                                  double* scratch) {
  const CostTypeInfo& t = TypeOf(p, rb);
  const int num_parameter_blocks = t.num_blocks;
  const int num_residuals = t.num_residuals;

  double* global_jacobians[10];
  if (jacobians != nullptr) {
    for (int i = 0; i < num_parameter_blocks; ++i) {
      if (jacobians[i] != nullptr && plus_jacobians[i] != nullptr) {
        global_jacobians[i] = scratch;
        scratch += num_residuals * t.sizes[i];
      } else {
        global_jacobians[i] = jacobians[i];
      }
    }
  }
  const bool outputting_residuals = (residuals != nullptr);
  if (!outputting_residuals) residuals = scratch;

  double** eval_jacobians = (jacobians != nullptr) ? global_jacobians : nullptr;
  // InvalidateEvaluation (residual_block_utils.cc)
  *cost = kImpossibleValue;
  for (int i = 0; i < num_residuals; ++i) residuals[i] = kImpossibleValue;
  if (eval_jacobians != nullptr)
    for (int i = 0; i < num_parameter_blocks; ++i)
      if (eval_jacobians[i] != nullptr)
        for (int k = 0; k < num_residuals * t.sizes[i]; ++k)
          eval_jacobians[i][k] = kImpossibleValue;

  const double* fdata = p.fdata.data() + p.rb_fdata_start[rb];
  if (!t.eval(fdata, params, residuals, eval_jacobians)) return false;

  // IsEvaluationValid (residual_block_utils.cc:70-95)
  if (!IsArrayValid(num_residuals, residuals)) return false;
  if (eval_jacobians != nullptr)
    for (int i = 0; i < num_parameter_blocks; ++i)
      if (!IsArrayValid(num_residuals * t.sizes[i], eval_jacobians[i])) return false;

  double squared_norm = 0.0;
  for (int i = 0; i < num_residuals; ++i) squared_norm += residuals[i] * residuals[i];

  if (jacobians != nullptr) {
    for (int i = 0; i < num_parameter_blocks; ++i) {
      if (jacobians[i] != nullptr && plus_jacobians[i] != nullptr) {
        // small_blas.h MatrixMatrixMultiply<..., 0>: C = A * B
        const int n = t.sizes[i], tn = tangent_sizes[i];
        for (int r = 0; r < num_residuals; ++r)
          for (int c = 0; c < tn; ++c) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k)
              acc += global_jacobians[i][r * n + k] * plus_jacobians[i][k * tn + c];
            jacobians[i][r * tn + c] = acc;
          }
      }
    }
  }

  const int loss_kind = p.rb_loss_kind[rb];
  if (loss_kind == kLossNone || !apply_loss_function) {
    *cost = 0.5 * squared_norm;
    return true;
  }
  double rho[3];
  LossEvaluate(loss_kind, p.rb_loss_a[rb], p.rb_loss_b[rb], squared_norm, rho);
  *cost = 0.5 * rho[0];
  if (jacobians == nullptr && !outputting_residuals) return true;

  Corrector correct(squared_norm, rho);
  if (jacobians != nullptr)
    for (int i = 0; i < num_parameter_blocks; ++i)
      if (jacobians[i] != nullptr)
        correct.CorrectJacobian(num_residuals, tangent_sizes[i], residuals, jacobians[i]);
  if (outputting_residuals) correct.CorrectResiduals(num_residuals, residuals);
  return true;
}

// --------------------------------------------------------------- build program
struct BuildOptions {
  int reduce;                // 1: Program::CreateReducedProgram; 0: program as is
  int schur_reorder;         // 1: LexicographicallyOrderResidualBlocks(num_eliminate_blocks)
  int num_eliminate_blocks;
  int jacobian_format;       // 0 BSM, 1 CRS
};

static bool Build(Problem& p, const BuildOptions& o) {
  p.jacobian_format = o.jacobian_format;
  p.num_eliminate_blocks = o.num_eliminate_blocks;
  p.program_pbs.clear();
  p.constant_pbs.clear();
  p.program_rbs.clear();
  p.fixed_cost = 0.0;
  const int npb = static_cast<int>(p.pbs.size());
  for (auto& pb : p.pbs) {
    pb.index = -1; pb.state_offset = -1; pb.delta_offset = -1; pb.in_constant_list = false;
  }

  p.max_residuals = p.max_params = p.max_blocks = p.max_scratch = 0;
  for (int rb = 0; rb < p.num_rb; ++rb) {
    const CostTypeInfo& t = TypeOf(p, rb);
    int derivs = 0, total = 0;
    for (int j = 0; j < t.num_blocks; ++j) {
      const ParameterBlock& pb = p.pbs[p.rb_pb[p.rb_pb_start[rb] + j]];
      total += t.sizes[j];
      if (pb.HasManifold()) derivs += t.num_residuals * t.sizes[j];
    }
    p.max_residuals = std::max(p.max_residuals, t.num_residuals);
    p.max_blocks = std::max(p.max_blocks, t.num_blocks);
    p.max_params = std::max(p.max_params, total);
    // residual_block.cc:206-230 NumScratchDoublesForEvaluate
    p.max_scratch = std::max(p.max_scratch, derivs + t.num_residuals);
  }

  if (o.reduce) {
    // program.cc:324-430 RemoveFixedBlocks
    std::vector<char> used(npb, 0);
    std::vector<double> scratch(p.max_scratch + 16);
    for (int rb = 0; rb < p.num_rb; ++rb) {
      const CostTypeInfo& t = TypeOf(p, rb);
      bool all_constant = true;
      for (int k = 0; k < t.num_blocks; ++k) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + k];
        if (!p.pbs[id].IsConstant()) {
          all_constant = false;
          used[id] = 1;
        }
      }
      if (!all_constant) {
        p.program_rbs.push_back(rb);
        continue;
      }
      const double* params[10];
      const double* plus_j[10];
      int tangents[10];
      for (int k = 0; k < t.num_blocks; ++k) {
        const ParameterBlock& pb = p.pbs[p.rb_pb[p.rb_pb_start[rb] + k]];
        params[k] = p.user_values.data() + pb.user_offset;
        plus_j[k] = nullptr;
        tangents[k] = pb.tangent_size;
      }
      double cost = 0.0;
      if (!EvaluateResidualBlock(p, rb, true, params, plus_j, tangents, &cost, nullptr,
                                 nullptr, scratch.data()))
        return false;
      p.fixed_cost += cost;
    }
    for (int id = 0; id < npb; ++id) {
      if (used[id]) p.program_pbs.push_back(id);
      else { p.constant_pbs.push_back(id); p.pbs[id].in_constant_list = true; }
    }
  } else {
    // The program straight from the problem (evaluator_test.cc:117-126): every
    // parameter block is a program block; callers guarantee none is constant.
    for (int id = 0; id < npb; ++id) p.program_pbs.push_back(id);
    for (int rb = 0; rb < p.num_rb; ++rb) p.program_rbs.push_back(rb);
  }

  // program.cc:152-177 SetParameterOffsetsAndIndex
  int state_offset = 0, delta_offset = 0;
  for (int i = 0; i < static_cast<int>(p.program_pbs.size()); ++i) {
    ParameterBlock& pb = p.pbs[p.program_pbs[i]];
    pb.index = i;
    pb.state_offset = state_offset;
    pb.delta_offset = delta_offset;
    state_offset += pb.size;
    delta_offset += pb.tangent_size;
  }
  p.num_parameters = state_offset;
  p.num_effective_parameters = delta_offset;
  state_offset = 0;
  for (int i = 0; i < static_cast<int>(p.constant_pbs.size()); ++i) {
    ParameterBlock& pb = p.pbs[p.constant_pbs[i]];
    pb.index = i;
    pb.state_offset = state_offset;
    state_offset += pb.size;
  }
  p.num_constant_parameters = state_offset;
  // program.cc:98-103 ConstantParameterBlocksToStateVector
  p.constant_state.assign(state_offset, 0.0);
  for (int id : p.constant_pbs) {
    const ParameterBlock& pb = p.pbs[id];
    std::copy_n(p.user_values.data() + pb.user_offset, pb.size,
                p.constant_state.data() + pb.state_offset);
  }

  auto is_active = [&](int id) {
    const ParameterBlock& pb = p.pbs[id];
    return !pb.IsConstant() && !pb.in_constant_list;
  };

  if (o.schur_reorder && o.num_eliminate_blocks > 0) {
    // reorder_program.cc:254-335.  MinParameterBlock (:78-97): smallest index
    // among non-constant blocks, capped at size_of_first_elimination_group.
    const int E = o.num_eliminate_blocks;
    const int n = static_cast<int>(p.program_rbs.size());
    std::vector<int> per_e(E + 1, 0), min_pos(n);
    for (int i = 0; i < n; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      int position = E;
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) position = std::min(position, p.pbs[id].index);
      }
      min_pos[i] = position;
      per_e[position]++;
    }
    std::vector<int> offsets(E + 1);
    std::partial_sum(per_e.begin(), per_e.end(), offsets.begin());
    std::vector<int> reordered(n, -1);
    for (int i = 0; i < n; ++i) {
      const int bucket = min_pos[i];
      offsets[bucket]--;
      reordered[offsets[bucket]] = p.program_rbs[i];
    }
    p.program_rbs.swap(reordered);
  }

  // program_evaluator.h:336-347 BuildResidualLayout
  const int nrb = static_cast<int>(p.program_rbs.size());
  p.residual_layout.resize(nrb);
  int residual_pos = 0;
  for (int i = 0; i < nrb; ++i) {
    p.residual_layout[i] = residual_pos;
    residual_pos += TypeOf(p, p.program_rbs[i]).num_residuals;
  }
  p.num_residuals = residual_pos;

  p.jacobian_per_residual_layout.assign(nrb, 0);
  p.jacobian_per_residual_offsets.clear();
  p.jacobian_layout_start.assign(nrb + 1, 0);
  p.jacobian_layout_storage.clear();
  p.col_block_size.clear(); p.col_block_pos.clear();
  p.row_block_size.clear(); p.row_block_pos.clear();
  p.row_cells_start.clear(); p.cell_block_id.clear(); p.cell_position.clear();
  p.crs_rows.clear(); p.crs_cols.clear();

  if (o.jacobian_format == 0) {
    // block_jacobian_writer.cc:72-150 BuildJacobianLayout
    int f_block_pos = 0;
    int total_residuals = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) {
          if (p.pbs[id].index < o.num_eliminate_blocks)
            f_block_pos += t.num_residuals * p.pbs[id].tangent_size;
          total_residuals += t.num_residuals;
        }
      }
    }
    p.jacobian_per_residual_offsets.resize(total_residuals);
    int e_block_pos = 0;
    int per_residual_index = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      p.jacobian_layout_start[i] = static_cast<int>(p.jacobian_layout_storage.size());
      p.jacobian_per_residual_layout[i] = per_residual_index;
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (!is_active(id)) continue;
        const ParameterBlock& pb = p.pbs[id];
        for (int k = 0; k < t.num_residuals; ++k) {
          if (pb.index < o.num_eliminate_blocks) {
            if (k == 0) p.jacobian_layout_storage.push_back(e_block_pos);
            p.jacobian_per_residual_offsets[per_residual_index] = e_block_pos;
            e_block_pos += pb.tangent_size;
          } else {
            if (k == 0) p.jacobian_layout_storage.push_back(f_block_pos);
            p.jacobian_per_residual_offsets[per_residual_index] = f_block_pos;
            f_block_pos += pb.tangent_size;
          }
          per_residual_index++;
        }
      }
    }
    p.jacobian_layout_start[nrb] = static_cast<int>(p.jacobian_layout_storage.size());
    p.num_jacobian_values = f_block_pos;

    // block_jacobian_writer.cc:192-250 CreateJacobian (block structure)
    int cursor = 0;
    for (int id : p.program_pbs) {
      p.col_block_size.push_back(p.pbs[id].tangent_size);
      p.col_block_pos.push_back(cursor);
      cursor += p.pbs[id].tangent_size;
    }
    int row_block_position = 0;
    p.row_cells_start.push_back(0);
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      p.row_block_size.push_back(t.num_residuals);
      p.row_block_pos.push_back(row_block_position);
      row_block_position += t.num_residuals;
      std::vector<std::pair<int, int>> cells;
      for (int j = 0, k = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) {
          cells.emplace_back(p.pbs[id].index,
                             p.jacobian_layout_storage[p.jacobian_layout_start[i] + k]);
          k++;
        }
      }
      // CellLessThan (block_structure.cc): by block_id, then position.
      std::sort(cells.begin(), cells.end());
      for (auto& c : cells) {
        p.cell_block_id.push_back(c.first);
        p.cell_position.push_back(c.second);
      }
      p.row_cells_start.push_back(static_cast<int>(p.cell_block_id.size()));
    }
  } else {
    // compressed_row_jacobian_writer.cc:93-193 CreateJacobian
    int nnz = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) nnz += t.num_residuals * p.pbs[id].tangent_size;
      }
    }
    p.crs_rows.assign(p.num_residuals + 1, 0);
    p.crs_cols.assign(nnz + p.num_effective_parameters, 0);
    int row_pos = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      int num_derivatives = 0;
      std::vector<int> parameter_indices;
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) {
          parameter_indices.push_back(p.pbs[id].index);
          num_derivatives += p.pbs[id].tangent_size;
        }
      }
      std::sort(parameter_indices.begin(), parameter_indices.end());
      for (int j = 0; j < t.num_residuals; ++j)
        p.crs_rows[row_pos + j + 1] = p.crs_rows[row_pos + j] + num_derivatives;
      int col_pos = 0;
      for (int parameter_index : parameter_indices) {
        const ParameterBlock& pb = p.pbs[p.program_pbs[parameter_index]];
        for (int r = 0; r < t.num_residuals; ++r) {
          const int column_block_begin = p.crs_rows[row_pos + r] + col_pos;
          for (int c = 0; c < pb.tangent_size; ++c)
            p.crs_cols[column_block_begin + c] = pb.delta_offset + c;
        }
        col_pos += pb.tangent_size;
      }
      row_pos += t.num_residuals;
    }
    // compressed_row_jacobian_writer.cc:240-300 CreateJacobianPerResidualLayout
    int total_residuals = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      for (int j = 0; j < t.num_blocks; ++j)
        if (is_active(p.rb_pb[p.rb_pb_start[rb] + j])) total_residuals += t.num_residuals;
    }
    p.jacobian_per_residual_offsets.assign(total_residuals, -1);
    p.num_jacobian_values = 0;
    int per_residual_pos = 0;
    int row_start = 0;
    for (int i = 0; i < nrb; ++i) {
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      p.jacobian_per_residual_layout[i] = per_residual_pos;
      // GetOrderedParameterBlocks(..., get_active_parameter_index = true) :71-91
      std::vector<std::pair<int, int>> blocks;
      int active_parameter_index = 0;
      for (int j = 0; j < t.num_blocks; ++j) {
        const int id = p.rb_pb[p.rb_pb_start[rb] + j];
        if (is_active(id)) {
          blocks.emplace_back(p.pbs[id].index, active_parameter_index);
          active_parameter_index++;
        }
      }
      std::sort(blocks.begin(), blocks.end());
      for (int r = 0; r < t.num_residuals; ++r) {
        int col_pos = 0;
        for (auto& b : blocks) {
          const ParameterBlock& pb = p.pbs[p.program_pbs[b.first]];
          p.jacobian_per_residual_offsets[per_residual_pos + r +
                                          t.num_residuals * b.second] = row_start + col_pos;
          col_pos += pb.tangent_size;
          p.num_jacobian_values += pb.tangent_size;
        }
        row_start += col_pos;
      }
      per_residual_pos += static_cast<int>(blocks.size()) * t.num_residuals;
    }
  }
  p.built = true;
  return true;
}

// ----------------------------------------------------------------- evaluate
struct EvalScratch {
  double cost = 0.0;
  std::vector<double> gradient;
  std::vector<double> block_residuals;
  std::vector<double> jac_storage;
  std::vector<double> evaluate_scratch;
  std::vector<double> plus_jac_storage;
};

// program_evaluator.h:134-283.  Any of residuals/gradient/jacobian_values may be
// null.  Returns false on abort.
static bool Evaluate(const Problem& p, const double* state, bool apply_loss_function,
                     int num_threads, double* cost, double* residuals, double* gradient,
                     double* jacobian_values) {
  const int nrb = static_cast<int>(p.program_rbs.size());
  num_threads = std::max(1, std::min(num_threads, std::max(1, nrb)));
  if (residuals) std::fill(residuals, residuals + p.num_residuals, 0.0);
  if (jacobian_values) {
    const int64_t nv = p.jacobian_format == 0
                           ? p.num_jacobian_values
                           : static_cast<int64_t>(p.crs_cols.size());
    // SetZero zeroes num_nonzeros values (BSM) / the whole values array (CRS).
    std::fill(jacobian_values, jacobian_values + nv, 0.0);
  }
  std::vector<EvalScratch> scratch(num_threads);
  for (auto& s : scratch) {
    s.cost = 0.0;
    if (gradient) s.gradient.assign(p.num_effective_parameters, 0.0);
    s.block_residuals.resize(p.max_residuals);
    s.jac_storage.resize(static_cast<size_t>(p.max_residuals) * p.max_params);
    s.evaluate_scratch.resize(p.max_scratch + 16);
    s.plus_jac_storage.resize(static_cast<size_t>(p.max_params) * p.max_params + 16);
  }
  std::atomic<bool> abort(false);

  auto work = [&](int thread_id, int begin, int end) {
    EvalScratch& s = scratch[thread_id];
    for (int i = begin; i < end; ++i) {
      if (abort) return;
      const int rb = p.program_rbs[i];
      const CostTypeInfo& t = TypeOf(p, rb);
      const double* params[10];
      const double* plus_j[10];
      int tangents[10];
      double* jac_ptrs[10];
      double* block_residuals = nullptr;
      if (residuals) block_residuals = residuals + p.residual_layout[i];
      else if (gradient) block_residuals = s.block_residuals.data();

      const bool want_j = (jacobian_values != nullptr || gradient != nullptr);
      double* jcursor = s.jac_storage.data();
      double* pcursor = s.plus_jac_storage.data();
      for (int j = 0; j < t.num_blocks; ++j) {
        const ParameterBlock& pb = p.pbs[p.rb_pb[p.rb_pb_start[rb] + j]];
        const bool constant = pb.IsConstant() || pb.in_constant_list;
        // program.cc:80-88: active blocks read the state vector; constant ones keep
        // their (user) state.
        params[j] = constant ? (pb.in_constant_list
                                    ? p.constant_state.data() + pb.state_offset
                                    : p.user_values.data() + pb.user_offset)
                             : state + pb.state_offset;
        tangents[j] = pb.tangent_size;
        plus_j[j] = nullptr;
        jac_ptrs[j] = nullptr;
        if (want_j && !constant) {
          // block_evaluate_preparer.cc:50-77 / scratch_evaluate_preparer.cc
          jac_ptrs[j] = jcursor;
          jcursor += t.num_residuals * pb.tangent_size;
          if (pb.HasManifold()) {
            // parameter_block.h:312-338 UpdatePlusJacobian at SetState.
            ManifoldPlusJacobian(pb.manifold_kind, pb.manifold_param, pb.size, params[j],
                                 pcursor);
            plus_j[j] = pcursor;
            pcursor += pb.size * pb.tangent_size;
          }
        }
      }
      double block_cost;
      if (!EvaluateResidualBlock(p, rb, apply_loss_function, params, plus_j, tangents,
                                 &block_cost, block_residuals,
                                 want_j ? jac_ptrs : nullptr, s.evaluate_scratch.data())) {
        abort = true;
        return;
      }
      s.cost += block_cost;

      if (jacobian_values != nullptr) {
        if (p.jacobian_format == 0) {
          // BlockJacobianWriter: blocks are evaluated in place (Write is a nop).
          for (int j = 0, k = 0; j < t.num_blocks; ++j) {
            if (jac_ptrs[j] == nullptr) continue;
            const int pos = p.jacobian_layout_storage[p.jacobian_layout_start[i] + k];
            std::copy_n(jac_ptrs[j], t.num_residuals * tangents[j], jacobian_values + pos);
            k++;
          }
        } else {
          // compressed_row_jacobian_writer.cc:195-238 Write
          std::pair<int, int> blocks[10];
          int nb = 0;
          for (int j = 0; j < t.num_blocks; ++j) {
            const ParameterBlock& pb = p.pbs[p.rb_pb[p.rb_pb_start[rb] + j]];
            if (jac_ptrs[j] != nullptr) blocks[nb++] = std::make_pair(pb.index, j);
          }
          std::sort(blocks, blocks + nb);
          int col_pos = 0;
          for (int bi = 0; bi < nb; ++bi) {
            const int argument = blocks[bi].second;
            const int size = tangents[argument];
            for (int r = 0; r < t.num_residuals; ++r) {
              const double* block_row_begin = jac_ptrs[argument] + r * size;
              double* column_block_begin =
                  jacobian_values + p.crs_rows[p.residual_layout[i] + r] + col_pos;
              std::copy(block_row_begin, block_row_begin + size, column_block_begin);
            }
            col_pos += size;
          }
        }
      }
      if (gradient != nullptr) {
        // program_evaluator.h:241-256, small_blas MatrixTransposeVectorMultiply<..., 1>
        for (int j = 0; j < t.num_blocks; ++j) {
          if (jac_ptrs[j] == nullptr) continue;
          const ParameterBlock& pb = p.pbs[p.rb_pb[p.rb_pb_start[rb] + j]];
          double* g = s.gradient.data() + pb.delta_offset;
          for (int c = 0; c < tangents[j]; ++c) {
            double acc = 0.0;
            for (int r = 0; r < t.num_residuals; ++r)
              acc += jac_ptrs[j][r * tangents[j] + c] * block_residuals[r];
            g[c] += acc;
          }
        }
      }
    }
  };

  if (num_threads == 1) {
    work(0, 0, nrb);
  } else {
    std::vector<std::thread> threads;
    for (int tid = 0; tid < num_threads; ++tid) {
      const int begin = static_cast<int>(static_cast<int64_t>(nrb) * tid / num_threads);
      const int end = static_cast<int>(static_cast<int64_t>(nrb) * (tid + 1) / num_threads);
      threads.emplace_back(work, tid, begin, end);
    }
    for (auto& th : threads) th.join();
  }
  if (abort) return false;
  // program_evaluator.h:258-277
  *cost = 0.0;
  if (gradient) std::fill(gradient, gradient + p.num_effective_parameters, 0.0);
  for (auto& s : scratch) {
    *cost += s.cost;
    if (gradient)
      for (int k = 0; k < p.num_effective_parameters; ++k) gradient[k] += s.gradient[k];
  }
  return true;
}

}  // namespace oracle

// ================================================================== C API
using oracle::Problem;

extern "C" {

int oracle_num_cost_types() { return static_cast<int>(oracle::CostTypes().size()); }

// out[0] = num_residuals, out[1] = num_blocks, out[2] = fdata_len, out[3..] = sizes
void oracle_cost_type_info(int type, int* out) {
  const auto& t = oracle::CostTypes()[type];
  out[0] = t.num_residuals;
  out[1] = t.num_blocks;
  out[2] = t.fdata_len;
  for (int i = 0; i < t.num_blocks; ++i) out[3 + i] = t.sizes[i];
}

void* oracle_problem_create(int num_pb, const int* pb_size, const double* pb_values,
                            const uint8_t* pb_constant, const int* pb_manifold_kind,
                            const int* pb_manifold_param, int num_rb, const int* rb_type,
                            const int* rb_pb, const int* rb_loss_kind,
                            const double* rb_loss_a, const double* rb_loss_b,
                            const double* fdata) {
  auto* p = new Problem;
  p->pbs.resize(num_pb);
  int64_t off = 0;
  for (int i = 0; i < num_pb; ++i) {
    auto& pb = p->pbs[i];
    pb.size = pb_size[i];
    pb.set_constant = pb_constant ? pb_constant[i] != 0 : false;
    pb.manifold_kind = pb_manifold_kind ? pb_manifold_kind[i] : 0;
    pb.manifold_param = pb_manifold_param ? pb_manifold_param[i] : 0;
    pb.tangent_size =
        oracle::ManifoldTangentSize(pb.manifold_kind, pb.manifold_param, pb.size);
    pb.user_offset = off;
    off += pb.size;
  }
  p->user_values.assign(pb_values, pb_values + off);
  p->num_rb = num_rb;
  p->rb_type.assign(rb_type, rb_type + num_rb);
  p->rb_pb_start.resize(num_rb + 1);
  p->rb_fdata_start.resize(num_rb + 1);
  int pbs = 0;
  int64_t fd = 0;
  for (int i = 0; i < num_rb; ++i) {
    const auto& t = oracle::CostTypes()[rb_type[i]];
    p->rb_pb_start[i] = pbs;
    p->rb_fdata_start[i] = fd;
    pbs += t.num_blocks;
    fd += t.fdata_len;
  }
  p->rb_pb_start[num_rb] = pbs;
  p->rb_fdata_start[num_rb] = fd;
  p->rb_pb.assign(rb_pb, rb_pb + pbs);
  if (fd > 0) p->fdata.assign(fdata, fdata + fd);
  p->rb_loss_kind.assign(num_rb, 0);
  p->rb_loss_a.assign(num_rb, 0.0);
  p->rb_loss_b.assign(num_rb, 0.0);
  if (rb_loss_kind) p->rb_loss_kind.assign(rb_loss_kind, rb_loss_kind + num_rb);
  if (rb_loss_a) p->rb_loss_a.assign(rb_loss_a, rb_loss_a + num_rb);
  if (rb_loss_b) p->rb_loss_b.assign(rb_loss_b, rb_loss_b + num_rb);
  return p;
}

void oracle_problem_destroy(void* h) { delete static_cast<Problem*>(h); }

int oracle_problem_build(void* h, int reduce, int schur_reorder, int num_eliminate_blocks,
                         int jacobian_format) {
  oracle::BuildOptions o{reduce, schur_reorder, num_eliminate_blocks, jacobian_format};
  return oracle::Build(*static_cast<Problem*>(h), o) ? 1 : 0;
}

// dims: [num_parameters, num_effective_parameters, num_residuals, num_program_rbs,
//        num_program_pbs, num_jacobian_values, values_size, num_cells,
//        per_residual_offsets_size, num_constant_parameters]
void oracle_problem_dims(void* h, int64_t* dims) {
  auto& p = *static_cast<Problem*>(h);
  dims[0] = p.num_parameters;
  dims[1] = p.num_effective_parameters;
  dims[2] = p.num_residuals;
  dims[3] = static_cast<int64_t>(p.program_rbs.size());
  dims[4] = static_cast<int64_t>(p.program_pbs.size());
  dims[5] = p.num_jacobian_values;
  dims[6] = p.jacobian_format == 0 ? p.num_jacobian_values
                                   : static_cast<int64_t>(p.crs_cols.size());
  dims[7] = static_cast<int64_t>(p.cell_block_id.size());
  dims[8] = static_cast<int64_t>(p.jacobian_per_residual_offsets.size());
  dims[9] = p.num_constant_parameters;
}

double oracle_problem_fixed_cost(void* h) { return static_cast<Problem*>(h)->fixed_cost; }

// program.cc:90-96 ParameterBlocksToStateVector
void oracle_problem_initial_state(void* h, double* state) {
  auto& p = *static_cast<Problem*>(h);
  for (int id : p.program_pbs) {
    const auto& pb = p.pbs[id];
    std::copy_n(p.user_values.data() + pb.user_offset, pb.size, state + pb.state_offset);
  }
}

// which: 0 residual_layout, 1 jacobian_per_residual_layout, 2 jacobian_per_residual_offsets,
// 3 program_rbs, 4 program_pbs, 5 col_block_size, 6 col_block_pos, 7 row_block_size,
// 8 row_block_pos, 9 row_cells_start, 10 cell_block_id, 11 cell_position, 12 crs_rows,
// 13 crs_cols, 14 constant_pbs, 15 jacobian_layout_storage
int64_t oracle_problem_get_ints(void* h, int which, int* out) {
  auto& p = *static_cast<Problem*>(h);
  const std::vector<int>* v = nullptr;
  switch (which) {
    case 0: v = &p.residual_layout; break;
    case 1: v = &p.jacobian_per_residual_layout; break;
    case 2: v = &p.jacobian_per_residual_offsets; break;
    case 3: v = &p.program_rbs; break;
    case 4: v = &p.program_pbs; break;
    case 5: v = &p.col_block_size; break;
    case 6: v = &p.col_block_pos; break;
    case 7: v = &p.row_block_size; break;
    case 8: v = &p.row_block_pos; break;
    case 9: v = &p.row_cells_start; break;
    case 10: v = &p.cell_block_id; break;
    case 11: v = &p.cell_position; break;
    case 12: v = &p.crs_rows; break;
    case 13: v = &p.crs_cols; break;
    case 14: v = &p.constant_pbs; break;
    case 15: v = &p.jacobian_layout_storage; break;
    default: return -1;
  }
  if (out) std::copy(v->begin(), v->end(), out);
  return static_cast<int64_t>(v->size());
}

// Per program parameter block: index, state_offset, delta_offset, tangent_size (4 ints each).
void oracle_problem_pb_table(void* h, int* out) {
  auto& p = *static_cast<Problem*>(h);
  for (size_t i = 0; i < p.program_pbs.size(); ++i) {
    const auto& pb = p.pbs[p.program_pbs[i]];
    out[4 * i + 0] = pb.index;
    out[4 * i + 1] = pb.state_offset;
    out[4 * i + 2] = pb.delta_offset;
    out[4 * i + 3] = pb.tangent_size;
  }
}

int oracle_problem_evaluate(void* h, const double* state, int apply_loss_function,
                            int num_threads, double* cost, double* residuals,
                            double* gradient, double* jacobian_values) {
  auto& p = *static_cast<Problem*>(h);
  if (!p.built) return 0;
  return oracle::Evaluate(p, state, apply_loss_function != 0, num_threads, cost, residuals,
                          gradient, jacobian_values)
             ? 1
             : 0;
}

// program.cc:121-150 Program::Plus
int oracle_problem_plus(void* h, const double* state, const double* delta,
                        double* state_plus_delta) {
  auto& p = *static_cast<Problem*>(h);
  for (int id : p.program_pbs) {
    const auto& pb = p.pbs[id];
    oracle::ManifoldPlus(pb.manifold_kind, pb.manifold_param, pb.size,
                         state + pb.state_offset, delta + pb.delta_offset,
                         state_plus_delta + pb.state_offset);
  }
  return 1;
}

// Stand-alone pieces for unit tests against the reference's known-answer tests.
int oracle_cost_evaluate(int type, const double* fdata, const double* params_concat,
                         double* residuals, double* jacobians_concat) {
  const auto& t = oracle::CostTypes()[type];
  const double* params[10];
  double* jac[10];
  int off = 0, joff = 0;
  for (int i = 0; i < t.num_blocks; ++i) {
    params[i] = params_concat + off;
    jac[i] = jacobians_concat ? jacobians_concat + joff : nullptr;
    off += t.sizes[i];
    joff += t.num_residuals * t.sizes[i];
  }
  return t.eval(fdata, params, residuals, jacobians_concat ? jac : nullptr) ? 1 : 0;
}

void oracle_loss_evaluate(int kind, double a, double b, double s, double* rho) {
  oracle::LossEvaluate(kind, a, b, s, rho);
}

void oracle_corrector(double sq_norm, const double* rho, int num_rows, int num_cols,
                      double* residuals, double* jacobian) {
  oracle::Corrector c(sq_norm, rho);
  if (jacobian) c.CorrectJacobian(num_rows, num_cols, residuals, jacobian);
  c.CorrectResiduals(num_rows, residuals);
}

void oracle_manifold_plus_jacobian(int kind, int param, int ambient, const double* x,
                                   double* jacobian) {
  oracle::ManifoldPlusJacobian(kind, param, ambient, x, jacobian);
}

void oracle_manifold_plus(int kind, int param, int ambient, const double* x,
                          const double* delta, double* out) {
  oracle::ManifoldPlus(kind, param, ambient, x, delta, out);
}

void oracle_angle_axis_rotate_point(const double* aa, const double* pt, double* out) {
  oracle::AngleAxisRotatePoint(aa, pt, out);
}

void oracle_quaternion_to_angle_axis_jet(const double* q, double* value, double* jac) {
  using J = oracle::Jet<4>;
  J x[4], out[3];
  for (int i = 0; i < 4; ++i) x[i] = J(q[i], i);
  oracle::QuaternionToAngleAxis(x, out);
  for (int r = 0; r < 3; ++r) {
    value[r] = out[r].a;
    for (int c = 0; c < 4; ++c) jac[r * 4 + c] = out[r].v[c];
  }
}

// Same battery, in the same order, as oracle/ref_arith.cc ref_jet_battery.
int oracle_jet_battery(const double* in, double* out) {
  using namespace oracle;
  using J = oracle::Jet<2>;
  const J x(in[0], 0), y(in[1], 1);
  int k = 0;
  auto put = [&](const J& j) { out[k++] = j.a; out[k++] = j.v[0]; out[k++] = j.v[1]; };
  put(x + y); put(x - y); put(x * y); put(x / y); put(-x); put(x + 1.5); put(1.5 - x);
  put(x * 1.5); put(1.5 / x); put(x / 1.5);
  put(sqrt(x)); put(exp(x)); put(log(x)); put(sin(x));
  put(cos(x)); put(tan(x)); put(asin(x / 3.0)); put(acos(x / 3.0));
  put(atan(x)); put(sinh(x)); put(cosh(x)); put(tanh(x));
  put(abs(-x)); put(atan2(y, x)); put(pow(x, 1.7)); put(pow(x, y));
  put(hypot(x, y)); put(hypot(x, y, x * y)); put(cbrt(x));
  put(exp2(x)); put(log2(x)); put(log10(x)); put(log1p(x));
  put(expm1(x)); put(fmax(x, y)); put(fmin(x, y)); put(erf(x));
  put(erfc(x)); put(copysign(x, -y)); put(fma(x, y, x));
  return k / 3;
}

}  // extern "C"
