// The evaluation kernels (sm_100a) and their launch thunk.
//
// One instantiation per <CostFunctor, LossFunctionCUDA, kNumResiduals, Ns...>,
// emitted in the user's translation unit by ProblemCUDA::AddResidualBlock; the
// precompiled engine (csrc/engine.cu) calls it through the C-ABI function pointer
// cb200_launch_fn.  Replaces the reference's EvaluateKernel
// (include/ceres/internal/cuda_evaluator_kernel.h:301-422) and the pieces it
// calls: AutoDifferentiate (include/ceres/internal/autodiff.h:318-381),
// MatrixMultiply J * PlusJacobian (:163-173,355-371), Corrector
// (include/ceres/internal/corrector.h:82-213), ComputeGradient (:193-217) and
// WriteJacobians (:264-294).
//
// Per residual block the order of operations is the CPU evaluator's
// (internal/ceres/residual_block.cc:68-204): autodiff -> finite check ->
// J * PlusJacobian -> s = ||r||^2 (uncorrected) -> rho -> cost = rho[0] / 2 ->
// CorrectJacobian (uses uncorrected r) -> CorrectResiduals -> gradient J^T r
// (both corrected) -> scatter.
//
// Design (DESIGN.md has the numbers):
//  * one thread per residual block, Jets with compile-time sparsity masks
//    (ceres/jet.h): ~500 FP64 instructions and < 128 registers for the BAL
//    functor, no local-memory spills, no per-thread Jacobian scratch in HBM;
//  * structure-of-arrays inputs, argument-major, so every per-block table is read
//    with unit-stride loads; parameters are gathered through L1/L2 (a camera is
//    72 B and hot, consecutive blocks share their point);
//  * each Jacobian cell is written exactly once, straight to its final position in
//    the BlockSparseMatrix / CompressedRowSparseMatrix values array: no memset of
//    the values, no second copy;
//  * cost: warp shuffle + one partial per thread block, summed in a fixed order;
//  * gradient: runs of blocks that share a parameter block (the points of a
//    Schur-ordered BAL problem) are pre-reduced with a segmented warp shuffle;
//    what is left goes out as fire-and-forget red.global.add.f64.
#ifndef CERES_B200_INTERNAL_EVALUATE_KERNEL_CUH_
#define CERES_B200_INTERNAL_EVALUATE_KERNEL_CUH_

#include <cuda_runtime.h>

#include <cstdint>
#include <utility>

#include "ceres/jet.h"
#include "ceres_b200.h"

namespace ceres {
namespace internal {

constexpr int kEvaluateThreads = 128;
// internal/ceres/array_utils.h: the value autodiff leaves in outputs a functor
// did not write; evaluations containing it are invalid.
constexpr double kImpossibleValue = 1e302;

template <int... Ns>
struct BlockDims {
  static constexpr int kNumBlocks = sizeof...(Ns);
  static constexpr int kNumParameters = (Ns + ... + 0);
  __host__ __device__ static constexpr int Size(int j) {
    constexpr int s[kNumBlocks > 0 ? kNumBlocks : 1] = {Ns...};
    return s[j];
  }
  __host__ __device__ static constexpr int Offset(int j) {
    int o = 0;
    for (int i = 0; i < j; ++i) o += Size(i);
    return o;
  }
  __host__ __device__ static constexpr int MaxSize() {
    int m = 0;
    for (int i = 0; i < kNumBlocks; ++i) m = Size(i) > m ? Size(i) : m;
    return m;
  }
};

template <typename Dims, typename Functor, typename T, std::size_t... Is>
__device__ __forceinline__ bool CallFunctor(const Functor& f, const T* x, T* out,
                                            std::index_sequence<Is...>) {
  return f((x + Dims::Offset(Is))..., out);
}

template <typename F, std::size_t... Js>
__device__ __forceinline__ void ForEachBlock(F&& f, std::index_sequence<Js...>) {
  (f(std::integral_constant<int, static_cast<int>(Js)>{}), ...);
}

__device__ __forceinline__ bool IsValidValue(double x) {
  return ::fabs(x) <= 1.7976931348623157e308 && x != 1e302;
}

// Sums `v` over runs of consecutive lanes; `run_end` is one past the last lane of
// the calling lane's run.  Afterwards the first lane of every run holds its total.
template <int kCount>
__device__ __forceinline__ void WarpSegmentedSum(int run_end, double (&v)[kCount], int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const bool take = lane + d < run_end;
#pragma unroll
    for (int c = 0; c < kCount; ++c) {
      const double o = __shfl_down_sync(0xffffffffu, v[c], d);
      if (take) v[c] += o;
    }
  }
}

__device__ __forceinline__ void RedAdd(double* address, double value) {
  // No return value wanted: red.global.add.f64 instead of an atom round trip.
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(address), "d"(value) : "memory");
}

// Everything that happens to one parameter block's Jacobian after autodiff.
// B is the kRes x kSize ambient block; on exit its first `tangent` columns hold
// the final (manifold-projected, loss-corrected) block.
template <int kRes, int kSize>
struct BlockEpilogue {
  // B <- B * P, P row-major kSize x tangent (cuda_evaluator_kernel.h:355-371).
  static __device__ __forceinline__ void MultiplyPlusJacobian(double (&B)[kRes][kSize],
                                                              const double* __restrict__ P,
                                                              int tangent) {
    double out[kRes][kSize];
#pragma unroll
    for (int c = 0; c < kSize; ++c) {
      if (c < tangent) {
        double p[kSize];
#pragma unroll
        for (int k = 0; k < kSize; ++k) p[k] = P[k * tangent + c];
#pragma unroll
        for (int r = 0; r < kRes; ++r) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < kSize; ++k) acc += B[r][k] * p[k];
          out[r][c] = acc;
        }
      } else {
#pragma unroll
        for (int r = 0; r < kRes; ++r) out[r][c] = 0.0;
      }
    }
#pragma unroll
    for (int r = 0; r < kRes; ++r)
#pragma unroll
      for (int c = 0; c < kSize; ++c) B[r][c] = out[r][c];
  }

  // corrector.h:174-213 CorrectJacobian, column by column.
  static __device__ __forceinline__ void Correct(double (&B)[kRes][kSize], int tangent,
                                                 const double (&res)[kRes], double sqrt_rho1,
                                                 double alpha_sq_norm) {
    if (alpha_sq_norm == 0.0) {
#pragma unroll
      for (int r = 0; r < kRes; ++r)
#pragma unroll
        for (int c = 0; c < kSize; ++c) B[r][c] *= sqrt_rho1;
      return;
    }
#pragma unroll
    for (int c = 0; c < kSize; ++c) {
      double r_transpose_j = 0.0;
#pragma unroll
      for (int r = 0; r < kRes; ++r) r_transpose_j += B[r][c] * res[r];
#pragma unroll
      for (int r = 0; r < kRes; ++r)
        B[r][c] = sqrt_rho1 * (B[r][c] - alpha_sq_norm * res[r] * r_transpose_j);
    }
  }
};

template <bool kWithJacobians, typename Functor, typename Loss, int kRes, int... Ns>
__global__ void __launch_bounds__(kEvaluateThreads)
    EvaluateKernel(const cb200_launch_args a) {
  using Dims = BlockDims<Ns...>;
  constexpr int kNB = Dims::kNumBlocks;
  constexpr int kNP = Dims::kNumParameters;

  const int t = blockIdx.x * kEvaluateThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = t < a.n;
  const int tt = valid ? t : a.n - 1;  // idle lanes shadow the last block, write nothing

  const Functor& functor = static_cast<const Functor*>(a.functors)[tt];
  const int4* __restrict__ pb_table = reinterpret_cast<const int4*>(a.parameter_block_table);

  int pb_id[kNB];
  int4 pb[kNB];  // x state_offset, y delta_offset, z tangent_size, w plus_jacobian_offset
#pragma unroll
  for (int j = 0; j < kNB; ++j) {
    pb_id[j] = __ldg(a.parameter_block + static_cast<size_t>(j) * a.n + tt);
    pb[j] = __ldg(pb_table + pb_id[j]);
  }

  double res[kRes];
  bool ok;
  double cost = 0.0;

  if constexpr (!kWithJacobians) {
    // Cost / residual only: plain doubles, no Jets
    // (AutoDiffCostFunction::Evaluate with jacobians == nullptr).
    double x[kNP];
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
      const double* __restrict__ src = a.state + pb[j].x;
#pragma unroll
      for (int i = 0; i < Dims::Size(j); ++i) x[Dims::Offset(j) + i] = __ldg(src + i);
    }
#pragma unroll
    for (int r = 0; r < kRes; ++r) res[r] = kImpossibleValue;
    ok = CallFunctor<Dims>(functor, x, res, std::make_index_sequence<kNB>{});
#pragma unroll
    for (int r = 0; r < kRes; ++r) ok = ok && IsValidValue(res[r]);

    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kRes; ++r) s += res[r] * res[r];
    if (a.apply_loss_function) {
      const Loss* __restrict__ losses = static_cast<const Loss*>(a.loss_table);
      const Loss& loss = losses[a.loss_index ? __ldg(a.loss_index + tt) : 0];
      double rho[3];
      loss.Evaluate(s, rho);
      cost = 0.5 * rho[0];
      if (a.output_residuals) {
        // corrector.h:82-147 then :159-166
        const double sqrt_rho1 = ::sqrt(rho[1]);
        double scaling = sqrt_rho1;
        if (!(s == 0.0 || rho[2] <= 0.0)) {
          const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
          const double alpha = 1.0 - ::sqrt(D);
          scaling = sqrt_rho1 / (1 - alpha);
        }
#pragma unroll
        for (int r = 0; r < kRes; ++r) res[r] *= scaling;
      }
    } else {
      cost = 0.5 * s;
    }
  } else {
    using JetT = Jet<double, kNP>;
    JetT x[kNP];
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
      const double* __restrict__ src = a.state + pb[j].x;
#pragma unroll
      for (int i = 0; i < Dims::Size(j); ++i)
        x[Dims::Offset(j) + i] = JetT(__ldg(src + i), Dims::Offset(j) + i);
    }
    JetT out[kRes];
#pragma unroll
    for (int r = 0; r < kRes; ++r) out[r] = JetT::Filled(kImpossibleValue, kImpossibleValue);
    ok = CallFunctor<Dims>(functor, x, out, std::make_index_sequence<kNB>{});

    // IsEvaluationValid (internal/ceres/residual_block_utils.cc:70-95): the CPU
    // evaluator rejects non-finite / unwritten values; so does this kernel.
#pragma unroll
    for (int r = 0; r < kRes; ++r) {
      res[r] = out[r].a;
      ok = ok && IsValidValue(out[r].a);
#pragma unroll
      for (int i = 0; i < kNP; ++i)
        if (out[r].lane(i)) ok = ok && IsValidValue(out[r].v[i]);
    }

    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kRes; ++r) s += res[r] * res[r];

    double sqrt_rho1 = 1.0, residual_scaling = 1.0, alpha_sq_norm = 0.0;
    bool correct = false;
    if (a.apply_loss_function) {
      const Loss* __restrict__ losses = static_cast<const Loss*>(a.loss_table);
      const Loss& loss = losses[a.loss_index ? __ldg(a.loss_index + tt) : 0];
      double rho[3];
      loss.Evaluate(s, rho);
      cost = 0.5 * rho[0];
      correct = true;
      sqrt_rho1 = ::sqrt(rho[1]);
      residual_scaling = sqrt_rho1;
      if (!(s == 0.0 || rho[2] <= 0.0)) {
        const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
        const double alpha = 1.0 - ::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / s;
      }
    } else {
      cost = 0.5 * s;
    }

    double res_corrected[kRes];
#pragma unroll
    for (int r = 0; r < kRes; ++r) res_corrected[r] = res[r] * residual_scaling;

    int row_stride_crs = 0;
    if (a.crs && a.output_jacobian) row_stride_crs = __ldg(a.jacobian_row_stride + tt);

    // Per parameter block: project, correct, accumulate the gradient, scatter.
    auto epilogue = [&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr int kSize = Dims::Size(j);
      constexpr int kOff = Dims::Offset(j);
      const bool active = pb[j].y >= 0;  // constant blocks have no Jacobian
      int tangent = kSize;
      double B[kRes][kSize];
#pragma unroll
      for (int r = 0; r < kRes; ++r)
#pragma unroll
        for (int c = 0; c < kSize; ++c) B[r][c] = out[r].v[kOff + c];
      if (active && pb[j].w >= 0) {
        tangent = pb[j].z;
        BlockEpilogue<kRes, kSize>::MultiplyPlusJacobian(B, a.plus_jacobians + pb[j].w,
                                                         tangent);
      }
      if (correct) BlockEpilogue<kRes, kSize>::Correct(B, tangent, res, sqrt_rho1, alpha_sq_norm);

      if (a.output_gradient) {
        double g[kSize];
#pragma unroll
        for (int c = 0; c < kSize; ++c) {
          double acc = 0.0;
#pragma unroll
          for (int r = 0; r < kRes; ++r) acc += B[r][c] * res_corrected[r];
          g[c] = (valid && ok && active) ? acc : 0.0;
        }
        // Runs of consecutive blocks sharing this parameter block are summed in
        // the warp first (warp-uniform test, so no divergence around shuffles).
        const int key = (valid && active) ? pb_id[j] : -1 - lane;
        const int prev_key = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev_key != key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (heads != 0xffffffffu) {
          const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
          const int run_end = above ? __ffs(above) - 1 : 32;
          WarpSegmentedSum<kSize>(run_end, g, lane);
        }
        if (head && valid && ok && active) {
          double* __restrict__ dst = a.gradient + pb[j].y;
#pragma unroll
          for (int c = 0; c < kSize; ++c)
            if (c < tangent) RedAdd(dst + c, g[c]);
        }
      }

      if (a.output_jacobian && valid && active) {
        const int pos = __ldg(a.jacobian_pos + static_cast<size_t>(j) * a.n + tt);
        double* __restrict__ dst = a.jacobian_values + pos;
        const int stride = a.crs ? row_stride_crs : tangent;
        if (stride == kSize && tangent == kSize && ((pos & 1) == 0) &&
            ((kRes * kSize) % 2 == 0)) {
          // Dense cell, 16-byte aligned: one run of kRes * kSize doubles.
          double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
          for (int e = 0; e < kRes * kSize; e += 2) {
            d2[e / 2] = make_double2(B[e / kSize][e % kSize],
                                     B[(e + 1) / kSize][(e + 1) % kSize]);
          }
        } else {
#pragma unroll
          for (int r = 0; r < kRes; ++r)
#pragma unroll
            for (int c = 0; c < kSize; ++c)
              if (c < tangent) dst[r * stride + c] = B[r][c];
        }
      }
    };
    if (a.output_jacobian || a.output_gradient) {
      ForEachBlock(epilogue, std::make_index_sequence<kNB>{});
    }
    if (correct) {
#pragma unroll
      for (int r = 0; r < kRes; ++r) res[r] = res_corrected[r];
    }
  }

  if (valid && !ok) *a.status = 1;

  if (a.output_residuals && valid) {
    double* __restrict__ dst = a.residuals + __ldg(a.residual_pos + tt);
#pragma unroll
    for (int r = 0; r < kRes; ++r) dst[r] = res[r];
  }

  // Cost: warp shuffle, then one partial per thread block (summed in a fixed order
  // by the engine, so the cost is reproducible run to run).
  double c = (valid && ok) ? cost : 0.0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  __shared__ double warp_cost[kEvaluateThreads / 32];
  if (lane == 0) warp_cost[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
#pragma unroll
    for (int w = 0; w < kEvaluateThreads / 32; ++w) total += warp_cost[w];
    a.cost_partials[blockIdx.x] = total;
  }
}

// The launch thunk whose address goes through the C ABI.
template <typename Functor, typename Loss, int kRes, int... Ns>
int LaunchEvaluate(const cb200_launch_args* args, void* stream) {
  if (args->n <= 0) return 0;
  const int grid = (args->n + kEvaluateThreads - 1) / kEvaluateThreads;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (args->output_jacobian || args->output_gradient) {
    EvaluateKernel<true, Functor, Loss, kRes, Ns...><<<grid, kEvaluateThreads, 0, s>>>(*args);
  } else {
    EvaluateKernel<false, Functor, Loss, kRes, Ns...><<<grid, kEvaluateThreads, 0, s>>>(*args);
  }
  return static_cast<int>(cudaGetLastError());
}

template <typename Functor, typename Loss, int kRes, int... Ns>
cb200_residual_type MakeResidualType() {
  static_assert(sizeof...(Ns) >= 1 && sizeof...(Ns) <= CB200_MAX_PARAMETER_BLOCKS,
                "between 1 and 10 parameter blocks per residual block");
  cb200_residual_type t{};
  t.num_residuals = kRes;
  t.num_parameter_blocks = sizeof...(Ns);
  const int sizes[] = {Ns...};
  for (int i = 0; i < static_cast<int>(sizeof...(Ns)); ++i) t.parameter_block_sizes[i] = sizes[i];
  t.functor_size = static_cast<int32_t>(sizeof(Functor));
  t.loss_size = static_cast<int32_t>(sizeof(Loss));
  t.threads_per_block = kEvaluateThreads;
  t.launch = &LaunchEvaluate<Functor, Loss, kRes, Ns...>;
  return t;
}

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_EVALUATE_KERNEL_CUH_
