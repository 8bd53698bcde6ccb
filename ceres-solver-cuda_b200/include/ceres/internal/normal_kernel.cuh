// y += J'(J x) on the Jacobian the evaluation kernel left in HBM, one instantiation per
// <kNumResiduals, Ns...> (the functor does not matter: the cells are plain numbers).  It is
// the product conjugate gradients on the normal equations spend their time in
// (reference: CgnrSolver / CudaCgnrSolver, internal/ceres/cgnr_solver.cc:190-330, two cuSPARSE
// SpMVs per iteration over a compressed-row copy of J).  Reaches the engine as a second thunk
// of cb200_residual_type; the engine's size-generic table-walk kernels (csrc/engine.cu) remain
// the fallback for every structure this kernel does not take.
//
// Preconditions (checked by the engine, cb200_normal_args): block-sparse values, no manifold
// and no constant block in the type (tangent == ambient size, gradient offset == state offset),
// and cell positions that are arithmetic progressions with step kRes * Size(j), i.e. the 32
// cells of a warp's tile are one contiguous run per argument (bundle adjustment after the
// Schur ordering).
//
// Design: a warp owns tiles of 32 consecutive residual blocks (grid stride).  Per tile it reads
// its two contiguous runs of cells with fully coalesced 16-byte cp.async copies into shared
// memory (the table-walk kernel reads a cell per thread at a 144-byte stride: 32 cache lines
// per instruction, L1-tag bound at 25 % of the HBM roof), gathers the x entries of its
// parameter blocks exactly like the evaluation kernel gathers parameters (aligned 16-byte
// windows, cooperative), keeps one tile in flight while the previous one is consumed, and
// adds J' t with the evaluation kernel's staged reductions (consecutive lanes on consecutive
// addresses).  Sizes are compile-time, so the 24 multiply-adds per block are straight-line.
#ifndef CERES_B200_INTERNAL_NORMAL_KERNEL_CUH_
#define CERES_B200_INTERNAL_NORMAL_KERNEL_CUH_

#include <cuda_runtime.h>

#include <cstdint>

#include "ceres_b200.h"
// (included from the end of ceres/internal/evaluate_kernel.cuh, whose helpers it uses)

namespace ceres {
namespace internal {

constexpr int kNormalThreads = 128;

// ---- TMA bulk load global -> shared, completion on an mbarrier (cp.async.bulk; SASS UBLKCP):
// the copy engine writes shared memory by itself, so the 6 KB of cells per tile cost no LSU
// wavefronts (with 16-byte cp.async they cost one wavefront per returned sector, four times
// the 128-byte ideal, and made this kernel L1-bound at 88 %: profiles/r2_normal_product_*).
__device__ __forceinline__ unsigned SharedAddress(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarrierInit(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SharedAddress(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void MbarrierExpect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SharedAddress(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarrierWait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      "  .reg .pred p;\n"
      "WAIT_%=:\n"
      "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "  @!p bra WAIT_%=;\n"
      "}" ::"r"(SharedAddress(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void BulkLoad(void* smem, const void* gmem, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(SharedAddress(smem)),
      "l"(gmem), "r"(bytes), "r"(SharedAddress(bar))
      : "memory");
}

template <int kRes, int... Ns>
struct NormalPlan {
  using Dims = BlockDims<Ns...>;
  static constexpr int kNB = Dims::kNumBlocks;
  static constexpr int kNP = Dims::kNumParameters;
  static constexpr int kXDoubles = 32 * 2 * Dims::PitchChunksBefore(kNB);  // one x stage of a warp
  static constexpr int kJDoubles = 32 * kRes * kNP;                    // one cell stage of a warp
  // staged reductions: [lane][pitch] sums and, per element, its destination (an int)
  static constexpr int kGradientDoubles = 48 * StagePitch(Dims::MaxSize());
  static constexpr int kStageDoubles = kJDoubles;
  // [x stage 0 | x stage 1 | cell stage 0 | cell stage 1 | reduction stage | two mbarriers]
  // (the reductions have their own stage: the scaling operation writes its cells back from
  // the cell stage with a bulk store that is still reading it while the sums are staged)
  static constexpr int kWarpDoubles = 2 * kXDoubles + 2 * kStageDoubles + kGradientDoubles + 2;
  static constexpr int kBytes = (kNormalThreads / 32) * kWarpDoubles * 8;
  static constexpr int kCtas = (228 * 1024) / (kBytes + 1024) >= 3 ? 3
                               : ((228 * 1024) / (kBytes + 1024) >= 2 ? 2 : 1);
  static constexpr bool kFits = kBytes <= 227 * 1024 && (kRes * kNP) % 2 == 0;
};

// Operations of the kernel (cb200_normal_args::op).  All of them walk the cells once.
constexpr int kNormalOpNormal = CB200_NORMAL_OP_NORMAL;          // y += J'(J x)
constexpr int kNormalOpLeft = CB200_NORMAL_OP_LEFT;              // y += J' w
constexpr int kNormalOpRight = CB200_NORMAL_OP_RIGHT;            // w  = J x
constexpr int kNormalOpColumnNorm = CB200_NORMAL_OP_COLUMN_NORM; // y += squared column norms
constexpr int kNormalOpScaleNorm = CB200_NORMAL_OP_SCALE_NORM;   // J <- J diag(x) in place, then
                                                                 // y += its squared column norms

template <int kOp, int kRes, int... Ns>
__global__ void __launch_bounds__(kNormalThreads, NormalPlan<kRes, Ns...>::kCtas)
    NormalProductKernel(const cb200_normal_args a) {
  constexpr bool kGatherX = kOp == kNormalOpNormal || kOp == kNormalOpRight || kOp == kNormalOpScaleNorm;
  constexpr bool kReduce = kOp != kNormalOpRight;
  using Dims = BlockDims<Ns...>;
  using Plan = NormalPlan<kRes, Ns...>;
  constexpr int kNB = Plan::kNB;
  constexpr int kNP = Plan::kNP;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* const wbuf = reinterpret_cast<double*>(smem) + warp * Plan::kWarpDoubles;
  auto xstage = [&](int s) { return wbuf + s * Plan::kXDoubles; };
  auto jstage = [&](int s) { return wbuf + 2 * Plan::kXDoubles + s * Plan::kStageDoubles; };
  double* const gbuf = wbuf + 2 * Plan::kXDoubles + 2 * Plan::kStageDoubles;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(
      wbuf + 2 * Plan::kXDoubles + 2 * Plan::kStageDoubles + Plan::kGradientDoubles);
  // bulk loads need 16-byte aligned sources: the cell size is even, so only the bases matter
  bool bulk = true;
#pragma unroll
  for (int j = 0; j < kNB; ++j)
    bulk = bulk && ((reinterpret_cast<uintptr_t>(a.values + a.base[j]) & 15) == 0);
  if (bulk && lane == 0) {
    MbarrierInit(bars, 1);
    MbarrierInit(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int n = a.n;
  const int num_tiles = (n + 31) / 32;
  const int warps = gridDim.x * (kNormalThreads / 32);
  const int first_tile = blockIdx.x * (kNormalThreads / 32) + warp;

  auto load_offsets = [&](int tile, int (&soff)[kNB]) {
    const int rb = min(tile * 32 + lane, n - 1);
#pragma unroll
    for (int j = 0; j < kNB; ++j)
      soff[j] = tile < num_tiles ? __ldg(a.offset + static_cast<size_t>(j) * n + rb) : 0;
  };
  // copies of one tile: the x windows of its parameter blocks and its runs of cells
  auto prefetch = [&](int s, int tile, const int (&soff)[kNB]) {
    if (tile < num_tiles) {
      double* xs = xstage(s);
#pragma unroll
      for (int j = 0; j < (kGatherX ? kNB : 0); ++j) {
        const int kW = Dims::WindowChunks(j);
#pragma unroll
        for (int it = 0; it < kW; ++it) {
          const int e = it * 32 + lane;
          const int owner = e / kW;
          const int c = e - owner * kW;
          const int so = __shfl_sync(0xffffffffu, soff[j], owner);
          CpAsync16(xs + 2 * (32 * Dims::PitchChunksBefore(j) + owner * Dims::WindowPitch(j) + c),
                    a.x + ((so & ~1) + 2 * c));
        }
      }
      if constexpr (kOp == kNormalOpLeft) {
        // the rows of the residual-space vector ride the same pipeline (the x stage is free)
        const int rbw = min(tile * 32 + lane, n - 1);
#pragma unroll
        for (int r = 0; r < kRes; ++r)
          CpAsync8(xs + lane * kRes + r, a.w + (a.residual_base + static_cast<int64_t>(rbw) * kRes + r));
      }
      double* js = jstage(s);
      const int rb0 = tile * 32;
      const int blocks = min(32, n - rb0);
      if (bulk && lane == 0) MbarrierExpect(bars + s, blocks * kRes * kNP * 8);
#pragma unroll
      for (int j = 0; j < kNB; ++j) {
        const int kCell = kRes * Dims::Size(j);  // doubles per cell, constant after unrolling
        const double* src = a.values + (a.base[j] + static_cast<int64_t>(rb0) * kCell);
        double* dst = js + 32 * kRes * Dims::Offset(j);
        const int doubles = blocks * kCell;
        if (bulk) {
          if (lane == 0) BulkLoad(dst, src, doubles * 8, bars + s);
        } else {
#pragma unroll
          for (int it = 0; it < kCell; ++it) {
            const int e = it * 32 + lane;
            if (e < doubles) CpAsync8(dst + e, src + e);
          }
        }
      }
    }
    CpAsyncCommit();
  };

  int soff_cur[kNB], soff_next[kNB];
  load_offsets(first_tile, soff_cur);
  prefetch(0, first_tile, soff_cur);
  load_offsets(first_tile + warps, soff_next);

  int k = 0;
  for (int tile = first_tile; tile < num_tiles; tile += warps, ++k) {
    const int s = k & 1;
    int soff_issue[kNB];
#pragma unroll
    for (int j = 0; j < kNB; ++j) soff_issue[j] = soff_next[j];
    if constexpr (kOp == kNormalOpScaleNorm) {
      // last tile's write-back still reads stage s ^ 1
      if (bulk && lane == 0) BulkWaitRead();
      __syncwarp();
    }
    prefetch(s ^ 1, tile + warps, soff_issue);
    load_offsets(tile + 2 * warps, soff_next);
    CpAsyncWait<1>();
    __syncwarp();
    if (bulk) MbarrierWait(bars + s, (k >> 1) & 1);  // stage s is used every other tile

    const int rb = tile * 32 + lane;
    const bool valid = rb < n;
    const double* xs = xstage(s);
    const double* js = jstage(s);
    // this lane's cells and x entries -> registers (block sizes reach the loops below as
    // compile-time constants: with a loop variable in their place the arrays are indexed
    // dynamically and live in local memory - 288 bytes of stack and a third of the stall
    // samples in profiles/r2_normal_product_ncu_summary.txt)
    double J[kRes][kNP], x[kNP];
    ForEachBlock(
        [&](auto jc) {
          constexpr int j = decltype(jc)::value;
          constexpr int kS = Dims::Size(j);
          constexpr int kO = Dims::Offset(j);
          // (16-byte reads: the lane stride kRes * kS * 8 is a multiple of 16, conflict free)
          const double2* cell =
              reinterpret_cast<const double2*>(js + 32 * kRes * kO + lane * kRes * kS);
#pragma unroll
          for (int e = 0; e < kRes * kS; e += 2) {
            const double2 v = cell[e / 2];
            J[e / kS][kO + e % kS] = valid ? v.x : 0.0;
            J[(e + 1) / kS][kO + (e + 1) % kS] = valid ? v.y : 0.0;
          }
          if constexpr (kGatherX) {
            // the x window in 16-byte pieces (odd lane pitch: conflict free), picked by parity
            constexpr int kW = Dims::WindowChunks(j);
            const double2* xw = reinterpret_cast<const double2*>(xs) +
                                (32 * Dims::PitchChunksBefore(j) + lane * Dims::WindowPitch(j));
            double2 w[kW];
#pragma unroll
            for (int c = 0; c < kW; ++c) w[c] = xw[c];
            const bool odd = soff_cur[j] & 1;
#pragma unroll
            for (int c = 0; c < kS; ++c) {
              const double even_pick = (c & 1) ? w[c / 2].y : w[c / 2].x;
              const double odd_pick = ((c + 1) & 1) ? w[(c + 1) / 2].y : w[(c + 1) / 2].x;
              x[kO + c] = valid ? (odd ? odd_pick : even_pick) : 0.0;
            }
          }
        },
        std::make_index_sequence<kNB>{});
    // rows of this block in the residual-space vector (residual positions are affine)
    double* const wrow = a.w + (a.residual_base + static_cast<int64_t>(rb) * kRes);
    double t[kRes];
#pragma unroll
    for (int r = 0; r < kRes; ++r) t[r] = 0.0;
    if constexpr (kOp == kNormalOpNormal || kOp == kNormalOpRight) {
#pragma unroll
      for (int r = 0; r < kRes; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < kNP; ++c) acc += J[r][c] * x[c];
        t[r] = acc;
      }
    } else if constexpr (kOp == kNormalOpLeft) {
      if (valid) {
#pragma unroll
        for (int r = 0; r < kRes; ++r) t[r] = xs[lane * kRes + r];
      }
    }
    if constexpr (kOp == kNormalOpRight) {
      if (valid) {
#pragma unroll
        for (int r = 0; r < kRes; ++r) wrow[r] = t[r];
      }
    }
    if constexpr (kOp == kNormalOpScaleNorm) {
      // scale the columns, put the cells back where they were staged and send the runs home
      ForEachBlock(
          [&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int kS = Dims::Size(j);
            constexpr int kO = Dims::Offset(j);
            double2* cell = reinterpret_cast<double2*>(jstage(s) + 32 * kRes * kO + lane * kRes * kS);
#pragma unroll
            for (int r = 0; r < kRes; ++r)
#pragma unroll
              for (int c = 0; c < kS; ++c) J[r][kO + c] *= x[kO + c];
            if (valid) {
#pragma unroll
              for (int e = 0; e < kRes * kS; e += 2)
                cell[e / 2] = make_double2(J[e / kS][kO + e % kS], J[(e + 1) / kS][kO + (e + 1) % kS]);
            }
          },
          std::make_index_sequence<kNB>{});
      if (bulk) {
        FenceProxyAsyncShared();
        __syncwarp();
        if (lane == 0) {
          const int rb0 = tile * 32;
          const int blocks = min(32, n - rb0);
#pragma unroll
          for (int j = 0; j < kNB; ++j) {
            const int kCell = kRes * Dims::Size(j);
            BulkStore(a.values + (a.base[j] + static_cast<int64_t>(rb0) * kCell),
                      jstage(s) + 32 * kRes * Dims::Offset(j), blocks * kCell * 8);
          }
          BulkCommit();
        }
      } else {
        __syncwarp();
        const int rb0 = tile * 32;
        const int blocks = min(32, n - rb0);
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
          const int kCell = kRes * Dims::Size(j);
          double* dst = a.values + (a.base[j] + static_cast<int64_t>(rb0) * kCell);
          const double* src = jstage(s) + 32 * kRes * Dims::Offset(j);
          for (int e = lane; e < blocks * kCell; e += 32) dst[e] = src[e];
        }
      }
    }
    if constexpr (kReduce) {
      int* obuf = reinterpret_cast<int*>(gbuf + 32 * StagePitch(Dims::MaxSize()));
      ForEachBlock(
          [&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int kS = Dims::Size(j);
            constexpr int kO = Dims::Offset(j);
            constexpr int kPitch = StagePitch(kS);
            // sums and, per element, its destination (-1: nothing to add; pad slots too), so
            // a round is two shared loads, one address and the red - no division
#pragma unroll
            for (int c = 0; c < kS; ++c) {
              double acc = 0.0;
#pragma unroll
              for (int r = 0; r < kRes; ++r)
                acc += J[r][kO + c] * ((kOp == kNormalOpColumnNorm || kOp == kNormalOpScaleNorm)
                                           ? J[r][kO + c]
                                           : t[r]);
              gbuf[lane * kPitch + c] = acc;
              obuf[lane * kPitch + c] = valid ? soff_cur[j] + c : -1;
            }
            if constexpr (kPitch > kS) obuf[lane * kPitch + kS] = -1;
            __syncwarp();
#pragma unroll
            for (int it = 0; it < kPitch; ++it) {
              const int d = obuf[it * 32 + lane];
              RedAddIf(d >= 0, a.y + d, gbuf[it * 32 + lane]);
            }
            __syncwarp();
          },
          std::make_index_sequence<kNB>{});
    }
    // the stage was read and rewritten through the generic proxy; the copy engine (async
    // proxy) overwrites it two tiles from now
    if (bulk) FenceProxyAsyncShared();
#pragma unroll
    for (int j = 0; j < kNB; ++j) soff_cur[j] = soff_issue[j];
  }
  CpAsyncWait<0>();
  if constexpr (kOp == kNormalOpScaleNorm) {
    if (bulk && lane == 0) BulkWaitAll();
  }
}

template <int kRes, int... Ns>
int LaunchNormalProduct(const cb200_normal_args* args, void* stream) {
  using Plan = NormalPlan<kRes, Ns...>;
  if constexpr (!Plan::kFits) {
    return -1;  // the engine falls back to its table-walk kernel
  } else {
    if (args->n <= 0) return 0;
    int device = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    using Kernel = void (*)(const cb200_normal_args);
    Kernel kernel = nullptr;
    switch (args->op) {
      case kNormalOpNormal: kernel = NormalProductKernel<kNormalOpNormal, kRes, Ns...>; break;
      case kNormalOpLeft: kernel = NormalProductKernel<kNormalOpLeft, kRes, Ns...>; break;
      case kNormalOpRight: kernel = NormalProductKernel<kNormalOpRight, kRes, Ns...>; break;
      case kNormalOpColumnNorm: kernel = NormalProductKernel<kNormalOpColumnNorm, kRes, Ns...>; break;
      case kNormalOpScaleNorm: kernel = NormalProductKernel<kNormalOpScaleNorm, kRes, Ns...>; break;
      default: return -1;
    }
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int tiles = (args->n + 31) / 32;
    const int needed = (tiles + kNormalThreads / 32 - 1) / (kNormalThreads / 32);
    const int wanted = (sms > 0 ? sms : 148) * Plan::kCtas;
    kernel<<<needed < wanted ? needed : wanted, kNormalThreads, Plan::kBytes,
             static_cast<cudaStream_t>(stream)>>>(*args);
    return static_cast<int>(cudaGetLastError());
  }
}

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_NORMAL_KERNEL_CUH_
