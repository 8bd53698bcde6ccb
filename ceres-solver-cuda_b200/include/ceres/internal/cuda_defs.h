// HOST_DEVICE / DEVICE / DEVICE_CODE, same contract as the reference's
// include/ceres/internal/cuda_defs.h:8-22: cost functors and loss functions mark
// every method that runs during evaluation with HOST_DEVICE.
#ifndef CERES_B200_INTERNAL_CUDA_DEFS_H_
#define CERES_B200_INTERNAL_CUDA_DEFS_H_

// In CUDA translation units HOST_DEVICE also forces inlining.  The evaluation kernel relies
// on the functor body being inlined into it: the Jets' sparsity masks (ceres/jet.h) and the
// unit seeds are compile-time constants only then.  Left to its heuristics the compiler keeps
// a large functor that several kernel variants call (e.g. RelativePoseError with two
// derivative passes) as a separate function taking Jets through local memory - measured 10x
// slower on the 10 M-edge pose graph (58 ms against 5.5 ms).
#ifdef __CUDACC__
// (the attribute, not the __forceinline__ keyword: user code may also say `inline`)
#define HOST_DEVICE __host__ __device__ __attribute__((always_inline))
#define DEVICE __device__
#define CERES_B200_INLINE __inline__
#else
#define HOST_DEVICE
#define DEVICE
#define CERES_B200_INLINE inline __attribute__((always_inline))
#endif

#ifdef __CUDA_ARCH__
#define DEVICE_CODE
#endif

#endif  // CERES_B200_INTERNAL_CUDA_DEFS_H_
