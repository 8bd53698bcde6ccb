// HOST_DEVICE / DEVICE / DEVICE_CODE, same contract as the reference's
// include/ceres/internal/cuda_defs.h:8-22: cost functors and loss functions mark
// every method that runs during evaluation with HOST_DEVICE.
#ifndef CERES_B200_INTERNAL_CUDA_DEFS_H_
#define CERES_B200_INTERNAL_CUDA_DEFS_H_

#ifdef __CUDACC__
#define HOST_DEVICE __host__ __device__
#define DEVICE __device__
#define CERES_B200_INLINE __forceinline__
#else
#define HOST_DEVICE
#define DEVICE
#define CERES_B200_INLINE inline __attribute__((always_inline))
#endif

#ifdef __CUDA_ARCH__
#define DEVICE_CODE
#endif

#endif  // CERES_B200_INTERNAL_CUDA_DEFS_H_
