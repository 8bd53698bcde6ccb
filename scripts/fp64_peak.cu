// FP64 FMA throughput of the device (the second roof of the evaluation kernel; BASELINE.md
// section 2 asks for a measured number).  Every thread runs 16 independent DFMA chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak scripts/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) Dfma(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    Dfma<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flop = 2.0 * 16 * iters * 256.0 * blocks;
  std::printf("{\"fp64_fma_tflops\": %.2f, \"sms\": %d, \"clock_mhz\": %d, \"ms\": %.3f}\n",
              flop / (best * 1e-3) / 1e12, p.multiProcessorCount, p.clockRate / 1000, best);
  return 0;
}
