// FP32 FMA peak micro-benchmark (BASELINE.md asks for a measured FP32 denominator for the direct convolution):
// every thread runs 16 independent FMA chains, fully unrolled; reports TFLOP/s (CUDA events).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fma_peak.cu -o tools/fma_peak.bin
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) fma_kernel(float *out, float a, float b, int iters) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 16; i++) r[i] = fmaf(r[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 4000;
  float *out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    fma_kernel<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * blocks * threads * (double)iters * 8 * 16;
    printf("{\"sms\": %d, \"ms\": %.3f, \"fp32_fma_tflops\": %.2f}\n", sms, ms, flop / ms / 1e9);
  }
  return 0;
}
