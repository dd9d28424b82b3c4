// Development probe: where do the cycles of fft_cluster_kernel go? Builds the kernel with per-phase
// clock64() counters (thread 0 of CTA 0) and prints the average cycles per transform of each phase.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB2F_PHASE_PROBE -I opencl_fft_b200/csrc \
//        tools/cluster_phase_probe.cu -o /tmp/probe && /tmp/probe
#include <cmath>
#include <cstdio>
#include <vector>

#include "fft_cluster.cuh"
using namespace b2f;

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                          \
      return 1;                                                               \
    }                                                                         \
  } while (0)

int main() {
  constexpr int L1 = 11;
#ifndef PROBE_S
#define PROBE_S 4
#endif
  using C = ClusterGeom<L1, PROBE_S>;
  const int N = C::N, batch = 1024;
  float2 *in, *out, *tw1, *twl, *w2;
  CK(cudaMalloc(&in, (size_t)batch * N * 8));
  CK(cudaMalloc(&out, (size_t)batch * N * 8));
  CK(cudaMemset(in, 0, (size_t)batch * N * 8));
  CK(cudaMalloc(&tw1, 1 << 20));
  CK(cudaMalloc(&twl, (size_t)N * 8));
  CK(cudaMalloc(&w2, (size_t)N * 8));
  CK(cudaMemset(tw1, 0, 1 << 20));
  CK(cudaMemset(twl, 0, (size_t)N * 8));
  CK(cudaMemset(w2, 0, (size_t)N * 8));
  auto kern = fft_cluster_kernel<L1, false, true, PROBE_S>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PROBE_S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(PROBE_S * 64);
  int ncl = 0;
  CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
  printf("max active clusters %d, smem %d B, threads %d\n", ncl, C::SMEM_BYTES, C::THREADS);
  cfg.gridDim = dim3(PROBE_S * ncl);
  const float scale = 1.0f / N;
  for (int it = 0; it < 3; it++) {
    unsigned long long zero[8] = {0};
    CK(cudaMemcpyToSymbol(g_phase_probe, zero, sizeof(zero)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, kern, (const float2 *)in, out, (const float2 *)tw1, (const float2 *)twl,
                          (const float2 *)w2, batch, scale));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long p[8];
    CK(cudaMemcpyFromSymbol(p, g_phase_probe, sizeof(p)));
    const int iters = (batch + ncl - 1) / ncl;
    printf("run %d: %.3f ms, %.0f GB/s; per transform on CTA 0 (cycles): step1 %llu | barrier1 %llu | gather+tw %llu | "
           "barrier2 %llu | step3+split+store %llu\n",
           it, ms, (double)batch * N * 16 / ms / 1e6, p[0] / iters, p[1] / iters, p[2] / iters, p[3] / iters, p[4] / iters);
  }
  return 0;
}
