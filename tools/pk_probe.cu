// Packed-FP32 issue-rate probe for sm_100a: FFMA vs FFMA2 (fma.rn.f32x2), FADD vs FADD2, and each mixed with
// shared-memory loads. Reports warp-instructions per cycle per SM (clock64 around the loop, one CTA per SM-slot).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/pk_probe.cu -o tools/pk_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(1024) probe(float *out, long long *cyc, float a, float b, int iters) {
  __shared__ float2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i * 1e-6f, 1.f);
  __syncthreads();
  u64 r[8];
  float s[16];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = pk(threadIdx.x * 0.001f + i, 1.f + i);
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = threadIdx.x * 0.001f + i;
  const u64 pa = pk(a, a), pb = pk(b, b);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = ffma1(s[i], a, b);
      } else if (MODE == 1) {  // 8 FFMA2 (same flops as mode 0)
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = ffma2(r[i], pa, pb);
      } else if (MODE == 2) {  // 16 scalar FADD
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = fadd1(s[i], b);
      } else if (MODE == 3) {  // 8 FADD2
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = fadd2(r[i], pb);
      } else if (MODE == 4) {  // 8 FFMA2 + 4 LDS.64 + 4 IADD-ish (issue-slot sharing)
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = ffma2(r[i], pa, pb);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float2 v = sm[(threadIdx.x + 32 * i + it) & 2047];
          s[i] += v.x;
        }
      } else if (MODE == 5) {  // 16 FFMA + same extra
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = ffma1(s[i], a, b);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float2 v = sm[(threadIdx.x + 32 * i + it) & 2047];
          s[i] += v.x;
        }
      } else if (MODE == 6) {  // 8 FMUL2
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = fmul2(r[i], pa);
      } else if (MODE == 7) {  // 4 FFMA2 + 8 FFMA interleaved
#pragma unroll
        for (int i = 0; i < 4; i++) { r[i] = ffma2(r[i], pa, pb); s[2*i] = ffma1(s[2*i], a, b); s[2*i+1] = ffma1(s[2*i+1], a, b); }
      }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) { float2 f = *reinterpret_cast<float2 *>(&r[i]); acc += f.x + f.y; }
#pragma unroll
  for (int i = 0; i < 16; i++) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int fp_per_iter, int other_per_iter, int threads) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  float *out; long long *cyc;
  cudaMalloc(&out, (size_t)sms * threads * 4);
  cudaMalloc(&cyc, sms * 8);
  for (int rep = 0; rep < 2; rep++) probe<MODE><<<sms, threads>>>(out, cyc, 0.999f, 0.001f, iters);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; i++) c += h[i]; c /= sms;
  const double warps = threads / 32.0;
  const double fp = warps * iters * 4.0 * fp_per_iter, oth = warps * iters * 4.0 * other_per_iter;
  printf("{\"mode\": \"%s\", \"threads\": %d, \"cycles\": %.0f, \"fp_warp_instr_per_clk_per_sm\": %.3f, \"all_warp_instr_per_clk_per_sm\": %.3f, \"err\": \"%s\"}\n",
         name, threads, c, fp / c, (fp + oth) / c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {512, 1024}) {
    run<0>("16xFFMA", 16, 0, threads);
    run<1>("8xFFMA2", 8, 0, threads);
    run<2>("16xFADD", 16, 0, threads);
    run<3>("8xFADD2", 8, 0, threads);
    run<6>("8xFMUL2", 8, 0, threads);
    run<5>("16xFFMA+4LDS64+4FADD", 16, 8, threads);
    run<4>("8xFFMA2+4LDS64+4FADD", 8, 8, threads);
    run<7>("4xFFMA2+8xFFMA", 12, 0, threads);
  }
  return 0;
}
