// fft_kernels.cuh -- batched C2C / R2C / C2R kernels for transforms that fit one CTA (N <= 16384).
//
// Replaces (reference, relative to /root/reference):
//   cfft_kernel      <- Clcfft::transform's reorder + log2(N) `fft` launches, cl_fft.cpp:138-161, 24-41
//   rfft_fwd_kernel  <- Clrfft forward: the above + `conv`, cl_fft.cpp:272-282, 178-191
//   rfft_inv_kernel  <- Clrfft inverse: `iconv` + the above, cl_fft.cpp:283-294, 192-205
// Each transform makes exactly one trip through HBM: 8N bytes in, 8N bytes out.
#pragma once

#include "fft_core.cuh"

namespace b2f {

// threads per CTA we aim for when several small transforms share a CTA
constexpr int kTargetThreads = 256;

template <int LOGN>
struct BatchGeom {
  using G = FftGeom<LOGN>;
  static constexpr int T = G::T;
  static constexpr int TPB = (T >= kTargetThreads) ? 1 : (kTargetThreads / T);  // transforms per CTA
  static constexpr int THREADS = T * TPB;
  static constexpr int SMEM_BYTES = TPB * G::SMEM * (int)sizeof(float2);
};

// ---- complex to complex ------------------------------------------------------------------------
// in/out: [batch][N] float2, may alias (each CTA gathers its whole transform before it scatters).
// scale: 1/N for the reference's forward transform (cl_fft.cpp:39-40), 1 for the inverse.
template <int LOGN, bool INV>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS)
    cfft_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, int batch, float scale) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * FftGeom<LOGN>::SMEM;
  auto load = [&](int idx, int) { return active ? src[idx] : make_float2(0.f, 0.f); };
  auto store = [&](int idx, float2 v, int) {
    if (active) dst[idx] = make_float2(v.x * scale, v.y * scale);
  };
  fft_run<LOGN, INV>(load, store, sm, tw, t, CtaSync());
}

// ---- real to complex (forward) -------------------------------------------------------------------
// in: [batch][2N] float (read as N packed float2), out: [batch][N] float2, may alias.
// Output convention of the reference: element 0 = (DC, Nyquist)/size packed, element k =
// 2 X[k]/size, element N/2 left as the plain FFT value (the reference's split never visits it, Q3).
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS)
    rfft_fwd_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ w2,
                    int batch) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * FftGeom<LOGN>::SMEM;
  const float scale = 1.0f / (float)N;
  auto load = [&](int idx, int) { return active ? src[idx] : make_float2(0.f, 0.f); };
  auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = make_float2(v.x * scale, v.y * scale); };
  fft_run<LOGN, false, true>(load, store, sm, tw, t, CtaSync());
  __syncthreads();
  if (!active) return;
  // split: pairs (i, N-i), i in [1, N/2); elements 0 and N/2 handled apart
  for (int i = t; i <= N / 2; i += B::T) {
    if (i == 0) {
      dst[0] = rfft_dc<false>(sm[pad_idx(0)]);
    } else if (i == N / 2) {
      dst[i] = sm[pad_idx(i)];
    } else {
      float2 ci = sm[pad_idx(i)], cj = sm[pad_idx(N - i)];
      rfft_pair<false>(ci, cj, __ldg(&w2[i]));
      dst[i] = ci;
      dst[N - i] = cj;
    }
  }
}

// ---- complex to real (inverse) -------------------------------------------------------------------
// in: [batch][N] float2 in the layout rfft_fwd_kernel writes, out: [batch][2N] float, may alias.
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS)
    rfft_inv_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ w2,
                    int batch) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * FftGeom<LOGN>::SMEM;
  if (active) {
    for (int i = t; i <= N / 2; i += B::T) {
      if (i == 0) {
        sm[pad_idx(0)] = rfft_dc<true>(src[0]);
      } else if (i == N / 2) {
        sm[pad_idx(i)] = src[i];
      } else {
        float2 ci = src[i], cj = src[N - i];
        rfft_pair<true>(ci, cj, __ldg(&w2[i]));
        sm[pad_idx(i)] = ci;
        sm[pad_idx(N - i)] = cj;
      }
    }
  }
  __syncthreads();
  auto load = [&](int idx, int) { return sm[pad_idx(idx)]; };
  auto store = [&](int idx, float2 v, int) {
    if (active) dst[idx] = v;
  };
  fft_run<LOGN, true, false, true>(load, store, sm, tw, t, CtaSync());
}

}  // namespace b2f
