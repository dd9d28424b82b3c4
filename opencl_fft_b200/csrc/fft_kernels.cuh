// fft_kernels.cuh -- batched C2C / R2C / C2R kernels for transforms that fit one CTA (N <= 16384).
//
// Replaces (reference, relative to /root/reference):
//   cfft_kernel (N >= 64), fft_thread_kernel<kThreadComplex> (N <= 32)
//                      <- Clcfft::transform's reorder + log2(N) `fft` launches, cl_fft.cpp:138-161, 24-41
//   rfft_fwd_reg_kernel, fft_thread_kernel<kThreadRealFwd>
//                      <- Clrfft forward: the above + `conv`, cl_fft.cpp:272-282, 178-191
//   rfft_inv_reg_kernel, fft_thread_kernel<kThreadRealInv>
//                      <- Clrfft inverse: `iconv` + the above, cl_fft.cpp:283-294, 192-205
// Each transform makes exactly one trip through HBM: 8N bytes in, 8N bytes out.
#pragma once

#include "fft_core.cuh"
#include "tma_utils.cuh"

#ifndef B2F_SPLIT_SHFL
#define B2F_SPLIT_SHFL 1
#endif

namespace b2f {

// threads per CTA we aim for when several small transforms share a CTA
constexpr int kTargetThreads = 256;

// N >= 64 (N <= 32: one thread per transform, ThreadGeom at the end of this file)
template <int LOGN>
struct BatchGeom {
  using G = FftGeom<LOGN>;
  static constexpr int T = G::T;
  static constexpr int TPB = (T >= kTargetThreads) ? 1 : (kTargetThreads / T);  // transforms per CTA
  static constexpr int THREADS = T * TPB;
  static constexpr int ROW = G::SMEM;
  static constexpr int SMEM_BYTES = TPB * ROW * (int)sizeof(float2);
  static constexpr int MIN_BLOCKS = 1024 / THREADS;  // caps registers at 64/thread: 32 resident warps per SM
  // real transforms of N >= 8192: the CTA asks the TMA engine for the transform of the CTA `ahead` further on (the one
  // that takes its place when it retires; the host passes the number of co-resident CTAs) while it works
  static constexpr bool PREFETCH = LOGN >= 13;
};
// A CTA's loads are all issued in its first few hundred cycles and then waited for; with 1-2 CTAs of 512-1024 threads
// per SM there is little else in flight meanwhile. Prefetched into L2 while the predecessor computes, they come back
// at L2 latency instead. Measured (fraction of the HBM peak, r2c / c2r): N = 8192 0.64 / 0.58 -> 0.67 / 0.61,
// N = 16384 0.49 / 0.46 -> 0.53 / 0.50; no gain at N <= 4096, and the complex kernels LOSE with it (N = 1024..4096:
// 1.04 -> 0.93..0.97; N = 8192: 0.72 -> 0.71 -- ncu of cfft_kernel<13>: the LSU data pipe, 59 % busy over the whole
// kernel and saturated in the load and store phases, is what limits them, not the latency of the loads), so only
// the real kernels use it.
template <int LOGN>
__device__ __forceinline__ void prefetch_successor(const float2 *in, int ahead, int batch) {
  using B = BatchGeom<LOGN>;
  if constexpr (B::PREFETCH) {
    constexpr int N = 1 << LOGN, CHUNKS = B::TPB * N * (int)sizeof(float2) / 4096;
    const long long first = ((long long)blockIdx.x + ahead) * B::TPB;  // first transform of the successor CTA
    if (ahead > 0 && (int)threadIdx.x < CHUNKS && first * N + ((int)threadIdx.x + 1) * 512 <= (long long)batch * N)
      tma::prefetch_l2(in + first * N + threadIdx.x * 512, 4096);
  }
}

// ---- complex to complex ------------------------------------------------------------------------
// in/out: [batch][N] float2, may alias (each CTA gathers its whole transform before it scatters).
// scale: 1/N for the reference's forward transform (cl_fft.cpp:39-40), 1 for the inverse.
template <int LOGN, bool INV>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS, BatchGeom<LOGN>::MIN_BLOCKS)
    cfft_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, int batch, float scale) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * BatchGeom<LOGN>::ROW;
  auto load = [&](int idx, int) { return active ? src[idx] : make_float2(0.f, 0.f); };
  auto store = [&](int idx, float2 v, int) {
    if (active) dst[idx] = cscale(v, scale);
  };
  fft_run<LOGN, INV>(load, store, sm, tw, t, CtaSync());
}

// =====================================================================================================
// Register-level real transforms for N >= 64 (schedules whose first and last passes leave every thread
// with the E = 16 (8 at N = 64) values X[t + m*T], m = 0..E-1, T = N/E; written below for E = 16). The pair partner of element (t, m) is element
// (T - t, 15 - m) [(0, 16 - m) for t = 0], so the split / unsplit needs ONE partner thread: the two swap
// half of their values through a small shared-memory staging area, each evaluates its 8 pairs once (the inverse
// reads the partner's inputs from global memory instead and only hands the results over).
// Against a split pass over the finished transform in shared memory this drops a full shared-memory round trip
// and halves the split arithmetic: 0.5, the 1/N scaling and the quarter turn are folded into the table
//   hw[i] = 0.5 * scale * i * w2[i]  (forward)      hw[i] = conj(0.5 * i * w2[i])  (inverse)
// so that   out_i = hs*S + hw*D,  out_j = conj(hs*S - hw*D),  S = A + conj(B), D = conj(B) - A,
// the same algebra as the reference's conv/iconv kernels (cl_fft.cpp:178-205), 12 instructions per pair.
// =====================================================================================================
// folded split twiddle of element t + m*T, T = N/E, from the one of element t: w2[t + m*T] = w2[t] * exp(-+ i pi m/E),
// a compile-time constant per m (one table load per thread instead of E/2; the loads share the LSU data pipe with
// the shared-memory exchanges, the multiplies go to the idle FP32 pipe)
template <bool INV, int E>
__device__ __forceinline__ float2 split_tw(float2 hw0, int m) {
  static_assert(E == 16 || E == 8, "");
  constexpr float kC[8] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                           0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                           0.19509032201612826785f};
  constexpr float kS[8] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                           0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                           0.98078528040323044913f};
  const int k = m * (16 / E);  // angle pi m/E in units of pi/16
  return m == 0 ? hw0 : cmulc<INV>(hw0, kC[k], kS[k]);
}

template <int LOGN>
struct RegSplitGeom {
  using G = FftGeom<LOGN>;
  static constexpr Sched S = G::S;
  static constexpr bool OK = (LOGN >= 6) && (G::E == 16 || G::E == 8) && S.radix[S.npass - 1] == G::E;
  static constexpr int T = G::T;
  static constexpr int R0 = S.radix[0];
};

// forward: in [batch][2N] float, out [batch][N] float2 (may alias). hw: folded table (scale included).
// ZEROPAD (Clpconv::push_ir, cl_conv.cpp:361-380): the transform's input is N reals followed by N zeros -- only the
// first N/2 packed elements are read, and they may be 4-byte aligned only (`pair_ok` false).
template <int LOGN, bool ZEROPAD>
__device__ __forceinline__ void rfft_fwd_reg_body(const float2 *src, float2 *dst, bool active, bool pair_ok, float2 *sm,
                                                  const float2 *__restrict__ tw, const float2 *__restrict__ hw, int t,
                                                  float scale) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T, E = FftGeom<LOGN>::E, H = E / 2;
  float2 x[E];
  auto load = [&](int idx, int) {
    if constexpr (ZEROPAD) {
      if (!active || idx >= N / 2) return make_float2(0.f, 0.f);
      const float *xr = reinterpret_cast<const float *>(src);
      return pair_ok ? __ldcs(src + idx) : make_float2(__ldcs(xr + 2 * idx), __ldcs(xr + 2 * idx + 1));
    } else {
      return active ? __ldcs(src + idx) : make_float2(0.f, 0.f);
    }
  };
  auto store = [&](int, float2 v, int slot) { x[slot] = v; };  // last pass: slot == m, value X[t + m*T]
  fft_run<LOGN, false>(load, store, sm, tw, t, CtaSync());
  const int pt = (t == 0) ? 0 : T - t;  // partner thread (itself for t = 0 and t = T/2)
  // T <= 32 (N <= 512): the partner sits in the same warp -- its high members arrive by shuffle (no staging in shared
  // memory, no CTA barriers; B2F_SPLIT_SHFL=0 keeps the staged exchange for comparison)
  constexpr bool SHFL = B2F_SPLIT_SHFL && T <= 32;
  float2 ph[H];  // ph[m]: the partner's value X[pt + pm * T], pm = E - 1 - m (E - m for t = 0: the thread's own)
  if constexpr (SHFL) {
    const int plane = ((int)threadIdx.x & 31 & ~(T - 1)) | pt;
#pragma unroll
    for (int m = 0; m < H; m++) {
      ph[m].x = __shfl_sync(0xffffffffu, x[E - 1 - m].x, plane);
      ph[m].y = __shfl_sync(0xffffffffu, x[E - 1 - m].y, plane);
      if (t == 0 && m > 0) ph[m] = x[E - m];
    }
  } else {
    __syncthreads();  // every thread is past its last gather: sm becomes the staging area [E/2][T]
#pragma unroll
    for (int m = H; m < E; m++) sm[(m - H) * T + t] = x[m];
    __syncthreads();
  }
  if (!active) return;
  const float hs = 0.5f * scale;
  // measured: deriving the split twiddles pays up to N = 4096 (r2c 4096: 89 -> 92 % of the HBM peak); above, the
  // 512- and 1024-thread CTAs are bound by their phases, not by the LSU pipe, and the table loads are faster
  constexpr bool DERIVE = LOGN <= 12;
  const float2 hw0 = __ldg(&hw[t]);
#pragma unroll
  for (int m = 0; m < H; m++) {
    if (m == 0 && t == 0) {
      dst[0] = make_float2((x[0].x + x[0].y) * hs, (x[0].x - x[0].y) * hs);   // packed (DC, Nyquist)
      dst[N / 2] = cscale(x[H], scale);               // never visited by the reference (Q3)
      continue;
    }
    const int pm = (t == 0) ? E - m : E - 1 - m;
    float2 a = x[m], bb;
    if constexpr (SHFL) bb = ph[m]; else bb = sm[(pm - H) * T + pt];
    rfft_pair_folded<false>(a, bb, DERIVE ? split_tw<false, E>(hw0, m) : __ldg(&hw[t + m * T]), hs);
    __stcs(dst + t + m * T, a);
    __stcs(dst + pt + pm * T, bb);
  }
}
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS, BatchGeom<LOGN>::MIN_BLOCKS)
    rfft_fwd_reg_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ hw,
                        int batch, float scale, int ahead) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / T, t = threadIdx.x % T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  prefetch_successor<LOGN>(in, ahead, batch);
  rfft_fwd_reg_body<LOGN, false>(in + (active ? b : 0) * N, out + (active ? b : 0) * N, active, true,
                                 smem + lt * BatchGeom<LOGN>::ROW, tw, hw, t, scale);
}

// inverse: in [batch][N] float2, out [batch][2N] float (may alias). hw: folded inverse table.
// OLA (Clpconv's last step, cl_conv_kernels.h:120-124): the transform's 2N reals are not stored; the last pass leaves
// them in shared memory and the CTA writes out = (y[0, N) + tail) * ola_scale, tail = y[N, 2N).
template <int LOGN, bool OLA>
__device__ __forceinline__ void rfft_inv_reg_body(const float2 *src, float2 *dst, bool active, float2 *sm,
                                                  const float2 *__restrict__ tw, const float2 *__restrict__ hw, int t,
                                                  float2 *tail2 = nullptr, float ola_scale = 1.f) {
  using B = BatchGeom<LOGN>;
  using RS = RegSplitGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T, R0 = RS::R0, E = FftGeom<LOGN>::E, H = E / 2;
  // a thread reads its 8 low members X[t + m*T] and, straight from global memory, their partners X[N - (t + m*T)]
  // (the values its partner thread will own): 16 loads as before, and no exchange before the unsplit
  float2 x[E], hi[H];  // hi: unsplit high members, owned by the partner thread
  const int pt = (t == 0) ? 0 : T - t;
  const float2 zero2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < H; m++) {
    const int i = t + m * T;
    x[m] = active ? __ldcs(src + i) : zero2;
    hi[m] = active ? __ldcs(src + (i == 0 ? N / 2 : N - i)) : zero2;  // element N/2 rides with (t, m) = (0, 0)
  }
#pragma unroll
  for (int m = 0; m < H; m++) {
    if (m == 0 && t == 0) {
      x[0] = make_float2(x[0].x + x[0].y, x[0].x - x[0].y);  // packed (DC, Nyquist); hi[0] = element N/2 passes through
      continue;
    }
    rfft_pair_folded<true>(x[m], hi[m], __ldg(&hw[t + m * T]), 0.5f);  // (derived twiddles measured slower here)
  }
  // the one exchange: hand the high members to their owners
  if constexpr (B2F_SPLIT_SHFL && T <= 32) {
    // the owner is in the same warp: slot m' of thread t is hi[E - 1 - m'] of its partner (t = 0 keeps its own: slot
    // E/2 is element N/2 = hi[0], slot m' is hi[E - m'])
    const int plane = ((int)threadIdx.x & 31 & ~(T - 1)) | pt;
#pragma unroll
    for (int m = H; m < E; m++) {
      x[m].x = __shfl_sync(0xffffffffu, hi[E - 1 - m].x, plane);
      x[m].y = __shfl_sync(0xffffffffu, hi[E - 1 - m].y, plane);
      if (t == 0) x[m] = hi[(E - m) & (H - 1)];  // m = H -> hi[0]; m > H -> hi[E - m]
    }
  } else {
#pragma unroll
    for (int m = 0; m < H; m++) {
      const int pm = (t == 0) ? ((E - m) & (E - 1)) : E - 1 - m;  // (t = 0, m = 0) parks element N/2 in slot E/2
      const int slot = (t == 0 && m == 0) ? H : pm;
      sm[(slot - H) * T + pt] = hi[m];
    }
    __syncthreads();
#pragma unroll
    for (int m = H; m < E; m++) x[m] = sm[(m - H) * T + t];
    __syncthreads();  // staging is the engine's work buffer from here on
  }
  // first pass: value index m of (idx, slot): idx = (t + q*T) + r*(N/R0), slot = q*R0 + r  ->  m = q + r*(E/R0)
  auto load = [&](int, int slot) { return x[(slot / R0) + (slot % R0) * (E / R0)]; };
  if constexpr (!OLA) {
    auto store = [&](int idx, float2 v, int) {
      if (active) __stcs(dst + idx, v);
    };
    fft_run<LOGN, true>(load, store, sm, tw, t, CtaSync());
  } else {
    static_assert(B::TPB == 1, "the overlap-add epilogue assumes one transform per CTA");
    auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = v; };
    fft_run<LOGN, true, true>(load, store, sm, tw, t, CtaSync());
    __syncthreads();
    // element m holds the reals (y[2m], y[2m+1]); the thread that reads a tail element is the one that replaces it
    for (int m = t; m < N / 2; m += T) {
      const float2 y = sm[pad_idx(m)], z = sm[pad_idx(m + N / 2)], tl = tail2[m];
      dst[m] = make_float2((y.x + tl.x) * ola_scale, (y.y + tl.y) * ola_scale);
      tail2[m] = z;
    }
  }
}
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS, BatchGeom<LOGN>::MIN_BLOCKS)
    rfft_inv_reg_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ hw,
                        int batch, int ahead) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / T, t = threadIdx.x % T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  prefetch_successor<LOGN>(in, ahead, batch);
  rfft_inv_reg_body<LOGN, false>(in + (active ? b : 0) * N, out + (active ? b : 0) * N, active,
                                 smem + lt * BatchGeom<LOGN>::ROW, tw, hw, t);
}

// =====================================================================================================
// One thread, one transform: N <= 32. The transform AND the real split / unsplit (cl_fft.cpp:178-205; the pair
// partner N - i sits in the same thread) run in registers; nothing but the coalescing copy touches shared memory.
// N = 2 / 4: a transform is one / two 16-byte words, read and written straight from global memory, four transforms
// per thread so that a CTA moves 16-32 KiB. N = 8, 16, 32: the CTA's transforms are one contiguous block, copied in
// and out with 128-bit accesses into rows of N/2 + 1 float4 (an odd stride: conflict-free for the copy and for the
// one-thread-per-row 128-bit accesses alike), and a thread holds its row in up to 64 registers.
// These replaced multi-thread schedules through the generic engine ({8, 4} at N = 32: four threads and a shared-
// memory exchange per transform) and 4 KiB CTAs at N = 2. Measured, fraction of the HBM copy peak, c2c / r2c / c2r:
//   N = 2:  0.52 / 0.42 / 0.43 -> 1.02 / 1.02 / 1.02      N = 4:  0.86 / 0.71 / 0.72 -> 1.01 / 1.02 / 1.01
//   N = 8:  1.00 / 0.85 / 0.86 -> 0.99 / 0.99 / 0.99      N = 16: 1.04 / 0.97 / 0.97 -> 1.04 / 1.03 / 1.03
//   N = 32: 0.75 / 0.57 / 0.58 -> 1.01 / 1.01 / 0.99
// (N = 8 wants many small CTAs: 128 threads x 8 per SM 0.99, 256 x 4 0.95, 512 x 2 0.87, unstaged 0.94.)
// =====================================================================================================
template <int LOGN>
struct ThreadGeom {
  static constexpr int N = 1 << LOGN, V = N / 2;  // V: float4 words per transform
  static constexpr bool OK = LOGN <= 5;
  static constexpr bool STAGED = LOGN >= 3;  // through shared memory (N >= 8)
  static constexpr int THREADS = (LOGN == 5 || LOGN == 3) ? 128 : 256;
  static constexpr int MIN_BLOCKS = LOGN == 3 ? 8 : 4;
  static constexpr int ITER = 4;       // transforms per thread (direct kernels)
  static constexpr int ROW4 = V + 1;   // row stride, float4 units
  static constexpr int SMEM_BYTES = STAGED ? THREADS * ROW4 * 16 : 0;
  static constexpr int PER_CTA = STAGED ? THREADS : THREADS * ITER;  // transforms per CTA
};
enum { kThreadComplex = 0, kThreadRealFwd = 1, kThreadRealInv = 2 };

// v: the N points of one transform. hw: folded split table of the plan (forward: scale folded in), see below.
template <int LOGN, int KIND, bool INV>
__device__ __forceinline__ void thread_transform(float2 (&v)[1 << LOGN], const float2 *__restrict__ hw, float scale) {
  constexpr int N = 1 << LOGN;
  if constexpr (KIND == kThreadRealInv) {  // unsplit first (cl_fft.cpp:192-205); element N/2 passes through (Q3)
    v[0] = rfft_dc<true>(v[0]);
#pragma unroll
    for (int i = 1; i < N / 2; i++) rfft_pair_folded<true>(v[i], v[N - i], __ldg(&hw[i]), 0.5f);
  }
  dftR<N, INV>(v);
  if constexpr (KIND == kThreadComplex) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = cscale(v[i], scale);
  }
  if constexpr (KIND == kThreadRealFwd) {  // split (cl_fft.cpp:178-191): packed (DC, Nyquist), bin N/2 only scaled
    const float hs = 0.5f * scale;
    v[0] = make_float2((v[0].x + v[0].y) * hs, (v[0].x - v[0].y) * hs);
    v[N / 2] = cscale(v[N / 2], scale);
#pragma unroll
    for (int i = 1; i < N / 2; i++) rfft_pair_folded<false>(v[i], v[N - i], __ldg(&hw[i]), hs);
  }
}

// in/out: [batch][N] float2 (real transforms: the packed layouts of rfft_fwd_kernel / rfft_inv_kernel), may alias:
// a CTA has read everything it owns before it writes.
template <int LOGN, int KIND, bool INV>
__global__ void __launch_bounds__(ThreadGeom<LOGN>::THREADS, ThreadGeom<LOGN>::MIN_BLOCKS)
    fft_thread_kernel(const float4 *in, float4 *out, const float2 *__restrict__ hw, long long batch, float scale) {
  using G = ThreadGeom<LOGN>;
  constexpr int N = G::N, V = G::V;
  const int tid = threadIdx.x;
  auto unpack = [](const float4 (&w)[V], float2 (&v)[N]) {
#pragma unroll
    for (int q = 0; q < V; q++) {
      v[2 * q] = make_float2(w[q].x, w[q].y);
      v[2 * q + 1] = make_float2(w[q].z, w[q].w);
    }
  };
  auto pack = [](const float2 (&v)[N], float4 (&w)[V]) {
#pragma unroll
    for (int q = 0; q < V; q++) w[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
  };
  if constexpr (!G::STAGED) {
    const long long b0 = (long long)blockIdx.x * G::PER_CTA + tid;
    float4 w[G::ITER][V];
#pragma unroll
    for (int it = 0; it < G::ITER; it++) {
      const long long b = b0 + (long long)it * G::THREADS;
#pragma unroll
      for (int q = 0; q < V; q++) w[it][q] = b < batch ? __ldcs(in + b * V + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < G::ITER; it++) {
      float2 v[N];
      unpack(w[it], v);
      thread_transform<LOGN, KIND, INV>(v, hw, scale);
      pack(v, w[it]);
    }
    __syncthreads();  // in place: the CTA's loads are all done before its first store
#pragma unroll
    for (int it = 0; it < G::ITER; it++) {
      const long long b = b0 + (long long)it * G::THREADS;
      if (b < batch) {
#pragma unroll
        for (int q = 0; q < V; q++) __stcs(out + b * V + q, w[it][q]);
      }
    }
  } else {
    extern __shared__ float4 rows4[];
    const long long base = (long long)blockIdx.x * G::PER_CTA;
    const long long left = batch - base;
    const int ntr = left < G::PER_CTA ? (int)left : G::PER_CTA;
    const float4 *src = in + base * V;
    float4 *dst = out + base * V;
    float4 w[V];
    if (ntr == G::PER_CTA) {  // (every CTA but the last) V independent 128-bit loads in flight per thread
#pragma unroll
      for (int q = 0; q < V; q++) w[q] = __ldcs(src + tid + q * G::THREADS);
#pragma unroll
      for (int q = 0; q < V; q++) {
        const int i = tid + q * G::THREADS;
        rows4[(i / V) * G::ROW4 + (i % V)] = w[q];
      }
    } else {
      for (int i = tid; i < ntr * V; i += G::THREADS) rows4[(i / V) * G::ROW4 + (i % V)] = __ldcs(src + i);
    }
    __syncthreads();
    if (tid < ntr) {
      float4 *row = rows4 + tid * G::ROW4;
#pragma unroll
      for (int q = 0; q < V; q++) w[q] = row[q];
      float2 v[N];
      unpack(w, v);
      thread_transform<LOGN, KIND, INV>(v, hw, scale);
      pack(v, w);
#pragma unroll
      for (int q = 0; q < V; q++) row[q] = w[q];
    }
    __syncthreads();
    if (ntr == G::PER_CTA) {
#pragma unroll
      for (int q = 0; q < V; q++) {
        const int i = tid + q * G::THREADS;
        __stcs(dst + i, rows4[(i / V) * G::ROW4 + (i % V)]);
      }
    } else {
      for (int i = tid; i < ntr * V; i += G::THREADS) __stcs(dst + i, rows4[(i / V) * G::ROW4 + (i % V)]);
    }
  }
}

}  // namespace b2f
