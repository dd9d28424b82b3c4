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

// REALK: geometry of the generic real-transform kernels (same as the complex ones; kept apart for re-measurement)
template <int LOGN, bool REALK = false>
struct BatchGeom {
  using G = FftGeom<LOGN>;
  static constexpr int T = G::T;
  static constexpr int TPB = (T >= kTargetThreads) ? 1 : (kTargetThreads / T);  // transforms per CTA
  static constexpr int THREADS = T * TPB;
  // Transforms of N <= 32 points are carried by one or two threads, whose own accesses to global memory would be
  // 128-byte strides across a warp (measured: 28 % / 48 % of the HBM peak at N = 16 / 32). The CTA's TPB
  // transforms are one contiguous block of memory instead: it is copied in and out cooperatively (coalesced),
  // straight into / out of the engine's per-transform work rows, which the transform then uses in place.
  static constexpr bool STAGED = LOGN <= 5;  // (N = 64, generic real kernels: staged 56 %, direct 71 % of the HBM peak)
  static constexpr int ROW = STAGED ? (G::SMEM | 1) : G::SMEM;  // odd row stride: one-thread-per-row accesses conflict-free
  static constexpr int SMEM_BYTES = TPB * ROW * (int)sizeof(float2);
  static constexpr int MIN_BLOCKS = 1024 / THREADS;  // caps registers at 64/thread: 32 resident warps per SM
};

// cooperative, coalesced copy of the CTA's block of transforms between global memory and the work rows
template <int LOGN, bool TO_SMEM, bool REALK = false>
__device__ __forceinline__ void stage_copy(const float2 *gin, float2 *gout, float2 *rows, long long total) {
  using B = BatchGeom<LOGN, REALK>;
  constexpr int N = 1 << LOGN;
  const long long base = (long long)blockIdx.x * B::TPB * N;
  const int n = (int)(total - base < (long long)B::TPB * N ? total - base : (long long)B::TPB * N);
  for (int i = threadIdx.x; i < n; i += B::THREADS) {
    float2 *cell = rows + (i >> LOGN) * B::ROW + pad_idx(i & (N - 1));
    if (TO_SMEM)
      *cell = gin[base + i];
    else
      gout[base + i] = *cell;
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
  if constexpr (B::STAGED) {
    stage_copy<LOGN, true>(in, nullptr, smem, (long long)batch * N);
    __syncthreads();
    auto load = [&](int idx, int) { return sm[pad_idx(idx)]; };
    auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = cscale(v, scale); };
    fft_run<LOGN, INV, true, true>(load, store, sm, tw, t, CtaSync());
    __syncthreads();
    stage_copy<LOGN, false>(nullptr, out, smem, (long long)batch * N);
    return;
  }
  auto load = [&](int idx, int) { return active ? src[idx] : make_float2(0.f, 0.f); };
  auto store = [&](int idx, float2 v, int) {
    if (active) dst[idx] = cscale(v, scale);
  };
  fft_run<LOGN, INV>(load, store, sm, tw, t, CtaSync());
}

// ---- real to complex (forward) -------------------------------------------------------------------
// in: [batch][2N] float (read as N packed float2), out: [batch][N] float2, may alias.
// Output convention of the reference: element 0 = (DC, Nyquist)/size packed, element k =
// 2 X[k]/size, element N/2 left as the plain FFT value (the reference's split never visits it, Q3).
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN, true>::THREADS, BatchGeom<LOGN, true>::MIN_BLOCKS)
    rfft_fwd_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ w2,
                    int batch, float scale) {
  using B = BatchGeom<LOGN, true>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * B::ROW;
  if constexpr (B::STAGED) {  // coalesced copy-in; the transform then works in its row, the split rewrites it in place
    stage_copy<LOGN, true, true>(in, nullptr, smem, (long long)batch * N);
    __syncthreads();
  }
  auto load = [&](int idx, int) {
    if constexpr (B::STAGED) return sm[pad_idx(idx)];
    return active ? src[idx] : make_float2(0.f, 0.f);
  };
  auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = cscale(v, scale); };
  fft_run<LOGN, false, true, B::STAGED>(load, store, sm, tw, t, CtaSync());
  __syncthreads();
  auto put = [&](int i, float2 v) {
    if constexpr (B::STAGED)
      sm[pad_idx(i)] = v;
    else
      dst[i] = v;
  };
  if (active) {
    // split: pairs (i, N-i), i in [1, N/2); elements 0 and N/2 handled apart
    for (int i = t; i <= N / 2; i += B::T) {
      if (i == 0) {
        put(0, rfft_dc<false>(sm[pad_idx(0)]));
      } else if (i == N / 2) {
        put(i, sm[pad_idx(i)]);
      } else {
        float2 ci = sm[pad_idx(i)], cj = sm[pad_idx(N - i)];
        rfft_pair<false>(ci, cj, __ldg(&w2[i]));
        put(i, ci);
        put(N - i, cj);
      }
    }
  }
  if constexpr (B::STAGED) {
    __syncthreads();
    stage_copy<LOGN, false, true>(nullptr, out, smem, (long long)batch * N);
  }
}

// ---- complex to real (inverse) -------------------------------------------------------------------
// in: [batch][N] float2 in the layout rfft_fwd_kernel writes, out: [batch][2N] float, may alias.
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN, true>::THREADS, BatchGeom<LOGN, true>::MIN_BLOCKS)
    rfft_inv_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ w2,
                    int batch) {
  using B = BatchGeom<LOGN, true>;
  constexpr int N = 1 << LOGN;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / B::T, t = threadIdx.x % B::T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * B::ROW;
  if constexpr (B::STAGED) {  // coalesced copy-in; unsplit and transform in place in the row, coalesced copy-out
    stage_copy<LOGN, true, true>(in, nullptr, smem, (long long)batch * N);
    __syncthreads();
  }
  auto get = [&](int i) {
    if constexpr (B::STAGED) return sm[pad_idx(i)];
    return src[i];
  };
  if (active) {
    for (int i = t; i <= N / 2; i += B::T) {
      if (i == 0) {
        sm[pad_idx(0)] = rfft_dc<true>(get(0));
      } else if (i == N / 2) {
        sm[pad_idx(i)] = get(i);
      } else {
        float2 ci = get(i), cj = get(N - i);
        rfft_pair<true>(ci, cj, __ldg(&w2[i]));
        sm[pad_idx(i)] = ci;
        sm[pad_idx(N - i)] = cj;
      }
    }
  }
  __syncthreads();
  auto load = [&](int idx, int) { return sm[pad_idx(idx)]; };
  auto store = [&](int idx, float2 v, int) {
    if constexpr (B::STAGED)
      sm[pad_idx(idx)] = v;
    else if (active)
      dst[idx] = v;
  };
  fft_run<LOGN, true, B::STAGED, true>(load, store, sm, tw, t, CtaSync());
  if constexpr (B::STAGED) {
    __syncthreads();
    stage_copy<LOGN, false, true>(nullptr, out, smem, (long long)batch * N);
  }
}


// =====================================================================================================
// Register-level real transforms for N >= 64 (schedules whose first and last passes leave every thread
// with the E = 16 (8 at N = 64) values X[t + m*T], m = 0..E-1, T = N/E; written below for E = 16). The pair partner of element (t, m) is element
// (T - t, 15 - m) [(0, 16 - m) for t = 0], so the split / unsplit needs ONE partner thread: the two swap
// half of their values through a small shared-memory staging area, each evaluates its 8 pairs once (the inverse
// reads the partner's inputs from global memory instead and only hands the results over).
// Against the generic kernels above this drops a full shared-memory round trip and halves the split
// arithmetic: 0.5, the 1/N scaling and the quarter turn are folded into the table
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
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS, BatchGeom<LOGN>::MIN_BLOCKS)
    rfft_fwd_reg_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ hw,
                        int batch, float scale) {
  using B = BatchGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T, E = FftGeom<LOGN>::E, H = E / 2;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / T, t = threadIdx.x % T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * BatchGeom<LOGN>::ROW;
  float2 x[E];
  auto load = [&](int idx, int) { return active ? __ldcs(src + idx) : make_float2(0.f, 0.f); };
  auto store = [&](int, float2 v, int slot) { x[slot] = v; };  // last pass: slot == m, value X[t + m*T]
  fft_run<LOGN, false>(load, store, sm, tw, t, CtaSync());
  __syncthreads();  // every thread is past its last gather: sm becomes the staging area [E/2][T]
#pragma unroll
  for (int m = H; m < E; m++) sm[(m - H) * T + t] = x[m];
  __syncthreads();
  if (!active) return;
  const int pt = (t == 0) ? 0 : T - t;  // partner thread (itself for t = 0 and t = T/2)
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
    float2 a = x[m], bb = sm[(pm - H) * T + pt];
    rfft_pair_folded<false>(a, bb, DERIVE ? split_tw<false, E>(hw0, m) : __ldg(&hw[t + m * T]), hs);
    __stcs(dst + t + m * T, a);
    __stcs(dst + pt + pm * T, bb);
  }
}

// inverse: in [batch][N] float2, out [batch][2N] float (may alias). hw: folded inverse table.
template <int LOGN>
__global__ void __launch_bounds__(BatchGeom<LOGN>::THREADS, BatchGeom<LOGN>::MIN_BLOCKS)
    rfft_inv_reg_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw, const float2 *__restrict__ hw,
                        int batch) {
  using B = BatchGeom<LOGN>;
  using RS = RegSplitGeom<LOGN>;
  constexpr int N = 1 << LOGN, T = B::T, R0 = RS::R0, E = FftGeom<LOGN>::E, H = E / 2;
  extern __shared__ float2 smem[];
  const int lt = threadIdx.x / T, t = threadIdx.x % T;
  const long long b = (long long)blockIdx.x * B::TPB + lt;
  const bool active = b < batch;
  const float2 *src = in + (active ? b : 0) * N;
  float2 *dst = out + (active ? b : 0) * N;
  float2 *sm = smem + lt * BatchGeom<LOGN>::ROW;
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
  // first pass: value index m of (idx, slot): idx = (t + q*T) + r*(N/R0), slot = q*R0 + r  ->  m = q + r*(E/R0)
  auto load = [&](int, int slot) { return x[(slot / R0) + (slot % R0) * (E / R0)]; };
  auto store = [&](int idx, float2 v, int) {
    if (active) __stcs(dst + idx, v);
  };
  fft_run<LOGN, true>(load, store, sm, tw, t, CtaSync());
}

}  // namespace b2f
