// fft_large.cuh -- transforms too long for one CTA's shared memory (N = 2^15, 2^16 complex, i.e. the
// 65536- and 131072-point real FFTs of BASELINE config 5), as a four-step factorisation N = N1 * N2:
//
//   columns kernel: for every column n2, an N1-point FFT over n1 of x[N2*n1 + n2], times W_N^(n2*k1),
//                   written to a scratch matrix T[k1][n2]. A CTA owns C adjacent columns (128-byte
//                   runs in HBM) and keeps its C*N1 inter-step twiddles in REGISTERS across the loop
//                   over the batch, so they cost no memory traffic.
//   rows kernel:    for every row k1, an N2-point FFT over n2 of T[k1][:], transposed through shared
//                   memory and written to X[k1 + N1*k2] in 128-byte runs, with the 1/N scaling fused.
//
// Both are built from the same register/shared-memory Stockham engine (fft_core.cuh) as the small
// kernels. Reference being replaced: the same reorder + log2(N) stage launches of Clcfft::fft()
// (cl_fft.cpp:138-151), which at N = 32768 re-read and re-write the whole array 16 times.
// HBM traffic here: one read + one write per step; the scratch matrix is written by step 1 and read by
// step 2 back to back, so it is served from the 126 MB L2 when the batch chunk fits.
//
// The real-FFT split/unsplit (cl_fft.cpp:178-205) runs as a separate element-wise pass for now.
#pragma once

#include "fft_core.cuh"

// Build-time knobs kept for re-measurement. Measured on B200 (2048 x 32768 complex): twiddles in registers or
// streamed from L2 at 2 CTAs/SM: 0.365 ms either way; 3 or 4 CTAs/SM (streamed twiddles, no spills): 0.43 ms --
// more resident column CTAs mean more concurrent 2 KB-strided streams and the memory system likes that less.
#ifndef LARGE_COLS_TW_REGS
#define LARGE_COLS_TW_REGS 1
#endif
#ifndef LARGE_COLS_MINB
#define LARGE_COLS_MINB 1  // measured: forcing 3 or 4 CTAs/SM (80/64 registers, spills) is 20-25 % slower
#endif

namespace b2f {

template <int LOG1, int LOG2>
struct LargeGeom {
  static constexpr int N1 = 1 << LOG1, N2 = 1 << LOG2, N = N1 * N2;
  using G1 = FftGeom<LOG1>;
  using G2 = FftGeom<LOG2>;
  static constexpr int C = 256 / G1::T;   // columns per CTA in the columns kernel
  static constexpr int RB = 256 / G2::T;  // rows per CTA in the rows kernel (16 for N2 = 256)
  static constexpr int THREADS = 256;
  static constexpr int SMEM_A = C * G1::SMEM * (int)sizeof(float2);
  // rows kernel: per-row stride in float2. Odd (== 1 mod 16) for the complex write-out, where a half-warp
  // reads one element from each of 16 rows; == 2 (mod 16) for the real write-out, where a half-warp reads two
  // adjacent elements from each of 8 rows
  static constexpr int ROWSTRIDE_C = G2::SMEM, ROWSTRIDE_R = G2::SMEM + 1;
  static constexpr int SMEM_B = RB * (G2::SMEM + 1) * (int)sizeof(float2);
};

// grid = (N2 / C, batch slots). twl: [N1][N2] table, twl[k1*N2 + n2] = W_N^(n2*k1) (forward sign).
template <int LOG1, int LOG2, bool INV>
__global__ void __launch_bounds__(256, LARGE_COLS_MINB)
    large_cols_kernel(const float2 *in, float2 *scratch, const float2 *__restrict__ tw1,
                      const float2 *__restrict__ twl, int batch) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N2 = L::N2, N = L::N, C = L::C, E = L::G1::E;
  extern __shared__ float2 smem[];
  const int c = threadIdx.x % C, t = threadIdx.x / C;
  const int n2 = blockIdx.x * C + c;
  float2 *sm = smem + c * L::G1::SMEM;
#if LARGE_COLS_TW_REGS
  // this thread's inter-step twiddles, fixed for every transform of the batch
  float2 wreg[E];
#pragma unroll
  for (int s = 0; s < E; s++) {
    const int k1 = last_pass_index<LOG1>(t, s);
    float2 w = __ldg(&twl[(size_t)k1 * N2 + n2]);
    if (INV) w.y = -w.y;
    wreg[s] = w;
  }
#endif
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float2 *src = in + (size_t)b * N + n2;
    float2 *dst = scratch + (size_t)b * N + n2;
    auto load = [&](int idx, int) { return src[(size_t)idx * N2]; };
#if LARGE_COLS_TW_REGS
    auto store = [&](int idx, float2 v, int slot) { dst[(size_t)idx * N2] = cmul(v, wreg[slot]); };
#else
    auto store = [&](int idx, float2 v, int) {
      float2 w = __ldg(&twl[(size_t)idx * N2 + n2]);
      if (INV) w.y = -w.y;
      dst[(size_t)idx * N2] = cmul(v, w);
    };
#endif
    fft_run<LOG1, INV>(load, store, sm, tw1, t, CtaSync());
    __syncthreads();  // shared memory is reused by the next transform
  }
}

// grid = (N1 / RB, batch slots). Reads scratch rows, writes out[k1 + N1*k2] * scale.
// REAL (forward real transform): the split of cl_fft.cpp:178-191 is fused into the write-out. A CTA then
// owns 8 rows {8g..8g+7} and their mirrors {N1-k} (row 0 mirrors itself; its slot hosts row N1/2), so both
// members of every pair (i, N-i) = ((k1,k2), (N1-k1, N2-1-k2)) sit in its shared memory; every pair is
// evaluated once (folded table hw, scale included) and both members are stored, in 64-byte runs.
template <int LOG1, int LOG2, bool INV, bool REAL>
__global__ void __launch_bounds__(256)
    large_rows_kernel(const float2 *scratch, float2 *out, const float2 *__restrict__ tw2,
                      const float2 *__restrict__ hw, int batch, float scale) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N1 = L::N1, N2 = L::N2, N = L::N, RB = L::RB, T2 = L::G2::T;
  static_assert(!REAL || RB == 16, "mirrored row groups are 8 + 8");
  extern __shared__ float2 smem[];
  const int t = threadIdx.x % T2, row = threadIdx.x / T2;
  const int g = blockIdx.x;
  auto row_of = [&](int rr) -> int {
    if (!REAL) return g * RB + rr;
    if (rr < 8) return g * 8 + rr;
    const int d = g * 8 + (rr - 8);
    return d == 0 ? N1 / 2 : N1 - d;
  };
  constexpr int RS = REAL ? L::ROWSTRIDE_R : L::ROWSTRIDE_C;
  float2 *sm = smem + row * RS;
  const int k1_fft = row_of(row);
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float2 *src = scratch + (size_t)b * N + (size_t)k1_fft * N2;
    auto load = [&](int idx, int) { return src[idx]; };
    auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = v; };
    fft_run<LOG2, INV, true>(load, store, sm, tw2, t, CtaSync());
    __syncthreads();
    float2 *dst = out + (size_t)b * N;
    if (!REAL) {
      // transposed write-out: consecutive threads take consecutive rows (k1), i.e. consecutive addresses
      const int rr = threadIdx.x % RB;
      const float2 *smr = smem + rr * RS;
      for (int k2 = threadIdx.x / RB; k2 < N2; k2 += L::THREADS / RB) {
        float2 v = smr[pad_idx(k2)];
        dst[(size_t)k2 * N1 + g * RB + rr] = make_float2(v.x * scale, v.y * scale);
      }
    } else {
      // one pair per (direct row rr < 8, k2): low member i = k1 + N1*k2 when i < N/2, else its partner is
      const int rr = threadIdx.x % 8;
      const int k1 = g * 8 + rr;              // direct row, in [0, N1/2)
      const bool zero = (k1 == 0);
      const int prr = zero ? 0 : rr + 8;      // row 0 pairs with itself
      const int k1p = zero ? 0 : N1 - k1;
      const float2 *smr = smem + rr * RS, *smp = smem + prr * RS;
      const float hs = 0.5f * scale;
      // rows 1..N1/2-1: all N2 values of k2, the pair's other member is on the mirror row.
      // row 0: pairs (0,k2) <-> (0,N2-k2) for k2 in [1, N2/2), plus the two self-paired elements.
      for (int k2 = threadIdx.x / 8; k2 < N2; k2 += L::THREADS / 8) {
        const int pk2 = zero ? N2 - k2 : N2 - 1 - k2;
        if (zero && (k2 == 0 || k2 >= N2 / 2)) {
          if (k2 == 0) {
            const float2 v = smr[pad_idx(0)];
            dst[0] = make_float2((v.x + v.y) * hs, (v.x - v.y) * hs);
          } else if (k2 == N2 / 2) {
            const float2 v = smr[pad_idx(k2)];
            dst[N / 2] = make_float2(v.x * scale, v.y * scale);  // never visited by the reference (Q3)
          }
          continue;
        }
        float2 a = smr[pad_idx(k2)], bb = smp[pad_idx(pk2)];
        const int i = k1 + N1 * k2, j = k1p + N1 * pk2;  // i + j == N
        if (i < j) {
          rfft_pair_folded<false>(a, bb, __ldg(&hw[i]), hs);
        } else {
          rfft_pair_folded<false>(bb, a, __ldg(&hw[j]), hs);
        }
        dst[i] = a;
        dst[j] = bb;
      }
      // the self-mirrored row N1/2 lives in slot 8 of group 0: pairs (N1/2,k2) <-> (N1/2, N2-1-k2)
      if (g == 0) {
        const float2 *smh = smem + 8 * RS;
        for (int k2 = threadIdx.x; k2 < N2 / 2; k2 += L::THREADS) {
          const int pk2 = N2 - 1 - k2;
          float2 a = smh[pad_idx(k2)], bb = smh[pad_idx(pk2)];
          const int i = N1 / 2 + N1 * k2, j = N1 / 2 + N1 * pk2;
          rfft_pair_folded<false>(a, bb, __ldg(&hw[i]), hs);
          dst[i] = a;
          dst[j] = bb;
        }
      }
    }
    __syncthreads();
  }
}

// element-wise real-FFT split (forward, after the complex transform) / unsplit (inverse, before it)
// over [batch][N] complex values; pairs (i, N-i), element N/2 left alone. May run in place.
template <bool INV>
__global__ void __launch_bounds__(256)
    rfft_split_kernel(const float2 *in, float2 *out, const float2 *__restrict__ w2, int N, long long pairs_total) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= pairs_total) return;
  const int half = N / 2;
  const long long b = g / half;
  const int i = (int)(g % half);
  const float2 *src = in + b * N;
  float2 *dst = out + b * N;
  if (i == 0) {
    dst[0] = rfft_dc<INV>(src[0]);
    dst[half] = src[half];
  } else {
    float2 ci = src[i], cj = src[N - i];
    rfft_pair<INV>(ci, cj, __ldg(&w2[i]));
    dst[i] = ci;
    dst[N - i] = cj;
  }
}

}  // namespace b2f
