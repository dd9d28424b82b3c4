// fft_large.cuh -- transforms too long for one CTA's shared memory as a four-step factorisation N = N1 * N2, two
// launches with a scratch matrix in between. Used for N = 2^16 complex (131072-point real) at any batch and for
// N = 2^15 (the 65536-point real FFT of BASELINE config 5) at SMALL batches, where spreading one transform over the
// whole GPU beats the one-SM kernel of fft_sm.cuh (which takes over from ~100 transforms per call):
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
// The real-FFT split (cl_fft.cpp:178-191) is fused into the rows kernel, the unsplit (192-205) into a columns kernel
// with mirrored column ownership; rfft_split_kernel is the unfused pass kept behind option separate_split.
// Two single-launch variants built in round 1 (4-CTA clusters with a DSMEM transposition; 8-CTA clusters with an
// L2-resident scratch) lost to this pair and were removed; profiles/r01_fft_cluster_notes.md has their numbers.
#pragma once

#include "fft_core.cuh"
#include "tma_utils.cuh"

// Build-time knobs kept for re-measurement. Measured on B200 (2048 x 32768 complex): twiddles in registers or
// streamed from L2 at 2 CTAs/SM: 0.365 ms either way; 3 or 4 CTAs/SM (streamed twiddles, no spills): 0.43 ms --
// more resident column CTAs mean more concurrent 2 KB-strided streams and the memory system likes that less.
#ifndef LARGE_COLS_TW_REGS
#define LARGE_COLS_TW_REGS 1
#endif
#ifndef LARGE_COLS_MINB
#define LARGE_COLS_MINB 1  // measured: forcing 3 or 4 CTAs/SM (80/64 registers, spills) is 20-25 % slower
#endif

// Resident warps per SM the rows kernel's register budget is set for. Measured with the register prefetch of the
// next transform (1024 x 65536 real / 32768 complex, ms): 32 warps (64 registers, spills) 0.263 / 0.186,
// 24 warps 0.203 / 0.190, 16 warps (128 registers) 0.204 / 0.181; without the prefetch 0.219 / 0.187.
#ifndef LARGE_ROWS_WARPS
#define LARGE_ROWS_WARPS 16
#endif
#ifndef LARGE_ROWS_PAIR_UNROLL
#define LARGE_ROWS_PAIR_UNROLL 8  // unroll of the real write-out's pair loop (8 = all pairs of a thread)
#endif
#define B2F_STR_(x) #x
#define B2F_UNROLL(n) _Pragma(B2F_STR_(unroll n))

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

// The inter-step twiddles of one thread: W_N^(n2*k1) for its column n2 and the E rows k1 its last pass stores.
template <int LOG1, int LOG2, bool INV>
__device__ __forceinline__ void large_cols_twiddles(float2 (&wreg)[FftGeom<LOG1>::E], const float2 *__restrict__ twl,
                                                    int n2, int t) {
  constexpr int N2 = 1 << LOG2;
#pragma unroll
  for (int s = 0; s < FftGeom<LOG1>::E; s++) {
    const int k1 = last_pass_index<LOG1>(t, s);
    float2 w = __ldg(&twl[(size_t)k1 * N2 + n2]);
    if (INV) w.y = -w.y;
    wreg[s] = w;
  }
}
// One column of one transform: src/dst already point at element n2 of the transform / scratch matrix.
template <int LOG1, int LOG2, bool INV>
__device__ __forceinline__ void large_cols_body(const float2 *src, float2 *dst, float2 *sm,
                                                const float2 *__restrict__ tw1,
                                                const float2 (&wreg)[FftGeom<LOG1>::E], int t) {
  constexpr int N2 = 1 << LOG2;
  auto load = [&](int idx, int) { return src[(size_t)idx * N2]; };
  auto store = [&](int idx, float2 v, int slot) { dst[(size_t)idx * N2] = cmul(v, wreg[slot]); };
  fft_run<LOG1, INV>(load, store, sm, tw1, t, CtaSync());
}

// grid = (N2 / C, batch slots). twl: [N1][N2] table, twl[k1*N2 + n2] = W_N^(n2*k1) (forward sign).
template <int LOG1, int LOG2, bool INV>
__global__ void __launch_bounds__(256, LARGE_COLS_MINB)
    large_cols_kernel(const float2 *in, float2 *scratch, const float2 *__restrict__ tw1,
                      const float2 *__restrict__ twl, int batch) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N = L::N, C = L::C, E = L::G1::E;
  extern __shared__ float2 smem[];
  const int c = threadIdx.x % C, t = threadIdx.x / C;
  const int n2 = blockIdx.x * C + c;
  float2 *sm = smem + c * L::G1::SMEM;
#if LARGE_COLS_TW_REGS
  // this thread's inter-step twiddles, fixed for every transform of the batch
  float2 wreg[E];
  large_cols_twiddles<LOG1, LOG2, INV>(wreg, twl, n2, t);
#endif
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float2 *src = in + (size_t)b * N + n2;
    float2 *dst = scratch + (size_t)b * N + n2;
#if LARGE_COLS_TW_REGS
    large_cols_body<LOG1, LOG2, INV>(src, dst, sm, tw1, wreg, t);
#else
    auto load = [&](int idx, int) { return src[(size_t)idx * N2]; };
    auto store = [&](int idx, float2 v, int) {
      float2 w = __ldg(&twl[(size_t)idx * N2 + n2]);
      if (INV) w.y = -w.y;
      dst[(size_t)idx * N2] = cmul(v, w);
    };
    fft_run<LOG1, INV>(load, store, sm, tw1, t, CtaSync());
#endif
    __syncthreads();  // shared memory is reused by the next transform
  }
}

// Inverse real transform: the unsplit of cl_fft.cpp:192-205 fused into the columns kernel. The partner of element
// (n1, n2) is (N1-1-n1, N2-n2) [(N1-n1, 0) in column 0], so a CTA owns C/2 columns {g*C/2 ...} and their mirrors
// {N2 - n2} (column 0 mirrors itself; its slot hosts column N2/2, which also pairs with itself): the whole tile is
// staged in shared memory, every pair is unsplit there once, then the column transforms run in place.
// grid = (N2 / C, batch slots). hw: folded inverse table (conj(0.5 i w2)), N/2 + 1 entries.
template <int LOG1, int LOG2>
__global__ void __launch_bounds__(256, 2)
    large_cols_unsplit_kernel(const float2 *in, float2 *scratch, const float2 *__restrict__ tw1,
                              const float2 *__restrict__ twl, const float2 *__restrict__ hw, int batch) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N1 = L::N1, N2 = L::N2, N = L::N, C = L::C, CH = C / 2, E = L::G1::E, T = L::G1::T;
  constexpr int RSTEP = 256 / CH, NPAIR = N1 / RSTEP;  // rows per round of the pair loop, rounds
  extern __shared__ float2 smem[];
  const int c = threadIdx.x % C, t = threadIdx.x / C, g = blockIdx.x;
  auto col_of = [&](int slot) -> int {
    if (slot < CH) return g * CH + slot;
    const int d = g * CH + (slot - CH);
    return d == 0 ? N2 / 2 : N2 - d;
  };
  const int n2 = col_of(c);
  float2 *smc = smem + c * L::G1::SMEM;
  float2 wreg[E];
  large_cols_twiddles<LOG1, LOG2, true>(wreg, twl, n2, t);
  // pair duty of this thread: direct column cd, rows n1 = r0 + RSTEP*m
  const int cd = threadIdx.x % CH, r0 = threadIdx.x / CH;
  const int n2d = g * CH + cd;
  const bool col0 = (n2d == 0);
  float2 *smd = smem + cd * L::G1::SMEM, *smm = smem + (cd + CH) * L::G1::SMEM;
  // split twiddle of the thread's first pair; pair m is element i_0 + N2*RSTEP*m and
  // w2[i_m] = w2[i_0] * exp(+i pi m RSTEP/N1) = w2[i_0] * exp(+i pi m/8): one table load for the whole launch
  static_assert(N1 / RSTEP == 8, "split-twiddle step constants are exp(i pi m/8)");
  const float2 hw0 = __ldg(&hw[N2 * r0 + n2d]);  // i_0 < N/8
  constexpr float kC[8] = {1.f, B2F_COS_PI_8, B2F_SQRT1_2, B2F_SIN_PI_8, 0.f, -B2F_SIN_PI_8, -B2F_SQRT1_2, -B2F_COS_PI_8};
  constexpr float kS[8] = {0.f, B2F_SIN_PI_8, B2F_SQRT1_2, B2F_COS_PI_8, 1.f, B2F_COS_PI_8, B2F_SQRT1_2, B2F_SIN_PI_8};
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float2 *src = in + (size_t)b * N;
#pragma unroll
    for (int s = 0; s < E; s++) smc[pad_idx(t + s * T)] = src[(size_t)(t + s * T) * N2 + n2];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NPAIR; m++) {
      const int n1 = r0 + RSTEP * m;
      const float2 h = m ? cmulc<true>(hw0, kC[m], kS[m]) : hw0;  // entry i_m of the table, extended past N/2
      if (!col0) {
        float2 a = smd[pad_idx(n1)], bb = smm[pad_idx(N1 - 1 - n1)];
        const int i = N2 * n1 + n2d;
        if (i < N - i) {
          rfft_pair_folded<true>(a, bb, h, 0.5f);
        } else {
          rfft_pair_folded<true>(bb, a, cconj(h), 0.5f);  // hw[N - i] == conj(h)
        }
        smd[pad_idx(n1)] = a;
        smm[pad_idx(N1 - 1 - n1)] = bb;
      } else if (n1 < N1 / 2) {
        // column 0: (n1, 0) <-> (N1 - n1, 0); element 0 is the packed (DC, Nyquist), element N/2 passes through (Q3)
        if (n1 == 0) {
          smd[pad_idx(0)] = rfft_dc<true>(smd[pad_idx(0)]);
        } else {
          float2 a = smd[pad_idx(n1)], bb = smd[pad_idx(N1 - n1)];
          rfft_pair_folded<true>(a, bb, h, 0.5f);
          smd[pad_idx(n1)] = a;
          smd[pad_idx(N1 - n1)] = bb;
        }
        // column N2/2 (hosted in column 0's mirror slot): (n1, N2/2) <-> (N1 - 1 - n1, N2/2)
        float2 a = smm[pad_idx(n1)], bb = smm[pad_idx(N1 - 1 - n1)];
        rfft_pair_folded<true>(a, bb, __ldg(&hw[N2 * n1 + N2 / 2]), 0.5f);
        smm[pad_idx(n1)] = a;
        smm[pad_idx(N1 - 1 - n1)] = bb;
      }
    }
    __syncthreads();
    float2 *dst = scratch + (size_t)b * N + n2;
    auto load = [&](int idx, int) { return smc[pad_idx(idx)]; };
    auto store = [&](int idx, float2 v, int slot) { dst[(size_t)idx * N2] = cmul(v, wreg[slot]); };
    fft_run<LOG1, true, false, true>(load, store, smc, tw1, t, CtaSync());
    __syncthreads();  // shared memory is reused by the next transform
  }
}

// grid = (N1 / RB, batch slots). Reads scratch rows, writes out[k1 + N1*k2] * scale.
// REAL (forward real transform): the split of cl_fft.cpp:178-191 is fused into the write-out. A CTA then
// owns 8 rows {8g..8g+7} and their mirrors {N1-k} (row 0 mirrors itself; its slot hosts row N1/2), so both
// members of every pair (i, N-i) = ((k1,k2), (N1-k1, N2-1-k2)) sit in its shared memory; every pair is
// evaluated once (folded table hw, scale included) and both members are stored, in 64-byte runs.
// One row group g of one transform: scratch_b / out_b point at the transform's scratch matrix / output.
// CG: read the scratch with ld.global.cg.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// after_reads(): called by every thread once the whole CTA holds its rows in shared memory, i.e. when the
// scratch has been consumed and before anything is written to `out`.
// RBT: rows per CTA (RBT * T2 threads). REAL with RBT = 32 (16 direct rows + 16 mirrors) makes the direct members'
// runs whole 128-byte lines (the mirrors' runs are 120 + 8 bytes); RBT = 16 gives 64-byte runs.
// scratch row transformed by slot rr of row group g (REAL: RBT/2 direct rows, then their mirrors; row 0 mirrors
// itself, its slot hosts row N1/2)
template <int LOG1, bool REAL, int RBT>
__device__ __forceinline__ int large_rows_k1(int g, int rr) {
  constexpr int N1 = 1 << LOG1, RH = RBT / 2;
  if (!REAL) return g * RBT + rr;
  if (rr < RH) return g * RH + rr;
  const int d = g * RH + (rr - RH);
  return d == 0 ? N1 / 2 : N1 - d;
}
// PRE: the thread's first-pass inputs are already in `xin` (fetched during the previous transform's write-out)
template <int LOG1, int LOG2, bool INV, bool REAL, bool CG, class Hook = NoHook, bool TWS = false,
          int RBT = LargeGeom<LOG1, LOG2>::RB, bool PRE = false>
__device__ __forceinline__ void large_rows_body(const float2 *scratch_b, float2 *out_b, float2 *smem,
                                                const float2 *__restrict__ tw2, const float2 *__restrict__ hw,
                                                float scale, int g, Hook after_reads = Hook(),
                                                const float2 *xin = nullptr) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N1 = L::N1, N2 = L::N2, N = L::N, RB = RBT, T2 = L::G2::T, NTHR = RBT * T2, RH = RBT / 2;
  static_assert(!REAL || RB == 16 || RB == 32, "mirrored row groups are 8 + 8 or 16 + 16");
  const int t = threadIdx.x % T2, row = threadIdx.x / T2;
  auto row_of = [&](int rr) -> int { return large_rows_k1<LOG1, REAL, RBT>(g, rr); };
  // per-row stride in float2: odd when a half-warp reads one element from each of 16 rows, == 2 (mod 16) when it
  // reads two adjacent elements from each of 8 rows
  constexpr int RS = (REAL && RH == 8) ? L::ROWSTRIDE_R : L::ROWSTRIDE_C;
  float2 *sm = smem + row * RS;
  const int k1_fft = row_of(row);
  const float2 *src = scratch_b + (size_t)k1_fft * N2;
  auto load = [&](int idx, int slot) { return PRE ? xin[slot] : (CG ? __ldcg(&src[idx]) : src[idx]); };
  auto store = [&](int idx, float2 v, int) { sm[pad_idx(idx)] = v; };
  // REAL: the folded split twiddle of this thread's first pair, requested before the transform so that its latency
  // is off the write-out's critical path. Pair m of the thread is element i_m = i_0 + N1*KSTEP*m, and
  // w2[i_m] = w2[i_0] * exp(-i pi m KSTEP/N2), a compile-time constant per m: one table load per thread instead of
  // N2/KSTEP, the products go to the idle FP32 pipe. For i_m > N/2 the pair is evaluated from its other member,
  // whose table entry hw[N - i_m] is the conjugate of the same product.
  constexpr int KSTEP = NTHR / RH;  // k2 values covered per round of the pair loop
  constexpr int NPAIR = REAL ? N2 / KSTEP : 1;
  static_assert(!REAL || NPAIR == 8, "split-twiddle step constants are exp(-i pi m/8)");
  float2 hw0 = make_float2(0.f, 0.f);
  if constexpr (REAL) hw0 = __ldg(&hw[g * RH + threadIdx.x % RH + N1 * (threadIdx.x / RH)]);  // i_0 < N1*KSTEP <= N/2
  fft_run<LOG2, INV, true, false, TWS>(load, store, sm, tw2, t, CtaSync());
  __syncthreads();
  after_reads();
  float2 *dst = out_b;
  if (!REAL) {
    // transposed write-out: consecutive threads take consecutive rows (k1), i.e. consecutive addresses
    const int rr = threadIdx.x % RB;
    const float2 *smr = smem + rr * RS;
    for (int k2 = threadIdx.x / RB; k2 < N2; k2 += NTHR / RB) {
      float2 v = smr[pad_idx(k2)];
      dst[(size_t)k2 * N1 + g * RB + rr] = cscale(v, scale);
    }
  } else {
    // one pair per (direct row rr < RH, k2): low member i = k1 + N1*k2 when i < N/2, else its partner is
    const int rr = threadIdx.x % RH;
    const int k1 = g * RH + rr;             // direct row, in [0, N1/2)
    const bool zero = (k1 == 0);
    const int prr = zero ? 0 : rr + RH;     // row 0 pairs with itself
    const int k1p = zero ? 0 : N1 - k1;
    const float2 *smr = smem + rr * RS, *smp = smem + prr * RS;
    const float hs = 0.5f * scale;
    // rows 1..N1/2-1: all N2 values of k2, the pair's other member is on the mirror row.
    // row 0: pairs (0,k2) <-> (0,N2-k2) for k2 in [1, N2/2), plus the two self-paired elements.
    B2F_UNROLL(LARGE_ROWS_PAIR_UNROLL)
    for (int m = 0; m < NPAIR; m++) {
      const int k2 = threadIdx.x / RH + m * KSTEP;
      const int pk2 = zero ? N2 - k2 : N2 - 1 - k2;
      if (zero && (k2 == 0 || k2 >= N2 / 2)) {
        if (k2 == 0) {
          const float2 v = smr[pad_idx(0)];
          dst[0] = make_float2((v.x + v.y) * hs, (v.x - v.y) * hs);
        } else if (k2 == N2 / 2) {
          const float2 v = smr[pad_idx(k2)];
          dst[N / 2] = cscale(v, scale);  // never visited by the reference (Q3)
        }
        continue;
      }
      float2 a = smr[pad_idx(k2)], bb = smp[pad_idx(pk2)];
      const int i = k1 + N1 * k2, j = k1p + N1 * pk2;  // i + j == N
      constexpr float kC[8] = {1.f, B2F_COS_PI_8, B2F_SQRT1_2, B2F_SIN_PI_8, 0.f, -B2F_SIN_PI_8, -B2F_SQRT1_2, -B2F_COS_PI_8};
      constexpr float kS[8] = {0.f, B2F_SIN_PI_8, B2F_SQRT1_2, B2F_COS_PI_8, 1.f, B2F_COS_PI_8, B2F_SQRT1_2, B2F_SIN_PI_8};
      const float2 h = m ? cmulc<false>(hw0, kC[m], kS[m]) : hw0;
      if (i < j) {
        rfft_pair_folded<false>(a, bb, h, hs);
      } else {
        rfft_pair_folded<false>(bb, a, cconj(h), hs);
      }
      dst[i] = a;
      dst[j] = bb;
    }
    // the self-mirrored row N1/2 lives in slot RH of group 0: pairs (N1/2,k2) <-> (N1/2, N2-1-k2)
    if (g == 0) {
      const float2 *smh = smem + RH * RS;
      for (int k2 = threadIdx.x; k2 < N2 / 2; k2 += NTHR) {
        const int pk2 = N2 - 1 - k2;
        float2 a = smh[pad_idx(k2)], bb = smh[pad_idx(pk2)];
        const int i = N1 / 2 + N1 * k2, j = N1 / 2 + N1 * pk2;
        rfft_pair_folded<false>(a, bb, __ldg(&hw[i]), hs);
        dst[i] = a;
        dst[j] = bb;
      }
    }
  }
}

// RBT rows per CTA: 16 (256 threads) for complex transforms; 32 (512 threads: 16 direct + 16 mirrored rows) for the
// fused real split, 16 kept for comparison (B2F_ROWS_RB16=1). grid.x = N1 / RBT.
template <int LOG1, int LOG2, int RBT>
struct RowsGeom {
  using L = LargeGeom<LOG1, LOG2>;
  static constexpr int THREADS = RBT * L::G2::T;
  static constexpr int GROUPS = L::N1 / RBT;
  static constexpr int SMEM = RBT * (L::G2::SMEM + 1) * (int)sizeof(float2);
};

template <int LOG1, int LOG2, bool INV, bool REAL, int RBT>
__global__ void __launch_bounds__(RowsGeom<LOG1, LOG2, RBT>::THREADS, LARGE_ROWS_WARPS * 32 / RowsGeom<LOG1, LOG2, RBT>::THREADS)
    large_rows_kernel(const float2 *scratch, float2 *out, const float2 *__restrict__ tw2,
                      const float2 *__restrict__ hw, int batch, float scale) {
  using L = LargeGeom<LOG1, LOG2>;
  constexpr int N = L::N, N2 = L::N2, T2 = L::G2::T, E = L::G2::E;
  extern __shared__ float2 smem[];
  // The thread's 16 first-pass inputs of the NEXT transform are requested while the current one is being written
  // out (their registers are free then): the scratch latency, a quarter of this kernel's stall samples when
  // every transform started with a cold load, is off the critical path.
  const float2 *row = scratch + (size_t)large_rows_k1<LOG1, REAL, RBT>(blockIdx.x, threadIdx.x / T2) * N2;
  const int t = threadIdx.x % T2;
  float2 xin[E];
  auto fetch = [&](int b) {
#pragma unroll
    for (int s = 0; s < E; s++) xin[s] = row[(size_t)b * N + first_pass_index<LOG2>(t, s)];
  };
  // last transform first: the tail of what the columns kernel has just written is still in the 126 MB L2
  int b = batch - 1 - (int)blockIdx.y;
  if (b >= 0) fetch(b);
  for (; b >= 0; b -= gridDim.y) {
    auto hook = [&]() {
      if (b - (int)gridDim.y >= 0) fetch(b - (int)gridDim.y);
    };
    large_rows_body<LOG1, LOG2, INV, REAL, false, decltype(hook), false, RBT, true>(
        scratch + (size_t)b * N, out + (size_t)b * N, smem, tw2, hw, scale, blockIdx.x, hook, xin);
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
