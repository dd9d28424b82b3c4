// fft_sm.cuh -- "one SM, one transform": the 32768-point complex FFT (= the 65536-point real FFT of BASELINE
// config 5a) in ONE pass over HBM, every transform carried start to finish by one persistent 512-thread CTA.
//
// What it replaces in the reference: the reorder launch + 15 radix-2 stage launches of Clcfft::fft()
// (cl_fft.cpp:138-151) and, for the real transform, the `conv` split launch behind them (cl_fft.cpp:178-191,
// 267-282), each of which re-reads and re-writes the whole array in global memory. Round 1 ran this length as a
// four-step pair of launches with a scratch matrix through HBM (fft_large.cuh: 2 x the algorithmic traffic).
//
// Factorisation N = 32 x 32 x 32, input index n = 1024 j1 + 32 j2 + j3, output index k = k1 + 32 k2 + 1024 k3:
//   P1      thread = column c = 32 j2 + j3 (two columns per thread): 32 loads at stride 1024 (a warp reads 256
//           contiguous bytes per j1), radix-32 over j1 -> k1, times W_N^(c k1).
//   pass A  warp = row k1, lane = j3: radix-32 over j2 -> k2, times W_1024^(j3 k2); the row is private to the
//           warp, so the pass needs no CTA barrier.
//   pass B  thread = (row k1, k2), a warp holds 16 rows x the two values {a, 31-a} of k2: radix-32 over j3 -> k3,
//           stores X[k1 + 32 k2 + 1024 k3] in 128-byte lines. Real transform: the split of cl_fft.cpp:178-191 is
//           fused here, see the comment at the kernel.
//
// A transform is 256 KiB; an SM has 227 KB of shared memory. So the rows are processed as two JOBS of 16 rows
// (139 KB of padded shared memory): P1 computes all 32 outputs of a column, stores the first job's 16 straight into
// the row buffers and parks the second job's 16 in TENSOR MEMORY (tcgen05.st, thread-private columns: 512 threads
// x 64 columns x 4 B = 128 KiB of the SM's 256 KiB TMEM), from where they are moved into the row buffers
// (tcgen05.ld) once the first job has left them. TMEM is used as what it physically is -- a second on-chip
// scratch-pad next to shared memory -- not for MMA. (A variant parking the same values in an L2-resident global
// scratch measured 2.34 vs 2.79 TB/s in the first version of this kernel and was dropped.)
// The next transform of the CTA is requested into L2 with cp.async.bulk.prefetch (TMA engine, no registers, no
// shared memory) while the current one is being computed, so P1's loads are L2 hits.
//
// HBM traffic: the algorithmic bytes, once. Shared-memory traffic per point: P1 store, A load/store, B load (+ TMEM
// round trip for half the points). Barriers: six per transform.
#pragma once

#include "fft_core.cuh"
#include "tma_utils.cuh"

// Timing experiments (results wrong by construction; never defined in a product build): -DB2F_SMX_NO_STG drops the
// global stores of pass B.
#ifdef B2F_SMX_NO_STG
#define B2F_SMX_STG(dst, val)                 \
  do {                                        \
    if ((val).x == 1.2345e30f) (dst) = (val); \
  } while (0)
#else
#define B2F_SMX_STG(dst, val) (dst) = (val)
#endif
// -DB2F_SMX_NO_FFT drops the butterflies and twiddles (data movement only); -DB2F_SMX_NO_LOAD drops the TMA loads.
#ifdef B2F_SMX_NO_FFT
#define B2F_SMX_FFT(x)
#else
#define B2F_SMX_FFT(x) x
#endif

#ifndef B2F_SMX_PDL
#define B2F_SMX_PDL 1
#endif
#ifndef B2F_SMX_FIRST_PF
#define B2F_SMX_FIRST_PF 1
#endif

namespace b2f {

// v[m] *= W^m for m = 1..31, given W^1, W^2, W^4, W^8, W^16 (base[b] = W^(2^b)): the other 26 powers are products
// built depth-first over the bits of m, so at most four partial products are alive at a time (26 complex
// multiplications, at most four roundings deep, ~3e-7) -- the full table of a pass would be 31 loads per
// butterfly through the LSU data pipe, which is the SM-side limit of the FFT kernels (fft_core.cuh).
template <int BIT, int M, bool HAVE, int OFF = 0>
__device__ __forceinline__ void tw_tree(float2 (&v)[32], float2 p, const float2 (&base)[5]) {
  if constexpr (BIT < 0) {
    if constexpr (HAVE) v[OFF + M] = cmul(v[OFF + M], p);
  } else {
    tw_tree<BIT - 1, M, HAVE, OFF>(v, p, base);
    float2 q = base[BIT];
    if constexpr (HAVE) q = cmul(p, q);
    tw_tree<BIT - 1, (M | (1 << BIT)), true, OFF>(v, q, base);
  }
}
// two independent 16-point butterflies on v[0..15] and v[16..31] (N = 2^14: two transforms per unit)
template <bool INV>
__device__ __forceinline__ void dft16x2(float2 (&v)[32]) {
  float2 a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    a[i] = v[i];
    b[i] = v[16 + i];
  }
  dft16<INV>(a);
  dft16<INV>(b);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    v[i] = a[i];
    v[16 + i] = b[i];
  }
}

// four independent 8-point butterflies (N = 2^13: four transforms per unit)
template <bool INV>
__device__ __forceinline__ void dft8x4(float2 (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = v[8 * g + i];
    dft8<INV>(a);
#pragma unroll
    for (int i = 0; i < 8; i++) v[8 * g + i] = a[i];
  }
}

// ---- tensor memory as a scratch-pad -----------------------------------------------------------------------------
// 32x32b shape: lane i of the issuing warp reads / writes 32-bit columns [col, col + n) of TMEM lane 32 (warp % 4) + i,
// i.e. storage private to the thread. Address = (lane << 16) | column.
namespace tmem {
__device__ __forceinline__ void alloc(uint32_t smem_dst, uint32_t cols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t taddr, uint32_t cols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// eight complex values <-> sixteen columns
__device__ __forceinline__ void st8(uint32_t taddr, const float2 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(__float_as_uint(v[0].x)), "r"(__float_as_uint(v[0].y)), "r"(__float_as_uint(v[1].x)), "r"(__float_as_uint(v[1].y)),
      "r"(__float_as_uint(v[2].x)), "r"(__float_as_uint(v[2].y)), "r"(__float_as_uint(v[3].x)), "r"(__float_as_uint(v[3].y)),
      "r"(__float_as_uint(v[4].x)), "r"(__float_as_uint(v[4].y)), "r"(__float_as_uint(v[5].x)), "r"(__float_as_uint(v[5].y)),
      "r"(__float_as_uint(v[6].x)), "r"(__float_as_uint(v[6].y)), "r"(__float_as_uint(v[7].x)), "r"(__float_as_uint(v[7].y))
      : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, float2 (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
        "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
      : "r"(taddr)
      : "memory");
}
// the values of ld8 may be used after this; passing them through the asm keeps the compiler from moving uses up
__device__ __forceinline__ void wait_ld8(float2 (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[2].x), "+f"(v[2].y), "+f"(v[3].x),
                 "+f"(v[3].y), "+f"(v[4].x), "+f"(v[4].y), "+f"(v[5].x), "+f"(v[5].y), "+f"(v[6].x), "+f"(v[6].y),
                 "+f"(v[7].x), "+f"(v[7].y)::"memory");
}
}  // namespace tmem

// ---- geometry ---------------------------------------------------------------------------------------------------------
struct SmGeom {
  static constexpr int LOGN = 15, N = 1 << LOGN;  // (LOG1 = 4: N = 2^14, two transforms per unit)
  static constexpr int THREADS = 512;
  // exchange layout of a row (written by pass A, read by pass B): [k2][j3] float2, 16 bytes of padding per k2 line
  // and 16 per row: pass B's 128-bit reads (8 lanes = 8 rows, or 8 rows of the other k2) and pass A's 64-bit writes
  // (32 consecutive j3) are both bank-conflict free. P1 writes the row as [j2][j3] into the first 8 KiB.
  static constexpr int K2S = 32 * 8 + 16;
  static constexpr int ROW = 32 * K2S + 16;
  static constexpr int RUN_C2R = 4096 + 16;           // staged bytes per run of the inverse real transform (514 columns)
  static constexpr int OFF_SP = 16 * ROW;             // staging of P1's round 1, runs j1 = 16..31 (64 KiB + 256 B)
  static constexpr int OFF_Z = OFF_SP + 16 * RUN_C2R; // row k1 = 0 of the real transform, [k2][k3] float2
  static constexpr int OFF_TWN = OFF_Z + 1024 * 8;    // P1 twiddle bases [5][32]: W_N^(j3 2^b)
  static constexpr int OFF_TWP = OFF_TWN + 5 * 32 * 8;  // P1 twiddle bases [5][32]: W_N^(32 j2 2^b)
  static constexpr int OFF_TWA = OFF_TWP + 5 * 32 * 8;  // pass-A twiddle bases [5][32]: W_1024^(j 2^b)
  static constexpr int OFF_HW = OFF_TWA + 5 * 32 * 8;   // folded split table, entries 0..1023
  static constexpr int OFF_HW32 = OFF_HW + 1024 * 8;    // folded split table, entries 32 m (row k1 = 0)
  static constexpr int OFF_MISC = OFF_HW32 + 512 * 8;
  static constexpr int SMEM = OFF_MISC + 32;  // TMEM base address, three mbarriers
  // tensor memory, per thread (4 warps share a lane quarter): columns [0, 64) hold the second job's P1 outputs of the
  // thread's two columns; real transform: columns [64, 128) park the first job's pass-B results
  static constexpr int TCOLS_THREAD_C = 64, TCOLS_THREAD_R = 128;
};

// exp(-i pi k / 32), k = 0..31: (kSplitC[k], -kSplitS[k])
__device__ constexpr float kSplitC[32] = {
    1.f,
    0.99518472667219693f,  0.98078528040323043f,  0.95694033573220882f,  0.92387953251128674f,  0.88192126434835505f,
    0.83146961230254524f,  0.77301045336273699f,  0.70710678118654757f,  0.63439328416364549f,  0.55557023301960229f,
    0.47139673682599770f,  0.38268343236508984f,  0.29028467725446233f,  0.19509032201612833f,  0.09801714032956077f,
    0.f,
    -0.09801714032956077f, -0.19509032201612833f, -0.29028467725446233f, -0.38268343236508984f, -0.47139673682599770f,
    -0.55557023301960229f, -0.63439328416364549f, -0.70710678118654757f, -0.77301045336273699f, -0.83146961230254524f,
    -0.88192126434835505f, -0.92387953251128674f, -0.95694033573220882f, -0.98078528040323043f, -0.99518472667219693f};
__device__ constexpr float kSplitS[32] = {
    0.f,
    0.09801714032956060f, 0.19509032201612825f, 0.29028467725446233f, 0.38268343236508978f, 0.47139673682599764f,
    0.55557023301960218f, 0.63439328416364549f, 0.70710678118654757f, 0.77301045336273699f, 0.83146961230254524f,
    0.88192126434835505f, 0.92387953251128674f, 0.95694033573220882f, 0.98078528040323043f, 0.99518472667219693f,
    1.f,
    0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f, 0.88192126434835505f,
    0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960218f,
    0.47139673682599764f, 0.38268343236508978f, 0.29028467725446233f, 0.19509032201612825f, 0.09801714032956060f};

// Job `job` holds rows k1 = 16 job + s, s = 0..15: pass B stores 128-byte lines.
//
// Real transform (the split of cl_fft.cpp:178-191 fused into pass B). The partner of element (k1, k2, k3) is
// (32 - k1, 31 - k2, 31 - k3): a row of the OTHER job, but the same warp (which holds k2 = a and 31 - a), the lane
// (16 - s, other k2) and the register 31 - k3. So job 0 stores nothing to global memory: its pass B parks the
// thread's 32 results in tensor memory; job 1 fetches them back eight at a time, shuffles them to the partner lane,
// evaluates every pair once and stores BOTH members -- its own (rows 16..31) and the partner's (rows 15..1) -- as
// full 128-byte lines. Row 16 is its own mirror (partner lane in the same job); row 0 is too, but with the irregular
// pattern (k2, k3) <-> (32 - k2, 31 - k3) and the two special elements (packed DC/Nyquist, untouched bin N/2:
// SURVEY Q2/Q3), so it is split in an 8 KiB shared-memory buffer Z and stored by the row-16 lanes, whose partner
// slot is free, completing the partner's lines.
//
// Inverse real transform (KIND = 2): the unsplit of cl_fft.cpp:192-205 is fused into P1's reads. The partner of input
// element (j1, c) is (31 - j1, 1024 - c) [(32 - j1, 0) in column 0], so the rounds are closed under c -> 1024 - c:
// round 0 = columns [0, 256) and [768, 1024), round 1 = [256, 768); the one pair that straddles them, (256, 768), is
// staged in both (runs of 514 columns). Every thread reads its own column and the mirrored one from the staged runs and
// keeps its own member of each pair. In the rows, the column groups j2 are stored permuted (0..7, 24..31, 8..23) so
// that round 0 still writes the first halves and round 1 the second.
//
// LOG1 = 4 (N = 16 x 32 x 32 = 2^14): the same kernel carries TWO transforms per unit of work -- the 32 staged runs
// are the 16 runs of each, P1 is two radix-16 butterflies per column, job 0 is the first transform's 16 rows and job 1
// the second's (parked in tensor memory meanwhile); passes A and B are unchanged. Real transforms: the unsplit pairs
// runs j1 <-> 15 - j1 of the same transform; the split stays inside a job (see pass B).
enum { kSmComplex = 0, kSmRealFwd = 1, kSmRealInv = 2 };
template <bool INV, int KIND, int LOG1 = 5>
__global__ void __launch_bounds__(SmGeom::THREADS, 1)
    fft_sm_kernel(const float2 *in, float2 *out, const float2 *__restrict__ twn_g, const float2 *__restrict__ twp_g,
                  const float2 *__restrict__ twa_g, const float2 *__restrict__ hw, int batch, float scale) {
  using G = SmGeom;
  constexpr bool REAL = (KIND == kSmRealFwd), C2R = (KIND == kSmRealInv);
  static_assert(!(REAL && INV) && !(C2R && !INV), "the split follows a forward, the unsplit precedes an inverse transform");
  static_assert(LOG1 == 5 || LOG1 == 4 || LOG1 == 3, "N = 2^15, two transforms of 2^14 or four of 2^13 per unit");
  static_assert(LOG1 != 3 || KIND != kSmRealFwd, "N = 2^13: complex and inverse real transforms only");
  constexpr int N1 = 1 << LOG1, NTR = 32 / N1;  // radix of P1, transforms per unit
  constexpr int N = N1 * 1024;
  const int units = (batch + NTR - 1) / NTR;
  constexpr int TCOLS = (REAL && LOG1 == 5) ? G::TCOLS_THREAD_R : G::TCOLS_THREAD_C;
  extern __shared__ __align__(16) unsigned char smraw[];
  unsigned char *rows = smraw;
  float2 *Z = reinterpret_cast<float2 *>(smraw + G::OFF_Z);
  unsigned char *sp = smraw + G::OFF_SP;
  float2 *twn = reinterpret_cast<float2 *>(smraw + G::OFF_TWN);
  float2 *twp = reinterpret_cast<float2 *>(smraw + G::OFF_TWP);
  float2 *twa = reinterpret_cast<float2 *>(smraw + G::OFF_TWA);
  float2 *hwb = reinterpret_cast<float2 *>(smraw + G::OFF_HW);
  float2 *hw32 = reinterpret_cast<float2 *>(smraw + G::OFF_HW32);
  uint32_t *misc = reinterpret_cast<uint32_t *>(smraw + G::OFF_MISC);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

#if B2F_SMX_PDL
  // Programmatic dependent launch: the next kernel of the stream may be scheduled from now on (its CTAs become
  // resident as ours exit and run their set-up); it reads and writes nothing the stream's earlier kernels touch
  // before its own griddepcontrol.wait below, which returns when those have completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  if (tid < 5 * 32) {
    twn[tid] = __ldg(&twn_g[tid]);
    twp[tid] = __ldg(&twp_g[tid]);
    twa[tid] = __ldg(&twa_g[tid]);
  }
  if constexpr (REAL || C2R)
    for (int i = tid; i < 1024; i += G::THREADS) hwb[i] = __ldg(&hw[i]);
  if constexpr (REAL) hw32[tid] = __ldg(&hw[N1 * ((tid >> 4) + 32 * (tid & 15))]);  // [k2][k3 < 16]: entry N1 m, m = k2 + 32 k3
  if (warp == 0) tmem::alloc((uint32_t)__cvta_generic_to_shared(misc), 4 * TCOLS);
  tmem::fence_before();
  __syncthreads();
  tmem::fence_after();
  // this thread's private columns: TMEM lane 32 (warp % 4) + lane, columns TCOLS (warp / 4) ...
  const uint32_t taddr = misc[0] + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TCOLS);
  const float hs = 0.5f * scale;
  // Both rounds of P1 are fed by the TMA engine (cp.async.bulk, SASS UBLKCP), 32 runs of 4 KiB each:
  //   round 0 (columns 0..511 of the NEXT transform) lands in the row buffers as soon as the last pass B of the
  //   current transform has read them (m_free: 16 warps arrived), so its HBM latency hides behind that pass's
  //   butterflies and stores;
  //   round 1 (columns 512..1023) is requested when round 0 is in registers and lands, while round 0 is being
  //   computed, where round 0's results do not go: runs 0..15 in the second halves of the 16 rows, runs 16..31 in
  //   a 64 KiB staging area.
  const uint32_t m_free = tma::smem_u32(misc + 2), m_full0 = tma::smem_u32(misc + 4), m_full1 = tma::smem_u32(misc + 6);
  constexpr int RUN = C2R ? G::RUN_C2R : 4096;
  // where run j1 of a round is staged
  auto run_base = [&](int round, int j) -> unsigned char * {
    return round == 0 ? rows + j * RUN : (j < 16 ? rows + j * G::ROW + 4096 : sp + (j - 16) * RUN);
  };
  auto stage_round = [&](int round, int un) {  // the 32 lanes of one warp, one run each
    // run `lane` of unit `un`: run lane % N1 of its transform lane / N1 (past the batch: the last transform again,
    // its results are not stored)
    int tr = un * NTR + lane / N1;
    tr = tr < batch ? tr : batch - 1;
    const float2 *nx = in + (size_t)tr * N + 1024 * (lane % N1);
    const uint32_t mb = round ? m_full1 : m_full0;
    const uint32_t d = tma::smem_u32(run_base(round, lane));
#ifdef B2F_SMX_NO_LOAD
    if (lane == 0) tma::mbar_arrive(mb);
    return;
#endif
    if (lane == 0) tma::mbar_expect_tx(mb, 32 * RUN);
    __syncwarp();
    if constexpr (!C2R) {
      tma::bulk_g2s(d, nx + 512 * round, 4096, mb);
    } else if (round == 0) {
      tma::bulk_g2s(d, nx, 258 * 8, mb);                   // columns [0, 258)
      tma::bulk_g2s(d + 258 * 8, nx + 768, 256 * 8, mb);   // columns [768, 1024)
    } else {
      tma::bulk_g2s(d, nx + 256, 514 * 8, mb);             // columns [256, 770)
    }
  };
  if (tid == 0) {
    tma::mbar_init(m_free, 16);
    tma::mbar_init(m_full0, 1);
    tma::mbar_init(m_full1, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
#if B2F_SMX_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  if (warp == 0) {
    stage_round(0, blockIdx.x);
#if B2F_SMX_FIRST_PF
    {  // the first unit's round 1 starts towards L2 together with round 0 (later units: during the previous unit's job 1)
      int tr = blockIdx.x * NTR + lane / N1;
      tr = tr < batch ? tr : batch - 1;
      tma::prefetch_l2(in + (size_t)tr * N + 1024 * (lane % N1) + 512, 4096);
    }
#endif
  }
  uint32_t par = 0;

  for (int t = blockIdx.x; t < units; t += gridDim.x, par ^= 1) {
    float2 *dst = out + (size_t)t * NTR * N;
    const bool more = t + (int)gridDim.x < units;

    // ---- P1: radix-32 over j1 for columns c = tid, tid + 512 ------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; r++) {
      // column of this thread and its position in the rows (complex / forward real: c itself)
      const int c = !C2R ? tid + 512 * r : (r == 0 ? (tid < 256 ? tid : tid + 512) : tid + 256);
      const int pos = !C2R ? c : (r == 0 ? tid : tid + 512);
      float2 v[32];
      tma::mbar_wait(r == 0 ? m_full0 : m_full1, par);
      if constexpr (!C2R) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = *reinterpret_cast<const float2 *>(run_base(r, j) + tid * 8);
      } else {
        // byte offset of a column inside a staged run
        auto off = [&](int col) { return r == 0 ? (col < 258 ? col * 8 : (col - 510) * 8) : (col - 256) * 8; };
        const bool col0 = (c == 0);
        const int oc = off(c), opc = off(col0 ? 0 : 1024 - c), pj0 = col0 ? N1 : N1 - 1;
        const float2 hbase = hwb[c];
        float2 raw[32];  // (only raw[0], raw[N1/2] of every transform are used, and only in column 0)
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int j1 = j % N1, jt = j - j1;  // run j1 of transform jt / N1 of the unit
          const float2 a = *reinterpret_cast<const float2 *>(run_base(r, j) + oc);
          // partner run N1 - 1 - j1 (N1 - j1 in column 0: a different run per lane there, so the address is computed)
          const int pj = jt + ((pj0 - j1) & (N1 - 1));
          const unsigned char *pb = r == 0 ? rows + pj * RUN : (pj < 16 ? rows + pj * G::ROW + 4096 : sp + (pj - 16) * RUN);
          const float2 b = *reinterpret_cast<const float2 *>(pb + opc);
          raw[j] = a;
          // hw(1024 j1 + c) = hw(c) exp(+i pi j1 / N1), extended analytically past N/2, where hw(N - i) = conj(hw(i))
          const float2 hk = j1 ? cmulc<true>(hbase, kSplitC[j1 * (32 / N1)], kSplitS[j1 * (32 / N1)]) : hbase;
          float2 A = a, B = b;
          if (j1 < N1 / 2) {
            rfft_pair_folded<true>(A, B, hk, 0.5f);
            v[j] = A;
          } else {
            rfft_pair_folded<true>(B, A, cconj(hk), 0.5f);
            v[j] = A;
          }
        }
        if (col0) {  // element 0: packed (DC, Nyquist); element N/2: never visited by the reference (cl_fft.cpp:286)
#pragma unroll
          for (int jt = 0; jt < 32; jt += N1) {
            v[jt] = rfft_dc<true>(raw[jt]);
            v[jt + N1 / 2] = raw[jt + N1 / 2];
          }
        }
      }
      __syncthreads();  // the round is in registers: its half of the row buffers may be written
      if (r == 0 && warp == 0) stage_round(1, t);
      // W_N^(c 2^b), c = 32 j2 + j3, as the product of two table entries
      float2 base[5];
#pragma unroll
      for (int b = 0; b < LOG1; b++) {
        base[b] = cmul(twp[b * 32 + (c >> 5)], twn[b * 32 + (c & 31)]);
        if (INV) base[b].y = -base[b].y;
      }
      if constexpr (LOG1 == 5) {
        B2F_SMX_FFT(dft32<INV>(v));
        B2F_SMX_FFT((tw_tree<4, 0, false>(v, make_float2(1.f, 0.f), base)));
      } else if constexpr (LOG1 == 4) {
        B2F_SMX_FFT(dft16x2<INV>(v));
        B2F_SMX_FFT((tw_tree<3, 0, false, 0>(v, make_float2(1.f, 0.f), base)));
        B2F_SMX_FFT((tw_tree<3, 0, false, 16>(v, make_float2(1.f, 0.f), base)));
      } else {
        B2F_SMX_FFT(dft8x4<INV>(v));
        B2F_SMX_FFT((tw_tree<2, 0, false, 0>(v, make_float2(1.f, 0.f), base)));
        B2F_SMX_FFT((tw_tree<2, 0, false, 8>(v, make_float2(1.f, 0.f), base)));
        B2F_SMX_FFT((tw_tree<2, 0, false, 16>(v, make_float2(1.f, 0.f), base)));
        B2F_SMX_FFT((tw_tree<2, 0, false, 24>(v, make_float2(1.f, 0.f), base)));
      }
#pragma unroll
      for (int s = 0; s < 16; s++) *reinterpret_cast<float2 *>(rows + s * G::ROW + pos * 8) = v[s];
#pragma unroll
      for (int g = 0; g < 2; g++) {
        float2 h8[8];
#pragma unroll
        for (int i = 0; i < 8; i++) h8[i] = v[16 + 8 * g + i];
        tmem::st8(taddr + 32 * r + 16 * g, h8);
      }
    }
    tmem::wait_st();
    __syncthreads();

#pragma unroll 1
    for (int job = 0; job < 2; job++) {
      // The CTA's next unit -> L2, one half per job (32 bulk prefetches of 4 KiB by the TMA engine, no registers, no
      // shared memory): the TMA copies that stage it later then find it in L2 instead of paying the HBM latency and
      // this SM's share of the HBM bandwidth in one burst. Columns 0..511 (staged at the end of this unit) during job
      // 0, columns 512..1023 (staged after the next unit's round 0 is in registers) during job 1 (measured: 0.144 ->
      // 0.138 ms real, 0.134 -> 0.122 ms complex per 1024 transforms; both halves during job 0: 0.139 / 0.124).
      if (tid < 32 && more) {
        int tr = (t + gridDim.x) * NTR + tid / N1;
        tr = tr < batch ? tr : batch - 1;
        const float2 *nx = in + (size_t)tr * N + 1024 * (tid % N1) + 512 * job;
        tma::prefetch_l2(nx, 4096);
      }
      if (job == 1) {
        // the second job's rows: tensor memory -> row buffers
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int c = tid + 512 * r;
          float2 ha[8], hb[8];
          tmem::ld8(taddr + 32 * r, ha);
          tmem::ld8(taddr + 32 * r + 16, hb);
          tmem::wait_ld8(ha);
          tmem::wait_ld8(hb);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            *reinterpret_cast<float2 *>(rows + i * G::ROW + c * 8) = ha[i];
            *reinterpret_cast<float2 *>(rows + (8 + i) * G::ROW + c * 8) = hb[i];
          }
        }
        __syncthreads();
      }

      // ---- pass A: warp = row slot, lane = j3, registers = j2 ---------------------------------------------------
      {
        unsigned char *row = rows + warp * G::ROW;
        float2 v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int pj = !C2R ? j : (j < 8 ? j : (j >= 24 ? j - 16 : j + 8));  // where column group j2 = j is stored
          v[j] = *reinterpret_cast<const float2 *>(row + (pj * 32 + lane) * 8);
        }
        float2 base[5];
#pragma unroll
        for (int b = 0; b < 5; b++) {
          base[b] = twa[b * 32 + lane];
          if (INV) base[b].y = -base[b].y;
        }
        __syncwarp();  // the whole row is in registers before its buffer is overwritten in the exchange layout
        B2F_SMX_FFT(dft32<INV>(v));
        B2F_SMX_FFT((tw_tree<4, 0, false>(v, make_float2(1.f, 0.f), base)));
#pragma unroll
        for (int k = 0; k < 32; k++) *reinterpret_cast<float2 *>(row + k * G::K2S + lane * 8) = v[k];
      }
      __syncthreads();

      // ---- pass B: thread = (row slot, k2), registers = j3 -> k3 --------------------------------------------------
      {
        const int s = lane & 15, h = lane >> 4;
        // Real transform, job 0: the lane takes the row and the k2 its job-1 self will need as partner, i.e. row
        // (16 - s) mod 16 and the warp's OTHER k2 -- what it parks in tensor memory is then its own partner data.
        const bool mirrored = REAL && LOG1 == 5 && job == 0;
        const int slot = mirrored ? ((16 - s) & 15) : s;
        const int k2 = (h != (int)mirrored) ? 31 - warp : warp;
        // LOG1 = 4: job = transform of the unit, rows 0..15 each; LOG1 = 3: two transforms per job, rows 0..7 each
        const int k1 = LOG1 == 5 ? 16 * job + slot : (LOG1 == 4 ? slot : (slot & 7));
        const int trj = LOG1 == 5 ? 0 : (LOG1 == 4 ? job : 2 * job + (slot >> 3));  // transform of the unit
        const unsigned char *p = rows + slot * G::ROW + k2 * G::K2S;
        float2 v[32];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const float4 f = *reinterpret_cast<const float4 *>(p + 16 * q);
          v[2 * q] = make_float2(f.x, f.y);
          v[2 * q + 1] = make_float2(f.z, f.w);
        }
        if (job == 1) {
          __syncwarp();
          if (lane == 0) tma::mbar_arrive(m_free);
          if (warp == 0 && more) {
            tma::mbar_wait(m_free, par);
            stage_round(0, t + gridDim.x);
          }
          __syncwarp();
        }
        B2F_SMX_FFT(dft32<INV>(v));
        float2 *o = dst + (size_t)trj * N + k1 + N1 * k2;
        if constexpr (!REAL) {
          if (LOG1 == 5 || t * NTR + trj < batch) {  // (the last unit of a batch that is not a multiple of NTR)
#pragma unroll
            for (int k3 = 0; k3 < 32; k3++) B2F_SMX_STG(o[32 * N1 * k3], cscale(v[k3], scale));
          }
        } else if constexpr (LOG1 == 4) {
          // Two transforms per unit: the partner (16 - k1, 31 - k2, 31 - k3) is in the SAME job -- the warp's other k2,
          // row slot 16 - s. Every lane hands the upper half of its results (k3 >= 16) over through the round-1 staging
          // area (idle during pass B; 16-byte chunk q of slot s sits at position q ^ (s & 7): conflict-free both ways),
          // evaluates the 16 pairs of its lower half and stores both members. Row 0 ((k2, k3) <-> (32 - k2, 31 - k3),
          // across warps) is split in Z between two barriers and stored by the s = 0 lanes in the same instructions.
          float4 *xch = reinterpret_cast<float4 *>(sp);  // [k2][slot][8 chunks]
          if (s == 0) {
#pragma unroll
            for (int q = 0; q < 16; q++)
              *reinterpret_cast<float4 *>(Z + 32 * k2 + 2 * q) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++)
              xch[(k2 * 16 + s) * 8 + (q ^ (s & 7))] = make_float4(v[16 + 2 * q].x, v[16 + 2 * q].y, v[17 + 2 * q].x, v[17 + 2 * q].y);
          }
          __syncthreads();
          {  // row k1 = 0: X[N1 m], m = k2 + 32 k3, pairs (m, 1024 - m); m = 0 packed (DC, Nyquist), m = 512 is bin N/2
            const int zk3 = tid & 15, zk2 = tid >> 4, m = zk2 + 32 * zk3;
            if (m == 0) {
              const float2 z0 = Z[0];
              Z[0] = make_float2((z0.x + z0.y) * hs, (z0.x - z0.y) * hs);
              Z[16] = cscale(Z[16], scale);
            } else {
              const int pi = zk2 ? 32 * (32 - zk2) + 31 - zk3 : 32 - zk3;
              float2 a = Z[32 * zk2 + zk3], b = Z[pi];
              rfft_pair_folded<false>(a, b, hw32[tid], hs);
              Z[32 * zk2 + zk3] = a;
              Z[pi] = b;
            }
          }
          __syncthreads();
          float2 pk[16];  // pk[i]: result 16 + i of the partner (row 0: of row 0, column group 31 - k2)
          if (s == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float4 f = *reinterpret_cast<const float4 *>(Z + 32 * k2 + 2 * q);
              const float4 g = *reinterpret_cast<const float4 *>(Z + 32 * (31 - k2) + 16 + 2 * q);
              v[2 * q] = make_float2(f.x, f.y);
              v[2 * q + 1] = make_float2(f.z, f.w);
              pk[2 * q] = make_float2(g.x, g.y);
              pk[2 * q + 1] = make_float2(g.z, g.w);
            }
          } else {
            const int ps = 16 - s;
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float4 g = xch[((31 - k2) * 16 + ps) * 8 + (q ^ (ps & 7))];
              pk[2 * q] = make_float2(g.x, g.y);
              pk[2 * q + 1] = make_float2(g.z, g.w);
            }
            const float2 hbase = hwb[k1 + N1 * k2];
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) {
              const float2 hk = k3 ? cmulc<false>(hbase, kSplitC[k3], kSplitS[k3]) : hbase;
              rfft_pair_folded<false>(v[k3], pk[15 - k3], hk, hs);
            }
          }
          if (t * NTR + job < batch) {
            float2 *om = dst + (size_t)job * N + ((16 - s) & 15) + N1 * (31 - k2) + 32 * N1 * 31;
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) {
              B2F_SMX_STG(o[32 * N1 * k3], v[k3]);
              B2F_SMX_STG(om[-32 * N1 * k3], pk[15 - k3]);
            }
          }
        } else if (job == 0) {
          if (s == 0) {  // row 0 -> Z[k2][k3]
#pragma unroll
            for (int q = 0; q < 16; q++)
              *reinterpret_cast<float4 *>(Z + 32 * k2 + 2 * q) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
          }
#pragma unroll
          for (int g = 0; g < 4; g++) {
            float2 h8[8];
#pragma unroll
            for (int i = 0; i < 8; i++) h8[i] = v[8 * g + i];
            tmem::st8(taddr + 64 + 16 * g, h8);
          }
          tmem::wait_st();
        } else {
          // k = k1 + 32 k2 + 1024 k3 (own, rows 16..31), N - k = (16 - s) + 32 (31 - k2) + 1024 (31 - k3) (partner).
          // hw(k) = hw(k1 + 32 k2) exp(-i pi k3 / 32), extended analytically past N/2, where hw(N - k) = conj(hw(k)).
          const float2 hbase = hwb[k1 + 32 * k2];
          const bool s0 = (s == 0);
          // Row 16 (s = 0) is its own mirror: the partner is the warp's other row-16 lane. The two swap their results
          // through Z16 (8 KiB of the round-1 staging area, idle now). Their partner slot in the stores is free and
          // takes row 0, element (31 - k2, 31 - k3), already split in Z.
          float2 *z16 = reinterpret_cast<float2 *>(sp);
          if (s0) {
#pragma unroll
            for (int q = 0; q < 16; q++)
              *reinterpret_cast<float4 *>(z16 + 32 * k2 + 2 * q) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
          }
          __syncwarp();
          float2 *om = dst + (N - (k1 + 32 * k2)) - (s0 ? 16 : 0);
          const float2 *zp = Z + 32 * (31 - k2);
#pragma unroll
          for (int g = 0; g < 4; g++) {
            float2 pk[8];  // pk[i] = the partner's result 24 - 8 g + i
            tmem::ld8(taddr + 64 + 2 * (24 - 8 * g), pk);
            tmem::wait_ld8(pk);
            if (s0) {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const float4 f = *reinterpret_cast<const float4 *>(z16 + 32 * (31 - k2) + 24 - 8 * g + 2 * q);
                pk[2 * q] = make_float2(f.x, f.y);
                pk[2 * q + 1] = make_float2(f.z, f.w);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int k3 = 8 * g + j;
              const float2 hk = k3 ? cmulc<false>(hbase, kSplitC[k3], kSplitS[k3]) : hbase;
              if (k3 < 16) {
                rfft_pair_folded<false>(v[k3], pk[7 - j], hk, hs);
              } else {
                rfft_pair_folded<false>(pk[7 - j], v[k3], cconj(hk), hs);
              }
            }
            if (s0) {  // partner slot of the row-16 lanes: row 0, elements (31 - k2, 24 - 8 g + i)
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const float4 f = *reinterpret_cast<const float4 *>(zp + 24 - 8 * g + 2 * q);
                pk[2 * q] = make_float2(f.x, f.y);
                pk[2 * q + 1] = make_float2(f.z, f.w);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int k3 = 8 * g + j;
              B2F_SMX_STG(o[1024 * k3], v[k3]);
              B2F_SMX_STG(om[-1024 * k3], pk[7 - j]);
            }
          }
        }
      }
      // the row buffers are free again and Z is complete (job 1: m_free, and P1's barrier; real LOG1 = 4: pass B's own)
      if (job == 0 && !(REAL && LOG1 == 4)) __syncthreads();

      if constexpr (REAL && LOG1 == 5) {
        if (job == 0) {
          // row k1 = 0: X[32 m], m = k2 + 32 k3, pairs (m, 1024 - m), split in place; m = 0 is the packed (DC, Nyquist)
          // element and m = 512 is bin N/2, which the reference's split never visits (cl_fft.cpp:278).
          // Read again in job 1's pass B, several barriers from here.
          const int k3 = tid & 15, k2 = tid >> 4, m = k2 + 32 * k3;  // m in [0, 512); Z is laid out [k2][k3]
          if (m == 0) {
            const float2 z0 = Z[0];
            Z[0] = make_float2((z0.x + z0.y) * hs, (z0.x - z0.y) * hs);
            Z[16] = cscale(Z[16], scale);  // m = 512
          } else {
            const int pi = k2 ? 32 * (32 - k2) + 31 - k3 : 32 - k3;  // 1024 - m
            float2 a = Z[32 * k2 + k3], b = Z[pi];
            rfft_pair_folded<false>(a, b, hw32[tid], hs);
            Z[32 * k2 + k3] = a;
            Z[pi] = b;
          }
        }
      }
    }
  }
  tmem::fence_before();
  __syncthreads();
  if (warp == 0) tmem::dealloc(misc[0], 4 * TCOLS);
}

}  // namespace b2f
