// pconv_kernels.cuh -- uniformly-partitioned convolution, one fused launch per streaming block.
//
// Replaces cl_conv::Clpconv's 24 launches + 2 copies per block (cl_conv.cpp:393-458; kernels
// cl_conv_kernels.h:46-124): reorder, 9x fft, r2c, convol (bsize work-items x 2 CAS atomics),
// c2r, reorder, 9x fft, olap. Here, per channel, one CTA (or one thread-block cluster that splits
// the partitions and reduces through distributed shared memory) does
//   forward real FFT of the new block -> frequency-domain delay line (FDL) frame `wp`
//   Y[n] = sum_p FDL[(wp+1+p) mod nparts][n] * IR[p][n]   streamed with 128-bit loads, accumulated in
//          registers in ascending p (deterministic; the reference's atomic order is unspecified, Q8)
//   unsplit -> inverse FFT -> /pts -> overlap-add with the saved tail -> out, new tail
// Algorithmic HBM traffic per channel-block: 8*pts*(2*nparts+3) bytes; the MAC loop moves
// all but 3/(2*nparts+3) of it.
//
// Data layout in HBM (channel-major, everything contiguous per channel):
//   fdl  [channels][nparts][pts] float2   the reference's spec1 ring, same frame order
//   irs  [channels][nparts][pts] float2   the reference's spec2 ring, same frame order
//   tail [channels][pts]         float    saved second half of the last inverse FFT, unnormalised
#pragma once

#include <cooperative_groups.h>

#include "fft_core.cuh"
#include "fft_kernels.cuh"
#include "tma_utils.cuh"

namespace b2f {

namespace cg = cooperative_groups;


constexpr int kPconvMaxLogP = 12;  // fused path: pts <= 4096 (shared-memory budget incl. time-varying)

template <int LOGP>
struct PconvGeom {
  using G = FftGeom<LOGP>;
  static constexpr int PTS = 1 << LOGP;
  static constexpr int T = G::T;                           // threads carrying the real transform
  static constexpr int FT = T < 32 ? 32 : T;               // FFT participants (whole warps)
  static constexpr int VT = FT / T;                        // virtual transforms (only #0 is real)
  static constexpr int HALF = PTS / 2 < 1 ? 1 : PTS / 2;   // float4 per frame
  static constexpr int THREADS = HALF < 32 ? 32 : (HALF > 256 ? 256 : HALF);
  static constexpr int NTHREADS = THREADS > FT ? THREADS : FT;
  static constexpr int TILES = (HALF + NTHREADS - 1) / NTHREADS;
  static constexpr int FFT_SMEM = VT * G::SMEM;            // float2 entries per FFT work buffer
  // TMA-fed MAC: one extra (producer) warp, a ring of STAGES x {FDL slice, IR slice} of SLICE float4 each
  // (4096-sample partitions leave room for one CTA per SM only: a deeper ring keeps enough bytes in flight)
  // pts 4096: 5 stages of 8 KB, so that TWO CTAs fit an SM (107 KB each) and one streams while the other runs its
  // transforms. Measured on one box, 1024 / 256 channels x 117 partitions, fraction of the HBM peak: one CTA per SM
  // with 10 stages 0.81 / 0.71, 20 stages 0.80 / 0.70, 20 stages + ring filled before the forward transform
  // 0.77 / 0.67 -- depth was never the limit, the serial phases of a lone CTA were.
#ifndef B2F_PCONV_STAGES12
#define B2F_PCONV_STAGES12 5  // (build-time knob kept for re-measurement)
#endif
  static constexpr int SLICE = HALF < NTHREADS ? HALF : NTHREADS;  // float4 per slice (16 B .. 4 KB)
  // DEEP ring, for launches with at most one CTA per SM (a few channels: the mono streams of csound/tests.py). A lone
  // CTA is paced by its per-stage hand-shake (wait, two 128-bit reads, one multiply-add, arrive: ~470 cycles per 8 KB
  // measured, 34-43 GB/s per CTA), not by the bytes in flight -- a 24-stage ring of the same 4 KB slices was SLOWER
  // (mono pts 2048 x 2048 partitions 121 -> 125 us, pts 512 x 8192 66 -> 127 us). So a deep stage holds DEEP_TILES
  // tiles of a partition (16 KB of FDL + 16 KB of IR): one hand-shake per 32 KB, and the ring takes the shared memory
  // that the FFT buffers and the cluster partials leave.
  static constexpr int DEEP_TILES = TILES > 4 ? 4 : TILES;
  __host__ __device__ static constexpr int deep_stages() {
    const int fixed = 2 * (FFT_SMEM + 1) * 8 + HALF * 16;  // both FFT buffers (time-varying) + cluster partials
    const int n = (224 * 1024 - fixed) / (2 * DEEP_TILES * SLICE * 16 + 16);
    return n > 12 ? 12 : (n < 2 ? 2 : n);
  }
  __host__ __device__ static constexpr int stages(bool deep) { return deep ? deep_stages() : (LOGP >= 12 ? B2F_PCONV_STAGES12 : 6); }
  __host__ __device__ static constexpr int ring_f4(bool deep) { return 2 * stages(deep) * SLICE * (deep ? DEEP_TILES : 1); }
  // frames wider than the CTA whose tiles are swept together, one accumulator per tile (see pconv_step_kernel)
  __host__ __device__ static constexpr bool pmajor(bool tma) { return TILES > 1 && (tma || TILES == 2 || TILES == 4); }
  // float4 entries between the FFT buffers and the TMA ring: the cluster-partials buffer (clusters only), then the
  // buffer the tile-by-tile sweep parks Y in
  __host__ __device__ static constexpr int partial_f4(bool tma, int S) {
    return (S > 1 ? HALF : 0) + (TILES > 1 && !pmajor(tma) ? HALF : 0);
  }
};

// barrier over the FFT participants only (warps 0 .. FT/32-1); id 1, the CTA barrier is id 0
template <int COUNT>
struct FftGroupSync {
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, %0;" ::"n"(COUNT) : "memory"); }
};

// R(x): zero-pad pts reals to 2*pts, pack as pts complex, unscaled forward FFT, real-FFT split.
// (cl_conv.cpp:399-419 / 361-380; kernels cl_conv_kernels.h:46-85.) Result left in sm (padded index).
// Called by ALL threads of the CTA; x may be nullptr for CTAs that only need the barriers.
// NT: threads of the CTA taking part in the split loop (the step kernel: its NTHREADS workers; push_ir: the FFT group)
template <int LOGP, int NT = PconvGeom<LOGP>::NTHREADS>
__device__ __forceinline__ void pconv_forward_frame(const float *x, float2 *sm, const float2 *__restrict__ tw,
                                                    const float2 *__restrict__ w2) {
  using P = PconvGeom<LOGP>;
  constexpr int N = P::PTS;
  const int tid = threadIdx.x;
  if (tid < P::FT) {
    const int vt = tid / P::T, t = tid % P::T;
    float2 *my = sm + vt * FftGeom<LOGP>::SMEM;
    const bool real = (vt == 0);
    // an IR pushed with an odd channel stride (push_ir_dev) leaves rows that are only 4-byte aligned
    const bool pair_ok = (reinterpret_cast<uintptr_t>(x) & 7) == 0;
    auto load = [&](int idx, int) {
      if (real && idx < N / 2)
        return pair_ok ? *reinterpret_cast<const float2 *>(x + 2 * idx) : make_float2(x[2 * idx], x[2 * idx + 1]);
      return make_float2(0.f, 0.f);
    };
    auto store = [&](int idx, float2 v, int) { my[pad_idx(idx)] = v; };
    fft_run<LOGP, false, true>(load, store, my, tw, t, FftGroupSync<P::FT>());
  }
  __syncthreads();
  for (int i = tid; i < N / 2 && tid < NT; i += NT) {  // (a TMA producer warp only syncs)
    if (i == 0) {
      sm[pad_idx(0)] = rfft_dc<false>(sm[pad_idx(0)]);
    } else {
      float2 ci = sm[pad_idx(i)], cj = sm[pad_idx(N - i)];
      rfft_pair<false>(ci, cj, __ldg(&w2[i]));
      sm[pad_idx(i)] = ci;
      sm[pad_idx(N - i)] = cj;
    }
  }
  __syncthreads();
}
// pts == 2 corner: N/2 == 1, the pair loop above is empty except i == 0; element 1 (= N/2) untouched. OK.

__device__ __forceinline__ void cmac2(float4 &acc, const float4 a, const float4 b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
  acc.z += a.z * b.z - a.w * b.w;
  acc.w += a.z * b.w + a.w * b.z;
}

// acc += sum_{p < count} f[p] (*) g[p]; f/g advance one frame (stride4 float4) per p.
// acc0 accumulates the reference's packed-bin product (re*re, im*im) (cl_conv_kernels.h:114-115) for
// the first of the two bins; only the thread owning bin 0 uses it.
template <int U>
__device__ __forceinline__ void mac_segment(float4 &acc, float2 &acc0, const float4 *f, const float4 *g, int count,
                                            size_t stride4) {
  int p = 0;
  for (; p + U <= count; p += U) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      a[u] = __ldcs(f + (size_t)(p + u) * stride4);
      b[u] = __ldcs(g + (size_t)(p + u) * stride4);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      cmac2(acc, a[u], b[u]);
      acc0.x += a[u].x * b[u].x;
      acc0.y += a[u].y * b[u].y;
    }
  }
  for (; p < count; p++) {
    float4 a = __ldcs(f + (size_t)p * stride4), b = __ldcs(g + (size_t)p * stride4);
    cmac2(acc, a, b);
    acc0.x += a.x * b.x;
    acc0.y += a.y * b.y;
  }
}

// The same for a frame wider than the CTA: TILES float4 per thread per frame (tile t at offset t*NT), walked
// partition by partition so that the CTA reads every frame front to back in one go (two sequential streams per
// CTA, as in the one-tile case) instead of sweeping all partitions once per tile with gaps of (TILES-1)/TILES of a
// frame between its reads. 16 loads in flight per thread as above.
template <int TILES, int NT>
__device__ __forceinline__ void mac_segment_tiles(float4 (&acc)[TILES], float2 &acc0, const float4 *f, const float4 *g,
                                                  int count, size_t stride4) {
  constexpr int U = TILES >= 8 ? 1 : 8 / TILES;
  int p = 0;
  for (; p + U <= count; p += U) {
    float4 a[U][TILES], b[U][TILES];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int t = 0; t < TILES; t++) {
        a[u][t] = __ldcs(f + (size_t)(p + u) * stride4 + t * NT);
        b[u][t] = __ldcs(g + (size_t)(p + u) * stride4 + t * NT);
      }
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int t = 0; t < TILES; t++) cmac2(acc[t], a[u][t], b[u][t]);
      acc0.x += a[u][0].x * b[u][0].x;
      acc0.y += a[u][0].y * b[u][0].y;
    }
  }
  for (; p < count; p++) {
#pragma unroll
    for (int t = 0; t < TILES; t++) {
      const float4 a = __ldcs(f + (size_t)p * stride4 + t * NT), b = __ldcs(g + (size_t)p * stride4 + t * NT);
      cmac2(acc[t], a, b);
      if (t == 0) {
        acc0.x += a.x * b.x;
        acc0.y += a.y * b.y;
      }
    }
  }
}

struct PconvArgs {
  float2 *fdl;         // [channels][nparts][pts]
  float2 *irs;         // [channels][nparts][pts]
  float *tail;         // [channels][pts]
  const float *in1;    // [channels][pts]
  const float *in2;    // [channels][pts] (time-varying only)
  float *out;          // [channels][pts]
  const float2 *tw;    // pass twiddles of the pts-point plan
  const float2 *w2;    // split twiddles exp(-i pi k / pts)
  int nparts;
  int wp;              // FDL write frame for this block (before the increment, cl_conv.cpp:406)
  int wp2;             // IR write frame for this block (time-varying, cl_conv.cpp:487)
};

// grid = (S, channels), cluster = (S,1,1). Rank r of a cluster handles partitions
// [r*nparts/S, (r+1)*nparts/S) minus the frames written in this very launch, which rank 0 takes
// from its shared memory instead.
// TMA: the partitions are streamed by the TMA engine (cp.async.bulk, one producer warp, an mbarrier ring in
// shared memory) instead of 128-bit loads into registers; block = NTHREADS + 32.
template <int LOGP, bool TV, bool TMA, bool DEEP = false>
__global__ void __launch_bounds__(PconvGeom<LOGP>::NTHREADS + (TMA ? 32 : 0)) pconv_step_kernel(PconvArgs a) {
  static_assert(TMA || !DEEP, "the deep ring is the TMA feed's");
  using P = PconvGeom<LOGP>;
  constexpr int PTS = P::PTS, HALF = P::HALF, NT = P::NTHREADS;
  extern __shared__ float4 smem4[];
  float2 *sX = reinterpret_cast<float2 *>(smem4);         // new input spectrum, later Y / IFFT buffer
  float2 *sG = sX + P::FFT_SMEM + (P::FFT_SMEM & 1);      // new IR spectrum (TV only)
  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  float4 *sP = reinterpret_cast<float4 *>(sG + (TV ? P::FFT_SMEM + (P::FFT_SMEM & 1) : 0));  // cluster partials [HALF], S > 1
  float4 *sPark = sP + (S > 1 ? HALF : 0);                // Y of the tile-by-tile sweep [HALF]
  constexpr int STAGES = P::stages(DEEP), RING_F4 = P::ring_f4(DEEP);
  constexpr bool PMAJOR = P::pmajor(TMA);
  float4 *ring = sP + P::partial_f4(TMA, S);              // TMA ring: [STAGES][2][SLICE] float4, then barriers
  const bool worker = !TMA || threadIdx.x < NT;           // false only for the producer warp
  uint32_t bar_full = 0, bar_empty = 0, slot = 0;         // slot: running ring position, identical in all threads
  if (TMA) {
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + RING_F4);
    bar_full = tma::smem_u32(bars);
    bar_empty = tma::smem_u32(bars + STAGES);
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; s++) {
        tma::mbar_init(bar_full + 8 * s, 1);         // the producer's expect_tx arrive + the bytes
        tma::mbar_init(bar_empty + 8 * s, NT / 32);  // one arrive per consumer warp
      }
      tma::fence_barrier_init();
    }
    __syncthreads();
  }

  const int ch = blockIdx.y;
  const int tid = threadIdx.x;
  const int nparts = a.nparts;
  const size_t chan_off = (size_t)ch * nparts * PTS;
  float2 *fdl = a.fdl + chan_off;
  float2 *irs = a.irs + chan_off;

  // After the reference's increment (cl_conv.cpp:424) the read base is rp = wp+1 (mod nparts): the
  // oldest frame. Partition p pairs FDL frame (rp+p) mod nparts with IR frame p; p = nparts-1 is the
  // frame just written. In TV mode IR frame wp2 is also new.
  const int rp = (a.wp + 1 == nparts) ? 0 : a.wp + 1;
  // Rank 0 also runs the serial part (forward FFT before, unsplit + inverse FFT + overlap-add after), so in a
  // cluster of 4 or more it takes no regular partitions at all: the other ranks share them and rank 0's FFT
  // overlaps their streaming instead of preceding its own.
  int p_lo, p_hi;
  if (S >= 4) {
    p_lo = rank == 0 ? 0 : (int)((long long)(rank - 1) * nparts / (S - 1));
    p_hi = rank == 0 ? 0 : (int)((long long)rank * nparts / (S - 1));
  } else {
    p_lo = (int)((long long)rank * nparts / S);
    p_hi = (int)((long long)(rank + 1) * nparts / S);
  }
  const int p_newx = nparts - 1;
  const int p_newg = TV ? a.wp2 : -1;
  const size_t stride4 = HALF;

  // ---- 1. forward transforms of the new block(s): rank 0 only -------------------------------------
  if (rank == 0) {
    pconv_forward_frame<LOGP>(a.in1 + (size_t)ch * PTS, sX, a.tw, a.w2);
    if (TV) pconv_forward_frame<LOGP>(a.in2 + (size_t)ch * PTS, sG, a.tw, a.w2);
    // store the new frames for future blocks (this launch never reads them back from HBM)
    float2 *fx = fdl + (size_t)a.wp * PTS;
    for (int i = tid; i < PTS && worker; i += NT) fx[i] = sX[pad_idx(i)];
    if (TV) {
      float2 *gx = irs + (size_t)a.wp2 * PTS;
      for (int i = tid; i < PTS && worker; i += NT) gx[i] = sG[pad_idx(i)];
    }
  }

  // ---- 2. spectral multiply-accumulate over this rank's partitions --------------------------------

  if constexpr (PMAJOR) {
    // ---- a frame several times wider than the CTA: all tiles of a partition together, one accumulator per tile.
    // Register-fed (measured, 256 channels x 480000 taps: pts 1024 4.48 -> 6.36 TB/s, pts 2048 4.14 -> 4.65; pts 4096
    // (8 tiles) 4.09 -> 3.78, so the register-fed 8-tile case keeps the tile-by-tile sweep below) or TMA-fed -----
    constexpr int TL = P::TILES;
    float4 acc[TL];
#pragma unroll
    for (int t = 0; t < TL; t++) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 acc0 = make_float2(0.f, 0.f);
    const float4 *F = reinterpret_cast<const float4 *>(fdl) + tid;
    const float4 *Gp = reinterpret_cast<const float4 *>(irs) + tid;
    if constexpr (TMA) {
      // TMA-fed, partition-major: the producer streams every frame front to back -- tile after tile of partition p,
      // then partition p + 1 -- so HBM sees two sequential streams per CTA, and the consumers keep one accumulator
      // per tile in registers. (Round 1 swept all partitions once per tile, 4 KB slices 32 KB apart at pts 4096:
      // 0.63-0.68 of the measured HBM peak.) Every thread walks the same (p, tile) sequence, so `slot` advances
      // identically in the producer and in the consumers.
      constexpr uint32_t kSliceBytes = P::SLICE * (uint32_t)sizeof(float4);
      const unsigned char *Fb = reinterpret_cast<const unsigned char *>(fdl);
      const unsigned char *Gb = reinterpret_cast<const unsigned char *>(irs);
      const size_t frame_bytes = (size_t)PTS * sizeof(float2);
      if constexpr (DEEP) {
        // one stage = DT tiles of a partition: a single hand-shake per 2 x DT x 4 KB
        constexpr int DT = P::DEEP_TILES, GROUPS = TL / DT;
        constexpr uint32_t kStageBytes = DT * kSliceBytes;
        for (int p = p_lo; p < p_hi; p++) {
          if (p == p_newx || p == p_newg) continue;
          const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
#pragma unroll
          for (int g = 0; g < GROUPS; g++) {
            const uint32_t s = slot % STAGES, round = slot / STAGES;
            float4 *stF = ring + (size_t)(2 * s) * DT * P::SLICE, *stG = stF + DT * P::SLICE;
            if (!worker) {
              if ((tid & 31) == 0) {
                if (round > 0) tma::mbar_wait(bar_empty + 8 * s, (round - 1) & 1);
                tma::mbar_expect_tx(bar_full + 8 * s, 2 * kStageBytes);
                tma::bulk_g2s(tma::smem_u32(stF), Fb + (size_t)frame * frame_bytes + g * kStageBytes, kStageBytes, bar_full + 8 * s);
                tma::bulk_g2s(tma::smem_u32(stG), Gb + (size_t)p * frame_bytes + g * kStageBytes, kStageBytes, bar_full + 8 * s);
              }
            } else {
              tma::mbar_wait(bar_full + 8 * s, round & 1);
#pragma unroll
              for (int u = 0; u < DT; u++) {
                const float4 fa = stF[u * P::SLICE + tid], gb = stG[u * P::SLICE + tid];
                cmac2(acc[g * DT + u], fa, gb);
                if (g == 0 && u == 0) {
                  acc0.x += fa.x * gb.x;
                  acc0.y += fa.y * gb.y;
                }
              }
              __syncwarp();
              if ((tid & 31) == 0) tma::mbar_arrive(bar_empty + 8 * s);
            }
            slot++;
          }
        }
      } else
      for (int p = p_lo; p < p_hi; p++) {
        if (p == p_newx || p == p_newg) continue;
        const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
#pragma unroll
        for (int t = 0; t < TL; t++) {
          const uint32_t s = slot % STAGES, round = slot / STAGES;
          if (!worker) {
            if ((tid & 31) == 0) {
              if (round > 0) tma::mbar_wait(bar_empty + 8 * s, (round - 1) & 1);  // consumers released the stage
              tma::mbar_expect_tx(bar_full + 8 * s, 2 * kSliceBytes);
              tma::bulk_g2s(tma::smem_u32(ring + (2 * s) * P::SLICE), Fb + (size_t)frame * frame_bytes + t * kSliceBytes,
                            kSliceBytes, bar_full + 8 * s);
              tma::bulk_g2s(tma::smem_u32(ring + (2 * s + 1) * P::SLICE), Gb + (size_t)p * frame_bytes + t * kSliceBytes,
                            kSliceBytes, bar_full + 8 * s);
            }
          } else {
            tma::mbar_wait(bar_full + 8 * s, round & 1);
            const float4 fa = ring[(2 * s) * P::SLICE + tid], gb = ring[(2 * s + 1) * P::SLICE + tid];
            cmac2(acc[t], fa, gb);
            if (t == 0) {
              acc0.x += fa.x * gb.x;
              acc0.y += fa.y * gb.y;
            }
            __syncwarp();
            if ((tid & 31) == 0) tma::mbar_arrive(bar_empty + 8 * s);
          }
          slot++;
        }
      }
      __syncwarp();  // the producer warp's lane 0 rejoins its warp before the next (aligned) barrier
    } else {
      int p = p_lo;
      while (p < p_hi) {
        if (p == p_newx || p == p_newg) {
          p++;
          continue;
        }
        int end = p_hi;
        if (p_newx > p && p_newx < end) end = p_newx;
        if (p_newg > p && p_newg < end) end = p_newg;
        const int wrap = nparts - rp;  // first p whose FDL frame index wraps to 0
        if (wrap > p && wrap < end) end = wrap;
        const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
        mac_segment_tiles<TL, NT>(acc, acc0, F + (size_t)frame * stride4, Gp + (size_t)p * stride4, end - p, stride4);
        p = end;
      }
    }
    if (worker) {
#pragma unroll
    for (int t = 0; t < TL; t++) {
      const int q = t * NT + tid;
      if (rank == 0) {
        // the terms that involve frames produced by this launch, from shared memory
        const int b0 = 2 * q;
        float2 x0 = sX[pad_idx(b0)], x1 = sX[pad_idx(b0 + 1)];
        float4 xn = make_float4(x0.x, x0.y, x1.x, x1.y);
        float4 gn;
        if (TV && p_newg == p_newx) {
          float2 g0 = sG[pad_idx(b0)], g1 = sG[pad_idx(b0 + 1)];
          gn = make_float4(g0.x, g0.y, g1.x, g1.y);
        } else {
          gn = __ldcs(Gp + t * NT + (size_t)p_newx * stride4);
        }
        cmac2(acc[t], xn, gn);
        if (t == 0) {
          acc0.x += xn.x * gn.x;
          acc0.y += xn.y * gn.y;
        }
        if (TV && p_newg != p_newx) {
          const int frame = (rp + p_newg < nparts) ? rp + p_newg : rp + p_newg - nparts;
          float4 fo = __ldcs(F + t * NT + (size_t)frame * stride4);
          float2 g0 = sG[pad_idx(b0)], g1 = sG[pad_idx(b0 + 1)];
          float4 gg = make_float4(g0.x, g0.y, g1.x, g1.y);
          cmac2(acc[t], fo, gg);
          if (t == 0) {
            acc0.x += fo.x * gg.x;
            acc0.y += fo.y * gg.y;
          }
        }
      }
      if (q == 0) {  // packed (DC, Nyquist) bin: component-wise product
        acc[t].x = acc0.x;
        acc[t].y = acc0.y;
      }
      if (S > 1) sP[q] = acc[t];
    }
    }
    if (S > 1) {
      cluster.sync();  // all partials visible cluster-wide
      if (rank == 0 && worker) {
#pragma unroll
        for (int t = 0; t < TL; t++) {
          for (int r = 1; r < S; r++) {
            const float4 o = cluster.map_shared_rank(sP, r)[t * NT + tid];
            acc[t].x += o.x;
            acc[t].y += o.y;
            acc[t].z += o.z;
            acc[t].w += o.w;
          }
        }
      }
      cluster.sync();  // remote reads done before anyone exits
    }
    if (rank == 0) {
      __syncthreads();  // the whole CTA is done with the new frames' spectra: Y takes their place in sX
      if (worker) {
#pragma unroll
        for (int t = 0; t < TL; t++) {
          const int q = t * NT + tid;
          sX[pad_idx(2 * q)] = make_float2(acc[t].x, acc[t].y);
          sX[pad_idx(2 * q + 1)] = make_float2(acc[t].z, acc[t].w);
        }
      }
    }
  } else
  for (int tile = 0; tile < P::TILES; tile++) {
    const int q = tile * NT + tid;  // float4 index inside a frame: bins 2q, 2q+1
    const bool owns = worker && q < HALF;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 acc0 = make_float2(0.f, 0.f);
    if (TMA) {
      // every thread walks the same sequence of partitions (ascending p, the frames written by this launch
      // left out), so `slot` advances identically in the producer and in the consumers
      constexpr uint32_t kSliceBytes = P::SLICE * (uint32_t)sizeof(float4);
      const unsigned char *Fb = reinterpret_cast<const unsigned char *>(reinterpret_cast<const float4 *>(fdl) + tile * NT);
      const unsigned char *Gb = reinterpret_cast<const unsigned char *>(reinterpret_cast<const float4 *>(irs) + tile * NT);
      const size_t frame_bytes = (size_t)PTS * sizeof(float2);
      for (int p = p_lo; p < p_hi; p++) {
        if (p == p_newx || p == p_newg) continue;
        const uint32_t s = slot % STAGES, round = slot / STAGES;
        slot++;
        if (!worker) {
          if ((tid & 31) == 0) {
            if (round > 0) tma::mbar_wait(bar_empty + 8 * s, (round - 1) & 1);  // consumers released the stage
            const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
            tma::mbar_expect_tx(bar_full + 8 * s, 2 * kSliceBytes);
            tma::bulk_g2s(tma::smem_u32(ring + (2 * s) * P::SLICE), Fb + (size_t)frame * frame_bytes, kSliceBytes, bar_full + 8 * s);
            tma::bulk_g2s(tma::smem_u32(ring + (2 * s + 1) * P::SLICE), Gb + (size_t)p * frame_bytes, kSliceBytes, bar_full + 8 * s);
          }
        } else {
          tma::mbar_wait(bar_full + 8 * s, round & 1);
          if (owns) {
            const float4 fa = ring[(2 * s) * P::SLICE + tid], gb = ring[(2 * s + 1) * P::SLICE + tid];
            cmac2(acc, fa, gb);
            acc0.x += fa.x * gb.x;
            acc0.y += fa.y * gb.y;
          }
          __syncwarp();
          if ((tid & 31) == 0) tma::mbar_arrive(bar_empty + 8 * s);
        }
      }
      __syncwarp();  // the producer warp's lane 0 rejoins its warp before the next (aligned) barrier
    }
    if (owns) {
      const float4 *F = reinterpret_cast<const float4 *>(fdl) + q;
      const float4 *Gp = reinterpret_cast<const float4 *>(irs) + q;
      // walk [p_lo, p_hi) in ascending p, cutting at the FDL wrap point and around the new frames
      int p = TMA ? p_hi : p_lo;
      while (p < p_hi) {
        if (p == p_newx || p == p_newg) {
          p++;
          continue;
        }
        int end = p_hi;
        if (p_newx > p && p_newx < end) end = p_newx;
        if (p_newg > p && p_newg < end) end = p_newg;
        const int wrap = nparts - rp;  // first p whose FDL frame index wraps to 0
        if (wrap > p && wrap < end) end = wrap;
        const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
        mac_segment<8>(acc, acc0, F + (size_t)frame * stride4, Gp + (size_t)p * stride4, end - p, stride4);
        p = end;
      }
      if (rank == 0) {
        // the terms that involve frames produced by this launch, from shared memory
        const int b0 = 2 * q;
        float2 x0 = sX[pad_idx(b0)], x1 = sX[pad_idx(b0 + 1)];
        float4 xn = make_float4(x0.x, x0.y, x1.x, x1.y);
        float4 gn;
        if (TV && p_newg == p_newx) {
          float2 g0 = sG[pad_idx(b0)], g1 = sG[pad_idx(b0 + 1)];
          gn = make_float4(g0.x, g0.y, g1.x, g1.y);
        } else {
          gn = __ldcs(Gp + (size_t)p_newx * stride4);
        }
        cmac2(acc, xn, gn);
        acc0.x += xn.x * gn.x;
        acc0.y += xn.y * gn.y;
        if (TV && p_newg != p_newx) {
          const int frame = (rp + p_newg < nparts) ? rp + p_newg : rp + p_newg - nparts;
          float4 fo = __ldcs(F + (size_t)frame * stride4);
          float2 g0 = sG[pad_idx(b0)], g1 = sG[pad_idx(b0 + 1)];
          float4 gg = make_float4(g0.x, g0.y, g1.x, g1.y);
          cmac2(acc, fo, gg);
          acc0.x += fo.x * gg.x;
          acc0.y += fo.y * gg.y;
        }
      }
      if (q == 0) {  // packed (DC, Nyquist) bin: component-wise product
        acc.x = acc0.x;
        acc.y = acc0.y;
      }
      if (S > 1) sP[q] = acc;
    }
    if (S > 1) {
      cluster.sync();  // all partials of this tile visible cluster-wide
      if (rank == 0 && owns) {
        for (int r = 1; r < S; r++) {
          const float4 *remote = cluster.map_shared_rank(sP, r);
          float4 o = remote[q];
          acc.x += o.x;
          acc.y += o.y;
          acc.z += o.z;
          acc.w += o.w;
        }
      }
      cluster.sync();  // remote reads done before anyone overwrites sP (next tile) or exits
    }
    if (rank == 0) {
      if (P::TILES == 1) __syncthreads();  // every thread has consumed sX/sG before Y overwrites sX
      // (TILES > 1: Y cannot go to sX yet; it is parked in sP-like scratch below)
      if (owns) {
        if (P::TILES == 1) {
          sX[pad_idx(2 * q)] = make_float2(acc.x, acc.y);
          sX[pad_idx(2 * q + 1)] = make_float2(acc.z, acc.w);
        } else {
          // park in the partial buffer area past the first HALF entries (sized for it by the host)
          sPark[q] = acc;
        }
      }
    }
  }
  if (rank != 0) return;
  __syncthreads();
  if (P::TILES > 1 && !PMAJOR) {
    for (int q = tid; q < HALF && worker; q += NT) {
      float4 y = sPark[q];
      sX[pad_idx(2 * q)] = make_float2(y.x, y.y);
      sX[pad_idx(2 * q + 1)] = make_float2(y.z, y.w);
    }
    __syncthreads();
  }

  // ---- 3. unsplit (cl_conv_kernels.h:87-100), inverse FFT, overlap-add (120-124) ------------------
  for (int i = tid; i < PTS / 2 && worker; i += NT) {
    if (i == 0) {
      sX[pad_idx(0)] = rfft_dc<true>(sX[pad_idx(0)]);
    } else {
      float2 ci = sX[pad_idx(i)], cj = sX[pad_idx(PTS - i)];
      rfft_pair<true>(ci, cj, __ldg(&a.w2[i]));
      sX[pad_idx(i)] = ci;
      sX[pad_idx(PTS - i)] = cj;
    }
  }
  __syncthreads();
  if (tid < P::FT) {
    const int vt = tid / P::T, t = tid % P::T;
    float2 *my = sX + vt * FftGeom<LOGP>::SMEM;
    auto load = [&](int idx, int) { return my[pad_idx(idx)]; };
    auto store = [&](int idx, float2 v, int) { my[pad_idx(idx)] = v; };
    fft_run<LOGP, true, true, true>(load, store, my, a.tw, t, FftGroupSync<P::FT>());
  }
  __syncthreads();
  // element m of the inverse transform holds reals (y[2m], y[2m+1]); y[0,pts) + old tail -> out,
  // y[pts, 2pts) -> new tail
  const float inv = 1.0f / (float)PTS;
  float2 *out2 = reinterpret_cast<float2 *>(a.out + (size_t)ch * PTS);
  float2 *tail2 = reinterpret_cast<float2 *>(a.tail + (size_t)ch * PTS);
  for (int m = tid; m < PTS / 2 && worker; m += NT) {
    float2 y = sX[pad_idx(m)], z = sX[pad_idx(m + PTS / 2)], tl = tail2[m];
    out2[m] = make_float2((y.x + tl.x) * inv, (y.y + tl.y) * inv);
    tail2[m] = z;
  }
}

// IR partition transform (Clpconv::push_ir, cl_conv.cpp:353-388), all partitions of all channels in
// one launch. A frame is carried by the FT threads of its FFT (one warp for pts <= 512); a CTA holds GROUPS such
// groups (256 threads in all), each with its own frame, shared-memory buffer and NAMED barrier, so the groups never
// wait for each other. History, 1024 x 937 partitions of 512: step-kernel-sized CTAs in which only the FFT group
// works 4.0 ms; one 32-thread CTA per frame 1.53 ms (3.9 TB/s); eight frames per 256-thread CTA, this kernel: 1.53 ms
// again -- not the CTA start rate but the four shared-memory passes over every frame (transform, split in place,
// copy out) bound it. pts >= 64 therefore runs pconv_push_ir_reg_kernel below (1.22 ms, 4.8 TB/s; pts 128: 0.83 ->
// 0.26 ms); this kernel serves pts < 64 and the `pconv_push_reg` = 0 comparison. Partition i goes to IR frame
// (wp2 - i) mod nparts. The arithmetic is that of pconv_forward_frame, instruction for instruction.
template <int LOGP>
struct PushGeom {
  using P = PconvGeom<LOGP>;
  static constexpr int GROUPS = P::FT >= 256 ? 1 : 256 / P::FT;
  static constexpr int THREADS = GROUPS * P::FT;
  static constexpr int GROUP_F2 = P::FFT_SMEM + (P::FFT_SMEM & 1) + 2;  // float2 per group buffer (16-byte multiple)
  static constexpr int SMEM_BYTES = GROUPS * GROUP_F2 * (int)sizeof(float2);
};
// barrier `id` (1..15) over COUNT threads
template <int COUNT>
struct NamedGroupSync {
  int id;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }
};
template <int LOGP>
__global__ void __launch_bounds__(PushGeom<LOGP>::THREADS)
    pconv_push_ir_kernel(const float *ir, size_t ir_stride, float2 *irs, const float2 *__restrict__ tw,
                         const float2 *__restrict__ w2, int nparts, int wp2) {
  using P = PconvGeom<LOGP>;
  using Q = PushGeom<LOGP>;
  constexpr int PTS = P::PTS, N = PTS, FT = P::FT;
  extern __shared__ float4 smem4[];
  const int grp = threadIdx.x / FT, gt = threadIdx.x % FT;
  const int i = blockIdx.x * Q::GROUPS + grp, ch = blockIdx.y;
  if (i >= nparts) return;  // whole groups leave: nobody else waits on their barrier
  float2 *sm = reinterpret_cast<float2 *>(smem4) + grp * Q::GROUP_F2;
  const NamedGroupSync<FT> sync{grp + 1};
  const float *x = ir + (size_t)ch * ir_stride + (size_t)i * PTS;
  {
    const int vt = gt / P::T, t = gt % P::T;
    float2 *my = sm + vt * FftGeom<LOGP>::SMEM;
    const bool real = (vt == 0);
    // an IR pushed with an odd channel stride (push_ir_dev) leaves rows that are only 4-byte aligned
    const bool pair_ok = (reinterpret_cast<uintptr_t>(x) & 7) == 0;
    auto load = [&](int idx, int) {
      if (real && idx < N / 2)
        return pair_ok ? *reinterpret_cast<const float2 *>(x + 2 * idx) : make_float2(x[2 * idx], x[2 * idx + 1]);
      return make_float2(0.f, 0.f);
    };
    auto store = [&](int idx, float2 v, int) { my[pad_idx(idx)] = v; };
    fft_run<LOGP, false, true>(load, store, my, tw, t, sync);
  }
  sync();
  for (int k = gt; k < N / 2; k += FT) {
    if (k == 0) {
      sm[pad_idx(0)] = rfft_dc<false>(sm[pad_idx(0)]);
    } else {
      float2 ci = sm[pad_idx(k)], cj = sm[pad_idx(N - k)];
      rfft_pair<false>(ci, cj, __ldg(&w2[k]));
      sm[pad_idx(k)] = ci;
      sm[pad_idx(N - k)] = cj;
    }
  }
  sync();
  int frame = (wp2 - i) % nparts;
  if (frame < 0) frame += nparts;
  float2 *g = irs + ((size_t)ch * nparts + frame) * PTS;
  for (int k = gt; k < PTS; k += FT) g[k] = sm[pad_idx(k)];
}


// pts >= 64: the same frames on the register-level real transform of the batched FFT (fft_kernels.cuh: split in
// registers with the folded table, one shared-memory exchange instead of four passes over the frame), reading the
// pts reals of a partition as the first half of a zero-padded 2 pts-point input and writing to the partition's ring
// frame. grid = (ceil(nparts / TPB), channels). hw: folded split table of the pts-point plan, scale 1.
template <int LOGP>
__global__ void __launch_bounds__(BatchGeom<LOGP>::THREADS, BatchGeom<LOGP>::MIN_BLOCKS)
    pconv_push_ir_reg_kernel(const float *ir, size_t ir_stride, float2 *irs, const float2 *__restrict__ tw,
                             const float2 *__restrict__ hw, int nparts, int wp2) {
  using B = BatchGeom<LOGP>;
  constexpr int PTS = 1 << LOGP, T = B::T;
  extern __shared__ float2 smem_push[];
  const int lt = threadIdx.x / T, t = threadIdx.x % T;
  const int i = blockIdx.x * B::TPB + lt, ch = blockIdx.y;
  const bool active = i < nparts;
  const float *x = ir + (size_t)ch * ir_stride + (size_t)(active ? i : 0) * PTS;
  int frame = (wp2 - (active ? i : 0)) % nparts;
  if (frame < 0) frame += nparts;
  rfft_fwd_reg_body<LOGP, true>(reinterpret_cast<const float2 *>(x), irs + ((size_t)ch * nparts + frame) * PTS, active,
                                (reinterpret_cast<uintptr_t>(x) & 7) == 0, smem_push + lt * B::ROW, tw, hw, t, 1.0f);
}

// General path, pts = 8192 / 16384 (below): the block's new frames -- R(in1) into FDL frame state[0] and, time-varying,
// R(in2) into IR frame state[1] -- in ONE launch on the same register-level transform (instead of pad, batched rFFT
// and frame copy, three launches per input). grid = (channels, 1 or 2).
template <int LOGP>
__global__ void __launch_bounds__(BatchGeom<LOGP>::THREADS, BatchGeom<LOGP>::MIN_BLOCKS)
    pconv_frames_reg_kernel(const float *in1, const float *in2, size_t in_stride, float2 *fdl, float2 *irs,
                            const float2 *__restrict__ tw, const float2 *__restrict__ hw, int nparts, const int *state) {
  static_assert(BatchGeom<LOGP>::TPB == 1, "one frame per CTA");
  constexpr int PTS = 1 << LOGP;
  extern __shared__ float2 smem_push[];
  const int ch = blockIdx.x, which = blockIdx.y;
  const float *x = (which ? in2 : in1) + (size_t)ch * in_stride;
  float2 *ring = which ? irs : fdl;
  rfft_fwd_reg_body<LOGP, true>(reinterpret_cast<const float2 *>(x), ring + ((size_t)ch * nparts + state[which]) * PTS, true,
                                (reinterpret_cast<uintptr_t>(x) & 7) == 0, smem_push, tw, hw, threadIdx.x, 1.0f);
}
// ... and the block's last three launches in one: inverse real transform of Y, overlap-add with the saved tail
// (cl_conv_kernels.h:120-124), ring positions advanced (cl_conv.cpp:424, 519). grid = channels.
template <int LOGP>
__global__ void __launch_bounds__(BatchGeom<LOGP>::THREADS, BatchGeom<LOGP>::MIN_BLOCKS)
    pconv_inverse_ola_kernel(const float2 *Y, float *tail, float *out, const float2 *__restrict__ tw,
                             const float2 *__restrict__ hw, int *state, int nparts, int tv) {
  static_assert(BatchGeom<LOGP>::TPB == 1, "one transform per CTA");
  constexpr int PTS = 1 << LOGP;
  extern __shared__ float2 smem_push[];
  const int ch = blockIdx.x;
  rfft_inv_reg_body<LOGP, true>(Y + (size_t)ch * PTS, reinterpret_cast<float2 *>(out + (size_t)ch * PTS), true, smem_push, tw,
                                hw, threadIdx.x, reinterpret_cast<float2 *>(tail + (size_t)ch * PTS), 1.0f / (float)PTS);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // nobody in this launch reads the positions
    const int wp = state[0], wp2 = state[1];
    state[0] = wp + 1 == nparts ? 0 : wp + 1;
    if (tv) state[1] = wp2 == 0 ? nparts - 1 : wp2 - 1;
  }
}

// =====================================================================================================
// General path for partitions too long for the fused kernel (pts = 8192 .. 32768, the upper half of the
// partition sizes swept by the reference's csound/tests.py:10). The FFTs run on the batched real-FFT
// plans (unscaled forward), the three kernels below do the rest. Per block: pad -> rFFT -> frame copy ->
// MAC -> inverse rFFT -> overlap-add; the MAC still moves all but a few percent of the bytes.
// =====================================================================================================

// The ring positions of the general path live in device memory -- state[0] = wp (frame the next input block is
// written to, cl_conv.cpp:144,424), state[1] = wp2 (frame the next time-varying IR block is written to, 385,519) --
// so that a block's launch sequence has no host-side parameter that changes from block to block and can be replayed
// as ONE CUDA graph (the host keeps a mirror for push_ir and the white-box reads).
// frame the MAC starts reading at: the reference increments wp before convol (cl_conv.cpp:424-429)
__device__ __forceinline__ int pconv_read_pos(const int *state, int nparts) {
  const int wp = state[0];
  return wp + 1 == nparts ? 0 : wp + 1;
}
// time-varying block: both inputs through ONE batched transform of 2 x channels rows (one single-transform latency
// instead of two: 19 us each at pts 32768). Row y < channels is in1's channel y, row channels + y is in2's.
__global__ void pconv_pad2_kernel(const float *in1, const float *in2, size_t in_stride, float *pad, int pts, int channels) {
  const int row = blockIdx.y, ch = row < channels ? row : row - channels;
  const float *in = row < channels ? in1 : in2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * pts) pad[(size_t)row * 2 * pts + i] = i < pts ? in[(size_t)ch * in_stride + i] : 0.f;
}
// Y [2 x channels][pts] -> FDL frame state[0] (rows < channels) and IR frame state[1] (the others)
__global__ void pconv_ring_store2_kernel(const float2 *Y, float2 *fdl, float2 *irs, int pts, int nparts, const int *state,
                                         int channels) {
  const int row = blockIdx.y, which = row < channels ? 0 : 1, ch = which ? row - channels : row;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  if (i < pts / 2)
    reinterpret_cast<float4 *>((which ? irs : fdl) + ((size_t)ch * nparts + state[which]) * pts)[i] =
        reinterpret_cast<const float4 *>(Y + (size_t)row * pts)[i];
}
// Y [channels][pts] -> ring frame state[which] of every channel
__global__ void pconv_ring_store_kernel(const float2 *Y, float2 *ring, int pts, int nparts, const int *state, int which) {
  const int ch = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  if (i < pts / 2)
    reinterpret_cast<float4 *>(ring + ((size_t)ch * nparts + state[which]) * pts)[i] =
        reinterpret_cast<const float4 *>(Y + (size_t)ch * pts)[i];
}
// end of a block: wp advances (cl_conv.cpp:424), wp2 retreats when the block was time-varying (519)
__global__ void pconv_advance_kernel(int *state, int nparts, int tv) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int wp = state[0], wp2 = state[1];
    state[0] = wp + 1 == nparts ? 0 : wp + 1;
    if (tv) state[1] = wp2 == 0 ? nparts - 1 : wp2 - 1;
  }
}

// in [channels][pts] -> pad [channels][2*pts], upper half zero (cl_conv.cpp:399: half of in1 is written)
__global__ void pconv_pad_kernel(const float *in, size_t in_stride, float *pad, int pts) {
  const int ch = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * pts) pad[(size_t)ch * 2 * pts + i] = i < pts ? in[(size_t)ch * in_stride + i] : 0.f;
}

// Y[ch][n] = sum_p FDL[(rp+p) mod nparts][n] (*) IR[p][n], ascending p; bin 0 component-wise
// (cl_conv_kernels.h:102-118). grid = (pts / 512, channels), 256 threads, two bins per thread.
__global__ void __launch_bounds__(256)
    pconv_mac_kernel(const float2 *fdl, const float2 *irs, float2 *Y, int pts, int nparts, const int *state) {
  const int rp = pconv_read_pos(state, nparts);
  const int ch = blockIdx.y;
  const int q = blockIdx.x * 256 + threadIdx.x;  // float4 index inside a frame
  const size_t stride4 = pts / 2;
  if (q >= (int)stride4) return;
  const size_t chan4 = (size_t)ch * nparts * stride4;
  const float4 *F = reinterpret_cast<const float4 *>(fdl) + chan4 + q;
  const float4 *G = reinterpret_cast<const float4 *>(irs) + chan4 + q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 acc0 = make_float2(0.f, 0.f);
  // this CTA's share of the partitions (grid.z splits them when there are too few CTAs otherwise; see pconv_mac_sum_kernel)
  const int p0 = (int)((long long)blockIdx.z * nparts / gridDim.z), p1 = (int)((long long)(blockIdx.z + 1) * nparts / gridDim.z);
  const int wrap = nparts - rp;  // partitions [0, wrap) read frames rp.., [wrap, nparts) read frames 0..
  const int m = wrap < p0 ? p0 : (wrap > p1 ? p1 : wrap);
  mac_segment<8>(acc, acc0, F + (size_t)(rp + p0) * stride4, G + (size_t)p0 * stride4, m - p0, stride4);
  mac_segment<8>(acc, acc0, F + (size_t)(m - wrap) * stride4, G + (size_t)m * stride4, p1 - m, stride4);
  if (q == 0) {
    acc.x = acc0.x;
    acc.y = acc0.y;
  }
  reinterpret_cast<float4 *>(Y)[((size_t)blockIdx.z * gridDim.y + ch) * stride4 + q] = acc;
}
// Y[ch][n] = sum_k part[k][ch][n], ascending k: the second half of a MAC whose partitions were split over grid.z
// (few channels, long IR: pts / 512 x channels CTAs streaming all partitions one after the other leave most of the GPU
// idle -- mono, pts 8192 x 512 partitions: 16 CTAs). total4 = channels * pts / 2 float4 per part.
__global__ void pconv_mac_sum_kernel(const float4 *part, float4 *Y, size_t total4, int K) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 a = part[i];
  for (int k = 1; k < K; k++) {
    const float4 b = part[(size_t)k * total4 + i];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
  }
  Y[i] = a;
}

// ---- the same MAC fed by the TMA engine -----------------------------------------------------------------------
// One producer thread issues `cp.async.bulk` copies (UBLKCP: global -> shared, completion counted on an mbarrier)
// of the 4 KB FDL slice and the 4 KB IR slice of partition p into an 8-stage ring; the 8 consumer warps wait on
// the stage's "full" barrier, multiply-accumulate out of shared memory and release the stage through its "empty"
// barrier. No registers are tied up by loads in flight (64 KB per CTA sits in the ring), address generation is
// off the SM's issue slots.

constexpr int kMacStages = 8;
constexpr int kMacTileBins = 512;                                  // bins per CTA: 4 KB per frame slice
constexpr int kMacSliceBytes = kMacTileBins * (int)sizeof(float2);
constexpr int kMacTmaSmem = kMacStages * 2 * kMacSliceBytes + 1024;  // ring + barriers (+ alignment slack)

// grid = (pts / 512, channels), 288 threads: warps 0-7 consume, warp 8 lane 0 produces.
__global__ void __launch_bounds__(288)
    pconv_mac_tma_kernel(const float2 *fdl, const float2 *irs, float2 *Y, int pts, int nparts, const int *state) {
  const int rp = pconv_read_pos(state, nparts);
  extern __shared__ __align__(128) unsigned char mac_smem[];
  float4 *ringF = reinterpret_cast<float4 *>(mac_smem);
  float4 *ringG = reinterpret_cast<float4 *>(mac_smem + kMacStages * kMacSliceBytes);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(mac_smem + 2 * kMacStages * kMacSliceBytes);
  const uint32_t full0 = tma::smem_u32(bars), empty0 = tma::smem_u32(bars + kMacStages);
  const int ch = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t frame_bytes = (size_t)pts * sizeof(float2);
  const unsigned char *Fb = reinterpret_cast<const unsigned char *>(fdl) + (size_t)ch * nparts * frame_bytes +
                            (size_t)tile * kMacSliceBytes;
  const unsigned char *Gb = reinterpret_cast<const unsigned char *>(irs) + (size_t)ch * nparts * frame_bytes +
                            (size_t)tile * kMacSliceBytes;
  if (tid == 0) {
    for (int s = 0; s < kMacStages; s++) {
      tma::mbar_init(full0 + 8 * s, 1);   // one arrive (the producer's expect_tx) + the bytes
      tma::mbar_init(empty0 + 8 * s, 8);  // one arrive per consumer warp
    }
    tma::fence_barrier_init();
  }
  __syncthreads();
  // this CTA's share of the partitions (grid.z; see pconv_mac_sum_kernel)
  const int p0 = (int)((long long)blockIdx.z * nparts / gridDim.z), p1 = (int)((long long)(blockIdx.z + 1) * nparts / gridDim.z);
  if (warp == 8) {
    if ((tid & 31) == 0) {
      for (int p = p0; p < p1; p++) {
        const int s = (p - p0) % kMacStages, round = (p - p0) / kMacStages;
        if (round > 0) tma::mbar_wait(empty0 + 8 * s, (round - 1) & 1);  // consumers released the stage
        const int frame = (rp + p < nparts) ? rp + p : rp + p - nparts;
        tma::mbar_expect_tx(full0 + 8 * s, 2 * kMacSliceBytes);
        tma::bulk_g2s(tma::smem_u32(ringF) + s * kMacSliceBytes, Fb + (size_t)frame * frame_bytes, kMacSliceBytes, full0 + 8 * s);
        tma::bulk_g2s(tma::smem_u32(ringG) + s * kMacSliceBytes, Gb + (size_t)p * frame_bytes, kMacSliceBytes, full0 + 8 * s);
      }
    }
    return;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 acc0 = make_float2(0.f, 0.f);
  for (int p = p0; p < p1; p++) {
    const int s = (p - p0) % kMacStages, round = (p - p0) / kMacStages;
    tma::mbar_wait(full0 + 8 * s, round & 1);
    const float4 a = ringF[s * (kMacSliceBytes / 16) + tid], b = ringG[s * (kMacSliceBytes / 16) + tid];
    cmac2(acc, a, b);
    acc0.x += a.x * b.x;
    acc0.y += a.y * b.y;
    __syncwarp();
    if ((tid & 31) == 0) tma::mbar_arrive(empty0 + 8 * s);
  }
  if (tile == 0 && tid == 0) {  // packed (DC, Nyquist) bin: component-wise product
    acc.x = acc0.x;
    acc.y = acc0.y;
  }
  reinterpret_cast<float4 *>(Y)[((size_t)blockIdx.z * gridDim.y + ch) * (pts / 2) + (size_t)tile * 256 + tid] = acc;
}

// y [channels][2*pts] (inverse transform, unnormalised) -> out = (y[0,pts) + tail) / pts, tail = y[pts, 2pts)
// (cl_conv_kernels.h:120-124)
__global__ void pconv_ola_kernel(const float *y, float *tail, float *out, int pts) {
  const int ch = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < pts) {
    const float inv = 1.0f / (float)pts;
    const size_t o = (size_t)ch * pts + n;
    out[o] = (y[(size_t)ch * 2 * pts + n] + tail[o]) * inv;
    tail[o] = y[(size_t)ch * 2 * pts + pts + n];
  }
}

}  // namespace b2f
