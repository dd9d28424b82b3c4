// dconv_kernels.cuh -- direct (time-domain) convolution as a register-tiled FP32 FMA kernel.
//
// Replaces cl_conv::Cldconv::convolution (cl_dconv.cpp:109-132) and its `convol` kernel (32-43):
// irsize*vsize work-items, each doing ONE multiply and a CAS-loop float atomic add into out[n]
// (4096-way contention per output at the BASELINE shape). Here every thread keeps TN (8 or 16) outputs
// in registers and slides over the taps: TN*8 FMAs per 8-tap step against 4 shared-memory vector loads;
// partial sums meet once through shared memory (and, when the taps are split over a thread-block
// cluster, once through distributed shared memory). No atomics, deterministic.
//
// Semantics: with xl = [last irsize samples of the stream | the new samples],
//   out[t] = sum_{h < irsize} xl[t + h] * g[h],   g[h] = coefs[irsize-1-h]      (== sum_c ir[c] x[t-1-c], Q9)
// which is what the reference's ring arithmetic del[(wp+n+h) mod L] evaluates to whenever its ring
// write is valid (irsize % vsize == 0), extended to any number of consecutive blocks per launch.
//
// Layout: hist [channels][irsize] float (double-buffered), coefs [channels][irsize+vsize] float (the
// reference's coefficient ring, same positions, so the time-varying variant re-records taps exactly
// as cl_dconv.cpp:134-148 does), grev [channels][irsize] float = the reversed taps g the kernel streams,
// in/out [channels][nblocks*vsize].
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>

// Two-lane FP32 (FFMA2 on register pairs, tap broadcast, a second window copy shifted by one sample for the odd
// taps) was built and measured here: 50.6 TFLOP/s against 53.4 for scalar FFMA at config 4. FFMA2 halves the
// instruction count of the inner loop but occupies the issue port like two FFMAs on B200 (tools/pk_probe.cu), so
// nothing is freed for the shared-memory loads. The scalar form is the one kept.

namespace b2f {

namespace cg = cooperative_groups;

#ifndef DCONV_WARPS
#define DCONV_WARPS 8  // measured: 4 warps per CTA (half the cross-warp reduction per FMA) 44.0 TFLOP/s, 8 warps 57.0
#endif
constexpr int kDcWarps = DCONV_WARPS;          // warps per CTA; each takes a slice of the tap chunk
constexpr int kDcThreads = kDcWarps * 32;
constexpr int kDcKC = 1024;                    // taps staged per chunk
constexpr int kDcWarpTaps = kDcKC / kDcWarps;  // taps per warp per chunk (multiple of 8)

template <int TN>
struct DconvGeom {
  static constexpr int TILE = 32 * TN;       // outputs per CTA tile (every warp covers all of them)
  static constexpr int XS = TILE + kDcKC;    // staged input window (logical floats)
  // Shared-memory index of logical float i: 4 pad floats after every TN. A lane owns TN consecutive
  // outputs, so its 128-bit window loads start TN floats from its neighbour's; unpadded that is a 4-way
  // (TN=16) / 2-way (TN=8) bank conflict and the kernel becomes shared-memory bound (measured: 50 instead
  // of 45 TFLOP/s was all TN=16 bought). With the pad the 8 lanes of a quarter-warp hit 8 distinct
  // 16-byte bank groups.
  static constexpr int XS_PADDED = XS + 4 * (XS / TN) + 4;
  __host__ __device__ static constexpr int xi(int i) { return i + 4 * (i / TN); }
};

struct DconvArgs {
  const float *hist_in;  // [channels][irsize]
  float *hist_out;       // [channels][irsize]  (other half of the double buffer)
  const float *grev;     // [channels][irsize]  reversed taps
  const float *in;       // [channels][nout]
  float *out;            // [channels][nout]
  int irsize, nout;
  int vec_ok;            // irsize % 4 == 0 && nout % 4 == 0: 128-bit staging loads are aligned
};

// grid = (S, tiles, channels), cluster = (S,1,1): rank r takes taps [r*irsize/S, (r+1)*irsize/S) (rounded to 8).
template <int TN>
__global__ void __launch_bounds__(kDcThreads) dconv_fir_kernel(DconvArgs a) {
  using G = DconvGeom<TN>;
  // double-buffered staging: chunk c+1 streams in (cp.async, no registers, no waiting) while chunk c is computed
  __shared__ __align__(16) float xs[2][G::XS_PADDED];
  __shared__ __align__(16) float gs[2][kDcKC];
  __shared__ __align__(16) float red[kDcWarps][G::TILE];

  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tile = blockIdx.y, ch = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int irsize = a.irsize, nout = a.nout;
  const float *hist = a.hist_in + (size_t)ch * irsize;
  const float *in = a.in + (size_t)ch * nout;
  const float *grev = a.grev + (size_t)ch * irsize;
  const int t0 = tile * G::TILE;
  const int total = irsize + nout;  // length of xl

  const int k_lo = (int)((long long)rank * irsize / S) & ~7;
  const int k_hi = (rank == S - 1) ? irsize : ((int)((long long)(rank + 1) * irsize / S) & ~7);

  float acc[TN];
#pragma unroll
  for (int i = 0; i < TN; i++) acc[i] = 0.f;

  // ---- stage xl[t0 + k0, +XS) and g[k0, +KC) into buffer b ------------------------------------------------
  auto stage = [&](int b, int k0) {
    if (a.vec_ok) {
      const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(&xs[b][0]), gs_s = (uint32_t)__cvta_generic_to_shared(&gs[b][0]);
      for (int i = tid * 4; i < G::XS; i += kDcThreads * 4) {
        const int xi = t0 + k0 + i;  // multiple of 4; irsize % 4 == 0, so 16 bytes never straddle the seam
        const float *src = xi < irsize ? hist + xi : in + (xi - irsize);
        const int nbytes = xi < total ? 16 : 0;  // 0: zero fill
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs_s + (uint32_t)G::xi(i) * 4u),
                     "l"(nbytes ? src : hist), "r"(nbytes)
                     : "memory");
      }
      for (int i = tid * 4; i < kDcKC; i += kDcThreads * 4) {
        const int k = k0 + i;  // k_hi is a multiple of 4 here: a group of four taps is all inside or all outside
        const int nbytes = k < k_hi ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(gs_s + (uint32_t)i * 4u),
                     "l"(nbytes ? grev + k : grev), "r"(nbytes)
                     : "memory");
      }
    } else {
      for (int i = tid; i < G::XS; i += kDcThreads) {
        const int xi = t0 + k0 + i;
        float v = 0.f;
        if (xi < total) v = xi < irsize ? hist[xi] : in[xi - irsize];
        xs[b][G::xi(i)] = v;
      }
      for (int i = tid; i < kDcKC; i += kDcThreads) {
        const int k = k0 + i;
        gs[b][i] = k < k_hi ? grev[k] : 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stage(0, k_lo);
  int buf = 0;
  for (int k0 = k_lo; k0 < k_hi; k0 += kDcKC, buf ^= 1) {
    if (k0 + kDcKC < k_hi) {
      stage(buf ^ 1, k0 + kDcKC);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // ---- this warp's slice of the chunk: TN outputs per lane, 8 taps per step ----------------------------
    const int kw = warp * kDcWarpTaps;
    if (k0 + kw < k_hi) {  // warp-uniform: skip slices that are all padding
      const int xb = lane * TN + kw;  // logical index of this lane's first window element
      // xb is a multiple of TN, so xi(xb + r) == xi(xb) + xi(r): one base register, every window load below
      // addresses it with a compile-time offset (no index arithmetic in the unrolled loop)
      const float *xwb = &xs[buf][0] + G::xi(xb);
      const float *gwb = &gs[buf][0] + kw;
      float xw[TN + 8];
#pragma unroll
      for (int i = 0; i < TN; i += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(xwb + G::xi(i));
        xw[i] = v.x, xw[i + 1] = v.y, xw[i + 2] = v.z, xw[i + 3] = v.w;
      }
#pragma unroll
      for (int kk = 0; kk < kDcWarpTaps; kk += 8) {  // fully unrolled: the sliding window never moves registers
        const float4 v2 = *reinterpret_cast<const float4 *>(xwb + G::xi(kk + TN)),
                     v3 = *reinterpret_cast<const float4 *>(xwb + G::xi(kk + TN + 4));
        xw[TN] = v2.x, xw[TN + 1] = v2.y, xw[TN + 2] = v2.z, xw[TN + 3] = v2.w;
        xw[TN + 4] = v3.x, xw[TN + 5] = v3.y, xw[TN + 6] = v3.z, xw[TN + 7] = v3.w;
        const float4 g0 = *reinterpret_cast<const float4 *>(gwb + kk), g1 = *reinterpret_cast<const float4 *>(gwb + kk + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int i = 0; i < TN; i++) acc[i] = fmaf(g[j], xw[i + j], acc[i]);
#pragma unroll
        for (int i = 0; i < TN; i++) xw[i] = xw[i + 8];
      }
    }
    __syncthreads();  // buffer `buf` is restaged by the next iteration
  }

  // ---- reduce the warps' tap slices ------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < TN; i++) red[warp][lane * TN + i] = acc[i];
  __syncthreads();
  for (int o = tid; o < G::TILE; o += kDcThreads) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kDcWarps; w++) sum += red[w][o];
    red[0][o] = sum;  // only this thread reads or writes column o from here on
  }
  if (S > 1) {
    cluster.sync();
    if (rank == 0) {
      for (int o = tid; o < G::TILE; o += kDcThreads) {
        float sum = red[0][o];
        for (int r = 1; r < S; r++) sum += cluster.map_shared_rank(&red[0][0], r)[o];
        red[0][o] = sum;
      }
    }
    cluster.sync();
  }
  if (rank == 0)
    for (int o = tid; o < G::TILE; o += kDcThreads)
      if (t0 + o < nout) a.out[(size_t)ch * nout + t0 + o] = red[0][o];

  // history for the next call: the last irsize samples of xl, written to the other buffer
  if (rank == 0 && tile == 0) {
    float *ho = a.hist_out + (size_t)ch * irsize;
    for (int i = tid; i < irsize; i += kDcThreads) {
      const int xi = i + nout;
      ho[i] = xi < irsize ? hist[xi] : in[xi - irsize];
    }
  }
}

// grev[k] = coefs[irsize-1-k] for all channels (after push_ir and after every time-varying ring write)
__global__ void dconv_reverse_kernel(float *grev, const float *coefs, int irsize, int L) {
  const int ch = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < irsize) grev[(size_t)ch * irsize + k] = coefs[(size_t)ch * L + irsize - 1 - k];
}

// ring write used by the time-varying variant (cl_dconv.cpp:134-148): coefs[(wp + i) mod L] = in2[i], and the
// same value into the reversed copy when it lands on a tap position (< irsize)
__global__ void dconv_coef_write_kernel(float *coefs, float *grev, const float *in2, int vsize, int irsize, int L, int wp) {
  const int ch = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < vsize) {
    int pos = wp + i;
    if (pos >= L) pos -= L;
    const float v = in2[(size_t)ch * vsize + i];
    coefs[(size_t)ch * L + pos] = v;
    if (pos < irsize) grev[(size_t)ch * irsize + irsize - 1 - pos] = v;
  }
}

}  // namespace b2f
