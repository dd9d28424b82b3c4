// dconv_kernels.cuh -- direct (time-domain) convolution as a register-tiled FP32 FMA kernel.
//
// Replaces cl_conv::Cldconv::convolution (cl_dconv.cpp:109-132) and its `convol` kernel (32-43):
// irsize*vsize work-items, each doing ONE multiply and a CAS-loop float atomic add into out[n]
// (4096-way contention per output at the BASELINE shape). Here every thread keeps 8 outputs in
// registers and slides over the taps, 64 FMAs per 4 shared-memory vector loads; partial sums meet
// once through shared memory (and, when the taps are split over a thread-block cluster, once
// through distributed shared memory). No atomics, deterministic.
//
// Semantics: with xl = [last irsize samples of the stream | the new samples],
//   out[t] = sum_{h < irsize} xl[t + h] * coefs[irsize-1-h]      (== sum_c ir[c] x[t-1-c], Q9)
// which is what the reference's ring arithmetic del[(wp+n+h) mod L] evaluates to whenever its ring
// write is valid (irsize % vsize == 0), extended to any number of consecutive blocks per launch.
//
// Layout: hist [channels][irsize] float (double-buffered), coefs [channels][irsize+vsize] float (the
// reference's coefficient ring, same positions, so the time-varying variant re-records taps exactly
// as cl_dconv.cpp:134-148 does), in/out [channels][nblocks*vsize].
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace b2f {

namespace cg = cooperative_groups;

constexpr int kDcWarps = 8;                    // warps per CTA; each takes a slice of the tap chunk
constexpr int kDcThreads = kDcWarps * 32;
constexpr int kDcTN = 8;                       // outputs per thread
constexpr int kDcTileOut = 32 * kDcTN;         // outputs per CTA tile (every warp covers all of them)
constexpr int kDcKC = 1024;                    // taps staged per chunk
constexpr int kDcWarpTaps = kDcKC / kDcWarps;  // taps per warp per chunk (multiple of 8)
constexpr int kDcXs = kDcTileOut + kDcKC;      // staged input window

struct DconvArgs {
  const float *hist_in;  // [channels][irsize]
  float *hist_out;       // [channels][irsize]  (other half of the double buffer)
  const float *coefs;    // [channels][irsize + vsize]
  const float *in;       // [channels][nout]
  float *out;            // [channels][nout]
  int irsize, vsize, nout;
  int coef_stride;       // irsize + vsize
};

// grid = (S, tiles, channels), cluster = (S,1,1): rank r takes taps [r*irsize/S, (r+1)*irsize/S) (rounded to 8).
__global__ void __launch_bounds__(kDcThreads) dconv_fir_kernel(DconvArgs a) {
  __shared__ __align__(16) float xs[kDcXs];
  __shared__ __align__(16) float gs[kDcKC];
  __shared__ __align__(16) float red[kDcWarps][kDcTileOut];

  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tile = blockIdx.y, ch = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int irsize = a.irsize, nout = a.nout;
  const float *hist = a.hist_in + (size_t)ch * irsize;
  const float *in = a.in + (size_t)ch * nout;
  const float *coefs = a.coefs + (size_t)ch * a.coef_stride;
  const int t0 = tile * kDcTileOut;
  const int total = irsize + nout;  // length of xl

  int k_lo = (int)((long long)rank * irsize / S) & ~7;
  int k_hi = (rank == S - 1) ? irsize : ((int)((long long)(rank + 1) * irsize / S) & ~7);

  float acc[kDcTN];
#pragma unroll
  for (int i = 0; i < kDcTN; i++) acc[i] = 0.f;

  for (int k0 = k_lo; k0 < k_hi; k0 += kDcKC) {
    // stage xl[t0 + k0, +kDcXs) and the reversed taps g[k0, +kDcKC)
    for (int i = tid; i < kDcXs; i += kDcThreads) {
      const int xi = t0 + k0 + i;
      float v = 0.f;
      if (xi < total) v = xi < irsize ? hist[xi] : in[xi - irsize];
      xs[i] = v;
    }
    for (int i = tid; i < kDcKC; i += kDcThreads) {
      const int k = k0 + i;
      gs[i] = k < k_hi ? coefs[irsize - 1 - k] : 0.f;
    }
    __syncthreads();
    const int kw = warp * kDcWarpTaps;
    if (k0 + kw < k_hi) {  // warp-uniform: skip slices that are all padding
      const float *xp = xs + lane * kDcTN + kw;
      float xw[16];
      {
        float4 v0 = *reinterpret_cast<const float4 *>(xp), v1 = *reinterpret_cast<const float4 *>(xp + 4);
        xw[0] = v0.x, xw[1] = v0.y, xw[2] = v0.z, xw[3] = v0.w, xw[4] = v1.x, xw[5] = v1.y, xw[6] = v1.z, xw[7] = v1.w;
      }
#pragma unroll 2
      for (int kk = 0; kk < kDcWarpTaps; kk += 8) {
        float4 v2 = *reinterpret_cast<const float4 *>(xp + kk + 8), v3 = *reinterpret_cast<const float4 *>(xp + kk + 12);
        xw[8] = v2.x, xw[9] = v2.y, xw[10] = v2.z, xw[11] = v2.w, xw[12] = v3.x, xw[13] = v3.y, xw[14] = v3.z, xw[15] = v3.w;
        float4 g0 = *reinterpret_cast<const float4 *>(gs + kw + kk), g1 = *reinterpret_cast<const float4 *>(gs + kw + kk + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int i = 0; i < kDcTN; i++) acc[i] = fmaf(g[j], xw[i + j], acc[i]);
#pragma unroll
        for (int i = 0; i < 8; i++) xw[i] = xw[i + 8];
      }
    }
    __syncthreads();
  }

  // reduce the warps' tap slices
#pragma unroll
  for (int i = 0; i < kDcTN; i++) red[warp][lane * kDcTN + i] = acc[i];
  __syncthreads();
  float sum = 0.f;  // thread `tid` owns output t0 + tid (kDcThreads == kDcTileOut)
#pragma unroll
  for (int w = 0; w < kDcWarps; w++) sum += red[w][tid];
  if (S > 1) {
    __syncthreads();
    red[0][tid] = sum;
    cluster.sync();
    if (rank == 0) {
      for (int r = 1; r < S; r++) sum += cluster.map_shared_rank(&red[0][0], r)[tid];
    }
    cluster.sync();
  }
  if (rank == 0 && t0 + tid < nout) a.out[(size_t)ch * nout + t0 + tid] = sum;

  // history for the next call: the last irsize samples of xl, written to the other buffer
  if (rank == 0 && tile == 0) {
    float *ho = a.hist_out + (size_t)ch * irsize;
    for (int i = tid; i < irsize; i += kDcThreads) {
      const int xi = i + nout;
      ho[i] = xi < irsize ? hist[xi] : in[xi - irsize];
    }
  }
}
static_assert(kDcThreads == kDcTileOut, "one thread per output in the epilogue");

// ring write used by the time-varying variant (cl_dconv.cpp:134-148): coefs[(wp + i) mod L] = in2[i]
__global__ void dconv_coef_write_kernel(float *coefs, const float *in2, int vsize, int L, int wp) {
  const int ch = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < vsize) {
    int pos = wp + i;
    if (pos >= L) pos -= L;
    coefs[(size_t)ch * L + pos] = in2[(size_t)ch * vsize + i];
  }
}

}  // namespace b2f
