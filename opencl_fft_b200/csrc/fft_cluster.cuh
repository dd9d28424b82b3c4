// fft_cluster.cuh -- single-HBM-pass FFT for transforms that do not fit one CTA, on a thread-block CLUSTER.
//
// N = N1 * 16 complex points (N1 = 1024 or 2048, i.e. the 32768- and 65536-point real FFTs; the latter
// is BASELINE config 5a) are carried by a cluster of 4 CTAs on 4 SMs. With n = 16*n1 + n2 and
// k = k1 + N1*k2:
//   step 1  every CTA owns 4 adjacent columns n2 (32-byte sectors in HBM, the 4 CTAs of the cluster
//           consume the 4 sectors of each 128-byte line at the same time) and runs their N1-point FFTs
//           over n1 in its shared memory with the usual Stockham engine;
//   step 2  the transposition never touches HBM or L2: after a cluster barrier every thread gathers
//           the 16 values T[k1][0..15] of ITS row k1 straight out of the 4 CTAs' shared memories
//           (distributed shared memory, coalesced along k1), multiplies by W_N^(n2*k1), and
//   step 3  finishes with one radix-16 butterfly in registers, which yields X[k1 + N1*k2], k2 = 0..15:
//           16 perfectly coalesced stores per warp.
// For the real transform the split (reference cl_fft.cpp:178-191) is fused in: row ownership is
// mirrored (a CTA owns rows k1 and N1-k1), so both members of every pair (i, N-i) live in the same
// CTA and meet through a shared-memory staging buffer before the single store.
// Per transform: 8N bytes read + 8N bytes written to HBM, nothing else (the 2-kernel four-step path in
// fft_large.cuh moves 2-3x that). Clusters are persistent: one cluster per 4 SMs loops over the batch.
//
// Replaces: Clcfft::fft() / Clrfft::transform forward (cl_fft.cpp:138-151, 272-282) at these sizes.
#pragma once

#include <cooperative_groups.h>

#include "fft_core.cuh"

namespace b2f {

namespace cg = cooperative_groups;

template <int LOG1, int CLUSTER = 4>
struct ClusterGeom {
  static constexpr int N1 = 1 << LOG1, N2 = 16, N = N1 * N2, S = CLUSTER, COLS = N2 / CLUSTER;
  using G1 = FftGeom<LOG1>;
  static constexpr int T1 = G1::T;                 // threads per column transform
  static constexpr int THREADS = COLS * T1;        // == N1 / 4 == rows owned by a CTA
  static constexpr int COLSTRIDE = G1::SMEM + 3;   // float2 per column region; == 4 (mod 16): the
                                                   // columns a half-warp touches land on distinct banks
  static constexpr int CTAS_PER_SM = (CLUSTER == 4) ? 1024 / THREADS : 1;  // caps registers at 64 per thread
  static constexpr int TW_ENTRIES = sched_tw_total(G1::S) + 1;  // pass twiddles kept in shared memory
  static constexpr int SMEM_BYTES = (COLS * COLSTRIDE + TW_ENTRIES) * (int)sizeof(float2);
  static_assert(THREADS == N1 / S, "one row per thread in step 3");
  static_assert(16 * THREADS <= COLS * COLSTRIDE, "split staging fits in the column buffer");
};

// exp(-i pi m / 16), m = 0..7, as (cos, sin)
__device__ __forceinline__ float2 w32_const(int m) {
  constexpr float c[8] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                          0.19509032201612826785f};
  constexpr float s[8] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                          0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                          0.98078528040323044913f};
  return make_float2(c[m], -s[m]);
}

// A note on the cluster barriers below. cg::cluster_group::sync() is barrier.cluster.arrive.release +
// wait.acquire; the release compiles to MEMBAR.ALL.GPU + ERRBAR and therefore also waits for the global
// stores of the previous transform to drain (ncu: 15% of warp time in the first version of this kernel).
// A fence-free variant (arrive.relaxed after a __syncthreads(), since only shared memory is exchanged) was
// tried and is WRONG on B200: with ~70 clusters resident a few transforms per batch came out corrupted
// and runs were not reproducible, for either of the two barriers (tests/test_fft_gpu.py::
// test_cluster_fft_every_transform_of_a_batch catches it). The fully fenced barrier is what is used.

// development aid: per-phase cycle counters (tools/cluster_phase_probe.cu); compiled out normally
#ifdef B2F_PHASE_PROBE
__device__ unsigned long long g_phase_probe[8];
#define B2F_PROBE(i)                                                          \
  do {                                                                        \
    if (blockIdx.x == 0 && tid == 0) {                                        \
      const long long now_ = clock64();                                       \
      atomicAdd(&g_phase_probe[i], (unsigned long long)(now_ - probe_last_)); \
      probe_last_ = now_;                                                     \
    }                                                                         \
  } while (0)
#else
#define B2F_PROBE(i)
#endif

// grid = 4 * clusters, cluster = (4,1,1). twl: [16][N1] (transposed so that a warp reads consecutive rows),
// twl[n2*N1 + k1] = W_N^(n2*k1) (forward sign). w2: split twiddles exp(-i pi i / N) (REAL only).
// scale: 1/N forward, 1 inverse. Two CTAs (of different clusters) share an SM: while one waits on HBM, a
// cluster barrier or distributed shared memory, the other one computes.
template <int LOG1, bool INV, bool REAL, int CLUSTER = 4>
__global__ void __launch_bounds__(ClusterGeom<LOG1, CLUSTER>::THREADS, ClusterGeom<LOG1, CLUSTER>::CTAS_PER_SM)
    fft_cluster_kernel(const float2 *in, float2 *out, const float2 *__restrict__ tw1, const float2 *__restrict__ twl,
                       const float2 *__restrict__ w2, int batch, float scale) {
  using C = ClusterGeom<LOG1, CLUSTER>;
  constexpr int N1 = C::N1, N = C::N, THREADS = C::THREADS, CS = C::COLSTRIDE;
  static_assert(!(REAL && INV), "the fused real path is forward only");
  extern __shared__ float2 sA[];
  float2 *stw = sA + C::COLS * CS;  // the N1-point plan's pass twiddles: L1 is flushed by every cluster
                                    // barrier (CCTL.IVALL), shared memory is not
  for (int i = threadIdx.x; i < C::TW_ENTRIES; i += THREADS) stw[i] = tw1[i];
  __syncthreads();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C::S, ncl = gridDim.x / C::S;
  const int tid = threadIdx.x;

  // step-1 role: 4 adjacent lanes take the 4 columns of this CTA (one 32-byte sector per n1)
  const int c = tid % C::COLS, t = tid / C::COLS;
  const int n2 = rank * C::COLS + c;
  float2 *smc = sA + c * CS;

  // step-3 role: one row k1 per thread
  int k1, ptid = tid;
  if (!REAL) {
    k1 = rank * THREADS + tid;
  } else {
    constexpr int H = THREADS / 2;
    if (tid < H) {
      k1 = rank * H + tid;
      ptid = tid + H;
    } else {
      const int d = rank * H + (tid - H);
      k1 = d == 0 ? N1 / 2 : N1 - d;  // the mirror of row 0 is row 0 itself; its slot hosts row N1/2
      ptid = tid - H;
    }
    if (k1 == 0 || k1 == N1 / 2) ptid = tid;  // self-paired rows
  }
  const bool k1zero = (k1 == 0);
  float2 w2own = make_float2(1.f, 0.f);
  if (REAL) w2own = __ldg(&w2[k1]);
  const float2 *twrow = twl + k1;

#ifdef B2F_PHASE_PROBE
  long long probe_last_ = clock64();
#endif
  for (int b = cid; b < batch; b += ncl) {
    const float2 *src = in + (size_t)b * N;
    float2 *dst = out + (size_t)b * N;

    // ---- step 1: N1-point FFT down this thread's column -------------------------------------------
    auto load = [&](int idx, int) { return __ldcs(src + (size_t)idx * 16 + n2); };
    auto store = [&](int idx, float2 v, int) { smc[pad_idx(idx)] = v; };
    fft_run<LOG1, INV, true, false, true>(load, store, smc, stw, t, CtaSync());
    // inter-step twiddles of this thread's row: issued now so that their L2 latency hides behind the barrier
    float2 tw3[16];
#pragma unroll
    for (int j = 1; j < 16; j++) {
      const float2 w = __ldg(twrow + (size_t)j * N1);
      tw3[j] = make_float2(w.x * scale, (INV ? -w.y : w.y) * scale);  // the 1/N scaling rides on the twiddle
    }
    __syncthreads();              // this CTA's columns are in shared memory (pending st.shared drained)
    B2F_PROBE(0);
    cluster.sync();               // ... and so are everybody else's (release/acquire: required, see above)
    B2F_PROBE(1);

    // ---- step 2: gather row k1 from the four CTAs' shared memories, twiddle --------------------------
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int off = (j % C::COLS) * CS + pad_idx(k1);
      if (j / C::COLS == rank)
        v[j] = sA[off];  // own columns: plain shared-memory load, stays off the cluster network
      else
        v[j] = cluster.map_shared_rank(sA, j / C::COLS)[off];
    }
    v[0] = cscale(v[0], scale);
#pragma unroll
    for (int j = 1; j < 16; j++) v[j] = cmul(v[j], tw3[j]);
    B2F_PROBE(2);
    cluster.sync();  // everyone has read everything: the column buffers may be overwritten
    B2F_PROBE(3);

    // the next transform of this cluster: pull its input into L2 now (one 128-byte line per thread of the
    // cluster, fire and forget), so that step 1's loads find it there instead of queueing on HBM
    if (b + ncl < batch) {
      const float2 *nxt = in + (size_t)(b + ncl) * N + (size_t)(rank * THREADS + tid) * 16;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
    }

    // ---- step 3: radix-16 over n2 in registers -> X[k1 + N1*k2], k2 = 0..15 --------------------------
    dft16<INV>(v);

    if (!REAL) {
#pragma unroll
      for (int j = 0; j < 16; j++) __stcs(dst + k1 + (size_t)j * N1, v[j]);
    } else {
      // real-FFT split. Element (k1, k2 < 8) is the low member i of a pair whose high member N-i is
      // element ((N1-k1) mod N1, 15-k2) [(0, 16-k2) on row 0], held by thread ptid. Every thread parks
      // its upper half, fetches its partner's, evaluates its 8 pairs ONCE and stores both members.
      float2 *st = sA;  // staging [8][THREADS]
#pragma unroll
      for (int j = 8; j < 16; j++) st[(j - 8) * THREADS + tid] = v[j];
      __syncthreads();
      const int krow = (N1 - k1) & (N1 - 1);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (j == 0 && k1zero) {
          __stcs(dst, rfft_dc<false>(v[0]));
          __stcs(dst + N / 2, v[8]);  // element N/2: the reference never touches it
          continue;
        }
        const int pj = k1zero ? 16 - j : 15 - j;
        float2 a = v[j], bb = st[(pj - 8) * THREADS + ptid];
        rfft_pair<false>(a, bb, cmul(w2own, w32_const(j)));
        __stcs(dst + k1 + (size_t)j * N1, a);
        __stcs(dst + krow + (size_t)pj * N1, bb);
      }
      __syncthreads();  // staging is the column buffer of the next transform
    }
    B2F_PROBE(4);
  }
  cluster.sync();  // nobody exits while a peer may still be reading its shared memory
}

}  // namespace b2f
