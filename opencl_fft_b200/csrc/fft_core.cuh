// fft_core.cuh -- the in-shared-memory Stockham FFT engine every kernel in this library is built on.
//
// One transform of N = 2^LOGN complex points is carried by T = N/E threads (E = widest radix of the
// schedule in fft_plan.h). Each pass gathers E values per thread at unit stride across threads,
// multiplies by per-pass twiddles read from a table laid out [r][k] (so a warp reads consecutive
// entries), runs a radix-R butterfly entirely in registers and scatters into the autosort position,
// so no bit-reversal pass exists (the reference's `reorder` kernel, cl_fft.cpp:24-27, disappears).
// The first pass reads through a caller-supplied functor (global memory, zero padding, real->complex
// packing ...) and the last pass writes through another one (global memory, 1/N scaling, or shared
// memory when a real-FFT split / overlap-add epilogue follows), which is how load/store fusion is
// expressed without copies of the engine.
//
// Numerics: IEEE float32, twiddles taken from a double-precision-generated table with the same
// formula as the reference (cl_fft.cpp:86-91), radix-internal constants are the float roundings of
// the same cosines. The result differs from the reference's radix-2 chain only by rounding order
// (measured ~1e-7 relative L2; the contract is 1e-5).
#pragma once

#include <cuda_runtime.h>

#include "fft_plan.h"

namespace b2f {

// ---- complex arithmetic on Blackwell's packed-FP32 instructions -------------------------------------------
// sm_100a has two-lane FP32 instructions on 64-bit register pairs (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 /
// FMUL2 / FFMA2). An interleaved complex value IS such a pair, and the SASS operands take free modifiers -- lane
// swap (.LO_HI), per-lane negate (.NP), scalar broadcast (.F32) -- which ptxas folds from the mov.b64 pack /
// unpack below. So a complex add is ONE instruction, a quarter turn folded into the following add costs none, a
// complex multiply is two (FMUL2 + FFMA2): the FP instruction count of a radix-16 pass halves (cfft_kernel<12>:
// 920 -> 608 SASS instructions). Measured effect on B200: none either way (tools/pk_probe.cu: an FFMA2 occupies
// pipe AND issue port like two FFMAs; every FFT size times the same within 1 %), which also shows that these
// kernels are not bound by instruction issue. Kept because the code is smaller; -DB2F_PACKED=0 builds the scalar
// forms.
#ifndef B2F_PACKED
#define B2F_PACKED 1
#endif
// 1: pass twiddles W^(r k) for r not a power of two are derived by multiplication instead of loaded (fft_pass)
#ifndef B2F_TW_DERIVE
#define B2F_TW_DERIVE 1
#endif
typedef unsigned long long b2f_u64;
__device__ __forceinline__ b2f_u64 pk2(float lo, float hi) {
  b2f_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 up2(b2f_u64 v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}
__device__ __forceinline__ b2f_u64 add2(b2f_u64 a, b2f_u64 b) {
  b2f_u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ b2f_u64 sub2(b2f_u64 a, b2f_u64 b) {
  b2f_u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ b2f_u64 mul2(b2f_u64 a, b2f_u64 b) {
  b2f_u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ b2f_u64 fma2(b2f_u64 a, b2f_u64 b, b2f_u64 c) {
  b2f_u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

#if B2F_PACKED
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up2(add2(pk2(a.x, a.y), pk2(b.x, b.y))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up2(sub2(pk2(a.x, a.y), pk2(b.x, b.y))); }
// a * b = b.x * (a.x, a.y) + b.y * (-a.y, a.x)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  const float2 t = up2(mul2(pk2(b.y, b.y), pk2(a.y, a.x)));
  return up2(fma2(pk2(b.x, b.x), pk2(a.x, a.y), pk2(-t.x, t.y)));
}
// s * a for a real s
__device__ __forceinline__ float2 cscale(float2 a, float s) { return up2(mul2(pk2(a.x, a.y), pk2(s, s))); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward quarter turn) or +i (inverse); a lane swap + negate the next packed add absorbs
template <bool INV>
__device__ __forceinline__ float2 cquarter(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by a compile-time unit-circle constant (c, -s) forward / (c, +s) inverse
template <bool INV>
__device__ __forceinline__ float2 cmulc(float2 a, float c, float s) {
#if B2F_PACKED
  const float2 t = up2(mul2(pk2(s, s), pk2(a.y, a.x)));  // (s a.y, s a.x)
  return up2(fma2(pk2(c, c), pk2(a.x, a.y), INV ? pk2(-t.x, t.y) : pk2(t.x, -t.y)));
#else
  return INV ? make_float2(a.x * c - a.y * s, a.x * s + a.y * c) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
#endif
}

#define B2F_SQRT1_2 0.70710678118654752440f
#define B2F_COS_PI_8 0.92387953251128675613f
#define B2F_SIN_PI_8 0.38268343236508977173f

// ---- radix butterflies: natural-order in, natural-order out, in registers ---------------------
template <bool INV>
__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
template <bool INV>
__device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
  float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = cquarter<INV>(csub(v1, v3));
  v0 = cadd(t0, t2);
  v2 = csub(t0, t2);
  v1 = cadd(t1, t3);
  v3 = csub(t1, t3);
}
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  // even / odd radix-4 sub-transforms, then one radix-2 combine with W8^k
  dft4<INV>(v[0], v[2], v[4], v[6]);
  dft4<INV>(v[1], v[3], v[5], v[7]);
  float2 o1 = cmulc<INV>(v[3], B2F_SQRT1_2, B2F_SQRT1_2);
  float2 o2 = cquarter<INV>(v[5]);
  float2 o3 = cmulc<INV>(v[7], -B2F_SQRT1_2, B2F_SQRT1_2);
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1);
  v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2);
  v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3);
  v[7] = csub(e3, o3);
}
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  float2 e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    e[i] = v[2 * i];
    o[i] = v[2 * i + 1];
  }
  dft8<INV>(e);
  dft8<INV>(o);
  o[1] = cmulc<INV>(o[1], B2F_COS_PI_8, B2F_SIN_PI_8);
  o[2] = cmulc<INV>(o[2], B2F_SQRT1_2, B2F_SQRT1_2);
  o[3] = cmulc<INV>(o[3], B2F_SIN_PI_8, B2F_COS_PI_8);
  o[4] = cquarter<INV>(o[4]);
  o[5] = cmulc<INV>(o[5], -B2F_SIN_PI_8, B2F_COS_PI_8);
  o[6] = cmulc<INV>(o[6], -B2F_SQRT1_2, B2F_SQRT1_2);
  o[7] = cmulc<INV>(o[7], -B2F_COS_PI_8, B2F_SIN_PI_8);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[i] = cadd(e[i], o[i]);
    v[i + 8] = csub(e[i], o[i]);
  }
}
// radix 32 (the one-SM kernel of fft_sm.cuh and the one-thread-per-transform kernel of fft_kernels.cuh)
template <bool INV>
__device__ __forceinline__ void dft32(float2 (&v)[32]) {
  constexpr float kC[16] = {1.f,
                            0.98078528040323043f,
                            0.92387953251128674f,
                            0.83146961230254524f,
                            0.70710678118654757f,
                            0.55557023301960229f,
                            0.38268343236508984f,
                            0.19509032201612833f,
                            0.f,
                            -0.19509032201612833f,
                            -0.38268343236508984f,
                            -0.55557023301960229f,
                            -0.70710678118654757f,
                            -0.83146961230254524f,
                            -0.92387953251128674f,
                            -0.98078528040323043f};
  constexpr float kS[16] = {0.f,
                            0.19509032201612825f,
                            0.38268343236508978f,
                            0.55557023301960218f,
                            0.70710678118654757f,
                            0.83146961230254524f,
                            0.92387953251128674f,
                            0.98078528040323043f,
                            1.f,
                            0.98078528040323043f,
                            0.92387953251128674f,
                            0.83146961230254524f,
                            0.70710678118654757f,
                            0.55557023301960218f,
                            0.38268343236508978f,
                            0.19509032201612825f};
  float2 e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    e[i] = v[2 * i];
    o[i] = v[2 * i + 1];
  }
  dft16<INV>(e);
  dft16<INV>(o);
#pragma unroll
  for (int i = 1; i < 16; i++) o[i] = (i == 8) ? cquarter<INV>(o[i]) : cmulc<INV>(o[i], kC[i], kS[i]);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    v[i] = cadd(e[i], o[i]);
    v[i + 16] = csub(e[i], o[i]);
  }
}
template <int R, bool INV>
__device__ __forceinline__ void dftR(float2 (&v)[R]) {
  if constexpr (R == 2) dft2<INV>(v[0], v[1]);
  if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
  if constexpr (R == 8) dft8<INV>(v);
  if constexpr (R == 16) dft16<INV>(v);
  if constexpr (R == 32) dft32<INV>(v);
}

// ---- geometry of one plan ----------------------------------------------------------------------
template <int LOGN>
struct FftGeom {
  static constexpr Sched S = sched_for(LOGN);
  static constexpr int N = 1 << LOGN;
  static constexpr int E = sched_max_radix(S);  // complex values per thread
  static constexpr int T = N / E;               // threads per transform
  static constexpr int SMEM = padded_len(N);    // float2 entries of shared memory per transform
};

// ---- one pass ------------------------------------------------------------------------------------
// SYNC is a functor performing the barrier that covers all threads sharing `sm` (usually
// __syncthreads()). `t` is this thread's index within the transform, in [0, T).
// FIRST_INPLACE / LAST_INPLACE: the load functor reads / the store functor writes the very `sm` the
// engine works in, so the pass needs the gather-before-scatter barrier it otherwise skips.
template <int LOGN, int P, bool INV, bool FIRST_INPLACE, bool LAST_INPLACE, bool TW_SMEM, class Load, class Store,
          class Sync>
__device__ __forceinline__ void fft_pass(Load &load, Store &store, float2 *sm, const float2 *tw, int t,
                                         Sync &sync) {
  using G = FftGeom<LOGN>;
  constexpr Sched S = G::S;
  constexpr int N = G::N, E = G::E, T = G::T;
  constexpr int R = S.radix[P];
  constexpr int NS = sched_stride(S, P);
  constexpr int Q = E / R;  // butterflies per thread in this pass
  constexpr bool FIRST = (P == 0), LAST = (P == S.npass - 1);
  constexpr int TWO = sched_tw_offset(S, P);

  float2 v[Q][R];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int j = t + q * T;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int idx = j + r * (N / R);
      if constexpr (FIRST)
        v[q][r] = load(idx, q * R + r);
      else
        v[q][r] = sm[pad_idx(idx)];
    }
  }
  // every gather done before the in-place scatter below
  if constexpr ((!FIRST && !LAST) || (FIRST && FIRST_INPLACE && !LAST) || (LAST && LAST_INPLACE && !FIRST)) sync();
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int j = t + q * T;
    const int k = j & (NS - 1);
    if constexpr (!FIRST) {
#if B2F_TW_DERIVE
      // Row r of the table holds W^(r k): only the power-of-two rows are LOADED, the others are products of two
      // of them (r = hi + lo, hi the top bit; at most three roundings deep for r = 15, ~2e-7). The loads go
      // through the same LSU data pipe as the shared-memory exchanges, which is the SM-side limit of these
      // kernels (ncu: 76-82 % busy at full HBM rate with all R-1 rows loaded); the FP32 pipe has the headroom.
      float2 wr[R];
#pragma unroll
      for (int r = 1; r < R; r <<= 1) {
        wr[r] = TW_SMEM ? tw[TWO + (r - 1) * NS + k] : __ldg(&tw[TWO + (r - 1) * NS + k]);
        if constexpr (INV) wr[r].y = -wr[r].y;
      }
#pragma unroll
      for (int r = 3; r < R; r++) {
        const int hi = r >= 8 ? 8 : (r >= 4 ? 4 : 2);
        if (r != hi) wr[r] = cmul(wr[hi], wr[r - hi]);
      }
#pragma unroll
      for (int r = 1; r < R; r++) v[q][r] = cmul(v[q][r], wr[r]);
#else
#pragma unroll
      for (int r = 1; r < R; r++) {
        float2 w = TW_SMEM ? tw[TWO + (r - 1) * NS + k] : __ldg(&tw[TWO + (r - 1) * NS + k]);
        if constexpr (INV) w.y = -w.y;
        v[q][r] = cmul(v[q][r], w);
      }
#endif
    }
    dftR<R, INV>(v[q]);
    const int base = (j - k) * R + k;  // (j / NS) * NS * R + k
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int idx = base + r * NS;
      if constexpr (LAST)
        store(idx, v[q][r], q * R + r);
      else
        sm[pad_idx(idx)] = v[q][r];
    }
  }
  if constexpr (!LAST) sync();
}

// ---- whole transform -------------------------------------------------------------------------------
// Load: float2 operator()(int idx, int slot)        -- element idx of the input, idx in [0, N)
// Store: void operator()(int idx, float2 v, int slot) -- element idx of the output (natural order)
// `slot` in [0, E) is a compile-time constant after unrolling: it names which of the thread's E
// values this is, so a functor can keep per-thread constants (e.g. twiddles) in a register array.
// last_pass_index<LOGN>(t, slot) gives the idx a given (t, slot) stores to.
// `sm` needs FftGeom<LOGN>::SMEM float2 (unused for single-pass sizes N <= 16).
// Only threads with t < T may call; all of them must (barriers inside).
// TW_SMEM: `tw` points to a copy of the pass-twiddle table in shared memory (plain loads) instead of
// global memory (read-only path).
template <int LOGN, bool INV, bool LAST_INPLACE = false, bool FIRST_INPLACE = false, bool TW_SMEM = false, class Load,
          class Store, class Sync>
__device__ __forceinline__ void fft_run(Load load, Store store, float2 *sm, const float2 *tw, int t,
                                        Sync sync) {
  constexpr Sched S = sched_for(LOGN);
  fft_pass<LOGN, 0, INV, FIRST_INPLACE, LAST_INPLACE, TW_SMEM>(load, store, sm, tw, t, sync);
  if constexpr (S.npass > 1) fft_pass<LOGN, 1, INV, FIRST_INPLACE, LAST_INPLACE, TW_SMEM>(load, store, sm, tw, t, sync);
  if constexpr (S.npass > 2) fft_pass<LOGN, 2, INV, FIRST_INPLACE, LAST_INPLACE, TW_SMEM>(load, store, sm, tw, t, sync);
  if constexpr (S.npass > 3) fft_pass<LOGN, 3, INV, FIRST_INPLACE, LAST_INPLACE, TW_SMEM>(load, store, sm, tw, t, sync);
}

// input index read by thread t's value `slot` in the first pass (same arithmetic as fft_pass)
template <int LOGN>
__device__ __forceinline__ int first_pass_index(int t, int slot) {
  using G = FftGeom<LOGN>;
  constexpr int R = G::S.radix[0];
  return t + (slot / R) * G::T + (slot % R) * (G::N / R);
}

// output index written by thread t's value `slot` in the last pass (same arithmetic as fft_pass)
template <int LOGN>
__device__ __forceinline__ int last_pass_index(int t, int slot) {
  using G = FftGeom<LOGN>;
  constexpr Sched S = G::S;
  constexpr int P = S.npass - 1;
  constexpr int R = S.radix[P];
  constexpr int NS = sched_stride(S, P);
  const int q = slot / R, r = slot % R;
  const int j = t + q * G::T;
  const int k = j & (NS - 1);
  return (j - k) * R + k + r * NS;
}

struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// ---- real-FFT split / unsplit on one pair (i, N-i) -------------------------------------------------
// Same algebra as the reference's conv/iconv kernels (cl_fft.cpp:178-205; cl_conv_kernels.h:70-100):
//   e = (c[i] + conj c[j]) / 2,  o = +-i (conj c[j] - c[i]) / 2,  c[i] = e + w o,  c[j] = conj(e - w o).
// w = exp(-i pi i/N) for the forward split, its conjugate for the inverse.
template <bool INV>
__device__ __forceinline__ void rfft_pair(float2 &ci, float2 &cj, float2 w) {
  float2 cjc = cconj(cj);
  float2 e = make_float2(.5f * (ci.x + cjc.x), .5f * (ci.y + cjc.y));
  float2 d = INV ? csub(ci, cjc) : csub(cjc, ci);
  float2 o = make_float2(.5f * -d.y, .5f * d.x);
  if (INV) w.y = -w.y;
  float2 p = cmul(w, o);
  ci = cadd(e, p);
  cj = cconj(csub(e, p));
}
// element 0: packed (DC, Nyquist). Forward halves it, inverse does not double it back -- that
// asymmetry is the reference's (cl_fft.cpp:181,195) and is kept.
template <bool INV>
__device__ __forceinline__ float2 rfft_dc(float2 c0) {
  return INV ? make_float2(c0.x + c0.y, c0.x - c0.y) : make_float2((c0.x + c0.y) * .5f, (c0.x - c0.y) * .5f);
}

// the same pair with 0.5, the scaling and the quarter turn folded into the table entry hw = 0.5*scale*i*w
// (forward) or its conjugate (inverse) and hs = 0.5*scale: out_i = hs*S + hw*D, out_j = conj(hs*S - hw*D) with
// S = A + conj(B), D = conj(B) - A. 6 packed instructions per pair (12 scalar).
template <bool INV>
__device__ __forceinline__ void rfft_pair_folded(float2 &A, float2 &B, float2 hw, float hs) {
#if B2F_PACKED
  const float2 Bc = make_float2(B.x, -B.y);
  const float2 S = cadd(A, Bc), D = csub(Bc, A), P = cmul(D, hw);
  A = up2(fma2(pk2(hs, hs), pk2(S.x, S.y), pk2(P.x, P.y)));
  B = up2(fma2(pk2(hs, -hs), pk2(S.x, S.y), pk2(-P.x, P.y)));
#else
  const float sx = A.x + B.x, sy = A.y - B.y;
  const float dx = B.x - A.x, dy = -B.y - A.y;
  const float px = hw.x * dx - hw.y * dy, py = hw.x * dy + hw.y * dx;
  A = make_float2(fmaf(hs, sx, px), fmaf(hs, sy, py));
  B = make_float2(fmaf(hs, sx, -px), fmaf(-hs, sy, py));
#endif
}

}  // namespace b2f
