// fft_plan.h -- compile-time pass schedules shared by host (twiddle-table generation) and device
// (the in-shared-memory Stockham engine in fft_core.cuh).
//
// A transform of N = 2^LOGN complex points is factored into at most four register-resident
// radix-R passes (R in {2,4,8,16}). Every thread owns E = max radix complex values, so one
// transform is carried by T = N/E threads. Radix-16 passes with the float2 shared-memory index
// padded as i + (i >> 4) are bank-conflict free for both the strided scatter of a pass and the
// unit-stride gather of the next one (checked offline for every schedule below).
//
// What it replaces in the reference: the log2(N) single-stage `fft` launches plus the `reorder`
// launch of Clcfft::fft() (cl_fft.cpp:138-151) and of cl_conv.cpp:53-67 -- one launch, one pass
// over HBM, instead of log2(N)+1 launches each re-reading the whole array.
#pragma once

namespace b2f {

constexpr int kMaxSmemLogN = 14;  // largest transform held by one CTA (16384 c64 = 128 KiB + padding)
constexpr int kMaxLogN = 16;      // reference limit: int32 index math overflows above 65536

struct Sched {
  int npass;
  int radix[4];
};

#if defined(__CUDACC__)
#define B2F_HD __host__ __device__
#else
#define B2F_HD
#endif

B2F_HD constexpr Sched sched_for(int logn) {
  switch (logn) {
    case 1: return {1, {2, 1, 1, 1}};
    case 2: return {1, {4, 1, 1, 1}};
    case 3: return {1, {8, 1, 1, 1}};
    case 4: return {1, {16, 1, 1, 1}};
    case 5: return {2, {8, 4, 1, 1}};
    case 6: return {2, {8, 8, 1, 1}};
    case 7: return {2, {8, 16, 1, 1}};
    case 8: return {2, {16, 16, 1, 1}};
    case 9: return {3, {2, 16, 16, 1}};
    case 10: return {3, {4, 16, 16, 1}};
    case 11: return {3, {8, 16, 16, 1}};
    case 12: return {3, {16, 16, 16, 1}};
    case 13: return {4, {2, 16, 16, 16}};
    case 14: return {4, {4, 16, 16, 16}};
    default: return {0, {1, 1, 1, 1}};
  }
}

B2F_HD constexpr int sched_max_radix(Sched s) {
  int m = 1;
  for (int i = 0; i < s.npass; i++) m = s.radix[i] > m ? s.radix[i] : m;
  return m;
}
// product of the radices before pass p (the sub-transform length entering pass p)
B2F_HD constexpr int sched_stride(Sched s, int p) {
  int n = 1;
  for (int i = 0; i < p; i++) n *= s.radix[i];
  return n;
}
// offset (in complex entries) of pass p's twiddle block inside the per-plan pass-twiddle table.
// Pass p (p >= 1) stores (R-1) rows of NS entries: row r-1, column k = W^(r*k), W = exp(-2*pi*i/(NS*R)).
B2F_HD constexpr int sched_tw_offset(Sched s, int p) {
  int off = 0;
  for (int i = 1; i < p; i++) off += (s.radix[i] - 1) * sched_stride(s, i);
  return off;
}
B2F_HD constexpr int sched_tw_total(Sched s) { return sched_tw_offset(s, s.npass); }

// shared-memory index padding (float2 units): one pad element every 16
B2F_HD constexpr int pad_idx(int i) { return i + (i >> 4); }
B2F_HD constexpr int padded_len(int n) { return n + (n >> 4) + 1; }

}  // namespace b2f
