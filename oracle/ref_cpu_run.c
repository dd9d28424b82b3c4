/*
 * oracle/ref_cpu_run.c -- timed multi-channel loops over the CPU restatement (ref_cpu.c).
 * TEST INFRASTRUCTURE ONLY. Used by bench.py as the "port" CPU baseline when the compiled
 * reference (oracle/_ref/libclfft_ref.so) is not available. One oracle object per channel,
 * one channel per OpenMP thread, mirroring how the reference would be deployed
 * (one Clpconv / Cldconv / Clrfft object per channel: SURVEY.md section 8e).
 */
#include <omp.h>
#include <stdlib.h>
#include <time.h>

#include "ref_cpu.h"

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double orc_pconv_run(int channels, int cvs, int pts, int nblocks, const float *ir, const float *in, float *out,
                     int threads) {
  orc_pconv **obj = (orc_pconv **)malloc(sizeof(orc_pconv *) * channels);
  for (int c = 0; c < channels; c++) {
    obj[c] = orc_pconv_create(cvs, pts);
    orc_pconv_push_ir(obj[c], ir + (size_t)c * cvs);
  }
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int c = 0; c < channels; c++)
    for (int b = 0; b < nblocks; b++) {
      size_t o = ((size_t)c * nblocks + b) * pts;
      orc_pconv_convolution(obj[c], out + o, in + o);
    }
  double t1 = now_s();
  for (int c = 0; c < channels; c++) orc_pconv_destroy(obj[c]);
  free(obj);
  return t1 - t0;
}

double orc_rfft_run(int size, int batch, float *data, int fwd, int threads) {
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int i = 0; i < batch; i++) orc_rfft(data + (size_t)i * size, size, fwd);
  return now_s() - t0;
}

double orc_cfft_run(int N, int batch, float *data, int fwd, int threads) {
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int i = 0; i < batch; i++) orc_cfft(data + (size_t)i * 2 * N, N, fwd);
  return now_s() - t0;
}

double orc_dconv_run(int channels, int irsize, int vsize, int nblocks, const float *ir, const float *in,
                     float *out, int threads) {
  orc_dconv **obj = (orc_dconv **)malloc(sizeof(orc_dconv *) * channels);
  for (int c = 0; c < channels; c++) {
    obj[c] = orc_dconv_create(irsize, vsize);
    orc_dconv_push_ir(obj[c], ir + (size_t)c * irsize);
  }
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int c = 0; c < channels; c++)
    for (int b = 0; b < nblocks; b++) {
      size_t o = ((size_t)c * nblocks + b) * vsize;
      orc_dconv_convolution(obj[c], out + o, in + o);
    }
  double t1 = now_s();
  for (int c = 0; c < channels; c++) orc_dconv_destroy(obj[c]);
  free(obj);
  return t1 - t0;
}
