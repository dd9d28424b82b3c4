/*
 * oracle/ref_capi.cpp -- extern "C" handles onto the UNMODIFIED reference classes.
 * TEST INFRASTRUCTURE ONLY (see oracle/ref_cpu.h). Compiled by oracle/Makefile together with
 * /root/reference/{cl_fft,cl_conv,cl_dconv}.cpp (read in place, never copied) and
 * oracle/minicl (the host-CPU OpenCL runtime) into oracle/_ref/libclfft_ref.so, so that tests
 * and bench.py can drive the real reference through ctypes:
 *   - to validate oracle/ref_cpu.c (bit-exact under minicl's sequential schedule),
 *   - to generate tests/golden/ fixtures (tests/golden/make_golden.py),
 *   - as the "reference" CPU baseline of bench.py (one reference object per channel, one
 *     channel per OpenMP thread -- the reference has no batching or threading of its own).
 */
#include <chrono>
#include <complex>
#include <cstring>
#include <vector>

#include <omp.h>

#include "cl_conv.h"
#include "cl_dconv.h"
#include "cl_fft.h"

static cl_device_id the_device() {
  cl_device_id ids[4];
  cl_uint num = 0;
  clGetDeviceIDs(NULL, CL_DEVICE_TYPE_ALL, 4, ids, &num);
  return ids[0];
}
static void quiet(std::string, void *) {}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

const char *ref_device_name() {
  static char name[128];
  clGetDeviceInfo(the_device(), CL_DEVICE_NAME, sizeof(name), name, NULL);
  return name;
}

/* ---- Clcfft / Clrfft ------------------------------------------------------------------- */
void *ref_cfft_create(int N, int fwd) { return new cl_fft::Clcfft(the_device(), N, fwd != 0); }
int ref_cfft_error(void *h) { return ((cl_fft::Clcfft *)h)->get_error(); }
int ref_cfft_transform(void *h, float *c) {
  return ((cl_fft::Clcfft *)h)->transform(reinterpret_cast<std::complex<float> *>(c));
}
void ref_cfft_destroy(void *h) { delete (cl_fft::Clcfft *)h; }

void *ref_rfft_create(int size, int fwd) { return new cl_fft::Clrfft(the_device(), size, fwd != 0); }
int ref_rfft_error(void *h) { return ((cl_fft::Clrfft *)h)->get_error(); }
/* in-place form (cl_fft.h:105-110) */
int ref_rfft_transform(void *h, float *c) {
  return ((cl_fft::Clrfft *)h)->transform(reinterpret_cast<std::complex<float> *>(c));
}
/* out-of-place form (cl_fft.h:97-103) */
int ref_rfft_transform2(void *h, float *c, float *r) {
  return ((cl_fft::Clrfft *)h)->transform(reinterpret_cast<std::complex<float> *>(c), r);
}
void ref_rfft_destroy(void *h) { delete (cl_fft::Clrfft *)h; }

/* ---- Clpconv ----------------------------------------------------------------------------- */
void *ref_pconv_create(int cvs, int pts) { return new cl_conv::Clpconv(the_device(), cvs, pts, quiet, NULL); }
int ref_pconv_error(void *h) { return ((cl_conv::Clpconv *)h)->get_cl_err(); }
int ref_pconv_push_ir(void *h, float *ir) { return ((cl_conv::Clpconv *)h)->push_ir(ir); }
int ref_pconv_convolution(void *h, float *out, float *in) {
  return ((cl_conv::Clpconv *)h)->convolution(out, in);
}
int ref_pconv_convolution_tv(void *h, float *out, float *in1, float *in2) {
  return ((cl_conv::Clpconv *)h)->convolution(out, in1, in2);
}
void ref_pconv_destroy(void *h) { delete (cl_conv::Clpconv *)h; }

/* ---- Cldconv ----------------------------------------------------------------------------- */
void *ref_dconv_create(int cvs, int vsize) { return new cl_conv::Cldconv(the_device(), cvs, vsize, quiet, NULL); }
int ref_dconv_error(void *h) { return ((cl_conv::Cldconv *)h)->get_cl_err(); }
int ref_dconv_push_ir(void *h, float *ir) { return ((cl_conv::Cldconv *)h)->push_ir(ir); }
int ref_dconv_convolution(void *h, float *out, float *in) {
  return ((cl_conv::Cldconv *)h)->convolution(out, in);
}
int ref_dconv_convolution_tv(void *h, float *out, float *in1, float *in2) {
  return ((cl_conv::Cldconv *)h)->convolution(out, in1, in2);
}
void ref_dconv_destroy(void *h) { delete (cl_conv::Cldconv *)h; }

/* ---- timed multi-channel loops for the CPU baseline ------------------------------------ */
/* Each returns the wall seconds of the processing loop only (construction / push_ir excluded),
 * with `threads` OpenMP threads, one reference object per channel / transform slot. */

/* channels x nblocks blocks of pts samples; ir: channels*cvs, in/out: channels*nblocks*pts */
double ref_pconv_run(int channels, int cvs, int pts, int nblocks, float *ir, float *in, float *out,
                     int threads) {
  std::vector<cl_conv::Clpconv *> obj(channels);
  for (int c = 0; c < channels; c++) {
    obj[c] = new cl_conv::Clpconv(the_device(), cvs, pts, quiet, NULL);
    obj[c]->push_ir(ir + (size_t)c * cvs);
  }
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int c = 0; c < channels; c++)
    for (int b = 0; b < nblocks; b++) {
      size_t o = ((size_t)c * nblocks + b) * pts;
      obj[c]->convolution(out + o, in + o);
    }
  double t1 = now_s();
  for (auto p : obj) delete p;
  return t1 - t0;
}

/* batch in-place rffts of `size` reals each; one plan per thread */
double ref_rfft_run(int size, int batch, float *data, int fwd, int threads) {
  std::vector<cl_fft::Clrfft *> plan(threads);
  for (int t = 0; t < threads; t++) plan[t] = new cl_fft::Clrfft(the_device(), size, fwd != 0);
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int i = 0; i < batch; i++)
    plan[omp_get_thread_num()]->transform(reinterpret_cast<std::complex<float> *>(data + (size_t)i * size));
  double t1 = now_s();
  for (auto p : plan) delete p;
  return t1 - t0;
}

/* batch in-place cffts of N complex points each */
double ref_cfft_run(int N, int batch, float *data, int fwd, int threads) {
  std::vector<cl_fft::Clcfft *> plan(threads);
  for (int t = 0; t < threads; t++) plan[t] = new cl_fft::Clcfft(the_device(), N, fwd != 0);
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int i = 0; i < batch; i++)
    plan[omp_get_thread_num()]->transform(reinterpret_cast<std::complex<float> *>(data + (size_t)i * 2 * N));
  double t1 = now_s();
  for (auto p : plan) delete p;
  return t1 - t0;
}

/* channels x nblocks blocks of vsize samples through Cldconv */
double ref_dconv_run(int channels, int irsize, int vsize, int nblocks, float *ir, float *in, float *out,
                     int threads) {
  std::vector<cl_conv::Cldconv *> obj(channels);
  for (int c = 0; c < channels; c++) {
    obj[c] = new cl_conv::Cldconv(the_device(), irsize, vsize, quiet, NULL);
    obj[c]->push_ir(ir + (size_t)c * irsize);
  }
  double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int c = 0; c < channels; c++)
    for (int b = 0; b < nblocks; b++) {
      size_t o = ((size_t)c * nblocks + b) * vsize;
      obj[c]->convolution(out + o, in + o);
    }
  double t1 = now_s();
  for (auto p : obj) delete p;
  return t1 - t0;
}

} /* extern "C" */
