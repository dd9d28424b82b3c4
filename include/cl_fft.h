// cl_fft.h -- 1-D radix-2-semantics FFT classes, B200 (sm_100a CUDA) build.
//
// Same public interface as the reference header of the same name (reference cl_fft.h:22-112):
// namespace cl_fft, PI, cl_error_string(), class Clcfft with a virtual transform(), class Clrfft
// derived from it. Callers written against the reference (test_cfft.cpp, test_rfft.cpp,
// csound/opcode.cpp) compile unchanged. What differs is everything private: no OpenCL objects, just
// an opaque plan of the C ABI in b200fft.h, which runs hand-written CUDA kernels.
//
// Behavioural contract kept from the reference:
//   forward complex transform scaled by 1/N, inverse unscaled (reference cl_fft.cpp:39-40)
//   real transform packing/scaling as in reference cl_fft.cpp:178-205, 267-296
//   get_error() == 0 after a successful construction; failures are reported as POSITIVE codes
//   (b200fft.h B2F_ERR_*), so the callers' `err > 0` tests (test_cfft.cpp:41) actually fire.
#ifndef __CL_FFT_H__
#define __CL_FFT_H__

#include <complex>
#include <iostream>

#include <CL/opencl.h>

struct b2f_cfft;
struct b2f_rfft;

namespace cl_fft {

const double PI = 3.141592653589793;

// text for a status code returned by these classes (0 or a positive B2F_ERR_* value)
const char *cl_error_string(int err);

/** Complex-to-complex FFT. One object = one plan on one device, like the reference. */
class Clcfft {
 protected:
  int N;         // number of complex points handled by the underlying plan
  bool forward;  // direction
  int cl_err;    // construction status
  char log[2048];
  b2f_cfft *cplan;

  struct no_plan_t {};
  // for Clrfft: set the bookkeeping up without creating a complex plan
  Clcfft(int size, bool fwd, no_plan_t);

 public:
  /** device_id: from clGetDeviceIDs (a CUDA device here); size: N complex points, power of two;
      fwd: true = forward (scaled by 1/N), false = inverse (unscaled) */
  Clcfft(cl_device_id device_id, int size, bool fwd = true);
  virtual ~Clcfft();
  Clcfft(const Clcfft &) = delete;
  Clcfft &operator=(const Clcfft &) = delete;

  /** in-place transform of N complex numbers held by the host */
  virtual int transform(std::complex<float> *c);

  /** construction status: 0 on success */
  int get_error() { return cl_err; }

  /** diagnostic text of a failed construction (the reference returned its OpenCL build log here) */
  const char *get_log() { return (const char *)log; }
};

/** Real-to-complex / complex-to-real FFT of `size` real points (size/2 packed complex bins). */
class Clrfft : public Clcfft {
  b2f_rfft *rplan;

 public:
  Clrfft(cl_device_id device_id, int size, bool fwd);
  virtual ~Clrfft();

  /** c: size/2 complex numbers, r: size reals. Same memory = in place. Forward reads r and writes c;
      inverse reads c and writes r (c is overwritten as well, as in the reference). */
  int transform(std::complex<float> *c, float *r);

  /** in-place form on one array viewed both ways */
  virtual int transform(std::complex<float> *c) {
    float *r = reinterpret_cast<float *>(c);
    return transform(c, r);
  }
};
}  // namespace cl_fft

#endif
