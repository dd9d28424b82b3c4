// cl_conv.h -- uniformly-partitioned (overlap-add, frequency-domain delay line) convolution class,
// B200 (sm_100a CUDA) build.
//
// Same public interface as the reference header of the same name (reference cl_conv.h:23-188):
// namespace cl_conv, cl_string(), class Clpconv with push_ir(), two convolution() overloads,
// get_cl_err(), cl_error_string(). The private OpenCL state is replaced by one opaque handle of the
// C ABI (b200fft.h); one fused CUDA launch per block replaces the reference's 24 kernel launches.
//
// Kept from the reference: nparts = cvs / pts truncating (cl_conv.cpp:143); IR partition i stored
// in ring frame nparts-1-i (385); per-block semantics of cl_conv.cpp:393-548 including the half-weight
// DC/Nyquist bin; error callback convention (cl_conv.h:137-145).
// Not kept: the optional host-memory constructor arguments are accepted and ignored -- in the
// reference they are dead code (cl_conv.cpp:151,232-237).
#ifndef __CL_CONV_H__
#define __CL_CONV_H__
#include <complex>
#include <iostream>
#include <string>

#include <CL/opencl.h>

struct b2f_pconv;
extern "C" const char *b2f_error_string(int code);

namespace cl_conv {

/** text for a status code: 0 -> "Success!", positive -> the engine's B2F_ERR_* text. Negative values
    were OpenCL codes in the reference; nothing in this build produces them. */
inline const char *cl_string(int err) {
  switch (err) {
  case CL_DEVICE_NOT_FOUND:
    return "Device not found.";
  case CL_INVALID_VALUE:
    return "Invalid value";
  case CL_INVALID_DEVICE:
    return "Invalid device";
  default:
    if (err < 0) return "OpenCL status code (this build has no OpenCL runtime)";
    return b2f_error_string(err);
  }
}

class Clpconv {
  int N, bins, bsize, nparts;
  b2f_pconv *handle;
  void (*err)(std::string s, void *uData);
  void *userData;
  int cl_err;

  static void msg(std::string str, void *userData) {
    if (userData == NULL) std::cout << str << std::endl;
  }

 public:
  /** device_id: device from clGetDeviceIDs; cvs: impulse response length; pts: partition size (power of
      two); errs/uData: error message callback and its user data; in1, in2, out: ignored (see above) */
  Clpconv(cl_device_id device_id, int cvs, int pts, void (*errs)(std::string s, void *d) = NULL,
          void *uData = NULL, void *in1 = NULL, void *in2 = NULL, void *out = 0);
  ~Clpconv();
  Clpconv(const Clpconv &) = delete;
  Clpconv &operator=(const Clpconv &) = delete;

  const char *cl_error_string(int err) { return cl_string(err); }

  /** set the impulse response: reads (cvs / pts) * pts floats */
  int push_ir(float *ir);

  /** one block: pts input samples -> pts output samples */
  int convolution(float *output, float *input);

  /** time-varying: input2's block is transformed into the impulse-response ring */
  int convolution(float *output, float *input1, float *input2);

  /** recorded status, CL_SUCCESS (0) if none */
  int get_cl_err() { return cl_err; }
};
}  // namespace cl_conv
#endif
