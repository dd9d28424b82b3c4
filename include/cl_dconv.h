// cl_dconv.h -- direct (time-domain) convolution class, B200 (sm_100a CUDA) build.
//
// Same public interface as the reference header of the same name (reference cl_dconv.h:14-67):
// class cl_conv::Cldconv with push_ir(), two convolution() overloads, get_cl_err(),
// cl_error_string(). Behind it: a register-tiled FP32 FMA kernel instead of one atomic add per
// multiply. Output convention kept: y[t] = sum_c ir[c] x[t-1-c], i.e. the linear convolution delayed by
// one sample, exactly what the reference's ring indexing yields (cl_dconv.cpp:40-41,124-125).
// State starts at zero (the reference leaves its buffers uninitialised, Q11).
#ifndef __CL_DCONV_H__
#define __CL_DCONV_H__
#include "cl_conv.h"

struct b2f_dconv;

namespace cl_conv {

class Cldconv {
  int irsize, vsize;
  b2f_dconv *handle;
  void (*err)(std::string s, void *uData);
  void *userData;
  int cl_err;

  static void msg(std::string str, void *userData) {
    if (userData == NULL) std::cout << str << std::endl;
  }

 public:
  /** device_id: device from clGetDeviceIDs; cvs: impulse response length; vsize: samples per call */
  Cldconv(cl_device_id device_id, int cvs, int vsize, void (*errs)(std::string s, void *d) = NULL,
          void *uData = NULL);
  ~Cldconv();
  Cldconv(const Cldconv &) = delete;
  Cldconv &operator=(const Cldconv &) = delete;

  const char *cl_error_string(int err) { return cl_string(err); }

  /** set the impulse response (cvs floats) */
  int push_ir(float *ir);

  /** vsize input samples -> vsize output samples */
  int convolution(float *output, float *input);

  /** time-varying: in2's block is written into the coefficient ring first */
  int convolution(float *out, float *in1, float *in2);

  int get_cl_err() { return cl_err; }
};
}  // namespace cl_conv
#endif
