// cl_classes.cpp -- the reference's C++ class interface (include/cl_fft.h, cl_conv.h, cl_dconv.h) and
// the device-enumeration compatibility calls (include/CL/opencl.h), implemented as thin host wrappers
// over the C ABI of libb200fft.so. Builds into libcl_fft.so -- the library name the reference's
// CMake target has (reference CMakeLists.txt:10), so its test programs link the same way.
//
// No arithmetic lives here: every transform()/convolution() is one synchronous b2f_*_host call, which
// mirrors the reference's blocking write / launch / blocking read (cl_fft.cpp:153-161).
#include <cstdio>
#include <cstring>

#include "b200fft.h"
#include "cl_conv.h"
#include "cl_dconv.h"
#include "cl_fft.h"

// ---- device enumeration (replaces the OpenCL platform layer for the callers) ----------------------
extern "C" cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint num_entries, cl_device_id *devices,
                                 cl_uint *num_devices) {
  int n = 0;
  if (b2f_device_count(&n) != B2F_OK || n <= 0) {
    if (num_devices) *num_devices = 0;
    return CL_DEVICE_NOT_FOUND;
  }
  if (devices)
    for (cl_uint i = 0; i < num_entries && (int)i < n; i++) devices[i] = b2f_cl_device_from_ordinal((int)i);
  if (num_devices) *num_devices = (cl_uint)n;
  return CL_SUCCESS;
}
extern "C" cl_int clGetDeviceInfo(cl_device_id device, cl_device_info param, size_t size, void *value,
                                  size_t *size_ret) {
  if (param != CL_DEVICE_NAME) return CL_INVALID_VALUE;
  char name[256];
  if (b2f_device_name(b2f_cl_device_ordinal(device), name, sizeof(name)) != B2F_OK) return CL_INVALID_DEVICE;
  const size_t need = strlen(name) + 1;
  if (value) {
    if (size == 0) return CL_INVALID_VALUE;
    snprintf((char *)value, size, "%s", name);
  }
  if (size_ret) *size_ret = need;
  return CL_SUCCESS;
}

namespace cl_fft {

const char *cl_error_string(int err) { return cl_conv::cl_string(err); }

static void note_failure(char *log, size_t n, int code) {
  snprintf(log, n, "%s (%s)", b2f_error_string(code), b2f_last_cuda_error());
}

Clcfft::Clcfft(int size, bool fwd, no_plan_t) : N(size), forward(fwd), cl_err(0), cplan(nullptr) { log[0] = 0; }

Clcfft::Clcfft(cl_device_id device_id, int size, bool fwd) : N(size), forward(fwd), cl_err(0), cplan(nullptr) {
  log[0] = 0;
  cl_err = b2f_cfft_create(&cplan, b2f_cl_device_ordinal(device_id), size, fwd ? 1 : 0, 1);
  if (cl_err) note_failure(log, sizeof(log), cl_err);
}
Clcfft::~Clcfft() { b2f_cfft_destroy(cplan); }

int Clcfft::transform(std::complex<float> *c) {
  if (!cplan) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return b2f_cfft_exec_host(cplan, reinterpret_cast<float *>(c), 1);
}

// the base class works on size/2 complex points, as in the reference (cl_fft.cpp:208-210)
Clrfft::Clrfft(cl_device_id device_id, int size, bool fwd) : Clcfft(size / 2, fwd, no_plan_t()), rplan(nullptr) {
  cl_err = b2f_rfft_create(&rplan, b2f_cl_device_ordinal(device_id), size, fwd ? 1 : 0, 1);
  if (cl_err) note_failure(log, sizeof(log), cl_err);
}
Clrfft::~Clrfft() { b2f_rfft_destroy(rplan); }

int Clrfft::transform(std::complex<float> *c, float *r) {
  if (!rplan) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return b2f_rfft_exec_host(rplan, reinterpret_cast<float *>(c), r, 1);
}
}  // namespace cl_fft

namespace cl_conv {

Clpconv::Clpconv(cl_device_id device_id, int cvs, int pts, void (*errs)(std::string s, void *d), void *uData, void *,
                 void *, void *)
    : N(pts << 1), bins(pts), bsize(pts > 0 ? (cvs / pts) * pts : 0), nparts(pts > 0 ? cvs / pts : 0), handle(nullptr),
      err(errs == NULL ? this->msg : errs), userData(uData), cl_err(CL_SUCCESS) {
  cl_err = b2f_pconv_create(&handle, b2f_cl_device_ordinal(device_id), cvs, pts, 1);
  if (cl_err) {
    err(cl_error_string(cl_err), userData);
    err(b2f_last_cuda_error(), userData);
  }
}
Clpconv::~Clpconv() { b2f_pconv_destroy(handle); }

int Clpconv::push_ir(float *ir) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return cl_err = b2f_pconv_push_ir_host(handle, ir, (size_t)bsize);
}
int Clpconv::convolution(float *output, float *input) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return cl_err = b2f_pconv_process_host(handle, output, input);
}
int Clpconv::convolution(float *output, float *input1, float *input2) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return cl_err = b2f_pconv_process_tv_host(handle, output, input1, input2);
}

Cldconv::Cldconv(cl_device_id device_id, int cvs, int vsiz, void (*errs)(std::string s, void *d), void *uData)
    : irsize(cvs), vsize(vsiz), handle(nullptr), err(errs == NULL ? this->msg : errs), userData(uData),
      cl_err(CL_SUCCESS) {
  cl_err = b2f_dconv_create(&handle, b2f_cl_device_ordinal(device_id), cvs, vsiz, 1, 1);
  if (cl_err) {
    err(cl_error_string(cl_err), userData);
    err(b2f_last_cuda_error(), userData);
  }
}
Cldconv::~Cldconv() { b2f_dconv_destroy(handle); }

int Cldconv::push_ir(float *ir) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  return b2f_dconv_push_ir_host(handle, ir, (size_t)irsize);
}
int Cldconv::convolution(float *output, float *input) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  cl_err = b2f_dconv_process_host(handle, output, input, 1);
  if (cl_err) err(cl_error_string(cl_err), userData);
  return cl_err;
}
int Cldconv::convolution(float *out, float *in1, float *in2) {
  if (!handle) return cl_err ? cl_err : B2F_ERR_INVALID_VALUE;
  cl_err = b2f_dconv_process_tv_host(handle, out, in1, in2);
  if (cl_err) err(cl_error_string(cl_err), userData);
  return cl_err;
}
}  // namespace cl_conv
