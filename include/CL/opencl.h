/*
 * include/CL/opencl.h -- compatibility names, NOT OpenCL.
 *
 * The reference's callers (test_cfft.cpp:24-38, test_rfft.cpp:24-38, csound/opcode.cpp:50-61,
 * 168-176, 269-277) pick a device with clGetDeviceIDs / clGetDeviceInfo and hand a cl_device_id to
 * the class constructors. This build has no OpenCL at all; so that those sources compile UNCHANGED,
 * this header supplies exactly the names they use, mapped onto CUDA device ordinals:
 *   cl_device_id  = opaque handle encoding (CUDA ordinal + 1)
 *   clGetDeviceIDs(NULL, CL_DEVICE_TYPE_ALL, n, ids, &num) -> the visible CUDA devices
 *   clGetDeviceInfo(id, CL_DEVICE_NAME, ...)               -> cudaDeviceProp::name
 * Status codes keep their Khronos values so that printed numbers keep their meaning.
 */
#ifndef B200FFT_CL_COMPAT_H
#define B200FFT_CL_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_uint cl_device_info;
typedef struct b2f_cl_platform *cl_platform_id;
typedef struct b2f_cl_device *cl_device_id;

#define CL_SUCCESS 0
#define CL_DEVICE_NOT_FOUND -1
#define CL_INVALID_VALUE -30
#define CL_INVALID_DEVICE -33
#define CL_DEVICE_TYPE_DEFAULT (1 << 0)
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_DEVICE_NAME 0x102B

cl_int clGetDeviceIDs(cl_platform_id platform, cl_device_type type, cl_uint num_entries, cl_device_id *devices,
                      cl_uint *num_devices);
cl_int clGetDeviceInfo(cl_device_id device, cl_device_info param, size_t size, void *value, size_t *size_ret);

/* the CUDA ordinal behind a handle returned by clGetDeviceIDs (-1 for NULL) */
static inline int b2f_cl_device_ordinal(cl_device_id id) { return (int)(intptr_t)id - 1; }
static inline cl_device_id b2f_cl_device_from_ordinal(int ordinal) { return (cl_device_id)(intptr_t)(ordinal + 1); }

#ifdef __cplusplus
}
#endif
#endif
