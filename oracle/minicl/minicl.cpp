/*
 * oracle/minicl/minicl.cpp -- a minimal host-CPU OpenCL runtime. TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: run the UNMODIFIED reference (host C++ and its OpenCL-C kernel strings) in an image
 * that has no OpenCL platform (no PoCL, no vendor ICD, no network). It implements exactly the
 * 21 entry points the reference calls (oracle/minicl/CL/opencl.h). clBuildProgram takes the
 * kernel source string the reference hands over, prepends a small C++ prelude that gives
 * OpenCL-C's float2 / get_global_id / atomic_cmpxchg their meaning, rewrites the one construct
 * C++ cannot parse (the vector literal `(cmplx)(a, b)` -> `mk2(a, b)`), compiles it with the
 * host g++ (-O2 -ffp-contract=off) into a shared object cached under MINICL_CACHE (default:
 * <dir of this library>/kcache) and dlopen()s it. An NDRange runs its work-items one after
 * another in ascending global-id order on the calling thread: a legal OpenCL schedule, and the
 * one oracle/ref_cpu.c restates, so both agree bit for bit. In-order queues execute eagerly.
 *
 * Nothing here is reference code; it is a runtime the reference runs ON.
 */
#include "CL/opencl.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

struct _cl_platform_id {
  int unused;
};
struct _cl_device_id {
  const char *name;
};
struct _cl_context {
  int refs;
};
struct _cl_command_queue {
  int refs;
};
struct _cl_mem {
  void *ptr;
  size_t size;
  bool owns;
};
struct KernelSig {
  std::vector<bool> is_ptr;
};
struct _cl_program {
  std::string source;
  std::string log;
  void *dl = nullptr;
  std::map<std::string, KernelSig> sigs;
  int refs = 1;
};
typedef void (*run_fn)(void **, size_t, size_t);
struct KArg {
  unsigned char bytes[16];
  size_t size = 0;
  bool set = false;
};
struct _cl_kernel {
  std::string name;
  run_fn run = nullptr;
  KernelSig sig;
  std::vector<KArg> args;
};

static _cl_device_id g_device = {"minicl host CPU (sequential work-items)"};

static const char *k_prelude = R"PRE(
#include <cstddef>
#include <cstdint>
typedef unsigned int uint;
struct float2 {
  float x, y;
  float2() = default;
  float2(float a, float b) : x(a), y(b) {}
  float2 &operator=(float s) { x = s; y = s; return *this; }
};
static inline float2 mk2(float a, float b) { return float2(a, b); }
static inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
static inline float2 operator-(float2 a, float2 b) { return float2(a.x - b.x, a.y - b.y); }
static inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
static inline float2 operator*(float s, float2 a) { return float2(s * a.x, s * a.y); }
static inline float2 operator*(float2 a, float s) { return float2(a.x * s, a.y * s); }
static inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
static inline float2 operator/(float2 a, float s) { return float2(a.x / s, a.y / s); }
static inline float2 operator/(float2 a, int s) { float f = (float)s; return float2(a.x / f, a.y / f); }
static thread_local size_t minicl_gid0;
static inline int get_global_id(int) { return (int)minicl_gid0; }
static inline uint atomic_cmpxchg(volatile uint *p, uint cmp, uint val) {
  return __sync_val_compare_and_swap(p, cmp, val);
}
#define kernel static
#define __kernel static
#define global
#define __global
#define constant const
#define __constant const
#line 1 "opencl_kernel_source"
)PRE";

static std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

/* strip address-space / cv words from one parameter and split "type name" */
static void parse_param(const std::string &param, std::string &type, bool &is_ptr) {
  std::string p = std::regex_replace(param, std::regex("\\b(__global|global|__constant|constant)\\b"), " ");
  p = trim(p);
  size_t star = p.rfind('*');
  if (star != std::string::npos) {
    is_ptr = true;
    type = trim(p.substr(0, star + 1));
  } else {
    is_ptr = false;
    size_t sp = p.find_last_of(" \t\r\n");
    type = trim(p.substr(0, sp));
  }
}

static std::string translate(_cl_program *prog) {
  std::string src = prog->source;
  /* OpenCL-C vector literal -> function call */
  src = std::regex_replace(src, std::regex("\\(\\s*cmplx\\s*\\)\\s*\\("), "mk2(");
  std::ostringstream out;
  out << k_prelude << src << "\n";
  /* one runner per kernel: unpack the argument block, loop over the global range */
  std::regex sig("kernel\\s+void\\s+(\\w+)\\s*\\(([^)]*)\\)");
  for (std::sregex_iterator it(prog->source.begin(), prog->source.end(), sig), end; it != end; ++it) {
    std::string name = (*it)[1], params = (*it)[2];
    KernelSig ks;
    std::ostringstream call;
    std::stringstream ss(params);
    std::string one;
    int idx = 0;
    while (std::getline(ss, one, ',')) {
      std::string type;
      bool is_ptr;
      parse_param(one, type, is_ptr);
      ks.is_ptr.push_back(is_ptr);
      if (idx) call << ", ";
      if (is_ptr)
        call << "(" << type << ")a[" << idx << "]";
      else
        call << "*(" << type << " *)a[" << idx << "]";
      idx++;
    }
    prog->sigs[name] = ks;
    out << "extern \"C\" void minicl_run_" << name << "(void **a, size_t g0, size_t g1) {\n"
        << "  for (size_t g = g0; g < g1; ++g) { minicl_gid0 = g; " << name << "(" << call.str()
        << "); }\n}\n";
  }
  return out.str();
}

static std::string self_dir() {
  Dl_info info;
  if (dladdr((void *)&self_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t s = p.rfind('/');
    if (s != std::string::npos) return p.substr(0, s);
  }
  return ".";
}

static uint64_t fnv1a(const std::string &s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

extern "C" {

cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint n, cl_device_id *ids, cl_uint *num) {
  if (ids && n > 0) ids[0] = &g_device;
  if (num) *num = 1;
  return CL_SUCCESS;
}

cl_int clGetDeviceInfo(cl_device_id d, cl_device_info what, size_t sz, void *out, size_t *ret) {
  if (what != CL_DEVICE_NAME) return CL_INVALID_VALUE;
  size_t need = strlen(d->name) + 1;
  if (out) {
    if (sz < need) return CL_INVALID_VALUE;
    memcpy(out, d->name, need);
  }
  if (ret) *ret = need;
  return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *err) {
  if (err) *err = CL_SUCCESS;
  return new _cl_context{1};
}
cl_int clReleaseContext(cl_context c) {
  delete c;
  return CL_SUCCESS;
}
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *err) {
  if (err) *err = CL_SUCCESS;
  return new _cl_command_queue{1};
}
cl_int clReleaseCommandQueue(cl_command_queue q) {
  delete q;
  return CL_SUCCESS;
}

cl_mem clCreateBuffer(cl_context, cl_mem_flags flags, size_t size, void *host, cl_int *err) {
  _cl_mem *m = new _cl_mem;
  m->size = size;
  if ((flags & CL_MEM_USE_HOST_PTR) && host) {
    m->ptr = host;
    m->owns = false;
  } else {
    m->ptr = calloc(size ? size : 1, 1); /* uninitialised in OpenCL; zeros here */
    m->owns = true;
  }
  if (err) *err = CL_SUCCESS;
  return m;
}
cl_int clReleaseMemObject(cl_mem m) {
  if (!m) return CL_INVALID_MEM_OBJECT;
  if (m->owns) free(m->ptr);
  delete m;
  return CL_SUCCESS;
}

cl_program clCreateProgramWithSource(cl_context, cl_uint count, const char **strings, const size_t *lengths,
                                     cl_int *err) {
  _cl_program *p = new _cl_program;
  for (cl_uint i = 0; i < count; i++)
    p->source += lengths && lengths[i] ? std::string(strings[i], lengths[i]) : std::string(strings[i]);
  if (err) *err = CL_SUCCESS;
  return p;
}

cl_int clBuildProgram(cl_program p, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *),
                      void *) {
  const char *cxx = getenv("MINICL_CXX") ? getenv("MINICL_CXX") : "g++";
  const char *flags = "-O2 -ffp-contract=off -fno-fast-math -shared -fPIC -std=c++14 -w";
  std::string code = translate(p);
  const char *envdir = getenv("MINICL_CACHE");
  std::string dir = envdir ? std::string(envdir) : self_dir() + "/kcache";
  mkdir(dir.c_str(), 0777);
  char tag[32];
  snprintf(tag, sizeof(tag), "%016llx", (unsigned long long)fnv1a(code + flags));
  std::string so = dir + "/k" + tag + ".so";
  if (access(so.c_str(), R_OK) != 0) {
    std::string stem = dir + "/k" + tag + "." + std::to_string((long)getpid());
    std::string cpp = stem + ".cpp", tmpso = stem + ".so", logf = stem + ".log";
    {
      std::ofstream f(cpp);
      f << code;
    }
    std::string cmd = std::string(cxx) + " " + flags + " -o " + tmpso + " " + cpp + " > " + logf + " 2>&1";
    int rc = system(cmd.c_str());
    std::ifstream lf(logf);
    std::stringstream ls;
    ls << lf.rdbuf();
    p->log = ls.str();
    unlink(logf.c_str());
    if (rc != 0) {
      unlink(tmpso.c_str());
      return CL_BUILD_PROGRAM_FAILURE;
    }
    unlink(cpp.c_str());
    rename(tmpso.c_str(), so.c_str());
  }
  p->dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!p->dl) {
    p->log = dlerror();
    return CL_BUILD_PROGRAM_FAILURE;
  }
  return CL_SUCCESS;
}

cl_int clGetProgramBuildInfo(cl_program p, cl_device_id, cl_program_build_info what, size_t sz, void *out,
                             size_t *ret) {
  if (what != CL_PROGRAM_BUILD_LOG) return CL_INVALID_VALUE;
  size_t n = p->log.size() + 1;
  if (out && sz) {
    size_t c = n < sz ? n : sz;
    memcpy(out, p->log.c_str(), c);
    ((char *)out)[c - 1] = 0;
  }
  if (ret) *ret = n;
  return CL_SUCCESS;
}
cl_int clReleaseProgram(cl_program p) {
  /* kernels keep raw function pointers into p->dl, so the library stays loaded */
  if (p && --p->refs == 0) delete p;
  return CL_SUCCESS;
}

cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *err) {
  if (!p || !p->dl) {
    if (err) *err = CL_INVALID_PROGRAM_EXECUTABLE;
    return nullptr;
  }
  auto it = p->sigs.find(name);
  std::string sym = std::string("minicl_run_") + name;
  void *f = dlsym(p->dl, sym.c_str());
  if (it == p->sigs.end() || !f) {
    if (err) *err = CL_INVALID_KERNEL_NAME;
    return nullptr;
  }
  _cl_kernel *k = new _cl_kernel;
  k->name = name;
  k->run = (run_fn)f;
  k->sig = it->second;
  k->args.resize(k->sig.is_ptr.size());
  if (err) *err = CL_SUCCESS;
  return k;
}
cl_int clReleaseKernel(cl_kernel k) {
  delete k;
  return CL_SUCCESS;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value) {
  if (!k) return CL_INVALID_KERNEL;
  if (idx >= k->args.size()) return CL_INVALID_ARG_INDEX;
  if (size > sizeof(k->args[idx].bytes) || !value) return CL_INVALID_ARG_SIZE;
  memcpy(k->args[idx].bytes, value, size);
  k->args[idx].size = size;
  k->args[idx].set = true;
  return CL_SUCCESS;
}

cl_int clGetKernelWorkGroupInfo(cl_kernel, cl_device_id, cl_kernel_work_group_info what, size_t sz, void *out,
                                size_t *ret) {
  if (what != CL_KERNEL_WORK_GROUP_SIZE) return CL_INVALID_VALUE;
  size_t v = 1024;
  if (out && sz >= sizeof(size_t)) memcpy(out, &v, sizeof(size_t));
  if (ret) *ret = sizeof(size_t);
  return CL_SUCCESS;
}

cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t n, const void *src, cl_uint,
                            const cl_event *, cl_event *) {
  if (!m) return CL_INVALID_MEM_OBJECT;
  if (off + n > m->size) return CL_INVALID_VALUE;
  memcpy((char *)m->ptr + off, src, n);
  return CL_SUCCESS;
}
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t n, void *dst, cl_uint,
                           const cl_event *, cl_event *) {
  if (!m) return CL_INVALID_MEM_OBJECT;
  if (off + n > m->size) return CL_INVALID_VALUE;
  memcpy(dst, (char *)m->ptr + off, n);
  return CL_SUCCESS;
}
cl_int clEnqueueFillBuffer(cl_command_queue, cl_mem m, const void *pattern, size_t psz, size_t off, size_t n,
                           cl_uint, const cl_event *, cl_event *) {
  if (!m) return CL_INVALID_MEM_OBJECT;
  if (off + n > m->size) return CL_INVALID_VALUE;
  for (size_t i = 0; i + psz <= n; i += psz) memcpy((char *)m->ptr + off + i, pattern, psz);
  return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t *goff, const size_t *gsz,
                              const size_t *, cl_uint, const cl_event *, cl_event *) {
  if (!k) return CL_INVALID_KERNEL;
  if (dim != 1) return CL_INVALID_WORK_DIMENSION;
  void *argv[16];
  if (k->args.size() > 16) return CL_INVALID_KERNEL_ARGS;
  for (size_t i = 0; i < k->args.size(); i++) {
    if (!k->args[i].set) return CL_INVALID_KERNEL_ARGS;
    if (k->sig.is_ptr[i]) {
      cl_mem m;
      memcpy(&m, k->args[i].bytes, sizeof(cl_mem));
      argv[i] = m ? m->ptr : nullptr;
    } else
      argv[i] = k->args[i].bytes;
  }
  size_t g0 = goff ? goff[0] : 0;
  k->run(argv, g0, g0 + gsz[0]);
  return CL_SUCCESS;
}

cl_int clFinish(cl_command_queue) { return CL_SUCCESS; }

} /* extern "C" */
