// capi.cu -- implementation of the C ABI declared in include/b200fft.h.
//
// Host-side responsibilities only: validating sizes, building twiddle tables in double precision
// exactly as the reference does (cl_fft.cpp:86-91, 233-238; cl_conv.cpp:263-287), owning device
// state, staging host buffers through pinned memory, and launching the sm_100a kernels in
// fft_kernels.cuh / fft_large.cuh / pconv_kernels.cuh / dconv_kernels.cuh. There is deliberately
// no CPU implementation behind these entry points: if CUDA is unavailable every create fails.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200fft.h"
#include "dconv_kernels.cuh"
#include "fft_kernels.cuh"
#include "fft_large.cuh"
#include "fft_sm.cuh"
#include "pconv_kernels.cuh"

using namespace b2f;

// ---- error plumbing -------------------------------------------------------------------------------
static thread_local std::string g_cuda_err = "";

static int cuda_fail(cudaError_t e, const char *where) {
  g_cuda_err = std::string(where) + ": " + cudaGetErrorString(e);
  (void)cudaGetLastError();  // clear the sticky-less error state
  if (e == cudaErrorMemoryAllocation) return B2F_ERR_ALLOC;
  if (e == cudaErrorNoDevice || e == cudaErrorInvalidDevice || e == cudaErrorInsufficientDriver)
    return B2F_ERR_NO_DEVICE;
  return B2F_ERR_CUDA;
}
#define CK(call)                                       \
  do {                                                 \
    cudaError_t e_ = (call);                           \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

extern "C" const char *b2f_error_string(int code) {
  switch (code) {
    case B2F_OK: return "Success!";
    case B2F_ERR_NO_DEVICE: return "CUDA device not found";
    case B2F_ERR_INVALID_VALUE: return "Invalid value";
    case B2F_ERR_UNSUPPORTED: return "Size not supported by this build";
    case B2F_ERR_ALLOC: return "Memory allocation failure";
    case B2F_ERR_CUDA: return "CUDA runtime failure";
    case B2F_ERR_BATCH: return "Batch exceeds the plan's max_batch";
    default: return "Unknown error";
  }
}
extern "C" const char *b2f_last_cuda_error(void) { return g_cuda_err.c_str(); }
extern "C" const char *b2f_version(void) { return "b200fft 0.1 (sm_100a)"; }

// ---- options ----------------------------------------------------------------------------------------
// Re-measurement knobs. They are process-wide DEFAULTS that a handle copies when it is created: no entry point
// reads the environment or this table afterwards, so a running handle never changes behaviour behind the
// caller's back. Initial values come from the B2F_* environment variables, read once, before the first create;
// b2f_set_option() changes them for handles created later (include/b200fft.h lists the names).
struct Options {
  long long fft_sm_min_batch = 96;   // N = 2^15: one-SM kernel from this batch up (0: never, 1: always)
  long long fft_sm_8192 = 1;         // complex N = 8192 and inverse real 16384 on the one-SM kernel too (four transforms per unit), from 4 x fft_sm_min_batch
  long long large_chunk_mb = 256;    // scratch chunk of the four-step launch pair
  long long rows_rb16 = 0;           // 16-row CTAs in the real rows kernel
  long long separate_split = 0;      // unfused real split / unsplit pass (four-step path)
  long long pconv_tma = -1;          // MAC feed of the partitioned convolution: -1 measured choice, 0 registers, 1 TMA
  long long pconv_cluster = 0;       // cluster split of the partitions: 0 measured choice, else 1 / 2 / 4 / 8
  long long pconv_pipeline = 1;      // two-stream host calls (pconv: halves of the channels; FFT: chunks of the batch)
  long long zerocopy_max = 65536;    // host calls up to this many bytes run on the pinned buffers directly
  long long graph = 1;               // CUDA graph for the multi-launch host paths
  long long pinned_direct = 1;       // pconv host calls on caller-pinned buffers: the kernel reads / writes them in place
  long long fft_prefetch = -1;       // real one-CTA transforms of N >= 8192: L2 prefetch distance in CTAs (-1: the resident CTAs, 0: off)
  long long pconv_cluster16_max_channels = 4;  // clusters of 16 CTAs for up to this many channels with long IRs (0: never)
  long long pconv_deep_ring = 1;     // launches of at most one CTA per SM (pts 2048 / 4096): TMA stages of 32 KB
  long long pconv_ksplit = 0;        // general path (pts >= 8192): partitions split over this many CTAs per tile (0: measured choice, -1: never)
  long long pconv_general_fused = 1; // pts 8192 / 16384: frames and inverse + overlap-add as fused launches (0: pad / rFFT / copy / ... one by one)
  long long pconv_deep_min_parts = 16;  // ... from this many partitions per CTA (measured: mono pts 2048 x 1024 partitions 65.9 -> 31.0 us, x 512: 39.2 -> 22.7)
  long long pconv_push_reg = 1;      // push_ir on the register-level real transform (pts >= 64); 0: the step kernel's frame routine
  long long verbose = 0;
};
struct OptionName {
  const char *name, *env;
  long long Options::*field;
};
static const OptionName kOptionNames[] = {
    {"fft_sm_min_batch", "B2F_FFT_SM_MIN_BATCH", &Options::fft_sm_min_batch},
    {"fft_sm_8192", "B2F_FFT_SM_8192", &Options::fft_sm_8192},
    {"large_chunk_mb", "B2F_LARGE_CHUNK_MB", &Options::large_chunk_mb},
    {"rows_rb16", "B2F_ROWS_RB16", &Options::rows_rb16},
    {"separate_split", "B2F_SEPARATE_SPLIT", &Options::separate_split},
    {"pconv_tma", "B2F_PCONV_TMA", &Options::pconv_tma},
    {"pconv_cluster", "B2F_PCONV_CLUSTER", &Options::pconv_cluster},
    {"pconv_pipeline", "B2F_PCONV_PIPELINE", &Options::pconv_pipeline},
    {"zerocopy_max", "B2F_ZEROCOPY_MAX", &Options::zerocopy_max},
    {"graph", "B2F_GRAPH", &Options::graph},
    {"pinned_direct", "B2F_PINNED_DIRECT", &Options::pinned_direct},
    {"fft_prefetch", "B2F_FFT_PREFETCH", &Options::fft_prefetch},
    {"pconv_cluster16_max_channels", "B2F_PCONV_CLUSTER16_MAX_CHANNELS", &Options::pconv_cluster16_max_channels},
    {"pconv_deep_ring", "B2F_PCONV_DEEP_RING", &Options::pconv_deep_ring},
    {"pconv_ksplit", "B2F_PCONV_KSPLIT", &Options::pconv_ksplit},
    {"pconv_general_fused", "B2F_PCONV_GENERAL_FUSED", &Options::pconv_general_fused},
    {"pconv_deep_min_parts", "B2F_PCONV_DEEP_MIN_PARTS", &Options::pconv_deep_min_parts},
    {"pconv_push_reg", "B2F_PCONV_PUSH_REG", &Options::pconv_push_reg},
    {"verbose", "B2F_VERBOSE", &Options::verbose},
};
static std::mutex g_opt_mutex;
static Options &options_locked() {
  static Options o = [] {
    Options v;
    for (const OptionName &n : kOptionNames)
      if (const char *e = getenv(n.env)) v.*(n.field) = atoll(e);
    return v;
  }();
  return o;
}
static Options current_options() {
  std::lock_guard<std::mutex> lk(g_opt_mutex);
  return options_locked();
}
extern "C" int b2f_set_option(const char *name, long long value) {
  if (!name) return B2F_ERR_INVALID_VALUE;
  std::lock_guard<std::mutex> lk(g_opt_mutex);
  for (const OptionName &n : kOptionNames)
    if (!strcmp(name, n.name)) {
      options_locked().*(n.field) = value;
      return B2F_OK;
    }
  return B2F_ERR_INVALID_VALUE;
}
extern "C" int b2f_get_option(const char *name, long long *value) {
  if (!name || !value) return B2F_ERR_INVALID_VALUE;
  std::lock_guard<std::mutex> lk(g_opt_mutex);
  for (const OptionName &n : kOptionNames)
    if (!strcmp(name, n.name)) {
      *value = options_locked().*(n.field);
      return B2F_OK;
    }
  return B2F_ERR_INVALID_VALUE;
}

extern "C" int b2f_device_count(int *count) {
  if (!count) return B2F_ERR_INVALID_VALUE;
  *count = 0;
  CK(cudaGetDeviceCount(count));
  return *count > 0 ? B2F_OK : B2F_ERR_NO_DEVICE;
}
extern "C" int b2f_device_name(int device, char *buf, size_t buflen) {
  if (!buf || buflen == 0) return B2F_ERR_INVALID_VALUE;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  snprintf(buf, buflen, "%s", prop.name);
  return B2F_OK;
}

// Make `device` current for the duration of an entry point and put the caller's device back afterwards: the
// library must not change the calling thread's current device behind the application's (or torch's) back.
static thread_local int g_device = -1;  // the device the running entry point made current (set_smem's cache key)
struct DeviceGuard {
  int prev = -1, rc = B2F_OK, prev_g = -1;
  explicit DeviceGuard(int device) {
    prev_g = g_device;
    g_device = device;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) {
      cudaError_t e = cudaSetDevice(device);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        rc = B2F_ERR_NO_DEVICE;
        prev = -1;
      }
    } else {
      prev = -1;  // nothing to restore
    }
  }
  ~DeviceGuard() {
    g_device = prev_g;
    if (prev >= 0 && cudaSetDevice(prev) != cudaSuccess) (void)cudaGetLastError();
  }
};
#define B2F_ON_DEVICE(dev)       \
  DeviceGuard device_guard_(dev); \
  if (device_guard_.rc) return device_guard_.rc

// device-pointer entry points use 64/128-bit accesses: base pointers must be 16-byte aligned (include/b200fft.h)
static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int ilog2_exact(int n) {
  if (n <= 0 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}
static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (device < 0 || device >= n) return B2F_ERR_NO_DEVICE;
  return B2F_OK;
}

// ---- twiddle tables ---------------------------------------------------------------------------------
static const double kPI = 3.141592653589793;  // cl_fft.h:24

// entry q of the reference's N-entry table (cl_fft.cpp:86-91), forward sign; the inverse is its conjugate
static inline float2 ref_twiddle(long long q, int N) {
  float2 w;
  w.x = (float)cos(q * 2 * kPI / N);
  w.y = (float)(-1.f * sin(q * 2 * kPI / N));
  return w;
}
// pass twiddles of the in-shared-memory engine: for pass p >= 1, rows r = 1..R-1, columns k < NS:
// W_N^(r * k * N / (NS * R)), i.e. entries of the reference's table at that index
static std::vector<float2> make_pass_twiddles(int logn) {
  Sched s = sched_for(logn);
  const int N = 1 << logn;
  std::vector<float2> tw((size_t)sched_tw_total(s) + 1);
  for (int p = 1; p < s.npass; p++) {
    const int R = s.radix[p], NS = sched_stride(s, p), off = sched_tw_offset(s, p);
    const int step = N / (NS * R);
    for (int r = 1; r < R; r++)
      for (int k = 0; k < NS; k++) tw[off + (r - 1) * NS + k] = ref_twiddle((long long)r * k * step, N);
  }
  return tw;
}
// split twiddles w2[i] = exp(-i pi i / N) (cl_fft.cpp:233-238, forward sign), N entries
static std::vector<float2> make_split_twiddles(int N) {
  std::vector<float2> w(N);
  for (int i = 0; i < N; i++) {
    w[i].x = (float)cos(i * kPI / N);
    w[i].y = (float)(-1.f * sin(i * kPI / N));
  }
  return w;
}
static int upload(const std::vector<float2> &h, float2 **d) {
  CK(cudaMalloc((void **)d, h.size() * sizeof(float2)));
  CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
  return B2F_OK;
}

// ---- kernel dispatch ----------------------------------------------------------------------------------
// Opt a kernel in to more than 48 KB of dynamic shared memory: once per (kernel, device), not once per launch --
// cudaFuncSetAttribute costs about as much as the launch itself on the single-block latency path.
static int set_smem_once(const void *fn, int bytes) {
  static std::mutex m;
  static std::unordered_map<uint64_t, int> done;
  const uint64_t key = (uint64_t)(uintptr_t)fn * 64u + (uint64_t)(g_device < 0 ? 63 : g_device & 63);
  {
    std::lock_guard<std::mutex> lk(m);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return B2F_OK;
  }
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  std::lock_guard<std::mutex> lk(m);
  done[key] = bytes;
  return B2F_OK;
}
// Clusters of 16 CTAs are beyond the portable limit of 8: opt the kernel in, once per (kernel, device)
static int allow_cluster16_once(const void *fn) {
  static std::mutex m;
  static std::unordered_map<uint64_t, int> done;
  const uint64_t key = (uint64_t)(uintptr_t)fn * 64u + (uint64_t)(g_device < 0 ? 63 : g_device & 63);
  {
    std::lock_guard<std::mutex> lk(m);
    if (done.count(key)) return B2F_OK;
  }
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  std::lock_guard<std::mutex> lk(m);
  done[key] = 1;
  return B2F_OK;
}
template <class K>
static int set_smem(K kernel, int bytes) {
  return bytes > 48 * 1024 ? set_smem_once((const void *)kernel, bytes) : B2F_OK;
}

// SM count of the device the running entry point made current
static int sm_count() {
  static std::mutex m;
  static int cached[64] = {0};
  const int d = g_device < 0 ? 0 : (g_device & 63);
  std::lock_guard<std::mutex> lk(m);
  if (!cached[d]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n < 1) n = 148;
    cached[d] = n;
  }
  return cached[d];
}
// L2 prefetch distance of the real one-transform-per-CTA kernels, in CTAs: option fft_prefetch, -1 = the CTAs resident at a time
template <class B>
static int prefetch_distance(long long opt) {
  if (!B::PREFETCH || opt == 0) return 0;
  return opt > 0 ? (int)opt : sm_count() * B::MIN_BLOCKS;
}

template <int LOGN>
static int launch_cfft_t(bool inv, const float2 *in, float2 *out, const float2 *tw, int batch, float scale,
                         cudaStream_t st) {
  if constexpr (ThreadGeom<LOGN>::OK) {  // one thread per transform (N = 2, 4, 32)
    using G = ThreadGeom<LOGN>;
    const int grid = (batch + G::PER_CTA - 1) / G::PER_CTA;
    if (inv) {
      int rc = set_smem(fft_thread_kernel<LOGN, kThreadComplex, true>, G::SMEM_BYTES);
      if (rc) return rc;
      fft_thread_kernel<LOGN, kThreadComplex, true><<<grid, G::THREADS, G::SMEM_BYTES, st>>>((const float4 *)in, (float4 *)out, nullptr, batch, scale);
    } else {
      int rc = set_smem(fft_thread_kernel<LOGN, kThreadComplex, false>, G::SMEM_BYTES);
      if (rc) return rc;
      fft_thread_kernel<LOGN, kThreadComplex, false><<<grid, G::THREADS, G::SMEM_BYTES, st>>>((const float4 *)in, (float4 *)out, nullptr, batch, scale);
    }
  } else {
    using B = BatchGeom<LOGN>;
    const int grid = (batch + B::TPB - 1) / B::TPB;
    if (inv) {
      int rc = set_smem(cfft_kernel<LOGN, true>, B::SMEM_BYTES);
      if (rc) return rc;
      cfft_kernel<LOGN, true><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(in, out, tw, batch, scale);
    } else {
      int rc = set_smem(cfft_kernel<LOGN, false>, B::SMEM_BYTES);
      if (rc) return rc;
      cfft_kernel<LOGN, false><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(in, out, tw, batch, scale);
    }
  }
  CK(cudaGetLastError());
  return B2F_OK;
}
template <int LOGN>
static int launch_rfft_t(bool inv, const float2 *in, float2 *out, const float2 *tw, const float2 *hw, int batch,
                         float fwd_scale, cudaStream_t st, long long prefetch) {
  if constexpr (ThreadGeom<LOGN>::OK) {  // one thread per transform, split / unsplit in registers (N = 2, 4, 32)
    using G = ThreadGeom<LOGN>;
    const int grid = (batch + G::PER_CTA - 1) / G::PER_CTA;
    if (inv) {
      int rc = set_smem(fft_thread_kernel<LOGN, kThreadRealInv, true>, G::SMEM_BYTES);
      if (rc) return rc;
      fft_thread_kernel<LOGN, kThreadRealInv, true><<<grid, G::THREADS, G::SMEM_BYTES, st>>>((const float4 *)in, (float4 *)out, hw, batch, 1.0f);
    } else {
      int rc = set_smem(fft_thread_kernel<LOGN, kThreadRealFwd, false>, G::SMEM_BYTES);
      if (rc) return rc;
      fft_thread_kernel<LOGN, kThreadRealFwd, false><<<grid, G::THREADS, G::SMEM_BYTES, st>>>((const float4 *)in, (float4 *)out, hw, batch, fwd_scale);
    }
  } else if constexpr (RegSplitGeom<LOGN>::OK) {
    using B = BatchGeom<LOGN>;
    const int grid = (batch + B::TPB - 1) / B::TPB;
    // register-level split / unsplit (fft_kernels.cuh, second half)
    const int ahead = prefetch_distance<B>(prefetch);
    if (inv) {
      int rc = set_smem(rfft_inv_reg_kernel<LOGN>, B::SMEM_BYTES);
      if (rc) return rc;
      rfft_inv_reg_kernel<LOGN><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(in, out, tw, hw, batch, ahead);
    } else {
      int rc = set_smem(rfft_fwd_reg_kernel<LOGN>, B::SMEM_BYTES);
      if (rc) return rc;
      rfft_fwd_reg_kernel<LOGN><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(in, out, tw, hw, batch, fwd_scale, ahead);
    }
  } else {
    static_assert(LOGN < 0, "every size has a real-transform kernel");
  }
  CK(cudaGetLastError());
  return B2F_OK;
}

#define B2F_DISPATCH_LOGN(logn, CALL)  \
  switch (logn) {                      \
    case 1: return CALL(1);            \
    case 2: return CALL(2);            \
    case 3: return CALL(3);            \
    case 4: return CALL(4);            \
    case 5: return CALL(5);            \
    case 6: return CALL(6);            \
    case 7: return CALL(7);            \
    case 8: return CALL(8);            \
    case 9: return CALL(9);            \
    case 10: return CALL(10);          \
    case 11: return CALL(11);          \
    case 12: return CALL(12);          \
    case 13: return CALL(13);          \
    case 14: return CALL(14);          \
    default: return B2F_ERR_UNSUPPORTED; \
  }

static int launch_cfft(int logn, bool inv, const float2 *in, float2 *out, const float2 *tw, int batch, float scale,
                       cudaStream_t st) {
#define CALL(L) launch_cfft_t<L>(inv, in, out, tw, batch, scale, st)
  B2F_DISPATCH_LOGN(logn, CALL)
#undef CALL
}
static int launch_rfft(int logn, bool inv, const float2 *in, float2 *out, const float2 *tw, const float2 *hw, int batch,
                       float fwd_scale, cudaStream_t st, long long prefetch) {
#define CALL(L) launch_rfft_t<L>(inv, in, out, tw, hw, batch, fwd_scale, st, prefetch)
  B2F_DISPATCH_LOGN(logn, CALL)
#undef CALL
}


// ---- one-SM plan (fft_sm.cuh): N = 2^15 (and complex N = 2^14) in one pass over HBM, one persistent CTA per SM -----------
struct SmPlan {
  int grid = 0, logn = 0;
  float2 *d_twn = nullptr, *d_twp = nullptr, *d_twa = nullptr;
  bool ok() const { return d_twn != nullptr; }
  // A transform takes one SM ~25 us whatever the batch, so below ~100 transforms per call the four-step launch pair, which
  // spreads every transform over the whole GPU, is the faster one (measured crossover, tools/fft_sm_probe.py).
  // N = 2^14 runs two transforms per CTA iteration: twice the batch before the GPU is full.
  long long min_batch = 96;
  bool use_for(int batch) const { return ok() && min_batch > 0 && batch >= min_batch * (32 >> (logn - 10)); }
  int init(int device, int logn_) {
    logn = logn_;
    const int N = 1 << logn;
    std::vector<float2> twn(5 * 32), twp(5 * 32), twa(5 * 32);
    for (int b = 0; b < 5; b++)
      for (int j = 0; j < 32; j++) {
        twn[b * 32 + j] = ref_twiddle(((long long)j << b) % N, N);                       // W_N^(j 2^b)
        twp[b * 32 + j] = ref_twiddle(((long long)(32 * j) << b) % N, N);                // W_N^(32 j 2^b)
        twa[b * 32 + j] = ref_twiddle(((long long)(N / 1024) * j << b) % N, N);          // W_1024^(j 2^b)
      }
    int rc;
    if ((rc = upload(twn, &d_twn)) || (rc = upload(twp, &d_twp)) || (rc = upload(twa, &d_twa))) return rc;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    grid = nsm > 0 ? nsm : 1;
    return B2F_OK;
  }
  void destroy() {
    for (void *p : {(void *)d_twn, (void *)d_twp, (void *)d_twa})
      if (p) cudaFree(p);
    d_twn = d_twp = d_twa = nullptr;
  }
  template <bool INV, int KIND, int LOG1>
  int run_t(const float2 *in, float2 *out, const float2 *hw, int batch, float scale, cudaStream_t st) {
    auto kern = fft_sm_kernel<INV, KIND, LOG1>;
    int rc = set_smem(kern, SmGeom::SMEM);
    if (rc) return rc;
    const int units = (batch + (32 >> LOG1) - 1) / (32 >> LOG1);
    const int g = units < grid ? units : grid;
#if B2F_SMX_PDL
    // programmatic dependent launch: scheduled while the stream's previous kernel drains (see the kernel's prologue)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g);
    cfg.blockDim = dim3(SmGeom::THREADS);
    cfg.dynamicSmemBytes = SmGeom::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const float2 *a_twn = d_twn, *a_twp = d_twp, *a_twa = d_twa;
    CK(cudaLaunchKernelEx(&cfg, kern, in, out, a_twn, a_twp, a_twa, hw, batch, scale));
#else
    kern<<<g, SmGeom::THREADS, SmGeom::SMEM, st>>>(in, out, d_twn, d_twp, d_twa, hw, batch, scale);
    CK(cudaGetLastError());
#endif
    return B2F_OK;
  }
  template <bool INV, int KIND>
  int run(const float2 *in, float2 *out, const float2 *hw, int batch, float scale, cudaStream_t st) {
    if (logn == 14) return run_t<INV, KIND, 4>(in, out, hw, batch, scale, st);
    if constexpr (KIND != kSmRealFwd) {
      if (logn == 13) return run_t<INV, KIND, 3>(in, out, hw, batch, scale, st);  // (init() leaves forward real plans of 2^13 alone)
    }
    return run_t<INV, KIND, 5>(in, out, hw, batch, scale, st);
  }
};

// ---- four-step plan for N > 2^kMaxSmemLogN ----------------------------------------------------------------
struct LargePlan {
  int logn = 0, log1 = 0, log2 = 0, chunk = 1;
  float2 *d_tw1 = nullptr, *d_tw2 = nullptr, *d_twl = nullptr, *d_scratch = nullptr;
  // Scratch matrix between the two steps. Measured on B200 (1024 x 65536-point and 2048 x 32768-point batches):
  // cutting the batch into L2-sized chunks (32-96 MB: 2.2-2.7 TB/s), pipelining the chunks over two streams
  // (2.3 TB/s) and a single persistent kernel with ticketed column/row items and an L2-resident double buffer
  // (2.4 TB/s) are all SLOWER than long launches whose scratch simply goes through HBM (2.9 TB/s): short launches
  // run as one wave of CTAs in phase lockstep, long ones desynchronise and overlap loads, butterflies and
  // stores. 256 MB per launch pair is where that saturates (complex 2048 x 32768: 0.369 ms at 256 MB and at
  // 1 GiB; real: 0.434 vs 0.463 ms, the rows kernel walks its chunk backwards and finds the tail in L2).
  static constexpr size_t kScratchBytes = 256u << 20;
  Options opt;
  int init(int logn_, int max_batch, const Options &o) {
    opt = o;
    logn = logn_;
    log1 = logn / 2;
    log2 = logn - log1;
    const int N = 1 << logn, N2 = 1 << log2, N1 = 1 << log1;
    int rc;
    if ((rc = upload(make_pass_twiddles(log1), &d_tw1))) return rc;
    if ((rc = upload(make_pass_twiddles(log2), &d_tw2))) return rc;
    std::vector<float2> twl((size_t)N);
    for (int k1 = 0; k1 < N1; k1++)
      for (int n2 = 0; n2 < N2; n2++) twl[(size_t)k1 * N2 + n2] = ref_twiddle((long long)n2 * k1, N);
    if ((rc = upload(twl, &d_twl))) return rc;
    const size_t scratch_bytes = opt.large_chunk_mb > 0 ? (size_t)opt.large_chunk_mb << 20 : kScratchBytes;
    chunk = (int)(scratch_bytes / ((size_t)N * sizeof(float2)));
    if (chunk < 1) chunk = 1;
    if (chunk > max_batch) chunk = max_batch < 1 ? 1 : max_batch;
    return B2F_OK;  // the scratch matrix is allocated by the first launch pair that needs it (ensure_scratch)
  }
  int ensure_scratch() {
    if (!d_scratch) CK(cudaMalloc((void **)&d_scratch, (size_t)chunk * ((size_t)1 << logn) * sizeof(float2)));
    return B2F_OK;
  }
  void destroy() {
    for (void *p : {(void *)d_tw1, (void *)d_tw2, (void *)d_twl, (void *)d_scratch})
      if (p) cudaFree(p);
    d_tw1 = d_tw2 = d_twl = d_scratch = nullptr;
  }
  // UNSPLIT: inverse real transform, the unsplit fused into the columns kernel (hw = folded inverse split table)
  template <int L1, int L2, bool INV, bool REAL, int RBT, bool UNSPLIT = false>
  int run_tt(const float2 *in, float2 *out, int batch, float scale, cudaStream_t st, const float2 *hw) {
    using L = LargeGeom<L1, L2>;
    using R = RowsGeom<L1, L2, RBT>;
    static_assert(!UNSPLIT || (INV && !REAL), "the fused unsplit precedes an inverse complex transform");
    int rc;
    if ((rc = ensure_scratch())) return rc;
    if ((rc = set_smem(large_cols_kernel<L1, L2, INV>, L::SMEM_A))) return rc;
    if (UNSPLIT && (rc = set_smem(large_cols_unsplit_kernel<L1, L2>, L::SMEM_A))) return rc;
    if ((rc = set_smem(large_rows_kernel<L1, L2, INV, REAL, RBT>, R::SMEM))) return rc;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
      const int nb = batch - b0 < chunk ? batch - b0 : chunk;
      const int gxa = L::N2 / L::C, gxb = R::GROUPS;
      // ~4 x 148 CTAs of 256 threads (2 x 148 of 512) per launch, each looping over its share of the batch
      int gya = (592 + gxa - 1) / gxa, gyb = (592 * 256 / R::THREADS + gxb - 1) / gxb;
      if (gya > nb) gya = nb;
      if (gyb > nb) gyb = nb;
      const float2 *src = in + (size_t)b0 * L::N;
      float2 *dst = out + (size_t)b0 * L::N;
      if constexpr (UNSPLIT)
        large_cols_unsplit_kernel<L1, L2><<<dim3(gxa, gya), L::THREADS, L::SMEM_A, st>>>(src, d_scratch, d_tw1, d_twl, hw, nb);
      else
        large_cols_kernel<L1, L2, INV><<<dim3(gxa, gya), L::THREADS, L::SMEM_A, st>>>(src, d_scratch, d_tw1, d_twl, nb);
      CK(cudaGetLastError());
      large_rows_kernel<L1, L2, INV, REAL, RBT><<<dim3(gxb, gyb), R::THREADS, R::SMEM, st>>>(d_scratch, dst, d_tw2, hw, nb, scale);
      CK(cudaGetLastError());
    }
    return B2F_OK;
  }
  template <int L1, int L2, bool INV, bool REAL = false>
  int run_t(const float2 *in, float2 *out, int batch, float scale, cudaStream_t st, const float2 *hw = nullptr) {
    constexpr int RB = LargeGeom<L1, L2>::RB;
    if constexpr (REAL) {
      if (!opt.rows_rb16) return run_tt<L1, L2, INV, true, 2 * RB>(in, out, batch, scale, st, hw);
    }
    return run_tt<L1, L2, INV, REAL, RB>(in, out, batch, scale, st, hw);
  }
  int run_c2c(bool inv, const float2 *in, float2 *out, int batch, float scale, cudaStream_t st) {
    if (logn == 15) return inv ? run_t<7, 8, true>(in, out, batch, scale, st) : run_t<7, 8, false>(in, out, batch, scale, st);
    if (logn == 16) return inv ? run_t<8, 8, true>(in, out, batch, scale, st) : run_t<8, 8, false>(in, out, batch, scale, st);
    return B2F_ERR_UNSUPPORTED;
  }
  int run_real(bool inv, const float2 *in, float2 *out, const float2 *w2, const float2 *hw, int batch, float fwd_scale,
               cudaStream_t st) {
    const int N = 1 << logn;
    if (!inv && !opt.separate_split) {  // forward: split fused into the rows kernel
      if (logn == 15) return run_t<7, 8, false, true>(in, out, batch, fwd_scale, st, hw);
      if (logn == 16) return run_t<8, 8, false, true>(in, out, batch, fwd_scale, st, hw);
    }
    if (inv && !opt.separate_split) {  // inverse: unsplit fused into the columns kernel
      if (logn == 15) return run_tt<7, 8, true, false, LargeGeom<7, 8>::RB, true>(in, out, batch, 1.0f, st, hw);
      if (logn == 16) return run_tt<8, 8, true, false, LargeGeom<8, 8>::RB, true>(in, out, batch, 1.0f, st, hw);
    }
    const long long pairs = (long long)batch * (N / 2);
    const int grid = (int)((pairs + 255) / 256);
    int rc;
    if (!inv) {
      if ((rc = run_c2c(false, in, out, batch, fwd_scale, st))) return rc;
      rfft_split_kernel<false><<<grid, 256, 0, st>>>(out, out, w2, N, pairs);
      CK(cudaGetLastError());
    } else {
      rfft_split_kernel<true><<<grid, 256, 0, st>>>(in, out, w2, N, pairs);
      CK(cudaGetLastError());
      if ((rc = run_c2c(true, out, out, batch, 1.0f, st))) return rc;
    }
    return B2F_OK;
  }
};

// ---- pinned staging helper -----------------------------------------------------------------------------
// Host entry points move data through a pinned bounce buffer when it is small (latency path: one
// block / one transform), and straight from the caller's memory when it is large (the DMA engine
// handles pageable or caller-pinned memory itself).
struct Staging {
  void *pin = nullptr;
  size_t cap = 0;
  // grow-only, sized to what the handle actually moves (an application following the reference's usage
  // creates one object per channel: 64 objects must not pin 64 x 3 MB)
  int ensure(size_t bytes) {
    if (bytes <= cap) return B2F_OK;
    if (pin) cudaFreeHost(pin);
    pin = nullptr;
    cap = 0;
    const size_t want = (bytes + 4095) & ~(size_t)4095;
    CK(cudaMallocHost(&pin, want));
    cap = want;
    return B2F_OK;
  }
  void release() {
    if (pin) cudaFreeHost(pin);
    pin = nullptr;
    cap = 0;
  }
};
static const size_t kBounceMax = 1u << 20;

// Low-latency path of the synchronous host entry points: when a call moves at most this
// many bytes each way (one block of a few channels, one small transform -- what a Csound performance thread
// does every k-cycle), the kernel reads its input from and writes its result to the PINNED bounce buffers
// directly (they are device-accessible under unified addressing), so the call is: memcpy in, ONE launch,
// one stream synchronise, memcpy out -- no DMA copies to set up and wait for. The threshold is the `zerocopy_max` option (0 disables the path).
static int h2d(void *dst, const void *src, size_t bytes, Staging &sg, cudaStream_t st) {
  if (bytes <= kBounceMax) {
    int rc = sg.ensure(bytes);
    if (rc) return rc;
    memcpy(sg.pin, src, bytes);
    CK(cudaMemcpyAsync(dst, sg.pin, bytes, cudaMemcpyHostToDevice, st));
  } else {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
  }
  return B2F_OK;
}
// blocking: returns when dst holds the data
static int d2h(void *dst, const void *src, size_t bytes, Staging &sg, cudaStream_t st) {
  if (bytes <= kBounceMax) {
    int rc = sg.ensure(bytes);
    if (rc) return rc;
    CK(cudaMemcpyAsync(sg.pin, src, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(dst, sg.pin, bytes);
  } else {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return B2F_OK;
}

// =====================================================================================================
// complex / real FFT plans
// =====================================================================================================
struct FftPlanCore {
  int device = 0, N = 0, logn = 0, fwd = 1, max_batch = 1;
  bool unscaled = false;  // forward without the 1/N (what Clpconv's frames use, cl_conv_kernels.h:54-68)
  float fwd_scale() const { return (fwd && !unscaled) ? 1.0f / (float)N : 1.0f; }
  float2 *d_tw = nullptr;   // pass twiddles (small path) or sub-plan twiddles (large path)
  float2 *d_w2 = nullptr;   // split twiddles (real plans)
  float2 *d_hw = nullptr;   // folded split table 0.5*scale*i*w2 (forward) / its conjugate, unscaled (inverse)
  float2 *d_buf = nullptr;  // device buffer backing the host entry points
  LargePlan large;          // N > 2^kMaxSmemLogN
  SmPlan sm;                // N = 2^15: one pass over HBM, one transform per SM (fft_sm.cuh)
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: odd chunks of a pipelined host call
  Staging sg_in, sg_out;
  Options opt;
  bool is_large() const { return logn > kMaxSmemLogN; }
  int init(int dev, int n, int f, int mb, bool real) {
    device = dev, N = n, fwd = f ? 1 : 0, max_batch = mb < 1 ? 1 : mb;
    logn = ilog2_exact(n);
    opt = current_options();
    int rc = check_device(dev);
    if (rc) return rc;
    B2F_ON_DEVICE(dev);
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (is_large()) {
      rc = large.init(logn, max_batch, opt);
      if (rc) return rc;
    } else {
      rc = upload(make_pass_twiddles(logn), &d_tw);
      if (rc) return rc;
    }
    if ((logn == SmGeom::LOGN || logn == 14 || (logn == 13 && (!real || !fwd) && opt.fft_sm_8192)) && opt.fft_sm_min_batch > 0) {
      sm.min_batch = opt.fft_sm_min_batch;
      rc = sm.init(dev, logn);
      if (rc) return rc;
    }
    if (real) {
      std::vector<float2> w2 = make_split_twiddles(N), hw((size_t)N / 2 + 1);
      rc = upload(w2, &d_w2);
      if (rc) return rc;
      const float s = fwd_scale();  // powers of two: the folding is exact
      for (int i = 0; i <= N / 2 && i < N; i++) {
        // 0.5 * i * w2[i] = 0.5 * (-w.y, w.x); the inverse table is its conjugate (cl_fft.cpp:233-238 sign)
        hw[i].x = -0.5f * s * w2[i].y;
        hw[i].y = (fwd ? 0.5f : -0.5f) * s * w2[i].x;
      }
      rc = upload(hw, &d_hw);
      if (rc) return rc;
    }
    return B2F_OK;
  }
  int ensure_buf() {
    if (!d_buf) CK(cudaMalloc((void **)&d_buf, (size_t)max_batch * N * sizeof(float2)));
    return B2F_OK;
  }
  void destroy() {
    DeviceGuard guard(device);  // a failed create may carry an invalid ordinal: the guard swallows that
    if (d_tw) cudaFree(d_tw);
    if (d_w2) cudaFree(d_w2);
    if (d_hw) cudaFree(d_hw);
    if (d_buf) cudaFree(d_buf);
    large.destroy();
    sm.destroy();
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
    sg_in.release();
    sg_out.release();
  }
  // Large synchronous host calls (the reference's blocking write + kernels + blocking read, cl_fft.cpp:155-158, for a
  // whole batch): the batch is cut into chunks that alternate between two streams, so that one chunk's upload,
  // another's transform and a third's download overlap -- PCIe is full duplex and the copy engines are separate
  // (1024 x 65536 real points: 9.8 -> ~5 ms per call). Only for paths without plan-wide scratch: the four-step launch
  // pair shares one scratch matrix per plan.
  bool can_pipeline(int batch, int nchunks) const {
    if (!opt.pconv_pipeline || batch < 2 * nchunks) return false;
    return !is_large() || sm.use_for(batch / nchunks);
  }
  template <class Run>
  int host_pipelined(const void *src, void *dst, int batch, int nchunks, Run run) {
    if (!stream2) CK(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    const size_t per = (size_t)N * sizeof(float2);
    auto enqueue = [&]() -> int {
      for (int i = 0; i < nchunks; i++) {
        const int b0 = (int)((long long)i * batch / nchunks), nb = (int)((long long)(i + 1) * batch / nchunks) - b0;
        cudaStream_t st = (i & 1) ? stream2 : stream;
        float2 *d = d_buf + (size_t)b0 * N;
        CK(cudaMemcpyAsync(d, (const char *)src + b0 * per, nb * per, cudaMemcpyHostToDevice, st));
        int r = run(d, nb, st);
        if (r) return r;
        CK(cudaMemcpyAsync((char *)dst + b0 * per, d, nb * per, cudaMemcpyDeviceToHost, st));
      }
      return B2F_OK;
    };
    int rc = enqueue();
    // whatever happened, nothing may touch the caller's buffers after we return
    const cudaError_t e1 = cudaStreamSynchronize(stream), e2 = cudaStreamSynchronize(stream2);
    if (!rc && e1 != cudaSuccess) rc = cuda_fail(e1, "cudaStreamSynchronize");
    if (!rc && e2 != cudaSuccess) rc = cuda_fail(e2, "cudaStreamSynchronize");
    return rc;
  }
  int run_c2c(const float2 *in, float2 *out, int batch, cudaStream_t st) {
    const float scale = fwd ? 1.0f / (float)N : 1.0f;
    if (sm.use_for(batch))
      return fwd ? sm.run<false, kSmComplex>(in, out, nullptr, batch, scale, st)
                 : sm.run<true, kSmComplex>(in, out, nullptr, batch, scale, st);
    if (is_large()) return large.run_c2c(!fwd, in, out, batch, scale, st);
    return launch_cfft(logn, !fwd, in, out, d_tw, batch, scale, st);
  }
  int run_real(const float2 *in, float2 *out, int batch, cudaStream_t st) {
    if (sm.use_for(batch))
      return fwd ? sm.run<false, kSmRealFwd>(in, out, d_hw, batch, fwd_scale(), st)
                 : sm.run<true, kSmRealInv>(in, out, d_hw, batch, 1.0f, st);
    if (is_large()) return large.run_real(!fwd, in, out, d_w2, d_hw, batch, fwd_scale(), st);
    return launch_rfft(logn, !fwd, in, out, d_tw, d_hw, batch, fwd_scale(), st, opt.fft_prefetch);
  }
};

struct b2f_cfft {
  FftPlanCore core;
};
struct b2f_rfft {
  FftPlanCore core;
};

extern "C" int b2f_cfft_create(b2f_cfft **plan, int device, int N, int fwd, int max_batch) {
  if (!plan) return B2F_ERR_INVALID_VALUE;
  *plan = nullptr;
  const int logn = ilog2_exact(N);
  if (logn < 1) return B2F_ERR_INVALID_VALUE;
  if (logn > kMaxLogN) return B2F_ERR_UNSUPPORTED;
  b2f_cfft *p = new (std::nothrow) b2f_cfft;
  if (!p) return B2F_ERR_ALLOC;
  int rc = p->core.init(device, N, fwd, max_batch, false);
  if (rc) {
    p->core.destroy();
    delete p;
    return rc;
  }
  *plan = p;
  return B2F_OK;
}
extern "C" int b2f_cfft_destroy(b2f_cfft *plan) {
  if (!plan) return B2F_OK;
  plan->core.destroy();
  delete plan;
  return B2F_OK;
}
extern "C" int b2f_cfft_exec_dev(b2f_cfft *plan, const void *d_in, void *d_out, int batch, void *stream) {
  if (!plan || !d_in || !d_out || batch < 0 || !al16(d_in) || !al16(d_out)) return B2F_ERR_INVALID_VALUE;
  if (batch == 0) return B2F_OK;
  FftPlanCore &c = plan->core;
  B2F_ON_DEVICE(c.device);
  return c.run_c2c((const float2 *)d_in, (float2 *)d_out, batch, (cudaStream_t)stream);
}
extern "C" int b2f_cfft_exec_host(b2f_cfft *plan, float *cdata, int batch) {
  if (!plan || !cdata || batch < 0) return B2F_ERR_INVALID_VALUE;
  if (batch == 0) return B2F_OK;
  FftPlanCore &c = plan->core;
  if (batch > c.max_batch) return B2F_ERR_BATCH;
  B2F_ON_DEVICE(c.device);
  int rc = c.ensure_buf();
  if (rc) return rc;
  const size_t bytes = (size_t)batch * c.N * sizeof(float2);
  if (bytes <= (size_t)c.opt.zerocopy_max) {
    if ((rc = c.sg_in.ensure(bytes)) || (rc = c.sg_out.ensure(bytes))) return rc;
    memcpy(c.sg_in.pin, cdata, bytes);
    if ((rc = c.run_c2c((const float2 *)c.sg_in.pin, (float2 *)c.sg_out.pin, batch, c.stream))) return rc;
    CK(cudaStreamSynchronize(c.stream));
    memcpy(cdata, c.sg_out.pin, bytes);
    return B2F_OK;
  }
  if (bytes > kBounceMax && c.can_pipeline(batch, 8))
    return c.host_pipelined(cdata, cdata, batch, 8, [&](float2 *d, int nb, cudaStream_t st) { return c.run_c2c(d, d, nb, st); });
  if ((rc = h2d(c.d_buf, cdata, bytes, c.sg_in, c.stream))) return rc;
  if ((rc = c.run_c2c(c.d_buf, c.d_buf, batch, c.stream))) return rc;
  return d2h(cdata, c.d_buf, bytes, c.sg_out, c.stream);
}

extern "C" int b2f_rfft_create(b2f_rfft **plan, int device, int size, int fwd, int max_batch) {
  if (!plan) return B2F_ERR_INVALID_VALUE;
  *plan = nullptr;
  const int logs = ilog2_exact(size);
  if (logs < 2) return B2F_ERR_INVALID_VALUE;
  if (logs - 1 > kMaxLogN) return B2F_ERR_UNSUPPORTED;
  b2f_rfft *p = new (std::nothrow) b2f_rfft;
  if (!p) return B2F_ERR_ALLOC;
  int rc = p->core.init(device, size / 2, fwd, max_batch, true);
  if (rc) {
    p->core.destroy();
    delete p;
    return rc;
  }
  *plan = p;
  return B2F_OK;
}
extern "C" int b2f_rfft_destroy(b2f_rfft *plan) {
  if (!plan) return B2F_OK;
  plan->core.destroy();
  delete plan;
  return B2F_OK;
}
extern "C" int b2f_rfft_exec_dev(b2f_rfft *plan, const void *d_in, void *d_out, int batch, void *stream) {
  if (!plan || !d_in || !d_out || batch < 0 || !al16(d_in) || !al16(d_out)) return B2F_ERR_INVALID_VALUE;
  if (batch == 0) return B2F_OK;
  FftPlanCore &c = plan->core;
  B2F_ON_DEVICE(c.device);
  return c.run_real((const float2 *)d_in, (float2 *)d_out, batch, (cudaStream_t)stream);
}
extern "C" int b2f_rfft_exec_host(b2f_rfft *plan, float *cdata, float *r, int batch) {
  if (!plan || !cdata || !r || batch < 0) return B2F_ERR_INVALID_VALUE;
  if (batch == 0) return B2F_OK;
  FftPlanCore &c = plan->core;
  if (batch > c.max_batch) return B2F_ERR_BATCH;
  B2F_ON_DEVICE(c.device);
  int rc = c.ensure_buf();
  if (rc) return rc;
  const size_t bytes = (size_t)batch * c.N * sizeof(float2);
  // forward reads the reals (cl_fft.cpp:273-275), inverse reads the spectrum (284)
  const void *src = c.fwd ? (const void *)r : (const void *)cdata;
  if (bytes <= (size_t)c.opt.zerocopy_max) {
    if ((rc = c.sg_in.ensure(bytes)) || (rc = c.sg_out.ensure(bytes))) return rc;
    memcpy(c.sg_in.pin, src, bytes);
    if ((rc = c.run_real((const float2 *)c.sg_in.pin, (float2 *)c.sg_out.pin, batch, c.stream))) return rc;
    CK(cudaStreamSynchronize(c.stream));
    memcpy(cdata, c.sg_out.pin, bytes);                                    // 281 / 290
    if (!c.fwd && (void *)r != (void *)cdata) memcpy(r, cdata, bytes);     // 292-293
    return B2F_OK;
  }
  if (bytes > kBounceMax && c.can_pipeline(batch, 8)) {
    if ((rc = c.host_pipelined(src, cdata, batch, 8, [&](float2 *d, int nb, cudaStream_t st) { return c.run_real(d, d, nb, st); })))
      return rc;
    if (!c.fwd && (void *)r != (void *)cdata) memcpy(r, cdata, bytes);        // 292-293
    return B2F_OK;
  }
  if ((rc = h2d(c.d_buf, src, bytes, c.sg_in, c.stream))) return rc;
  if ((rc = c.run_real(c.d_buf, c.d_buf, batch, c.stream))) return rc;
  if ((rc = d2h(cdata, c.d_buf, bytes, c.sg_out, c.stream))) return rc;       // 281 / 290
  if (!c.fwd && (void *)r != (void *)cdata) memcpy(r, cdata, bytes);          // 292-293
  return B2F_OK;
}

// =====================================================================================================
// partitioned convolution
// =====================================================================================================
struct b2f_pconv {
  int device = 0, cvs = 0, pts = 0, logp = 0, nparts = 0, channels = 1;
  int wp = 0, wp2 = 0;  // ring positions, cl_conv.cpp:144
  float2 *d_fdl = nullptr, *d_irs = nullptr, *d_tw = nullptr, *d_w2 = nullptr;
  float2 *d_hw = nullptr;   // folded split table 0.5 i w2 (push_ir's register-level transform, pts >= 64)
  float *d_tail = nullptr, *d_in1 = nullptr, *d_in2 = nullptr, *d_out = nullptr, *d_ir = nullptr;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: second half of the channels in the pipelined host call
  Staging sg_in, sg_in2, sg_out;
  int cluster = 1;
  bool deep = false;        // whole-handle launches have at most one CTA per SM: TMA feed with the deep ring
  Options opt;
  int failed = 0;           // sticky: a multi-stream host call broke off half way, the state is not trustworthy
  // general path (pts > 2^kPconvMaxLogP): batched real-FFT plans + pad / MAC / overlap-add kernels; ring positions in
  // device memory (d_state), one CUDA graph per block kind for the host call
  FftPlanCore *gfwd = nullptr, *ginv = nullptr;
  float *d_pad = nullptr;   // [2][channels][2*pts] (second half: in2 of a time-varying block)
  float2 *d_Y = nullptr;    // [2][channels][pts]
  float2 *d_Ypart = nullptr;  // [ksplit][channels][pts]: partial sums of a MAC whose partitions are split over CTAs
  int ksplit = 1;
  int *d_state = nullptr;   // {wp, wp2}
  cudaGraphExec_t graph[2] = {nullptr, nullptr};  // [time-varying]: H2D, the block's launches, D2H on fixed buffers
  bool general() const { return gfwd != nullptr; }
  size_t ring_elems() const { return (size_t)channels * nparts * pts; }
  void destroy() {
    DeviceGuard guard(device);  // a failed create may carry an invalid ordinal: the guard swallows that
    for (cudaGraphExec_t g : graph)
      if (g) cudaGraphExecDestroy(g);
    for (void *p : {(void *)d_fdl, (void *)d_irs, (void *)d_tw, (void *)d_w2, (void *)d_hw, (void *)d_tail,
                    (void *)d_in1, (void *)d_in2, (void *)d_out, (void *)d_ir, (void *)d_pad, (void *)d_Y, (void *)d_Ypart, (void *)d_state})
      if (p) cudaFree(p);
    for (FftPlanCore *p : {gfwd, ginv})
      if (p) {
        p->destroy();
        delete p;
      }
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
    sg_in.release();
    sg_in2.release();
    sg_out.release();
  }
};

template <int LOGP>
static int pconv_smem_bytes(bool tv, bool tma, int S, bool deep = false) {
  using P = PconvGeom<LOGP>;
  const int fft = P::FFT_SMEM + (P::FFT_SMEM & 1);
  const int partial4 = P::partial_f4(tma, S);
  const int ring = tma ? P::ring_f4(deep) * (int)sizeof(float4) + 2 * P::stages(deep) * 8 : 0;
  return (tv ? 2 : 1) * fft * (int)sizeof(float2) + partial4 * (int)sizeof(float4) + ring;
}

template <int LOGP, bool TV, bool TMA, bool DEEP = false>
static int launch_pconv_step_tt(const PconvArgs &a, int channels, int S, cudaStream_t st) {
  using P = PconvGeom<LOGP>;
  const int smem = pconv_smem_bytes<LOGP>(TV, TMA, S, DEEP);
  int rc = set_smem(pconv_step_kernel<LOGP, TV, TMA, DEEP>, smem);
  if (rc) return rc;
  if (S > 8 && (rc = allow_cluster16_once((const void *)pconv_step_kernel<LOGP, TV, TMA, DEEP>))) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, channels, 1);
  cfg.blockDim = dim3(P::NTHREADS + (TMA ? 32 : 0), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, pconv_step_kernel<LOGP, TV, TMA, DEEP>, a));
  return B2F_OK;
}
template <int LOGP, bool TV>
static int launch_pconv_step_t(const PconvArgs &a, int channels, int S, int tma_opt, cudaStream_t st, bool deep = false) {
  // At most one CTA per SM, partitions of 2048 / 4096 samples, long streams per CTA: the TMA feed with whole-frame stages (PconvGeom::deep_stages)
  if constexpr (LOGP >= 11) {
    if (deep && tma_opt != 0) return launch_pconv_step_tt<LOGP, TV, true, true>(a, channels, S, st);
  }
  // Which MAC feeds the fused kernel: registers (128-bit loads, 16 in flight per thread) or the TMA ring. Measured on
  // B200 (480000 taps, fraction of the measured HBM peak, registers vs TMA; tools/pconv_sweep.py --feed-sweep,
  // profiles/r02_pconv_feed_sweep.txt): pts 512 1.10 vs 1.05; pts 1024 0.47 / 0.76 / 0.85 / 1.02 vs 0.40 / 0.66 /
  // 0.75 / 0.95 at 16 / 64 / 256 / 1024 channels; pts 2048 0.29 / 0.53 / 0.71 / 0.84 vs 0.40 / 0.76 / 0.97 / 0.97;
  // pts 4096 0.21 / 0.44 / 0.63 / 0.63 vs 0.23 / 0.46 / 0.68 / 0.78 -- the TMA feed walks every frame front to back
  // (partition-major), the register feed of an 8-tile frame cannot. Option pconv_tma = 0 | 1 forces one or the other.
  const bool use_tma = tma_opt >= 0 ? tma_opt == 1 : LOGP >= 11;
  return use_tma ? launch_pconv_step_tt<LOGP, TV, true>(a, channels, S, st)
                 : launch_pconv_step_tt<LOGP, TV, false>(a, channels, S, st);
}
// Can the device co-schedule a cluster of S CTAs of the step kernel this handle would launch? (Clusters of 16 need
// 16 free SMs of ONE GPC at up to 198 KB of shared memory each: true on a full B200, not guaranteed on every part.)
template <int LOGP, bool TV, bool TMA, bool DEEP>
static bool pconv_cluster_fits_tt(int S) {
  using P = PconvGeom<LOGP>;
  auto kern = pconv_step_kernel<LOGP, TV, TMA, DEEP>;
  const int smem = pconv_smem_bytes<LOGP>(TV, TMA, S, DEEP);
  if (set_smem(kern, smem)) return false;
  if (S > 8 && allow_cluster16_once((const void *)kern)) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, 1, 1);
  cfg.blockDim = dim3(P::NTHREADS + (TMA ? 32 : 0), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return n >= 1;
}
template <int LOGP>
static int pconv_cluster_fits_t(int S, int tma_opt, bool deep) {  // both block kinds (static and time-varying)
  if constexpr (LOGP >= 11) {
    if (deep && tma_opt != 0) return pconv_cluster_fits_tt<LOGP, false, true, true>(S) && pconv_cluster_fits_tt<LOGP, true, true, true>(S);
  }
  const bool use_tma = tma_opt >= 0 ? tma_opt == 1 : LOGP >= 11;
  return use_tma ? pconv_cluster_fits_tt<LOGP, false, true, false>(S) && pconv_cluster_fits_tt<LOGP, true, true, false>(S)
                 : pconv_cluster_fits_tt<LOGP, false, false, false>(S) && pconv_cluster_fits_tt<LOGP, true, false, false>(S);
}
template <int LOGP>
static int launch_pconv_push_t(const float *ir, size_t stride, b2f_pconv *h, cudaStream_t st) {
  if constexpr (RegSplitGeom<LOGP>::OK) {
    if (h->opt.pconv_push_reg) {
      using B = BatchGeom<LOGP>;
      int rc = set_smem(pconv_push_ir_reg_kernel<LOGP>, B::SMEM_BYTES);
      if (rc) return rc;
      dim3 grid((h->nparts + B::TPB - 1) / B::TPB, h->channels, 1);
      pconv_push_ir_reg_kernel<LOGP><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(ir, stride, h->d_irs, h->d_tw, h->d_hw, h->nparts, h->wp2);
      CK(cudaGetLastError());
      return B2F_OK;
    }
  }
  using Q = PushGeom<LOGP>;
  int rc = set_smem(pconv_push_ir_kernel<LOGP>, Q::SMEM_BYTES);
  if (rc) return rc;
  dim3 grid((h->nparts + Q::GROUPS - 1) / Q::GROUPS, h->channels, 1);
  pconv_push_ir_kernel<LOGP><<<grid, Q::THREADS, Q::SMEM_BYTES, st>>>(ir, stride, h->d_irs, h->d_tw, h->d_w2, h->nparts, h->wp2);
  CK(cudaGetLastError());
  return B2F_OK;
}

#define B2F_DISPATCH_LOGP(logp, CALL)    \
  switch (logp) {                        \
    case 1: return CALL(1);              \
    case 2: return CALL(2);              \
    case 3: return CALL(3);              \
    case 4: return CALL(4);              \
    case 5: return CALL(5);              \
    case 6: return CALL(6);              \
    case 7: return CALL(7);              \
    case 8: return CALL(8);              \
    case 9: return CALL(9);              \
    case 10: return CALL(10);            \
    case 11: return CALL(11);            \
    case 12: return CALL(12);            \
    default: return B2F_ERR_UNSUPPORTED; \
  }

static int launch_pconv_step(int logp, bool tv, const PconvArgs &a, int channels, int S, int tma_opt, cudaStream_t st,
                             bool deep = false) {
  if (tv) {
#define CALL(L) launch_pconv_step_t<L, true>(a, channels, S, tma_opt, st, deep)
    B2F_DISPATCH_LOGP(logp, CALL)
#undef CALL
  } else {
#define CALL(L) launch_pconv_step_t<L, false>(a, channels, S, tma_opt, st, deep)
    B2F_DISPATCH_LOGP(logp, CALL)
#undef CALL
  }
}
static int pconv_cluster_fits(int logp, int S, int tma_opt, bool deep) {
#define CALL(L) pconv_cluster_fits_t<L>(S, tma_opt, deep)
  B2F_DISPATCH_LOGP(logp, CALL)
#undef CALL
}
static int launch_pconv_push(int logp, const float *ir, size_t stride, b2f_pconv *h, cudaStream_t st) {
#define CALL(L) launch_pconv_push_t<L>(ir, stride, h, st)
  B2F_DISPATCH_LOGP(logp, CALL)
#undef CALL
}

extern "C" int b2f_pconv_create(b2f_pconv **out, int device, int cvs, int pts, int channels) {
  if (!out) return B2F_ERR_INVALID_VALUE;
  *out = nullptr;
  const int logp = ilog2_exact(pts);
  if (logp < 1 || cvs < pts || channels < 1) return B2F_ERR_INVALID_VALUE;
  if (logp > 15 || channels > 65535) return B2F_ERR_UNSUPPORTED;  // frame = pts complex points <= 32768
  int rc = check_device(device);
  if (rc) return rc;
  B2F_ON_DEVICE(device);
  b2f_pconv *h = new (std::nothrow) b2f_pconv;
  if (!h) return B2F_ERR_ALLOC;
  h->device = device, h->cvs = cvs, h->pts = pts, h->logp = logp, h->channels = channels;
  h->opt = current_options();
  h->nparts = cvs / pts;  // truncating, cl_conv.cpp:143
  h->wp = 0;
  h->wp2 = h->nparts - 1;
  // Cluster split of the partitions (portable cluster limit 8), from measurements on B200 (tools/pconv_sweep.py
  // --cluster-sweep, profiles/r02_pconv_cluster_sweep.txt; fractions of the measured HBM peak). Partitions up to 1024
  // samples: enough CTAs for ~3.5 per SM (64 ch x 937 partitions: S=8 0.98 vs 0.94 at S=4 and 0.47 at S=1; 128 ch:
  // S=4 1.04 vs 1.02 at S=2 -- the regime of the 8-GPU strong split; 256 ch x 187: S=2 0.97 vs 0.96). Longer
  // partitions: one CTA per SM-slot is enough and more only costs (256 ch x 4096: S=1 0.94, S=2 0.88, S=8 0.73; 64 ch x
  // 4096: S=4 0.71, S=2 0.56, S=8 0.56). Many channels with long IRs: up to ~2048 CTAs while every CTA keeps >= 256
  // partitions to stream (1024 ch x 937: S=2 is 3 % faster than S=1 or 4).
  int S = 1;
  const int target_ctas = pts <= 1024 ? 512 : 148;
  while (S < 8 && channels * S < target_ctas && S * 2 <= h->nparts) S *= 2;
  while (S < 8 && channels * S < 2048 && h->nparts / (2 * S) >= 256) S *= 2;
  // A handful of channels with a long IR (the mono rows of csound/tests.py's grid): every CTA streams at what ONE SM
  // can pull (~60-100 GB/s), so the step time is nparts / S; a cluster of 16 (beyond the portable limit, one per GPC
  // at a time) halves it once more.
  // Measured, mono (tools/rt_ratio_grid.py): pts 512 x 8192 partitions 147 -> 82 us per block, pts 2048 x 2048: 256 -> 136;
  // nothing to gain below ~8 MB of rings per channel (pts 512 x 1024: 31.0 vs 29.5 us).
  // More channels at pts <= 1024 (register feed; 8 channels x 8192 partitions of 512: 198 -> 109 us; 8 channels x 2048 of
  // 2048: 121 us with 8 CTAs per cluster, 139 with 16; tools/pconv_cluster16_probe.py).
  // 16 channels x 937 partitions of 512: 30.2 -> 25.1 us, x 468 of 1024: 39.2 -> 31.0; 32 channels: 41 -> 54 us, so four times
  // the channel limit at pts <= 1024 and no further (tools/pconv_cluster16_probe2.py).
  const bool small = pts <= 1024;
  const long long max16 = h->opt.pconv_cluster16_max_channels * (small ? 4 : 1);
  if (S == 8 && channels <= max16 && (long long)h->nparts * pts >= (small ? (1 << 18) : (1 << 19)) && h->nparts >= 64) S = 16;
  auto fail = [&](int code) {
    h->destroy();
    delete h;
    return code;
  };
  if (h->opt.pconv_cluster) {  // forced split: a power of two within the portable cluster size, at most one CTA per partition
    const long long f = h->opt.pconv_cluster;
    if ((f != 1 && f != 2 && f != 4 && f != 8 && f != 16) || f > h->nparts) return fail(B2F_ERR_INVALID_VALUE);
    S = (int)f;
  }
  h->cluster = S;
  // Measured (tools/pconv_few_channels_probe.py, profiles/r02_pconv_few_channels.txt): mono, pts 2048 x 2048 partitions
  // 121 -> 55 us per block, x 1024: 65.9 -> 31.0, x 512: 39.2 -> 22.7 (tools/pconv_deep_threshold_probe.py); nothing to gain at
  // 16 channels x 234 (45.3 vs 43.8 us); a loss at pts 1024 (register feed: 19.6 vs 24.8 us)
  h->deep = h->opt.pconv_deep_ring != 0 && logp >= 11 && (long long)channels * S <= sm_count() && h->nparts / S >= h->opt.pconv_deep_min_parts;
  if (S == 16 && !h->opt.pconv_cluster && logp <= kPconvMaxLogP && pconv_cluster_fits(logp, 16, (int)h->opt.pconv_tma, h->deep) != 1) {
    // this device cannot co-schedule 16 such CTAs: the portable limit it is
    S = h->cluster = 8;
    h->deep = h->opt.pconv_deep_ring != 0 && logp >= 11 && (long long)channels * S <= sm_count() && h->nparts / S >= h->opt.pconv_deep_min_parts;
  }
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cuda_fail(e, "stream"));
  if (logp <= kPconvMaxLogP) {
    if ((rc = upload(make_pass_twiddles(logp), &h->d_tw))) return fail(rc);
    if ((rc = upload(make_split_twiddles(pts), &h->d_w2))) return fail(rc);
    {
      const std::vector<float2> w2 = make_split_twiddles(pts);
      std::vector<float2> hw((size_t)pts / 2 + 1);
      for (int i = 0; i <= pts / 2 && i < pts; i++) hw[i] = make_float2(-0.5f * w2[i].y, 0.5f * w2[i].x);  // 0.5 i w2[i]
      if ((rc = upload(hw, &h->d_hw))) return fail(rc);
    }
  } else {
    h->gfwd = new (std::nothrow) FftPlanCore;
    h->ginv = new (std::nothrow) FftPlanCore;
    if (!h->gfwd || !h->ginv) return fail(B2F_ERR_ALLOC);
    h->gfwd->unscaled = true;  // Clpconv's frames are never scaled (cl_conv_kernels.h:54-68)
    if ((rc = h->gfwd->init(device, pts, 1, 2 * channels, true))) return fail(rc);  // (time-varying blocks: both inputs in one batch)
    if ((rc = h->ginv->init(device, pts, 0, channels, true))) return fail(rc);
    if ((e = cudaMalloc((void **)&h->d_pad, (size_t)2 * channels * 2 * pts * sizeof(float))) != cudaSuccess)
      return fail(cuda_fail(e, "cudaMalloc pad"));
    if ((e = cudaMalloc((void **)&h->d_Y, (size_t)2 * channels * pts * sizeof(float2))) != cudaSuccess)
      return fail(cuda_fail(e, "cudaMalloc Y"));
    {
      // enough MAC CTAs for two per SM, every CTA keeping at least 8 partitions (measured, mono, pts 8192: 512
      // partitions 170 -> see profiles/r02_rt_ratio_grid.txt)
      const long long tiles = (long long)(pts / kMacTileBins) * channels;
      long long K = h->opt.pconv_ksplit > 0 ? h->opt.pconv_ksplit : (2LL * sm_count() + tiles - 1) / tiles;
      if (K > h->nparts / 8) K = h->nparts / 8;
      if (K > 64) K = 64;
      if (K < 1 || h->opt.pconv_ksplit < 0) K = 1;
      h->ksplit = (int)K;
      if (K > 1 && (e = cudaMalloc((void **)&h->d_Ypart, (size_t)K * channels * pts * sizeof(float2))) != cudaSuccess)
        return fail(cuda_fail(e, "cudaMalloc Ypart"));
    }
    if ((e = cudaMalloc((void **)&h->d_state, 2 * sizeof(int))) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc state"));
    const int st0[2] = {h->wp, h->wp2};
    if ((e = cudaMemcpy(h->d_state, st0, sizeof(st0), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(cuda_fail(e, "state"));
    // nothing may be allocated while a block's launches are being captured into a graph
    for (FftPlanCore *pl : {h->gfwd, h->ginv})
      if (pl->is_large() && (rc = pl->large.ensure_scratch())) return fail(rc);
  }
  const size_t ring = h->ring_elems() * sizeof(float2), blk = (size_t)channels * pts * sizeof(float);
  if ((e = cudaMalloc((void **)&h->d_fdl, ring)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc fdl"));
  if ((e = cudaMalloc((void **)&h->d_irs, ring)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc irs"));
  if ((e = cudaMalloc((void **)&h->d_tail, blk)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc tail"));
  // zero state, cl_conv.cpp:303-313
  if ((e = cudaMemsetAsync(h->d_fdl, 0, ring, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  if ((e = cudaMemsetAsync(h->d_irs, 0, ring, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  if ((e = cudaMemsetAsync(h->d_tail, 0, blk, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(cuda_fail(e, "sync"));
  *out = h;
  return B2F_OK;
}
extern "C" int b2f_pconv_destroy(b2f_pconv *h) {
  if (!h) return B2F_OK;
  h->destroy();
  delete h;
  return B2F_OK;
}
extern "C" int b2f_pconv_nparts(const b2f_pconv *h) { return h ? h->nparts : 0; }

extern "C" int b2f_pconv_reset(b2f_pconv *h) {
  if (!h) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  CK(cudaMemsetAsync(h->d_fdl, 0, h->ring_elems() * sizeof(float2), h->stream));
  CK(cudaMemsetAsync(h->d_tail, 0, (size_t)h->channels * h->pts * sizeof(float), h->stream));
  h->wp = 0;
  h->wp2 = h->nparts - 1;
  if (h->d_state) {
    const int st0[2] = {h->wp, h->wp2};
    CK(cudaMemcpyAsync(h->d_state, st0, sizeof(st0), cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  if (h->stream2) CK(cudaStreamSynchronize(h->stream2));
  h->failed = 0;
  return B2F_OK;
}

// general path: R(x) of one block per channel (x: [channels] rows of `stride` floats) into ring frame `frame`
static int pconv_general_frame(b2f_pconv *h, const float *x, size_t stride, float2 *ring, int frame, cudaStream_t st) {
  const int pts = h->pts;
  dim3 grid((2 * pts + 255) / 256, h->channels, 1);
  pconv_pad_kernel<<<grid, 256, 0, st>>>(x, stride, h->d_pad, pts);
  CK(cudaGetLastError());
  int rc = h->gfwd->run_real((const float2 *)h->d_pad, h->d_Y, h->channels, st);
  if (rc) return rc;
  CK(cudaMemcpy2DAsync(ring + (size_t)frame * pts, (size_t)h->nparts * pts * sizeof(float2), h->d_Y,
                       (size_t)pts * sizeof(float2), (size_t)pts * sizeof(float2), h->channels,
                       cudaMemcpyDeviceToDevice, st));
  return B2F_OK;
}
// pts = 8192 / 16384 with fewer channels than the one-SM FFT kernel wants: the general path's transforms run on the
// register-level kernels fused with their neighbours (pconv_kernels.cuh: frames, inverse + overlap-add + advance, and
// push_ir in one launch) -- 11-12 launches per time-varying block become 3-4
static bool pconv_general_fused(const b2f_pconv *h) {
  return h->opt.pconv_general_fused != 0 && (h->logp == 13 || h->logp == 14) && !h->gfwd->sm.use_for(h->channels);
}
template <int LOGP>
static int pconv_general_push_fused_t(b2f_pconv *h, const float *ir, size_t stride, cudaStream_t st) {
  using B = BatchGeom<LOGP>;
  int rc = set_smem(pconv_push_ir_reg_kernel<LOGP>, B::SMEM_BYTES);
  if (rc) return rc;
  dim3 grid((h->nparts + B::TPB - 1) / B::TPB, h->channels, 1);
  pconv_push_ir_reg_kernel<LOGP><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(ir, stride, h->d_irs, h->gfwd->d_tw, h->gfwd->d_hw, h->nparts, h->wp2);
  CK(cudaGetLastError());
  return B2F_OK;
}
template <int LOGP>
static int pconv_general_frames_fused_t(b2f_pconv *h, bool tv, const float *d_in1, const float *d_in2, cudaStream_t st) {
  using B = BatchGeom<LOGP>;
  int rc = set_smem(pconv_frames_reg_kernel<LOGP>, B::SMEM_BYTES);
  if (rc) return rc;
  dim3 grid(h->channels, tv ? 2 : 1, 1);
  pconv_frames_reg_kernel<LOGP><<<grid, B::THREADS, B::SMEM_BYTES, st>>>(d_in1, d_in2, (size_t)h->pts, h->d_fdl, h->d_irs, h->gfwd->d_tw,
                                                                       h->gfwd->d_hw, h->nparts, h->d_state);
  CK(cudaGetLastError());
  return B2F_OK;
}
template <int LOGP>
static int pconv_general_inverse_fused_t(b2f_pconv *h, bool tv, float *d_out, cudaStream_t st) {
  using B = BatchGeom<LOGP>;
  int rc = set_smem(pconv_inverse_ola_kernel<LOGP>, B::SMEM_BYTES);
  if (rc) return rc;
  pconv_inverse_ola_kernel<LOGP><<<h->channels, B::THREADS, B::SMEM_BYTES, st>>>(h->d_Y, h->d_tail, d_out, h->ginv->d_tw, h->ginv->d_hw,
                                                                              h->d_state, h->nparts, tv ? 1 : 0);
  CK(cudaGetLastError());
  return B2F_OK;
}
static int pconv_general_push(b2f_pconv *h, const float *ir, size_t stride, cudaStream_t st) {
  if (pconv_general_fused(h))
    return h->logp == 13 ? pconv_general_push_fused_t<13>(h, ir, stride, st) : pconv_general_push_fused_t<14>(h, ir, stride, st);
  for (int i = 0; i < h->nparts; i++) {
    int frame = (h->wp2 - i) % h->nparts;
    if (frame < 0) frame += h->nparts;
    int rc = pconv_general_frame(h, ir + (size_t)i * h->pts, stride, h->d_irs, frame, st);
    if (rc) return rc;
  }
  return B2F_OK;
}
// R(x) of one block per channel into the ring frame named by the device-side position `which` (0: wp, 1: wp2)
static int pconv_general_frame_state(b2f_pconv *h, const float *x, float2 *ring, int which, cudaStream_t st) {
  const int pts = h->pts;
  dim3 grid((2 * pts + 255) / 256, h->channels, 1);
  pconv_pad_kernel<<<grid, 256, 0, st>>>(x, (size_t)pts, h->d_pad, pts);
  CK(cudaGetLastError());
  int rc = h->gfwd->run_real((const float2 *)h->d_pad, h->d_Y, h->channels, st);
  if (rc) return rc;
  dim3 gs((pts / 2 + 255) / 256, h->channels, 1);
  pconv_ring_store_kernel<<<gs, 256, 0, st>>>(h->d_Y, ring, pts, h->nparts, h->d_state, which);
  CK(cudaGetLastError());
  return B2F_OK;
}
// One block on the general path. Every launch takes the ring positions from h->d_state, the last one advances
// them: the sequence is the same for every block, which is what lets the host call replay it as a CUDA graph.
static int pconv_general_step(b2f_pconv *h, bool tv, float *d_out, const float *d_in1, const float *d_in2, cudaStream_t st) {
  const int pts = h->pts;
  const bool fused = pconv_general_fused(h);
  int rc;
  if (fused) {
    rc = h->logp == 13 ? pconv_general_frames_fused_t<13>(h, tv, d_in1, d_in2, st) : pconv_general_frames_fused_t<14>(h, tv, d_in1, d_in2, st);
    if (rc) return rc;
  } else {
    if (tv) {
      // both inputs as ONE batch of 2 x channels transforms: pad, batched rFFT, frame stores
      dim3 gp((2 * pts + 255) / 256, 2 * h->channels, 1);
      pconv_pad2_kernel<<<gp, 256, 0, st>>>(d_in1, d_in2, (size_t)pts, h->d_pad, pts, h->channels);
      CK(cudaGetLastError());
      if ((rc = h->gfwd->run_real((const float2 *)h->d_pad, h->d_Y, 2 * h->channels, st))) return rc;
      dim3 gs((pts / 2 + 255) / 256, 2 * h->channels, 1);
      pconv_ring_store2_kernel<<<gs, 256, 0, st>>>(h->d_Y, h->d_fdl, h->d_irs, pts, h->nparts, h->d_state, h->channels);
      CK(cudaGetLastError());
    } else if ((rc = pconv_general_frame_state(h, d_in1, h->d_fdl, 0, st))) {
      return rc;
    }
  }
  // TMA-fed MAC by default on this path (measured 3-16 % faster with many channels, 2x for a mono 4M-tap IR);
  // option pconv_tma = 0 selects the register-fed kernel.
  // Few channels: the partitions are split over grid.z (ksplit CTAs per tile, partial sums in d_Ypart, added in
  // ascending order by a second launch), so that a mono stream with a long IR still covers the GPU
  const int K = h->ksplit;
  float2 *macY = K > 1 ? h->d_Ypart : h->d_Y;
  if (h->opt.pconv_tma != 0) {
    if ((rc = set_smem(pconv_mac_tma_kernel, kMacTmaSmem))) return rc;
    dim3 gt(pts / kMacTileBins, h->channels, K);
    pconv_mac_tma_kernel<<<gt, 288, kMacTmaSmem, st>>>(h->d_fdl, h->d_irs, macY, pts, h->nparts, h->d_state);
  } else {
    dim3 gm((pts / 2 + 255) / 256, h->channels, K);
    pconv_mac_kernel<<<gm, 256, 0, st>>>(h->d_fdl, h->d_irs, macY, pts, h->nparts, h->d_state);
  }
  CK(cudaGetLastError());
  if (K > 1) {
    const size_t total4 = (size_t)h->channels * pts / 2;
    pconv_mac_sum_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>((const float4 *)h->d_Ypart, (float4 *)h->d_Y, total4, K);
    CK(cudaGetLastError());
  }
  if (fused)
    return h->logp == 13 ? pconv_general_inverse_fused_t<13>(h, tv, d_out, st) : pconv_general_inverse_fused_t<14>(h, tv, d_out, st);
  if ((rc = h->ginv->run_real(h->d_Y, h->d_Y, h->channels, st))) return rc;
  dim3 go((pts + 255) / 256, h->channels, 1);
  pconv_ola_kernel<<<go, 256, 0, st>>>((const float *)h->d_Y, h->d_tail, d_out, pts);
  CK(cudaGetLastError());
  pconv_advance_kernel<<<1, 32, 0, st>>>(h->d_state, h->nparts, tv ? 1 : 0);
  CK(cudaGetLastError());
  return B2F_OK;
}

extern "C" int b2f_pconv_push_ir_dev(b2f_pconv *h, const void *d_ir, size_t ir_stride, void *stream) {
  if (!h || !d_ir || ir_stride < (size_t)h->nparts * h->pts) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  if (h->general()) return pconv_general_push(h, (const float *)d_ir, ir_stride, (cudaStream_t)stream);
  int rc = launch_pconv_push(h->logp, (const float *)d_ir, ir_stride, h, (cudaStream_t)stream);
  // after nparts decrements the write position is back where it started (cl_conv.cpp:385)
  return rc;
}
extern "C" int b2f_pconv_push_ir_host(b2f_pconv *h, const float *ir, size_t ir_stride) {
  if (!h || !ir || ir_stride < (size_t)h->nparts * h->pts) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  const size_t per = (size_t)h->nparts * h->pts;
  // the time-domain IRs are needed on the device only for the duration of this call (1.96 GB at 1024 x 937 x 512)
  CK(cudaMalloc((void **)&h->d_ir, (size_t)h->channels * per * sizeof(float)));
  int rc = B2F_OK;
  cudaError_t e = cudaMemcpy2DAsync(h->d_ir, per * sizeof(float), ir, ir_stride * sizeof(float), per * sizeof(float),
                                    h->channels, cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy2DAsync ir");
  if (!rc) rc = h->general() ? pconv_general_push(h, h->d_ir, per, h->stream) : launch_pconv_push(h->logp, h->d_ir, per, h, h->stream);
  e = cudaStreamSynchronize(h->stream);
  if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
  cudaFree(h->d_ir);
  h->d_ir = nullptr;
  return rc;
}

static int pconv_enqueue(b2f_pconv *h, bool tv, float *d_out, const float *d_in1, const float *d_in2, cudaStream_t st) {
  PconvArgs a;
  a.fdl = h->d_fdl, a.irs = h->d_irs, a.tail = h->d_tail;
  a.in1 = d_in1, a.in2 = d_in2, a.out = d_out;
  a.tw = h->d_tw, a.w2 = h->d_w2;
  a.nparts = h->nparts, a.wp = h->wp, a.wp2 = h->wp2;
  int rc = h->general() ? pconv_general_step(h, tv, d_out, d_in1, d_in2, st)
                        : launch_pconv_step(h->logp, tv, a, h->channels, h->cluster, (int)h->opt.pconv_tma, st, h->deep);
  if (rc) return rc;
  h->wp = h->wp != h->nparts - 1 ? h->wp + 1 : 0;             // cl_conv.cpp:424
  if (tv) h->wp2 = h->wp2 == 0 ? h->nparts - 1 : h->wp2 - 1;  // cl_conv.cpp:519
  return B2F_OK;
}
// fused path, static IR: one step for channels [c0, c0 + n) only (ring positions are NOT advanced here)
static int pconv_launch_range(b2f_pconv *h, float *d_out, const float *d_in, int c0, int n, cudaStream_t st) {
  const size_t ring = (size_t)c0 * h->nparts * h->pts, blk = (size_t)c0 * h->pts;
  PconvArgs a;
  a.fdl = h->d_fdl + ring, a.irs = h->d_irs + ring, a.tail = h->d_tail + blk;
  a.in1 = d_in + blk, a.in2 = nullptr, a.out = d_out + blk;
  a.tw = h->d_tw, a.w2 = h->d_w2;
  a.nparts = h->nparts, a.wp = h->wp, a.wp2 = h->wp2;
  return launch_pconv_step(h->logp, false, a, n, h->cluster, (int)h->opt.pconv_tma, st);
}
extern "C" int b2f_pconv_process_dev(b2f_pconv *h, void *d_out, const void *d_in, void *stream) {
  if (!h || !d_out || !d_in || !al16(d_out) || !al16(d_in)) return B2F_ERR_INVALID_VALUE;
  if (h->failed) return h->failed;
  B2F_ON_DEVICE(h->device);
  return pconv_enqueue(h, false, (float *)d_out, (const float *)d_in, nullptr, (cudaStream_t)stream);
}
extern "C" int b2f_pconv_process_tv_dev(b2f_pconv *h, void *d_out, const void *d_in1, const void *d_in2, void *stream) {
  if (!h || !d_out || !d_in1 || !d_in2 || !al16(d_out) || !al16(d_in1) || !al16(d_in2)) return B2F_ERR_INVALID_VALUE;
  if (h->failed) return h->failed;
  B2F_ON_DEVICE(h->device);
  return pconv_enqueue(h, true, (float *)d_out, (const float *)d_in1, (const float *)d_in2, (cudaStream_t)stream);
}
// device buffers behind the host entry points, allocated by the first call that does not run on the pinned
// buffers directly
static int pconv_host_bufs(b2f_pconv *h, bool tv) {
  const size_t blk = (size_t)h->channels * h->pts * sizeof(float);
  if (!h->d_in1) CK(cudaMalloc((void **)&h->d_in1, blk));
  if (!h->d_out) CK(cudaMalloc((void **)&h->d_out, blk));
  if (tv && !h->d_in2) CK(cudaMalloc((void **)&h->d_in2, blk));
  return B2F_OK;
}
// General path (7-9 launches per block): the host call replays H2D + launches + D2H as ONE CUDA graph, captured
// from the very code the device entry points run (pconv_general_step) on the handle's fixed buffers. The ring
// positions are device-resident, so no node parameter changes between blocks. Blocks above the bounce limit (the
// throughput regime) and option graph = 0 take the plain stream path.
static const size_t kGraphZeroCopyMax = 128u << 10;
static int pconv_graph_call(b2f_pconv *h, bool tv, float *out, const float *in1, const float *in2) {
  const size_t blk = (size_t)h->channels * h->pts * sizeof(float);
  int rc;
  if ((rc = pconv_host_bufs(h, tv))) return rc;
  if ((rc = h->sg_in.ensure(blk)) || (rc = h->sg_out.ensure(blk)) || (tv && (rc = h->sg_in2.ensure(blk)))) return rc;
  cudaGraphExec_t &ge = h->graph[tv ? 1 : 0];
  if (!ge) {
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    cudaError_t e = cudaSuccess;
    if (blk <= kGraphZeroCopyMax) {
      // a few channels: the block's first kernel reads the pinned staging buffers and its last one writes the pinned
      // result directly (three copy nodes fewer: a mono 8192-sample block is 32 KB each way)
      rc = pconv_general_step(h, tv, (float *)h->sg_out.pin, (const float *)h->sg_in.pin, (const float *)h->sg_in2.pin, h->stream);
    } else {
      e = cudaMemcpyAsync(h->d_in1, h->sg_in.pin, blk, cudaMemcpyHostToDevice, h->stream);
      if (e == cudaSuccess && tv) e = cudaMemcpyAsync(h->d_in2, h->sg_in2.pin, blk, cudaMemcpyHostToDevice, h->stream);
      rc = e == cudaSuccess ? pconv_general_step(h, tv, h->d_out, h->d_in1, h->d_in2, h->stream) : cuda_fail(e, "capture H2D");
      if (!rc && (e = cudaMemcpyAsync(h->sg_out.pin, h->d_out, blk, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
        rc = cuda_fail(e, "capture D2H");
    }
    e = cudaStreamEndCapture(h->stream, &g);
    if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamEndCapture");
    if (!rc && (e = cudaGraphInstantiate(&ge, g, 0)) != cudaSuccess) rc = cuda_fail(e, "cudaGraphInstantiate");
    if (g) cudaGraphDestroy(g);
    if (rc) return rc;
  }
  memcpy(h->sg_in.pin, in1, blk);
  if (tv) memcpy(h->sg_in2.pin, in2, blk);
  CK(cudaGraphLaunch(ge, h->stream));
  h->wp = h->wp != h->nparts - 1 ? h->wp + 1 : 0;             // host mirror of d_state
  if (tv) h->wp2 = h->wp2 == 0 ? h->nparts - 1 : h->wp2 - 1;
  CK(cudaStreamSynchronize(h->stream));
  memcpy(out, h->sg_out.pin, blk);
  return B2F_OK;
}
// Device-side alias of a host buffer the CALLER has page-locked (cudaHostAlloc / cudaHostRegister: mapped into every
// device's address space under unified addressing), or nullptr for pageable memory and for pointers the kernels
// cannot take (16-byte alignment).
static void *pinned_alias(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer || !al16(a.devicePointer)) return nullptr;
  return a.devicePointer;
}
// Blocks above the bounce-buffer size whose buffers the caller has page-locked: ONE launch that reads the input blocks
// from and writes the output blocks to the caller's memory over PCIe (2 KB per channel each way, against 7.7 MB of HBM
// traffic per channel at configuration 5b), instead of upload -> kernel -> download. Returns -1 when not applicable.
static int pconv_pinned_direct(b2f_pconv *h, bool tv, float *out, const float *in1, const float *in2) {
  if (!h->opt.pinned_direct || h->general()) return -1;
  float *o = (float *)pinned_alias(out);
  const float *a = (const float *)pinned_alias(in1), *b = tv ? (const float *)pinned_alias(in2) : nullptr;
  if (!o || !a || (tv && !b)) return -1;
  int rc = pconv_enqueue(h, tv, o, a, b, h->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return B2F_OK;
}
extern "C" int b2f_pconv_process_host(b2f_pconv *h, float *out, const float *in) {
  if (!h || !out || !in) return B2F_ERR_INVALID_VALUE;
  if (h->failed) return h->failed;
  B2F_ON_DEVICE(h->device);
  int rc;
  const size_t blk = (size_t)h->channels * h->pts * sizeof(float);
  if (h->general() && h->opt.graph && blk <= kBounceMax) return pconv_graph_call(h, false, out, in, nullptr);
  if (blk <= (size_t)h->opt.zerocopy_max) {
    if ((rc = h->sg_in.ensure(blk)) || (rc = h->sg_out.ensure(blk))) return rc;
    memcpy(h->sg_in.pin, in, blk);
    if ((rc = pconv_enqueue(h, false, (float *)h->sg_out.pin, (const float *)h->sg_in.pin, nullptr, h->stream))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->sg_out.pin, blk);
    return B2F_OK;
  }
  if ((rc = pconv_pinned_direct(h, false, out, in, nullptr)) >= 0) return rc;
  if ((rc = pconv_host_bufs(h, false))) return rc;
  // Many channels, blocks too large for the bounce buffer (the copies go straight from / to the caller's memory,
  // asynchronously when it is pinned): the two halves of the channels run on two streams, so that the second
  // half's upload overlaps the first half's kernel and the first half's download the second half's kernel.
  if (!h->general() && h->channels >= 128 && blk > kBounceMax && h->opt.pconv_pipeline) {
    if (!h->stream2) CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    const int n0 = h->channels / 2, n1 = h->channels - n0;
    const size_t b0 = (size_t)n0 * h->pts, bytes0 = b0 * sizeof(float), bytes1 = (size_t)n1 * h->pts * sizeof(float);
    // Every channel's frame `wp` and overlap tail are rewritten by this block: if anything fails after the first
    // enqueue, both streams are drained (the caller's buffers must not be touched after we return) and the handle
    // is marked failed -- half its channels would be one block ahead of the other half -- until b2f_pconv_reset().
    // Pageable caller memory (or page-locked but unaligned): the halves go through the handle's own pinned buffers with
    // plain memcpy, and the kernels read / write those in place -- the second half's memcpy overlaps the first half's
    // kernel, the first half's copy-out the second half's kernel (1024 channels: 1.33 -> ~1.2 ms per block against
    // cudaMemcpyAsync from pageable memory, which the driver stages synchronously).
    auto run_staged = [&]() -> int {
      int r;
      if ((r = h->sg_in.ensure(blk)) || (r = h->sg_out.ensure(blk))) return r;
      float *pin = (float *)h->sg_in.pin, *pout = (float *)h->sg_out.pin;
      memcpy(pin, in, bytes0);
      if ((r = pconv_launch_range(h, pout, pin, 0, n0, h->stream))) return r;
      memcpy(pin + b0, in + b0, bytes1);
      if ((r = pconv_launch_range(h, pout, pin, n0, n1, h->stream2))) return r;
      CK(cudaStreamSynchronize(h->stream));
      memcpy(out, pout, bytes0);
      CK(cudaStreamSynchronize(h->stream2));
      memcpy(out + b0, pout + b0, bytes1);
      return B2F_OK;
    };
    auto run = [&]() -> int {
      int r;
      CK(cudaMemcpyAsync(h->d_in1, in, bytes0, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(h->d_in1 + b0, in + b0, bytes1, cudaMemcpyHostToDevice, h->stream2));
      if ((r = pconv_launch_range(h, h->d_out, h->d_in1, 0, n0, h->stream))) return r;
      if ((r = pconv_launch_range(h, h->d_out, h->d_in1, n0, n1, h->stream2))) return r;
      CK(cudaMemcpyAsync(out, h->d_out, bytes0, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(out + b0, h->d_out + b0, bytes1, cudaMemcpyDeviceToHost, h->stream2));
      return B2F_OK;
    };
    rc = h->opt.pinned_direct ? run_staged() : run();
    const cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->stream2);
    if (!rc && e1 != cudaSuccess) rc = cuda_fail(e1, "cudaStreamSynchronize");
    if (!rc && e2 != cudaSuccess) rc = cuda_fail(e2, "cudaStreamSynchronize");
    if (rc) {
      h->failed = rc;
      return rc;
    }
    h->wp = h->wp != h->nparts - 1 ? h->wp + 1 : 0;  // cl_conv.cpp:424
    return B2F_OK;
  }
  if ((rc = h2d(h->d_in1, in, blk, h->sg_in, h->stream))) return rc;
  if ((rc = pconv_enqueue(h, false, h->d_out, h->d_in1, nullptr, h->stream))) return rc;
  return d2h(out, h->d_out, blk, h->sg_out, h->stream);
}
extern "C" int b2f_pconv_process_tv_host(b2f_pconv *h, float *out, const float *in1, const float *in2) {
  if (!h || !out || !in1 || !in2) return B2F_ERR_INVALID_VALUE;
  if (h->failed) return h->failed;
  B2F_ON_DEVICE(h->device);
  int rc;
  const size_t blk = (size_t)h->channels * h->pts * sizeof(float);
  if (h->general() && h->opt.graph && blk <= kBounceMax) return pconv_graph_call(h, true, out, in1, in2);
  if (blk <= (size_t)h->opt.zerocopy_max) {
    if ((rc = h->sg_in.ensure(blk)) || (rc = h->sg_in2.ensure(blk)) || (rc = h->sg_out.ensure(blk))) return rc;
    memcpy(h->sg_in.pin, in1, blk);
    memcpy(h->sg_in2.pin, in2, blk);
    if ((rc = pconv_enqueue(h, true, (float *)h->sg_out.pin, (const float *)h->sg_in.pin, (const float *)h->sg_in2.pin,
                            h->stream)))
      return rc;
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->sg_out.pin, blk);
    return B2F_OK;
  }
  if ((rc = pconv_pinned_direct(h, true, out, in1, in2)) >= 0) return rc;
  if ((rc = pconv_host_bufs(h, true))) return rc;
  if ((rc = h2d(h->d_in1, in1, blk, h->sg_in, h->stream))) return rc;
  if ((rc = h2d(h->d_in2, in2, blk, h->sg_in2, h->stream))) return rc;
  if ((rc = pconv_enqueue(h, true, h->d_out, h->d_in1, h->d_in2, h->stream))) return rc;
  return d2h(out, h->d_out, blk, h->sg_out, h->stream);
}
extern "C" int b2f_pconv_read_spectra(b2f_pconv *h, int which, int channel, float *dst) {
  if (!h || !dst || channel < 0 || channel >= h->channels || (which != 1 && which != 2)) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  const size_t per = (size_t)h->nparts * h->pts;
  const float2 *src = (which == 1 ? h->d_fdl : h->d_irs) + (size_t)channel * per;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(dst, src, per * sizeof(float2), cudaMemcpyDeviceToHost));
  return B2F_OK;
}

// =====================================================================================================
// direct convolution
// =====================================================================================================
struct b2f_dconv {
  int device = 0, irsize = 0, vsize = 0, channels = 1, max_blocks = 1;
  int wp = 0;  // the reference's ring position (cl_dconv.cpp:124), kept for the coefficient ring
  int cur = 0; // which history buffer is current
  float *d_hist[2] = {nullptr, nullptr};
  float *d_coefs = nullptr, *d_grev = nullptr, *d_in1 = nullptr, *d_in2 = nullptr, *d_out = nullptr;
  cudaStream_t stream = nullptr;
  Staging sg_in, sg_in2, sg_out;
  Options opt;
  int L() const { return irsize + vsize; }
  void destroy() {
    DeviceGuard guard(device);  // a failed create may carry an invalid ordinal: the guard swallows that
    for (void *p : {(void *)d_hist[0], (void *)d_hist[1], (void *)d_coefs, (void *)d_grev, (void *)d_in1, (void *)d_in2,
                    (void *)d_out})
      if (p) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
    sg_in.release();
    sg_in2.release();
    sg_out.release();
  }
};

extern "C" int b2f_dconv_create(b2f_dconv **out, int device, int irsize, int vsize, int channels, int max_blocks) {
  if (!out) return B2F_ERR_INVALID_VALUE;
  *out = nullptr;
  if (irsize < 1 || vsize < 1 || channels < 1) return B2F_ERR_INVALID_VALUE;
  if (channels > 65535) return B2F_ERR_UNSUPPORTED;
  int rc = check_device(device);
  if (rc) return rc;
  B2F_ON_DEVICE(device);
  b2f_dconv *h = new (std::nothrow) b2f_dconv;
  if (!h) return B2F_ERR_ALLOC;
  h->device = device, h->irsize = irsize, h->vsize = vsize, h->channels = channels;
  h->max_blocks = max_blocks < 1 ? 1 : max_blocks;
  h->opt = current_options();
  auto fail = [&](int code) {
    h->destroy();
    delete h;
    return code;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cuda_fail(e, "stream"));
  const size_t hb = (size_t)channels * irsize * sizeof(float), cb = (size_t)channels * h->L() * sizeof(float);
  for (int i = 0; i < 2; i++) {
    if ((e = cudaMalloc((void **)&h->d_hist[i], hb)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc hist"));
    if ((e = cudaMemsetAsync(h->d_hist[i], 0, hb, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  }
  if ((e = cudaMalloc((void **)&h->d_coefs, cb)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc coefs"));
  if ((e = cudaMemsetAsync(h->d_coefs, 0, cb, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  if ((e = cudaMalloc((void **)&h->d_grev, hb)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc grev"));
  if ((e = cudaMemsetAsync(h->d_grev, 0, hb, h->stream)) != cudaSuccess) return fail(cuda_fail(e, "memset"));
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(cuda_fail(e, "sync"));
  *out = h;
  return B2F_OK;
}
extern "C" int b2f_dconv_destroy(b2f_dconv *h) {
  if (!h) return B2F_OK;
  h->destroy();
  delete h;
  return B2F_OK;
}
extern "C" int b2f_dconv_reset(b2f_dconv *h) {
  if (!h) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  for (int i = 0; i < 2; i++) CK(cudaMemsetAsync(h->d_hist[i], 0, (size_t)h->channels * h->irsize * sizeof(float), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->wp = 0;
  return B2F_OK;
}
// keep the reversed tap copy the FIR kernel streams in step with the coefficient ring
static int dconv_reverse(b2f_dconv *h, cudaStream_t st) {
  dim3 grid((h->irsize + 255) / 256, h->channels, 1);
  dconv_reverse_kernel<<<grid, 256, 0, st>>>(h->d_grev, h->d_coefs, h->irsize, h->L());
  CK(cudaGetLastError());
  return B2F_OK;
}
extern "C" int b2f_dconv_push_ir_dev(b2f_dconv *h, const void *d_ir, size_t ir_stride, void *stream) {
  if (!h || !d_ir || ir_stride < (size_t)h->irsize) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  CK(cudaMemcpy2DAsync(h->d_coefs, (size_t)h->L() * sizeof(float), d_ir, ir_stride * sizeof(float),
                       (size_t)h->irsize * sizeof(float), h->channels, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return dconv_reverse(h, (cudaStream_t)stream);
}
extern "C" int b2f_dconv_push_ir_host(b2f_dconv *h, const float *ir, size_t ir_stride) {
  if (!h || !ir || ir_stride < (size_t)h->irsize) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  CK(cudaMemcpy2DAsync(h->d_coefs, (size_t)h->L() * sizeof(float), ir, ir_stride * sizeof(float),
                       (size_t)h->irsize * sizeof(float), h->channels, cudaMemcpyHostToDevice, h->stream));
  int rc = dconv_reverse(h, h->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return B2F_OK;
}

static bool tiles_too_many(long long nout) { return (nout + 255) / 256 > 65535; }
template <int TN>
static int dconv_launch_t(const DconvArgs &a, int channels, cudaStream_t st) {
  using G = DconvGeom<TN>;
  const int tiles = (a.nout + G::TILE - 1) / G::TILE;
  // split the taps over a cluster when the grid would not cover the GPU
  int S = 1;
  while (S < 8 && (long long)tiles * channels * S < 296 && a.irsize / (S * 2) >= 64) S *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, tiles, channels);
  cfg.blockDim = dim3(kDcThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, dconv_fir_kernel<TN>, a));
  return B2F_OK;
}
static int dconv_enqueue(b2f_dconv *h, float *d_out, const float *d_in, int nblocks, cudaStream_t st) {
  DconvArgs a;
  a.hist_in = h->d_hist[h->cur];
  a.hist_out = h->d_hist[h->cur ^ 1];
  a.grev = h->d_grev;
  a.in = d_in, a.out = d_out;
  a.irsize = h->irsize, a.nout = nblocks * h->vsize;
  a.vec_ok = (h->irsize % 4 == 0) && (a.nout % 4 == 0) && ((uintptr_t)d_in % 16 == 0);
  // 16 outputs per thread when there is enough stream per launch to fill 512-output tiles, 8 otherwise
  int rc = a.nout >= 512 ? dconv_launch_t<16>(a, h->channels, st) : dconv_launch_t<8>(a, h->channels, st);
  if (rc) return rc;
  h->cur ^= 1;
  h->wp = (int)(((long long)h->wp + (long long)nblocks * h->vsize) % h->L());  // cl_dconv.cpp:124
  return B2F_OK;
}
extern "C" int b2f_dconv_process_dev(b2f_dconv *h, void *d_out, const void *d_in, int nblocks, void *stream) {
  if (!h || !d_out || !d_in || nblocks < 1) return B2F_ERR_INVALID_VALUE;
  if (tiles_too_many((long long)nblocks * h->vsize)) return B2F_ERR_UNSUPPORTED;
  B2F_ON_DEVICE(h->device);
  return dconv_enqueue(h, (float *)d_out, (const float *)d_in, nblocks, (cudaStream_t)stream);
}
static int dconv_coef_write(b2f_dconv *h, const float *d_in2, cudaStream_t st) {
  dim3 grid((h->vsize + 255) / 256, h->channels, 1);
  dconv_coef_write_kernel<<<grid, 256, 0, st>>>(h->d_coefs, h->d_grev, d_in2, h->vsize, h->irsize, h->L(), h->wp);
  CK(cudaGetLastError());
  return B2F_OK;
}
extern "C" int b2f_dconv_process_tv_dev(b2f_dconv *h, void *d_out, const void *d_in1, const void *d_in2, void *stream) {
  if (!h || !d_out || !d_in1 || !d_in2) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  int rc = dconv_coef_write(h, (const float *)d_in2, (cudaStream_t)stream);
  if (rc) return rc;
  return dconv_enqueue(h, (float *)d_out, (const float *)d_in1, 1, (cudaStream_t)stream);
}
static int dconv_host_bufs(b2f_dconv *h, bool tv) {
  const size_t blk = (size_t)h->channels * h->max_blocks * h->vsize * sizeof(float);
  if (!h->d_in1) CK(cudaMalloc((void **)&h->d_in1, blk));
  if (!h->d_out) CK(cudaMalloc((void **)&h->d_out, blk));
  if (tv && !h->d_in2) CK(cudaMalloc((void **)&h->d_in2, (size_t)h->channels * h->vsize * sizeof(float)));
  return B2F_OK;
}
extern "C" int b2f_dconv_process_host(b2f_dconv *h, float *out, const float *in, int nblocks) {
  if (!h || !out || !in || nblocks < 1) return B2F_ERR_INVALID_VALUE;
  if (nblocks > h->max_blocks) return B2F_ERR_BATCH;
  B2F_ON_DEVICE(h->device);
  int rc = dconv_host_bufs(h, false);
  if (rc) return rc;
  const size_t bytes = (size_t)h->channels * nblocks * h->vsize * sizeof(float);
  if (bytes <= (size_t)h->opt.zerocopy_max) {
    if ((rc = h->sg_in.ensure(bytes)) || (rc = h->sg_out.ensure(bytes))) return rc;
    memcpy(h->sg_in.pin, in, bytes);
    if ((rc = dconv_enqueue(h, (float *)h->sg_out.pin, (const float *)h->sg_in.pin, nblocks, h->stream))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->sg_out.pin, bytes);
    return B2F_OK;
  }
  if ((rc = h2d(h->d_in1, in, bytes, h->sg_in, h->stream))) return rc;
  if ((rc = dconv_enqueue(h, h->d_out, h->d_in1, nblocks, h->stream))) return rc;
  return d2h(out, h->d_out, bytes, h->sg_out, h->stream);
}
extern "C" int b2f_dconv_process_tv_host(b2f_dconv *h, float *out, const float *in1, const float *in2) {
  if (!h || !out || !in1 || !in2) return B2F_ERR_INVALID_VALUE;
  B2F_ON_DEVICE(h->device);
  int rc = dconv_host_bufs(h, true);
  if (rc) return rc;
  const size_t bytes = (size_t)h->channels * h->vsize * sizeof(float);
  if (bytes <= (size_t)h->opt.zerocopy_max) {
    if ((rc = h->sg_in.ensure(bytes)) || (rc = h->sg_in2.ensure(bytes)) || (rc = h->sg_out.ensure(bytes))) return rc;
    memcpy(h->sg_in.pin, in1, bytes);
    memcpy(h->sg_in2.pin, in2, bytes);
    if ((rc = dconv_coef_write(h, (const float *)h->sg_in2.pin, h->stream))) return rc;
    if ((rc = dconv_enqueue(h, (float *)h->sg_out.pin, (const float *)h->sg_in.pin, 1, h->stream))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->sg_out.pin, bytes);
    return B2F_OK;
  }
  if ((rc = h2d(h->d_in1, in1, bytes, h->sg_in, h->stream))) return rc;
  if ((rc = h2d(h->d_in2, in2, bytes, h->sg_in2, h->stream))) return rc;
  if ((rc = dconv_coef_write(h, h->d_in2, h->stream))) return rc;
  if ((rc = dconv_enqueue(h, h->d_out, h->d_in1, 1, h->stream))) return rc;
  return d2h(out, h->d_out, bytes, h->sg_out, h->stream);
}

#include "multi_gpu.inl"
