/*
 * b200fft.h -- C ABI of the B200-native FFT / convolution engine (libb200fft.so).
 *
 * This is the drop-in boundary for the hot path of vlazzarini/opencl_fft. Everything is
 * extern "C": opaque handles, plain pointers, ints. The reference's own C++ classes
 * (include/cl_fft.h, include/cl_conv.h, include/cl_dconv.h -- same public interface as the
 * reference's headers of the same name) are thin inline wrappers over these entry points, so
 * test_cfft.cpp, test_rfft.cpp and csound/opcode.cpp compile unchanged; any other host language
 * binds the same symbols (INTEGRATION.md shows the ctypes binding shipped in opencl_fft_b200/).
 *
 * Each entry point names the reference interface it replaces (file:line under the reference
 * repository). Semantics -- scaling, packing, ring order, quirks Q1..Q6, Q9, Q12 of SURVEY.md --
 * are the reference's; where the reference is undefined or broken (Q10, Q11, Q13, Q15) the
 * behaviour is defined here and documented at the function.
 *
 * Conventions
 *   - return value: 0 = success (== CL_SUCCESS), > 0 = B2F_ERR_* (the reference's callers test
 *     `err > 0`, test_cfft.cpp:41, so positive codes make those checks live). b2f_error_string()
 *     maps a code to text. There is NO CPU fallback: without a CUDA device every create fails.
 *   - "host" entry points take host pointers and are synchronous (they return after the result
 *     is in the caller's buffer), like the reference's blocking reads (cl_fft.cpp:158).
 *   - "dev" entry points take device pointers and a cudaStream_t (passed as void*) and only
 *     enqueue work; this is the batched, HBM-resident path the roofline numbers are measured on.
 *     FFT and partitioned-convolution block pointers must be 16-byte aligned (B2F_ERR_INVALID_VALUE
 *     otherwise); impulse responses and direct-convolution blocks need only float alignment, and
 *     `ir_stride` may be any value >= the IR length.
 *   - complex data is interleaved float32 (re, im) == std::complex<float> == OpenCL float2.
 *   - batch / channels: the reference has one object per transform / channel; a handle here
 *     carries `max_batch` transforms or `channels` independent convolver states and runs them in
 *     one launch. batch = channels = 1 is exactly the reference object.
 */
#ifndef B200FFT_H
#define B200FFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define B2F_OK 0
#define B2F_ERR_NO_DEVICE 1      /* no CUDA device / bad ordinal (cf. CL_DEVICE_NOT_FOUND) */
#define B2F_ERR_INVALID_VALUE 2  /* bad size / NULL pointer (cf. CL_INVALID_VALUE) */
#define B2F_ERR_UNSUPPORTED 3    /* size outside what this build implements */
#define B2F_ERR_ALLOC 4          /* device or pinned-host allocation failed */
#define B2F_ERR_CUDA 5           /* a CUDA call / launch failed; see b2f_last_cuda_error() */
#define B2F_ERR_BATCH 6          /* batch larger than the handle's max_batch */

const char *b2f_error_string(int code);
/* text of the last CUDA runtime error seen by the calling thread's most recent failing call */
const char *b2f_last_cuda_error(void);
/* library version string, e.g. "b200fft 0.1 (sm_100a)" */
const char *b2f_version(void);

/* ---- options ---------------------------------------------------------------------------------
 * Re-measurement knobs: process-wide defaults that a handle COPIES when it is created; no entry
 * point consults them (or the environment) afterwards. Initial values come from the environment
 * variables B2F_<NAME IN CAPITALS>, read once before the first create. The reference has no
 * runtime configuration: the constructor arguments stay the only per-object input and
 * every default below is the measured winner.
 *   fft_sm_min_batch  96     N = 32768 complex / 65536 real: the one-SM kernel from this batch up
 *                            (0 never, 1 always); below it the four-step launch pair. N = 16384 / 32768 real:
 *                            from twice this batch up (two transforms per CTA iteration); below, one CTA each
 *   fft_sm_8192       1      complex N = 8192 and the inverse 16384-point real transform on the one-SM kernel as well
 *                            (four transforms per CTA iteration), from four times fft_sm_min_batch; 0: one CTA each
 *   large_chunk_mb    256    scratch chunk of the four-step launch pair
 *   rows_rb16         0      16-row CTAs in the four-step real rows kernel
 *   separate_split    0      unfused real split / unsplit pass on the four-step path
 *   pconv_tma         -1     partitioned-convolution MAC feed: -1 measured choice, 0 registers, 1 TMA
 *   pconv_cluster     0      cluster split of the partitions: 0 measured choice, else 1 | 2 | 4 | 8 | 16
 *                            (anything else, or more than nparts: create fails with INVALID_VALUE)
 *   pconv_pipeline    1      two-stream host calls: partitioned convolution with >= 128 channels (halves of the
 *                            channels), FFT batches above 1 MB (eight chunks of the batch)
 *   zerocopy_max      65536  host calls moving at most this many bytes run on pinned buffers directly
 *   graph             1      CUDA graph replay for the multi-launch host paths (pts >= 8192)
 *   pinned_direct     1      partitioned-convolution host calls above zerocopy_max whose buffers the caller has
 *                            page-locked: one launch reads / writes them in place over PCIe (no staging copies)
 *   fft_prefetch      -1     real transforms of 8192 / 16384 complex points, one CTA each: L2 prefetch of the
 *                            transform this many CTAs ahead (-1: the co-resident CTAs, 0: off)
 *   pconv_cluster16_max_channels 4  clusters of 16 CTAs (non-portable size) for handles of up to this many channels
 *                            (four times as many at pts <= 1024)
 *                            and at least 8 MB of rings per channel; 0 never
 *   pconv_deep_ring   1      handles whose launches put at most one CTA on an SM (mono and few-channel streams of 2048 /
 *                            4096-sample partitions): TMA stages of 32 KB; 0: the usual 4 KB slices
 *   pconv_deep_min_parts 16  ... when a CTA streams at least this many partitions
 *   pconv_ksplit      0      pts >= 8192 with few channels: the partitions of the spectral multiply-accumulate are split
 *                            over this many CTAs per 512-bin tile and the partial sums added by a second launch
 *                            (0: enough for two CTAs per SM, -1: never)
 *   pconv_general_fused 1    pts 8192 / 16384 below the one-SM FFT's batch: a block is new frames -> multiply-accumulate
 *                            (-> partial sums) -> inverse transform + overlap-add, 3-4 launches; 0: the 11-12 separate ones
 *   pconv_push_reg    1      push_ir (64 <= pts <= 4096) on the register-level real transform of the batched FFT;
 *                            0: the step kernel's frame routine (shared-memory split)
 *   verbose           0
 * Unknown names return B2F_ERR_INVALID_VALUE. */
int b2f_set_option(const char *name, long long value);
int b2f_get_option(const char *name, long long *value);

/* ---- devices (replaces clGetDeviceIDs / clGetDeviceInfo as used at test_cfft.cpp:31-38,
 *      csound/opcode.cpp:57-61) ------------------------------------------------------------- */
int b2f_device_count(int *count);
int b2f_device_name(int device, char *buf, size_t buflen);

/* ---- complex FFT: cl_fft::Clcfft (cl_fft.h:29-70, cl_fft.cpp:44-161) ----------------------
 * N: power of two, 2 <= N <= 65536 (the reference's int32 index math overflows above, Q13).
 * fwd != 0: X[k] = (1/N) sum x[n] exp(-2 pi i k n / N)   (scaled, cl_fft.cpp:39-40)
 * fwd == 0: x[n] = sum X[k] exp(+2 pi i k n / N)         (unscaled)                            */
typedef struct b2f_cfft b2f_cfft;
int b2f_cfft_create(b2f_cfft **plan, int device, int N, int fwd, int max_batch);
int b2f_cfft_destroy(b2f_cfft *plan);
/* Clcfft::transform (cl_fft.cpp:153-161): in place on `batch` consecutive N-point host arrays */
int b2f_cfft_exec_host(b2f_cfft *plan, float *c, int batch);
/* device-resident: in/out [batch][N] complex, may alias; asynchronous on `stream`. Stream order is kept as for any
 * kernel launch. (The 16384 / 32768-point kernels are launched with programmatic stream serialisation: a following
 * launch of theirs may be SCHEDULED while the previous one drains, but reads and writes nothing before the previous
 * kernel has completed and flushed -- griddepcontrol.wait -- so dependent calls on one stream need no extra care.) */
int b2f_cfft_exec_dev(b2f_cfft *plan, const void *d_in, void *d_out, int batch, void *stream);

/* ---- real FFT: cl_fft::Clrfft (cl_fft.h:74-111, cl_fft.cpp:208-296) -----------------------
 * size: number of real points, power of two, 4 <= size <= 131072; N = size/2 complex bins.
 * forward output: c[0] = (X[0], X[size/2]) / size packed; c[k] = 2 X[k] / size;
 * c[size/4] is the conjugate of that (the reference's split skips it, Q3). Inverse undoes it. */
typedef struct b2f_rfft b2f_rfft;
int b2f_rfft_create(b2f_rfft **plan, int device, int size, int fwd, int max_batch);
int b2f_rfft_destroy(b2f_rfft *plan);
/* Clrfft::transform(c, r) (cl_fft.cpp:267-296). c: size/2 complex, r: size reals, per transform,
 * `batch` of each back to back. r == (float*)c is the in-place form (cl_fft.h:105-110).
 * forward: reads r, writes c. inverse: reads c, writes r AND overwrites c with the same reals
 * (the reference's c is its transfer buffer, cl_fft.cpp:290-293). */
int b2f_rfft_exec_host(b2f_rfft *plan, float *c, float *r, int batch);
/* device-resident: forward d_in = [batch][size] float, d_out = [batch][size/2] complex;
 * inverse the other way round; may alias; asynchronous on `stream` */
int b2f_rfft_exec_dev(b2f_rfft *plan, const void *d_in, void *d_out, int batch, void *stream);

/* ---- partitioned convolution: cl_conv::Clpconv (cl_conv.h:124-188, cl_conv.cpp:140-548) ----
 * cvs: impulse-response length, pts: partition size (power of two). nparts = cvs / pts
 * TRUNCATING (cl_conv.cpp:143, Q4). `channels` independent convolvers (own IR, own delay line,
 * own overlap tail), laid out channel-major in every buffer.
 * The reference's optional host-memory constructor arguments (cl_conv.h:156-158) are dead code
 * there (Q15) and have no counterpart. */
typedef struct b2f_pconv b2f_pconv;
int b2f_pconv_create(b2f_pconv **h, int device, int cvs, int pts, int channels);
int b2f_pconv_destroy(b2f_pconv *h);
int b2f_pconv_nparts(const b2f_pconv *h);
/* Zero the delay line and the overlap tail and put both ring positions back to their start
 * (wp = 0, wp2 = nparts - 1, cl_conv.cpp:144). The IR spectra are KEPT, frame for frame: after
 * static push_ir that is the pushed IR; after time-varying use it is whatever the ring held, which
 * later time-varying blocks re-record from frame nparts - 1 downwards exactly as a fresh object
 * would. Waits for the handle's own streams; work queued through *_dev entry points on a
 * caller's stream must have been synchronised by the caller. Clears a sticky failure (below). */
int b2f_pconv_reset(b2f_pconv *h);
/* Clpconv::push_ir (cl_conv.cpp:353-388). ir: per channel nparts*pts floats, channel c starting
 * at ir + c*ir_stride (ir_stride in floats; pass cvs for back-to-back IRs). */
int b2f_pconv_push_ir_host(b2f_pconv *h, const float *ir, size_t ir_stride);
int b2f_pconv_push_ir_dev(b2f_pconv *h, const void *d_ir, size_t ir_stride, void *stream);
/* Clpconv::convolution(out, in) (cl_conv.cpp:393-458): one block of pts samples per channel.
 * in/out: [channels][pts] floats.
 * Stream ordering: a handle's state is advanced by every call, *_host calls on the handle's own
 * stream(s), *_dev calls on the caller's; the caller orders the two kinds (synchronise the stream
 * before switching). If a multi-stream host call fails half way the handle turns FAILED -- every
 * later process call returns the same code -- until b2f_pconv_reset(). */
int b2f_pconv_process_host(b2f_pconv *h, float *out, const float *in);
int b2f_pconv_process_dev(b2f_pconv *h, void *d_out, const void *d_in, void *stream);
/* Clpconv::convolution(out, in1, in2) (cl_conv.cpp:460-548): time-varying; in2's block is
 * transformed into the IR ring at the descending write position (Q12). */
int b2f_pconv_process_tv_host(b2f_pconv *h, float *out, const float *in1, const float *in2);
int b2f_pconv_process_tv_dev(b2f_pconv *h, void *d_out, const void *d_in1, const void *d_in2, void *stream);
/* white-box access for parity tests: copy the frequency-domain delay line (`which` = 1, the
 * reference's spec1) or the IR spectra (`which` = 2, spec2) of one channel to the host,
 * nparts*pts complex values in the reference's frame order. */
int b2f_pconv_read_spectra(b2f_pconv *h, int which, int channel, float *dst);

/* ---- direct convolution: cl_conv::Cldconv (cl_dconv.h:17-66, cl_dconv.cpp:46-153) ----------
 * y[t] = sum_{c < irsize} ir[c] x[t-1-c]: the exact linear convolution delayed by ONE sample,
 * as the reference computes it (Q9). State starts at zero (the reference leaves it
 * uninitialised, Q11). When irsize % vsize != 0 the reference's ring write is broken (Q10);
 * here the stream semantics above simply continue to hold. */
typedef struct b2f_dconv b2f_dconv;
int b2f_dconv_create(b2f_dconv **h, int device, int irsize, int vsize, int channels, int max_blocks);
int b2f_dconv_destroy(b2f_dconv *h);
int b2f_dconv_reset(b2f_dconv *h);
/* Cldconv::push_ir (cl_dconv.cpp:150-153): irsize floats per channel */
int b2f_dconv_push_ir_host(b2f_dconv *h, const float *ir, size_t ir_stride);
int b2f_dconv_push_ir_dev(b2f_dconv *h, const void *d_ir, size_t ir_stride, void *stream);
/* Cldconv::convolution(out, in) (cl_dconv.cpp:109-132), `nblocks` consecutive blocks of vsize
 * samples per channel in one call (nblocks = 1 is the reference call).
 * in/out: [channels][nblocks*vsize] floats. */
int b2f_dconv_process_host(b2f_dconv *h, float *out, const float *in, int nblocks);
int b2f_dconv_process_dev(b2f_dconv *h, void *d_out, const void *d_in, int nblocks, void *stream);
/* Cldconv::convolution(out, in1, in2) (cl_dconv.cpp:134-148): in2's block is written into the
 * coefficient ring at the delay line's write position before the block is computed (Q12).
 * One block per call. in1/in2/out: [channels][vsize]. */
int b2f_dconv_process_tv_host(b2f_dconv *h, float *out, const float *in1, const float *in2);
int b2f_dconv_process_tv_dev(b2f_dconv *h, void *d_out, const void *d_in1, const void *d_in2, void *stream);

/* ---- several GPUs behind one handle ------------------------------------------------
 * The reference is one object = one channel = one device (cl_conv.cpp:154, cl_fft.cpp:49) and its
 * callers are single processes (csound/opcode.cpp:172-203). A *_multi handle shards `channels`
 * convolvers (or the transforms of a batch) over `ndev` CUDA devices in contiguous ranges --
 * channels [g*channels/ndev, (g+1)*channels/ndev) live on devices[g] for their whole life -- with
 * one stream and one host worker thread per device and NO inter-device communication. The host
 * entry points have the single-device signatures and semantics; one call fans out
 * H2D -> kernels -> D2H on all devices at once and returns when every device has finished.
 * devices[] must be distinct ordinals; channels (max_batch) >= ndev. */
typedef struct b2f_pconv_multi b2f_pconv_multi;
int b2f_pconv_multi_create(b2f_pconv_multi **h, const int *devices, int ndev, int cvs, int pts, int channels);
int b2f_pconv_multi_destroy(b2f_pconv_multi *h);
int b2f_pconv_multi_nparts(const b2f_pconv_multi *h);
int b2f_pconv_multi_reset(b2f_pconv_multi *h);
int b2f_pconv_multi_push_ir_host(b2f_pconv_multi *h, const float *ir, size_t ir_stride);
/* the same for the channels of devices[g] only (ir: that shard's IRs), e.g. to upload a large set shard by shard */
int b2f_pconv_multi_push_ir_shard_host(b2f_pconv_multi *h, int g, const float *ir, size_t ir_stride);
int b2f_pconv_multi_process_host(b2f_pconv_multi *h, float *out, const float *in);
int b2f_pconv_multi_process_tv_host(b2f_pconv_multi *h, float *out, const float *in1, const float *in2);
typedef struct b2f_dconv_multi b2f_dconv_multi;
int b2f_dconv_multi_create(b2f_dconv_multi **h, const int *devices, int ndev, int irsize, int vsize, int channels,
                           int max_blocks);
int b2f_dconv_multi_destroy(b2f_dconv_multi *h);
int b2f_dconv_multi_reset(b2f_dconv_multi *h);
int b2f_dconv_multi_push_ir_host(b2f_dconv_multi *h, const float *ir, size_t ir_stride);
int b2f_dconv_multi_process_host(b2f_dconv_multi *h, float *out, const float *in, int nblocks);
int b2f_dconv_multi_process_tv_host(b2f_dconv_multi *h, float *out, const float *in1, const float *in2);
/* transform b of a call's batch runs on devices[g], g*batch/ndev <= b < (g+1)*batch/ndev */
typedef struct b2f_cfft_multi b2f_cfft_multi;
int b2f_cfft_multi_create(b2f_cfft_multi **h, const int *devices, int ndev, int N, int fwd, int max_batch);
int b2f_cfft_multi_destroy(b2f_cfft_multi *h);
int b2f_cfft_multi_exec_host(b2f_cfft_multi *h, float *c, int batch);
typedef struct b2f_rfft_multi b2f_rfft_multi;
int b2f_rfft_multi_create(b2f_rfft_multi **h, const int *devices, int ndev, int size, int fwd, int max_batch);
int b2f_rfft_multi_destroy(b2f_rfft_multi *h);
int b2f_rfft_multi_exec_host(b2f_rfft_multi *h, float *c, float *r, int batch);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_H */
