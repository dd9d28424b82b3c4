/*
 * oracle/ref_cpu.h -- CPU restatement of the opencl_fft hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain-C, single-threaded, float32, work-item-by-work-item
 * restatement of what the reference's OpenCL kernels and host dispatch code compute.
 * Nothing under opencl_fft_b200/ (the product) may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle.py) against
 *   (a) the two known answers implied by the reference's own test programs
 *       (test_cfft.cpp:54-56, test_rfft.cpp:54-57), and
 *   (b) the UNMODIFIED reference sources executed on the CPU through oracle/minicl
 *       (oracle/_ref/libclfft_ref.so, built by oracle/Makefile from /root/reference) --
 *       bit-exact, since both run work-items in ascending global-id order with
 *       -ffp-contract=off, and
 *   (c) the fixtures under tests/golden/ that (b) generated (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines (relative to /root/reference) it follows.
 */
#ifndef ORACLE_REF_CPU_H
#define ORACLE_REF_CPU_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Clcfft (cl_fft.cpp:44-161) ------------------------------------------------------ */
/* In-place N-point complex transform on interleaved float pairs. fwd!=0: scaled 1/N in the
 * last stage (cl_fft.cpp:39-40); fwd==0: unscaled inverse. Returns 0. */
int orc_cfft(float *c, int N, int fwd);

/* ---- Clrfft (cl_fft.cpp:208-296) ----------------------------------------------------- */
/* size real points <-> size/2 packed complex, in place on the same storage.
 * fwd: pack, scaled C2C, split for i in [0, size/4) (cl_fft.cpp:272-282).
 * inv: unsplit for i in [0, size/4), unscaled C2C (cl_fft.cpp:283-294). */
int orc_rfft(float *c, int size, int fwd);

/* ---- Clpconv (cl_conv.cpp:140-548, cl_conv_kernels.h:46-124) ------------------------- */
typedef struct orc_pconv orc_pconv;
orc_pconv *orc_pconv_create(int cvs, int pts);
void orc_pconv_destroy(orc_pconv *p);
int orc_pconv_nparts(const orc_pconv *p);
/* reads nparts*pts floats (cl_conv.cpp:353-388) */
int orc_pconv_push_ir(orc_pconv *p, const float *ir);
/* one block: pts floats in -> pts floats out (cl_conv.cpp:393-458) */
int orc_pconv_convolution(orc_pconv *p, float *out, const float *in);
/* time-varying block (cl_conv.cpp:460-548) */
int orc_pconv_convolution_tv(orc_pconv *p, float *out, const float *in1, const float *in2);
/* raw state access for white-box parity tests: FDL ring (spec1), IR ring (spec2),
 * overlap buffer (olap, 2*pts floats: [0,pts) last output, [pts,2pts) saved tail) */
const float *orc_pconv_spec1(const orc_pconv *p);
const float *orc_pconv_spec2(const orc_pconv *p);
const float *orc_pconv_olap(const orc_pconv *p);

/* ---- Cldconv (cl_dconv.cpp:46-153) --------------------------------------------------- */
typedef struct orc_dconv orc_dconv;
orc_dconv *orc_dconv_create(int irsize, int vsize);
void orc_dconv_destroy(orc_dconv *d);
int orc_dconv_push_ir(orc_dconv *d, const float *ir);
int orc_dconv_convolution(orc_dconv *d, float *out, const float *in);
int orc_dconv_convolution_tv(orc_dconv *d, float *out, const float *in1, const float *in2);

/* ---- float64 ground truth (not a restatement; used to report absolute accuracy) ------- */
/* naive O(N^2) DFT in double, sign=-1 forward / +1 inverse, no scaling */
void orc_dft64(const double *in_ri, double *out_ri, int N, int sign);

#ifdef __cplusplus
}
#endif
#endif
