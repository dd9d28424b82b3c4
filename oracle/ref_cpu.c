/*
 * oracle/ref_cpu.c -- CPU restatement of the opencl_fft hot path. TEST INFRASTRUCTURE ONLY:
 * see ref_cpu.h for the rules (never linked into the product) and the parity-pinning status.
 *
 * Style: every OpenCL kernel of the reference becomes a "work-item" function taking the
 * global id, and every clEnqueueNDRangeKernel becomes a loop over ascending global ids.
 * That is one legal schedule of the reference (its kernels either touch disjoint elements
 * per work-item or combine through float atomics, whose order is unspecified), and it is
 * the schedule oracle/minicl uses when it runs the real reference sources, so the two agree
 * bit for bit. Arithmetic is IEEE float32 with no contraction (build with -ffp-contract=off).
 *
 * Citations are file:line relative to /root/reference.
 */
#include "ref_cpu.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float x, y;
} cmplx; /* OpenCL float2 (cl_fft.cpp:18, cl_conv_kernels.h:15) */

static const double PI = 3.141592653589793; /* cl_fft.h:24; cl_conv.cpp:21 uses M_PI (same double) */

/* cl_fft.cpp:20-22 / cl_conv_kernels.h:17-19 */
static inline cmplx prod(cmplx a, cmplx b) {
  cmplx r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
/* cl_fft.cpp:170-172 */
static inline cmplx conjg(cmplx a) {
  cmplx r = {a.x, -a.y};
  return r;
}
/* cl_fft.cpp:174-176 ("rotation by pi" in the source; it is a multiplication by i) */
static inline cmplx rot(cmplx a) {
  cmplx r = {-a.y, a.x};
  return r;
}
static inline cmplx cadd(cmplx a, cmplx b) {
  cmplx r = {a.x + b.x, a.y + b.y};
  return r;
}
static inline cmplx csub(cmplx a, cmplx b) {
  cmplx r = {a.x - b.x, a.y - b.y};
  return r;
}
static inline cmplx cscale(float s, cmplx a) {
  cmplx r = {s * a.x, s * a.y};
  return r;
}

/* ---- tables ---------------------------------------------------------------------------- */

/* cl_fft.cpp:86-91 (sign=-1 fwd / +1 inv), cl_conv.cpp:263-276: w[i] = cos(i*2*PI/N) + sign*sin(..) i,
 * evaluated in double and rounded to float, N entries. */
static void make_twiddle(cmplx *w, int N, float sign) {
  for (int i = 0; i < N; i++) {
    w[i].x = (float)cos(i * 2 * PI / N);
    w[i].y = (float)(sign * sin(i * 2 * PI / N));
  }
}
/* cl_fft.cpp:233-238, cl_conv.cpp:277-287: split twiddle w2[i] = cos(i*PI/N) + sign*sin(i*PI/N) i */
static void make_twiddle2(cmplx *w, int N, float sign) {
  for (int i = 0; i < N; i++) {
    w[i].x = (float)cos(i * PI / N);
    w[i].y = (float)(sign * sin(i * PI / N));
  }
}
/* cl_fft.cpp:96-101, cl_conv.cpp:290-295: bit-reversal by the doubling recurrence */
static void make_bitrev(int *b, int N) {
  for (int i = 0; i < N; i++) b[i] = i;
  for (int i = 1, n = N / 2; i < N; i = i << 1, n = n >> 1)
    for (int j = 0; j < i; j++) b[i + j] = b[j] + n;
}

/* ---- fft_code kernels (cl_fft.cpp:24-41) ---------------------------------------------- */

/* cl_fft.cpp:24-27 */
static void k_reorder(cmplx *out, const cmplx *in, const int *b, int gid) { out[gid] = in[b[gid]]; }

/* cl_fft.cpp:29-41. scale: forward && last stage divides by N (float2 / int -> float division) */
static void k_fft(cmplx *s, const cmplx *w, int N, int n2, int fwd, int gid) {
  int k, i, m, n;
  cmplx e, o;
  k = gid * n2;
  m = k / N;
  n = n2 >> 1;
  k = k % N + m;
  i = k + n;
  e = s[k];
  o = prod(s[i], w[m * N / n2]);
  if (n2 == N && fwd) {
    float fN = (float)N;
    s[k].x = (e.x + o.x) / fN;
    s[k].y = (e.y + o.y) / fN;
    s[i].x = (e.x - o.x) / fN;
    s[i].y = (e.y - o.y) / fN;
  } else {
    s[k] = cadd(e, o);
    s[i] = csub(e, o);
  }
}

/* Clcfft::fft(), cl_fft.cpp:138-151: reorder data1 -> data2, then log2N stage launches on data2 */
static void cfft_core(cmplx *data2, const cmplx *data1, const cmplx *w, const int *b, int N, int fwd) {
  for (int g = 0; g < N; g++) k_reorder(data2, data1, b, g);
  for (int n = 1; n < N; n *= 2) {
    int n2 = n << 1;
    for (int g = 0; g < (N >> 1); g++) k_fft(data2, w, N, n2, fwd, g);
  }
}

int orc_cfft(float *c, int N, int fwd) {
  cmplx *w = (cmplx *)malloc(sizeof(cmplx) * N);
  int *b = (int *)malloc(sizeof(int) * N);
  cmplx *d1 = (cmplx *)malloc(sizeof(cmplx) * N);
  cmplx *d2 = (cmplx *)malloc(sizeof(cmplx) * N);
  make_twiddle(w, N, fwd ? -1.f : 1.f);
  make_bitrev(b, N);
  memcpy(d1, c, sizeof(cmplx) * N); /* cl_fft.cpp:155 */
  cfft_core(d2, d1, w, b, N, fwd);  /* cl_fft.cpp:157 */
  memcpy(c, d2, sizeof(cmplx) * N); /* cl_fft.cpp:158 */
  free(w);
  free(b);
  free(d1);
  free(d2);
  return 0;
}

/* ---- r2c_code kernels (cl_fft.cpp:178-205) -------------------------------------------- */

/* cl_fft.cpp:178-191 */
static void k_conv(cmplx *c, const cmplx *w, int N, int i) {
  if (!i) {
    cmplx z = c[0];
    c[0].x = (z.x + z.y) * .5f;
    c[0].y = (z.x - z.y) * .5f;
    return;
  }
  int j = N - i;
  cmplx e, o, cj = conjg(c[j]), p;
  e = cscale(.5f, cadd(c[i], cj));
  o = cscale(.5f, rot(csub(cj, c[i])));
  p = prod(w[i], o);
  c[i] = cadd(e, p);
  c[j] = conjg(csub(e, p));
}
/* cl_fft.cpp:192-205 */
static void k_iconv(cmplx *c, const cmplx *w, int N, int i) {
  if (!i) {
    cmplx z = c[0];
    c[0].x = (z.x + z.y);
    c[0].y = (z.x - z.y);
    return;
  }
  int j = N - i;
  cmplx e, o, cj = conjg(c[j]), p;
  e = cscale(.5f, cadd(c[i], cj));
  o = cscale(.5f, rot(csub(c[i], cj)));
  p = prod(w[i], o);
  c[i] = cadd(e, p);
  c[j] = conjg(csub(e, p));
}

/* Clrfft::transform, cl_fft.cpp:267-296 (N = size/2, cl_fft.cpp:208-210). The split kernels
 * run with threads = N>>1 (lines 278, 286), so element N/2 is never visited. */
int orc_rfft(float *c, int size, int fwd) {
  int N = size / 2;
  cmplx *w = (cmplx *)malloc(sizeof(cmplx) * N);
  cmplx *w2 = (cmplx *)malloc(sizeof(cmplx) * N);
  int *b = (int *)malloc(sizeof(int) * N);
  cmplx *d1 = (cmplx *)malloc(sizeof(cmplx) * N);
  cmplx *d2 = (cmplx *)malloc(sizeof(cmplx) * N);
  make_twiddle(w, N, fwd ? -1.f : 1.f);
  make_twiddle2(w2, N, fwd ? -1.f : 1.f);
  make_bitrev(b, N);
  memcpy(d1, c, sizeof(cmplx) * N);
  if (fwd) {
    cfft_core(d2, d1, w, b, N, 1);                          /* 277 */
    for (int g = 0; g < (N >> 1); g++) k_conv(d2, w2, N, g); /* 278-280 on data2 */
  } else {
    for (int g = 0; g < (N >> 1); g++) k_iconv(d1, w2, N, g); /* 286-288 on data1 */
    cfft_core(d2, d1, w, b, N, 0);                            /* 289 */
  }
  memcpy(c, d2, sizeof(cmplx) * N);
  free(w);
  free(w2);
  free(b);
  free(d1);
  free(d2);
  return 0;
}

/* ---- pconvcode kernels (cl_conv_kernels.h:46-124) ------------------------------------- */

/* cl_conv_kernels.h:46-52: gather into frame `offs` and zero the source */
static void pk_reorder(cmplx *out, cmplx *in, const int *b, int offs, int k) {
  out += offs;
  out[k] = in[b[k]];
  in[b[k]].x = 0.f;
  in[b[k]].y = 0.f;
}
/* cl_conv_kernels.h:54-68: unscaled stage on frame `offs` */
static void pk_fft(cmplx *s, const cmplx *w, int N, int n2, int offs, int gid) {
  int k, i, m, n;
  cmplx e, o;
  s += offs;
  k = gid * n2;
  m = k / N;
  n = n2 >> 1;
  k = k % N + m;
  i = k + n;
  e = s[k];
  o = prod(s[i], w[m * N / n2]);
  s[k] = cadd(e, o);
  s[i] = csub(e, o);
}
/* cl_conv_kernels.h:70-85. `if(!i%N)` is (!i)%N, i.e. i==0 */
static void pk_r2c(cmplx *c, const cmplx *w, int N, int offs, int i) {
  int j = N - i;
  c += offs;
  if ((!i) % N) {
    cmplx z = c[0];
    c[0].x = (z.x + z.y) * .5f;
    c[0].y = (z.x - z.y) * .5f;
    return;
  }
  cmplx e, o, cj = conjg(c[j]), p;
  e = cscale(.5f, cadd(c[i], cj));
  o = cscale(.5f, rot(csub(cj, c[i])));
  p = prod(w[i], o);
  c[i] = cadd(e, p);
  c[j] = conjg(csub(e, p));
}
/* cl_conv_kernels.h:87-100 */
static void pk_c2r(cmplx *c, const cmplx *w, int N, int i) {
  if (!i) {
    cmplx z = c[0];
    c[0].x = (z.x + z.y);
    c[0].y = (z.x - z.y);
    return;
  }
  int j = N - i;
  cmplx e, o, cj = conjg(c[j]), p;
  e = cscale(.5f, cadd(c[i], cj));
  o = cscale(.5f, rot(csub(c[i], cj)));
  p = prod(w[i], o);
  c[i] = cadd(e, p);
  c[j] = conjg(csub(e, p));
}
/* cl_conv_kernels.h:102-118; AtomicAdd (29-44) is a plain += under the sequential schedule */
static void pk_convol(float *out, const cmplx *in, const cmplx *coef, int rp, int b, int nparts, int k) {
  int n = k % b;
  int n2 = n << 1;
  cmplx s;
  rp += k / b;
  in += (rp < nparts ? rp : rp - nparts) * b;
  if (n) {
    s = prod(in[n], coef[k]);
  } else {
    s.x = in[0].x * coef[k].x;
    s.y = in[0].y * coef[k].y;
  }
  out[n2] = out[n2] + s.x;
  out[n2 + 1] = out[n2 + 1] + s.y;
}
/* cl_conv_kernels.h:120-124 (float / int -> float division) */
static void pk_olap(float *buf, const float *in, int parts, int n) {
  buf[n] = (in[n] + buf[parts + n]) / (float)parts;
  buf[parts + n] = in[parts + n];
}

struct orc_pconv {
  int N, bins, bsize, nparts, wp, wp2; /* cl_conv.cpp:143-144 */
  cmplx *w[2], *w2[2];
  int *b;
  cmplx *in1, *in2, *out; /* bins c64 each (cl_conv.cpp:232-240) */
  float *olap;            /* 2*bins floats */
  cmplx *spec1, *spec2;   /* bsize c64 each (243-246) */
};

orc_pconv *orc_pconv_create(int cvs, int pts) {
  orc_pconv *p = (orc_pconv *)calloc(1, sizeof(orc_pconv));
  p->N = pts << 1;
  p->bins = pts;
  p->nparts = cvs / pts; /* truncating, cl_conv.cpp:143 */
  p->bsize = p->nparts * p->bins;
  p->wp = 0;
  p->wp2 = p->nparts - 1;
  int bins = p->bins;
  for (int d = 0; d < 2; d++) {
    p->w[d] = (cmplx *)malloc(sizeof(cmplx) * bins);
    p->w2[d] = (cmplx *)malloc(sizeof(cmplx) * bins);
    make_twiddle(p->w[d], bins, d ? 1.f : -1.f);   /* cl_conv.cpp:263-276 */
    make_twiddle2(p->w2[d], bins, d ? 1.f : -1.f); /* cl_conv.cpp:277-287 */
  }
  p->b = (int *)malloc(sizeof(int) * bins);
  make_bitrev(p->b, bins);
  /* zero-filled state, cl_conv.cpp:303-313 (`out` is written before it is read) */
  p->in1 = (cmplx *)calloc(bins, sizeof(cmplx));
  p->in2 = (cmplx *)calloc(bins, sizeof(cmplx));
  p->out = (cmplx *)calloc(bins, sizeof(cmplx));
  p->olap = (float *)calloc(2 * bins, sizeof(float));
  p->spec1 = (cmplx *)calloc(p->bsize > 0 ? p->bsize : 1, sizeof(cmplx));
  p->spec2 = (cmplx *)calloc(p->bsize > 0 ? p->bsize : 1, sizeof(cmplx));
  return p;
}

void orc_pconv_destroy(orc_pconv *p) {
  if (!p) return;
  for (int d = 0; d < 2; d++) {
    free(p->w[d]);
    free(p->w2[d]);
  }
  free(p->b);
  free(p->in1);
  free(p->in2);
  free(p->out);
  free(p->olap);
  free(p->spec1);
  free(p->spec2);
  free(p);
}

int orc_pconv_nparts(const orc_pconv *p) { return p->nparts; }
const float *orc_pconv_spec1(const orc_pconv *p) { return (const float *)p->spec1; }
const float *orc_pconv_spec2(const orc_pconv *p) { return (const float *)p->spec2; }
const float *orc_pconv_olap(const orc_pconv *p) { return p->olap; }

/* dispatch helpers, cl_conv.cpp:36-135, as loops over work-items */
static void d_reorder(cmplx *out, cmplx *in, const int *b, int offs, int threads) {
  for (int g = 0; g < threads; g++) pk_reorder(out, in, b, offs, g);
}
static void d_fft(cmplx *data, const cmplx *w, int bins, int offs, int threads) {
  for (int n = 1; n < bins; n *= 2) { /* cl_conv.cpp:55-65 */
    int n2 = n << 1;
    for (int g = 0; g < threads; g++) pk_fft(data, w, bins, n2, offs, g);
  }
}
static void d_real_cmplx(cmplx *data, const cmplx *w, int bins, int offs, int threads) {
  for (int g = 0; g < threads; g++) pk_r2c(data, w, bins, offs, g);
}
static void d_cmplx_real(cmplx *data, const cmplx *w, int bins, int threads) {
  for (int g = 0; g < threads; g++) pk_c2r(data, w, bins, g);
}
static void d_convol(cmplx *out, const cmplx *in, const cmplx *coefs, int wp, int bins, int nparts,
                     int threads) {
  for (int g = 0; g < threads; g++) pk_convol((float *)out, in, coefs, wp, bins, nparts, g);
}
static void d_ola(float *out, const cmplx *in, int parts, int threads) {
  for (int g = 0; g < threads; g++) pk_olap(out, (const float *)in, parts, g);
}

/* cl_conv.cpp:353-388 */
int orc_pconv_push_ir(orc_pconv *p, const float *ir) {
  int bins = p->bins;
  for (int i = 0; i < p->nparts; i++) {
    memcpy(p->in2, &ir[i * bins], sizeof(float) * bins);          /* 361: half of in2 */
    d_reorder(p->spec2, p->in2, p->b, p->wp2 * bins, bins);       /* 367 */
    d_fft(p->spec2, p->w[0], bins, p->wp2 * bins, bins >> 1);     /* 373 */
    d_real_cmplx(p->spec2, p->w2[0], bins, p->wp2 * bins, bins >> 1); /* 379 */
    p->wp2 = p->wp2 == 0 ? p->nparts - 1 : p->wp2 - 1;            /* 385 */
  }
  return 0;
}

/* shared tail of both convolution variants: cl_conv.cpp:428-456 == 526-546 */
static void pconv_tail(orc_pconv *p, float *output) {
  int bins = p->bins;
  d_convol(p->in1, p->spec1, p->spec2, p->wp, bins, p->nparts, p->bsize); /* 428 */
  d_cmplx_real(p->in1, p->w2[1], bins, bins >> 1);                        /* 434 */
  d_reorder(p->out, p->in1, p->b, 0, bins);                               /* 439 */
  d_fft(p->out, p->w[1], bins, 0, bins >> 1);                             /* 444 */
  d_ola(p->olap, p->out, bins, bins);                                     /* 449 */
  memcpy(output, p->olap, sizeof(float) * bins);                          /* 455 */
}

/* cl_conv.cpp:393-458 */
int orc_pconv_convolution(orc_pconv *p, float *output, const float *input) {
  int bins = p->bins;
  memcpy(p->in1, input, sizeof(float) * bins);                         /* 399 */
  d_reorder(p->spec1, p->in1, p->b, p->wp * bins, bins);               /* 406 */
  d_fft(p->spec1, p->w[0], bins, p->wp * bins, bins >> 1);             /* 412 */
  d_real_cmplx(p->spec1, p->w2[0], bins, p->wp * bins, bins >> 1);     /* 418 */
  p->wp = p->wp != p->nparts - 1 ? p->wp + 1 : 0;                      /* 424 */
  pconv_tail(p, output);
  return 0;
}

/* cl_conv.cpp:460-548 (queue 2 work is simply done in program order) */
int orc_pconv_convolution_tv(orc_pconv *p, float *output, const float *input1, const float *input2) {
  int bins = p->bins;
  memcpy(p->in1, input1, sizeof(float) * bins);                        /* 465 */
  memcpy(p->in2, input2, sizeof(float) * bins);                        /* 472 */
  d_reorder(p->spec1, p->in1, p->b, p->wp * bins, bins);               /* 480 */
  d_reorder(p->spec2, p->in2, p->b, p->wp2 * bins, bins);              /* 487 */
  d_fft(p->spec1, p->w[0], bins, p->wp * bins, bins >> 1);             /* 493 */
  d_fft(p->spec2, p->w[0], bins, p->wp2 * bins, bins >> 1);            /* 499 */
  d_real_cmplx(p->spec1, p->w2[0], bins, p->wp * bins, bins >> 1);     /* 504 */
  d_real_cmplx(p->spec2, p->w2[0], bins, p->wp2 * bins, bins >> 1);    /* 510 */
  p->wp = p->wp != p->nparts - 1 ? p->wp + 1 : 0;                      /* 516 */
  p->wp2 = p->wp2 == 0 ? p->nparts - 1 : p->wp2 - 1;                   /* 519 */
  pconv_tail(p, output);
  return 0;
}

/* ---- dconvcode (cl_dconv.cpp:32-43) and Cldconv host code (46-153) --------------------- */

static void dk_convol(float *out, const float *del, const float *coefs, int irsize, int rp, int vsize,
                      int t) {
  float tap;
  if (t >= irsize * vsize) return;
  int n = t % vsize;
  int h = t / vsize;
  int end = irsize + vsize;
  rp += n + h;
  tap = del[rp < end ? rp : rp % end] * coefs[irsize - 1 - h];
  out[n] = out[n] + tap; /* AtomicAdd, cl_dconv.cpp:17-31 */
}

struct orc_dconv {
  int irsize, vsize, wp;
  float *buff, *coefs, *del;
};

/* cl_dconv.cpp:46-98. The reference never initialises del/coefs; zeros here,
 * which is also what oracle/minicl's clCreateBuffer hands out. */
orc_dconv *orc_dconv_create(int irsize, int vsize) {
  orc_dconv *d = (orc_dconv *)calloc(1, sizeof(orc_dconv));
  d->irsize = irsize;
  d->vsize = vsize;
  d->wp = 0;
  d->buff = (float *)calloc(vsize, sizeof(float));
  d->del = (float *)calloc(irsize + vsize, sizeof(float));
  d->coefs = (float *)calloc(irsize + vsize, sizeof(float));
  return d;
}
void orc_dconv_destroy(orc_dconv *d) {
  if (!d) return;
  free(d->buff);
  free(d->del);
  free(d->coefs);
  free(d);
}
/* cl_dconv.cpp:150-153 */
int orc_dconv_push_ir(orc_dconv *d, const float *ir) {
  memcpy(d->coefs, ir, sizeof(float) * d->irsize);
  return 0;
}
/* ring write used for both del (112-122) and coefs (136-146); returns the float count the
 * reference leaves in its `bytes` variable (the clobbered value when the ring wraps, Q10) */
static int ring_write(float *ring, int wp, int irsize, int vsize, const float *in) {
  int count = vsize;
  if (wp > irsize) {
    int front = wp - irsize;
    count = vsize - front;
    memcpy(ring + wp, in, sizeof(float) * count);
    count = front;
    memcpy(ring, &in[vsize - front], sizeof(float) * count);
  } else
    memcpy(ring + wp, in, sizeof(float) * count);
  return count;
}
/* cl_dconv.cpp:109-132, literal (including the short fill/read-back when the ring wraps) */
int orc_dconv_convolution(orc_dconv *d, float *out, const float *in) {
  int threads = d->irsize * d->vsize;
  int count = ring_write(d->del, d->wp, d->irsize, d->vsize, in);
  memset(d->buff, 0, sizeof(float) * count);        /* 123 */
  d->wp = (d->wp + d->vsize) % (d->irsize + d->vsize); /* 124 */
  for (int t = 0; t < threads; t++) dk_convol(d->buff, d->del, d->coefs, d->irsize, d->wp, d->vsize, t);
  memcpy(out, d->buff, sizeof(float) * count);      /* 130 */
  return 0;
}
/* cl_dconv.cpp:134-148 */
int orc_dconv_convolution_tv(orc_dconv *d, float *out, const float *in1, const float *in2) {
  ring_write(d->coefs, d->wp, d->irsize, d->vsize, in2);
  return orc_dconv_convolution(d, out, in1);
}

/* ---- float64 ground truth --------------------------------------------------------------- */
void orc_dft64(const double *in, double *out, int N, int sign) {
  for (int k = 0; k < N; k++) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; n++) {
      /* reduce k*n mod N first so the angle stays accurate for large N */
      long long kn = ((long long)k * n) % N;
      double a = sign * 2.0 * PI * (double)kn / N;
      double c = cos(a), s = sin(a);
      sr += in[2 * n] * c - in[2 * n + 1] * s;
      si += in[2 * n] * s + in[2 * n + 1] * c;
    }
    out[2 * k] = sr;
    out[2 * k + 1] = si;
  }
}
