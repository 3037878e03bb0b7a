/*
 * FFT behind the 7 FFTW symbols the reference uses -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference links libfftw3f (burst_detect.c:238-244,684;
 * burst_downmix.c:198-214,316-354,492,549-560; main.c:62-78).  That library is
 * not installed here, so oracle/_ref is "reference sources + this shim".
 * Semantics kept: out-of-place, unnormalised in both directions, plan bound to
 * the in/out pointers given at plan time, power-of-two sizes only.
 *
 * Algorithm: Stockham autosort, radix-4 passes plus one radix-2 pass when
 * log2(n) is odd, planar work buffers, twiddles evaluated in double and
 * rounded once.  Built with -DSHIM_DOUBLE the whole transform runs in double
 * and is rounded to float at the end (the "ideal DFT" used to bound FFT
 * rounding effects).  This is deliberately NOT the radix-2 DIF schedule the
 * CUDA kernels and oracle/ir_oracle.c share, so agreement between the two
 * chains is evidence rather than construction.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef SHIM_DOUBLE
typedef double real_t;
#else
typedef float real_t;
#endif

struct shim_pass {
    int radix;      /* 4 or 2 */
    int n;          /* sub-transform length entering this pass */
    int s;          /* stride entering this pass */
    real_t *wr;     /* (radix-1) * (n/radix) twiddle reals  */
    real_t *wi;
};

struct shim_plan_s {
    int n;
    int sign;
    fftwf_complex *in;
    fftwf_complex *out;
    int npass;
    struct shim_pass pass[16];
    real_t *buf[4];   /* xr, xi, yr, yi */
};

void *fftwf_alloc_complex(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n * sizeof(fftwf_complex) + 64) != 0)
        return NULL;
    return p;
}

void fftwf_free(void *p) { free(p); }

int fftwf_import_wisdom_from_filename(const char *filename) {
    (void)filename;
    return 0;
}

int fftwf_export_wisdom_to_filename(const char *filename) {
    (void)filename;
    return 0;
}

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
                             int sign, unsigned flags) {
    (void)flags;
    if (n < 1 || (n & (n - 1)) != 0)
        return NULL;
    struct shim_plan_s *p = calloc(1, sizeof(*p));
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    for (int b = 0; b < 4; b++)
        if (posix_memalign((void **)&p->buf[b], 64, sizeof(real_t) * (size_t)n + 64) != 0)
            return NULL;

    int len = n, stride = 1;
    while (len > 1) {
        struct shim_pass *ps = &p->pass[p->npass++];
        ps->radix = (len % 4 == 0) ? 4 : 2;
        ps->n = len;
        ps->s = stride;
        int m = len / ps->radix;
        int nt = (ps->radix - 1) * m;
        ps->wr = malloc(sizeof(real_t) * (size_t)(nt > 0 ? nt : 1));
        ps->wi = malloc(sizeof(real_t) * (size_t)(nt > 0 ? nt : 1));
        for (int k = 1; k < ps->radix; k++) {
            for (int q = 0; q < m; q++) {
                double ang = (double)sign * 2.0 * M_PI * (double)k * (double)q / (double)len;
                ps->wr[(k - 1) * m + q] = (real_t)cos(ang);
                ps->wi[(k - 1) * m + q] = (real_t)sin(ang);
            }
        }
        len /= ps->radix;
        stride *= ps->radix;
    }
    return p;
}

void fftwf_destroy_plan(fftwf_plan p) {
    if (!p) return;
    for (int i = 0; i < p->npass; i++) {
        free(p->pass[i].wr);
        free(p->pass[i].wi);
    }
    for (int b = 0; b < 4; b++)
        free(p->buf[b]);
    free(p);
}

/* One radix-4 Stockham pass: x (length n blocks, stride s) -> y. */
static void pass4(const struct shim_pass *ps, int sign,
                  const real_t *restrict xr, const real_t *restrict xi,
                  real_t *restrict yr, real_t *restrict yi) {
    const int m = ps->n / 4, s = ps->s;
    const real_t *w1r = ps->wr, *w2r = ps->wr + m, *w3r = ps->wr + 2 * m;
    const real_t *w1i = ps->wi, *w2i = ps->wi + m, *w3i = ps->wi + 2 * m;
    const real_t sg = (real_t)sign;   /* forward: multiply (b-d) by -j */
    for (int q = 0; q < m; q++) {
        const real_t a1r = w1r[q], a1i = w1i[q];
        const real_t a2r = w2r[q], a2i = w2i[q];
        const real_t a3r = w3r[q], a3i = w3i[q];
        const real_t *x0r = xr + s * (q + 0 * m), *x0i = xi + s * (q + 0 * m);
        const real_t *x1r = xr + s * (q + 1 * m), *x1i = xi + s * (q + 1 * m);
        const real_t *x2r = xr + s * (q + 2 * m), *x2i = xi + s * (q + 2 * m);
        const real_t *x3r = xr + s * (q + 3 * m), *x3i = xi + s * (q + 3 * m);
        real_t *y0r = yr + s * (4 * q + 0), *y0i = yi + s * (4 * q + 0);
        real_t *y1r = yr + s * (4 * q + 1), *y1i = yi + s * (4 * q + 1);
        real_t *y2r = yr + s * (4 * q + 2), *y2i = yi + s * (4 * q + 2);
        real_t *y3r = yr + s * (4 * q + 3), *y3i = yi + s * (4 * q + 3);
        for (int t = 0; t < s; t++) {
            real_t apcr = x0r[t] + x2r[t], apci = x0i[t] + x2i[t];
            real_t amcr = x0r[t] - x2r[t], amci = x0i[t] - x2i[t];
            real_t bpdr = x1r[t] + x3r[t], bpdi = x1i[t] + x3i[t];
            real_t bmdr = x1r[t] - x3r[t], bmdi = x1i[t] - x3i[t];
            /* sign*j*(b-d) */
            real_t jr = -sg * bmdi, ji = sg * bmdr;
            real_t t1r = amcr + jr, t1i = amci + ji;
            real_t t2r = apcr - bpdr, t2i = apci - bpdi;
            real_t t3r = amcr - jr, t3i = amci - ji;
            y0r[t] = apcr + bpdr;
            y0i[t] = apci + bpdi;
            y1r[t] = t1r * a1r - t1i * a1i;
            y1i[t] = t1r * a1i + t1i * a1r;
            y2r[t] = t2r * a2r - t2i * a2i;
            y2i[t] = t2r * a2i + t2i * a2r;
            y3r[t] = t3r * a3r - t3i * a3i;
            y3i[t] = t3r * a3i + t3i * a3r;
        }
    }
}

static void pass2(const struct shim_pass *ps,
                  const real_t *restrict xr, const real_t *restrict xi,
                  real_t *restrict yr, real_t *restrict yi) {
    const int m = ps->n / 2, s = ps->s;
    for (int q = 0; q < m; q++) {
        const real_t ar = ps->wr[q], ai = ps->wi[q];
        const real_t *x0r = xr + s * q, *x0i = xi + s * q;
        const real_t *x1r = xr + s * (q + m), *x1i = xi + s * (q + m);
        real_t *y0r = yr + s * (2 * q), *y0i = yi + s * (2 * q);
        real_t *y1r = yr + s * (2 * q + 1), *y1i = yi + s * (2 * q + 1);
        for (int t = 0; t < s; t++) {
            real_t dr = x0r[t] - x1r[t], di = x0i[t] - x1i[t];
            y0r[t] = x0r[t] + x1r[t];
            y0i[t] = x0i[t] + x1i[t];
            y1r[t] = dr * ar - di * ai;
            y1i[t] = dr * ai + di * ar;
        }
    }
}

#ifdef SHIM_DIF
/* Variant used only to prove the restatement's non-FFT arithmetic: route the reference's
 * FFT calls through oracle/ir_oracle.c's radix-2 DIF so that reference+DIF must equal the
 * restatement bit for bit (tests/test_oracle_ref.py::test_port_bit_exact_with_shared_fft). */
#include "../ir_oracle.h"
void fftwf_execute(const fftwf_plan p) {
    memcpy(p->out, p->in, sizeof(fftwf_complex) * (size_t)p->n);
    orc_fft((orc_cf32 *)p->out, p->n, p->sign > 0);
}
#else
void fftwf_execute(const fftwf_plan p) {
    const int n = p->n;
    const float *src = (const float *)p->in;
    float *dst = (float *)p->out;
    real_t *xr = p->buf[0], *xi = p->buf[1], *yr = p->buf[2], *yi = p->buf[3];
    for (int i = 0; i < n; i++) {
        xr[i] = (real_t)src[2 * i];
        xi[i] = (real_t)src[2 * i + 1];
    }
    for (int k = 0; k < p->npass; k++) {
        const struct shim_pass *ps = &p->pass[k];
        if (ps->radix == 4)
            pass4(ps, p->sign, xr, xi, yr, yi);
        else
            pass2(ps, xr, xi, yr, yi);
        real_t *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    for (int i = 0; i < n; i++) {
        dst[2 * i] = (float)xr[i];
        dst[2 * i + 1] = (float)xi[i];
    }
}
#endif
