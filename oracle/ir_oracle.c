/*
 * ir_oracle.c -- CPU restatement of the reference hot path (detect -> downmix ->
 * DQPSK demod -> RAW line).  TEST INFRASTRUCTURE ONLY: see ir_oracle.h for who may
 * load it and for the arithmetic contract.
 *
 * Parity status: PINNED against oracle/_ref (unmodified reference + FFT shim) by
 * tests/test_oracle_ref.py; golden vectors under tests/golden/.
 *
 * Layout differs from the reference on purpose: the detector works on a whole
 * recording with an explicit emulation of the feed-block cadence and ring buffer,
 * the downmix keeps every intermediate for stage-level comparison, and all three
 * stages are single functions over flat arrays.  Each block cites the reference
 * lines whose arithmetic it reproduces.  Build: see oracle/Makefile (needs
 * -ffp-contract=off so that only the fmaf() calls written here are fused -- the
 * reference's AVX2 kernels use explicit FMA intrinsics, its scalar C does not).
 */
#define _GNU_SOURCE
#include "ir_oracle.h"

#include <complex.h>
#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef float complex cf;

#define SYMBOL_RATE 25000            /* iridium.h:17 */
#define UW_LEN 12                    /* iridium.h:18 */
static const int UW_DL[UW_LEN] = {0, 2, 2, 2, 2, 0, 0, 0, 2, 0, 0, 2};   /* iridium.h:30 */
static const int UW_UL[UW_LEN] = {2, 2, 0, 0, 0, 2, 0, 0, 2, 0, 2, 2};   /* iridium.h:31 */
#define PI_F ((float)M_PI)

void orc_free(void *p) { free(p); }

static double cpu_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ===================================================================== FFT */

typedef struct { int n; float *wr, *wi; } tw_table;
static tw_table g_tw[24];

static const tw_table *twiddles(int n) {
    int lg = 0;
    while ((1 << lg) < n) lg++;
    tw_table *t = &g_tw[lg];
    if (t->n == n) return t;
    int h = n / 2 > 0 ? n / 2 : 1;
    float *wr = malloc(sizeof(float) * h), *wi = malloc(sizeof(float) * h);
    for (int k = 0; k < h; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        wr[k] = (float)cos(a);
        wi[k] = (float)(-sin(a));
    }
    wr[0] = 1.0f; wi[0] = 0.0f;
    if (n >= 4) { wr[n / 4] = 0.0f; wi[n / 4] = -1.0f; }
    t->wr = wr; t->wi = wi; t->n = n;
    return t;
}

void orc_fft(orc_cf32 *x, int n, int inverse) {
    const tw_table *t = twiddles(n);
    /* radix-2 DIF stages */
    for (int half = n / 2, step = 1; half >= 1; half >>= 1, step <<= 1) {
        for (int blk = 0; blk < n; blk += 2 * half) {
            for (int j = 0; j < half; j++) {
                orc_cf32 a = x[blk + j], b = x[blk + j + half];
                float wr = t->wr[j * step];
                float wi = inverse ? -t->wi[j * step] : t->wi[j * step];
                float dr = a.re - b.re, di = a.im - b.im;
                x[blk + j].re = a.re + b.re;
                x[blk + j].im = a.im + b.im;
                float p = di * wi;
                float q = di * wr;
                x[blk + j + half].re = fmaf(dr, wr, -p);
                x[blk + j + half].im = fmaf(dr, wi, q);
            }
        }
    }
    /* bit-reversal to natural order */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { orc_cf32 tmp = x[i]; x[i] = x[j]; x[j] = tmp; }
    }
}

/* ================================================================ detector */

/* burst_detect.c:174-299 */
void orc_det_params_init(orc_det_params *p, double center_frequency, int sample_rate,
                         int fft_size, int burst_width_hz, float threshold_db) {
    memset(p, 0, sizeof(*p));
    p->center_frequency = center_frequency;
    p->sample_rate = sample_rate;
    if (fft_size > 0) p->fft_size = fft_size;
    else p->fft_size = 1 << (int)round(log2(sample_rate / 1000.0));
    p->burst_pre_len = 2 * p->fft_size;
    p->burst_post_len = (int)(sample_rate * 16e-3);
    if (burst_width_hz <= 0) burst_width_hz = 40000;
    p->burst_width_bins = burst_width_hz / (sample_rate / p->fft_size);
    p->max_bursts = (int)((sample_rate / (float)burst_width_hz) * 0.8f);
    p->max_burst_len = (int)(sample_rate * 0.09);
    p->history_size = 512;
    p->threshold_db = threshold_db > 0 ? threshold_db : 16.0f;
    float enbw = 1.72f;
    p->threshold_lin = powf(10.0f, p->threshold_db / 10.0f) / p->history_size / enbw;
    p->ringbuf_size = (size_t)p->max_burst_len + p->burst_pre_len + p->burst_post_len
                      + (size_t)p->fft_size * 4;
    if (p->ringbuf_size < (size_t)(2 * sample_rate))
        p->ringbuf_size = 2 * (size_t)sample_rate;
}

/* window_func.c:19-24 then burst_detect.c:249-250 */
static void blackman(float *w, int n) {
    for (int i = 0; i < n; i++)
        w[i] = 0.42f - 0.5f * cosf(2.0f * PI_F * i / (n - 1))
                     + 0.08f * cosf(4.0f * PI_F * i / (n - 1));
}

void orc_det_window(float *w, int n) {
    blackman(w, n);
    for (int i = 0; i < n; i++) w[i] /= 0.42f;
}

/* burst_detect.c:679-687 with simd_avx2.c:145-165 (window) and :177-218 (shift+mag, one FMA) */
void orc_det_frame_mag(const orc_cf32 *frame, const float *window, int n, float *mag_out) {
    orc_cf32 *buf = malloc(sizeof(orc_cf32) * n);
    for (int i = 0; i < n; i++) {
        buf[i].re = frame[i].re * window[i];
        buf[i].im = frame[i].im * window[i];
    }
    orc_fft(buf, n, 0);
    int half = n / 2;
    for (int i = 0; i < n; i++) {
        orc_cf32 v = buf[(i + half) % n];
        mag_out[i] = fmaf(v.re, v.re, v.im * v.im);
    }
    free(buf);
}

void orc_convert_ci8(const int8_t *iq, size_t n, orc_cf32 *dst) {
    for (size_t i = 0; i < n; i++) {
        dst[i].re = iq[2 * i] / 128.0f;
        dst[i].im = iq[2 * i + 1] / 128.0f;
    }
}

void orc_convert_ci16(const int16_t *iq, size_t n, orc_cf32 *dst) {
    for (size_t i = 0; i < n; i++) {
        int8_t a = (int8_t)(iq[2 * i] >> 8), b = (int8_t)(iq[2 * i + 1] >> 8);
        dst[i].re = a / 128.0f;
        dst[i].im = b / 128.0f;
    }
}

typedef struct { int bin; float rel; } peak;

typedef struct {
    const orc_det_params *p;
    float *window, *hist, *base, *mag, *rel;
    unsigned char *mask;          /* 1 = free, 0 = covered by an active burst */
    int hist_idx, primed;
    orc_burst *act; int n_act, cap_act;
    orc_burst *gone; size_t n_gone, cap_gone;
    orc_burst *pending; size_t n_pending, cap_pending;
    peak *peaks; int n_peaks;
    uint64_t next_id, index;
    int squelch_count, n_squelch;
} det_state;

static void blist_push(orc_burst **arr, size_t *n, size_t *cap, const orc_burst *b) {
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 64;
        *arr = realloc(*arr, *cap * sizeof(orc_burst));
    }
    (*arr)[(*n)++] = *b;
}

static void act_push(det_state *s, const orc_burst *b) {
    if (s->n_act == s->cap_act) {
        s->cap_act = s->cap_act ? s->cap_act * 2 : 64;
        s->act = realloc(s->act, s->cap_act * sizeof(orc_burst));
    }
    s->act[s->n_act++] = *b;
}

static void act_remove(det_state *s, int i) {
    memmove(&s->act[i], &s->act[i + 1], (size_t)(s->n_act - i - 1) * sizeof(orc_burst));
    s->n_act--;
}

/* burst_detect.c:438-454 (simd_avx2.c:221-236: subtract, then add -- two roundings) */
static void baseline_push(det_state *s) {
    const int n = s->p->fft_size;
    float *h = s->hist + (size_t)s->hist_idx * n;
    for (int i = 0; i < n; i++) {
        float v = s->base[i] - h[i];
        s->base[i] = v + s->mag[i];
        h[i] = s->mag[i];
    }
    if (++s->hist_idx == s->p->history_size) { s->primed = 1; s->hist_idx = 0; }
}

/* burst_detect.c:473-486 */
static void mask_range(det_state *s, int center) {
    const int n = s->p->fft_size, hw = s->p->burst_width_bins / 2;
    int lo = center - hw, hi = center + hw;
    if (lo < 0) lo = 0;
    if (hi >= n) hi = n - 1;
    memset(s->mask + lo, 0, (size_t)(hi - lo + 1));
}

static void mask_rebuild(det_state *s) {
    memset(s->mask, 1, (size_t)s->p->fft_size);
    for (int i = 0; i < s->n_act; i++) mask_range(s, s->act[i].center_bin);
}

static int peak_desc(const void *a, const void *b) {
    const peak *x = a, *y = b;
    if (x->rel > y->rel) return -1;
    if (x->rel < y->rel) return 1;
    return x->bin - y->bin;     /* ties: ascending bin (what a stable sort of the scan order gives) */
}

/* One frame of the state machine given s->mag (burst_detect.c:689-698). */
static void det_frame(det_state *s) {
    const orc_det_params *p = s->p;
    const int n = p->fft_size;
    const float thr = p->threshold_lin;
    if (s->primed) {
        /* :426-434, simd_avx2.c:239-257 */
        for (int i = 0; i < n; i++)
            s->rel[i] = s->base[i] > 0 ? s->mag[i] / s->base[i] : 0.0f;
        /* :458-469 hysteresis on centre +-1 */
        for (int i = 0; i < s->n_act; i++) {
            int cb = s->act[i].center_bin;
            if ((cb > 0 && s->rel[cb - 1] > thr) || s->rel[cb] > thr ||
                (cb < n - 1 && s->rel[cb + 1] > thr))
                s->act[i].last_active = s->index;
        }
        /* :522-552 masked peaks, DC notch +-3, edges burst_width/2 */
        const int hw = p->burst_width_bins / 2, dc = n / 2;
        s->n_peaks = 0;
        for (int b = hw; b < n - hw; b++) {
            if (b >= dc - 3 && b <= dc + 3) continue;
            float r = s->mask[b] ? s->rel[b] : 0.0f;
            if (r > thr) { s->peaks[s->n_peaks].bin = b; s->peaks[s->n_peaks].rel = r; s->n_peaks++; }
        }
        qsort(s->peaks, (size_t)s->n_peaks, sizeof(peak), peak_desc);
        /* :490-518 retire */
        int force = 0;
        for (int i = 0; i < s->n_act;) {
            orc_burst *b = &s->act[i];
            int too_long = 0;
            if (p->max_burst_len > 0 && b->last_active - b->start > (uint64_t)p->max_burst_len) {
                force = 1; too_long = 1;
            }
            if (b->last_active + (uint64_t)p->burst_post_len <= s->index || too_long) {
                b->stop = s->index;
                blist_push(&s->gone, &s->n_gone, &s->cap_gone, b);
                act_remove(s, i);
            } else {
                i++;
            }
        }
        if (force) baseline_push(s);
        mask_rebuild(s);
        /* :556-591 new bursts, strongest first */
        int n_before = s->n_act;
        (void)n_before;
        for (int i = 0; i < s->n_peaks; i++) {
            int bin = s->peaks[i].bin;
            if (!s->mask[bin]) continue;
            orc_burst b;
            memset(&b, 0, sizeof(b));
            b.id = s->next_id;
            s->next_id += 10;
            b.center_bin = bin;
            b.peak_rel = s->peaks[i].rel;
            b.base_at_create = s->base[bin];
            b.magnitude = 10.0f * log10f(s->peaks[i].rel * p->history_size * 1.72f);
            b.start = s->index - (uint64_t)p->burst_pre_len;
            b.last_active = b.start;
            b.noise = 10.0f * log10f(s->base[bin] / p->history_size
                                     / ((float)n * n) / 1.72f / ((float)p->sample_rate / n));
            act_push(s, &b);
            mask_range(s, bin);
        }
        /* :593-631 squelch */
        if (p->max_bursts > 0 && s->n_act > p->max_bursts) {
            s->n_squelch++;
            for (int i = 0; i < s->n_act;) {
                if (s->act[i].start != s->index - (uint64_t)p->burst_pre_len) {
                    s->act[i].stop = s->index;
                    blist_push(&s->gone, &s->n_gone, &s->cap_gone, &s->act[i]);
                    act_remove(s, i);
                } else {
                    i++;
                }
            }
            s->n_act = 0;
            mask_rebuild(s);
            s->squelch_count += 3;
            if (s->squelch_count >= 10) {
                s->hist_idx = 0; s->primed = 0;
                memset(s->hist, 0, sizeof(float) * (size_t)n * p->history_size);
                memset(s->base, 0, sizeof(float) * (size_t)n);
                s->squelch_count = 0;
            }
        } else if (s->squelch_count > 0) {
            s->squelch_count--;
        }
    }
    if (s->n_act == 0) baseline_push(s);
}

size_t orc_detect(const orc_det_params *p, const orc_cf32 *iq, size_t n, size_t feed_block,
                  orc_burst **out, float *mag_dump, int *n_squelch) {
    const int N = p->fft_size;
    det_state s;
    memset(&s, 0, sizeof(s));
    s.p = p;
    s.window = malloc(sizeof(float) * N);
    orc_det_window(s.window, N);
    s.hist = calloc((size_t)N * p->history_size, sizeof(float));
    s.base = calloc(N, sizeof(float));
    s.mag = calloc(N, sizeof(float));
    s.rel = calloc(N, sizeof(float));
    s.mask = malloc(N);
    memset(s.mask, 1, N);
    s.peaks = malloc(sizeof(peak) * N);
    if (feed_block == 0) feed_block = 32768;

    uint64_t sample_count = 0, ring_start = 0;
    size_t frame_no = 0;
    for (size_t off = 0; off < n; off += feed_block) {
        size_t m = n - off < feed_block ? n - off : feed_block;
        /* ringbuf_write computes ringbuf_start from the count BEFORE this block (:396-398 vs :774) */
        if (sample_count > p->ringbuf_size) ring_start = sample_count - p->ringbuf_size;
        sample_count += m;
        while (s.index + (uint64_t)N <= sample_count) {
            orc_det_frame_mag(iq + s.index, s.window, N, s.mag);
            if (mag_dump) memcpy(mag_dump + frame_no * N, s.mag, sizeof(float) * N);
            det_frame(&s);
            s.index += N;
            frame_no++;
        }
        /* emit at the end of the feed call (:839-841) */
        for (size_t i = 0; i < s.n_gone; i++) {
            orc_burst b = s.gone[i];
            b.emit_count = sample_count;
            b.ring_start = ring_start;
            if (orc_burst_num_samples(p, &b) == 0) continue;
            blist_push(&s.pending, &s.n_pending, &s.cap_pending, &b);
        }
        s.n_gone = 0;
    }
    if (n_squelch) *n_squelch = s.n_squelch;
    *out = s.pending;
    free(s.window); free(s.hist); free(s.base); free(s.mag); free(s.rel);
    free(s.mask); free(s.peaks); free(s.act); free(s.gone);
    return s.n_pending;
}

/* burst_detect.c:401-411, :708-712 */
size_t orc_burst_num_samples(const orc_det_params *p, const orc_burst *b) {
    uint64_t start = b->start, stop = b->stop + (uint64_t)p->burst_pre_len;
    if (start < b->ring_start) start = b->ring_start;
    if (stop <= start) return 0;
    return (size_t)(stop - start);
}

size_t orc_burst_extract(const orc_det_params *p, const orc_cf32 *iq, size_t n,
                         const orc_burst *b, orc_cf32 *dst) {
    uint64_t start = b->start;
    if (start < b->ring_start) start = b->ring_start;
    size_t len = orc_burst_num_samples(p, b);
    const uint64_t R = p->ringbuf_size;
    for (size_t i = 0; i < len; i++) {
        uint64_t pos = start + i;
        if (pos >= b->emit_count) {
            /* not yet written: the ring slot still holds the sample one lap back, or the
             * untouched (zero) allocation during the first lap (SURVEY.md D10 ii) */
            if (pos >= R) pos -= R;
            else { dst[i].re = 0; dst[i].im = 0; continue; }
        }
        if (pos < n) dst[i] = iq[pos];
        else { dst[i].re = 0; dst[i].im = 0; }
    }
    return len;
}

/* ================================================================= filters */

/* fir_filter.c:143-182 */
static float *design_lpf(int *ntaps_out, float gain, float fs, float cutoff, float trans) {
    int nt = (int)(4.0f / (trans / fs));
    nt |= 1;
    float *h = malloc(sizeof(float) * nt);
    int c = nt / 2;
    float wc = 2.0f * PI_F * cutoff / fs;
    float total = 0;
    for (int i = 0; i < nt; i++) {
        float k = i - c;
        float s = fabsf(k) < 1e-10f ? wc / PI_F : sinf(wc * k) / (PI_F * k);
        float w = 0.35875f - 0.48829f * cosf(2.0f * PI_F * i / (nt - 1))
                           + 0.14128f * cosf(4.0f * PI_F * i / (nt - 1))
                           - 0.01168f * cosf(6.0f * PI_F * i / (nt - 1));
        h[i] = s * w;
        total += h[i];
    }
    if (fabsf(total) > 0) {
        float sc = gain / total;
        for (int i = 0; i < nt; i++) h[i] *= sc;
    }
    *ntaps_out = nt;
    return h;
}

/* fir_filter.c:74-111 */
static float *design_rrc(int *ntaps_out, float gain, float fs, float sym, float alpha, int nt) {
    nt |= 1;
    float *h = malloc(sizeof(float) * nt);
    float sps = fs / sym;
    int c = nt / 2;
    float e = 0;
    for (int i = 0; i < nt; i++) {
        float t = (i - c) / sps;
        if (fabsf(t) < 1e-10f) {
            h[i] = (1.0f - alpha + 4.0f * alpha / PI_F);
        } else if (fabsf(fabsf(t) - 1.0f / (4.0f * alpha)) < 1e-6f) {
            h[i] = alpha / sqrtf(2.0f) *
                   ((1.0f + 2.0f / PI_F) * sinf(PI_F / (4.0f * alpha)) +
                    (1.0f - 2.0f / PI_F) * cosf(PI_F / (4.0f * alpha)));
        } else {
            float num = sinf(PI_F * t * (1.0f - alpha)) +
                        4.0f * alpha * t * cosf(PI_F * t * (1.0f + alpha));
            float den = PI_F * t * (1.0f - (4.0f * alpha * t) * (4.0f * alpha * t));
            h[i] = num / den;
        }
        e += h[i] * h[i];
    }
    float sc = gain / sqrtf(e);
    for (int i = 0; i < nt; i++) h[i] *= sc;
    *ntaps_out = nt;
    return h;
}

static float sinc_f(float x) {                      /* fir_filter.c:67-70 */
    if (fabsf(x) < 1e-10f) return 1.0f;
    return sinf(PI_F * x) / (PI_F * x);
}

/* fir_filter.c:115-139 */
static float *design_rc(int *ntaps_out, float fs, float sym, float alpha, int nt) {
    nt |= 1;
    float *h = malloc(sizeof(float) * nt);
    float sps = fs / sym;
    int c = nt / 2;
    for (int i = 0; i < nt; i++) {
        float t = (i - c) / sps;
        if (fabsf(t) < 1e-10f) {
            h[i] = 1.0f;
        } else if (alpha > 0 && fabsf(fabsf(t) - 1.0f / (2.0f * alpha)) < 1e-6f) {
            h[i] = PI_F / (4.0f) * sinc_f(1.0f / (2.0f * alpha));
        } else {
            float ct = cosf(PI_F * alpha * t);
            float den = 1.0f - (2.0f * alpha * t) * (2.0f * alpha * t);
            h[i] = sinc_f(t) * ct / den;
        }
    }
    *ntaps_out = nt;
    return h;
}

/* Tail arithmetic of the reference's AVX2 FIR kernels AS COMPILED (gcc 13, -O3 -mavx2 -mfma):
 * the scalar remainder loops (simd_avx2.c:45-54, :130-136) are vectorised in-order
 * ("fold-left"): taps are consumed in chunks of 8, then one chunk of 4 if at least 4 remain,
 * each product rounded on its own and added to the accumulator in tap order; only the last
 * <4 taps are fused.  This affects the final n%4 (complex) / n%8 (real) outputs of a call. */
static inline float tail_mac(const float *h, int nt, const float *x, int stride) {
    float a = 0;
    int k = 0, k8 = nt & ~7;
    for (; k < k8; k++) a = a + h[k] * x[(size_t)k * stride];
    if (nt - k >= 4)
        for (int e = k + 4; k < e; k++) a = a + h[k] * x[(size_t)k * stride];
    for (; k < nt; k++) a = fmaf(h[k], x[(size_t)k * stride], a);
    return a;
}

/* "valid" complex FIR: one sequential FMA chain per output for the 4-at-a-time body
 * (simd_avx2.c:28-44), tail_mac for the remainder outputs. */
static void fir_cc(const float *h, int nt, const cf *in, cf *out, int n) {
    int body = n & ~3;
    for (int i = 0; i < body; i++) {
        float ar = 0, ai = 0;
        for (int k = 0; k < nt; k++) {
            ar = fmaf(h[k], crealf(in[i + k]), ar);
            ai = fmaf(h[k], cimagf(in[i + k]), ai);
        }
        out[i] = ar + ai * I;
    }
    for (int i = body; i < n; i++) {
        const float *x = (const float *)(in + i);
        out[i] = tail_mac(h, nt, x, 2) + tail_mac(h, nt, x + 1, 2) * I;
    }
}

/* real FIR, 8-at-a-time body (simd_avx2.c:117-128), tail_mac for the rest */
static void fir_ff(const float *h, int nt, const float *in, float *out, int n) {
    int body = n & ~7;
    for (int i = 0; i < body; i++) {
        float a = 0;
        for (int k = 0; k < nt; k++) a = fmaf(h[k], in[i + k], a);
        out[i] = a;
    }
    for (int i = body; i < n; i++) out[i] = tail_mac(h, nt, in + i, 1);
}

/* Decimating FIR: four interleaved FMA chains over taps k = j mod 4, combined as
 * (c0+c2)+(c1+c3), then the leftover taps one FMA each (simd_avx2.c:62-110). */
static void fir_cc_dec(const float *h, int nt, const cf *in, cf *out, int n_out, int dec) {
    for (int o = 0; o < n_out; o++) {
        const cf *p = in + (size_t)o * dec;
        float cr[4] = {0, 0, 0, 0}, ci[4] = {0, 0, 0, 0};
        int k = 0;
        for (; k + 3 < nt; k += 4)
            for (int j = 0; j < 4; j++) {
                cr[j] = fmaf(h[k + j], crealf(p[k + j]), cr[j]);
                ci[j] = fmaf(h[k + j], cimagf(p[k + j]), ci[j]);
            }
        float ar = (cr[0] + cr[2]) + (cr[1] + cr[3]);
        float ai = (ci[0] + ci[2]) + (ci[1] + ci[3]);
        for (; k < nt; k++) {
            ar = fmaf(h[k], crealf(p[k]), ar);
            ai = fmaf(h[k], cimagf(p[k]), ai);
        }
        out[o] = ar + ai * I;
    }
}

/* rotator.h:36-46: sequential phase recurrence, plain (unfused) complex products */
static cf rotate_run(cf phase, cf incr, cf *dst, const cf *src, int n) {
    for (int i = 0; i < n; i++) {
        dst[i] = src[i] * phase;
        phase *= incr;
    }
    return phase;
}

/* ================================================================= downmix */

#define CFO_OVERSAMPLE 16
#define DM_WORK (2 * 1024 * 1024)     /* burst_downmix.c:366 */

struct orc_downmix {
    int out_rate, search_depth, pre_start;
    float sps;
    float *h_in, *h_noise, *h_box, *h_rrc, *h_rc;
    int n_in, n_noise, n_box, n_rrc, n_rc;
    int cfo_n, cfo_total;
    float *cfo_win;
    int corr_n, sync_search;
    orc_cf32 *sync_dl, *sync_ul;
    int sync_dl_len, sync_ul_len;
    cf *wa, *wb;
    float *mf, *mff;
    orc_cf32 *fa, *fb, *fc;
};

/* burst_downmix.c:138-219 */
static orc_cf32 *make_sync(orc_downmix *dm, const int *uw, int pre, int uplink, int *len_out) {
    const cf s0 = 1.0f + 1.0f * I, s1 = -1.0f - 1.0f * I;
    int nsym = pre + UW_LEN, isps = (int)roundf(dm->sps);
    int plen = nsym * isps - (isps - 1);
    int half = (dm->n_rc - 1) / 2;
    cf *buf = calloc((size_t)plen + dm->n_rc - 1, sizeof(cf));
    for (int i = 0; i < nsym; i++) {
        cf v;
        if (i < pre) v = uplink ? ((i % 2 == 0) ? s1 : s0) : s0;
        else v = uw[i - pre] == 0 ? s0 : s1;
        buf[half + i * isps] = v;
    }
    cf *shaped = malloc(sizeof(cf) * plen);
    fir_cc(dm->h_rc, dm->n_rc, buf, shaped, plen);
    free(buf);
    orc_cf32 *tpl = calloc(dm->corr_n, sizeof(orc_cf32));
    for (int i = 0; i < plen && i < dm->corr_n; i++) {     /* reversed + conjugated */
        cf v = conjf(shaped[plen - 1 - i]);
        tpl[i].re = crealf(v); tpl[i].im = cimagf(v);
    }
    free(shaped);
    orc_fft(tpl, dm->corr_n, 0);
    *len_out = plen;
    return tpl;
}

static int pow2_ceil(int n) { int p = 1; while (p < n) p <<= 1; return p; }

orc_downmix *orc_downmix_create(void) {
    orc_downmix *dm = calloc(1, sizeof(*dm));
    dm->out_rate = 10 * SYMBOL_RATE;                         /* burst_downmix.c:227-234 */
    dm->sps = (float)dm->out_rate / SYMBOL_RATE;
    dm->search_depth = dm->out_rate;
    dm->pre_start = (int)(100 * 1e-6f * dm->out_rate);       /* :241 */
    dm->h_in = design_lpf(&dm->n_in, 1.0f, 10000000.0f, dm->out_rate * 0.4f, dm->out_rate * 0.2f);
    dm->h_noise = design_lpf(&dm->n_noise, 1.0f, (float)dm->out_rate, 40000.0f / 2.0f, 40000.0f);
    {
        int bl = (int)(dm->sps * 2);
        if (bl < 3) bl = 3;
        dm->n_box = bl;
        dm->h_box = malloc(sizeof(float) * bl);
        float v = 1.0f / bl;
        for (int i = 0; i < bl; i++) dm->h_box[i] = v;
    }
    dm->h_rrc = design_rrc(&dm->n_rrc, 1.0f, (float)dm->out_rate, (float)SYMBOL_RATE, 0.4f, 51);
    dm->h_rc = design_rc(&dm->n_rc, (float)dm->out_rate, (float)SYMBOL_RATE, 0.4f, 51);
    {
        int raw = (int)(dm->sps * 26);                      /* :309-315 */
        dm->cfo_n = 1;
        while (dm->cfo_n * 2 <= raw) dm->cfo_n *= 2;
        dm->cfo_total = dm->cfo_n * CFO_OVERSAMPLE;
        dm->cfo_win = malloc(sizeof(float) * dm->cfo_n);
        blackman(dm->cfo_win, dm->cfo_n);
    }
    dm->sync_search = (int)((64 + UW_LEN + 8) * dm->sps);    /* :324-330 */
    dm->corr_n = pow2_ceil(dm->sync_search + (int)((16 + UW_LEN) * dm->sps));
    dm->sync_dl = make_sync(dm, UW_DL, 16, 0, &dm->sync_dl_len);
    dm->sync_ul = make_sync(dm, UW_UL, 16, 1, &dm->sync_ul_len);
    dm->wa = malloc(sizeof(cf) * DM_WORK);
    dm->wb = malloc(sizeof(cf) * DM_WORK);
    dm->mf = malloc(sizeof(float) * DM_WORK / 8);
    dm->mff = malloc(sizeof(float) * DM_WORK / 8);
    dm->fa = malloc(sizeof(orc_cf32) * 4096);
    dm->fb = malloc(sizeof(orc_cf32) * 4096);
    dm->fc = malloc(sizeof(orc_cf32) * 4096);
    return dm;
}

void orc_downmix_destroy(orc_downmix *dm) {
    if (!dm) return;
    free(dm->h_in); free(dm->h_noise); free(dm->h_box); free(dm->h_rrc); free(dm->h_rc);
    free(dm->cfo_win); free(dm->sync_dl); free(dm->sync_ul);
    free(dm->wa); free(dm->wb); free(dm->mf); free(dm->mff);
    free(dm->fa); free(dm->fb); free(dm->fc);
    free(dm);
}

const float *orc_downmix_taps(const orc_downmix *dm, int which, int *ntaps) {
    switch (which) {
    case 0: *ntaps = dm->n_in; return dm->h_in;
    case 1: *ntaps = dm->n_noise; return dm->h_noise;
    case 2: *ntaps = dm->n_box; return dm->h_box;
    case 3: *ntaps = dm->n_rrc; return dm->h_rrc;
    default: *ntaps = dm->n_rc; return dm->h_rc;
    }
}

const orc_cf32 *orc_downmix_sync_fft(const orc_downmix *dm, int uplink, int *sync_len) {
    *sync_len = uplink ? dm->sync_ul_len : dm->sync_dl_len;
    return uplink ? dm->sync_ul : dm->sync_dl;
}

const float *orc_downmix_cfo_window(const orc_downmix *dm, int *n) {
    *n = dm->cfo_n;
    return dm->cfo_win;
}

static float quad_peak(float a, float b, float c) {      /* burst_downmix.c:526-528 / :623-625 */
    float den = a - 2.0f * b + c;
    if (fabsf(den) > 1e-10f) return 0.5f * (a - c) / den;
    return 0;
}

int orc_downmix_process(orc_downmix *dm, const orc_burst_hdr *hdr, const orc_cf32 *samples,
                        size_t num_samples, orc_frame_info *info, orc_cf32 *frame_out,
                        orc_cf32 *dec_out, orc_cf32 *nlpf_out, orc_cf32 *rrc_out) {
    memset(info, 0, sizeof(*info));
    info->id = hdr->id;
    if (num_samples < 100) { info->fail_stage = 1; return 0; }           /* :645 */
    int n = (int)num_samples;
    if (n > DM_WORK) n = DM_WORK;
    const cf *src = (const cf *)samples;
    double cfreq = hdr->center_frequency;
    const int fs = hdr->sample_rate;
    uint64_t ts = hdr->start_time_ns + (uint64_t)((double)hdr->start / fs * 1e9);   /* :659-660 */

    /* 1: coarse shift (:663-672) */
    float rel = (hdr->center_bin - hdr->fft_size / 2) / (float)hdr->fft_size;
    {
        float ph = -2.0f * PI_F * rel;
        cf incr = cexpf(ph * I);
        info->incr_coarse_re = crealf(incr); info->incr_coarse_im = cimagf(incr);
        rotate_run(1.0f, incr, dm->wa, src, n);
        cfreq += rel * fs;
    }
    /* 2: decimate (:417-437) */
    int dec = (int)roundf((float)fs / dm->out_rate);
    if (dec < 1) dec = 1;
    int dlen = (n - dm->n_in + 1) / dec;
    if (dlen > DM_WORK) dlen = DM_WORK;
    if (dlen > 0) {
        fir_cc_dec(dm->h_in, dm->n_in, dm->wa, dm->wb, dlen, dec);
        ts += (uint64_t)((dm->n_in / 2) * 1000000000ULL / fs);
    } else {
        dlen = 0;
    }
    info->dec_len = dlen;
    if (dlen < 100) { info->fail_stage = 2; return 0; }
    if (dec_out) memcpy(dec_out, dm->wb, sizeof(cf) * dlen);
    /* 2b: noise LPF, centred (:683-698) */
    {
        int hn = (dm->n_noise - 1) / 2;
        memset(dm->wa, 0, sizeof(cf) * ((size_t)dlen + dm->n_noise - 1));
        memcpy(dm->wa + hn, dm->wb, sizeof(cf) * dlen);
        fir_cc(dm->h_noise, dm->n_noise, dm->wa, dm->wb, dlen);
        memcpy(dm->wa, dm->wb, sizeof(cf) * dlen);
    }
    if (nlpf_out) memcpy(nlpf_out, dm->wa, sizeof(cf) * dlen);
    /* 3: burst start (:441-478) */
    int start;
    {
        int search = dm->search_depth < dlen ? dm->search_depth : dlen;
        int mlen = search + dm->n_box - 1;
        if (mlen > dlen) mlen = dlen;
        for (int i = 0; i < mlen; i++) {                      /* simd_avx2.c:297-318 */
            float re = crealf(dm->wa[i]), im = cimagf(dm->wa[i]);
            dm->mf[i] = fmaf(re, re, im * im);
        }
        int flen = mlen - dm->n_box + 1;
        if (flen <= 0) {
            start = 0;
        } else {
            if (flen > search) flen = search;
            fir_ff(dm->h_box, dm->n_box, dm->mf, dm->mff, flen);
            float mx = -1e30f;
            for (int i = 0; i < flen; i++) if (dm->mff[i] > mx) mx = dm->mff[i];
            float th = 0.45f * mx;
            for (start = 0; start < flen; start++)
                if (dm->mff[start] >= th) break;
            if (start > 0) {
                start = start + (dm->n_box - 1) / 2 - dm->pre_start;
                if (start < 0) start = 0;
            }
        }
    }
    info->start = start;
    if (start >= dlen - 100) { info->fail_stage = 3; return 0; }
    int flen = dlen - start;
    /* 4: fine CFO from the squared signal (:482-535) */
    float coff;
    {
        int m = dm->cfo_n < flen ? dm->cfo_n : flen;
        memset(dm->fa, 0, sizeof(orc_cf32) * dm->cfo_total);
        for (int i = 0; i < m; i++) {                         /* simd_avx2.c:345-388 */
            float a = crealf(dm->wa[start + i]), b = cimagf(dm->wa[start + i]);
            float sr = fmaf(a, a, -(b * b));   /* gcc fuses the mul/sub intrinsics: vfmsub231ps */
            float si = 2.0f * (a * b);
            dm->fa[i].re = sr * dm->cfo_win[i];
            dm->fa[i].im = si * dm->cfo_win[i];
        }
        orc_fft(dm->fa, dm->cfo_total, 0);
        float best = 0; int bi = 0;
        for (int i = 0; i < dm->cfo_total; i++) {
            float re = dm->fa[i].re, im = dm->fa[i].im;
            float v = re * re + im * im;
            if (v > best) { best = v; bi = i; }
        }
        int T = dm->cfo_total;
        int ui = bi >= T / 2 ? bi - T : bi;
        float corr = 0;
        if (bi > 0 && bi < T - 1) {
            int im1 = ui - 1 < 0 ? ui - 1 + T : ui - 1;
            int ip1 = ui + 1 < 0 ? ui + 1 + T : ui + 1;
            float a = dm->fa[im1].re * dm->fa[im1].re + dm->fa[im1].im * dm->fa[im1].im;
            float c = dm->fa[ip1].re * dm->fa[ip1].re + dm->fa[ip1].im * dm->fa[ip1].im;
            corr = quad_peak(a, best, c);
        }
        coff = (ui + corr) / T / 2.0f;
        info->cfo_peak_bin = bi;
    }
    info->center_offset = coff;
    /* 5: fine shift (:713-720) */
    {
        float ph = -2.0f * PI_F * coff;
        cf incr = cexpf(ph * I);
        info->incr_fine_re = crealf(incr); info->incr_fine_im = cimagf(incr);
        rotate_run(1.0f, incr, dm->wb, dm->wa + start, flen);
        cfreq += coff * dm->out_rate;
    }
    /* 6: matched filter, centred (:723-734) */
    {
        int hr = (dm->n_rrc - 1) / 2;
        memset(dm->wa, 0, sizeof(cf) * ((size_t)flen + dm->n_rrc - 1));
        memcpy(dm->wa + hr, dm->wb, sizeof(cf) * flen);
        fir_cc(dm->h_rrc, dm->n_rrc, dm->wa, dm->wb, flen);
    }
    if (rrc_out) memcpy(rrc_out, dm->wb, sizeof(cf) * flen);
    /* 7: sync correlation (:539-639) */
    int dir, uw_start;
    float uwc;
    cf cres;
    {
        int sl = dm->sync_search < flen ? dm->sync_search : flen;
        memset(dm->fa, 0, sizeof(orc_cf32) * dm->corr_n);
        memcpy(dm->fa, dm->wb, sizeof(cf) * sl);
        orc_fft(dm->fa, dm->corr_n, 0);
        const cf *F = (const cf *)dm->fa, *TD = (const cf *)dm->sync_dl, *TU = (const cf *)dm->sync_ul;
        cf *PD = (cf *)dm->fb, *PU = (cf *)dm->fc;
        for (int i = 0; i < dm->corr_n; i++) { PD[i] = F[i] * TD[i]; PU[i] = F[i] * TU[i]; }
        orc_fft(dm->fb, dm->corr_n, 1);
        orc_fft(dm->fc, dm->corr_n, 1);
        float md = 0, mu = 0; int od = 0, ou = 0;
        for (int i = 0; i < sl; i++) {
            float v = dm->fb[i].re * dm->fb[i].re + dm->fb[i].im * dm->fb[i].im;
            if (v > md) { md = v; od = i; }
        }
        for (int i = 0; i < sl; i++) {
            float v = dm->fc[i].re * dm->fc[i].re + dm->fc[i].im * dm->fc[i].im;
            if (v > mu) { mu = v; ou = i; }
        }
        const orc_cf32 *R; int co, slen;
        if (md >= mu) { dir = 1; co = od; R = dm->fb; slen = dm->sync_dl_len; }
        else { dir = 2; co = ou; R = dm->fc; slen = dm->sync_ul_len; }
        cres = R[co].re + R[co].im * I;
        uwc = 0;
        if (co > 0 && co < sl - 1) {
            float a = R[co - 1].re * R[co - 1].re + R[co - 1].im * R[co - 1].im;
            float b = R[co].re * R[co].re + R[co].im * R[co].im;
            float c = R[co + 1].re * R[co + 1].re + R[co + 1].im * R[co + 1].im;
            uwc = quad_peak(a, b, c);
        }
        int pre_syms = dir == 1 ? 16 : 32;                    /* :633-634 quirk kept */
        uw_start = co - slen + 1 + (int)(pre_syms * dm->sps);
        info->corr_offset = co;
    }
    info->uw_start_idx = uw_start;
    info->corr_re = crealf(cres); info->corr_im = cimagf(cres);
    info->direction = dir;
    info->uw_start = uwc;
    if (uw_start < 0 || uw_start >= flen) { info->fail_stage = 7; return 0; }
    /* 8: phase alignment (:750-760) */
    {
        float mg = cabsf(cres);
        cf pc = mg > 0 ? conjf(cres / mg) : 1.0f;
        rotate_run(pc, 1.0f, dm->wa, dm->wb, flen);
    }
    /* 9: extraction (:763-793) */
    int maxl, minl;
    if (cfreq > 1626000000) { maxl = (int)(444 * dm->sps); minl = (int)(80 * dm->sps); }
    else { maxl = (int)(191 * dm->sps); minl = (int)(131 * dm->sps); }
    int avail = flen - uw_start;
    if (avail < minl) { info->fail_stage = 9; return 0; }
    int xl = avail < maxl ? avail : maxl;
    info->ok = 1;
    info->timestamp = ts + (uint64_t)((double)start / dm->out_rate * 1e9);
    info->center_frequency = cfreq;
    info->sample_rate = (float)dm->out_rate;
    info->samples_per_symbol = dm->sps;
    info->magnitude = hdr->magnitude;
    info->noise = hdr->noise;
    info->num_samples = xl;
    memcpy(frame_out, dm->wa + uw_start, sizeof(cf) * xl);
    return 1;
}

/* =================================================================== demod */

/* qpsk_demod.c:56-81 */
static cf catmull(const cf *in, int n, float pos) {
    int idx = (int)pos;
    float mu = pos - idx;
    if (idx < 1) idx = 1;
    if (idx >= n - 2) idx = n - 3;
    cf s0 = in[idx - 1], s1 = in[idx], s2 = in[idx + 1], s3 = in[idx + 2];
    float mu2 = mu * mu, mu3 = mu2 * mu;
    cf a = -0.5f * s0 + 1.5f * s1 - 1.5f * s2 + 0.5f * s3;
    cf b = s0 - 2.5f * s1 + 2.0f * s2 - 0.5f * s3;
    cf c = -0.5f * s0 + 0.5f * s2;
    cf d = s1;
    return a * mu3 + b * mu2 + c * mu + d;
}

/* qpsk_demod.c:85-130 */
static int timing_gardner(const cf *in, int n, float sps, cf *out) {
    int k = 0;
    float pos = 0.0f, integ = 0.0f;
    cf prev = 0;
    while (pos < n - 3) {
        cf now = catmull(in, n, pos);
        out[k] = now;
        if (k > 0) {
            float mp = pos - sps * 0.5f;
            if (mp >= 1.0f) {
                cf mid = catmull(in, n, mp);
                cf df = prev - now;
                float e = crealf(df * conjf(mid));
                if (e > 1.0f) e = 1.0f;
                if (e < -1.0f) e = -1.0f;
                integ += 0.0002f * e;
                float adj = 0.02f * e + integ;
                if (adj > 0.5f) adj = 0.5f;
                if (adj < -0.5f) adj = -0.5f;
                pos += adj;
            }
        }
        prev = now;
        k++;
        pos += sps;
    }
    return k;
}

int orc_demod(const orc_cf32 *frame, int num_samples, float samples_per_symbol,
              double center_frequency, int direction, int use_gardner,
              orc_demod_info *info, uint8_t *bits_out, float *llr_out, orc_cf32 *pll_dump) {
    memset(info, 0, sizeof(*info));
    info->direction = direction;
    const cf *in = (const cf *)frame;
    int sps = (int)(samples_per_symbol + 0.5f);
    if (sps < 1) sps = 1;
    int cap = num_samples / sps + 1;
    cf *sy = malloc(sizeof(cf) * cap), *pl = malloc(sizeof(cf) * cap);
    int *sym = malloc(sizeof(int) * cap);
    float *off = malloc(sizeof(float) * cap), *mg = malloc(sizeof(float) * cap);
    int ns;
    if (use_gardner) {
        ns = timing_gardner(in, num_samples, samples_per_symbol, sy);
    } else {                                                   /* :134-141 */
        ns = 0;
        for (int i = 0; i < num_samples; i += (int)samples_per_symbol) sy[ns++] = in[i];
    }
    info->n_raw_symbols = ns;
    /* PLL (:145-195) */
    float total = 0.0f;
    {
        const float r = 0.70710678118654752f;
        cf ph = 1.0f + 0.0f * I;
        for (int i = 0; i < ns; i++) {
            pl[i] = sy[i] * ph;
            float re = crealf(pl[i]), im = cimagf(pl[i]);
            cf ideal;
            if (re >= 0 && im >= 0) ideal = r + r * I;
            else if (re >= 0) ideal = r - r * I;
            else if (im < 0) ideal = -r - r * I;
            else ideal = -r + r * I;
            cf er = conjf(ideal) * pl[i];
            float em = cabsf(er);
            if (em < 1e-10f) continue;
            cf unit = er / em;
            float ang = cargf(unit);
            float sa = 0.2f * ang;
            cf corr = cosf(sa) + sinf(sa) * I;
            total += sa;
            ph = conjf(corr) * ph;
            float pm = cabsf(ph);
            if (pm > 0) ph /= pm;
        }
    }
    info->total_phase = total;
    if (pll_dump) memcpy(pll_dump, pl, sizeof(cf) * ns);
    /* hard decisions, end of frame, confidence (:199-260) */
    int nv = 0, conf; float level;
    {
        float mx = 0; int low = 0;
        for (int i = 0; i < ns; i++) {
            float re = crealf(pl[i]), im = cimagf(pl[i]);
            float m = sqrtf(re * re + im * im);
            mg[i] = m;
            if (m > mx) mx = m;
            if (re >= 0 && im >= 0) sym[i] = 0;
            else if (re < 0 && im >= 0) sym[i] = 1;
            else if (re < 0) sym[i] = 2;
            else sym[i] = 3;
            float phs = (atan2f(im, re) + PI_F) * 180.0f / PI_F;
            off[i] = 45.0f - fmodf(phs, 90.0f);
            nv++;
            if (m < mx / 8.0f) {
                if (++low >= 3) { nv -= 3; break; }
            } else {
                low = 0;
            }
        }
        int okc = 0; float sum = 0;
        for (int i = 0; i < nv; i++) {
            sum += mg[i];
            if (fabsf(off[i]) <= 22) okc++;
        }
        level = nv > 0 ? sum / nv : 0;
        conf = nv > 0 ? (100 * okc) / nv : 0;
    }
    /* unique word, hard then soft (:277-325, :429-465) */
    int accept = 1;
    {
        int okd = 0, oku = 0;
        if (nv >= UW_LEN) {
            int dd = 0, du = 0;
            for (int i = 0; i < UW_LEN; i++) {
                int a = abs(sym[i] - UW_DL[i]); if (a == 3) a = 1; dd += a;
                int b = abs(sym[i] - UW_UL[i]); if (b == 3) b = 1; du += b;
            }
            okd = dd <= 2; oku = du <= 2;
        }
        if (!okd && !oku) {
            float ed = 999.0f, eu = 999.0f;
            if (nv >= UW_LEN) {
                ed = 0; eu = 0;
                for (int i = 0; i < UW_LEN; i++) {
                    float act = cargf(pl[i]);
                    if (act < 0) act += 2.0f * PI_F;
                    float xd = PI_F * 0.25f + UW_DL[i] * PI_F * 0.5f;
                    float d1 = act - xd;
                    if (d1 > PI_F) d1 -= 2.0f * PI_F;
                    if (d1 < -PI_F) d1 += 2.0f * PI_F;
                    ed += fabsf(d1) * (float)(2.0 / M_PI);
                    float xu = PI_F * 0.25f + UW_UL[i] * PI_F * 0.5f;
                    float d2 = act - xu;
                    if (d2 > PI_F) d2 -= 2.0f * PI_F;
                    if (d2 < -PI_F) d2 += 2.0f * PI_F;
                    eu += fabsf(d2) * (float)(2.0 / M_PI);
                }
            }
            float em = ed < eu ? ed : eu;
            if (em > 3.0f) accept = 0;
            else info->direction = eu < ed ? 2 : 1;
        } else if (oku && !okd) {
            info->direction = 2;
        } else if (okd && !oku) {
            info->direction = 1;
        }
    }
    if (accept) {
        static const int dq[4] = {0, 2, 3, 1};                 /* :46, :264-273 */
        int old = 0;
        for (int i = 0; i < nv; i++) {
            int s = sym[i], d = (s - old + 4) % 4;
            old = s;
            int v = dq[d];
            bits_out[2 * i] = (v >> 1) & 1;                    /* :329-335 */
            bits_out[2 * i + 1] = v & 1;
        }
        if (llr_out) {                                         /* :489-503 */
            float sm = 0;
            for (int i = 0; i < nv; i++) sm += cabsf(pl[i]);
            float sc = (nv > 0 && sm > 0) ? (0.70710678118654752f / (sm / nv)) : 1.0f;
            for (int i = 0; i < nv; i++) {
                llr_out[2 * i] = fabsf(crealf(pl[i])) * sc;
                llr_out[2 * i + 1] = fabsf(cimagf(pl[i])) * sc;
            }
        }
        info->ok = 1;
        info->confidence = conf;
        info->level = level;
        info->n_symbols = nv;
        info->n_payload_symbols = nv - UW_LEN;
        info->n_bits = 2 * nv;
        if (nv > 0) {                                          /* :521-527 */
            double dur = (double)nv / SYMBOL_RATE;
            info->center_frequency = center_frequency + total / dur / M_PI / 2.0;
        } else {
            info->center_frequency = center_frequency;
        }
    }
    free(sy); free(pl); free(sym); free(off); free(mg);
    return accept;
}

/* ================================================================ RAW line */

int orc_format_raw(char *dst, size_t cap, const char *file_info, uint64_t t0,
                   uint64_t timestamp, double center_frequency, float magnitude, float noise,
                   uint64_t id, int confidence, float level, int n_payload_symbols,
                   const uint8_t *bits, int n_bits) {
    double ts_ms = (double)(timestamp - t0) / 1000000.0;
    int fhz = (int)(center_frequency + 0.5);
    if (n_payload_symbols < 0) n_payload_symbols = 0;
    int k = snprintf(dst, cap, "RAW: %s %012.4f %010d N:%05.2f%+06.2f I:%011" PRIu64 " %3d%% %.5f %3d ",
                     file_info, ts_ms, fhz, magnitude, noise, id, confidence, level,
                     n_payload_symbols);
    if (k < 0) return k;
    size_t pos = (size_t)k < cap ? (size_t)k : cap - 1;
    for (int i = 0; i < n_bits && pos + 2 < cap; i++) dst[pos++] = (char)('0' + bits[i]);
    if (pos + 1 < cap) dst[pos++] = '\n';
    dst[pos] = 0;
    return (int)pos;
}

/* ============================================================== whole path */

int orc_run_recording(const orc_cf32 *iq, size_t n, double center_frequency, int sample_rate,
                      float threshold_db, size_t feed_block, uint64_t start_time_ns,
                      int use_gardner, orc_run *out) {
    memset(out, 0, sizeof(*out));
    orc_det_params P;
    orc_det_params_init(&P, center_frequency, sample_rate, 0, 0, threshold_db);
    orc_burst *bursts = NULL;
    double t0 = cpu_now();
    size_t nb = orc_detect(&P, iq, n, feed_block, &bursts, NULL, NULL);
    double t1 = cpu_now();
    out->t_detect_s = t1 - t0;
    out->n_bursts = nb;
    out->results = malloc(sizeof(orc_result) * (nb ? nb : 1));
    size_t bits_cap = nb * 900 + 16;
    out->bits = malloc(bits_cap);
    orc_downmix *dm = orc_downmix_create();
    orc_cf32 *buf = malloc(sizeof(orc_cf32) * DM_WORK);
    orc_cf32 frame[4480];
    uint8_t bits[900];
    float llr[900];
    for (size_t i = 0; i < nb; i++) {
        double a = cpu_now();
        size_t len = orc_burst_extract(&P, iq, n, &bursts[i], buf);
        orc_burst_hdr h = {bursts[i].id, bursts[i].start, bursts[i].center_bin, P.fft_size,
                           sample_rate, bursts[i].magnitude, bursts[i].noise, center_frequency,
                           start_time_ns};
        if (bursts[i].start < bursts[i].ring_start) h.start = bursts[i].start; /* info.start is unclamped */
        orc_frame_info fi;
        int ok = orc_downmix_process(dm, &h, buf, len, &fi, frame, NULL, NULL, NULL);
        double b = cpu_now();
        out->t_downmix_s += b - a;
        if (!ok) continue;
        out->n_frames++;
        orc_demod_info di;
        int acc = orc_demod(frame, fi.num_samples, fi.samples_per_symbol, fi.center_frequency,
                            fi.direction, use_gardner, &di, bits, llr, NULL);
        out->t_demod_s += cpu_now() - b;
        if (!acc) continue;
        orc_result *r = &out->results[out->n_results++];
        r->id = fi.id; r->timestamp = fi.timestamp; r->center_frequency = di.center_frequency;
        r->direction = di.direction; r->magnitude = fi.magnitude; r->noise = fi.noise;
        r->confidence = di.confidence; r->level = di.level; r->n_symbols = di.n_symbols;
        r->n_payload_symbols = di.n_payload_symbols; r->n_bits = di.n_bits;
        r->bits_offset = (uint32_t)out->bits_len;
        memcpy(out->bits + out->bits_len, bits, (size_t)di.n_bits);
        out->bits_len += (size_t)di.n_bits;
    }
    free(buf);
    orc_downmix_destroy(dm);
    free(bursts);
    return 0;
}

void orc_run_free(orc_run *r) {
    free(r->results);
    free(r->bits);
    memset(r, 0, sizeof(*r));
}
