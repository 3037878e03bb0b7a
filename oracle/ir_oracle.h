/*
 * ir_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (iridium-sniffer_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_ref.py checks this restatement against
 * oracle/_ref (the unmodified reference sources + FFT shim) stage by stage and
 * against the reference's in-tree known answers (unique words iridium.h:30-31,
 * access codes frame_decode.c:51-56, derived constants printed by -v); the
 * resulting vectors are committed under tests/golden/.
 *
 * Arithmetic contract shared with the CUDA kernels (so GPU == oracle bit for bit
 * on every float that does not go through libm):
 *   - all FFTs: radix-2 decimation-in-frequency, float, twiddle table
 *     W[k] = ((float)cos(2*pi*k/N), (float)-sin(2*pi*k/N)) with W[N/4] forced to (0,-1),
 *     butterfly  a' = a + b;  d = a - b;  b' = (fma(d.re,w.re,-(d.im*w.im)),
 *                                            fma(d.re,w.im,  d.im*w.re));
 *     natural-order output.  (The reference uses FFTW; FFT-level parity is unpinned
 *     by the reference itself)
 *   - everything else follows the reference's AVX2 build operation for operation
 *     (file:line cited at each function in ir_oracle.c).
 */
#ifndef IR_ORACLE_H
#define IR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } orc_cf32;

/* ---------------- FFT ---------------- */
/* In-place n-point DFT (n power of two), forward (inverse=0) or unnormalised backward. */
void orc_fft(orc_cf32 *x, int n, int inverse);

/* ---------------- detector ---------------- */
typedef struct {
    double center_frequency;
    int sample_rate;
    int fft_size;
    int burst_pre_len;
    int burst_post_len;
    int burst_width_bins;
    int max_bursts;
    int max_burst_len;
    int history_size;
    float threshold_db;
    float threshold_lin;
    size_t ringbuf_size;
} orc_det_params;

/* Fill derived parameters exactly as burst_detector_create does (burst_detect.c:174-299).
 * Pass 0 for "auto" fields. */
void orc_det_params_init(orc_det_params *p, double center_frequency, int sample_rate,
                         int fft_size, int burst_width_hz, float threshold_db);

typedef struct {
    uint64_t id;
    uint64_t start;
    uint64_t stop;
    uint64_t last_active;
    int32_t center_bin;
    float magnitude;
    float noise;
    float peak_rel;        /* relative magnitude of the creating peak */
    float base_at_create;  /* baseline_sum[center_bin] when created */
    uint64_t emit_count;   /* detector sample_count when the burst was emitted */
    uint64_t ring_start;   /* ringbuf_start in force at emission */
} orc_burst;

/* Blackman/0.42 window the detector applies (burst_detect.c:247-250, window_func.c:19-24). */
void orc_det_window(float *w, int n);

/* One detector frame: window -> FFT -> fftshift -> |X|^2 (burst_detect.c:679-687). */
void orc_det_frame_mag(const orc_cf32 *frame, const float *window, int n, float *mag_out);

/* Run the detector over a whole recording fed `feed_block` samples at a time
 * (burst_detect.c:746-925 driven as main.c:223-284 does; feed_block=32768 for files).
 * Returns the number of emitted bursts; *out is malloc'd (free with orc_free).
 * mag_dump (optional, n_frames*fft_size floats) receives every frame's magnitudes.
 * n_squelch (optional) counts squelch events. */
size_t orc_detect(const orc_det_params *p, const orc_cf32 *iq, size_t n, size_t feed_block,
                  orc_burst **out, float *mag_dump, int *n_squelch);

/* Number of samples ringbuf_extract would return for this burst (burst_detect.c:401-422,703-717). */
size_t orc_burst_num_samples(const orc_det_params *p, const orc_burst *b);

/* Copy the burst's IQ exactly as the reference's ring buffer would deliver it, including
 * the stale-tail quirk: positions >= emit_count read the sample one ring
 * lap earlier, or zero when nothing was ever written there. */
size_t orc_burst_extract(const orc_det_params *p, const orc_cf32 *iq, size_t n,
                         const orc_burst *b, orc_cf32 *dst);

/* int8 -> float conversion of the feed path (simd_avx2.c:264-294) and the file reader's
 * ci16 -> int8 step (main.c:245-246). */
void orc_convert_ci8(const int8_t *iq, size_t n, orc_cf32 *dst);
void orc_convert_ci16(const int16_t *iq, size_t n, orc_cf32 *dst);

/* ---------------- downmix ---------------- */
typedef struct orc_downmix orc_downmix;
orc_downmix *orc_downmix_create(void);            /* burst_downmix.c:223-373, default config */
void orc_downmix_destroy(orc_downmix *dm);
/* tap / template access for the tests */
const float *orc_downmix_taps(const orc_downmix *dm, int which, int *ntaps); /* 0 input,1 noise,2 box,3 rrc,4 rc */
const orc_cf32 *orc_downmix_sync_fft(const orc_downmix *dm, int uplink, int *sync_len);
const float *orc_downmix_cfo_window(const orc_downmix *dm, int *n);

typedef struct {
    /* inputs copied from burst_data_t */
    uint64_t id;
    uint64_t start;
    int32_t center_bin;
    int32_t fft_size;
    int32_t sample_rate;
    float magnitude;
    float noise;
    double center_frequency;
    uint64_t start_time_ns;
} orc_burst_hdr;

typedef struct {
    int32_t ok;                 /* 1 = frame produced */
    int32_t fail_stage;         /* 0 none, 2 decimate, 3 start, 7 sync, 9 extract */
    uint64_t id;
    uint64_t timestamp;
    double center_frequency;
    float sample_rate;
    float samples_per_symbol;
    int32_t direction;          /* 1 DL, 2 UL */
    float magnitude;
    float noise;
    float uw_start;             /* sub-sample correction */
    int32_t num_samples;        /* extracted frame length */
    /* trace */
    int32_t dec_len;
    int32_t start;              /* find_burst_start result */
    float center_offset;        /* fine CFO, cycles/sample */
    int32_t cfo_peak_bin;
    int32_t corr_offset;
    int32_t uw_start_idx;
    float corr_re, corr_im;
    float incr_coarse_re, incr_coarse_im;
    float incr_fine_re, incr_fine_im;
} orc_frame_info;

/* Process one burst (burst_downmix.c:643-797).  frame_out must hold >= 4440 samples.
 * dec_out / nlpf_out / rrc_out (optional) receive the decimated burst, the noise-filtered
 * burst and the matched-filtered frame (each up to samples/dec entries). */
int orc_downmix_process(orc_downmix *dm, const orc_burst_hdr *hdr, const orc_cf32 *samples,
                        size_t num_samples, orc_frame_info *info, orc_cf32 *frame_out,
                        orc_cf32 *dec_out, orc_cf32 *nlpf_out, orc_cf32 *rrc_out);

/* ---------------- demod ---------------- */
typedef struct {
    int32_t ok;
    int32_t direction;          /* possibly updated (qpsk_demod.c:429-465) */
    int32_t confidence;
    float level;
    int32_t n_symbols;
    int32_t n_payload_symbols;
    int32_t n_bits;
    double center_frequency;
    float total_phase;
    int32_t n_raw_symbols;      /* symbols out of the timing-recovery stage */
} orc_demod_info;

/* qpsk_demod.c:393-535.  bits_out >= 2*(num_samples/10+1) bytes, llr_out same count of floats. */
int orc_demod(const orc_cf32 *frame, int num_samples, float samples_per_symbol,
              double center_frequency, int direction, int use_gardner,
              orc_demod_info *info, uint8_t *bits_out, float *llr_out,
              orc_cf32 *pll_out /* optional, >= num_samples/10+1 */);

/* ---------------- RAW line ---------------- */
/* frame_output.c:160-199; t0 handling of ensure_initialized (:144-158) is the caller's:
 * pass t0 = (first_timestamp / 1e9) * 1e9.  Returns the line length (no trailing NUL counted). */
int orc_format_raw(char *dst, size_t cap, const char *file_info, uint64_t t0,
                   uint64_t timestamp, double center_frequency, float magnitude, float noise,
                   uint64_t id, int confidence, float level, int n_payload_symbols,
                   const uint8_t *bits, int n_bits);

/* ---------------- whole path ---------------- */
typedef struct {
    uint64_t id;
    uint64_t timestamp;
    double center_frequency;
    int32_t direction;
    float magnitude;
    float noise;
    int32_t confidence;
    float level;
    int32_t n_symbols;
    int32_t n_payload_symbols;
    int32_t n_bits;
    uint32_t bits_offset;        /* into the shared bits buffer */
} orc_result;

typedef struct {
    size_t n_bursts;
    size_t n_frames;       /* downmix produced a frame */
    size_t n_results;      /* demod accepted */
    orc_result *results;   /* malloc'd */
    uint8_t *bits;         /* malloc'd, one byte per bit */
    size_t bits_len;
    double t_detect_s, t_downmix_s, t_demod_s;   /* CPU seconds per stage (single thread) */
} orc_run;

/* detect -> downmix -> demod over a cf32 recording; start_time_ns as the detector would
 * have stamped it.  Free with orc_run_free. */
int orc_run_recording(const orc_cf32 *iq, size_t n, double center_frequency, int sample_rate,
                      float threshold_db, size_t feed_block, uint64_t start_time_ns,
                      int use_gardner, orc_run *out);
void orc_run_free(orc_run *r);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
