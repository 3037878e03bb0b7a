/*
 * ir_ref_api.h -- the reference's own per-item C interface, served by libiridium_b200.so.
 *
 * Same function names, argument meaning, ownership and error behaviour as the reference, and
 * the same struct layouts, so that main.c (and frame_output.c / frame_decode.c / ida_decode.c,
 * which consume demod_frame_t) link against this library in place of burst_detect.c,
 * burst_downmix.c and qpsk_demod.c:
 *
 *   burst_detect.h:29-97     burst_info_t, burst_data_t, burst_config_t, burst_detector_*
 *   burst_downmix.h:32-72    ir_direction_t, downmix_frame_t, downmix_config_t, burst_downmix_*
 *   qpsk_demod.h:24-42       demod_frame_t, qpsk_demod
 *
 * Differences, all deliberate:
 *   - every call runs on the GPU (H2D, kernels, D2H inside the call); there is no CPU path.
 *     A missing device makes the create functions return NULL (the reference never returns
 *     NULL; its callers do not check -- with this library they crash early instead of silently
 *     computing on the CPU).
 *   - burst_config_t.use_gpu is ignored (always GPU).
 *   - burst_detector_thread / burst_downmix_thread (burst_detect.h:97, burst_downmix.h:76) are
 *     exported too.  They shuttle items between main.c's queues (samples_queue, burst_queue,
 *     frame_queue; blocking_queue.h) and these calls, so the library refers to those queues, to
 *     blocking_queue_take/put/add and to stat_n_detected / stat_n_dropped as WEAK symbols: linked
 *     into the reference's program they bind to main.c's definitions, stand-alone they are null and
 *     the thread functions return at once.  (qpsk_demod_thread is declared by the reference but
 *     never defined, qpsk_demod.h:45.)
 *   - qpsk_demod reads `use_gardner` (main.c:143) exactly like the reference; the library
 *     carries a weak definition (=1) so it also loads stand-alone.
 * C99 only (float complex); C++ callers use include/iridium_b200.h instead.
 */
#ifndef IR_REF_API_H
#define IR_REF_API_H

#include <complex.h>
#include <stddef.h>
#include <stdint.h>

struct _burst_detector;
typedef struct _burst_detector burst_detector_t;

typedef struct {                 /* burst_detect.h:29-37 */
    uint64_t id;
    uint64_t start;
    uint64_t stop;
    uint64_t last_active;
    int center_bin;
    float magnitude;
    float noise;
} burst_info_t;

typedef struct {                 /* burst_detect.h:40-48 */
    burst_info_t info;
    double center_frequency;
    int sample_rate;
    int fft_size;
    uint64_t start_time_ns;
    size_t num_samples;
    float complex *samples;      /* malloc'd; the receiver frees samples and the struct */
} burst_data_t;

typedef struct {                 /* burst_detect.h:51-63 */
    double center_frequency;
    int sample_rate;
    int fft_size;
    int burst_pre_len;
    int burst_post_len;
    int burst_width;
    int max_bursts;
    int max_burst_len;
    float threshold;
    int history_size;
    int use_gpu;
} burst_config_t;

typedef void (*burst_callback_t)(burst_data_t *burst, void *user);

burst_detector_t *burst_detector_create(burst_config_t *config);                 /* burst_detect.h:67 */
void burst_detector_feed(burst_detector_t *det, const int8_t *iq, size_t num_samples,
                         burst_callback_t cb, void *user);                       /* :74 */
void burst_detector_feed_cf32(burst_detector_t *det, const float *iq, size_t num_samples,
                              burst_callback_t cb, void *user);                  /* :78 */
int burst_detector_active_count(burst_detector_t *det);                          /* :82 */
uint64_t burst_detector_total_count(burst_detector_t *det);                      /* :85 */
float burst_detector_noise_floor(burst_detector_t *det);                         /* :88 */
float burst_detector_peak_signal(burst_detector_t *det);                         /* :91 */
void burst_detector_destroy(burst_detector_t *det);                              /* :94 */
void *burst_detector_thread(void *arg);      /* :97; arg = burst_detector_t*, destroyed on exit */

typedef enum { DIR_UNDEF = 0, DIR_DOWNLINK = 1, DIR_UPLINK = 2 } ir_direction_t; /* burst_downmix.h:32-36 */

typedef struct {                 /* burst_downmix.h:39-51 */
    uint64_t id;
    uint64_t timestamp;
    double center_frequency;
    float sample_rate;
    float samples_per_symbol;
    ir_direction_t direction;
    float magnitude;
    float noise;
    float uw_start;
    size_t num_samples;
    float complex *samples;      /* malloc'd */
} downmix_frame_t;

typedef struct _burst_downmix burst_downmix_t;

typedef struct {                 /* burst_downmix.h:57-61 */
    int output_sample_rate;
    int search_depth;
    int handle_multiple_frames;
} downmix_config_t;

burst_downmix_t *burst_downmix_create(downmix_config_t *config);                 /* burst_downmix.h:64 */
int burst_downmix_process(burst_downmix_t *dm, burst_data_t *burst,
                          downmix_frame_t **frames_out);                         /* :69 */
void burst_downmix_destroy(burst_downmix_t *dm);                                 /* :73 */
void *burst_downmix_thread(void *arg);       /* :76; arg = burst_downmix_t*, destroyed on exit */

typedef struct {                 /* qpsk_demod.h:24-38 */
    uint64_t id;
    uint64_t timestamp;
    double center_frequency;
    ir_direction_t direction;
    float magnitude;
    float noise;
    int confidence;
    float level;
    int n_symbols;
    int n_payload_symbols;
    uint8_t *bits;               /* malloc'd, one byte per bit */
    float *llr;                  /* malloc'd */
    int n_bits;
} demod_frame_t;

int qpsk_demod(downmix_frame_t *in, demod_frame_t **out);                        /* qpsk_demod.h:42 */

#endif
