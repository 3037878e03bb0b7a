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

/* ---- next row: frame_decode.h / ida_decode.h, so that frame_decode.c and ida_decode.c
 * can leave the build as well.  frame_decode() and ida_decode() classify their one frame on the GPU
 * (k_classify_frames); the reassembly and the bit helpers are host bookkeeping, restated. ---- */
typedef enum { FRAME_UNKNOWN = 0, FRAME_IRA, FRAME_IBC } frame_type_t;           /* frame_decode.h:20-24 */

typedef struct {                 /* frame_decode.h:26-38 */
    int sat_id;
    int beam_id;
    double lat, lon;
    int alt;
    int pos_xyz[3];
    int n_pages;
    struct { uint32_t tmsi; int msc_id; } pages[12];
} ira_data_t;

typedef struct {                 /* frame_decode.h:40-47 */
    int sat_id;
    int beam_id;
    int timeslot;
    int sv_blocking;
    int bc_type;
    uint32_t iri_time;
} ibc_data_t;

typedef struct {                 /* frame_decode.h:49-57 */
    frame_type_t type;
    uint64_t timestamp;
    double frequency;
    union { ira_data_t ira; ibc_data_t ibc; };
} decoded_frame_t;

void frame_decode_init(void);                                                    /* frame_decode.h:60 (no-op here) */
int frame_decode(const demod_frame_t *frame, decoded_frame_t *out);              /* :63 */
uint32_t gf2_remainder(uint32_t poly, uint32_t val);                             /* :66-68 */
uint32_t bits_to_uint(const uint8_t *bits, int n);
void uint_to_bits(uint32_t val, uint8_t *bits, int n);
int bch_31_21_correct(uint32_t syndrome, uint32_t *locator);                     /* :72 */

typedef struct {                 /* ida_decode.h:19-26 */
    int ft;
    int lcw_ok;
    int lcw_ft;
    int lcw_code;
    uint32_t lcw3_val;
    int ec_lcw;
} lcw_t;

typedef struct {                 /* ida_decode.h:29-56 */
    uint64_t timestamp;
    double frequency;
    ir_direction_t direction;
    float magnitude;
    float noise;
    float level;
    int confidence;
    int n_symbols;
    int da_ctr;
    int da_len;
    int cont;
    uint8_t payload[32];
    int payload_len;
    int crc_ok;
    uint16_t stored_crc;
    uint16_t computed_crc;
    int fixederrs;
    uint8_t bch_stream[256];
    int bch_len;
    lcw_t lcw;
    char lcw_header[128];
} ida_burst_t;

typedef struct {                 /* ida_decode.h:59-67 */
    int active;
    ir_direction_t direction;
    double frequency;
    uint64_t last_timestamp;
    int last_ctr;
    uint8_t data[256];
    int data_len;
} ida_reassembly_t;

#define IDA_MAX_REASSEMBLY 16
typedef struct { ida_reassembly_t slots[IDA_MAX_REASSEMBLY]; } ida_context_t;   /* ida_decode.h:72-74 */

typedef void (*ida_message_cb)(const uint8_t *data, int len, uint64_t timestamp, double frequency,
                               ir_direction_t direction, float magnitude, void *user);   /* :77-80 */

void ida_decode_init(void);                                                      /* ida_decode.h:83 (no-op here) */
int ida_decode(const demod_frame_t *frame, ida_burst_t *burst);                  /* :87 */
int ida_reassemble(ida_context_t *ctx, const ida_burst_t *burst, ida_message_cb cb, void *user);   /* :91 */
void ida_reassemble_flush(ida_context_t *ctx, uint64_t now_ns);                  /* :95 */

/* simd_kernels.h:105 -- main.c:567 calls it at start-up; with this library there are no CPU kernels to select. */
void simd_init(int force_generic);

/* ---- frame_output.h:20-29: the per-line sinks (one fwrite + fflush per line, like the reference).  They read
 * main.c's diagnostic_mode / acars_enabled as weak symbols.  frame_output_zmq_* (HAVE_ZMQ builds) are not
 * provided: keep frame_output.c if ZMQ publishing is wanted. ---- */
void frame_output_init(const char *file_info);
void frame_output_print(demod_frame_t *frame);
void frame_output_print_ida(const ida_burst_t *burst);

#endif
