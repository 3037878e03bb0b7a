/*
 * iridium_b200.h -- C ABI of libiridium_b200.so: the B200-native Iridium
 * burst-detect -> downmix -> DQPSK-demod path.
 *
 * Plain C, plain pointers and sizes, no torch / CUDA types in any signature.
 * Every entry point names the reference interface it replaces (file:line relative
 * to alphafox02/iridium-sniffer @ 99453295).  INTEGRATION.md shows the bindings a
 * maintainer of the reference would add.
 *
 * Three layers are exported:
 *   1. ir_pipeline_*      batched whole-path API (what bench.py and the tests drive);
 *                         replaces the detector/downmix/demod thread trio of
 *                         main.c:667-685 for file or block input.
 *   2. ir_ref_*.h-shaped  the reference's own per-item functions with identical
 *                         signatures and struct layouts (include/ir_ref_api.h).
 *   3. gpu_burst_fft_*    the reference's existing accelerator plug-in ABI
 *                         (opencl/burst_fft.h:35-47), see include/burst_fft.h.
 * Around layer 1: ir_plan_blocks / ir_merge_blocks / ir_multi_* put one long stream on several
 * GPUs by contiguous time blocks (host bookkeeping, no collective).
 *
 * There is no CPU fallback anywhere: if no CUDA device / kernel image is usable the
 * create calls return NULL and ir_last_error() says why.
 */
#ifndef IRIDIUM_B200_H
#define IRIDIUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IR_ABI_VERSION 1

/* Sample formats of the input stream (options.c:326-337, main.c:236-262). */
enum { IR_FMT_CF32 = 0, IR_FMT_CI16 = 1, IR_FMT_CI8 = 2 };

/* Direction, same values as ir_direction_t (burst_downmix.h:32-36). */
enum { IR_DIR_UNDEF = 0, IR_DIR_DOWNLINK = 1, IR_DIR_UPLINK = 2 };

typedef struct ir_pipeline ir_pipeline_t;

/* Configuration: the union of burst_config_t (burst_detect.h:51-63), downmix_config_t
 * (burst_downmix.h:57-61) and the globals qpsk_demod reads (qpsk_demod.c:34-35).
 * Zero means "reference default" exactly as in burst_detector_create (burst_detect.c:180-226). */
typedef struct {
    uint32_t abi_version;       /* IR_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    double center_frequency;    /* Hz (main.c:645) */
    int32_t sample_rate;        /* Hz (main.c:646) */
    int32_t fft_size;           /* 0 = auto (~1 ms, power of two) */
    int32_t burst_width_hz;     /* 0 = 40000 */
    float threshold_db;         /* 0 = 16.0 */
    int32_t use_gardner;        /* main.c:143 default 1 */
    int32_t feed_block;         /* samples per emulated burst_detector_feed call; 0 = 32768
                                   (main.c:225).  Decides the emit cadence and therefore the
                                   stale-tail samples of SURVEY.md D10. */
    uint64_t start_time_ns;     /* timestamp of sample 0; 0 = CLOCK_REALTIME at first feed
                                   (burst_detect.c:755-759) */
    uint64_t max_samples;       /* capacity of the resident IQ buffer; 0 = size of first feed */
    int32_t h2d_chunk;          /* samples per pinned->device copy in ir_pipeline_run_host;
                                   0 = 32 Mi */
    int32_t reserved[7];
} ir_config_t;

/* One demodulated frame == demod_frame_t (qpsk_demod.h:24-38) without the pointers. */
typedef struct {
    uint64_t id;
    uint64_t timestamp;         /* ns */
    double center_frequency;    /* Hz */
    int32_t direction;
    float magnitude;            /* dB */
    float noise;                /* dBFS/Hz */
    int32_t confidence;         /* 0..100 */
    float level;
    int32_t n_symbols;
    int32_t n_payload_symbols;
    int32_t n_bits;
    uint32_t bits_offset;       /* into the bits / llr arrays of the result set */
} ir_frame_t;

/* Detected burst == burst_info_t + the fields of burst_data_t (burst_detect.h:29-48). */
typedef struct {
    uint64_t id;
    uint64_t start;
    uint64_t stop;
    uint64_t last_active;
    int32_t center_bin;
    float magnitude;
    float noise;
    uint64_t num_samples;
    uint64_t emit_count;        /* detector sample_count when the burst was emitted */
    int32_t downmix_status;     /* 0 = frame produced, else stage that dropped it */
    int32_t demod_ok;
    float center_offset;        /* fine CFO, cycles/sample at 250 kHz */
    int32_t dm_start;           /* find_burst_start result */
    int32_t uw_start;           /* sample index of the unique word in the frame */
    int32_t frame_len;          /* extracted samples */
    float uw_start_frac;        /* sub-sample correction (downmix_frame_t.uw_start) */
    int32_t dm_direction;
    int32_t dec_len;
} ir_burst_t;

typedef struct {
    size_t n_bursts;
    const ir_burst_t *bursts;
    size_t n_frames;
    const ir_frame_t *frames;
    const uint8_t *bits;        /* one byte per bit (0/1), like demod_frame_t.bits */
    const float *llr;
    size_t n_bits_total;
    /* device time of the last run, from CUDA events on the launching streams (ms) */
    float ms_total;             /* first kernel start .. last result ready */
    float ms_detect_fft, ms_detect_scan, ms_downmix_fir, ms_downmix_chain, ms_demod;
    uint64_t kernel_launches;   /* launches of this library's kernels in the last run */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t alg_bytes;         /* algorithmic bytes of the run (SURVEY.md 8d formula) */
} ir_results_t;

const char *ir_last_error(void);
int ir_device_count(void);

/* Replaces burst_detector_create + burst_downmix_create x4 (main.c:644-664). */
ir_pipeline_t *ir_pipeline_create(const ir_config_t *cfg);
void ir_pipeline_destroy(ir_pipeline_t *p);

/* Forget all stream state (fresh detector, empty buffer). */
int ir_pipeline_reset(ir_pipeline_t *p);

/* Whole path over a block of IQ in HOST memory (pinned or pageable): chunked H2D copies
 * overlapped with detection, then downmix + demod of every emitted burst, results copied
 * back.  The block is treated like a file handed to the reference (fresh detector).
 * Replaces spewer_thread -> burst_detector_thread -> burst_downmix_thread ->
 * frame_consumer_thread (main.c:223-385) up to, not including, frame_output_print.
 * Returns 0 on success. */
int ir_pipeline_run_host(ir_pipeline_t *p, const void *iq, size_t n_samples, int fmt);

/* Same, IQ already resident in DEVICE memory (device pointer of this pipeline's device). */
int ir_pipeline_run_device(ir_pipeline_t *p, const void *iq_dev, size_t n_samples, int fmt);

/* Results of the last run; pointers stay valid until the next run/reset/destroy. */
int ir_pipeline_results(ir_pipeline_t *p, ir_results_t *out);

/* Counters of the detector's state machine over the last run: out[0] launches the streaming
 * state machine completed, [1] launches it handed to the general cluster kernel (squelch,
 * too-long burst, baseline outside the guard band ...), [2] baseline commands, [3] event frames,
 * [4] bitmap words resolved with the exact divide, [5] waits for the baseline workers, [6] frame
 * of the last hand-over.  Returns 1 if the streaming state machine is in use, 0 if IR_SCAN
 * selected another variant, -1 on error. */
int ir_pipeline_scan_stats(ir_pipeline_t *p, uint64_t *out, int n);

/* Debug / parity taps on the last run (host copies; caller provides the buffers). */
int ir_pipeline_copy_mag(ir_pipeline_t *p, size_t frame0, size_t n_frames, float *dst);
int ir_pipeline_copy_frame_samples(ir_pipeline_t *p, size_t burst_index, float *dst_cf32,
                                   size_t cap_samples);
int ir_pipeline_copy_decimated(ir_pipeline_t *p, size_t burst_index, float *dst_cf32,
                               size_t cap_samples);
int ir_pipeline_copy_burst_samples(ir_pipeline_t *p, size_t burst_index, float *dst_cf32,
                                   size_t cap_samples);

/* frame_output_print's line (frame_output.c:160-199) into dst; t0 per ensure_initialized
 * (frame_output.c:144-158).  Returns the length, or -1. */
int ir_format_raw(char *dst, size_t cap, const char *file_info, uint64_t t0,
                  const ir_frame_t *frame, const uint8_t *bits);

/* printf's "%[+][0]<width>.<decimals>f" of x into dst (>= 400 bytes), computed exactly from the double's bits (integer
 * part, 52-bit fraction * 10^decimals as a 128-bit integer, round to nearest, ties to even) -- the conversion the
 * RAW: line formatter uses instead of printf (frame_output.c:176-190 prints four of them per line).  Returns the
 * length.  Exposed for the tests that hold it against printf. */
int ir_format_fixed(char *dst, double x, int decimals, int width, int zero_pad, int plus);

/* All RAW: lines of the last run in one call, in frame order (the batched sink of SURVEY.md 8f
 * rank 2).  t0 = 0 selects frame_output.c:144-158's rule (first frame's timestamp floored to
 * 1 s).  Returns the number of bytes written (no trailing NUL counted); with dst == NULL the
 * size a buffer must have; -1 on error. */
long ir_pipeline_format_raw_all(ir_pipeline_t *p, const char *file_info, uint64_t t0, char *dst,
                                size_t cap);

/* ---- next row of the path: what frame_consumer_thread does with each frame ---- */

/* frame_type_t (frame_decode.h:20-24). */
enum { IR_FRAME_UNKNOWN = 0, IR_FRAME_IRA = 1, IR_FRAME_IBC = 2 };

/* Outcome of frame_decode() (frame_decode.c:414-598) AND ida_decode() (ida_decode.c:543-662) on one
 * demodulated frame -- main.c:320-350 calls both on every frame.  The first half mirrors
 * decoded_frame_t's ira_data_t / ibc_data_t (frame_decode.h:26-58; timestamp and frequency are the
 * ir_frame_t's), the second ida_burst_t + lcw_t (ida_decode.h:19-56) without the fields copied from the
 * frame and without lcw_header (text, ida_decode.c:398-541: stays with the host). */
typedef struct ir_frame_class {
    int32_t frame_type;         /* IR_FRAME_*; != 0 <=> frame_decode() returned 1 */
    int32_t sat_id, beam_id;    /* IRA and IBC */
    double lat, lon;            /* IRA */
    int32_t alt;
    int32_t pos_xyz[3];
    int32_t n_pages;
    uint32_t tmsi[12];
    int32_t msc_id[12];
    int32_t timeslot, sv_blocking, bc_type;   /* IBC */
    uint32_t iri_time;
    int32_t ida_ok;             /* != 0 <=> ida_decode() returned 1 */
    int32_t lcw_ft, lcw_code, ec_lcw;
    uint32_t lcw3_val;
    int32_t da_ctr, da_len, cont, payload_len, crc_ok, fixederrs, bch_len;
    uint16_t stored_crc, computed_crc;
    uint8_t payload[32];
    uint8_t bch_stream[256];
} ir_frame_class_t;

/* Classify n_frames demodulated frames on the GPU (one launch): frames[i].n_bits / .direction /
 * .bits_offset select the frame's bits and LLRs in the two arrays (host memory, n_bits_total
 * elements each; llr may be NULL = hard decisions only, like a demod_frame_t without llr).
 * Replaces the per-frame frame_decode() + ida_decode() calls of main.c:320-350.  Returns 0, or -1
 * (ir_last_error()); there is no CPU path. */
int ir_classify_frames(int device, const ir_frame_t *frames, size_t n_frames, const uint8_t *bits,
                       const float *llr, size_t n_bits_total, ir_frame_class_t *out);

/* The same over the frames of the pipeline's last run, reading bits and LLRs where the demod kernel
 * left them in device memory (no copy of the bit arrays back up).  out[i] belongs to
 * ir_results_t.frames[i].  Returns the number of frames classified, or -1. */
long ir_pipeline_classify(ir_pipeline_t *p, ir_frame_class_t *out, size_t cap);

/* Device time (ms, CUDA events on the launching stream) of the classification kernel in the last
 * ir_pipeline_classify call; -1 without a pipeline. */
float ir_pipeline_last_classify_ms(ir_pipeline_t *p);

/* The reference's --parsed sink.  ir_format_lcw: the "LCW(...)" header ida_decode() leaves in
 * ida_burst_t.lcw_header (ida_decode.c:398-541; 110 columns + one space).  ir_format_ida: the whole
 * "IDA: ..." line of frame_output_print_ida() (frame_output.c:203-357) for a frame whose class has
 * ida_ok set; t0 as for ir_format_raw.  Return the length, or -1 (not an IDA frame / buffer too small).
 * Host text formatting, byte-identical to the reference for bch_len <= 256 (beyond that the reference
 * prints past its own bch_stream array; this prints the 256 bits it holds). */
int ir_format_lcw(char *dst, size_t cap, const ir_frame_class_t *cls);
int ir_format_ida(char *dst, size_t cap, uint64_t t0, const ir_frame_t *frame,
                  const ir_frame_class_t *cls);
/* the same with the header text given (what frame_output_print_ida() does with ida_burst_t.lcw_header) */
int ir_format_ida_hdr(char *dst, size_t cap, uint64_t t0, const ir_frame_t *frame,
                      const ir_frame_class_t *cls, const char *lcw_header);

/* All output lines of the last run the way `--parsed` prints them (main.c:328-331): the IDA line for a
 * frame ida_decode() accepted, its RAW line otherwise.  cls = what ir_pipeline_classify returned for this
 * run (n_cls entries), or NULL to classify here.  dst == NULL returns the size needed.  Returns bytes
 * written, or -1. */
long ir_pipeline_format_parsed_all(ir_pipeline_t *p, const char *file_info, uint64_t t0,
                                   const ir_frame_class_t *cls, size_t n_cls, char *dst, size_t cap);

/* ir_frame_t + ir_frame_class_t -> the reference's own structs, exactly as frame_decode() / ida_decode() leave
 * them: decoded_frame_t (frame_decode.h:49-57) and ida_burst_t (ida_decode.h:29-56, lcw_header included); the
 * pointers are void here so that this header needs no reference types (include/ir_ref_api.h declares them).
 * ir_fill_ida_burst returns 0 and a zeroed struct when ida_ok is not set, like ida_decode(). */
void ir_fill_decoded_frame(const ir_frame_t *frame, const ir_frame_class_t *cls, void *decoded_frame_out);
int ir_fill_ida_burst(const ir_frame_t *frame, const ir_frame_class_t *cls, void *ida_burst_out);

/* How ir_pipeline_run_* cuts a block of n samples into pieces (end offsets into `ends`, returns their
 * number or -1): full chunks of `chunk` samples (rounded down to whole detector frames), then the last
 * chunk in halves down to 1 Mi samples, so that little work is left after the last copy.  Pure host
 * arithmetic, exported for tests and for callers that want to align their reads with it. */
long ir_plan_chunks(size_t n_samples, size_t chunk, size_t fft_size, size_t *ends, size_t cap);

/* ---- one long stream over several pipelines / GPUs by contiguous time blocks ----
 * Each block is processed exactly like a separate file handed to the reference: a fresh detector whose first 512
 * frames build the noise baseline and detect nothing (burst_detect.c:426-428).  So that nothing is lost, block k is
 * fed from `halo` samples before the range it owns -- 512*N for the baseline, plus the longest burst, its post_len
 * and 2*pre_len (burst_detect.c:181-213), so that a burst already on the air when detection goes live is over before
 * the owned range begins -- and `tail` samples past its end, so that a burst starting in the owned range is finished,
 * declared gone and emitted (emission happens at the end of a feed call, burst_detect.c:839-841; bursts still active
 * at end of input are never emitted: burst_detector_destroy just frees).  Halo, tail and block length are multiples of
 * lcm(fft_size, feed_block), so detector frames and feed calls fall on the same samples as in the unsharded run.
 * No exchange between blocks, no collective: the host merges the frames by time stamp.
 * What differs from the unsharded run, by construction: burst ids (per block, see ir_merge_blocks), and the noise
 * baseline a block starts from (its own first 512 frames instead of the last 512 quiet frames), which moves the
 * magnitude / noise fields by hundredths of a dB and can flip a marginal detection. */
typedef struct {
    uint64_t feed_first, feed_end;   /* samples [feed_first, feed_end) of the stream go to the block's pipeline */
    uint64_t own_first, own_end;     /* the block keeps frames whose time stamp falls on samples [own_first, own_end) */
} ir_block_t;

#define IR_BLOCK_ID_STRIDE 1000000000ULL   /* merged id = block * stride + the block's own id (`I:%011`) */

/* Samples a block is fed before / after the range it owns (0 on a bad configuration). */
size_t ir_block_halo(const ir_config_t *cfg);
size_t ir_block_tail(const ir_config_t *cfg);

/* Cuts [0, n_samples) into at most n_blocks blocks of equal owned length (the last takes the remainder; fewer
 * blocks come back when n_samples is too short for the owned length to exceed the halo).  Block 0 starts at
 * sample 0 like the unsharded run.  Returns the number of blocks written, -1 on error. */
long ir_plan_blocks(const ir_config_t *cfg, size_t n_samples, int n_blocks, ir_block_t *blocks, size_t cap);

/* Absolute index of sample 0 of the following runs of this pipeline (= ir_block_t.feed_first): enters the frame time
 * stamps only -- (start + origin) / fs like burst_downmix.c:659-660 on the whole stream -- with cfg.start_time_ns
 * staying the time of the STREAM's sample 0.  Sticky until set again; 0 at creation. */
int ir_pipeline_set_origin(ir_pipeline_t *p, uint64_t sample_origin);
/* Replaces cfg.start_time_ns for the following runs (0 = CLOCK_REALTIME at each run again). */
int ir_pipeline_set_start_time(ir_pipeline_t *p, uint64_t start_time_ns);

/* Merge of the blocks' frame lists: keeps from block k the frames whose time stamp lies in its owned range (plus
 * 1 ms beyond its end, where a frame also reported by block k+1 -- within 1 ms and 200 Hz -- is kept once, the
 * earlier block's), orders them by time stamp (stable: ties keep block, then frame order) and gives them unique ids.
 * frames[k] / n_frames[k] = ir_results_t.frames / .n_frames of block k (run with ir_pipeline_set_origin(feed_first)
 * and the stream's start_time_ns).  out[i] is a copy of the frame with the merged id, bits_offset still into its
 * own block's bit array; out_block[i] says which block.  Returns the number kept, -1 on error (cap too small). */
long ir_merge_blocks(const ir_config_t *cfg, uint64_t start_time_ns, const ir_block_t *blocks, int n_blocks,
                     const ir_frame_t *const *frames, const size_t *n_frames, ir_frame_t *out,
                     uint32_t *out_block, size_t cap);

/* The above in one call, for a single process that owns several GPUs (the reference is one process, main.c): one
 * pipeline and one host thread per device, the stream cut into n_blocks time blocks (0 = one per device) dealt
 * round-robin to the devices, each device working through its blocks in order, the frame lists merged on the calling
 * thread.  cfg.device is ignored (devices[] decides), cfg.start_time_ns = 0 is resolved once (CLOCK_REALTIME, like
 * burst_detect.c:755-759) and shared by all blocks.  Results stay valid until the next run / destroy. */
typedef struct ir_multi ir_multi_t;
ir_multi_t *ir_multi_create(const ir_config_t *cfg, const int *devices, int n_devices);
void ir_multi_destroy(ir_multi_t *m);
int ir_multi_run_host(ir_multi_t *m, const void *iq, size_t n_samples, int fmt, int n_blocks);
/* n_streams INDEPENDENT streams (BASELINE config 5: e.g. one recording or receiver per GPU): stream s goes to device
 * s % n_devices, each device working through its streams in order; no halo, no merge.  The result lists the streams
 * one after the other, each in its pipeline's own frame order; `block` = the stream's index, ids = stream *
 * IR_BLOCK_ID_STRIDE + id.  All streams share cfg (rate, centre frequency, start time). */
int ir_multi_run_streams_host(ir_multi_t *m, const void *const *iq, const size_t *n_samples, int n_streams, int fmt);
/* merged frames of the last run; block[i] = the time block frame i came from, bits[block] + frames[i].bits_offset its bits */
typedef struct {
    size_t n_frames;
    const ir_frame_t *frames;
    const uint32_t *block;
    const uint32_t *index;           /* frame i is frame index[i] of block[i]'s own list */
    const ir_frame_class_t *const *classes;   /* per block, parallel to its frame list; NULL unless ir_multi_set_classify */
    size_t n_blocks;
    const ir_block_t *blocks;
    const uint8_t *const *bits;      /* per block */
    const float *const *llr;         /* per block */
    uint64_t start_time_ns;
    uint64_t kernel_launches;        /* all blocks */
    uint64_t samples_fed;            /* sum of the blocks' feed ranges (>= n_samples: halos and tails are read twice) */
} ir_multi_results_t;
int ir_multi_results(ir_multi_t *m, ir_multi_results_t *out);
/* on != 0: the following runs also classify every frame (ir_pipeline_classify per block, while its bits and LLRs are
 * still in that GPU's memory) -- what frame_consumer_thread does with `--parsed` (main.c:320-350) */
int ir_multi_set_classify(ir_multi_t *m, int on);
/* the `--parsed` text of the merged run: conventions of ir_pipeline_format_parsed_all */
long ir_multi_format_parsed_all(ir_multi_t *m, const char *file_info, uint64_t t0, char *dst, size_t cap);
/* every RAW: line of the merged run, in time order (frame_output.c:160-199); conventions of ir_pipeline_format_raw_all */
long ir_multi_format_raw_all(ir_multi_t *m, const char *file_info, uint64_t t0, char *dst, size_t cap);

/* Pinned host allocations for callers that want full-rate H2D. */
void *ir_host_alloc(size_t bytes);
void ir_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
