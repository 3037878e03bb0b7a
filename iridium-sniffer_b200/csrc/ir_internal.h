// ir_internal.h -- structures and launcher prototypes shared between the kernels
// (k_*.cu), the host tables (host_tables.cpp) and the pipeline (pipeline.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>


#define IR_MAX_ACTIVE 1024      // device-side cap on simultaneously tracked bursts
#define IR_SCAN_THREADS 1024
#define IR_STREAM_CL 8          // CTAs of the streaming state machine's cluster
#define IR_STREAM_MAX_FRAMES 4096   // frames per launch of the streaming state machine
#define IR_SEG_MAX_FRAMES 16384     // frames per chunk of the segmented state machine (k_detect_seg.cu)
#define IR_SEG_LEN 64              // frames per segment (one warp each); multiple of 32
#define IR_SEG_GONE 512             // gone records a segment can hold (more: the chunk falls back)
#define IR_SEG_LIST 256             // bursts alive at a segment cut / inside a segment walked by the generic walker
                                    // (its list lives in the walker's shared memory; max_bursts squelches at 200-240)
#define IR_SEG_OVF (IR_SEG_LIST - 32)   // ... of which 32 live in SegState, the rest in the overflow lists
#define IR_SEG_PCAP 2048            // candidate peaks of one frame (generic walker)
#define IR_SEG_ROUNDS 12            // rounds enqueued per chunk (a chunk without a fixed point by then falls back)
// guard band of the bitmaps: valid while every baseline stays inside [LO, HI] x its reference value.
// LO also sets how often pure noise lands in the uncertain band (e^(-23.1*LO) per bin and frame at 16 dB).
#define IR_GUARD_LO 0.65f
#define IR_GUARD_HI 1.5f
#define IR_ROT_G 16             // samples between NCO phase checkpoints
#define IR_FIR_TILE_OF(dec) ((dec) == 48 ? 240 : 256)   // decimated outputs per FIR tile (k_downmix.cu)
#define IR_FIR_R 8              // outputs per lane
#define IR_INPUT_NTAPS 801      // burst_downmix.c:252-259 (always designed for 10 MHz)
#define IR_DM_WORK (2 * 1024 * 1024)   // burst_downmix.c:366
#define IR_OUT_RATE 250000
#define IR_MAX_FRAME 4440       // 444 symbols * 10 (iridium.h:27)
#define IR_MAX_SYMS 480
#define IR_CFO_N 256
#define IR_CFO_TOTAL 4096
#define IR_CORR_N 2048
#define IR_SYNC_SEARCH 840
#define IR_SYNC_LEN 271

struct ir_frame_class;     // include/iridium_b200.h

namespace ir {

// ------------------------------------------------------------------ detector
struct DetConfig {                 // derived as burst_detector_create does (burst_detect.c:174-226)
    int L, N;
    int hist_size;
    int half_bw;                   // burst_width / 2 (bins)
    int burst_width;
    int max_bursts;
    int pre_len, post_len, max_burst_len;
    float thr;                     // linear threshold
    int sample_rate;
    uint64_t ringbuf_size;
};

struct ActBurst {
    uint64_t id, start, last_active;
    int center_bin;
    float peak_rel;
    float base_at_create;
    int pad;
};

struct GoneBurst {
    uint64_t id, start, stop, last_active;
    int center_bin;
    float peak_rel;
    float base_at_create;
    int pad;
};

struct DetState {
    int hist_idx, primed, n_act, squelch_count;
    uint64_t next_id;
    uint64_t index;                // absolute sample index of the next frame
    uint32_t n_gone;               // entries written to the gone list so far
    uint32_t n_squelch;
    uint32_t overflow;             // IR_MAX_ACTIVE or gone list exceeded
    uint32_t pad;
    unsigned long long dbg[24];    // cycle counters of the cluster state machine (IR_SCAN_DEBUG)
    ActBurst act[IR_MAX_ACTIVE];
};

// what a bailed streaming launch is undone from: the history values its baseline updates overwrote
// (one row per quiet frame, in order), the baseline at its start, and how many rows (in the control block)
struct ScanSnapshot {
    const float *undo = nullptr;
    const float *base = nullptr;
    const struct StreamCtl *ctl = nullptr;
};

// control block of the streaming state machine (k_detect_stream.cu), device memory
struct StreamCtl {
    unsigned long long cmd[IR_STREAM_MAX_FRAMES + 8];   // leader -> workers: ranges of quiet frames
    unsigned int done[IR_STREAM_CL];                    // workers -> leader: commands finished per CTA
    unsigned int guard_bad;                             // a baseline left the band the bitmaps were made for
    int bailed;                                         // last launch gave up: restore + fallback must run
    int undo_frames;                                    // rows of the undo log the last launch wrote
    int reason;                                         // why (1 primed priming launch, 2 guard band, 3 too-long burst,
                                                        //  4 peak list, 5/7 burst table, 6 squelch, 8 primed in mid-launch)
    unsigned long long stats[24];                       // 0 launches kept, 1 bailed, 2 commands, 3 event frames,
                                                        // 4 words resolved exactly, 5 waits for the workers, 6 frame of last bail,
                                                        // 7 ns inside the kernel, 8-11 leader cycles: ring wait, bitmap pass,
                                                        // wait for workers, event body
};

// ---- segmented state machine (k_detect_seg.cu)
struct SegBurst {                  // an active burst at a segment cut; times in frames of the chunk
    unsigned long long id;         // real id, or (1<<63 | segment<<32 | ordinal) for a burst created in this chunk
    unsigned long long start, last0;
    int cb;
    float rel, base;
    int dl, lah, tl;               // deletion deadline, frame of the latest hit (or NONE), last frame it cannot be too long
};
struct SegState { int n_act; int pad[3]; SegBurst b[32]; };
struct SegCtl {
    int finished, converged, changed, round, cur, hard_bail, F, S, nq, n_slots, sq0, hist_idx0, skip_base, q_keep;
    int qfc[2];                    // per round parity: first frame whose quiet flag changed in that round (INT_MAX: none)
    int bailed;                    // the chunk was not kept: the fallback must run
    int reason;                    // why (1 not primed, 2 guard band, 3 too-long burst, 4 peak list, 5/7 burst table, 6 squelch,
                                   //  9 missing snapshot, 10 gone pool, 11 snapshot slots, 12 no fixed point)
    unsigned int guard_bad, n_gone0, reclass, reclass_prev;
    unsigned long long index0, next_id0;
    unsigned long long stats[8];   // 0 chunks kept, 1 chunks bailed, 2 rounds, 3 event frames (all rounds), 4 snapshots, 5 quiet frames,
                                   // 6 segments walked by the generic walker (all rounds), 7 bitmap rebuilds (a baseline left its band)
};
struct SegBuffers {                // device memory of the segmented scan, owned by the pipeline
    SegCtl *ctl = nullptr;
    SegState *stA = nullptr, *stB = nullptr;
    uint32_t *qw = nullptr, *valid = nullptr;
    int *wpre = nullptr, *qlist = nullptr, *slotv = nullptr, *fslot = nullptr, *ncreate = nullptr, *ngone = nullptr;
    int *segbail = nullptr, *stch = nullptr, *cpre = nullptr, *gpre = nullptr;
    GoneBurst *glist = nullptr;
    SegBurst *ovfA = nullptr, *ovfB = nullptr;                        // bursts 32.. of the lists at the segment cuts
    unsigned long long *gkeys = nullptr;                              // [S][IR_SEG_PCAP] candidate peaks of the generic walker
    uint32_t *seggen = nullptr;                                       // [s]: segment s went to the generic walker
    float *snap = nullptr, *bfinal = nullptr, *qmag = nullptr, *glo = nullptr, *ghi = nullptr;
    int slot_cap = 0, frames_cap = 0;
};

// ------------------------------------------------------------------ downmix
struct BurstParam {                // one per emitted burst, built on the host
    int64_t start;                 // first sample of the extract (after ring clamp)
    int64_t emit_count;            // samples the detector had seen at emission
    int32_t n;                     // samples in the extract (<= IR_DM_WORK)
    int32_t dec_len;               // decimated length
    int64_t dec_off;               // offset into the per-run decimated / scratch arrays
    float2 incr_coarse;            // cexpf(-j*2*pi*rel) from the host libm (burst_downmix.c:668-669)
    const float2 *rot_table;       // phase checkpoints every IR_ROT_G samples for this bin
    int32_t tile0;                 // first FIR tile of this burst
    int32_t simplex;               // extraction limits of burst_downmix.c:764-770 decided later
    double cfreq_coarse;           // center_frequency after the coarse shift
};

struct ChainOut {                  // what the 250 kHz chain reports per burst
    int32_t status;                // 0 ok, else failing stage (2,3,7,9)
    int32_t start;                 // find_burst_start
    float center_offset;
    int32_t cfo_peak_bin;
    int32_t direction;
    int32_t corr_offset;
    int32_t uw_start;
    float uw_corr;                 // sub-sample correction
    float corr_re, corr_im;
    int32_t frame_len;
    float2 incr_fine;
};

struct DemodOut {
    int32_t ok;
    int32_t direction;
    int32_t confidence;
    float level;
    int32_t n_symbols;
    int32_t n_raw_symbols;
    float total_phase;
    int32_t pad;
};

// Filters / windows / templates exactly as the reference designs them (host, float math).
struct HostTables {
    std::vector<float> det_window;        // blackman/0.42 (burst_detect.c:247-250)
    std::vector<float> h_input;           // 801 taps (burst_downmix.c:252-259)
    std::vector<float> h_noise;           // 25 taps  (:264-277)
    std::vector<float> h_box;             // 20 taps  (:281-287)
    std::vector<float> h_rrc;             // 51 taps  (:290-296)
    std::vector<float> h_rc;              // 51 taps  (:299-305)
    std::vector<float> cfo_window;        // blackman(256) (:320-321)
    std::vector<float2> sync_dl_fft;      // 2048 (:358-363)
    std::vector<float2> sync_ul_fft;
};
void build_host_tables(HostTables &t, int det_fft_size);
void set_last_error(const std::string &s);   // what ir_last_error() returns on this thread (pipeline.cu)
void derive_det_config(DetConfig &c, int sample_rate, int fft_size, int burst_width_hz, float threshold_db);
std::vector<float2> build_twiddle_image(int L);      // layout of ir_device.cuh
void host_fft(std::vector<float2> &x, bool inverse);   // same radix-2 DIF, for the templates

// ------------------------------------------------------------------ launchers
// k_detect.cu
cudaError_t launch_detect_fft(int L, int fmt, const void *iq, int64_t first_sample, const float *window,
                              const float2 *tw, float *mag, int64_t n_frames, int sm_count,
                              cudaStream_t st);
cudaError_t launch_detect_scan(const DetConfig &c, DetState *state, float *base, float *hist,
                               const float *mag, int64_t n_frames, GoneBurst *gone,
                               uint32_t gone_cap, cudaStream_t st);
// k_detect_cluster.cu: same state machine on an 8-CTA cluster with speculative frame batches
cudaError_t launch_detect_scan_cluster(const DetConfig &c, DetState *state, float *base, float *hist,
                                       const float *mag, int64_t n_frames, GoneBurst *gone,
                                       uint32_t gone_cap, cudaStream_t st);
// the same, but a no-op unless *run_if != 0 (device flag; the fallback of the streaming scan)
cudaError_t launch_detect_scan_cluster_if(const DetConfig &c, DetState *state, float *base, float *hist,
                                          const float *mag, int64_t n_frames, GoneBurst *gone,
                                          uint32_t gone_cap, const int *run_if, const ScanSnapshot &snap,
                                          cudaStream_t st);
// k_detect_stream.cu: bitmaps against a reference baseline on all SMs + a one-warp state machine
// with baseline workers (see the file header).  Caller snapshots / restores around it.
bool stream_scan_supported(const DetConfig &c);
cudaError_t launch_detect_classify(const float *mag, const float *base, float thr, int N, int n_frames,
                                   uint32_t *xu, float *ref_out, int sm_count, cudaStream_t st,
                                   unsigned char *rowany = nullptr);
// k_detect_seg.cu: the same state machine cut into segments walked concurrently, exact by fixed point
bool seg_scan_supported(const DetConfig &c);
cudaError_t launch_detect_seg_prime(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                    int n_frames, cudaStream_t st);
#ifdef IR_SEG_TIMING
void seg_timing_dump();
#endif
size_t seg_walk_smem(const DetConfig &c);
cudaError_t launch_detect_scan_seg(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                   uint32_t *xu, unsigned char *rowany, const float *ref, int n_frames,
                                   GoneBurst *gone, uint32_t gone_cap, const SegBuffers &b, int *n_launches, cudaStream_t st);
cudaError_t launch_detect_scan_stream(const DetConfig &c, DetState *state, float *base, float *hist,
                                      const float *mag, const uint32_t *xu, const float *ref, int n_frames,
                                      GoneBurst *gone, uint32_t gone_cap, StreamCtl *ctl, unsigned epoch,
                                      float *undo, float *base_snap, cudaStream_t st);
// picks the cluster kernel for N >= 2048 unless IR_SCAN=single is set in the environment
cudaError_t launch_detect_scan_auto(const DetConfig &c, DetState *state, float *base, float *hist,
                                    const float *mag, int64_t n_frames, GoneBurst *gone,
                                    uint32_t gone_cap, cudaStream_t st);
// k_downmix.cu
cudaError_t upload_input_taps(const float *taps, int ntaps);
cudaError_t upload_chain_tables(const HostTables &t);
cudaError_t launch_rot_tables(const float2 *incr, float2 *const *tables, const int *lens, int n,
                              cudaStream_t st);
// tile_burst[tile] = index of the burst the tile belongs to (nullptr: the one-tile-per-CTA kernel with its search)
cudaError_t launch_fir(int fmt, int dec, const void *iq, int64_t n_total, uint64_t ring,
                       const BurstParam *bp, const int *tile_start, const int *tile_burst, int n_bursts, int n_tiles,
                       float2 *dec_out, cudaStream_t st);
cudaError_t launch_chain(const BurstParam *bp, int n_bursts, const float2 *dec, float2 *scrA,
                         float2 *scrB, const float2 *tw4096, const float2 *tw2048,
                         const float2 *sync_dl, const float2 *sync_ul, ChainOut *out,
                         float2 *frames, cudaStream_t st);
// k_demod.cu
// max_frame: an upper bound of the frame lengths of this launch (1910 unless a burst lies in the simplex band), 0 = 4440
cudaError_t launch_demod(const ChainOut *co, int n_bursts, const float2 *frames, int use_gardner,
                         DemodOut *out, uint8_t *bits, float *llr, int max_frame, cudaStream_t st);
// k_classify.cu: frame classification (frame_decode() + ida_decode() per frame), one warp per frame
struct FrameSrc {
    const uint8_t *bits;     // one byte per bit, device memory
    const float *llr;        // may be null
    int32_t n_bits;
    int32_t direction;
};
// syndrome tables of the five BCH codes, built once per process and uploaded once per device
cudaError_t classify_tables(int device, const void **d_tables);
cudaError_t launch_classify(const void *d_tables, const FrameSrc *src, int n_frames,
                            ::ir_frame_class *out, cudaStream_t st);

}  // namespace ir
