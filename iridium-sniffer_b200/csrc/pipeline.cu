// pipeline.cu -- host side of libiridium_b200.so: owns device memory, streams and the
// burst bookkeeping between the kernels, and exports the C ABI of include/iridium_b200.h.
//
// Data layout in HBM:
//   iq      [n]            resident input block in its native format (cf32 / ci16 / ci8)
//   mag     [frames][N]    f32 |X|^2 per detector frame (written by k_detect_fft, read once by
//                          k_detect_scan)
//   hist    [512][N], base [N]   detector noise floor
//   dec / scrA / scrB [sum dec_len]   250 kHz working arrays of all bursts of the run
//   frames  [bursts][4440] extracted frames,  bits/llr [bursts][2*480]
// Bursts are (offset, length) pairs into iq -- the reference's ring-buffer copy
// (burst_detect.c:401-422) does not exist here.
#include <math.h>
#include <cmath>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/iridium_b200.h"
#include "frame_classify.cuh"
#include "ir_device.cuh"
#include "ir_internal.h"

using namespace ir;

static thread_local std::string g_err;
static void set_err(const std::string &s) { g_err = s; }
namespace ir { void set_last_error(const std::string &s) { g_err = s; } }
extern "C" const char *ir_last_error(void) { return g_err.c_str(); }

#define CK(expr)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            set_err(std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

extern "C" int ir_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

namespace {

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = n + n / 4 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { cap = 0; set_err(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct RotTable { float2 *d = nullptr; int len = 0; float2 incr; };

struct EvPair { cudaEvent_t a, b; };

// Bump allocator over one big allocation (device memory or pinned host memory).  The per-wave
// burst buffers come from here, so nothing is freed or reallocated while earlier waves are
// still in flight; a request that does not fit gets its own allocation, and the next run (device
// idle) regrows the arena to the size the previous run needed.
struct Arena {
    bool host = false;
    unsigned char *base = nullptr;
    size_t cap = 0, used = 0;
    std::vector<std::pair<void *, size_t>> extra;
    void *raw_alloc(size_t n) {
        void *q = nullptr;
        cudaError_t e = host ? cudaHostAlloc(&q, n, cudaHostAllocDefault) : cudaMalloc(&q, n);
        if (e != cudaSuccess) { set_err(std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e)); return nullptr; }
        return q;
    }
    void raw_free(void *q) { if (!q) return; if (host) cudaFreeHost(q); else cudaFree(q); }
    int begin(size_t min_cap) {
        size_t want = cap;
        for (auto &e : extra) { raw_free(e.first); want += e.second; }
        const bool grow = !extra.empty() || cap < min_cap;
        extra.clear();
        if (grow) {
            want = std::max(want + want / 4, min_cap);
            raw_free(base);
            base = (unsigned char *)raw_alloc(want);
            cap = base ? want : 0;
            if (!base) return -1;
        }
        used = 0;
        return 0;
    }
    template <class T>
    T *take(size_t count) {
        size_t n = (count * sizeof(T) + 255) & ~(size_t)255;
        if (n == 0) n = 256;
        if (used + n <= cap) { void *q = base + used; used += n; return (T *)q; }
        void *q = raw_alloc(n);
        if (q) extra.push_back({q, n});
        return (T *)q;
    }
    void release() { for (auto &e : extra) raw_free(e.first); extra.clear(); raw_free(base); base = nullptr; cap = used = 0; }
};

// One wave = the bursts the detector emitted while scanning one chunk of the input.  Its
// downmix / demod kernels run on st_burst while the detector is busy with the next chunk.
struct Wave {
    size_t b0 = 0, nb = 0;
    int n_tiles = 0;
    int64_t dec_total = 0;
    BurstParam *d_bp = nullptr, *h_bp = nullptr;
    int *d_tile_start = nullptr, *h_tile_start = nullptr, *d_tile_burst = nullptr, *h_tile_burst = nullptr;
    float2 *d_dec = nullptr, *d_scrA = nullptr, *d_scrB = nullptr, *d_frames = nullptr;
    ChainOut *d_co = nullptr, *h_co = nullptr;
    DemodOut *d_do = nullptr, *h_do = nullptr;
    unsigned char *d_bits = nullptr, *h_bits = nullptr;
    float *d_llr = nullptr, *h_llr = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e1a = nullptr, e1b = nullptr, e2 = nullptr, e3 = nullptr, e_done = nullptr;
};

struct Chunk { size_t end = 0; cudaEvent_t e_hdr = nullptr; EvPair fft, scan; };

constexpr size_t kHdrBytes = offsetof(DetState, act);

}  // namespace

struct ir_pipeline {
    ir_config_t cfg;
    DetConfig dc;
    HostTables tab;
    int dev = 0, sm_count = 148;
    int dec = 40;
    cudaStream_t st_copy = nullptr, st_fft = nullptr, st_scan = nullptr, st_burst = nullptr, st_chain = nullptr;
    // the symbol slicer is a serial recurrence per frame (long latency, few warps): its launches go
    // round-robin over a few streams of their own so that they overlap the FIR / chain of later waves
    static constexpr int kDemodStreams = 2;   // (8 streams in all: the default number of hardware queues -- a ninth
                                              //  shares one with another stream, and sharing the copy stream's queue,
                                              //  filled up front, kept every 4th wave's slicer waiting for the last copy)
    cudaStream_t st_demod[kDemodStreams] = {nullptr, nullptr};
    // constants
    DevBuf<float> d_window;
    DevBuf<float2> d_tw_det, d_tw12, d_tw11, d_sync_dl, d_sync_ul;
    // stream
    DevBuf<unsigned char> d_iq;
    DevBuf<float> d_mag, d_base, d_hist;
    DevBuf<DetState> d_state;
    // streaming state machine (k_detect_stream.cu): bitmaps of one launch, the reference baseline
    // they were made against, the snapshot a bailed launch is undone from, the control block
    DevBuf<uint32_t> d_xu[2];                // double-buffered: the bitmap pass of launch k+1 overlaps launch k
    DevBuf<float> d_ref[2], d_undo, d_base_snap;
    cudaStream_t st_cls = nullptr;           // bitmap passes
    std::vector<cudaEvent_t> ev_scan_done;   // per state-machine launch of the current run
    DevBuf<StreamCtl> d_ctl;
    unsigned scan_epoch = 1;
    bool scan_dbg = false;                   // IR_SCAN_DEBUG: events between the operations of every launch
    std::vector<cudaEvent_t> scan_dbg_ev;
    int scan_mode = 0;                       // 0 = streaming, 1 = cluster / single, 2 = segmented (default where supported); IR_SCAN
    // segmented state machine (k_detect_seg.cu)
    SegBuffers seg;
    DevBuf<unsigned char> d_rowany[2], d_seg_raw;
    DevBuf<float> d_seg_snap, d_seg_qmag;
    size_t seg_frames_cap = 0;
    uint64_t scan_stats[24] = {0};
    // burst list: pinned host memory mapped into the device; the scan kernel stores the (few,
    // 56-byte) records straight into it, the host reads them after the chunk's event
    GoneBurst *h_gone = nullptr, *d_gone = nullptr;
    uint32_t gone_cap = 0;
    unsigned char *h_hdr = nullptr;          // pinned: DetState header after every chunk
    size_t hdr_slots = 0;
    // bursts
    Arena dev_arena, pin_arena;
    std::vector<Wave> waves;
    std::vector<Chunk> chunks;
    std::unordered_map<int, RotTable> rot;
    std::vector<std::pair<unsigned char *, size_t>> rot_slabs;   // (base, used); never freed before destroy
    size_t rot_slab_cap = 0;                                      // capacity of rot_slabs.back()
    // last run (host)
    const void *last_iq = nullptr;
    size_t last_n = 0;
    int last_fmt = 0;
    std::vector<BurstParam> h_bp;
    std::vector<float2 *> frame_ptr, dec_ptr;                     // per burst, device
    std::vector<ir_burst_t> bursts;
    std::vector<ir_frame_t> frames;
    std::vector<uint8_t> bits;
    std::vector<float> llr;
    std::vector<FrameSrc> frame_src;                              // per frame: where k_demod left its bits / LLRs (device)
    DevBuf<FrameSrc> d_frame_src;
    DevBuf<ir_frame_class_t> d_class;
    float ms_classify = 0.0f;                                     // device time of the last k_classify_frames launch
    cudaEvent_t ev_cls[2] = {nullptr, nullptr};
    // RAW: lines formatted while the run is still in flight (everything after the file_info field,
    // with t0 = frame_output.c:144-158's rule), so that the batched sink is a copy
    std::string raw_rest;
    std::vector<size_t> raw_off;
    uint64_t raw_t0 = 0;
    uint64_t alg = 0, alg_sink = 0;                             // (alg_sink: added by the thread that assembles the waves)
    ir_results_t res;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    uint64_t start_time_ns = 0;
    uint64_t sample_origin = 0;                                  // ir_pipeline_set_origin: index of sample 0 in the whole stream
    int64_t n_frames_last = 0;

    cudaEvent_t ev() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
};

static int upload_f2(DevBuf<float2> &b, const std::vector<float2> &v) {
    if (b.ensure(v.size())) return -1;
    CK(cudaMemcpy(b.p, v.data(), v.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" ir_pipeline_t *ir_pipeline_create(const ir_config_t *cfg) {
    if (!cfg || cfg->abi_version != IR_ABI_VERSION) { set_err("ir_pipeline_create: bad config / ABI version"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        set_err(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
        return nullptr;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_err("device ordinal out of range"); return nullptr; }
    if (cfg->sample_rate <= 0) { set_err("sample_rate must be > 0"); return nullptr; }
    ir_pipeline *p = new ir_pipeline();
    p->cfg = *cfg;
    p->dev = cfg->device;
    auto fail = [&](const std::string &m) -> ir_pipeline_t * { set_err(m); ir_pipeline_destroy(p); return nullptr; };
    if (cudaSetDevice(p->dev) != cudaSuccess) return fail("cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->dev) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
    p->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) return fail(std::string("device '") + prop.name + "' is not sm_100; kernels are built for sm_100a only");
    derive_det_config(p->dc, cfg->sample_rate, cfg->fft_size, cfg->burst_width_hz, cfg->threshold_db);
    if (p->dc.L < 10 || p->dc.L > 14) return fail("detector FFT size must be 1024..16384 (sample rates ~0.7-23 MHz)");
    if (p->dc.N < IR_SCAN_THREADS) return fail("detector FFT size below 1024 not supported");
    p->dec = (int)roundf((float)cfg->sample_rate / IR_OUT_RATE);       // burst_downmix.c:420
    if (p->dec != 40 && p->dec != 48) return fail("decimation ratio (sample_rate/250k) must be 40 or 48 in this build");
    build_host_tables(p->tab, p->dc.N);
    if ((int)p->tab.h_input.size() != IR_INPUT_NTAPS) return fail("unexpected input filter length");
    // the state machine is the serial spine of a run: its launches go first when SMs free up
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithFlags(&p->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->st_fft, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&p->st_scan, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->st_burst, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->st_chain, cudaStreamNonBlocking) != cudaSuccess)
        return fail("stream creation failed");
    for (int i = 0; i < ir_pipeline::kDemodStreams; i++)
        if (cudaStreamCreateWithFlags(&p->st_demod[i], cudaStreamNonBlocking) != cudaSuccess) return fail("stream creation failed");
    p->pin_arena.host = true;
    if (p->d_window.ensure(p->dc.N)) return fail(g_err);
    if (cudaMemcpy(p->d_window.p, p->tab.det_window.data(), sizeof(float) * p->dc.N, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail("window upload failed");
    if (upload_f2(p->d_tw_det, build_twiddle_image(p->dc.L)) || upload_f2(p->d_tw12, build_twiddle_image(12)) ||
        upload_f2(p->d_tw11, build_twiddle_image(11)) || upload_f2(p->d_sync_dl, p->tab.sync_dl_fft) ||
        upload_f2(p->d_sync_ul, p->tab.sync_ul_fft))
        return fail(g_err);
    if (upload_input_taps(p->tab.h_input.data(), (int)p->tab.h_input.size()) != cudaSuccess ||
        upload_chain_tables(p->tab) != cudaSuccess)
        return fail("constant upload failed");
    if (p->d_base.ensure(p->dc.N) || p->d_hist.ensure((size_t)p->dc.N * p->dc.hist_size) || p->d_state.ensure(1))
        return fail(g_err);
    {
        const char *env = getenv("IR_SCAN");
        if (!env || !*env || strcmp(env, "seg") == 0) p->scan_mode = seg_scan_supported(p->dc) ? 2 : (stream_scan_supported(p->dc) ? 0 : 1);
        else p->scan_mode = strcmp(env, "stream") == 0 && stream_scan_supported(p->dc) ? 0 : 1;
        if (p->scan_mode == 2) {
            if (p->d_ref[0].ensure(p->dc.N) || p->d_ref[1].ensure(p->dc.N) ||
                cudaStreamCreateWithFlags(&p->st_cls, cudaStreamNonBlocking) != cudaSuccess)
                return fail(g_err);
        }
        if (p->scan_mode == 0) {
            const size_t W2 = (size_t)p->dc.N / 16;
            if (p->d_xu[0].ensure((size_t)IR_STREAM_MAX_FRAMES * W2) || p->d_xu[1].ensure((size_t)IR_STREAM_MAX_FRAMES * W2) ||
                p->d_ref[0].ensure(p->dc.N) || p->d_ref[1].ensure(p->dc.N) ||
                cudaStreamCreateWithFlags(&p->st_cls, cudaStreamNonBlocking) != cudaSuccess ||
                p->d_undo.ensure((size_t)p->dc.N * IR_STREAM_MAX_FRAMES) || p->d_base_snap.ensure(p->dc.N) ||
                p->d_ctl.ensure(1))
                return fail(g_err);
            if (cudaMemset(p->d_ctl.p, 0, sizeof(StreamCtl)) != cudaSuccess) return fail("control block init failed");
        }
    }
    memset(&p->res, 0, sizeof(p->res));
    return p;
}

extern "C" void ir_pipeline_destroy(ir_pipeline_t *p) {
    if (!p) return;
    cudaSetDevice(p->dev);
    cudaDeviceSynchronize();
    for (auto &s : p->rot_slabs) cudaFree(s.first);
    p->d_window.release(); p->d_tw_det.release(); p->d_tw12.release(); p->d_tw11.release();
    p->d_sync_dl.release(); p->d_sync_ul.release(); p->d_iq.release(); p->d_mag.release();
    p->d_base.release(); p->d_hist.release(); p->d_state.release();
    p->d_xu[0].release(); p->d_xu[1].release(); p->d_ref[0].release(); p->d_ref[1].release();
    p->d_undo.release(); p->d_base_snap.release();
    p->d_rowany[0].release(); p->d_rowany[1].release(); p->d_seg_raw.release(); p->d_seg_snap.release(); p->d_seg_qmag.release();
    p->d_frame_src.release(); p->d_class.release();
    for (auto e : p->ev_cls) if (e) cudaEventDestroy(e);
    if (p->st_cls) cudaStreamDestroy(p->st_cls);
    p->d_ctl.release();
    p->dev_arena.release(); p->pin_arena.release();
    if (p->h_gone) cudaFreeHost(p->h_gone);
    if (p->h_hdr) cudaFreeHost(p->h_hdr);
    for (auto e : p->ev_pool) cudaEventDestroy(e);
    if (p->st_copy) cudaStreamDestroy(p->st_copy);
    if (p->st_fft) cudaStreamDestroy(p->st_fft);
    if (p->st_scan) cudaStreamDestroy(p->st_scan);
    if (p->st_burst) cudaStreamDestroy(p->st_burst);
    if (p->st_chain) cudaStreamDestroy(p->st_chain);
    for (int i = 0; i < ir_pipeline::kDemodStreams; i++)
        if (p->st_demod[i]) cudaStreamDestroy(p->st_demod[i]);
    delete p;
}

extern "C" int ir_pipeline_reset(ir_pipeline_t *p) {
    if (!p) return -1;
    CK(cudaSetDevice(p->dev));
    CK(cudaMemsetAsync(p->d_state.p, 0, sizeof(DetState), p->st_scan));
    CK(cudaMemsetAsync(p->d_base.p, 0, sizeof(float) * p->dc.N, p->st_scan));
    return 0;
}

// ------------------------------------------------------------------------------------------
// NCO phase-checkpoint tables, one per detector bin that ever carried a burst; carved from slabs
// that live until the pipeline is destroyed (no cudaFree while kernels are in flight).
static float2 *rot_alloc(ir_pipeline *p, size_t count) {
    const size_t bytes = (count * sizeof(float2) + 255) & ~(size_t)255;
    if (p->rot_slabs.empty() || p->rot_slab_cap - p->rot_slabs.back().second < bytes) {
        const size_t cap = std::max<size_t>((size_t)16 << 20, bytes);
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, cap);
        if (e != cudaSuccess) { set_err(std::string("cudaMalloc(rot tables): ") + cudaGetErrorString(e)); return nullptr; }
        p->rot_slabs.push_back({(unsigned char *)q, 0});
        p->rot_slab_cap = cap;
    }
    auto &s = p->rot_slabs.back();
    float2 *r = (float2 *)(s.first + s.second);
    s.second += bytes;
    return r;
}

// Small parameter blocks go pinned host -> device through a kernel that reads the pinned memory
// directly (UVA), NOT through cudaMemcpyAsync: the H2D copy engine's queue holds the bulk IQ copies of
// the whole run, and a parameter copy queued behind them would hold the burst kernels back until the
// last sample has arrived (measured: with 32 Mi-sample chunks every FIR started after the last copy).
__global__ void k_fetch_words(const uint32_t *__restrict__ src_pinned, uint32_t *__restrict__ dst, size_t n_words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src_pinned[i];
}
static cudaError_t fetch_to_device(void *dst, const void *src_pinned, size_t bytes, cudaStream_t st) {
    const size_t n_words = (bytes + 3) / 4;                    // (arena blocks are 256-byte multiples)
    if (n_words == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n_words + 255) / 256, 64);
    k_fetch_words<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(src_pinned), reinterpret_cast<uint32_t *>(dst), n_words);
    return cudaGetLastError();
}

// Bursts [b0, b1) of the gone list: host bookkeeping (burst_data_t equivalents, decimation
// geometry), then FIR + chain + demod + result copies enqueued on st_burst.  Does not wait.
static int launch_wave(ir_pipeline *p, size_t b0, size_t b1, const void *iq_dev, size_t n, int fmt) {
    const DetConfig &dc = p->dc;
    cudaStream_t st = p->st_burst;
    const size_t nb = b1 - b0;
    const uint64_t B = p->cfg.feed_block > 0 ? (uint64_t)p->cfg.feed_block : 32768;
    const uint64_t R = dc.ringbuf_size;
    const int fs = dc.sample_rate, N = dc.N;
    Wave w;
    w.b0 = b0; w.nb = nb;
    const size_t nsym2 = 2 * (size_t)IR_MAX_SYMS;
    w.h_bp = p->pin_arena.take<BurstParam>(nb);
    w.h_tile_start = p->pin_arena.take<int>(nb + 1);
    w.h_co = p->pin_arena.take<ChainOut>(nb);
    w.h_do = p->pin_arena.take<DemodOut>(nb);
    w.h_bits = p->pin_arena.take<unsigned char>(nb * nsym2);
    w.h_llr = p->pin_arena.take<float>(nb * nsym2);
    if (!w.h_bp || !w.h_tile_start || !w.h_co || !w.h_do || !w.h_bits || !w.h_llr) return -1;
    p->bursts.resize(b1, ir_burst_t{});
    p->h_bp.resize(b1, BurstParam{});
    p->frame_ptr.resize(b1, nullptr);
    p->dec_ptr.resize(b1, nullptr);
    std::vector<size_t> need_bins;
    int64_t dec_total = 0;
    int n_tiles = 0;
    // ---- burst_data_t equivalents (burst_detect.c:703-742) and decimation geometry
    for (size_t i = b0; i < b1; i++) {
        const GoneBurst &g = p->h_gone[i];
        ir_burst_t &ob = p->bursts[i];
        ob.id = g.id; ob.start = g.start; ob.stop = g.stop; ob.last_active = g.last_active;
        ob.center_bin = g.center_bin;
        ob.magnitude = 10.0f * log10f(g.peak_rel * dc.hist_size * 1.72f);            // :572
        ob.noise = 10.0f * log10f(g.base_at_create / dc.hist_size / ((float)N * N) / 1.72f /
                                  ((float)fs / N));                                 // :583-586
        // feed call in which the frame at `stop` was processed, and what the ring looked like
        uint64_t calls = (g.stop + (uint64_t)N + B - 1) / B;
        uint64_t emit = calls * B;
        if (emit > n) emit = n;
        uint64_t before = (calls - 1) * B;                     // sample_count seen by ringbuf_write
        uint64_t ring_start = before > R ? before - R : 0;     // :396-398
        uint64_t xs = g.start < ring_start ? ring_start : g.start;
        uint64_t xe = g.stop + (uint64_t)dc.pre_len;
        uint64_t ns = xe > xs ? xe - xs : 0;
        ob.num_samples = ns;
        ob.emit_count = emit;
        BurstParam &bp = p->h_bp[i];
        bp.start = (int64_t)xs;
        bp.emit_count = (int64_t)emit;
        int nn = (int)std::min<uint64_t>(ns, IR_DM_WORK);      // burst_downmix.c:650-651
        bp.n = nn;
        int dlen = 0;
        if (ns >= 100) {                                       // :645
            dlen = (nn - IR_INPUT_NTAPS + 1) / p->dec;         // :423
            if (dlen < 0) dlen = 0;
        }
        bp.dec_len = dlen;
        bp.dec_off = dec_total;                                // relative to this wave's arrays
        bp.tile0 = n_tiles;
        w.h_tile_start[i - b0] = n_tiles;
        const float rel = (g.center_bin - N / 2) / (float)N;   // :663-664
        const float ph = -2.0f * (float)M_PI * rel;
        float sn, cs;
        sincosf(ph, &sn, &cs);                                 // what cexpf(ph*I) evaluates (:669)
        bp.incr_coarse = make_float2(cs, sn);
        bp.cfreq_coarse = p->cfg.center_frequency + (double)(rel * fs);   // :671
        bp.simplex = 0;
        ob.dec_len = dlen;
        if (dlen >= 100) {
            dec_total += dlen;
            n_tiles += (dlen + IR_FIR_TILE_OF(p->dec) - 1) / IR_FIR_TILE_OF(p->dec);
            auto it = p->rot.find(g.center_bin);
            if (it == p->rot.end() || it->second.len < nn + IR_ROT_G) need_bins.push_back(i);
            p->alg += 8ull * (uint64_t)nn + 8ull * (uint64_t)dlen;
        } else {
            bp.dec_len = 0;        // chain reports status 2
        }
    }
    w.h_tile_start[nb] = n_tiles;
    w.n_tiles = n_tiles; w.dec_total = dec_total;
    // ---- NCO checkpoint tables for bins not cached yet (or cached too short)
    if (!need_bins.empty()) {
        std::unordered_map<int, int> want;     // bin -> samples
        for (size_t i : need_bins) {
            int bin = p->h_gone[i].center_bin;
            int len = ((p->h_bp[i].n + IR_ROT_G + 65535) / 65536) * 65536;
            auto wi = want.find(bin);
            if (wi == want.end() || wi->second < len) want[bin] = len;
        }
        const size_t k = want.size();
        float2 *h_incr = p->pin_arena.take<float2>(k), *d_incr = p->dev_arena.take<float2>(k);
        float2 **h_ptrs = p->pin_arena.take<float2 *>(k), **d_ptrs = p->dev_arena.take<float2 *>(k);
        int *h_lens = p->pin_arena.take<int>(k), *d_lens = p->dev_arena.take<int>(k);
        if (!h_incr || !d_incr || !h_ptrs || !d_ptrs || !h_lens || !d_lens) return -1;
        size_t j = 0;
        for (auto &kv : want) {
            RotTable &rt = p->rot[kv.first];
            rt.len = kv.second;
            rt.d = rot_alloc(p, (size_t)(rt.len / IR_ROT_G + 1));
            if (!rt.d) return -1;
            const float rel = (kv.first - N / 2) / (float)N;
            const float ph = -2.0f * (float)M_PI * rel;
            float sn, cs;
            sincosf(ph, &sn, &cs);
            rt.incr = make_float2(cs, sn);
            h_incr[j] = rt.incr; h_ptrs[j] = rt.d; h_lens[j] = rt.len;
            j++;
        }
        CK(fetch_to_device(d_incr, h_incr, k * sizeof(float2), st));
        CK(fetch_to_device(d_ptrs, h_ptrs, k * sizeof(float2 *), st));
        CK(fetch_to_device(d_lens, h_lens, k * sizeof(int), st));
        CK(launch_rot_tables(d_incr, d_ptrs, d_lens, (int)k, st));
        p->res.kernel_launches += 4;                           // 3 parameter fetches + the tables
    }
    for (size_t i = b0; i < b1; i++) {
        auto it = p->rot.find(p->h_gone[i].center_bin);
        p->h_bp[i].rot_table = it != p->rot.end() ? it->second.d : nullptr;
        w.h_bp[i - b0] = p->h_bp[i];
    }
    // ---- device work for the bursts
    w.e0 = p->ev(); w.e1 = p->ev(); w.e1a = p->ev(); w.e1b = p->ev(); w.e2 = p->ev(); w.e3 = p->ev(); w.e_done = p->ev();
    w.d_bp = p->dev_arena.take<BurstParam>(nb);
    w.d_tile_start = p->dev_arena.take<int>(nb + 1);
    w.h_tile_burst = p->pin_arena.take<int>((size_t)n_tiles + 1);
    w.d_tile_burst = p->dev_arena.take<int>((size_t)n_tiles + 1);
    if (!w.h_tile_burst || !w.d_tile_burst) return -1;
    for (size_t i = b0; i < b1; i++)
        for (int t = w.h_tile_start[i - b0]; t < w.h_tile_start[i - b0 + 1]; t++) w.h_tile_burst[t] = (int)(i - b0);
    w.d_dec = p->dev_arena.take<float2>((size_t)dec_total + 16);
    w.d_scrA = p->dev_arena.take<float2>((size_t)dec_total + 16);
    w.d_scrB = p->dev_arena.take<float2>((size_t)dec_total + 16);
    w.d_frames = p->dev_arena.take<float2>(nb * (size_t)IR_MAX_FRAME);
    w.d_co = p->dev_arena.take<ChainOut>(nb);
    w.d_do = p->dev_arena.take<DemodOut>(nb);
    w.d_bits = p->dev_arena.take<unsigned char>(nb * nsym2);
    w.d_llr = p->dev_arena.take<float>(nb * nsym2);
    if (!w.d_bp || !w.d_tile_start || !w.d_dec || !w.d_scrA || !w.d_scrB || !w.d_frames || !w.d_co || !w.d_do ||
        !w.d_bits || !w.d_llr)
        return -1;
    for (size_t i = b0; i < b1; i++) {
        p->frame_ptr[i] = w.d_frames + (i - b0) * (size_t)IR_MAX_FRAME;
        p->dec_ptr[i] = w.d_dec + p->h_bp[i].dec_off;
    }
    CK(fetch_to_device(w.d_bp, w.h_bp, nb * sizeof(BurstParam), st));
    CK(fetch_to_device(w.d_tile_start, w.h_tile_start, (nb + 1) * sizeof(int), st));
    if (n_tiles > 0) CK(fetch_to_device(w.d_tile_burst, w.h_tile_burst, (size_t)n_tiles * sizeof(int), st));
    p->res.h2d_bytes += nb * sizeof(BurstParam) + (nb + 1) * sizeof(int);
    CK(cudaEventRecord(w.e0, st));
    CK(launch_fir(fmt, p->dec, iq_dev, (int64_t)n, R, w.d_bp, w.d_tile_start, w.d_tile_burst, (int)nb, n_tiles, w.d_dec, st));
    CK(cudaEventRecord(w.e1, st));
    // the 250 kHz chain of this wave overlaps the FIR of the next one
    CK(cudaStreamWaitEvent(p->st_chain, w.e1, 0));
    CK(cudaEventRecord(w.e1a, p->st_chain));
    CK(launch_chain(w.d_bp, (int)nb, w.d_dec, w.d_scrA, w.d_scrB, p->d_tw12.p, p->d_tw11.p, p->d_sync_dl.p,
                    p->d_sync_ul.p, w.d_co, w.d_frames, p->st_chain));
    CK(cudaEventRecord(w.e1b, p->st_chain));
    cudaStream_t sd = p->st_demod[p->waves.size() % ir_pipeline::kDemodStreams];
    CK(cudaStreamWaitEvent(sd, w.e1b, 0));
    CK(cudaEventRecord(w.e2, sd));
    // frames are 191 symbols at most outside the simplex band (burst_downmix.c:764-770; the fine CFO moves the centre
    // by less than 125 kHz): the slicer's shared-memory footprint follows
    int max_frame = 1910;
    for (size_t i = b0; i < b1; i++)
        if (p->h_bp[i].cfreq_coarse > 1626000000.0 - 200000.0) max_frame = IR_MAX_FRAME;
    CK(launch_demod(w.d_co, (int)nb, w.d_frames, p->cfg.use_gardner, w.d_do, w.d_bits, w.d_llr, max_frame, sd));
    CK(cudaEventRecord(w.e3, sd));
    p->res.kernel_launches += (n_tiles > 0 ? 2 : 0) + 2 + 2;   // FIR, chain, demod + 3 parameter fetches
    CK(cudaMemcpyAsync(w.h_co, w.d_co, nb * sizeof(ChainOut), cudaMemcpyDeviceToHost, sd));
    CK(cudaMemcpyAsync(w.h_do, w.d_do, nb * sizeof(DemodOut), cudaMemcpyDeviceToHost, sd));
    CK(cudaMemcpyAsync(w.h_bits, w.d_bits, nb * nsym2, cudaMemcpyDeviceToHost, sd));
    CK(cudaMemcpyAsync(w.h_llr, w.d_llr, nb * nsym2 * sizeof(float), cudaMemcpyDeviceToHost, sd));
    p->res.d2h_bytes += nb * (sizeof(GoneBurst) + sizeof(ChainOut) + sizeof(DemodOut) + nsym2 * 5);
    CK(cudaEventRecord(w.e_done, sd));
    p->waves.push_back(w);
    return 0;
}

// ---- printf's %f, exactly, without printf: the RAW: line has four floating conversions and glibc's printf_fp costs
// ~0.25 us each; at 10^5 frames/s of sink rate that is what the host spends its time on.  A double is m * 2^e: the
// integer part and the 52-bit fraction are taken apart exactly, fraction * 10^decimals is an exact 128-bit integer,
// and the division by 2^52 rounds to nearest, ties to even -- the rule printf applies to the exact binary value.
// Falls back to snprintf outside the range this path needs (|x| >= 2^53, NaN, inf, decimals > 9).
static inline int fmt_fixed(char *dst, double x, int decimals, int width, bool zero_pad, bool plus) {
    const bool neg = std::signbit(x);
    const double ax = neg ? -x : x;
    if (!(ax < 9007199254740992.0) || decimals > 9) {
        char f[16];
        snprintf(f, sizeof f, "%%%s%s%d.%df", plus ? "+" : "", zero_pad ? "0" : "", width, decimals);
        return snprintf(dst, 400, f, x);
    }
    static const uint64_t P10[10] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull};
    uint64_t ip = (uint64_t)ax;                              // exact: ax < 2^53
    const double fr = ax - (double)ip;                       // exact, in [0, 1)
    const uint64_t f52 = (uint64_t)(fr * 4503599627370496.0);   // fraction * 2^52: exact (fr has at most 52 fraction bits... 
    // ... when ax >= 1; below 1 the fraction can have more bits, handled by the remainder test below)
    const double rem_lo = fr * 4503599627370496.0 - (double)f52;   // what f52 dropped (non-zero only for tiny ax), in [0, 1)
    unsigned __int128 prod = (unsigned __int128)f52 * P10[decimals];
    uint64_t q = (uint64_t)(prod >> 52);
    const uint64_t r = (uint64_t)(prod & ((1ull << 52) - 1ull));
    const uint64_t half = 1ull << 51;
    // the dropped low bits add rem_lo * 10^decimals / 2^52 < 10^decimals * 2^-52 to the remainder: they can only matter at an exact tie
    bool up = r > half || (r == half && ((q & 1ull) || rem_lo > 0.0));
    if (r < half && rem_lo > 0.0) {                           // could the dropped bits lift r over the half?  (only for |x| < 1)
        const double extra = rem_lo * (double)P10[decimals];  // in units of 2^-52 of the remainder... compared exactly enough:
        if ((double)(half - r) <= extra) {                    // too close to call with this arithmetic: let printf decide
            char f[16];
            snprintf(f, sizeof f, "%%%s%s%d.%df", plus ? "+" : "", zero_pad ? "0" : "", width, decimals);
            return snprintf(dst, 400, f, x);
        }
    }
    if (up) { q++; if (q == P10[decimals]) { q = 0; ip++; } }
    char tmp[48];
    int n = 0;
    for (int i = 0; i < decimals; i++) { tmp[n++] = (char)('0' + q % 10); q /= 10; }
    if (decimals > 0) tmp[n++] = '.';
    do { tmp[n++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    const bool sign = neg || plus;
    const int len = n + (sign ? 1 : 0);
    int pos = 0;
    if (!zero_pad) for (int i = len; i < width; i++) dst[pos++] = ' ';
    if (sign) dst[pos++] = neg ? '-' : '+';
    if (zero_pad) for (int i = len; i < width; i++) dst[pos++] = '0';
    while (n > 0) dst[pos++] = tmp[--n];
    dst[pos] = 0;
    return pos;
}
static inline int fmt_uint(char *dst, unsigned long long v, int width, char pad) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    int pos = 0;
    for (int i = n; i < width; i++) dst[pos++] = pad;
    while (n > 0) dst[pos++] = tmp[--n];
    return pos;
}

// frame_output_print's line (frame_output.c:160-199) from the field after file_info on:
//   " %012.4f %010d N:%05.2f%+06.2f I:%011llu %3d%% %.5f %3d " + bits + newline
static int format_raw_rest(char *dst, size_t cap, uint64_t t0, const ir_frame_t *f, const uint8_t *bits) {
    const double ts_ms = (double)(f->timestamp - t0) / 1000000.0;
    const int fhz = (int)(f->center_frequency + 0.5);
    int pay = f->n_payload_symbols;
    if (pay < 0) pay = 0;
    if (cap < 160 || fhz < 0 || f->confidence < 0) {         // (room for the fixed fields; odd values: printf's own rules)
        int k = snprintf(dst, cap, " %012.4f %010d N:%05.2f%+06.2f I:%011llu %3d%% %.5f %3d ", ts_ms, fhz, f->magnitude,
                         f->noise, (unsigned long long)f->id, f->confidence, f->level, pay);
        if (k < 0) return -1;
        size_t pos = (size_t)k < cap ? (size_t)k : cap - 1;
        for (int i = 0; i < f->n_bits && pos + 2 < cap; i++) dst[pos++] = (char)('0' + bits[i]);
        if (pos + 1 < cap) dst[pos++] = '\n';
        dst[pos] = 0;
        return (int)pos;
    }
    size_t pos = 0;
    dst[pos++] = ' ';
    pos += (size_t)fmt_fixed(dst + pos, ts_ms, 4, 12, true, false);
    dst[pos++] = ' ';
    pos += (size_t)fmt_uint(dst + pos, (unsigned long long)fhz, 10, '0');
    dst[pos++] = ' '; dst[pos++] = 'N'; dst[pos++] = ':';
    pos += (size_t)fmt_fixed(dst + pos, (double)f->magnitude, 2, 5, true, false);
    pos += (size_t)fmt_fixed(dst + pos, (double)f->noise, 2, 6, true, true);
    dst[pos++] = ' '; dst[pos++] = 'I'; dst[pos++] = ':';
    pos += (size_t)fmt_uint(dst + pos, (unsigned long long)f->id, 11, '0');
    dst[pos++] = ' ';
    pos += (size_t)fmt_uint(dst + pos, (unsigned long long)f->confidence, 3, ' ');
    dst[pos++] = '%'; dst[pos++] = ' ';
    pos += (size_t)fmt_fixed(dst + pos, (double)f->level, 5, 0, false, false);
    dst[pos++] = ' ';
    pos += (size_t)fmt_uint(dst + pos, (unsigned long long)pay, 3, ' ');
    dst[pos++] = ' ';
    for (int i = 0; i < f->n_bits && pos + 2 < cap; i++) dst[pos++] = (char)('0' + bits[i]);
    if (pos + 1 < cap) dst[pos++] = '\n';
    dst[pos] = 0;
    return (int)pos;
}

// test hook: printf's "%[+][0]<width>.<decimals>f" by the exact path above
extern "C" int ir_format_fixed(char *dst, double x, int decimals, int width, int zero_pad, int plus) {
    return fmt_fixed(dst, x, decimals, width, zero_pad != 0, plus != 0);
}

// demod_frame_t equivalents (qpsk_demod.c:505-527) of one finished wave, on the host, in double
static int assemble_wave(ir_pipeline *p, const Wave &w) {
    CK(cudaEventSynchronize(w.e_done));
    const int fs = p->dc.sample_rate;
    const uint64_t delay_ns = (uint64_t)((IR_INPUT_NTAPS / 2) * 1000000000ULL / fs);      // burst_downmix.c:431-433
    const size_t nsym2 = 2 * (size_t)IR_MAX_SYMS;
    for (size_t i = w.b0; i < w.b0 + w.nb; i++) {
        ir_burst_t &ob = p->bursts[i];
        const ChainOut &c = w.h_co[i - w.b0];
        const DemodOut &d = w.h_do[i - w.b0];
        ob.downmix_status = p->h_bp[i].dec_len >= 100 ? c.status : (ob.num_samples < 100 ? 1 : 2);
        ob.demod_ok = 0;
        ob.center_offset = c.center_offset; ob.dm_start = c.start; ob.uw_start = c.uw_start;
        ob.frame_len = c.frame_len; ob.uw_start_frac = c.uw_corr; ob.dm_direction = c.direction;
        if (ob.downmix_status != 0) continue;
        p->alg_sink += 8ull * (uint64_t)c.frame_len;
        if (!d.ok) continue;
        ob.demod_ok = 1;
        ir_frame_t f;
        f.id = ob.id;
        uint64_t ts = p->start_time_ns + (uint64_t)((double)(ob.start + p->sample_origin) / fs * 1e9);          // :659-660
        ts += delay_ns;
        f.timestamp = ts + (uint64_t)((double)c.start / IR_OUT_RATE * 1e9);               // :783
        double cf = p->h_bp[i].cfreq_coarse;
        cf += (double)(c.center_offset * (float)IR_OUT_RATE);                             // :719
        if (d.n_symbols > 0) {                                                            // qpsk_demod.c:521-527
            double dur = (double)d.n_symbols / 25000;
            cf = cf + d.total_phase / dur / M_PI / 2.0;
        }
        f.center_frequency = cf;
        f.direction = d.direction;
        f.magnitude = ob.magnitude; f.noise = ob.noise;
        f.confidence = d.confidence; f.level = d.level;
        f.n_symbols = d.n_symbols; f.n_payload_symbols = d.n_symbols - 12;
        f.n_bits = 2 * d.n_symbols;
        f.bits_offset = (uint32_t)p->bits.size();
        const uint8_t *br = w.h_bits + (i - w.b0) * nsym2;
        const float *lr = w.h_llr + (i - w.b0) * nsym2;
        p->bits.insert(p->bits.end(), br, br + f.n_bits);
        p->llr.insert(p->llr.end(), lr, lr + f.n_bits);
        p->frames.push_back(f);
        p->frame_src.push_back(FrameSrc{w.d_bits + (i - w.b0) * nsym2, w.d_llr + (i - w.b0) * nsym2, f.n_bits, f.direction});
        p->alg_sink += 8ull * (uint64_t)c.frame_len + (uint64_t)f.n_bits;
        {   // its RAW: line, while the GPU is busy with later waves
            if (p->frames.size() == 1) p->raw_t0 = (f.timestamp / 1000000000ULL) * 1000000000ULL;
            char line[160 + 2 * IR_MAX_SYMS];
            const int ln = format_raw_rest(line, sizeof(line), p->raw_t0, &f, br);
            p->raw_off.push_back(p->raw_rest.size());
            if (ln > 0) p->raw_rest.append(line, (size_t)ln);
        }
    }
    return 0;
}

static float span(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// Frames [f0, f1) of the run through the streaming state machine, on st_scan.  The run starts
// from a reset detector, so the first hist_size frames are the priming launch (every frame quiet,
// no bitmaps); after that each launch is: bitmaps against the current baseline, snapshot, state
// machine, and -- a no-op unless the launch bailed -- the cluster kernel, which first restores the snapshot.
static int scan_stream_range(ir_pipeline *p, int64_t f0, int64_t f1, cudaEvent_t fft_done) {
    const DetConfig &dc = p->dc;
    const int N = dc.N;
    cudaStream_t st = p->st_scan;
    const bool dbg = p->scan_dbg;
    for (int64_t a = f0; a < f1;) {
        const bool priming = a < dc.hist_size;
        const int64_t b = priming ? std::min<int64_t>(f1, dc.hist_size) : std::min<int64_t>(f1, a + IR_STREAM_MAX_FRAMES);
        const int nf = (int)(b - a);
        const float *mag = p->d_mag.p + a * N;
        cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (dbg) { for (auto &x : e) x = p->ev(); CK(cudaEventRecord(e[0], st)); }
        // Bitmaps of this launch, on their own stream.  They are made against the baseline as it
        // stands when the pass runs -- any recent value is as good as another inside the guard band,
        // and the workers check theirs against exactly the reference the pass recorded -- so the pass
        // of launch k only waits for launch k-2 (its buffers are free, the detector is primed) and
        // overlaps launch k-1; right after priming it has to wait for the priming launch itself.
        const size_t k = p->ev_scan_done.size();
        const int buf = (int)(k & 1);
        if (!priming) {
            const bool after_priming = k == 0 || a <= dc.hist_size;
            if (k >= 1) CK(cudaStreamWaitEvent(p->st_cls, p->ev_scan_done[after_priming || k < 2 ? k - 1 : k - 2], 0));
            CK(cudaStreamWaitEvent(p->st_cls, fft_done, 0));
            CK(launch_detect_classify(mag, p->d_base.p, dc.thr, N, nf, p->d_xu[buf].p, p->d_ref[buf].p, p->sm_count, p->st_cls));
            p->res.kernel_launches++;
            cudaEvent_t ec = p->ev();
            CK(cudaEventRecord(ec, p->st_cls));
            CK(cudaStreamWaitEvent(st, ec, 0));
        }
        if (dbg) { CK(cudaEventRecord(e[1], st)); CK(cudaEventRecord(e[2], st)); }
        CK(launch_detect_scan_stream(dc, p->d_state.p, p->d_base.p, p->d_hist.p, mag, priming ? nullptr : p->d_xu[buf].p,
                                     p->d_ref[buf].p, nf, p->d_gone, p->gone_cap, p->d_ctl.p, p->scan_epoch++, p->d_undo.p,
                                     p->d_base_snap.p, st));
        if (dbg) CK(cudaEventRecord(e[3], st));
        ScanSnapshot snap;
        snap.undo = p->d_undo.p; snap.base = p->d_base_snap.p; snap.ctl = p->d_ctl.p;
        CK(launch_detect_scan_cluster_if(dc, p->d_state.p, p->d_base.p, p->d_hist.p, mag, nf, p->d_gone, p->gone_cap,
                                         &p->d_ctl.p->bailed, snap, st));
        {
            cudaEvent_t ed = p->ev();
            CK(cudaEventRecord(ed, st));
            p->ev_scan_done.push_back(ed);
        }
        if (dbg) { CK(cudaEventRecord(e[4], st)); for (auto x : e) p->scan_dbg_ev.push_back(x); }
        p->res.kernel_launches += 2;
        a = b;
    }
    return 0;
}

// Device memory of the segmented state machine for chunks of up to `fc` frames.
static int seg_ensure(ir_pipeline *p, size_t fc) {
    fc = std::min<size_t>(std::max<size_t>(fc, IR_SEG_LEN), IR_SEG_MAX_FRAMES);
    fc = (fc + IR_SEG_LEN - 1) / IR_SEG_LEN * IR_SEG_LEN;
    if (fc <= p->seg_frames_cap) return 0;
    const size_t N = (size_t)p->dc.N, S = fc / IR_SEG_LEN;
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_ctl = carve(sizeof(SegCtl)), o_a = carve((S + 1) * sizeof(SegState)), o_b = carve((S + 1) * sizeof(SegState)),
                 o_qw = carve((fc / 32 + 2) * 4), o_valid = carve(N / 32 * 4), o_wpre = carve((fc / 32 + 4) * 4),
                 o_ql = carve((fc + 2) * 4), o_sv = carve((fc + 4) * 4), o_fs = carve((fc + 2) * 4), o_nc = carve((S + 2) * 4),
                 o_ng = carve((S + 2) * 4), o_sb = carve((S + 2) * 4), o_sc = carve(2 * (S + 2) * 4), o_cp = carve((S + 2) * 4),
                 o_gp = carve((S + 2) * 4), o_gl = carve(S * (size_t)IR_SEG_GONE * sizeof(GoneBurst)), o_bf = carve(N * 4), o_glo = carve(N * 4), o_ghi = carve(N * 4),
                 o_oa = carve((S + 1) * (size_t)IR_SEG_OVF * sizeof(SegBurst)), o_ob = carve((S + 1) * (size_t)IR_SEG_OVF * sizeof(SegBurst)),
                 o_pr = carve(S * (size_t)IR_SEG_PCAP * 8), o_fv = carve((S + 2) * 4);
    if (p->d_seg_raw.ensure(off) || p->d_seg_snap.ensure((fc + 1) * N) || p->d_seg_qmag.ensure((fc + (size_t)p->dc.hist_size + 128) * N) || p->d_rowany[0].ensure(fc) || p->d_rowany[1].ensure(fc) ||
        p->d_xu[0].ensure(fc * (N / 16)) || p->d_xu[1].ensure(fc * (N / 16)))
        return -1;
    CK(cudaMemset(p->d_seg_raw.p, 0, off));
    unsigned char *b = p->d_seg_raw.p;
    SegBuffers &g = p->seg;
    g.ctl = (SegCtl *)(b + o_ctl); g.stA = (SegState *)(b + o_a); g.stB = (SegState *)(b + o_b);
    g.qw = (uint32_t *)(b + o_qw); g.valid = (uint32_t *)(b + o_valid); g.wpre = (int *)(b + o_wpre);
    g.qlist = (int *)(b + o_ql); g.slotv = (int *)(b + o_sv); g.fslot = (int *)(b + o_fs);
    g.ncreate = (int *)(b + o_nc); g.ngone = (int *)(b + o_ng); g.segbail = (int *)(b + o_sb); g.stch = (int *)(b + o_sc);
    g.cpre = (int *)(b + o_cp); g.gpre = (int *)(b + o_gp); g.glist = (GoneBurst *)(b + o_gl);
    g.bfinal = (float *)(b + o_bf); g.glo = (float *)(b + o_glo); g.ghi = (float *)(b + o_ghi);
    g.ovfA = (SegBurst *)(b + o_oa); g.ovfB = (SegBurst *)(b + o_ob);
    g.gkeys = (unsigned long long *)(b + o_pr); g.seggen = (uint32_t *)(b + o_fv);
    g.qmag = p->d_seg_qmag.p;
    g.snap = p->d_seg_snap.p; g.slot_cap = (int)fc + 1; g.frames_cap = (int)fc;
    p->seg_frames_cap = fc;
    return 0;
}

// Frames [f0, f1) of the run through the segmented state machine, on st_scan.  The first hist_size frames
// of a run prime the detector (k_seg_prime).  After that every chunk of <= IR_SEG_MAX_FRAMES frames is:
// bitmaps against the baseline as it stood two chunks ago (a copy taken in stream order, so the pass of
// chunk k+1 overlaps chunk k and nobody reads a baseline that is being written), the rounds of
// k_detect_seg.cu, and -- a no-op unless the chunk bailed -- the cluster kernel.
static int scan_seg_range(ir_pipeline *p, int64_t f0, int64_t f1, cudaEvent_t fft_done) {
    const DetConfig &dc = p->dc;
    const int N = dc.N;
    cudaStream_t st = p->st_scan;
    for (int64_t a = f0; a < f1;) {
        const bool priming = a < dc.hist_size;
        const int64_t b = priming ? std::min<int64_t>(f1, dc.hist_size) : std::min<int64_t>(f1, a + (int64_t)p->seg_frames_cap);
        const int nf = (int)(b - a);
        const float *mag = p->d_mag.p + a * N;
        const size_t k = p->ev_scan_done.size();
        if (priming) {
            CK(launch_detect_seg_prime(dc, p->d_state.p, p->d_base.p, p->d_hist.p, mag, nf, st));
            p->res.kernel_launches += 2;
            if (b >= dc.hist_size) {       // primed: both reference buffers start from this baseline
                CK(cudaMemcpyAsync(p->d_ref[0].p, p->d_base.p, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(p->d_ref[1].p, p->d_base.p, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
            }
        } else {
            const int buf = (int)(k & 1);
            if (k >= 2) CK(cudaStreamWaitEvent(p->st_cls, p->ev_scan_done[k - 2], 0));
            else if (k >= 1) CK(cudaStreamWaitEvent(p->st_cls, p->ev_scan_done[0], 0));
            CK(cudaStreamWaitEvent(p->st_cls, fft_done, 0));
            CK(launch_detect_classify(mag, p->d_ref[buf].p, dc.thr, N, nf, p->d_xu[buf].p, nullptr, p->sm_count, p->st_cls,
                                      p->d_rowany[buf].p));
            p->res.kernel_launches++;
            cudaEvent_t ec = p->ev();
            CK(cudaEventRecord(ec, p->st_cls));
            CK(cudaStreamWaitEvent(st, ec, 0));
            int nl = 0;
            CK(launch_detect_scan_seg(dc, p->d_state.p, p->d_base.p, p->d_hist.p, mag, p->d_xu[buf].p, p->d_rowany[buf].p,
                                      p->d_ref[buf].p, nf, p->d_gone, p->gone_cap, p->seg, &nl, st));
            CK(launch_detect_scan_cluster_if(dc, p->d_state.p, p->d_base.p, p->d_hist.p, mag, nf, p->d_gone, p->gone_cap,
                                             &p->seg.ctl->bailed, ScanSnapshot{}, st));
            p->res.kernel_launches += nl + 1;
            // the baseline the bitmaps of chunk k+2 are made against
            CK(cudaMemcpyAsync(p->d_ref[buf].p, p->d_base.p, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
        }
        cudaEvent_t ed = p->ev();
        CK(cudaEventRecord(ed, st));
        p->ev_scan_done.push_back(ed);
        a = b;
    }
    return 0;
}

// Whole path over one block.  Every chunk's copy / FFT / scan is enqueued up front; the host then
// follows the detector chunk by chunk and launches the downmix + demod of the bursts each chunk
// emitted (a "wave") on a fourth stream, so that only the last wave runs after the detector is
// done.  The state machine occupies 8 SMs; the waves and the FFTs of later chunks use the rest.
static int run_common(ir_pipeline *p, const void *host_iq, const void *dev_iq, size_t n, int fmt) {
    if (!p) { set_err("null pipeline"); return -1; }
    if (fmt < 0 || fmt > 2) { set_err("bad sample format"); return -1; }
    CK(cudaSetDevice(p->dev));
    const auto t_run0 = std::chrono::steady_clock::now();
    auto run_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_run0).count(); };
    const DetConfig &dc = p->dc;
    const int N = dc.N;
    const size_t bps = (size_t)fmt_bytes(fmt);
    p->ev_used = 0;
    p->res.kernel_launches = 0; p->res.h2d_bytes = 0; p->res.d2h_bytes = 0;
    p->alg = (uint64_t)n * bps;
    p->alg_sink = 0;
    if (p->cfg.start_time_ns) p->start_time_ns = p->cfg.start_time_ns;
    else {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);                     // burst_detect.c:755-759
        p->start_time_ns = ts.tv_sec * 1000000000ULL + ts.tv_nsec;
    }
    const int64_t n_frames = (int64_t)(n / (size_t)N);
    p->n_frames_last = n_frames;
    if (p->d_mag.ensure((size_t)std::max<int64_t>(n_frames, 1) * N)) return -1;
    // burst list: the reference limits the bursts that are ACTIVE (max_bursts), not how many it emits; the most a
    // stream can emit is max_bursts per post_len samples (a burst lives at least that long)
    const size_t per_post = (size_t)std::max(dc.max_bursts, 32);
    const uint32_t gone_cap = (uint32_t)std::min<size_t>((size_t)1 << 24,
        std::max<size_t>(4096, (n / (size_t)std::max(dc.post_len, 1) + 2) * per_post));
    if (gone_cap > p->gone_cap) {
        if (p->h_gone) cudaFreeHost(p->h_gone);
        p->h_gone = nullptr; p->gone_cap = 0;
        CK(cudaHostAlloc((void **)&p->h_gone, (size_t)gone_cap * sizeof(GoneBurst), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void **)&p->d_gone, p->h_gone, 0));
        p->gone_cap = gone_cap;
    }
    const void *iq_dev = dev_iq;
    if (host_iq) {
        if (p->d_iq.ensure(n * bps + 64)) return -1;
        iq_dev = p->d_iq.p;
    }
    // chunking: copies (host input only) overlap the detector kernels of earlier chunks
    size_t chunk = p->cfg.h2d_chunk > 0 ? (size_t)p->cfg.h2d_chunk : ((size_t)32 << 20);
    if (const char *env = getenv("IR_CHUNK_MI")) { const long v = atol(env); if (v > 0 && v <= 1024) chunk = (size_t)v << 20; }
    // device-resident input in segmented mode: nothing to overlap with copies, so the chunks are as long as
    // one pass of the segmented state machine takes (more segments side by side per round)
    if (!host_iq && p->scan_mode == 2 && p->cfg.h2d_chunk == 0 && !getenv("IR_CHUNK_MI")) chunk = (size_t)IR_SEG_MAX_FRAMES * N;
    chunk = std::max<size_t>(chunk / N, 1) * N;
    if (p->scan_mode == 2 && seg_ensure(p, std::min<size_t>(chunk / N, (size_t)std::max<int64_t>(n_frames, 1)))) return -1;
    // chunk boundaries: full chunks, then the last stretch in halves (ir_plan_chunks)
    std::vector<size_t> bounds((n / std::max<size_t>(chunk, 1)) + 64);
    {
        long nb = 0;
        if (host_iq) {
            nb = ir_plan_chunks(n, chunk, (size_t)N, bounds.data(), bounds.size());
        } else {
            // device-resident input: no copy to hide the tail behind -- equal chunks, after a short first one (the
            // downmix of its bursts starts while the state machine is busy with the second)
            size_t off = 0;
            if (p->scan_mode == 2 && n > 2 * chunk) {
                off = std::min(n, (size_t)(p->dc.hist_size + 4096) * N);
                bounds[(size_t)nb++] = off;
            }
            for (; off < n; off += chunk) bounds[(size_t)nb++] = std::min(n, off + chunk);
        }
        if (nb < 0) { set_err("chunk plan does not fit"); return -1; }
        bounds.resize((size_t)nb);
        if (host_iq && p->scan_mode == 2 && bounds.size() > 1) {
            // a chunk of the segmented state machine costs its rounds whatever its length (~0.4 ms): the tail of
            // halves stops at 4 Mi samples
            const size_t min_seg = std::max<size_t>(((size_t)4 << 20) / N, 1) * N;
            std::vector<size_t> kept;
            size_t prev = 0;
            for (size_t i = 0; i + 1 < bounds.size(); i++)
                if (bounds[i] - prev >= min_seg && bounds.back() - bounds[i] >= min_seg / 2) { kept.push_back(bounds[i]); prev = bounds[i]; }
            kept.push_back(bounds.back());
            bounds.swap(kept);
        }
        if (bounds.empty()) bounds.push_back(0);
    }
    const size_t n_chunks = bounds.size();
    if (n_chunks > p->hdr_slots) {
        if (p->h_hdr) cudaFreeHost(p->h_hdr);
        p->h_hdr = nullptr; p->hdr_slots = 0;
        CK(cudaHostAlloc((void **)&p->h_hdr, (n_chunks + 8) * kHdrBytes, cudaHostAllocDefault));
        p->hdr_slots = n_chunks + 8;
    }
    if (p->dev_arena.begin((size_t)64 << 20) || p->pin_arena.begin((size_t)8 << 20)) return -1;
    p->waves.clear(); p->chunks.clear();
    p->waves.reserve(n_chunks + 8);                          // (the sink thread reads earlier entries while later ones are pushed)
    p->bursts.clear(); p->h_bp.clear(); p->frame_ptr.clear(); p->dec_ptr.clear();
    p->bursts.reserve(p->gone_cap); p->h_bp.reserve(p->gone_cap);
    p->frames.clear(); p->bits.clear(); p->llr.clear(); p->frame_src.clear();
    p->raw_rest.clear(); p->raw_off.clear(); p->raw_t0 = 0;
    p->scan_dbg = getenv("IR_SCAN_DEBUG") != nullptr;
    p->scan_dbg_ev.clear();
    p->ev_scan_done.clear();
    if (ir_pipeline_reset(p)) return -1;
    if (p->scan_mode == 0) CK(cudaMemsetAsync(p->d_ctl.p, 0, sizeof(StreamCtl), p->st_scan));
    if (p->scan_mode == 2) CK(cudaMemsetAsync(p->seg.ctl, 0, sizeof(SegCtl), p->st_scan));
    cudaEvent_t ev_first_copy = nullptr, ev_last_copy = nullptr;
    cudaEvent_t ev_begin = p->ev();
    CK(cudaEventRecord(ev_begin, p->st_scan));          // after the state reset
    CK(cudaStreamWaitEvent(p->st_fft, ev_begin, 0));
    for (size_t bi = 0, off = 0; bi < bounds.size() && off < n; off = bounds[bi], bi++) {
        const size_t m = bounds[bi] - off;
        if (host_iq) {
            CK(cudaMemcpyAsync((unsigned char *)p->d_iq.p + off * bps, (const unsigned char *)host_iq + off * bps,
                               m * bps, cudaMemcpyHostToDevice, p->st_copy));
            cudaEvent_t e = p->ev();
            CK(cudaEventRecord(e, p->st_copy));
            CK(cudaStreamWaitEvent(p->st_fft, e, 0));
            p->res.h2d_bytes += m * bps;
            ev_last_copy = e;
            if (!ev_first_copy) ev_first_copy = e;
        }
        const int64_t f0 = (int64_t)(off / N);
        const int64_t f1 = std::min<int64_t>((int64_t)((off + m) / N), n_frames);
        Chunk c;
        c.end = off + m;
        c.fft = EvPair{p->ev(), p->ev()};
        c.scan = EvPair{p->ev(), p->ev()};
        c.e_hdr = p->ev();
        CK(cudaEventRecord(c.fft.a, p->st_fft));
        if (f1 > f0) {
            CK(launch_detect_fft(dc.L, fmt, iq_dev, f0 * N, p->d_window.p, p->d_tw_det.p, p->d_mag.p + f0 * N,
                                 f1 - f0, p->sm_count, p->st_fft));
            p->res.kernel_launches++;
        }
        CK(cudaEventRecord(c.fft.b, p->st_fft));
        CK(cudaStreamWaitEvent(p->st_scan, c.fft.b, 0));
        CK(cudaEventRecord(c.scan.a, p->st_scan));
        if (f1 > f0 && p->scan_mode == 1) {
            CK(launch_detect_scan_auto(dc, p->d_state.p, p->d_base.p, p->d_hist.p, p->d_mag.p + f0 * N, f1 - f0,
                                       p->d_gone, p->gone_cap, p->st_scan));
            p->res.kernel_launches++;
        } else if (f1 > f0 && p->scan_mode == 2) {
            if (scan_seg_range(p, f0, f1, c.fft.b)) return -1;
        } else if (f1 > f0) {
            if (scan_stream_range(p, f0, f1, c.fft.b)) return -1;
        }
        CK(cudaEventRecord(c.scan.b, p->st_scan));
        CK(cudaMemcpyAsync(p->h_hdr + p->chunks.size() * kHdrBytes, p->d_state.p, kHdrBytes, cudaMemcpyDeviceToHost,
                           p->st_scan));
        CK(cudaEventRecord(c.e_hdr, p->st_scan));
        p->res.d2h_bytes += kHdrBytes;
        p->chunks.push_back(c);
    }
    // ---- follow the detector: one wave of bursts per chunk
    const uint64_t B = p->cfg.feed_block > 0 ? (uint64_t)p->cfg.feed_block : 32768;
    size_t done = 0;
    DetState hs;
    memset(&hs, 0, kHdrBytes);
    // The sink: a second host thread turns every finished wave into demod_frame_t equivalents and RAW: text
    // (snprintf per frame, ~1 us each) while this thread does nothing but follow the detector and launch the next
    // wave -- with the formatting in line, the FIR of wave k+1 waited for the text of wave k-1.
    std::atomic<size_t> n_waves_pub{0};
    std::atomic<int> sink_stop{0}, sink_err{0};
    std::string sink_msg;
    std::thread sink([&]() {
        cudaSetDevice(p->dev);
        size_t wi = 0;
        for (;;) {
            while (wi >= n_waves_pub.load(std::memory_order_acquire)) {
                if (sink_stop.load(std::memory_order_acquire) && wi >= n_waves_pub.load(std::memory_order_acquire)) return;
                std::this_thread::yield();
            }
            if (assemble_wave(p, p->waves[wi])) { sink_msg = g_err; sink_err.store(1); return; }
            wi++;
        }
    });
    auto finish_sink = [&]() { sink_stop.store(1, std::memory_order_release); if (sink.joinable()) sink.join(); };
    struct SinkGuard { decltype(finish_sink) &f; ~SinkGuard() { f(); } } sink_guard{finish_sink};   // (every early return joins)
    int follow_rc = 0;
    const bool host_dbg = getenv("IR_CHUNK_DEBUG") != nullptr;
    const double t_enqueued = run_ms();
    const auto t_host0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count(); };
    for (size_t ci = 0; ci < p->chunks.size() && follow_rc == 0; ci++) {
        if (cudaEventSynchronize(p->chunks[ci].e_hdr) != cudaSuccess) { set_err("cudaEventSynchronize(header) failed"); follow_rc = -1; break; }
        const double t_seen = host_dbg ? host_ms() : 0.0;
        memcpy(&hs, p->h_hdr + ci * kHdrBytes, kHdrBytes);
        if (hs.overflow || hs.n_gone > p->gone_cap) { set_err("detector capacity exceeded (IR_MAX_ACTIVE or burst list)"); follow_rc = -1; break; }
        const bool last = ci + 1 == p->chunks.size();
        // a burst can be processed once every sample its extract reads is resident: its emit point
        // (end of the emulated feed call) must lie inside what has been copied and scanned
        size_t hi = done;
        while (hi < hs.n_gone) {
            if (!last) {
                const uint64_t calls = (p->h_gone[hi].stop + (uint64_t)N + B - 1) / B;
                if (std::min<uint64_t>(calls * B, n) > p->chunks[ci].end) break;
            }
            hi++;
        }
        if (hi > done) {
            if (launch_wave(p, done, hi, iq_dev, n, fmt)) { follow_rc = -1; break; }
            n_waves_pub.store(p->waves.size(), std::memory_order_release);
            done = hi;
        }
        if (host_dbg) fprintf(stderr, "host: chunk %zu header seen at %.3f ms (after the follow loop began), wave enqueued by %.3f ms\n", ci, t_seen, host_ms());
    }
    if (follow_rc != 0) { finish_sink(); return -1; }
    cudaEvent_t e_end = p->ev();
    if (!p->chunks.empty()) CK(cudaStreamWaitEvent(p->st_burst, p->chunks.back().e_hdr, 0));
    else CK(cudaStreamWaitEvent(p->st_burst, ev_begin, 0));
    for (size_t wi = p->waves.size() > (size_t)ir_pipeline::kDemodStreams ? p->waves.size() - ir_pipeline::kDemodStreams : 0;
         wi < p->waves.size(); wi++)
        CK(cudaStreamWaitEvent(p->st_burst, p->waves[wi].e_done, 0));     // the last wave of every demod stream
    CK(cudaEventRecord(e_end, p->st_burst));
    const double t_followed = run_ms();
    finish_sink();
    if (sink_err.load()) { set_err(sink_msg); return -1; }
    const double t_sunk = run_ms();
    if (host_iq) CK(cudaStreamSynchronize(p->st_copy));
    CK(cudaStreamSynchronize(p->st_burst));
    if (host_dbg)
        fprintf(stderr, "host: everything enqueued at %.3f ms after the call, last wave launched at %.3f, sink thread done at %.3f, "
                        "device idle at %.3f\n", t_enqueued, t_followed, t_sunk, run_ms());
    memset(p->scan_stats, 0, sizeof(p->scan_stats));
    if (p->scan_mode == 0)
        CK(cudaMemcpy(p->scan_stats, p->d_ctl.p->stats, sizeof(p->scan_stats), cudaMemcpyDeviceToHost));
    else if (p->scan_mode == 2) {
        SegCtl hc;
        CK(cudaMemcpy(&hc, p->seg.ctl, sizeof(SegCtl), cudaMemcpyDeviceToHost));
        memcpy(p->scan_stats, hc.stats, sizeof(hc.stats));
        p->scan_stats[6] = ((uint64_t)hc.reason & 0xffu) | (hc.stats[6] << 8);     // + segments the generic walker took
        if (getenv("IR_SCAN_DEBUG"))
            fprintf(stderr, "seg scan: chunks kept %llu bailed %llu (last reason %d) rounds %llu event frames (all rounds) %llu snapshots %llu quiet frames %llu bitmap rebuilds %llu generic segment walks %llu\n",
                    (unsigned long long)hc.stats[0], (unsigned long long)hc.stats[1], hc.reason, (unsigned long long)hc.stats[2],
                    (unsigned long long)hc.stats[3], (unsigned long long)hc.stats[4], (unsigned long long)hc.stats[5],
                    (unsigned long long)hc.stats[7], (unsigned long long)hc.stats[6]);
#ifdef IR_SEG_TIMING
        if (getenv("IR_SCAN_DEBUG")) seg_timing_dump();
#endif
    }
    if (p->scan_dbg && p->scan_mode == 0 && !p->scan_dbg_ev.empty()) {
        double t[4] = {0, 0, 0, 0};
        for (size_t i = 0; i + 4 < p->scan_dbg_ev.size() + 1; i += 5)
            for (int k = 0; k < 4; k++) t[k] += span(p->scan_dbg_ev[i + k], p->scan_dbg_ev[i + k + 1]);
        fprintf(stderr, "stream scan launches: %zu; ms in classify+snapshot %.3f, (-) %.3f, state machine %.3f, fallback (no-op) %.3f\n",
                p->scan_dbg_ev.size() / 5, t[0], t[1], t[2], t[3]);
    }
    if (getenv("IR_SCAN_DEBUG") && p->scan_mode == 0) {
        int reason = 0;
        cudaMemcpy(&reason, &p->d_ctl.p->reason, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "stream scan: launches kept %llu bailed %llu (last reason %d at frame %llu) commands %llu events %llu exact words %llu waits %llu\n",
                (unsigned long long)p->scan_stats[0], (unsigned long long)p->scan_stats[1], reason,
                (unsigned long long)p->scan_stats[6], (unsigned long long)p->scan_stats[2],
                (unsigned long long)p->scan_stats[3], (unsigned long long)p->scan_stats[4],
                (unsigned long long)p->scan_stats[5]);
        fprintf(stderr, "stream scan: kernel %.3f ms; leader Mcycles (IR_SCAN_TIMING builds): frames %.2f list+prefetch %.2f worker wait %.2f hyst+peaks %.2f delete %.2f create+reload %.2f\n",
                p->scan_stats[7] * 1e-6, p->scan_stats[8] * 1e-6, p->scan_stats[9] * 1e-6, p->scan_stats[10] * 1e-6,
                p->scan_stats[11] * 1e-6, p->scan_stats[12] * 1e-6, p->scan_stats[13] * 1e-6);
        fprintf(stderr, "stream scan: deadline-only events handled in the fast loop %llu\n", (unsigned long long)p->scan_stats[16]);
        fprintf(stderr, "stream scan: ring refill+wait %.2f Mcycles, loop trips %llu; pair path: loads+logic %.2f votes %.2f commit %.2f\n",
                p->scan_stats[14] * 1e-6, (unsigned long long)p->scan_stats[15], p->scan_stats[17] * 1e-6,
                p->scan_stats[18] * 1e-6, p->scan_stats[19] * 1e-6);
    }
    if (getenv("IR_SCAN_DEBUG") && p->scan_mode != 0) {
        fprintf(stderr, "scan cycles leader: p1 %llu waitA %llu p2 %llu waitB %llu p3 %llu waitC %llu batches %llu qbatches %llu\n",
                hs.dbg[0], hs.dbg[1], hs.dbg[2], hs.dbg[3], hs.dbg[4], hs.dbg[5], hs.dbg[6], hs.dbg[7]);
        fprintf(stderr, "scan leader p2 (type F): precheck %llu old-replay %llu | per round: eligible %llu peaks %llu fetch %llu greedy %llu new-replay %llu | apply %llu | frame replay %llu | rounds %llu planned %llu replayed %llu\n",
                hs.dbg[8], hs.dbg[9], hs.dbg[10], hs.dbg[11], hs.dbg[12], hs.dbg[13], hs.dbg[14], hs.dbg[15], hs.dbg[16],
                hs.dbg[17], hs.dbg[18], hs.dbg[19]);
        fprintf(stderr, "scan owner(rank3) p1 split: setup %llu chunks %llu ship %llu\n", hs.dbg[20], hs.dbg[21], hs.dbg[22]);
    }
    // ---- timings: CUDA events on the launching streams
    p->res.ms_detect_fft = 0; p->res.ms_detect_scan = 0;
    p->res.ms_downmix_fir = 0; p->res.ms_downmix_chain = 0; p->res.ms_demod = 0;
    for (auto &c : p->chunks) {
        p->res.ms_detect_fft += span(c.fft.a, c.fft.b);
        p->res.ms_detect_scan += span(c.scan.a, c.scan.b);
    }
    for (auto &w : p->waves) {
        p->res.ms_downmix_fir += span(w.e0, w.e1);
        p->res.ms_downmix_chain += span(w.e1a, w.e1b);
        p->res.ms_demod += span(w.e2, w.e3);
    }
    p->res.ms_total = span(ev_begin, e_end);
    if (getenv("IR_CHUNK_DEBUG")) {
        for (size_t ci = 0; ci < p->chunks.size(); ci++) {
            const Chunk &c = p->chunks[ci];
            fprintf(stderr, "chunk %2zu end %10zu: fft %.3f..%.3f ms, scan %.3f..%.3f ms, header on host %.3f ms\n", ci, c.end,
                    span(ev_begin, c.fft.a), span(ev_begin, c.fft.b), span(ev_begin, c.scan.a), span(ev_begin, c.scan.b),
                    span(ev_begin, c.e_hdr));
        }
        for (size_t wi = 0; wi < p->waves.size(); wi++) {
            const Wave &w = p->waves[wi];
            fprintf(stderr, "wave %2zu (%4zu bursts): fir %.3f..%.3f chain ..%.3f demod %.3f..%.3f results %.3f ms\n", wi, w.nb,
                    span(ev_begin, w.e0), span(ev_begin, w.e1), span(ev_begin, w.e1b), span(ev_begin, w.e2),
                    span(ev_begin, w.e3), span(ev_begin, w.e_done));
        }
    }
    if (getenv("IR_SCAN_DEBUG") && ev_last_copy)
        fprintf(stderr, "run_host: first copy done at %.3f ms, last copy done at %.3f ms, end of device work at %.3f ms (%zu chunks)\n",
                span(ev_begin, ev_first_copy), span(ev_begin, ev_last_copy), p->res.ms_total, p->chunks.size());
    p->res.alg_bytes = p->alg + p->alg_sink;
    p->res.n_bursts = p->bursts.size(); p->res.bursts = p->bursts.data();
    p->res.n_frames = p->frames.size(); p->res.frames = p->frames.data();
    p->res.bits = p->bits.data(); p->res.llr = p->llr.data(); p->res.n_bits_total = p->bits.size();
    p->last_iq = iq_dev; p->last_n = n; p->last_fmt = fmt;
    return 0;
}

extern "C" int ir_pipeline_run_host(ir_pipeline_t *p, const void *iq, size_t n_samples, int fmt) {
    if (!iq) { set_err("null input"); return -1; }
    return run_common(p, iq, nullptr, n_samples, fmt);
}

extern "C" int ir_pipeline_run_device(ir_pipeline_t *p, const void *iq_dev, size_t n_samples, int fmt) {
    if (!iq_dev) { set_err("null input"); return -1; }
    return run_common(p, nullptr, iq_dev, n_samples, fmt);
}

extern "C" int ir_pipeline_results(ir_pipeline_t *p, ir_results_t *out) {
    if (!p || !out) return -1;
    *out = p->res;
    return 0;
}

// End offsets of the pieces a block of n samples is processed in: full chunks, then the last stretch
// in halves (16, 8, 4, 2, 1, 1 Mi samples for the default 32 Mi): what runs after the last copy / the last
// state-machine launch -- one wave of FIR + chain + demod, the result copies, the RAW text -- shrinks
// with the last piece.  Every piece but the last is a whole number of detector frames.
extern "C" long ir_plan_chunks(size_t n, size_t chunk, size_t fft_size, size_t *ends, size_t cap) {
    if (!ends || fft_size == 0 || chunk == 0) return -1;
    const size_t N = fft_size;
    chunk = std::max<size_t>(chunk / N, 1) * N;
    const size_t min_piece = std::max<size_t>(((size_t)1 << 20) / N, 1) * N;
    size_t off = 0, k = 0;
    while (off < n) {
        size_t m = std::min(chunk, n - off);
        if (n - off <= chunk)                                  // inside the last chunk
            while (m > min_piece && m / 2 >= min_piece && (n - off) - m < m) m = std::max<size_t>(m / 2 / N, 1) * N;
        off += m;
        if (k >= cap) return -1;
        ends[k++] = off;
    }
    return (long)k;
}

extern "C" int ir_pipeline_set_origin(ir_pipeline_t *p, uint64_t sample_origin) {
    if (!p) { set_err("null pipeline"); return -1; }
    p->sample_origin = sample_origin;
    return 0;
}

extern "C" int ir_pipeline_set_start_time(ir_pipeline_t *p, uint64_t start_time_ns) {
    if (!p) { set_err("null pipeline"); return -1; }
    p->cfg.start_time_ns = start_time_ns;
    return 0;
}

extern "C" int ir_pipeline_scan_stats(ir_pipeline_t *p, uint64_t *out, int n) {
    if (!p || !out) return -1;
    for (int i = 0; i < n && i < 8; i++) out[i] = p->scan_stats[i];
    return p->scan_mode == 0 ? 1 : (p->scan_mode == 2 ? 2 : 0);
}

extern "C" int ir_pipeline_copy_mag(ir_pipeline_t *p, size_t frame0, size_t n_frames, float *dst) {
    if (!p || !dst) return -1;
    if ((int64_t)(frame0 + n_frames) > p->n_frames_last) { set_err("frame range out of bounds"); return -1; }
    CK(cudaSetDevice(p->dev));
    CK(cudaMemcpy(dst, p->d_mag.p + frame0 * p->dc.N, n_frames * p->dc.N * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int ir_pipeline_copy_frame_samples(ir_pipeline_t *p, size_t bi, float *dst, size_t cap) {
    if (!p || !dst || bi >= p->bursts.size()) return -1;
    const size_t nfl = (size_t)p->bursts[bi].frame_len;
    if (p->bursts[bi].downmix_status != 0 || nfl > cap) return -1;
    CK(cudaSetDevice(p->dev));
    CK(cudaMemcpy(dst, p->frame_ptr[bi], nfl * sizeof(float2), cudaMemcpyDeviceToHost));
    return (int)nfl;
}

extern "C" int ir_pipeline_copy_decimated(ir_pipeline_t *p, size_t bi, float *dst, size_t cap) {
    if (!p || !dst || bi >= p->bursts.size()) return -1;
    const size_t dl = (size_t)p->h_bp[bi].dec_len;
    if (dl == 0 || dl > cap) return -1;
    CK(cudaSetDevice(p->dev));
    CK(cudaMemcpy(dst, p->dec_ptr[bi], dl * sizeof(float2), cudaMemcpyDeviceToHost));
    return (int)dl;
}

// Gather the burst's IQ the way the reference's ringbuf_extract would have delivered it.
template <int FMT>
__global__ void k_gather_burst(const void *__restrict__ iq, int64_t n_total, uint64_t ring, BurstParam P,
                               float2 *__restrict__ dst, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = P.start + i;
        if (q >= P.emit_count) q -= (int64_t)ring;
        dst[i] = (q >= 0 && q < n_total) ? load_sample<FMT>(iq, q) : make_float2(0.0f, 0.0f);
    }
}

extern "C" int ir_pipeline_copy_burst_samples(ir_pipeline_t *p, size_t bi, float *dst, size_t cap) {
    if (!p || !dst || bi >= p->bursts.size() || !p->last_iq) return -1;
    const size_t ns = (size_t)p->bursts[bi].num_samples;
    if (ns == 0 || ns > cap) return -1;
    CK(cudaSetDevice(p->dev));
    float2 *tmp;
    CK(cudaMalloc(&tmp, ns * sizeof(float2)));
    BurstParam P = p->h_bp[bi];
    const int blocks = (int)std::min<size_t>((ns + 255) / 256, 4096);
    if (p->last_fmt == IR_FMT_CF32) k_gather_burst<IR_FMT_CF32><<<blocks, 256>>>(p->last_iq, (int64_t)p->last_n, p->dc.ringbuf_size, P, tmp, (int64_t)ns);
    else if (p->last_fmt == IR_FMT_CI16) k_gather_burst<IR_FMT_CI16><<<blocks, 256>>>(p->last_iq, (int64_t)p->last_n, p->dc.ringbuf_size, P, tmp, (int64_t)ns);
    else k_gather_burst<IR_FMT_CI8><<<blocks, 256>>>(p->last_iq, (int64_t)p->last_n, p->dc.ringbuf_size, P, tmp, (int64_t)ns);
    cudaError_t e = cudaMemcpy(dst, tmp, ns * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return -1; }
    return (int)ns;
}

// frame_output_print (frame_output.c:160-199)
extern "C" int ir_format_raw(char *dst, size_t cap, const char *file_info, uint64_t t0,
                             const ir_frame_t *f, const uint8_t *bits) {
    if (!dst || !f || cap < 64) return -1;
    int k = snprintf(dst, cap, "RAW: %s", file_info ? file_info : "");
    if (k < 0) return -1;
    const size_t pos = (size_t)k < cap ? (size_t)k : cap - 1;
    if (cap - pos < 8) { dst[pos] = 0; return (int)pos; }
    const int r = format_raw_rest(dst + pos, cap - pos, t0, f, bits);
    return r < 0 ? -1 : (int)pos + r;
}

extern "C" long ir_pipeline_format_raw_all(ir_pipeline_t *p, const char *file_info, uint64_t t0, char *dst,
                                           size_t cap) {
    if (!p) return -1;
    const size_t head = 128 + (file_info ? strlen(file_info) : 0);   // fixed-width fields < 100 chars
    if (!dst) return (long)(p->frames.size() * head + p->bits.size() + 64);
    if (p->frames.empty()) return 0;
    if (t0 == 0) t0 = (p->frames[0].timestamp / 1000000000ULL) * 1000000000ULL;
    size_t pos = 0;
    if (t0 == p->raw_t0 && p->raw_off.size() == p->frames.size()) {
        // the lines were formatted during the run: "RAW: " + file_info + the cached remainder
        const size_t fl = file_info ? strlen(file_info) : 0;
        for (size_t i = 0; i < p->frames.size(); i++) {
            const size_t a = p->raw_off[i], b = i + 1 < p->raw_off.size() ? p->raw_off[i + 1] : p->raw_rest.size();
            if (pos + 5 + fl + (b - a) + 1 > cap) { set_err("ir_pipeline_format_raw_all: buffer too small"); return -1; }
            memcpy(dst + pos, "RAW: ", 5); pos += 5;
            if (fl) { memcpy(dst + pos, file_info, fl); pos += fl; }
            memcpy(dst + pos, p->raw_rest.data() + a, b - a); pos += b - a;
        }
        dst[pos < cap ? pos : cap - 1] = 0;
        return (long)pos;
    }
    for (const ir_frame_t &f : p->frames) {
        if (pos + head + (size_t)f.n_bits + 2 > cap) { set_err("ir_pipeline_format_raw_all: buffer too small"); return -1; }
        int n = ir_format_raw(dst + pos, cap - pos, file_info, t0, &f, p->bits.data() + f.bits_offset);
        if (n < 0) return -1;
        pos += (size_t)n;
    }
    return (long)pos;
}

// ---- frame classification: frame_decode() + ida_decode() of main.c:320-350, batched
extern "C" long ir_pipeline_classify(ir_pipeline_t *p, ir_frame_class_t *out, size_t cap) {
    if (!p || (!out && cap)) { set_err("ir_pipeline_classify: null argument"); return -1; }
    const size_t n = p->frames.size();
    if (n > cap) { set_err("ir_pipeline_classify: output array too small"); return -1; }
    if (n == 0) return 0;
    CK(cudaSetDevice(p->dev));
    const void *d_tab = nullptr;
    CK(classify_tables(p->dev, &d_tab));
    if (p->d_frame_src.ensure(n) || p->d_class.ensure(n)) return -1;
    CK(cudaMemcpyAsync(p->d_frame_src.p, p->frame_src.data(), n * sizeof(FrameSrc), cudaMemcpyHostToDevice, p->st_burst));
    for (auto &e : p->ev_cls)
        if (!e) CK(cudaEventCreate(&e));
    cudaEvent_t e0 = p->ev_cls[0], e1 = p->ev_cls[1];
    CK(cudaEventRecord(e0, p->st_burst));
    CK(launch_classify(d_tab, p->d_frame_src.p, (int)n, p->d_class.p, p->st_burst));
    CK(cudaEventRecord(e1, p->st_burst));
    CK(cudaMemcpyAsync(out, p->d_class.p, n * sizeof(ir_frame_class_t), cudaMemcpyDeviceToHost, p->st_burst));
    CK(cudaStreamSynchronize(p->st_burst));
    CK(cudaEventElapsedTime(&p->ms_classify, e0, e1));
    p->res.kernel_launches++;
    return (long)n;
}

extern "C" float ir_pipeline_last_classify_ms(ir_pipeline_t *p) { return p ? p->ms_classify : -1.0f; }

extern "C" int ir_classify_frames(int device, const ir_frame_t *frames, size_t n_frames, const uint8_t *bits,
                                  const float *llr, size_t n_bits_total, ir_frame_class_t *out) {
    if (n_frames == 0) return 0;
    if (!frames || !bits || !out) { set_err("ir_classify_frames: null argument"); return -1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        set_err(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
        return -1;
    }
    if (device < 0 || device >= ndev) { set_err("device ordinal out of range"); return -1; }
    CK(cudaSetDevice(device));
    const void *d_tab = nullptr;
    CK(classify_tables(device, &d_tab));
    // the calling thread's buffers on this device, kept between calls (the reference-named frame_decode() /
    // ida_decode() come here once per frame: four allocations and four frees each time cost more than the kernel);
    // a call with more than a few MB of bits gives them back
    struct ClassifyBufs {
        DevBuf<uint8_t> bits;
        DevBuf<float> llr;
        DevBuf<FrameSrc> src;
        DevBuf<ir_frame_class_t> out;
        int device = -1;
        void drop() { if (device >= 0 && cudaSetDevice(device) == cudaSuccess) { bits.release(); llr.release(); src.release(); out.release(); } device = -1; }
        ~ClassifyBufs() { drop(); }
    };
    static thread_local ClassifyBufs t_bufs;
    if (t_bufs.device != device) {
        int cur = device;
        t_bufs.drop();
        CK(cudaSetDevice(cur));
        t_bufs.device = cur;
    }
    DevBuf<uint8_t> &d_bits = t_bufs.bits;
    DevBuf<float> &d_llr = t_bufs.llr;
    DevBuf<FrameSrc> &d_src = t_bufs.src;
    DevBuf<ir_frame_class_t> &d_out = t_bufs.out;
    struct Guard { ClassifyBufs &b; bool big; ~Guard() { if (big) b.drop(); } } guard{t_bufs, n_bits_total > (size_t)(4u << 20)};
    std::vector<FrameSrc> src(n_frames);
    if (d_bits.ensure(n_bits_total + 1) || (llr && d_llr.ensure(n_bits_total + 1)) || d_src.ensure(n_frames) || d_out.ensure(n_frames))
        return -1;
    for (size_t i = 0; i < n_frames; i++) {
        const ir_frame_t &f = frames[i];
        if (f.n_bits < 0 || (size_t)f.bits_offset + (size_t)f.n_bits > n_bits_total) { set_err("ir_classify_frames: frame outside the bit array"); return -1; }
        src[i] = FrameSrc{d_bits.p + f.bits_offset, llr ? d_llr.p + f.bits_offset : nullptr, f.n_bits, f.direction};
    }
    CK(cudaMemcpy(d_bits.p, bits, n_bits_total, cudaMemcpyHostToDevice));
    if (llr) CK(cudaMemcpy(d_llr.p, llr, n_bits_total * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_src.p, src.data(), n_frames * sizeof(FrameSrc), cudaMemcpyHostToDevice));
    CK(launch_classify(d_tab, d_src.p, (int)n_frames, d_out.p, 0));
    CK(cudaMemcpy(out, d_out.p, n_frames * sizeof(ir_frame_class_t), cudaMemcpyDeviceToHost));
    return 0;
}

// `--parsed` output of the whole run (main.c:328-331): the IDA line where ida_decode() accepted the frame, the RAW line otherwise
extern "C" long ir_pipeline_format_parsed_all(ir_pipeline_t *p, const char *file_info, uint64_t t0, const ir_frame_class_t *cls,
                                              size_t n_cls, char *dst, size_t cap) {
    if (!p) return -1;
    const size_t n = p->frames.size();
    const size_t head = 512 + (file_info ? strlen(file_info) : 0);   // an IDA line is < 450 characters
    if (!dst) return (long)(n * head + p->bits.size() + 64);
    if (n == 0) return 0;
    std::vector<ir_frame_class_t> own;
    if (!cls) {
        own.resize(n);
        if (ir_pipeline_classify(p, own.data(), n) < 0) return -1;
        cls = own.data();
        n_cls = n;
    }
    if (n_cls != n) { set_err("ir_pipeline_format_parsed_all: class array does not match the last run"); return -1; }
    if (t0 == 0) t0 = (p->frames[0].timestamp / 1000000000ULL) * 1000000000ULL;
    size_t pos = 0;
    for (size_t i = 0; i < n; i++) {
        const ir_frame_t &f = p->frames[i];
        if (pos + head + (size_t)f.n_bits + 2 > cap) { set_err("ir_pipeline_format_parsed_all: buffer too small"); return -1; }
        const int k = cls[i].ida_ok ? ir_format_ida(dst + pos, cap - pos, t0, &f, &cls[i])
                                    : ir_format_raw(dst + pos, cap - pos, file_info, t0, &f, p->bits.data() + f.bits_offset);
        if (k < 0) { set_err("ir_pipeline_format_parsed_all: formatting failed"); return -1; }
        pos += (size_t)k;
    }
    return (long)pos;
}

extern "C" void *ir_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void ir_host_free(void *p) { if (p) cudaFreeHost(p); }
