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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/iridium_b200.h"
#include "ir_device.cuh"
#include "ir_internal.h"

using namespace ir;

static thread_local std::string g_err;
static void set_err(const std::string &s) { g_err = s; }
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

}  // namespace

struct ir_pipeline {
    ir_config_t cfg;
    DetConfig dc;
    HostTables tab;
    int dev = 0, sm_count = 148;
    int dec = 40;
    cudaStream_t st_copy = nullptr, st_fft = nullptr, st_scan = nullptr;
    // constants
    DevBuf<float> d_window;
    DevBuf<float2> d_tw_det, d_tw12, d_tw11, d_sync_dl, d_sync_ul;
    // stream
    DevBuf<unsigned char> d_iq;
    DevBuf<float> d_mag, d_base, d_hist;
    DevBuf<DetState> d_state;
    DevBuf<GoneBurst> d_gone;
    // bursts
    DevBuf<BurstParam> d_bp;
    DevBuf<int> d_tile_start;
    DevBuf<float2> d_dec, d_scrA, d_scrB, d_frames;
    DevBuf<ChainOut> d_co;
    DevBuf<DemodOut> d_do;
    DevBuf<unsigned char> d_bits;
    DevBuf<float> d_llr;
    std::unordered_map<int, RotTable> rot;
    // last run (host)
    const void *last_iq = nullptr;
    size_t last_n = 0;
    int last_fmt = 0;
    std::vector<GoneBurst> h_gone;
    std::vector<BurstParam> h_bp;
    std::vector<ChainOut> h_co;
    std::vector<DemodOut> h_do;
    std::vector<ir_burst_t> bursts;
    std::vector<ir_frame_t> frames;
    std::vector<uint8_t> bits;
    std::vector<float> llr;
    std::vector<uint8_t> h_bits_raw;
    std::vector<float> h_llr_raw;
    ir_results_t res;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    uint64_t start_time_ns = 0;
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
    if (cudaStreamCreateWithFlags(&p->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->st_fft, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->st_scan, cudaStreamNonBlocking) != cudaSuccess)
        return fail("stream creation failed");
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
    memset(&p->res, 0, sizeof(p->res));
    return p;
}

extern "C" void ir_pipeline_destroy(ir_pipeline_t *p) {
    if (!p) return;
    cudaSetDevice(p->dev);
    cudaDeviceSynchronize();
    for (auto &kv : p->rot) if (kv.second.d) cudaFree(kv.second.d);
    p->d_window.release(); p->d_tw_det.release(); p->d_tw12.release(); p->d_tw11.release();
    p->d_sync_dl.release(); p->d_sync_ul.release(); p->d_iq.release(); p->d_mag.release();
    p->d_base.release(); p->d_hist.release(); p->d_state.release(); p->d_gone.release();
    p->d_bp.release(); p->d_tile_start.release(); p->d_dec.release(); p->d_scrA.release();
    p->d_scrB.release(); p->d_frames.release(); p->d_co.release(); p->d_do.release();
    p->d_bits.release(); p->d_llr.release();
    for (auto e : p->ev_pool) cudaEventDestroy(e);
    if (p->st_copy) cudaStreamDestroy(p->st_copy);
    if (p->st_fft) cudaStreamDestroy(p->st_fft);
    if (p->st_scan) cudaStreamDestroy(p->st_scan);
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
// Everything after the detector: host bookkeeping + FIR + chain + demod + result assembly.
static int finish_run(ir_pipeline *p, const void *iq_dev, size_t n, int fmt, cudaEvent_t ev_begin,
                      std::vector<EvPair> &ev_fft, std::vector<EvPair> &ev_scan) {
    const DetConfig &dc = p->dc;
    cudaStream_t st = p->st_scan;
    DetState hs;                       // only the header is needed
    CK(cudaMemcpyAsync(&hs, p->d_state.p, offsetof(DetState, act), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    p->res.d2h_bytes += offsetof(DetState, act);
    if (hs.overflow) { set_err("detector capacity exceeded (IR_MAX_ACTIVE or burst list)"); return -1; }
    if (getenv("IR_SCAN_DEBUG")) {
        fprintf(stderr, "scan cycles leader: p1 %llu waitA %llu p2 %llu waitB %llu p3 %llu waitC %llu batches %llu qbatches %llu\n",
                hs.dbg[0], hs.dbg[1], hs.dbg[2], hs.dbg[3], hs.dbg[4], hs.dbg[5], hs.dbg[6], hs.dbg[7]);
        fprintf(stderr, "scan leader p2 split: search %llu preflags %llu deletion %llu creation %llu squelch/end %llu | event frames %llu creates %llu deletes %llu\n",
                hs.dbg[8], hs.dbg[9], hs.dbg[10], hs.dbg[11], hs.dbg[12], hs.dbg[13], hs.dbg[14], hs.dbg[15]);
        fprintf(stderr, "scan owner(rank3) p1 split: issue %llu oldloads %llu wait+sync %llu lds+sync %llu compute %llu\n",
                hs.dbg[16], hs.dbg[17], hs.dbg[18], hs.dbg[19], hs.dbg[20]);
    }
    const size_t nb = hs.n_gone;
    p->h_gone.resize(nb);
    if (nb) {
        CK(cudaMemcpyAsync(p->h_gone.data(), p->d_gone.p, nb * sizeof(GoneBurst), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        p->res.d2h_bytes += nb * sizeof(GoneBurst);
    }
    // ---- burst_data_t equivalents (burst_detect.c:703-742) and decimation geometry
    const uint64_t B = p->cfg.feed_block > 0 ? (uint64_t)p->cfg.feed_block : 32768;
    const uint64_t R = dc.ringbuf_size;
    const int fs = dc.sample_rate, N = dc.N;
    p->bursts.assign(nb, ir_burst_t{});
    p->h_bp.assign(nb, BurstParam{});
    std::vector<int> tile_start(nb + 1, 0);
    std::vector<int> need_bins;
    int64_t dec_total = 0;
    int n_tiles = 0;
    uint64_t alg = (uint64_t)n * fmt_bytes(fmt);
    for (size_t i = 0; i < nb; i++) {
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
        bp.dec_off = dec_total;
        bp.tile0 = n_tiles;
        tile_start[i] = n_tiles;
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
            n_tiles += (dlen + IR_FIR_TILE - 1) / IR_FIR_TILE;
            auto it = p->rot.find(g.center_bin);
            if (it == p->rot.end() || it->second.len < nn + IR_ROT_G) need_bins.push_back((int)i);
            alg += 8ull * (uint64_t)nn + 8ull * (uint64_t)dlen;
        } else {
            bp.dec_len = 0;        // chain reports status 2
        }
    }
    tile_start[nb] = n_tiles;
    // ---- NCO checkpoint tables for bins not cached yet (or cached too short)
    if (!need_bins.empty()) {
        std::unordered_map<int, int> want;     // bin -> samples
        for (int i : need_bins) {
            int bin = p->h_gone[i].center_bin;
            int len = ((p->h_bp[i].n + IR_ROT_G + 65535) / 65536) * 65536;
            auto w = want.find(bin);
            if (w == want.end() || w->second < len) want[bin] = len;
        }
        std::vector<float2> incr;
        std::vector<float2 *> ptrs;
        std::vector<int> lens;
        for (auto &kv : want) {
            RotTable &rt = p->rot[kv.first];
            if (rt.d) cudaFree(rt.d);
            rt.len = kv.second;
            CK(cudaMalloc(&rt.d, sizeof(float2) * (size_t)(rt.len / IR_ROT_G + 1)));
            const float rel = (kv.first - N / 2) / (float)N;
            const float ph = -2.0f * (float)M_PI * rel;
            float sn, cs;
            sincosf(ph, &sn, &cs);
            rt.incr = make_float2(cs, sn);
            incr.push_back(rt.incr); ptrs.push_back(rt.d); lens.push_back(rt.len);
        }
        float2 *d_incr; float2 **d_ptrs; int *d_lens;
        const size_t k = incr.size();
        CK(cudaMalloc(&d_incr, k * sizeof(float2)));
        CK(cudaMalloc(&d_ptrs, k * sizeof(float2 *)));
        CK(cudaMalloc(&d_lens, k * sizeof(int)));
        CK(cudaMemcpyAsync(d_incr, incr.data(), k * sizeof(float2), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_ptrs, ptrs.data(), k * sizeof(float2 *), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_lens, lens.data(), k * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(launch_rot_tables(d_incr, d_ptrs, d_lens, (int)k, st));
        p->res.kernel_launches++;
        CK(cudaStreamSynchronize(st));
        cudaFree(d_incr); cudaFree(d_ptrs); cudaFree(d_lens);
    }
    for (size_t i = 0; i < nb; i++) {
        auto it = p->rot.find(p->h_gone[i].center_bin);
        p->h_bp[i].rot_table = it != p->rot.end() ? it->second.d : nullptr;
    }
    // ---- device work for the bursts
    cudaEvent_t e_fir0 = p->ev(), e_fir1 = p->ev(), e_ch1 = p->ev(), e_dm1 = p->ev();
    p->h_co.assign(nb, ChainOut{});
    p->h_do.assign(nb, DemodOut{});
    if (nb) {
        if (p->d_bp.ensure(nb) || p->d_tile_start.ensure(nb + 1) || p->d_dec.ensure((size_t)dec_total + 16) ||
            p->d_scrA.ensure((size_t)dec_total + 16) || p->d_scrB.ensure((size_t)dec_total + 16) ||
            p->d_frames.ensure(nb * (size_t)IR_MAX_FRAME) || p->d_co.ensure(nb) || p->d_do.ensure(nb) ||
            p->d_bits.ensure(nb * 2 * (size_t)IR_MAX_SYMS) || p->d_llr.ensure(nb * 2 * (size_t)IR_MAX_SYMS))
            return -1;
        CK(cudaMemcpyAsync(p->d_bp.p, p->h_bp.data(), nb * sizeof(BurstParam), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(p->d_tile_start.p, tile_start.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        p->res.h2d_bytes += nb * sizeof(BurstParam) + (nb + 1) * sizeof(int);
        CK(cudaEventRecord(e_fir0, st));
        CK(launch_fir(fmt, p->dec, iq_dev, (int64_t)n, R, p->d_bp.p, p->d_tile_start.p, (int)nb, n_tiles, p->d_dec.p, st));
        CK(cudaEventRecord(e_fir1, st));
        CK(launch_chain(p->d_bp.p, (int)nb, p->d_dec.p, p->d_scrA.p, p->d_scrB.p, p->d_tw12.p, p->d_tw11.p,
                        p->d_sync_dl.p, p->d_sync_ul.p, p->d_co.p, p->d_frames.p, st));
        CK(cudaEventRecord(e_ch1, st));
        CK(launch_demod(p->d_co.p, (int)nb, p->d_frames.p, p->cfg.use_gardner, p->d_do.p, p->d_bits.p, p->d_llr.p, st));
        CK(cudaEventRecord(e_dm1, st));
        p->res.kernel_launches += (n_tiles > 0 ? 1 : 0) + 2;
        p->h_bits_raw.resize(nb * 2 * (size_t)IR_MAX_SYMS);
        p->h_llr_raw.resize(nb * 2 * (size_t)IR_MAX_SYMS);
        CK(cudaMemcpyAsync(p->h_co.data(), p->d_co.p, nb * sizeof(ChainOut), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(p->h_do.data(), p->d_do.p, nb * sizeof(DemodOut), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(p->h_bits_raw.data(), p->d_bits.p, p->h_bits_raw.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(p->h_llr_raw.data(), p->d_llr.p, p->h_llr_raw.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        p->res.d2h_bytes += nb * (sizeof(ChainOut) + sizeof(DemodOut)) + p->h_bits_raw.size() + p->h_llr_raw.size() * sizeof(float);
    } else {
        CK(cudaEventRecord(e_fir0, st)); CK(cudaEventRecord(e_fir1, st));
        CK(cudaEventRecord(e_ch1, st)); CK(cudaEventRecord(e_dm1, st));
    }
    cudaEvent_t e_end = p->ev();
    CK(cudaEventRecord(e_end, st));
    CK(cudaStreamSynchronize(st));
    // ---- assemble demod_frame_t equivalents (qpsk_demod.c:505-527) on the host, in double
    p->frames.clear(); p->bits.clear(); p->llr.clear();
    const uint64_t delay_ns = (uint64_t)((IR_INPUT_NTAPS / 2) * 1000000000ULL / fs);      // burst_downmix.c:431-433
    for (size_t i = 0; i < nb; i++) {
        ir_burst_t &ob = p->bursts[i];
        const ChainOut &c = p->h_co[i];
        const DemodOut &d = p->h_do[i];
        ob.downmix_status = p->h_bp[i].dec_len >= 100 ? c.status : (ob.num_samples < 100 ? 1 : 2);
        ob.demod_ok = 0;
        ob.center_offset = c.center_offset; ob.dm_start = c.start; ob.uw_start = c.uw_start;
        ob.frame_len = c.frame_len; ob.uw_start_frac = c.uw_corr; ob.dm_direction = c.direction;
        if (ob.downmix_status != 0) continue;
        alg += 8ull * (uint64_t)c.frame_len;
        if (!d.ok) continue;
        ob.demod_ok = 1;
        ir_frame_t f;
        f.id = ob.id;
        uint64_t ts = p->start_time_ns + (uint64_t)((double)ob.start / fs * 1e9);          // :659-660
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
        const uint8_t *br = p->h_bits_raw.data() + i * 2 * (size_t)IR_MAX_SYMS;
        const float *lr = p->h_llr_raw.data() + i * 2 * (size_t)IR_MAX_SYMS;
        p->bits.insert(p->bits.end(), br, br + f.n_bits);
        p->llr.insert(p->llr.end(), lr, lr + f.n_bits);
        p->frames.push_back(f);
        alg += 8ull * (uint64_t)c.frame_len + (uint64_t)f.n_bits;
    }
    // ---- timings
    auto span = [&](cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; };
    p->res.ms_detect_fft = 0; p->res.ms_detect_scan = 0;
    for (auto &e : ev_fft) p->res.ms_detect_fft += span(e.a, e.b);
    for (auto &e : ev_scan) p->res.ms_detect_scan += span(e.a, e.b);
    p->res.ms_downmix_fir = span(e_fir0, e_fir1);
    p->res.ms_downmix_chain = span(e_fir1, e_ch1);
    p->res.ms_demod = span(e_ch1, e_dm1);
    p->res.ms_total = span(ev_begin, e_end);
    p->res.alg_bytes = alg;
    p->res.n_bursts = p->bursts.size(); p->res.bursts = p->bursts.data();
    p->res.n_frames = p->frames.size(); p->res.frames = p->frames.data();
    p->res.bits = p->bits.data(); p->res.llr = p->llr.data(); p->res.n_bits_total = p->bits.size();
    p->last_iq = iq_dev; p->last_n = n; p->last_fmt = fmt;
    return 0;
}

static int run_common(ir_pipeline *p, const void *host_iq, const void *dev_iq, size_t n, int fmt) {
    if (!p) { set_err("null pipeline"); return -1; }
    if (fmt < 0 || fmt > 2) { set_err("bad sample format"); return -1; }
    CK(cudaSetDevice(p->dev));
    const DetConfig &dc = p->dc;
    const int N = dc.N;
    const size_t bps = (size_t)fmt_bytes(fmt);
    p->ev_used = 0;
    p->res.kernel_launches = 0; p->res.h2d_bytes = 0; p->res.d2h_bytes = 0;
    if (p->cfg.start_time_ns) p->start_time_ns = p->cfg.start_time_ns;
    else {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);                     // burst_detect.c:755-759
        p->start_time_ns = ts.tv_sec * 1000000000ULL + ts.tv_nsec;
    }
    const int64_t n_frames = (int64_t)(n / (size_t)N);
    p->n_frames_last = n_frames;
    if (p->d_mag.ensure((size_t)std::max<int64_t>(n_frames, 1) * N)) return -1;
    const uint32_t gone_cap = (uint32_t)std::max<size_t>(4096, n / 20000 + 1024);
    if (p->d_gone.ensure(gone_cap)) return -1;
    const void *iq_dev = dev_iq;
    if (host_iq) {
        if (p->d_iq.ensure(n * bps + 64)) return -1;
        iq_dev = p->d_iq.p;
    }
    if (ir_pipeline_reset(p)) return -1;
    cudaEvent_t ev_begin = p->ev();
    CK(cudaEventRecord(ev_begin, p->st_scan));          // after the state reset
    CK(cudaStreamWaitEvent(p->st_fft, ev_begin, 0));
    std::vector<EvPair> ev_fft, ev_scan;
    // chunking: copies (host input only) overlap the detector kernels of earlier chunks
    size_t chunk = p->cfg.h2d_chunk > 0 ? (size_t)p->cfg.h2d_chunk : ((size_t)16 << 20);
    chunk = std::max<size_t>(chunk / N, 1) * N;
    if (!host_iq) chunk = std::max<size_t>((size_t)64 << 20, chunk);    // resident input: few big launches
    for (size_t off = 0; off < n; off += chunk) {
        const size_t m = std::min(chunk, n - off);
        if (host_iq) {
            CK(cudaMemcpyAsync((unsigned char *)p->d_iq.p + off * bps, (const unsigned char *)host_iq + off * bps,
                               m * bps, cudaMemcpyHostToDevice, p->st_copy));
            cudaEvent_t e = p->ev();
            CK(cudaEventRecord(e, p->st_copy));
            CK(cudaStreamWaitEvent(p->st_fft, e, 0));
            p->res.h2d_bytes += m * bps;
        }
        const int64_t f0 = (int64_t)(off / N);
        const int64_t f1 = std::min<int64_t>((int64_t)((off + m) / N), n_frames);
        if (f1 <= f0) continue;
        EvPair a{p->ev(), p->ev()}, b{p->ev(), p->ev()};
        CK(cudaEventRecord(a.a, p->st_fft));
        CK(launch_detect_fft(dc.L, fmt, iq_dev, f0 * N, p->d_window.p, p->d_tw_det.p, p->d_mag.p + f0 * N,
                             f1 - f0, p->sm_count, p->st_fft));
        CK(cudaEventRecord(a.b, p->st_fft));
        CK(cudaStreamWaitEvent(p->st_scan, a.b, 0));
        CK(cudaEventRecord(b.a, p->st_scan));
        CK(launch_detect_scan_auto(dc, p->d_state.p, p->d_base.p, p->d_hist.p, p->d_mag.p + f0 * N, f1 - f0,
                              p->d_gone.p, gone_cap, p->st_scan));
        CK(cudaEventRecord(b.b, p->st_scan));
        ev_fft.push_back(a); ev_scan.push_back(b);
        p->res.kernel_launches += 2;
    }
    if (host_iq) {                                        // the FIR reads iq on st_scan: all copies must be in
        cudaEvent_t e = p->ev();
        CK(cudaEventRecord(e, p->st_copy));
        CK(cudaStreamWaitEvent(p->st_scan, e, 0));
    }
    return finish_run(p, iq_dev, n, fmt, ev_begin, ev_fft, ev_scan);
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
    CK(cudaMemcpy(dst, p->d_frames.p + bi * (size_t)IR_MAX_FRAME, nfl * sizeof(float2), cudaMemcpyDeviceToHost));
    return (int)nfl;
}

extern "C" int ir_pipeline_copy_decimated(ir_pipeline_t *p, size_t bi, float *dst, size_t cap) {
    if (!p || !dst || bi >= p->bursts.size()) return -1;
    const size_t dl = (size_t)p->h_bp[bi].dec_len;
    if (dl == 0 || dl > cap) return -1;
    CK(cudaSetDevice(p->dev));
    CK(cudaMemcpy(dst, p->d_dec.p + p->h_bp[bi].dec_off, dl * sizeof(float2), cudaMemcpyDeviceToHost));
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
    const double ts_ms = (double)(f->timestamp - t0) / 1000000.0;
    const int fhz = (int)(f->center_frequency + 0.5);
    int pay = f->n_payload_symbols;
    if (pay < 0) pay = 0;
    int k = snprintf(dst, cap, "RAW: %s %012.4f %010d N:%05.2f%+06.2f I:%011llu %3d%% %.5f %3d ",
                     file_info ? file_info : "", ts_ms, fhz, f->magnitude, f->noise,
                     (unsigned long long)f->id, f->confidence, f->level, pay);
    if (k < 0) return -1;
    size_t pos = (size_t)k < cap ? (size_t)k : cap - 1;
    for (int i = 0; i < f->n_bits && pos + 2 < cap; i++) dst[pos++] = (char)('0' + bits[i]);
    if (pos + 1 < cap) dst[pos++] = '\n';
    dst[pos] = 0;
    return (int)pos;
}

extern "C" long ir_pipeline_format_raw_all(ir_pipeline_t *p, const char *file_info, uint64_t t0, char *dst,
                                           size_t cap) {
    if (!p) return -1;
    const size_t head = 128 + (file_info ? strlen(file_info) : 0);   // fixed-width fields < 100 chars
    if (!dst) return (long)(p->frames.size() * head + p->bits.size() + 64);
    if (p->frames.empty()) return 0;
    if (t0 == 0) t0 = (p->frames[0].timestamp / 1000000000ULL) * 1000000000ULL;
    size_t pos = 0;
    for (const ir_frame_t &f : p->frames) {
        if (pos + head + (size_t)f.n_bits + 2 > cap) { set_err("ir_pipeline_format_raw_all: buffer too small"); return -1; }
        int n = ir_format_raw(dst + pos, cap - pos, file_info, t0, &f, p->bits.data() + f.bits_offset);
        if (n < 0) return -1;
        pos += (size_t)n;
    }
    return (long)pos;
}

extern "C" void *ir_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void ir_host_free(void *p) { if (p) cudaFreeHost(p); }
