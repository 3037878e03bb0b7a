// refapi.cu -- the reference's per-item C interface on top of the CUDA kernels
// (include/ir_ref_api.h, include/burst_fft.h).  Each call moves its item to the device, runs
// the same kernels the batched pipeline uses, and returns malloc'd results with the
// reference's ownership rules.  No CPU arithmetic on the path; failures return NULL / 0 / -1.
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "ir_device.cuh"
#include "ir_internal.h"

using namespace ir;

// ---- layouts of include/ir_ref_api.h (float complex == two floats)
// the host program's statistics counter (main.c:181, atomic_ulong); weak: null when loaded stand-alone
extern "C" { extern unsigned long stat_n_detected __attribute__((weak)); }

extern "C" {
typedef struct { uint64_t id, start, stop, last_active; int center_bin; float magnitude, noise; } burst_info_t;
typedef struct {
    burst_info_t info; double center_frequency; int sample_rate; int fft_size;
    uint64_t start_time_ns; size_t num_samples; float *samples;
} burst_data_t;
typedef struct {
    double center_frequency; int sample_rate, fft_size, burst_pre_len, burst_post_len, burst_width,
        max_bursts, max_burst_len; float threshold; int history_size, use_gpu;
} burst_config_t;
typedef void (*burst_callback_t)(burst_data_t *, void *);
typedef struct {
    uint64_t id, timestamp; double center_frequency; float sample_rate, samples_per_symbol;
    int direction; float magnitude, noise, uw_start; size_t num_samples; float *samples;
} downmix_frame_t;
typedef struct { int output_sample_rate, search_depth, handle_multiple_frames; } downmix_config_t;
typedef struct {
    uint64_t id, timestamp; double center_frequency; int direction; float magnitude, noise;
    int confidence; float level; int n_symbols, n_payload_symbols; uint8_t *bits; float *llr; int n_bits;
} demod_frame_t;

// globals of main.c the reference's stages read (main.c:101,141,143); weak so that the
// library also loads without main.c
__attribute__((weak)) int use_gardner = 1;
__attribute__((weak)) int verbose = 0;
}

// simd_kernels.h:105: main.c:567 calls it before any DSP to pick the CPU kernels of simd_*.c.  Those files are
// not part of a build that links this library (every stage runs on the GPU), so there is nothing to select.
extern "C" void simd_init(int force_generic) {
    (void)force_generic;
    fprintf(stderr, "iridium-sniffer: DSP stages run on the GPU (libiridium_b200, sm_100a kernels); no CPU SIMD kernels in this build\n");
}

#define RCK(expr, ret)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            fprintf(stderr, "iridium_b200: %s: %s\n", #expr, cudaGetErrorString(_e));       \
            return ret;                                                                    \
        }                                                                                  \
    } while (0)

static bool device_ok() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        fprintf(stderr, "iridium_b200: no CUDA device (this library has no CPU fallback)\n");
        return false;
    }
    return true;
}

// =========================================================================== plug-in ABI
struct gpu_burst_fft {
    int N, L, batch;
    float *d_window = nullptr, *d_mag = nullptr;
    float2 *d_tw = nullptr, *d_in = nullptr;
    int sm_count = 148;
    cudaStream_t st = nullptr;
};

extern "C" void gpu_burst_fft_destroy(gpu_burst_fft *g) {
    if (!g) return;
    cudaFree(g->d_window); cudaFree(g->d_mag); cudaFree(g->d_tw); cudaFree(g->d_in);
    if (g->st) cudaStreamDestroy(g->st);
    delete g;
}

extern "C" gpu_burst_fft *gpu_burst_fft_create(int fft_size, int batch_size, const float *window) {
    if (!device_ok() || !window || batch_size < 1) return nullptr;
    int L = 0;
    while ((1 << L) < fft_size) L++;
    if ((1 << L) != fft_size || L < 10 || L > 14) {
        fprintf(stderr, "iridium_b200: gpu_burst_fft: fft_size %d unsupported (1024..16384)\n", fft_size);
        return nullptr;
    }
    gpu_burst_fft *g = new gpu_burst_fft();
    g->N = fft_size; g->L = L; g->batch = batch_size;
    cudaDeviceProp prop;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);                                   // the caller's device, not device 0
    if (cudaGetDeviceProperties(&prop, cur_dev) == cudaSuccess) g->sm_count = prop.multiProcessorCount;
    std::vector<float2> tw = build_twiddle_image(L);
    auto bail = [&]() -> gpu_burst_fft * { gpu_burst_fft_destroy(g); return nullptr; };
    if (cudaStreamCreate(&g->st) != cudaSuccess) return bail();
    if (cudaMalloc(&g->d_window, sizeof(float) * fft_size) != cudaSuccess) return bail();
    if (cudaMalloc(&g->d_tw, sizeof(float2) * tw.size()) != cudaSuccess) return bail();
    if (cudaMalloc(&g->d_in, sizeof(float2) * (size_t)fft_size * batch_size) != cudaSuccess) return bail();
    if (cudaMalloc(&g->d_mag, sizeof(float) * (size_t)fft_size * batch_size) != cudaSuccess) return bail();
    if (cudaMemcpy(g->d_window, window, sizeof(float) * fft_size, cudaMemcpyHostToDevice) != cudaSuccess) return bail();
    if (cudaMemcpy(g->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice) != cudaSuccess) return bail();
    if (getenv("IR_PLUGIN_LOG")) fprintf(stderr, "iridium_b200: gpu_burst_fft_create(%d, %d): k_detect_fft on the GPU\n", fft_size, batch_size);
    return g;
}

extern "C" int gpu_burst_fft_process(gpu_burst_fft *g, const float *input, float *output, int batch_count) {
    if (!g || !input || !output || batch_count < 0 || batch_count > g->batch) return -1;
    if (batch_count == 0) return 0;
    const size_t ns = (size_t)batch_count * g->N;
    RCK(cudaMemcpyAsync(g->d_in, input, ns * sizeof(float2), cudaMemcpyHostToDevice, g->st), -1);
    RCK(launch_detect_fft(g->L, IR_FMT_CF32, g->d_in, 0, g->d_window, g->d_tw, g->d_mag, batch_count,
                          g->sm_count, g->st), -1);
    RCK(cudaMemcpyAsync(output, g->d_mag, ns * sizeof(float), cudaMemcpyDeviceToHost, g->st), -1);
    RCK(cudaStreamSynchronize(g->st), -1);
    return 0;
}

// =========================================================================== detector
__global__ void k_convert_ci8(const char2 *__restrict__ in, float2 *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        char2 v = in[i];
        out[i] = make_float2((float)v.x / 128.0f, (float)v.y / 128.0f);       // simd_avx2.c:264-294
    }
}

template <int FMT>
__global__ void k_gather(const void *__restrict__ iq, int64_t n_total, uint64_t ring, int64_t start,
                         int64_t emit, float2 *__restrict__ dst, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = start + i;
        if (q >= emit) q -= (int64_t)ring;
        dst[i] = (q >= 0 && q < n_total) ? load_sample<FMT>(iq, q) : make_float2(0.0f, 0.0f);
    }
}

struct _burst_detector {
    DetConfig dc;
    double center_frequency;
    float *d_window = nullptr, *d_mag = nullptr, *d_base = nullptr, *d_hist = nullptr;
    float2 *d_tw = nullptr;
    DetState *d_state = nullptr;
    GoneBurst *d_gone = nullptr;
    float2 *d_buf = nullptr, *d_buf2 = nullptr, *d_out = nullptr;
    char2 *d_stage = nullptr;
    size_t cap = 0, mag_cap = 0, stage_cap = 0, out_cap = 0;
    uint64_t buf_base = 0, sample_count = 0, index = 0, start_time_ns = 0, n_tagged = 0;
    float peak_db = 0;
    int n_act = 0, sm_count = 148;
    uint32_t gone_cap = 8192;
    cudaStream_t st = nullptr;
};

extern "C" void burst_detector_destroy(_burst_detector *d) {
    if (!d) return;
    cudaFree(d->d_window); cudaFree(d->d_mag); cudaFree(d->d_base); cudaFree(d->d_hist); cudaFree(d->d_tw);
    cudaFree(d->d_state); cudaFree(d->d_gone); cudaFree(d->d_buf); cudaFree(d->d_buf2); cudaFree(d->d_out);
    cudaFree(d->d_stage);
    if (d->st) cudaStreamDestroy(d->st);
    fprintf(stderr, "burst_detect: tagged %lu bursts total\n", (unsigned long)d->n_tagged);   // burst_detect.c:350
    delete d;
}

extern "C" _burst_detector *burst_detector_create(burst_config_t *cfg) {
    if (!cfg || !device_ok()) return nullptr;
    _burst_detector *d = new _burst_detector();
    derive_det_config(d->dc, cfg->sample_rate, cfg->fft_size, cfg->burst_width, cfg->threshold);
    // explicit overrides the reference honours (burst_detect.c:189-217)
    if (cfg->burst_pre_len > 0) d->dc.pre_len = cfg->burst_pre_len;
    if (cfg->burst_post_len > 0) d->dc.post_len = cfg->burst_post_len;
    if (cfg->max_bursts > 0) d->dc.max_bursts = cfg->max_bursts;
    if (cfg->max_burst_len > 0) d->dc.max_burst_len = cfg->max_burst_len;
    if (cfg->history_size > 0 && cfg->history_size != 512) {
        fprintf(stderr, "iridium_b200: history_size %d unsupported (512 only)\n", cfg->history_size);
        delete d;
        return nullptr;
    }
    {
        uint64_t ring = (uint64_t)d->dc.max_burst_len + d->dc.pre_len + d->dc.post_len + (uint64_t)d->dc.N * 4;
        if (ring < (uint64_t)(2 * (int64_t)cfg->sample_rate)) ring = 2 * (uint64_t)cfg->sample_rate;
        d->dc.ringbuf_size = ring;
    }
    d->center_frequency = cfg->center_frequency;
    if (d->dc.L < 10 || d->dc.L > 14) {
        fprintf(stderr, "iridium_b200: detector FFT size %d unsupported\n", d->dc.N);
        delete d;
        return nullptr;
    }
    cudaDeviceProp prop;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cudaGetDeviceProperties(&prop, cur_dev) == cudaSuccess) d->sm_count = prop.multiProcessorCount;
    HostTables tab;
    build_host_tables(tab, d->dc.N);
    std::vector<float2> tw = build_twiddle_image(d->dc.L);
    const size_t N = d->dc.N;
    d->cap = 2 * d->dc.ringbuf_size + 8 * N;
    bool ok = cudaStreamCreate(&d->st) == cudaSuccess &&
              cudaMalloc(&d->d_window, sizeof(float) * N) == cudaSuccess &&
              cudaMalloc(&d->d_tw, sizeof(float2) * tw.size()) == cudaSuccess &&
              cudaMalloc(&d->d_base, sizeof(float) * N) == cudaSuccess &&
              cudaMalloc(&d->d_hist, sizeof(float) * N * 512) == cudaSuccess &&
              cudaMalloc(&d->d_state, sizeof(DetState)) == cudaSuccess &&
              cudaMalloc(&d->d_gone, sizeof(GoneBurst) * d->gone_cap) == cudaSuccess &&
              cudaMalloc(&d->d_buf, sizeof(float2) * d->cap) == cudaSuccess &&
              cudaMemcpy(d->d_window, tab.det_window.data(), sizeof(float) * N, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemset(d->d_state, 0, sizeof(DetState)) == cudaSuccess &&
              cudaMemset(d->d_base, 0, sizeof(float) * N) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "iridium_b200: burst_detector_create: CUDA allocation failed\n");
        burst_detector_destroy(d);
        return nullptr;
    }
    return d;
}

static void detector_feed(_burst_detector *d, const void *host, size_t n, bool is_float,
                          burst_callback_t cb, void *user) {
    if (!d || n == 0) return;
    if (d->start_time_ns == 0) {                               // burst_detect.c:755-759
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        d->start_time_ns = ts.tv_sec * 1000000000ULL + ts.tv_nsec;
    }
    const DetConfig &dc = d->dc;
    const uint64_t N = dc.N, R = dc.ringbuf_size;
    auto die = [&](const char *what, cudaError_t e) {
        fprintf(stderr, "iridium_b200: burst_detector_feed: %s: %s\n", what, cudaGetErrorString(e));
        abort();                                                // no silent CPU fallback
    };
    cudaError_t e;
    // burst list of this call: at most max_bursts bursts can end per post_len samples (the reference caps the ACTIVE
    // bursts, not the emitted ones); grown before the call that could need more
    {
        const size_t need = ((size_t)n / (size_t)std::max(dc.post_len, 1) + 2) * (size_t)std::max(dc.max_bursts, 32);
        if (need > d->gone_cap) {
            cudaFree(d->d_gone);
            d->d_gone = nullptr;
            if ((e = cudaMalloc(&d->d_gone, sizeof(GoneBurst) * need)) != cudaSuccess) die("cudaMalloc(burst list)", e);
            d->gone_cap = (uint32_t)need;
        }
    }
    // room: keep the last R + pre + 2N samples when the linear buffer would overflow
    uint64_t used = d->sample_count - d->buf_base;
    if (used + n > d->cap) {
        uint64_t keep = R + (uint64_t)dc.pre_len + 2 * N;
        if (keep > used) keep = used;
        size_t need = keep + n + 4 * N;
        size_t newcap = need > d->cap ? need : d->cap;
        float2 *dst = nullptr;
        if ((e = cudaMalloc(&dst, sizeof(float2) * newcap)) != cudaSuccess) die("cudaMalloc", e);
        if ((e = cudaMemcpyAsync(dst, d->d_buf + (used - keep), keep * sizeof(float2), cudaMemcpyDeviceToDevice, d->st)) != cudaSuccess) die("compact", e);
        cudaStreamSynchronize(d->st);
        cudaFree(d->d_buf);
        d->d_buf = dst; d->cap = newcap;
        d->buf_base = d->sample_count - keep;
        used = keep;
    }
    // ingest
    if (is_float) {
        if ((e = cudaMemcpyAsync(d->d_buf + used, host, n * sizeof(float2), cudaMemcpyHostToDevice, d->st)) != cudaSuccess) die("H2D", e);
    } else {
        if (n > d->stage_cap) {
            cudaFree(d->d_stage);
            d->stage_cap = n + n / 2;
            if ((e = cudaMalloc(&d->d_stage, sizeof(char2) * d->stage_cap)) != cudaSuccess) die("cudaMalloc", e);
        }
        if ((e = cudaMemcpyAsync(d->d_stage, host, n * sizeof(char2), cudaMemcpyHostToDevice, d->st)) != cudaSuccess) die("H2D", e);
        k_convert_ci8<<<(unsigned)((n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256), 256, 0, d->st>>>(d->d_stage, d->d_buf + used, n);
    }
    const uint64_t before = d->sample_count;
    d->sample_count += n;
    // frames that became complete (burst_detect.c:821-836)
    const int64_t nf = (int64_t)((d->sample_count - d->index) / N);
    if (nf > 0) {
        if ((size_t)nf * N > d->mag_cap) {
            cudaFree(d->d_mag);
            d->mag_cap = (size_t)nf * N * 2;
            if ((e = cudaMalloc(&d->d_mag, sizeof(float) * d->mag_cap)) != cudaSuccess) die("cudaMalloc", e);
        }
        if ((e = launch_detect_fft(dc.L, IR_FMT_CF32, d->d_buf, (int64_t)(d->index - d->buf_base), d->d_window, d->d_tw,
                                   d->d_mag, nf, d->sm_count, d->st)) != cudaSuccess) die("k_detect_fft", e);
        if ((e = launch_detect_scan_auto(dc, d->d_state, d->d_base, d->d_hist, d->d_mag, nf, d->d_gone, d->gone_cap, d->st)) != cudaSuccess) die("k_detect_scan", e);
        d->index += (uint64_t)nf * N;
    }
    DetState hs;
    if ((e = cudaMemcpyAsync(&hs, d->d_state, offsetof(DetState, act), cudaMemcpyDeviceToHost, d->st)) != cudaSuccess) die("D2H", e);
    if ((e = cudaStreamSynchronize(d->st)) != cudaSuccess) die("sync", e);
    d->n_act = hs.n_act;
    if (hs.overflow) { fprintf(stderr, "iridium_b200: detector capacity exceeded\n"); abort(); }
    if (hs.n_gone == 0) return;
    // emit_gone_bursts (burst_detect.c:703-742)
    std::vector<GoneBurst> gone(hs.n_gone);
    if ((e = cudaMemcpy(gone.data(), d->d_gone, gone.size() * sizeof(GoneBurst), cudaMemcpyDeviceToHost)) != cudaSuccess) die("D2H", e);
    uint32_t zero = 0;
    cudaMemcpy((char *)d->d_state + offsetof(DetState, n_gone), &zero, sizeof(zero), cudaMemcpyHostToDevice);
    const uint64_t ring_start = before > R ? before - R : 0;                     // :396-398
    for (const GoneBurst &g : gone) {
        uint64_t xs = g.start < ring_start ? ring_start : g.start;
        uint64_t xe = g.stop + (uint64_t)dc.pre_len;
        if (xe <= xs) continue;
        const size_t ns = (size_t)(xe - xs);
        if (ns > d->out_cap) {
            cudaFree(d->d_out);
            d->out_cap = ns + ns / 4;
            if ((e = cudaMalloc(&d->d_out, sizeof(float2) * d->out_cap)) != cudaSuccess) die("cudaMalloc", e);
        }
        const int blocks = (int)((ns + 255) / 256 > 4096 ? 4096 : (ns + 255) / 256);
        // buffer-relative coordinates; positions before the stream start (first ring lap) read zero
        int64_t rel_start = (int64_t)xs - (int64_t)d->buf_base;
        k_gather<IR_FMT_CF32><<<blocks, 256, 0, d->st>>>(d->d_buf, (int64_t)(d->sample_count - d->buf_base), R,
                                                        rel_start, (int64_t)(d->sample_count - d->buf_base),
                                                        d->d_out, (int64_t)ns);
        burst_data_t *bd = (burst_data_t *)malloc(sizeof(*bd));
        float *smp = (float *)malloc(sizeof(float2) * ns);
        if ((e = cudaMemcpyAsync(smp, d->d_out, ns * sizeof(float2), cudaMemcpyDeviceToHost, d->st)) != cudaSuccess) die("D2H", e);
        cudaStreamSynchronize(d->st);
        bd->info.id = g.id; bd->info.start = g.start; bd->info.stop = g.stop; bd->info.last_active = g.last_active;
        bd->info.center_bin = g.center_bin;
        bd->info.magnitude = 10.0f * log10f(g.peak_rel * dc.hist_size * 1.72f);
        bd->info.noise = 10.0f * log10f(g.base_at_create / dc.hist_size / ((float)dc.N * dc.N) / 1.72f /
                                        ((float)dc.sample_rate / dc.N));
        if (bd->info.magnitude > d->peak_db) d->peak_db = bd->info.magnitude;
        bd->center_frequency = d->center_frequency;
        bd->sample_rate = dc.sample_rate; bd->fft_size = dc.N;
        bd->start_time_ns = d->start_time_ns;
        bd->num_samples = ns; bd->samples = smp;
        if (&stat_n_detected) __atomic_fetch_add(&stat_n_detected, 1ul, __ATOMIC_SEQ_CST);   // burst_detect.c:739
        cb(bd, user);
        d->n_tagged++;
    }
}

extern "C" void burst_detector_feed(_burst_detector *d, const int8_t *iq, size_t n, burst_callback_t cb, void *user) {
    detector_feed(d, iq, n, false, cb, user);
}
extern "C" void burst_detector_feed_cf32(_burst_detector *d, const float *iq, size_t n, burst_callback_t cb, void *user) {
    detector_feed(d, iq, n, true, cb, user);
}
extern "C" int burst_detector_active_count(_burst_detector *d) { return d ? d->n_act : 0; }
extern "C" uint64_t burst_detector_total_count(_burst_detector *d) { return d ? d->n_tagged : 0; }
extern "C" float burst_detector_peak_signal(_burst_detector *d) { return d ? d->peak_db : 0.0f; }
extern "C" float burst_detector_noise_floor(_burst_detector *d) {                 // burst_detect.c:363-380
    if (!d) return 0.0f;
    std::vector<float> b(d->dc.N);
    if (cudaMemcpy(b.data(), d->d_base, sizeof(float) * d->dc.N, cudaMemcpyDeviceToHost) != cudaSuccess) return -120.0f;
    double sum = 0;
    for (float v : b) sum += v;
    float avg = (float)(sum / ((double)d->dc.N * d->dc.hist_size));
    float bw = (float)d->dc.sample_rate / d->dc.N;
    if (avg > 0 && bw > 0) return 10.0f * log10f(avg / bw);
    return -120.0f;
}

// =========================================================================== downmix
struct _burst_downmix {
    HostTables tab;
    float2 *d_tw12 = nullptr, *d_tw11 = nullptr, *d_sync_dl = nullptr, *d_sync_ul = nullptr;
    float2 *d_in = nullptr, *d_dec = nullptr, *d_a = nullptr, *d_b = nullptr, *d_frame = nullptr;
    BurstParam *d_bp = nullptr;
    int *d_tiles = nullptr;
    ChainOut *d_co = nullptr;
    std::unordered_map<long long, std::pair<float2 *, int>> rot;   // (fft_size<<20 | bin) -> table
    cudaStream_t st = nullptr;
};

extern "C" void burst_downmix_destroy(_burst_downmix *dm) {
    if (!dm) return;
    for (auto &kv : dm->rot) cudaFree(kv.second.first);
    cudaFree(dm->d_tw12); cudaFree(dm->d_tw11); cudaFree(dm->d_sync_dl); cudaFree(dm->d_sync_ul);
    cudaFree(dm->d_in); cudaFree(dm->d_dec); cudaFree(dm->d_a); cudaFree(dm->d_b); cudaFree(dm->d_frame);
    cudaFree(dm->d_bp); cudaFree(dm->d_tiles); cudaFree(dm->d_co);
    if (dm->st) cudaStreamDestroy(dm->st);
    delete dm;
}

static bool up(float2 **dst, const std::vector<float2> &v) {
    return cudaMalloc(dst, sizeof(float2) * v.size()) == cudaSuccess &&
           cudaMemcpy(*dst, v.data(), sizeof(float2) * v.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

extern "C" _burst_downmix *burst_downmix_create(downmix_config_t *cfg) {
    if (!device_ok()) return nullptr;
    if (cfg && cfg->output_sample_rate > 0 && cfg->output_sample_rate != IR_OUT_RATE) {
        fprintf(stderr, "iridium_b200: output_sample_rate %d unsupported (250000 only)\n", cfg->output_sample_rate);
        return nullptr;
    }
    _burst_downmix *dm = new _burst_downmix();
    build_host_tables(dm->tab, 1024);
    const size_t dmax = IR_DM_WORK / 40 + 64;
    bool ok = cudaStreamCreate(&dm->st) == cudaSuccess && up(&dm->d_tw12, build_twiddle_image(12)) &&
              up(&dm->d_tw11, build_twiddle_image(11)) && up(&dm->d_sync_dl, dm->tab.sync_dl_fft) &&
              up(&dm->d_sync_ul, dm->tab.sync_ul_fft) &&
              upload_input_taps(dm->tab.h_input.data(), (int)dm->tab.h_input.size()) == cudaSuccess &&
              upload_chain_tables(dm->tab) == cudaSuccess &&
              cudaMalloc(&dm->d_in, sizeof(float2) * IR_DM_WORK) == cudaSuccess &&
              cudaMalloc(&dm->d_dec, sizeof(float2) * dmax) == cudaSuccess &&
              cudaMalloc(&dm->d_a, sizeof(float2) * dmax) == cudaSuccess &&
              cudaMalloc(&dm->d_b, sizeof(float2) * dmax) == cudaSuccess &&
              cudaMalloc(&dm->d_frame, sizeof(float2) * IR_MAX_FRAME) == cudaSuccess &&
              cudaMalloc(&dm->d_bp, sizeof(BurstParam)) == cudaSuccess &&
              cudaMalloc(&dm->d_tiles, sizeof(int) * 2) == cudaSuccess &&
              cudaMalloc(&dm->d_co, sizeof(ChainOut)) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "iridium_b200: burst_downmix_create: CUDA setup failed\n");
        burst_downmix_destroy(dm);
        return nullptr;
    }
    return dm;
}

extern "C" int burst_downmix_process(_burst_downmix *dm, burst_data_t *burst, downmix_frame_t **frames_out) {
    if (frames_out) *frames_out = nullptr;
    if (!dm || !burst || !frames_out || burst->num_samples < 100) return 0;        // burst_downmix.c:645-648
    const int fs = burst->sample_rate, N = burst->fft_size;
    const int dec = (int)roundf((float)fs / IR_OUT_RATE);
    if (dec != 40 && dec != 48) {
        fprintf(stderr, "iridium_b200: decimation %d unsupported (40 or 48)\n", dec);
        return 0;
    }
    int n = (int)burst->num_samples;
    if (n > IR_DM_WORK) n = IR_DM_WORK;
    const int dlen = (n - IR_INPUT_NTAPS + 1) / dec;
    if (dlen < 100) return 0;                                                      // :677-680
    const float rel = (burst->info.center_bin - N / 2) / (float)N;
    const float ph = -2.0f * (float)M_PI * rel;
    float sn, cs;
    sincosf(ph, &sn, &cs);
    // NCO checkpoints for this (fft size, bin)
    const long long key = ((long long)N << 20) | (unsigned)burst->info.center_bin;
    auto it = dm->rot.find(key);
    if (it == dm->rot.end() || it->second.second < n + IR_ROT_G) {
        if (it != dm->rot.end()) cudaFree(it->second.first);
        int len = ((n + IR_ROT_G + 65535) / 65536) * 65536;
        float2 *tbl = nullptr, *d_incr = nullptr, **d_ptr = nullptr;
        int *d_len = nullptr;
        float2 incr = make_float2(cs, sn);
        RCK(cudaMalloc(&tbl, sizeof(float2) * (size_t)(len / IR_ROT_G + 1)), 0);
        RCK(cudaMalloc(&d_incr, sizeof(float2)), 0);
        RCK(cudaMalloc(&d_ptr, sizeof(float2 *)), 0);
        RCK(cudaMalloc(&d_len, sizeof(int)), 0);
        RCK(cudaMemcpy(d_incr, &incr, sizeof(incr), cudaMemcpyHostToDevice), 0);
        RCK(cudaMemcpy(d_ptr, &tbl, sizeof(tbl), cudaMemcpyHostToDevice), 0);
        RCK(cudaMemcpy(d_len, &len, sizeof(len), cudaMemcpyHostToDevice), 0);
        RCK(launch_rot_tables(d_incr, d_ptr, d_len, 1, dm->st), 0);
        RCK(cudaStreamSynchronize(dm->st), 0);
        cudaFree(d_incr); cudaFree(d_ptr); cudaFree(d_len);
        dm->rot[key] = {tbl, len};
        it = dm->rot.find(key);
    }
    BurstParam bp;
    memset(&bp, 0, sizeof(bp));
    bp.start = 0; bp.emit_count = n; bp.n = n; bp.dec_len = dlen; bp.dec_off = 0;
    bp.incr_coarse = make_float2(cs, sn);
    bp.rot_table = it->second.first;
    bp.tile0 = 0;
    bp.cfreq_coarse = burst->center_frequency + (double)(rel * fs);
    const int n_tiles = (dlen + IR_FIR_TILE_OF(dec) - 1) / IR_FIR_TILE_OF(dec);
    int tiles[2] = {0, n_tiles};
    RCK(cudaMemcpyAsync(dm->d_in, burst->samples, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, dm->st), 0);
    RCK(cudaMemcpyAsync(dm->d_bp, &bp, sizeof(bp), cudaMemcpyHostToDevice, dm->st), 0);
    RCK(cudaMemcpyAsync(dm->d_tiles, tiles, sizeof(tiles), cudaMemcpyHostToDevice, dm->st), 0);
    RCK(launch_fir(IR_FMT_CF32, dec, dm->d_in, n, 1ull << 40, dm->d_bp, dm->d_tiles, nullptr, 1, n_tiles, dm->d_dec, dm->st), 0);
    RCK(launch_chain(dm->d_bp, 1, dm->d_dec, dm->d_a, dm->d_b, dm->d_tw12, dm->d_tw11, dm->d_sync_dl, dm->d_sync_ul,
                     dm->d_co, dm->d_frame, dm->st), 0);
    ChainOut co;
    RCK(cudaMemcpyAsync(&co, dm->d_co, sizeof(co), cudaMemcpyDeviceToHost, dm->st), 0);
    RCK(cudaStreamSynchronize(dm->st), 0);
    if (co.status != 0) return 0;
    downmix_frame_t *f = (downmix_frame_t *)malloc(sizeof(*f));
    f->samples = (float *)malloc(sizeof(float2) * (size_t)co.frame_len);
    RCK(cudaMemcpy(f->samples, dm->d_frame, sizeof(float2) * (size_t)co.frame_len, cudaMemcpyDeviceToHost), 0);
    uint64_t ts = burst->start_time_ns + (uint64_t)((double)burst->info.start / fs * 1e9);   // :659-660
    ts += (uint64_t)((IR_INPUT_NTAPS / 2) * 1000000000ULL / fs);                            // :431-433
    f->id = burst->info.id;
    f->timestamp = ts + (uint64_t)((double)co.start / IR_OUT_RATE * 1e9);                   // :783
    f->center_frequency = bp.cfreq_coarse + (double)(co.center_offset * (float)IR_OUT_RATE);
    f->sample_rate = (float)IR_OUT_RATE;
    f->samples_per_symbol = 10.0f;
    f->direction = co.direction;
    f->magnitude = burst->info.magnitude;
    f->noise = burst->info.noise;
    f->uw_start = co.uw_corr;
    f->num_samples = (size_t)co.frame_len;
    *frames_out = f;
    return 1;
}

// =========================================================================== demod
namespace {
struct DemodCtx {
    float2 *d_frame = nullptr;
    ChainOut *d_co = nullptr;
    DemodOut *d_do = nullptr;
    uint8_t *d_bits = nullptr;
    float *d_llr = nullptr;
    cudaStream_t st = nullptr;
    bool ok = false;
};
std::mutex g_demod_mu;
DemodCtx g_demod;
bool demod_ctx() {
    if (g_demod.ok) return true;
    if (!device_ok()) return false;
    g_demod.ok = cudaStreamCreate(&g_demod.st) == cudaSuccess &&
                 cudaMalloc(&g_demod.d_frame, sizeof(float2) * IR_MAX_FRAME) == cudaSuccess &&
                 cudaMalloc(&g_demod.d_co, sizeof(ChainOut)) == cudaSuccess &&
                 cudaMalloc(&g_demod.d_do, sizeof(DemodOut)) == cudaSuccess &&
                 cudaMalloc(&g_demod.d_bits, 2 * IR_MAX_SYMS) == cudaSuccess &&
                 cudaMalloc(&g_demod.d_llr, sizeof(float) * 2 * IR_MAX_SYMS) == cudaSuccess;
    return g_demod.ok;
}
}  // namespace

extern "C" int qpsk_demod(downmix_frame_t *in, demod_frame_t **out) {
    if (out) *out = nullptr;
    if (!in || !out || !in->samples) return 0;
    std::lock_guard<std::mutex> lk(g_demod_mu);
    if (!demod_ctx()) return 0;
    if ((int)(in->samples_per_symbol + 0.5f) != 10 || in->num_samples > IR_MAX_FRAME) {
        fprintf(stderr, "iridium_b200: qpsk_demod supports 10 samples/symbol frames of <= %d samples\n", IR_MAX_FRAME);
        return 0;
    }
    DemodCtx &c = g_demod;
    ChainOut co;
    memset(&co, 0, sizeof(co));
    co.status = 0; co.frame_len = (int)in->num_samples; co.direction = in->direction;
    RCK(cudaMemcpyAsync(c.d_frame, in->samples, sizeof(float2) * in->num_samples, cudaMemcpyHostToDevice, c.st), 0);
    RCK(cudaMemcpyAsync(c.d_co, &co, sizeof(co), cudaMemcpyHostToDevice, c.st), 0);
    RCK(launch_demod(c.d_co, 1, c.d_frame, use_gardner, c.d_do, c.d_bits, c.d_llr, 0, c.st), 0);
    DemodOut d;
    RCK(cudaMemcpyAsync(&d, c.d_do, sizeof(d), cudaMemcpyDeviceToHost, c.st), 0);
    RCK(cudaStreamSynchronize(c.st), 0);
    if (!d.ok) return 0;                                                       // qpsk_demod.c:441-451
    in->direction = d.direction;                                               // :453-463
    demod_frame_t *f = (demod_frame_t *)calloc(1, sizeof(*f));
    f->n_bits = 2 * d.n_symbols;
    f->bits = (uint8_t *)malloc((size_t)(f->n_bits > 0 ? f->n_bits : 1));
    f->llr = (float *)malloc(sizeof(float) * (size_t)(f->n_bits > 0 ? f->n_bits : 1));
    RCK(cudaMemcpy(f->bits, c.d_bits, (size_t)f->n_bits, cudaMemcpyDeviceToHost), 0);
    RCK(cudaMemcpy(f->llr, c.d_llr, sizeof(float) * (size_t)f->n_bits, cudaMemcpyDeviceToHost), 0);
    f->id = in->id; f->timestamp = in->timestamp; f->direction = d.direction;
    f->magnitude = in->magnitude; f->noise = in->noise;
    f->confidence = d.confidence; f->level = d.level;
    f->n_symbols = d.n_symbols; f->n_payload_symbols = d.n_symbols - 12;
    if (d.n_symbols > 0) {                                                     // :521-527
        double dur = (double)d.n_symbols / 25000;
        f->center_frequency = in->center_frequency + d.total_phase / dur / M_PI / 2.0;
    } else {
        f->center_frequency = in->center_frequency;
    }
    *out = f;
    return 1;
}

// ------------------------------------------------------------------------------------------
// Thread functions (burst_detect.c:927-960, burst_downmix.c:801-828).  They only move items between
// the host program's queues (main.c:176-178, blocking_queue.h) and the calls above, so the library
// refers to those queues, the queue functions and the two statistics counters as WEAK symbols: linked
// into the reference's program they bind to main.c's definitions; loaded stand-alone (tests, bench.py)
// they stay null and the thread functions return at once with a message.
extern "C" {
struct Blocking_Queue;                                                  // opaque here: the host program's type
extern struct Blocking_Queue samples_queue __attribute__((weak));       // main.c:176
extern struct Blocking_Queue burst_queue __attribute__((weak));         // main.c:177
extern struct Blocking_Queue frame_queue __attribute__((weak));         // main.c:178
int blocking_queue_take(struct Blocking_Queue *bq, void *element) __attribute__((weak));   // blocking_queue.h:215
int blocking_queue_put(struct Blocking_Queue *bq, void *element) __attribute__((weak));    // :194
int blocking_queue_add(struct Blocking_Queue *bq, void *element) __attribute__((weak));    // :184
extern unsigned long stat_n_dropped __attribute__((weak));              // atomic_ulong, main.c:185
}
#define IR_BQ_FULL 2                                                    // blocking_queue.h:124

struct ir_sample_buf {                                                  // sample_buf_t, sdr.h:13-17
    unsigned num;
    int format;                                                         // 0 = int8 pairs, 1 = float pairs
    int8_t samples[1];
};

static bool host_queues_present() {
    return &samples_queue && &burst_queue && &frame_queue && blocking_queue_take && blocking_queue_put &&
           blocking_queue_add;
}

static void burst_to_queue(burst_data_t *burst, void *user) {          // burst_detect.c:929-937
    if (blocking_queue_put((struct Blocking_Queue *)user, burst) != 0) {
        free(burst->samples);
        free(burst);
        if (&stat_n_dropped) __atomic_fetch_add(&stat_n_dropped, 1ul, __ATOMIC_SEQ_CST);
    }
}

extern "C" void *burst_detector_thread(void *arg) {                     // burst_detect.c:941-960
    _burst_detector *det = (_burst_detector *)arg;
    if (!host_queues_present()) {
        fprintf(stderr, "burst_detector_thread: the host program's queues (samples_queue, burst_queue, blocking_queue_*) are not linked in\n");
        return nullptr;
    }
    for (;;) {
        ir_sample_buf *samples = nullptr;
        if (blocking_queue_take(&samples_queue, &samples) != 0) break;
        if (det) {
            if (samples->format == 1)
                burst_detector_feed_cf32(det, (const float *)samples->samples, samples->num, burst_to_queue, &burst_queue);
            else
                burst_detector_feed(det, samples->samples, samples->num, burst_to_queue, &burst_queue);
        }
        free(samples);
    }
    burst_detector_destroy(det);
    return nullptr;
}

extern "C" void *burst_downmix_thread(void *arg) {                      // burst_downmix.c:801-828
    _burst_downmix *dm = (_burst_downmix *)arg;
    if (!host_queues_present()) {
        fprintf(stderr, "burst_downmix_thread: the host program's queues (burst_queue, frame_queue, blocking_queue_*) are not linked in\n");
        return nullptr;
    }
    for (;;) {
        burst_data_t *burst = nullptr;
        if (blocking_queue_take(&burst_queue, &burst) != 0) break;
        downmix_frame_t *frames = nullptr;
        const int n_frames = dm ? burst_downmix_process(dm, burst, &frames) : 0;
        if (n_frames > 0 && frames) {
            if (blocking_queue_add(&frame_queue, frames) == IR_BQ_FULL) {   // live capture: drop, never block
                free(frames->samples);
                free(frames);
            }
        } else {
            free(frames);
        }
        free(burst->samples);
        free(burst);
    }
    burst_downmix_destroy(dm);
    return nullptr;
}
