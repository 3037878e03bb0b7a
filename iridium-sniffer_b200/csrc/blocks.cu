// blocks.cu -- one long IQ stream over several pipelines (one per GPU) by contiguous time blocks: the plan and the
// merge (include/iridium_b200.h).  Host bookkeeping only -- the path has no exchange step, so
// there is no collective and nothing here touches a device: each block runs through ir_pipeline_run_* like a file of
// its own (what the reference does with a file cut in pieces: a fresh detector whose first 512 frames only build the
// baseline, burst_detect.c:426-428), and the host puts the frame lists together.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <string>
#include <thread>
#include <time.h>
#include <string.h>
#include <vector>

#include "../../include/iridium_b200.h"
#include "ir_internal.h"

namespace ir {
namespace {

struct BlockGeom { size_t unit, halo, tail; int fs; };

bool block_geom(const ir_config_t *cfg, BlockGeom &g) {
    if (!cfg || cfg->sample_rate <= 0 || cfg->fft_size < 0) return false;
    DetConfig dc;
    derive_det_config(dc, cfg->sample_rate, cfg->fft_size, cfg->burst_width_hz, cfg->threshold_db);
    if (dc.N <= 0) return false;
    const size_t N = (size_t)dc.N, feed = cfg->feed_block > 0 ? (size_t)cfg->feed_block : 32768;
    g.unit = N / std::gcd(N, feed) * feed;                     // detector frames AND feed calls stay aligned
    auto up = [&](size_t v) { return (v + g.unit - 1) / g.unit * g.unit; };
    // before: baseline priming, then a burst on the air when detection goes live must be over -- and its successor on
    // the same channel detectable from its first sample (start = index - pre_len, burst_detect.c:609) -- before the
    // owned range begins
    g.halo = up((size_t)dc.hist_size * N + (size_t)dc.max_burst_len + (size_t)dc.post_len + 2 * (size_t)dc.pre_len);
    // after: the longest burst, the wait until it is declared gone, the samples kept after its end, and the feed
    // call at whose end it is emitted (:703-742, :839-841)
    g.tail = up((size_t)dc.max_burst_len + (size_t)dc.post_len + (size_t)dc.pre_len + N + 2 * feed);
    g.fs = cfg->sample_rate;
    return true;
}

uint64_t sample_ns(uint64_t start_time_ns, uint64_t sample, int fs) {                     // burst_downmix.c:659-660
    return start_time_ns + (uint64_t)((double)sample / (double)fs * 1e9);
}

}  // namespace
}  // namespace ir

using namespace ir;

static void set_err(const char *s) { set_last_error(s); }

extern "C" size_t ir_block_halo(const ir_config_t *cfg) {
    BlockGeom g;
    return block_geom(cfg, g) ? g.halo : 0;
}

extern "C" size_t ir_block_tail(const ir_config_t *cfg) {
    BlockGeom g;
    return block_geom(cfg, g) ? g.tail : 0;
}

extern "C" long ir_plan_blocks(const ir_config_t *cfg, size_t n, int n_blocks, ir_block_t *blocks, size_t cap) {
    BlockGeom g;
    if (!block_geom(cfg, g)) { set_err("ir_plan_blocks: bad configuration"); return -1; }
    if (!blocks || n_blocks <= 0) { set_err("ir_plan_blocks: null argument"); return -1; }
    if (n == 0) return 0;
    // equal owned lengths, whole units; a block that owns less than it re-reads is not worth a GPU
    size_t own = (n + (size_t)n_blocks - 1) / (size_t)n_blocks;
    own = (own + g.unit - 1) / g.unit * g.unit;
    if (own <= g.halo) own = g.halo + g.unit;
    size_t k = 0;
    for (size_t first = 0; first < n; first += own, k++) {
        if (k >= cap) { set_err("ir_plan_blocks: block array too small"); return -1; }
        ir_block_t &b = blocks[k];
        b.own_first = first;
        b.own_end = std::min(first + own, n);
        b.feed_first = first >= g.halo ? first - g.halo : 0;
        b.feed_end = std::min<uint64_t>(b.own_end + g.tail, n);
        if (n - b.own_end <= g.unit) { b.own_end = n; b.feed_end = n; k++; break; }       // no sliver of a last block
    }
    return (long)k;
}

static long merge_impl(const ir_config_t *cfg, uint64_t start_time_ns, const ir_block_t *blocks, int n_blocks,
                       const ir_frame_t *const *frames, const size_t *n_frames, ir_frame_t *out,
                       uint32_t *out_block, uint32_t *out_index, size_t cap) {
    if (!cfg || cfg->sample_rate <= 0 || !blocks || !frames || !n_frames || n_blocks < 0 || (cap && (!out || !out_block))) {
        set_err("ir_merge_blocks: null argument");
        return -1;
    }
    const int fs = cfg->sample_rate;
    const uint64_t kOverlapNs = 1000000;          // 1 ms past the owned range, settled against the next block
    const double kSameHz = 200.0;
    struct Kept { uint64_t ts; uint32_t block, index; };
    std::vector<Kept> kept;
    size_t prev_begin = 0;                         // kept[] range of the previous block
    for (int k = 0; k < n_blocks; k++) {
        const size_t this_begin = kept.size();
        if (n_frames[k] && !frames[k]) { set_err("ir_merge_blocks: null frame list"); return -1; }
        const uint64_t t_first = sample_ns(start_time_ns, blocks[k].own_first, fs);
        const bool last = k == n_blocks - 1;
        const uint64_t t_end = sample_ns(start_time_ns, blocks[k].own_end, fs) + (last ? 0 : kOverlapNs);
        for (size_t i = 0; i < n_frames[k]; i++) {
            const ir_frame_t &f = frames[k][i];
            if (f.timestamp < t_first || (!last && f.timestamp >= t_end)) continue;
            if (k > 0 && f.timestamp < t_first + 2 * kOverlapNs) {                        // the previous block's, if it has it
                // (twice the overlap: two blocks stamp the same frame slightly differently -- each has its own
                //  baseline -- and a copy just under the edge in one must still meet its twin just over it in the other)
                bool dup = false;
                for (size_t j = prev_begin; j < this_begin && !dup; j++) {
                    const ir_frame_t &o = frames[kept[j].block][kept[j].index];
                    const uint64_t dt = o.timestamp > f.timestamp ? o.timestamp - f.timestamp : f.timestamp - o.timestamp;
                    dup = dt < kOverlapNs && std::fabs(o.center_frequency - f.center_frequency) < kSameHz;
                }
                if (dup) continue;
            }
            kept.push_back(Kept{f.timestamp, (uint32_t)k, (uint32_t)i});
        }
        prev_begin = this_begin;
    }
    std::stable_sort(kept.begin(), kept.end(), [](const Kept &a, const Kept &b) { return a.ts < b.ts; });
    if (kept.size() > cap) { set_err("ir_merge_blocks: output arrays too small"); return -1; }
    for (size_t i = 0; i < kept.size(); i++) {
        out[i] = frames[kept[i].block][kept[i].index];
        out[i].id = (uint64_t)kept[i].block * IR_BLOCK_ID_STRIDE + out[i].id % IR_BLOCK_ID_STRIDE;
        out_block[i] = kept[i].block;
        if (out_index) out_index[i] = kept[i].index;
    }
    return (long)kept.size();
}

extern "C" long ir_merge_blocks(const ir_config_t *cfg, uint64_t start_time_ns, const ir_block_t *blocks, int n_blocks,
                                const ir_frame_t *const *frames, const size_t *n_frames, ir_frame_t *out,
                                uint32_t *out_block, size_t cap) {
    return merge_impl(cfg, start_time_ns, blocks, n_blocks, frames, n_frames, out, out_block, nullptr, cap);
}

// ---- one process, several GPUs: a pipeline and a host thread per device, blocks dealt round-robin
struct ir_multi {
    ir_config_t cfg;
    std::vector<int> devices;
    std::vector<ir_pipeline_t *> pipes;
    // last run
    std::vector<ir_block_t> blocks;
    std::vector<std::vector<ir_frame_t>> frames;       // per block
    std::vector<std::vector<uint8_t>> bits;
    std::vector<std::vector<float>> llr;
    std::vector<const uint8_t *> bits_ptr;
    std::vector<const float *> llr_ptr;
    std::vector<ir_frame_t> merged;
    std::vector<uint32_t> merged_block, merged_index;
    bool classify = false;                             // ir_multi_set_classify: frame_decode() + ida_decode() per frame too
    bool classified_last = false;                      // ... and whether the LAST run was (what results / parsed text go by)
    std::vector<std::vector<ir_frame_class_t>> cls;    // per block, parallel to frames[k]
    std::vector<const ir_frame_class_t *> cls_ptr;
    uint64_t start_time_ns = 0, launches = 0, fed = 0;
};

extern "C" int ir_multi_set_classify(ir_multi_t *m, int on) {
    if (!m) { set_err("ir_multi_set_classify: null argument"); return -1; }
    m->classify = on != 0;
    return 0;
}

extern "C" void ir_multi_destroy(ir_multi_t *m) {
    if (!m) return;
    for (ir_pipeline_t *p : m->pipes)
        if (p) ir_pipeline_destroy(p);
    delete m;
}

extern "C" ir_multi_t *ir_multi_create(const ir_config_t *cfg, const int *devices, int n_devices) {
    if (!cfg || !devices || n_devices <= 0) { set_err("ir_multi_create: null argument"); return nullptr; }
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) { set_err("ir_multi_create: a device is listed twice (one pipeline per GPU)"); return nullptr; }
    ir_multi_t *m = new ir_multi();
    m->cfg = *cfg;
    m->devices.assign(devices, devices + n_devices);
    for (int i = 0; i < n_devices; i++) {
        ir_config_t c = *cfg;
        c.device = devices[i];
        ir_pipeline_t *p = ir_pipeline_create(&c);              // (sets ir_last_error on failure: no device, no fallback)
        if (!p) { ir_multi_destroy(m); return nullptr; }
        m->pipes.push_back(p);
    }
    return m;
}

extern "C" int ir_multi_run_host(ir_multi_t *m, const void *iq, size_t n, int fmt, int n_blocks) {
    if (!m || (!iq && n)) { set_err("ir_multi_run_host: null argument"); return -1; }
    if (fmt < 0 || fmt > 2) { set_err("bad sample format"); return -1; }
    const size_t bps = fmt == IR_FMT_CF32 ? 8 : fmt == IR_FMT_CI16 ? 4 : 2;
    const int nd = (int)m->pipes.size();
    if (n_blocks <= 0) n_blocks = nd;
    m->classified_last = m->classify;
    m->start_time_ns = m->cfg.start_time_ns;
    if (!m->start_time_ns) {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);                     // burst_detect.c:755-759, once for the whole stream
        m->start_time_ns = (uint64_t)ts.tv_sec * 1000000000ULL + (uint64_t)ts.tv_nsec;
    }
    m->blocks.assign((size_t)n_blocks, ir_block_t{});
    const long nb = ir_plan_blocks(&m->cfg, n, n_blocks, m->blocks.data(), m->blocks.size());
    if (nb < 0) return -1;
    m->blocks.resize((size_t)nb);
    m->frames.assign((size_t)nb, {}); m->bits.assign((size_t)nb, {}); m->llr.assign((size_t)nb, {}); m->cls.assign((size_t)nb, {});
    m->merged.clear(); m->merged_block.clear(); m->merged_index.clear();
    m->launches = 0; m->fed = 0;
    std::vector<std::string> errs((size_t)nd);
    std::vector<uint64_t> launches((size_t)nd, 0);
    auto work = [&](int d) {
        ir_pipeline_t *p = m->pipes[(size_t)d];
        for (long k = d; k < nb; k += nd) {
            const ir_block_t &b = m->blocks[(size_t)k];
            ir_results_t r;
            if (ir_pipeline_set_start_time(p, m->start_time_ns) || ir_pipeline_set_origin(p, b.feed_first) ||
                ir_pipeline_run_host(p, (const unsigned char *)iq + b.feed_first * bps, (size_t)(b.feed_end - b.feed_first), fmt) ||
                ir_pipeline_results(p, &r)) {
                errs[(size_t)d] = std::string("block ") + std::to_string(k) + " on device " + std::to_string(m->devices[(size_t)d]) +
                                  ": " + ir_last_error();
                break;
            }
            m->frames[(size_t)k].assign(r.frames, r.frames + r.n_frames);
            m->bits[(size_t)k].assign(r.bits, r.bits + r.n_bits_total);
            m->llr[(size_t)k].assign(r.llr, r.llr + r.n_bits_total);
            launches[(size_t)d] += r.kernel_launches;
            if (m->classify && r.n_frames) {            // while the block's bits and LLRs are still in device memory
                m->cls[(size_t)k].resize(r.n_frames);
                if (ir_pipeline_classify(p, m->cls[(size_t)k].data(), r.n_frames) < 0) {
                    errs[(size_t)d] = std::string("classifying block ") + std::to_string(k) + " on device " +
                                      std::to_string(m->devices[(size_t)d]) + ": " + ir_last_error();
                    break;
                }
                launches[(size_t)d] += 1;
            }
        }
        ir_pipeline_set_origin(p, 0);
    };
    std::vector<std::thread> th;
    for (int d = 1; d < nd; d++) th.emplace_back(work, d);
    work(0);                                                    // the calling thread takes the first device
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (!errs[(size_t)d].empty()) { set_last_error("ir_multi_run_host: " + errs[(size_t)d]); return -1; }
    std::vector<const ir_frame_t *> fl((size_t)nb);
    std::vector<size_t> nf((size_t)nb);
    size_t total = 0;
    m->bits_ptr.resize((size_t)nb); m->llr_ptr.resize((size_t)nb); m->cls_ptr.resize((size_t)nb);
    for (long k = 0; k < nb; k++) {
        fl[(size_t)k] = m->frames[(size_t)k].data(); nf[(size_t)k] = m->frames[(size_t)k].size(); total += nf[(size_t)k];
        m->bits_ptr[(size_t)k] = m->bits[(size_t)k].data(); m->llr_ptr[(size_t)k] = m->llr[(size_t)k].data();
        m->cls_ptr[(size_t)k] = m->cls[(size_t)k].empty() ? nullptr : m->cls[(size_t)k].data();
        m->fed += m->blocks[(size_t)k].feed_end - m->blocks[(size_t)k].feed_first;
    }
    for (uint64_t l : launches) m->launches += l;
    m->merged.resize(total + 1); m->merged_block.resize(total + 1); m->merged_index.resize(total + 1);
    const long kept = merge_impl(&m->cfg, m->start_time_ns, m->blocks.data(), (int)nb, fl.data(), nf.data(),
                                 m->merged.data(), m->merged_block.data(), m->merged_index.data(), total);
    if (kept < 0) return -1;
    m->merged.resize((size_t)kept); m->merged_block.resize((size_t)kept); m->merged_index.resize((size_t)kept);
    return 0;
}

extern "C" int ir_multi_results(ir_multi_t *m, ir_multi_results_t *out) {
    if (!m || !out) { set_err("ir_multi_results: null argument"); return -1; }
    out->n_frames = m->merged.size();
    out->frames = m->merged.data();
    out->block = m->merged_block.data();
    out->index = m->merged_index.data();
    out->classes = m->classified_last ? m->cls_ptr.data() : nullptr;
    out->n_blocks = m->blocks.size();
    out->blocks = m->blocks.data();
    out->bits = m->bits_ptr.data();
    out->llr = m->llr_ptr.data();
    out->start_time_ns = m->start_time_ns;
    out->kernel_launches = m->launches;
    out->samples_fed = m->fed;
    return 0;
}

extern "C" long ir_multi_format_raw_all(ir_multi_t *m, const char *file_info, uint64_t t0, char *dst, size_t cap) {
    if (!m) return -1;
    const size_t head = 128 + (file_info ? strlen(file_info) : 0);
    size_t need = 64;
    for (const ir_frame_t &f : m->merged) need += head + (size_t)f.n_bits + 2;
    if (!dst) return (long)need;
    if (m->merged.empty()) return 0;
    if (t0 == 0) t0 = (m->merged[0].timestamp / 1000000000ULL) * 1000000000ULL;       // frame_output.c:144-158
    size_t pos = 0;
    for (size_t i = 0; i < m->merged.size(); i++) {
        const ir_frame_t &f = m->merged[i];
        if (pos + head + (size_t)f.n_bits + 2 > cap) { set_err("ir_multi_format_raw_all: buffer too small"); return -1; }
        const int k = ir_format_raw(dst + pos, cap - pos, file_info, t0, &f, m->bits[m->merged_block[i]].data() + f.bits_offset);
        if (k < 0) { set_err("ir_multi_format_raw_all: formatting failed"); return -1; }
        pos += (size_t)k;
    }
    return (long)pos;
}

// `--parsed` output of the merged run (main.c:328-331): the IDA line where ida_decode() accepted the frame, the RAW line otherwise
extern "C" long ir_multi_format_parsed_all(ir_multi_t *m, const char *file_info, uint64_t t0, char *dst, size_t cap) {
    if (!m) return -1;
    if (!m->classified_last) { set_err("ir_multi_format_parsed_all: the last run was not classified (ir_multi_set_classify before the run)"); return -1; }
    const size_t head = 512 + (file_info ? strlen(file_info) : 0);
    size_t need = 64;
    for (const ir_frame_t &f : m->merged) need += head + (size_t)f.n_bits + 2;
    if (!dst) return (long)need;
    if (m->merged.empty()) return 0;
    if (t0 == 0) t0 = (m->merged[0].timestamp / 1000000000ULL) * 1000000000ULL;
    size_t pos = 0;
    for (size_t i = 0; i < m->merged.size(); i++) {
        const ir_frame_t &f = m->merged[i];
        const uint32_t b = m->merged_block[i];
        if (b >= m->cls.size() || m->merged_index[i] >= m->cls[b].size()) { set_err("ir_multi_format_parsed_all: no class record for a frame"); return -1; }
        const ir_frame_class_t &c = m->cls[b][m->merged_index[i]];
        if (pos + head + (size_t)f.n_bits + 2 > cap) { set_err("ir_multi_format_parsed_all: buffer too small"); return -1; }
        const int k = c.ida_ok ? ir_format_ida(dst + pos, cap - pos, t0, &f, &c)
                               : ir_format_raw(dst + pos, cap - pos, file_info, t0, &f, m->bits[b].data() + f.bits_offset);
        if (k < 0) { set_err("ir_multi_format_parsed_all: formatting failed"); return -1; }
        pos += (size_t)k;
    }
    return (long)pos;
}

// Independent streams (BASELINE config 5: one recording per GPU, nothing shared): stream s on device s % n_devices, each
// device working through its streams in order.  No halo, no merge -- the result lists the streams one after the other,
// each in its pipeline's own order, `block` = the stream's index, ids = stream * IR_BLOCK_ID_STRIDE + id.
extern "C" int ir_multi_run_streams_host(ir_multi_t *m, const void *const *iq, const size_t *n_samples, int n_streams, int fmt) {
    if (!m || !iq || !n_samples || n_streams <= 0) { set_err("ir_multi_run_streams_host: null argument"); return -1; }
    if (fmt < 0 || fmt > 2) { set_err("bad sample format"); return -1; }
    const int nd = (int)m->pipes.size();
    const size_t ns = (size_t)n_streams;
    m->classified_last = m->classify;
    m->start_time_ns = m->cfg.start_time_ns;
    if (!m->start_time_ns) {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        m->start_time_ns = (uint64_t)ts.tv_sec * 1000000000ULL + (uint64_t)ts.tv_nsec;
    }
    m->blocks.assign(ns, ir_block_t{});
    for (size_t k = 0; k < ns; k++) { m->blocks[k].feed_end = n_samples[k]; m->blocks[k].own_end = n_samples[k]; }
    m->frames.assign(ns, {}); m->bits.assign(ns, {}); m->llr.assign(ns, {}); m->cls.assign(ns, {});
    m->merged.clear(); m->merged_block.clear(); m->merged_index.clear();
    m->launches = 0; m->fed = 0;
    std::vector<std::string> errs((size_t)nd);
    std::vector<uint64_t> launches((size_t)nd, 0);
    auto work = [&](int d) {
        ir_pipeline_t *p = m->pipes[(size_t)d];
        for (size_t k = (size_t)d; k < ns; k += (size_t)nd) {
            ir_results_t r;
            if (ir_pipeline_set_start_time(p, m->start_time_ns) || ir_pipeline_set_origin(p, 0) ||
                ir_pipeline_run_host(p, iq[k], n_samples[k], fmt) || ir_pipeline_results(p, &r)) {
                errs[(size_t)d] = std::string("stream ") + std::to_string(k) + " on device " + std::to_string(m->devices[(size_t)d]) +
                                  ": " + ir_last_error();
                break;
            }
            m->frames[k].assign(r.frames, r.frames + r.n_frames);
            m->bits[k].assign(r.bits, r.bits + r.n_bits_total);
            m->llr[k].assign(r.llr, r.llr + r.n_bits_total);
            launches[(size_t)d] += r.kernel_launches;
            if (m->classify && r.n_frames) {
                m->cls[k].resize(r.n_frames);
                if (ir_pipeline_classify(p, m->cls[k].data(), r.n_frames) < 0) {
                    errs[(size_t)d] = std::string("classifying stream ") + std::to_string(k) + ": " + ir_last_error();
                    break;
                }
                launches[(size_t)d] += 1;
            }
        }
    };
    std::vector<std::thread> th;
    for (int d = 1; d < nd; d++) th.emplace_back(work, d);
    work(0);
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (!errs[(size_t)d].empty()) { set_last_error("ir_multi_run_streams_host: " + errs[(size_t)d]); return -1; }
    m->bits_ptr.resize(ns); m->llr_ptr.resize(ns); m->cls_ptr.resize(ns);
    for (size_t k = 0; k < ns; k++) {
        m->bits_ptr[k] = m->bits[k].data(); m->llr_ptr[k] = m->llr[k].data();
        m->cls_ptr[k] = m->cls[k].empty() ? nullptr : m->cls[k].data();
        m->fed += n_samples[k];
        for (size_t i = 0; i < m->frames[k].size(); i++) {
            ir_frame_t f = m->frames[k][i];
            f.id = (uint64_t)k * IR_BLOCK_ID_STRIDE + f.id % IR_BLOCK_ID_STRIDE;
            m->merged.push_back(f); m->merged_block.push_back((uint32_t)k); m->merged_index.push_back((uint32_t)i);
        }
    }
    for (uint64_t l : launches) m->launches += l;
    return 0;
}
