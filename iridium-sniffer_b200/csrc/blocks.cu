// blocks.cu -- one long IQ stream over several pipelines (one per GPU) by contiguous time blocks: the plan and the
// merge (include/iridium_b200.h).  Host bookkeeping only -- the path has no exchange step, so
// there is no collective and nothing here touches a device: each block runs through ir_pipeline_run_* like a file of
// its own (what the reference does with a file cut in pieces: a fresh detector whose first 512 frames only build the
// baseline, burst_detect.c:426-428), and the host puts the frame lists together.
#include <algorithm>
#include <cmath>
#include <numeric>
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

extern "C" long ir_merge_blocks(const ir_config_t *cfg, uint64_t start_time_ns, const ir_block_t *blocks, int n_blocks,
                                const ir_frame_t *const *frames, const size_t *n_frames, ir_frame_t *out,
                                uint32_t *out_block, size_t cap) {
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
            if (k > 0 && f.timestamp < t_first + kOverlapNs) {                            // the previous block's, if it has it
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
    }
    return (long)kept.size();
}
