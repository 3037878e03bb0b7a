// k_detect_cluster.cu -- the burst state machine (burst_detect.c:426-632) on a thread-block
// cluster of 8 CTAs, frames processed in speculative batches.
//
// Why: one CTA walking the frames one by one is bound by barrier / dependency latency and by a
// single SM's ingest rate (4N bytes per frame).  Here every CTA owns N/8 bins (baseline in
// registers, its slice of every magnitude row) and the frame-to-frame dependency is broken by
// the observation that the noise baseline changes only on "quiet" frames (no active burst,
// burst_detect.c:438-454), on which nothing else happens:
//
//   batch type F (bursts active): the baseline is frozen, so the owners compute the
//     "above threshold" bitmaps of K frames at once; the leader CTA then replays the K frames
//     through the sparse state machine (hysteresis, deletion, masks, peak picking, squelch).
//     The batch is cut after the first frame that ends with a baseline update.
//   batch type Q (no burst active): the owners assume every frame is quiet, update their
//     baselines frame by frame (per-bin serial, exactly the reference's two roundings) and
//     produce the bitmaps along the way; the leader only has to confirm that no valid bin crossed
//     the threshold.  The first frame that does cuts the batch: owners rewind to that frame
//     (recompute from the batch start) and the next batch is of type F.
//
// Two cluster barriers per batch instead of two CTA barriers per frame; bitmaps travel to the
// leader through distributed shared memory.  Results are identical to the single-CTA kernel in
// k_detect.cu (kept for N < 2048 and as a cross-check; tests compare both with the CPU oracle).
#include <cooperative_groups.h>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace cg = cooperative_groups;

namespace ir {

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int CT = 256;        // threads per CTA
constexpr int KB = 32;         // frames per batch
static_assert(KB == 32, "the leader's frame search maps one frame to one lane");
constexpr int MAXW = 512;      // bitmap words per frame (N <= 16384)
constexpr int MAXC = 2048;     // candidate peaks kept in shared memory per frame

struct ClShared {
    uint32_t words[KB][MAXW];      // leader only: above-threshold bitmaps of the batch
    uint32_t myw[KB][MAXW / CL];   // every CTA: its own words of the batch, shipped once per batch
    uint32_t free_mask[MAXW];      // 1 = bin not covered by an active burst
    uint32_t valid[MAXW];
    uint32_t cand[MAXW];
    ArgMax red[32];
    // control block, written by the leader, read by every CTA after barrier B
    int ctl_commit;                // frames of this batch that stand
    int ctl_push_forced;           // baseline updates to apply at the last committed frame (type F)
    int ctl_push_normal;
    int ctl_reset_noise;           // squelch reset happened at the last committed frame
    int ctl_next_type;             // 0 = F, 1 = Q
    // leader machine state
    int n_cand;                    // candidate peaks of the frame being processed
    int cbin[MAXC];
    float crel[MAXC];
    float cbase[MAXC];
    int n_act;
    int flags;
    int ff_frame;
    int squelch_count;
    unsigned long long next_id;
    uint32_t n_gone, n_squelch, overflow;
    ActBurst act[IR_MAX_ACTIVE];
};

__device__ __forceinline__ bool cbit(const uint32_t *bm, int bin) { return (bm[bin >> 5] >> (bin & 31)) & 1u; }

__device__ __forceinline__ void cclear(uint32_t *bm, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi >> 5); w++) {
        int a = max(lo, w << 5) & 31, b = min(hi, (w << 5) + 31) & 31;
        uint32_t m = (b == 31 ? 0xffffffffu : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
        atomicAnd(&bm[w], ~m);
    }
}

__device__ __forceinline__ void cgone(ClShared &S, GoneBurst *gone, uint32_t cap, const ActBurst &b, uint64_t index) {
    if (S.n_gone < cap) {
        GoneBurst g;
        g.id = b.id; g.start = b.start; g.stop = index; g.last_active = b.last_active;
        g.center_bin = b.center_bin; g.peak_rel = b.peak_rel; g.base_at_create = b.base_at_create; g.pad = 0;
        gone[S.n_gone] = g;
    } else {
        S.overflow = 1;
    }
    S.n_gone++;
}

}  // namespace

// BPT = bins per thread = N / (CL * CT)
template <int BPT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 1)
k_detect_scan_cluster(DetConfig c, DetState *__restrict__ gs, float *__restrict__ base_g,
                      float *__restrict__ hist, const float *__restrict__ mag, int64_t n_frames,
                      GoneBurst *__restrict__ gone, uint32_t gone_cap) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClShared &S = *reinterpret_cast<ClShared *>(smem_raw);
    ClShared &LS = *cluster.map_shared_rank(&S, 0);           // the leader's copy (DSMEM)
    const int rank = (int)cluster.block_rank();
    const bool leader = rank == 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, W = N >> 5;
    constexpr int NB = BPT * CT;                              // bins per CTA
    constexpr int WPC = NB / 32;                              // words per CTA per frame
    const int bin0 = rank * NB;                               // first bin of this CTA
    const float thr = c.thr;
    const float thr_lo = thr * 0.99999f, thr_hi = thr * 1.00001f;
    const float INF = __int_as_float(0x7f800000);

    // ---- per-bin state: bin(u) = bin0 + u*CT + tid
    float base[BPT];
#pragma unroll
    for (int u = 0; u < BPT; u++) base[u] = base_g[bin0 + u * CT + tid];
    int hist_idx = gs->hist_idx, primed = gs->primed;         // replicated in every CTA
    uint64_t index = gs->index;                               // sample index of frame k0

    if (leader) {
        if (tid == 0) {
            S.n_act = gs->n_act; S.squelch_count = gs->squelch_count; S.next_id = gs->next_id;
            S.n_gone = gs->n_gone; S.n_squelch = gs->n_squelch; S.overflow = gs->overflow; S.flags = 0;
            S.ctl_next_type = gs->n_act > 0 ? 0 : 1;
        }
        for (int i = tid; i < gs->n_act && i < IR_MAX_ACTIVE; i += CT) S.act[i] = gs->act[i];
        for (int w = tid; w < W; w += CT) {
            uint32_t v = 0;
            for (int b = 0; b < 32; b++) {
                int bin = (w << 5) + b;
                bool ok = bin >= c.half_bw && bin < N - c.half_bw && !(bin >= N / 2 - 3 && bin <= N / 2 + 3);
                v |= ok ? (1u << b) : 0u;
            }
            S.valid[w] = v;
            S.free_mask[w] = 0xffffffffu;
        }
        __syncthreads();
        if (tid < S.n_act) cclear(S.free_mask, max(S.act[tid].center_bin - c.half_bw, 0),
                                  min(S.act[tid].center_bin + c.half_bw, N - 1));
    }
    cluster.sync();
    int type = LS.ctl_next_type;

    // one baseline update of the owned bins with magnitude row `row` (burst_detect.c:438-454,
    // simd_avx2.c:221-236); `write_hist` = also store the row into the history
    auto push_row = [&](const float *__restrict__ row, bool write_hist) {
        float *h = hist + (size_t)hist_idx * N + bin0;
#pragma unroll
        for (int u = 0; u < BPT; u++) {
            const int o = u * CT + tid;
            const float mv = row[bin0 + o];
            const float old = primed ? h[o] : 0.0f;            // untouched history is zero
            const float v = base[u] - old;
            base[u] = v + mv;
            if (write_hist) h[o] = mv;
        }
        if (++hist_idx == c.hist_size) { primed = 1; hist_idx = 0; }
    };
    // bitmap words of one frame (values already in registers) for the owned bins -> leader's words[j]
    auto screen_vals = [&](const float (&mv)[BPT], int j) {
        bool pass_any = false;
#pragma unroll
        for (int u = 0; u < BPT; u++) {
            const float lim = base[u] > 0.0f ? base[u] * thr_lo : INF;
            pass_any = pass_any | (mv[u] > lim);
        }
        const bool warp_pass = __any_sync(0xffffffffu, pass_any);
#pragma unroll
        for (int u = 0; u < BPT; u++) {
            uint32_t b = 0;
            if (warp_pass) {
                // rel = mag/base > thr (IEEE divide, simd_avx2.c:239-257).  Outside the band
                // base*thr*(1 -+ 1e-5) the outcome is decided by a product; the divide runs only
                // for values inside it.
                bool ab = false;
                if (base[u] > 0.0f) {
                    if (mv[u] > base[u] * thr_hi) ab = true;
                    else if (mv[u] > base[u] * thr_lo) ab = mv[u] / base[u] > thr;
                }
                b = __ballot_sync(0xffffffffu, ab);
            }
            if (lane == 0) S.myw[j][u * (CT / 32) + warp] = b;
        }
    };
    // Frames are consumed in groups of G whose values (and, for quiet batches, the history rows
    // they replace) are all requested before the first is used: one memory round trip per group
    // instead of one per frame.
    constexpr int G = BPT >= 8 ? 4 : (BPT == 4 ? 8 : 16);
    float *stage = reinterpret_cast<float *>(smem_raw + ((sizeof(ClShared) + 127) / 128) * 128);   // [2][G][NB]

    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long sub[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // leader phase-2 breakdown
    unsigned long long p1s[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // owner phase-1 breakdown (rank 3)
    long long tsub = 0;
    long long tprev = clock64();
// Cycle counters per phase (tools/scan_debug.py): compiled in only with -DIR_SCAN_TIMING.
#ifdef IR_SCAN_TIMING
#define IR_TICK(i) do { long long _t = clock64(); tacc[i] += (unsigned long long)(_t - tprev); tprev = _t; } while (0)
#define IR_SUB(i) do { long long _t = clock64(); sub[i] += (unsigned long long)(_t - tsub); tsub = _t; } while (0)
#define IR_P1(i) do { long long _t = clock64(); p1s[i] += (unsigned long long)(_t - tq); tq = _t; } while (0)
#define IR_COUNT(x) do { x; } while (0)
#else
#define IR_TICK(i) do { } while (0)
#define IR_SUB(i) do { } while (0)
#define IR_P1(i) do { } while (0)
#define IR_COUNT(x) do { } while (0)
#endif
    int64_t k0 = 0;
    while (k0 < n_frames) {
        tprev = clock64();
        const int Kb = (int)min((int64_t)KB, n_frames - k0);
        const float *rows = mag + (size_t)k0 * N;
        // ---------------- phase 1: owners
        float base0[BPT];
        const int idx0 = hist_idx, primed0 = primed;
#pragma unroll
        for (int u = 0; u < BPT; u++) base0[u] = base[u];
        // The owned slices of the magnitude rows are staged through shared memory with 16-byte
        // asynchronous copies, one group of G frames ahead of the group being consumed.
        auto stage_group = [&](int j0s) {
            float *dst = stage + (size_t)((j0s / G) & 1) * G * NB;
            for (int g = 0; g < G; g++) {
                if (j0s + g < Kb) {
                    const float *src = rows + (size_t)(j0s + g) * N + bin0;
                    for (int ch = tid; ch < NB / 4; ch += CT) cp_async_16(dst + (size_t)g * NB + 4 * ch, src + 4 * ch);
                }
            }
            cp_async_commit();
        };
        long long tq = clock64();
        stage_group(0);
        for (int j0 = 0; j0 < Kb; j0 += G) {
            IR_P1(4);
            const bool more = j0 + G < Kb;
            if (more) stage_group(j0 + G);
            IR_P1(0);
            float mv[G][BPT], ov[G][BPT];
            if (type == 1) {
#pragma unroll
                for (int g = 0; g < G; g++) {       // history rows of a quiet batch: plain loads, in flight
                    const bool in = j0 + g < Kb;     // while the staged magnitudes are awaited
                    int hrow = hist_idx + g;
                    if (hrow >= c.hist_size) hrow -= c.hist_size;
                    const float *h = hist + (size_t)hrow * N + bin0;
                    // a history row is live if the detector is primed now or wraps before reaching it
                    const bool live = primed || (hist_idx + g >= c.hist_size);
#pragma unroll
                    for (int u = 0; u < BPT; u++) ov[g][u] = (in && live) ? h[u * CT + tid] : 0.0f;
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int u = 0; u < BPT; u++) ov[g][u] = 0.0f;
            }
            IR_P1(1);
            if (more) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
            __syncthreads();
            IR_P1(2);
            {
                const float *srcs = stage + (size_t)((j0 / G) & 1) * G * NB;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const bool in = j0 + g < Kb;
#pragma unroll
                    for (int u = 0; u < BPT; u++) mv[g][u] = in ? srcs[(size_t)g * NB + u * CT + tid] : 0.0f;
                }
            }
            __syncthreads();                         // buffer may be refilled two groups later
            IR_P1(3);
            // Frozen baseline: one vote decides for the whole group whether this warp's bins need
            // the per-frame test at all (they do only where a burst sits).
            bool group_quiet = false;
            if (type == 0) {
                bool pass_any = false;
#pragma unroll
                for (int u = 0; u < BPT; u++) {
                    const float lim = base[u] > 0.0f ? base[u] * thr_lo : INF;
#pragma unroll
                    for (int g = 0; g < G; g++) pass_any = pass_any | (mv[g][u] > lim);
                }
                group_quiet = !__any_sync(0xffffffffu, pass_any);
                if (group_quiet) {                      // G*BPT == 32 zero words, one per lane
                    const int g = lane / BPT, u = lane % BPT;
                    if (G * BPT == 32) {
                        if (j0 + g < Kb) S.myw[j0 + g][u * (CT / 32) + warp] = 0;
                    } else {
                        for (int q = lane; q < G * BPT; q += 32)
                            if (j0 + q / BPT < Kb) S.myw[j0 + q / BPT][(q % BPT) * (CT / 32) + warp] = 0;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int j = j0 + g;
                if (j < Kb) {
                    if (type == 0) {
                        if (!group_quiet) screen_vals(mv[g], j);
                    } else {
                        if (primed) {
                            screen_vals(mv[g], j);
                        } else if (lane == 0) {
#pragma unroll
                            for (int u = 0; u < BPT; u++) S.myw[j][u * (CT / 32) + warp] = 0;
                        }
                        // speculative baseline update (history written at commit)
#pragma unroll
                        for (int u = 0; u < BPT; u++) {
                            const float v = base[u] - ov[g][u];
                            base[u] = v + mv[g][u];
                        }
                        if (++hist_idx == c.hist_size) { primed = 1; hist_idx = 0; }
                    }
                }
            }
        }
        IR_P1(4);
        // ship this CTA's words of the whole batch to the leader: 16-byte DSMEM stores
        __syncthreads();
        if (WPC >= 4) {
            constexpr int CH = WPC >= 4 ? WPC / 4 : 1;                 // 16-byte chunks per frame
            for (int i = tid; i < Kb * CH; i += CT) {
                const int j = i / CH, k = i % CH;
                const uint4 v = *reinterpret_cast<const uint4 *>(&S.myw[j][4 * k]);
                *reinterpret_cast<uint4 *>(&LS.words[j][rank * WPC + 4 * k]) = v;
            }
        } else {
            for (int i = tid; i < Kb * WPC; i += CT) LS.words[i / WPC][rank * WPC + i % WPC] = S.myw[i / WPC][i % WPC];
        }
        IR_TICK(0);
        cluster.sync();                                        // (A) bitmaps are at the leader
        IR_TICK(1);
        // ---------------- phase 2: leader replays the batch
        if (leader) {
            int commit = Kb, push_forced = 0, push_normal = 0, reset_noise = 0;
            if (type == 1) {
                // confirm quietness: no valid bin may cross (the mask is all-free: n_act == 0)
                if (tid == 0) S.ff_frame = Kb;
                __syncthreads();
                int mine = Kb;
                for (int w = tid; w < W; w += CT) {
                    const uint32_t v = S.valid[w];
                    for (int j = 0; j < mine; j++)
                        if (S.words[j][w] & v) { mine = j; break; }
                }
                mine = __reduce_min_sync(0xffffffffu, mine);
                if (lane == 0 && mine < Kb) atomicMin(&S.ff_frame, mine);
                __syncthreads();
                const int first = S.ff_frame;
                commit = first;
                if (tid == 0) {
                    // create_new_bursts' else-branch runs on every primed frame (:628-631)
                    int prim = primed0, idx = idx0;
                    for (int j = 0; j < commit; j++) {
                        if (prim && S.squelch_count > 0) S.squelch_count--;
                        if (++idx == c.hist_size) { prim = 1; idx = 0; }
                    }
                    S.ctl_next_type = commit < Kb ? 0 : 1;
                }
            } else {
                uint64_t fidx = index;
                tsub = clock64();
                for (int j = 0; j < Kb; j++, fidx += (uint64_t)N) {
                    // Fast-forward: warp 0 alone walks the frames on which nothing happens (no
                    // candidate peak, no burst ending, some burst still active), applying their
                    // only effects (hysteresis refresh, squelch count-down), and stops at the
                    // first frame that needs the full machinery.
                    // The search is parallel: one thread per bitmap word looks for the first frame
                    // with an unmasked valid crossing, one thread per active burst replays its
                    // hysteresis to find the frame on which it ends; the earliest wins.
                    if (tid == 0) S.ff_frame = Kb;
                    __syncthreads();
                    const int na_ff = S.n_act;
                    // frames [j, Kb) of the batch as a bit range (KB == 32 == warp width)
                    const uint32_t range = (Kb >= 32 ? 0xffffffffu : ((1u << Kb) - 1u)) & ~((1u << j) - 1u);
                    {
                        int mine = na_ff == 0 ? j : Kb;
                        // candidates: one thread per word reads that word of all 32 frames (independent
                        // loads, fully pipelined) and builds the mask of frames with an eligible bit
                        for (int w = tid; w < W; w += CT) {
                            const uint32_t mk = S.free_mask[w] & S.valid[w];
                            if (mk) {
                                uint32_t fm = 0;
#pragma unroll
                                for (int jj = 0; jj < KB; jj++) fm |= ((S.words[jj][w] & mk) != 0u) ? (1u << jj) : 0u;
                                fm &= range;
                                if (fm) mine = min(mine, __ffs(fm) - 1);
                            }
                        }
                        // burst endings: one warp per burst, lane = frame.  H = frames with a
                        // hysteresis hit; the frame's last_active is the latest hit at or before it.
                        for (int i = warp; i < na_ff; i += CT / 32) {
                            const ActBurst b = S.act[i];
                            const int cb = b.center_bin;
                            const uint32_t *Wf = S.words[lane];
                            const bool inr = (range >> lane) & 1u;
                            const bool hit = inr && ((cb > 0 && cbit(Wf, cb - 1)) || cbit(Wf, cb) || (cb < N - 1 && cbit(Wf, cb + 1)));
                            const uint32_t H = __ballot_sync(0xffffffffu, hit);
                            const uint32_t Hle = H & (0xffffffffu >> (31 - lane));
                            const uint64_t fi = index + (uint64_t)lane * (uint64_t)N;
                            const uint64_t la = Hle ? index + (uint64_t)(31 - __clz(Hle)) * (uint64_t)N : b.last_active;
                            const bool too_long = c.max_burst_len > 0 && la - b.start > (uint64_t)c.max_burst_len;
                            const bool done = inr && ((la + (uint64_t)c.post_len <= fi) || too_long);
                            const uint32_t D = __ballot_sync(0xffffffffu, done);
                            if (D) mine = min(mine, __ffs(D) - 1);
                        }
                        mine = __reduce_min_sync(0xffffffffu, mine);
                        if (lane == 0 && mine < Kb) atomicMin(&S.ff_frame, mine);
                    }
                    __syncthreads();
                    {
                        const int jn = S.ff_frame;
                        // effects of the skipped frames [j, jn): hysteresis refresh, squelch count-down
                        const uint32_t skipped = range & (jn >= 32 ? 0xffffffffu : ((1u << jn) - 1u));
                        for (int i = warp; i < na_ff; i += CT / 32) {
                            const int cb = S.act[i].center_bin;
                            const uint32_t *Wf = S.words[lane];
                            const bool hit = ((skipped >> lane) & 1u) &&
                                             ((cb > 0 && cbit(Wf, cb - 1)) || cbit(Wf, cb) || (cb < N - 1 && cbit(Wf, cb + 1)));
                            const uint32_t H = __ballot_sync(0xffffffffu, hit);
                            if (lane == 0 && H) S.act[i].last_active = index + (uint64_t)(31 - __clz(H)) * (uint64_t)N;
                        }
                        if (tid == 0 && primed) {
                            const int d = jn - j;
                            S.squelch_count = S.squelch_count > d ? S.squelch_count - d : 0;
                        }
                        fidx += (uint64_t)(jn - j) * (uint64_t)N;
                        j = jn;
                    }
                    IR_SUB(0);
                    if (j >= Kb) break;
                    IR_COUNT(sub[5] += 1);
                    const uint32_t *Wd = S.words[j];
                    if (tid == 0) S.flags = 0;
                    __syncthreads();
                    int n_act = S.n_act;
                    int fl = 0;
                    // update_bursts (:458-469)
                    for (int i = tid; i < n_act; i += CT) {
                        ActBurst &b = S.act[i];
                        const int cb = b.center_bin;
                        bool hit = (cb > 0 && cbit(Wd, cb - 1)) || cbit(Wd, cb) || (cb < N - 1 && cbit(Wd, cb + 1));
                        if (hit) b.last_active = fidx;
                        bool too_long = c.max_burst_len > 0 && b.last_active - b.start > (uint64_t)c.max_burst_len;
                        bool done = (b.last_active + (uint64_t)c.post_len <= fidx) || too_long;
                        if (done) fl |= 2;
                        if (too_long) fl |= 4;
                    }
                    // peaks: above & mask of the previous frame & search range (:522-548)
                    for (int w = tid; w < W; w += CT) {
                        uint32_t cw = Wd[w] & S.free_mask[w] & S.valid[w];
                        S.cand[w] = cw;
                        if (cw) fl |= 1;
                    }
                    if (fl) atomicOr(&S.flags, fl);
                    __syncthreads();
                    const int flags = S.flags;
                    bool forced = false;
                    IR_SUB(1);
                    if (flags & 2) {                           // delete_gone_bursts (:490-518)
                        if (tid == 0) {
                            int k = 0;
                            for (int i = 0; i < n_act; i++) {
                                ActBurst b = S.act[i];
                                bool too_long = c.max_burst_len > 0 && b.last_active - b.start > (uint64_t)c.max_burst_len;
                                if ((b.last_active + (uint64_t)c.post_len <= fidx) || too_long) cgone(S, gone, gone_cap, b, fidx);
                                else S.act[k++] = b;
                            }
                            S.n_act = k;
                        }
                        __syncthreads();
                        n_act = S.n_act;
                        forced = (flags & 4) != 0;             // update_filters_post(d, 1): applied by the owners
                        for (int w = tid; w < W; w += CT) S.free_mask[w] = 0xffffffffu;
                        __syncthreads();
                        for (int i = tid; i < n_act; i += CT)
                            cclear(S.free_mask, max(S.act[i].center_bin - c.half_bw, 0), min(S.act[i].center_bin + c.half_bw, N - 1));
                        __syncthreads();
                    }
                    // A forced update (too-long burst) changes the baseline between peak extraction
                    // and create_new_bursts: peaks keep their pre-update relative magnitude, the
                    // noise field reads the updated sum (:583).  The leader recomputes that one
                    // value per new burst itself; the owners apply the update after the batch.
                    IR_SUB(2);
                    IR_COUNT(if (flags & 1) sub[6] += 1);
                    IR_COUNT(if (flags & 2) sub[7] += 1);
                    bool created_fast = false;
                    if (flags & 1) {
                        // Gather the candidate peaks once (bin, relative magnitude, baseline the
                        // noise field will read) so that the greedy strongest-first selection runs
                        // out of shared memory in a single warp.
                        const float *row = rows + (size_t)j * N;
                        const float *hold = hist + (size_t)hist_idx * N;
                        // one candidate per thread, so that their global reads overlap: list the
                        // bins first (cheap, shared memory only), then fetch in parallel
                        if (tid == 0) S.n_cand = 0;
                        __syncthreads();
                        for (int w = tid; w < W; w += CT) {
                            uint32_t cw = S.cand[w];
                            if (cw) {
                                int slot = atomicAdd(&S.n_cand, __popc(cw));
                                while (cw) {
                                    const int b = __ffs(cw) - 1;
                                    cw &= cw - 1;
                                    if (slot < MAXC) S.cbin[slot] = (w << 5) + b;
                                    slot++;
                                }
                            }
                        }
                        __syncthreads();
                        const int nc = S.n_cand;
                        for (int i = tid; i < nc && i < MAXC; i += CT) {
                            const int bin = S.cbin[i];
                            const float bs = base_g[bin], mv = row[bin];
                            float bc = bs;
                            if (forced) {
                                const float old = primed ? hold[bin] : 0.0f;
                                const float v = bs - old;
                                bc = v + mv;
                            }
                            S.crel[i] = mv / bs;
                            S.cbase[i] = bc;
                        }
                        __syncthreads();
                        if (nc <= MAXC) {
                            created_fast = true;
                            if (warp == 0) {
                                for (;;) {
                                    ArgMax best{-1.0f, 0x7fffffff};
                                    int bslot = -1;
                                    for (int i = lane; i < nc; i += 32) {
                                        const int bin = S.cbin[i];
                                        if (bin >= 0) {
                                            const ArgMax cur{S.crel[i], bin};
                                            const ArgMax nb = argmax_pick(best, cur);
                                            if (nb.i != best.i) bslot = i;
                                            best = nb;
                                        }
                                    }
                                    const ArgMax wbest = warp_argmax(best);
                                    if (wbest.v < 0.0f) break;
                                    const int bin = wbest.i;
                                    // the lane that holds the winner publishes its baseline
                                    const unsigned owner = __ballot_sync(0xffffffffu, best.i == bin && bslot >= 0);
                                    const int src = __ffs(owner) - 1;
                                    const float bc = __shfl_sync(0xffffffffu, bslot >= 0 ? S.cbase[bslot] : 0.0f, src);
                                    if (lane == 0) {
                                        const int slot = S.n_act;
                                        if (slot < IR_MAX_ACTIVE) {
                                            ActBurst nb;
                                            nb.id = S.next_id;
                                            nb.start = fidx - (uint64_t)c.pre_len;
                                            nb.last_active = nb.start;
                                            nb.center_bin = bin;
                                            nb.peak_rel = wbest.v;
                                            nb.base_at_create = bc;
                                            nb.pad = 0;
                                            S.act[slot] = nb;
                                            S.n_act = slot + 1;
                                        } else {
                                            S.overflow = 1;
                                        }
                                        S.next_id += 10;
                                        cclear(S.free_mask, max(bin - c.half_bw, 0), min(bin + c.half_bw, N - 1));
                                    }
                                    for (int i = lane; i < nc; i += 32) {
                                        const int bb = S.cbin[i];
                                        if (bb >= bin - c.half_bw && bb <= bin + c.half_bw) S.cbin[i] = -1;
                                    }
                                    __syncwarp();
                                }
                            }
                            __syncthreads();
                        }
                    }
                    if ((flags & 1) && !created_fast) {         // create_new_bursts (:556-591), overflow path
                        const float *row = rows + (size_t)j * N;
                        const float *hold = hist + (size_t)hist_idx * N;
                        for (;;) {
                            ArgMax best{-1.0f, 0x7fffffff};
                            for (int w = tid; w < W; w += CT) {
                                uint32_t cw = S.cand[w];
                                while (cw) {
                                    const int b = __ffs(cw) - 1;
                                    cw &= cw - 1;
                                    const int bin = (w << 5) + b;
                                    best = argmax_pick(best, ArgMax{row[bin] / base_g[bin], bin});
                                }
                            }
                            best = block_argmax(best, S.red);
                            if (best.v < 0.0f) break;
                            const int bin = best.i;
                            if (tid == 0) {
                                const int slot = S.n_act;
                                if (slot < IR_MAX_ACTIVE) {
                                    ActBurst nb;
                                    nb.id = S.next_id;
                                    nb.start = fidx - (uint64_t)c.pre_len;
                                    nb.last_active = nb.start;
                                    nb.center_bin = bin;
                                    nb.peak_rel = best.v;
                                    float bs = base_g[bin];
                                    if (forced) {
                                        const float old = primed ? hold[bin] : 0.0f;
                                        const float v = bs - old;
                                        bs = v + row[bin];
                                    }
                                    nb.base_at_create = bs;
                                    nb.pad = 0;
                                    S.act[slot] = nb;
                                    S.n_act = slot + 1;
                                } else {
                                    S.overflow = 1;
                                }
                                S.next_id += 10;
                                const int lo = max(bin - c.half_bw, 0), hi = min(bin + c.half_bw, N - 1);
                                cclear(S.free_mask, lo, hi);
                                cclear(S.cand, lo, hi);
                            }
                            __syncthreads();
                        }
                    }
                    IR_SUB(3);
                    // squelch (:593-631)
                    n_act = S.n_act;
                    {
                        if (c.max_bursts > 0 && n_act > c.max_bursts) {
                            __syncthreads();
                            if (tid == 0) {
                                for (int i = 0; i < n_act; i++) {
                                    const ActBurst &b = S.act[i];
                                    if (b.start != fidx - (uint64_t)c.pre_len) cgone(S, gone, gone_cap, b, fidx);
                                }
                                S.n_act = 0;
                                S.n_squelch++;
                                S.squelch_count += 3;
                                if (S.squelch_count >= 10) { S.squelch_count = 0; S.flags |= 8; }
                            }
                            for (int w = tid; w < W; w += CT) S.free_mask[w] = 0xffffffffu;
                            __syncthreads();
                            if (S.flags & 8) reset_noise = 1;
                        } else if (tid == 0 && S.squelch_count > 0) {
                            S.squelch_count--;
                        }
                    }
                    __syncthreads();
                    // does this frame end the batch?  any baseline update does.
                    IR_SUB(4);
                    const bool quiet_after = S.n_act == 0;
                    if (forced || quiet_after || reset_noise) {
                        commit = j + 1;
                        push_forced = forced ? 1 : 0;
                        push_normal = quiet_after ? 1 : 0;
                        break;
                    }
                }
                if (tid == 0) S.ctl_next_type = push_normal ? 1 : 0;
            }
            if (tid == 0) {
                S.ctl_commit = commit;
                S.ctl_push_forced = push_forced;
                S.ctl_push_normal = push_normal;
                S.ctl_reset_noise = reset_noise;
            }
        }
        IR_TICK(2);
        cluster.sync();                                        // (B) verdict is published
        IR_TICK(3);
        const int commit = LS.ctl_commit;
        const int pushF = LS.ctl_push_forced, pushN = LS.ctl_push_normal, rst = LS.ctl_reset_noise;
        const int next_type = LS.ctl_next_type;
        // ---------------- phase 3: owners make their state match the verdict
        bool publish = false;
        if (type == 1) {
            if (commit == Kb) {
                // all quiet: the speculative baselines stand; write the history rows now
                int idx = idx0;
                for (int j0 = 0; j0 < Kb; j0 += G) {
                    float mv[G][BPT];
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        const float *row = rows + (size_t)(j0 + g) * N + bin0;
#pragma unroll
                        for (int u = 0; u < BPT; u++) mv[g][u] = (j0 + g < Kb) ? row[u * CT + tid] : 0.0f;
                    }
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        if (j0 + g < Kb) {
                            float *h = hist + (size_t)idx * N + bin0;
#pragma unroll
                            for (int u = 0; u < BPT; u++) h[u * CT + tid] = mv[g][u];
                            if (++idx == c.hist_size) idx = 0;
                        }
                    }
                }
            } else {
                // rewind to the batch start and redo the frames that stand, this time for real
#pragma unroll
                for (int u = 0; u < BPT; u++) base[u] = base0[u];
                hist_idx = idx0; primed = primed0;
                for (int j0 = 0; j0 < commit; j0 += G) {
                    float mv[G][BPT], ov[G][BPT];
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        const bool in = j0 + g < commit;
                        const float *row = rows + (size_t)(j0 + g) * N + bin0;
                        int hrow = hist_idx + g;
                        if (hrow >= c.hist_size) hrow -= c.hist_size;
                        const float *h = hist + (size_t)hrow * N + bin0;
                        const bool live = primed || (hist_idx + g >= c.hist_size);
#pragma unroll
                        for (int u = 0; u < BPT; u++) {
                            mv[g][u] = in ? row[u * CT + tid] : 0.0f;
                            ov[g][u] = (in && live) ? h[u * CT + tid] : 0.0f;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        if (j0 + g < commit) {
                            float *h = hist + (size_t)hist_idx * N + bin0;
#pragma unroll
                            for (int u = 0; u < BPT; u++) {
                                const float v = base[u] - ov[g][u];
                                base[u] = v + mv[g][u];
                                h[u * CT + tid] = mv[g][u];
                            }
                            if (++hist_idx == c.hist_size) { primed = 1; hist_idx = 0; }
                        }
                    }
                }
                publish = true;                                // next batch is type F
            }
        } else {
            // order inside the frame: forced update (:516-517), squelch reset (:618-627), regular
            // update (:698)
            if (pushF) { push_row(rows + (size_t)(commit - 1) * N, true); publish = true; }
            if (rst) {
                hist_idx = 0; primed = 0;
#pragma unroll
                for (int u = 0; u < BPT; u++) base[u] = 0.0f;
                publish = true;
            }
            if (pushN) { push_row(rows + (size_t)(commit - 1) * N, true); publish = true; }
        }
        if (publish) {
#pragma unroll
            for (int u = 0; u < BPT; u++) base_g[bin0 + u * CT + tid] = base[u];
            __threadfence();
        }
        k0 += commit;
        index += (uint64_t)commit * (uint64_t)N;
        type = next_type;
        IR_TICK(4);
        cluster.sync();                                        // base_g / history visible before the next batch
        IR_TICK(5);
        IR_COUNT(tacc[6] += 1; tacc[7] += (unsigned long long)(type == 1));
    }

    // ---- store state
#pragma unroll
    for (int u = 0; u < BPT; u++) base_g[bin0 + u * CT + tid] = base[u];
    if (leader) {
        __syncthreads();
        for (int i = tid; i < S.n_act; i += CT) gs->act[i] = S.act[i];
        if (tid == 0) {
            gs->hist_idx = hist_idx; gs->primed = primed; gs->n_act = S.n_act;
            gs->squelch_count = S.squelch_count; gs->next_id = S.next_id; gs->index = index;
            gs->n_gone = S.n_gone; gs->n_squelch = S.n_squelch; gs->overflow = S.overflow;
            for (int i = 0; i < 8; i++) gs->dbg[i] += tacc[i];
        }
    }
    if (rank == 0 && tid == 0) for (int i = 0; i < 8; i++) gs->dbg[8 + i] += sub[i];
    if (rank == 3 && tid == 0) for (int i = 0; i < 8; i++) gs->dbg[16 + i] += p1s[i];
}

template <int BPT>
static cudaError_t launch_cluster_t(const DetConfig &c, DetState *state, float *base, float *hist,
                                    const float *mag, int64_t n_frames, GoneBurst *gone,
                                    uint32_t gone_cap, cudaStream_t st) {
    constexpr int G = BPT >= 8 ? 4 : (BPT == 4 ? 8 : 16);
    const size_t smem = ((sizeof(ClShared) + 127) / 128) * 128 + sizeof(float) * 2 * G * BPT * CT;
    cudaError_t e = cudaFuncSetAttribute(k_detect_scan_cluster<BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_detect_scan_cluster<BPT><<<CL, CT, smem, st>>>(c, state, base, hist, mag, n_frames, gone, gone_cap);
    return cudaGetLastError();
}

cudaError_t launch_detect_scan_cluster(const DetConfig &c, DetState *state, float *base, float *hist,
                                       const float *mag, int64_t n_frames, GoneBurst *gone,
                                       uint32_t gone_cap, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    switch (c.N / (CL * CT)) {
    case 1: return launch_cluster_t<1>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 2: return launch_cluster_t<2>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 4: return launch_cluster_t<4>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 8: return launch_cluster_t<8>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ir

#include <stdlib.h>
#include <string.h>
namespace ir {
cudaError_t launch_detect_scan_auto(const DetConfig &c, DetState *state, float *base, float *hist,
                                    const float *mag, int64_t n_frames, GoneBurst *gone,
                                    uint32_t gone_cap, cudaStream_t st) {
    const char *env = getenv("IR_SCAN");
    const bool force_single = env && strcmp(env, "single") == 0;
    if (c.N >= 2048 && !force_single)
        return launch_detect_scan_cluster(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    return launch_detect_scan(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
}
}  // namespace ir
