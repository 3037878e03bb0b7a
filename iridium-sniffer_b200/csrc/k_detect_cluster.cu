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
//     "above threshold" bitmaps of K frames at once; the leader CTA then works out what the
//     state machine does on those K frames.  The batch is cut after the first frame that ends
//     with a baseline update.
//   batch type Q (no burst active): the owners assume every frame is quiet, update their
//     baselines frame by frame (per-bin serial, exactly the reference's two roundings) and
//     produce the bitmaps along the way; the leader only has to confirm that no valid bin crossed
//     the threshold.  The first frame that does cuts the batch: owners rewind to that frame
//     (recompute from the batch start) and the next batch is of type F.
//
// Owners (phase 1) stream their slices straight from global memory into registers, one chunk of
// frames ahead, and test them against per-bin thresholds; only a warp that sees a crossing builds
// bitmap words (ballots) and lists them as (frame, word) entries.  Everything the leader does
// is driven by those sparse entries.
//
// Leader, type F (phase 2): the frames of a batch are NOT replayed one by one.  With a frozen
// baseline a burst's end frame depends only on its own bins (hysteresis replay: one warp per
// burst, one lane per frame), so deletions need no sequential step at all.  What is sequential is
// burst creation -- a new burst masks its neighbourhood in later frames -- so the leader iterates
// "rounds": find the earliest frame with an eligible crossing given every burst known so far,
// create that frame's bursts (strongest first), compute their end frames, repeat.  The plan is
// applied (burst list, gone records, ids, mask) only at the end; anything unusual -- squelch,
// a burst that could exceed max_burst_len, table overflow -- drops the plan and runs the plain
// frame-by-frame replay instead, which is also the cross-check path (IR_SCAN=cluster_dense).
//
// Three cluster barriers per batch instead of two CTA barriers per frame; bitmaps travel to the
// leader through distributed shared memory.  Results are identical to the single-CTA kernel in
// k_detect.cu (kept for N < 2048 and as a cross-check; tests compare all variants with the CPU
// oracle).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace cg = cooperative_groups;

namespace ir {

namespace {

constexpr int CL = 8;          // CTAs per cluster
constexpr int CT = 256;        // threads per CTA
constexpr int KB = 32;         // frames per batch
static_assert(KB == 32, "the leader maps one frame to one lane");
constexpr int MAXW = 512;      // bitmap words per frame (N <= 16384)
constexpr int MAXC = 1024;     // candidate peaks kept in shared memory per frame
constexpr int ENT_PER = 256;   // non-zero bitmap words listed per CTA and batch
constexpr int ENT_MAX = CL * ENT_PER;
constexpr int PB_MAX = CT;     // bursts a plan can hold (one thread each when it is applied)
constexpr int NOF = 64;        // "no frame"
constexpr uint32_t FULL = 0xffffffffu;

struct ClShared {
    uint32_t words[KB][MAXW];      // leader only: above-threshold bitmaps of the batch
    uint32_t myw[KB][MAXW / CL];   // every CTA: its own words of the batch, shipped once per batch
    // sparse view of the same bitmaps: (frame << 16 | word) of every non-zero word
    uint32_t my_ent[ENT_PER];      // every CTA: entries of its own words
    uint32_t ent_r[CL][ENT_PER];   // leader only: as shipped by each CTA
    uint32_t ent[ENT_MAX];         // leader only: compacted
    int my_n_ent;
    int n_ent_r[CL];
    int fbits[KB];                 // leader: set valid bits per frame (bound on candidate peaks)
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
    // plan of a type-F batch
    int pb_cb[PB_MAX];             // center bin
    int pb_create[PB_MAX];         // frame of creation, -1 = active before the batch
    int pb_end[PB_MAX];            // frame of deletion, NOF = survives the batch
    uint32_t pb_hm[PB_MAX];        // frames with a hysteresis hit
    float nb_rel[PB_MAX];          // new bursts: peak relative magnitude, baseline at creation
    float nb_base[PB_MAX];
    int pl_rank[PB_MAX];
    int fstar, round_new, plan_abort;
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

// bits of word w that lie in [lo, hi] (bin numbers)
__device__ __forceinline__ uint32_t range_bits(int w, int lo, int hi) {
    int a = max(lo, w << 5), b = min(hi, (w << 5) + 31);
    if (a > b) return 0u;
    a &= 31; b &= 31;
    return (b == 31 ? FULL : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
}

__device__ __forceinline__ void cclear(uint32_t *bm, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi >> 5); w++) atomicAnd(&bm[w], ~range_bits(w, lo, hi));
}

__device__ __forceinline__ bool hyst_hit(const uint32_t *Wf, int cb, int N) {
    return (cb > 0 && cbit(Wf, cb - 1)) || cbit(Wf, cb) || (cb < N - 1 && cbit(Wf, cb + 1));
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
                      GoneBurst *__restrict__ gone, uint32_t gone_cap, int force_dense, const int *run_if,
                      ScanSnapshot snap) {
    if (run_if != nullptr && *run_if == 0) return;            // uniform over the cluster: nobody reaches a barrier
    cg::cluster_group cluster = cg::this_cluster();
    if (run_if != nullptr && snap.undo != nullptr) {
        // fallback of a bailed streaming launch: first put back what its baseline updates overwrote
        // (latest first, so that a history row written twice ends with its oldest value) and the
        // baseline it started from; the state block was never touched by it
        const size_t stride = (size_t)CL * CT, t = (size_t)cluster.block_rank() * CT + threadIdx.x;
        const int nq = snap.ctl->undo_frames, h0 = gs->hist_idx, H = c.hist_size;
        for (size_t bin = t; bin < (size_t)c.N; bin += stride) {
            for (int q = nq - 1; q >= 0; q--) hist[(size_t)((h0 + q) % H) * c.N + bin] = snap.undo[(size_t)q * c.N + bin];
            base_g[bin] = snap.base[bin];
        }
        __threadfence();
        cluster.sync();
    }
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
            S.free_mask[w] = FULL;
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
    // rel = mag/base > thr (IEEE divide, simd_avx2.c:239-257).  Outside the band
    // base*thr*(1 -+ 1e-5) the outcome is decided by a product; the divide runs only inside it.
    auto above = [&](float mv, float bs) -> bool {
        bool ab = false;
        if (bs > 0.0f) {
            if (mv > bs * thr_hi) ab = true;
            else if (mv > bs * thr_lo) ab = mv / bs > thr;
        }
        return ab;
    };
    // one bitmap word (32 consecutive bins of this warp) of frame j: store + list it if non-zero
    auto put_word = [&](int j, int u, bool ab) {
        const uint32_t b = __ballot_sync(FULL, ab);
        if (lane == 0 && b) {
            S.myw[j][u * (CT / 32) + warp] = b;
            const int s = atomicAdd(&S.my_n_ent, 1);
            if (s < ENT_PER) S.my_ent[s] = ((uint32_t)j << 16) | (uint32_t)(rank * WPC + u * (CT / 32) + warp);
        }
    };
    constexpr int G = BPT >= 8 ? 4 : 8;                       // frames per register chunk

    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long sub[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // leader phase-2 breakdown
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
        long long tq = clock64();
        if (tid == 0) S.my_n_ent = 0;
        {
            uint4 *mz = reinterpret_cast<uint4 *>(&S.myw[0][0]);
            for (int i = tid; i < KB * (MAXW / CL) / 4; i += CT) mz[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        auto load_chunk = [&](float (&dst)[G][BPT], int j0) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                const float *row = rows + (size_t)(j0 + g) * N + bin0 + tid;
#pragma unroll
                for (int u = 0; u < BPT; u++) dst[g][u] = (j0 + g < Kb) ? row[u * CT] : 0.0f;
            }
        };
        float lim[BPT];                                       // type F: per-bin pre-screen threshold
#pragma unroll
        for (int u = 0; u < BPT; u++) lim[u] = base[u] > 0.0f ? base[u] * thr_lo : INF;
        float nx[G][BPT];
        load_chunk(nx, 0);
        IR_P1(0);
#pragma unroll
        for (int j0 = 0; j0 < KB; j0 += G) {
            if (j0 >= Kb) break;
            float mv[G][BPT];
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int u = 0; u < BPT; u++) mv[g][u] = nx[g][u];
            if (j0 + G < Kb) load_chunk(nx, j0 + G);          // next chunk in flight while this one is used
            if (type == 0) {
                // frozen baseline: frames (bits) on which this thread's bins pass the pre-screen
                uint32_t fm[BPT];
                uint32_t anyb = 0;
#pragma unroll
                for (int u = 0; u < BPT; u++) {
                    fm[u] = 0;
#pragma unroll
                    for (int g = 0; g < G; g++) fm[u] |= (mv[g][u] > lim[u]) ? (1u << g) : 0u;
                    anyb |= fm[u];
                }
                if (__any_sync(FULL, anyb != 0u)) {              // only where a burst sits
#pragma unroll
                    for (int u = 0; u < BPT; u++) {
                        const uint32_t fr = __reduce_or_sync(FULL, fm[u]);
#pragma unroll
                        for (int g = 0; g < G; g++)
                            if ((fr >> g) & 1u) put_word(j0 + g, u, above(mv[g][u], base[u]));
                    }
                }
            } else {
                float ov[G][BPT];
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const bool in = j0 + g < Kb;
                    int hrow = hist_idx + g;
                    if (hrow >= c.hist_size) hrow -= c.hist_size;
                    const float *h = hist + (size_t)hrow * N + bin0 + tid;
                    // a history row is live if the detector is primed now or wraps before reaching it
                    const bool live = primed || (hist_idx + g >= c.hist_size);
#pragma unroll
                    for (int u = 0; u < BPT; u++) ov[g][u] = (in && live) ? h[u * CT] : 0.0f;
                }
                // speculative pass: update the baselines frame by frame, note pre-screen passes
                float bs0[BPT];
#pragma unroll
                for (int u = 0; u < BPT; u++) bs0[u] = base[u];
                const int hi0 = hist_idx, pr0 = primed;
                uint32_t anyb = 0;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (j0 + g < Kb) {
#pragma unroll
                        for (int u = 0; u < BPT; u++) {
                            // (a non-positive baseline passes here and is sorted out by the exact test)
                            if (primed) anyb |= (mv[g][u] > base[u] * thr_lo) ? 1u : 0u;
                            const float v = base[u] - ov[g][u];
                            base[u] = v + mv[g][u];
                        }
                        if (++hist_idx == c.hist_size) { primed = 1; hist_idx = 0; }
                    }
                }
                if (__any_sync(FULL, anyb != 0u)) {              // rare: redo this chunk with the exact test
#pragma unroll
                    for (int u = 0; u < BPT; u++) base[u] = bs0[u];
                    hist_idx = hi0; primed = pr0;
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        if (j0 + g < Kb) {
                            if (primed) {
#pragma unroll
                                for (int u = 0; u < BPT; u++) {
                                    const bool pass = base[u] > 0.0f ? (mv[g][u] > base[u] * thr_lo) : false;
                                    if (__any_sync(FULL, pass)) put_word(j0 + g, u, above(mv[g][u], base[u]));
                                }
                            }
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
        }
        IR_P1(1);
        // ship this CTA's words and entries of the whole batch to the leader (DSMEM)
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
        {
            const int my_n = S.my_n_ent;
            for (int i = tid; i < my_n && i < ENT_PER; i += CT) LS.ent_r[rank][i] = S.my_ent[i];
            if (tid == 0) LS.n_ent_r[rank] = my_n;
        }
        IR_P1(2);
        IR_TICK(0);
        cluster.sync();                                        // (A) bitmaps are at the leader
        IR_TICK(1);
        // ---------------- phase 2: leader
        if (leader) {
            int commit = Kb, push_forced = 0, push_normal = 0, reset_noise = 0;
            // compact the per-CTA lists of non-zero words; sparse = they are complete
            bool sparse = true;
            int n_ent = 0;
#pragma unroll
            for (int r = 0; r < CL; r++) sparse = sparse && S.n_ent_r[r] <= ENT_PER;
            if (sparse) {
#pragma unroll
                for (int r = 0; r < CL; r++) {
                    const int nr = S.n_ent_r[r];
                    for (int i = tid; i < nr; i += CT) S.ent[n_ent + i] = S.ent_r[r][i];
                    n_ent += nr;
                }
            }
            if (type == 1) {
                // confirm quietness: no valid bin may cross (the mask is all-free: n_act == 0)
                if (tid == 0) S.ff_frame = Kb;
                __syncthreads();
                int mine = Kb;
                if (sparse) {
                    for (int e = tid; e < n_ent; e += CT) {
                        const uint32_t en = S.ent[e];
                        const int ej = (int)(en >> 16), w = (int)(en & 0xffffu);
                        if (ej < mine && (S.words[ej][w] & S.valid[w])) mine = ej;
                    }
                } else {
                    for (int w = tid; w < W; w += CT) {
                        const uint32_t v = S.valid[w];
                        for (int j = 0; j < mine; j++)
                            if (S.words[j][w] & v) { mine = j; break; }
                    }
                }
                mine = __reduce_min_sync(FULL, mine);
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
                tsub = clock64();
                // ======== plan: rounds over burst creations (see the header)
                bool planned = false;
                {
                    const int n_old = S.n_act;
                    bool ok = sparse && primed && !force_dense && n_old <= PB_MAX - 64 &&
                              (c.max_burst_len <= 0 ||
                               (uint64_t)Kb * (uint64_t)N + (uint64_t)c.pre_len <= (uint64_t)c.max_burst_len);
                    const uint64_t batch_end = index + (uint64_t)Kb * (uint64_t)N;
                    int bad = 0;
                    if (ok && c.max_burst_len > 0)
                        for (int i = tid; i < n_old; i += CT)
                            if (batch_end - S.act[i].start > (uint64_t)c.max_burst_len) bad = 1;   // could become too long
                    if (tid < KB) S.fbits[tid] = 0;
                    if (tid == 0) { S.plan_abort = 0; S.round_new = 0; }
                    __syncthreads();                              // also orders the compaction of S.ent
                    if (ok)
                        for (int e = tid; e < n_ent; e += CT) {
                            const uint32_t en = S.ent[e];
                            const int ej = (int)(en >> 16), w = (int)(en & 0xffffu);
                            atomicAdd(&S.fbits[ej], __popc(S.words[ej][w] & S.valid[w]));
                        }
                    __syncthreads();
                    if (tid < KB && S.fbits[tid] > MAXC) bad = 1;  // a frame could overflow the peak list
                    ok = !__syncthreads_or(bad) && ok;
                    IR_SUB(0);
                    if (ok) {
                        const uint32_t *Wf = S.words[lane];       // lane = frame
                        const uint64_t fi = index + (uint64_t)lane * (uint64_t)N;
                        // bursts active before the batch: hysteresis replay -> end frame
                        for (int i = warp; i < n_old; i += CT / 32) {
                            const int cb = S.act[i].center_bin;
                            const uint64_t b_la = S.act[i].last_active;
                            const bool inr = lane < Kb;
                            const uint32_t H = __ballot_sync(FULL, inr && hyst_hit(Wf, cb, N));
                            const uint32_t Hle = H & (FULL >> (31 - lane));
                            const uint64_t la = Hle ? index + (uint64_t)(31 - __clz(Hle)) * (uint64_t)N : b_la;
                            const uint32_t D = __ballot_sync(FULL, inr && (la + (uint64_t)c.post_len <= fi));
                            if (lane == 0) {
                                S.pb_cb[i] = cb; S.pb_create[i] = -1; S.pb_end[i] = D ? __ffs(D) - 1 : NOF; S.pb_hm[i] = H;
                            }
                        }
                        int n_pb = n_old, f_done = -1, n_new = 0, q = NOF;
                        bool abort_plan = false;
                        __syncthreads();
                        IR_SUB(1);
                        for (;;) {
                            IR_COUNT(sub[9] += 1);
                            if (tid == 0) { S.fstar = NOF; S.n_cand = 0; }
                            // first frame after which no burst is left (every warp works it out)
                            bool alive = lane >= Kb;
                            for (int k = 0; k < n_pb; k++) alive = alive || (S.pb_create[k] <= lane && lane < S.pb_end[k]);
                            const uint32_t am = __ballot_sync(FULL, alive);
                            q = (~am) ? __ffs(~am) - 1 : NOF;
                            const int flim = min(q, Kb - 1);
                            __syncthreads();
                            // earliest frame after f_done with a crossing no known burst masks.  A burst
                            // masks the peaks of frames (create, end]: the mask is the previous frame's.
                            for (int e = tid; e < n_ent; e += CT) {
                                const uint32_t en = S.ent[e];
                                const int f = (int)(en >> 16), w = (int)(en & 0xffffu);
                                if (f > f_done && f <= flim) {
                                    uint32_t m = S.words[f][w] & S.valid[w];
                                    for (int k = 0; k < n_pb && m; k++)
                                        if (S.pb_create[k] < f && f <= S.pb_end[k])
                                            m &= ~range_bits(w, S.pb_cb[k] - c.half_bw, S.pb_cb[k] + c.half_bw);
                                    if (m) atomicMin(&S.fstar, f);
                                }
                            }
                            __syncthreads();
                            const int fs = S.fstar;
                            IR_SUB(2);
                            if (fs >= NOF) break;
                            // peaks of frame fs (:522-548)
                            for (int e = tid; e < n_ent; e += CT) {
                                const uint32_t en = S.ent[e];
                                const int f = (int)(en >> 16), w = (int)(en & 0xffffu);
                                if (f == fs) {
                                    uint32_t m = S.words[f][w] & S.valid[w];
                                    for (int k = 0; k < n_pb && m; k++)
                                        if (S.pb_create[k] < f && f <= S.pb_end[k])
                                            m &= ~range_bits(w, S.pb_cb[k] - c.half_bw, S.pb_cb[k] + c.half_bw);
                                    if (m) {
                                        int slot = atomicAdd(&S.n_cand, __popc(m));
                                        while (m) {
                                            const int b = __ffs(m) - 1;
                                            m &= m - 1;
                                            if (slot < MAXC) S.cbin[slot] = (w << 5) + b;
                                            slot++;
                                        }
                                    }
                                }
                            }
                            __syncthreads();
                            const int nc = min(S.n_cand, MAXC);
                            IR_SUB(3);
                            {
                                const float *row = rows + (size_t)fs * N;
                                for (int i = tid; i < nc; i += CT) {
                                    const int bin = S.cbin[i];
                                    const float bs = base_g[bin];
                                    S.crel[i] = row[bin] / bs;
                                    S.cbase[i] = bs;
                                }
                            }
                            __syncthreads();
                            IR_SUB(4);
                            if (warp == 0) {                      // create_new_bursts (:556-591): strongest first
                                int t = 0;
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
                                    const unsigned owner = __ballot_sync(FULL, best.i == bin && bslot >= 0);
                                    const int src = __ffs(owner) - 1;
                                    const float bc = __shfl_sync(FULL, bslot >= 0 ? S.cbase[bslot] : 0.0f, src);
                                    if (n_pb + t >= PB_MAX) { if (lane == 0) S.plan_abort = 1; break; }
                                    if (lane == 0) {
                                        const int k = n_pb + t;
                                        S.pb_cb[k] = bin; S.pb_create[k] = fs; S.pb_end[k] = NOF; S.pb_hm[k] = 0u;
                                        S.nb_rel[n_new + t] = wbest.v; S.nb_base[n_new + t] = bc;
                                    }
                                    t++;
                                    for (int i = lane; i < nc; i += 32) {
                                        const int bb = S.cbin[i];
                                        if (bb >= bin - c.half_bw && bb <= bin + c.half_bw) S.cbin[i] = -1;
                                    }
                                    __syncwarp();
                                }
                                if (lane == 0) S.round_new = t;
                            }
                            __syncthreads();
                            const int t_new = S.round_new;
                            IR_SUB(5);
                            if (S.plan_abort) { abort_plan = true; break; }
                            // end frames of the new bursts: they are updated from the next frame on
                            for (int t = warp; t < t_new; t += CT / 32) {
                                const int k = n_pb + t;
                                const int cb = S.pb_cb[k];
                                const bool inr = lane > fs && lane < Kb;
                                const uint32_t H = __ballot_sync(FULL, inr && hyst_hit(Wf, cb, N));
                                const uint32_t Hle = H & (FULL >> (31 - lane));
                                const uint64_t start = index + (uint64_t)fs * (uint64_t)N - (uint64_t)c.pre_len;
                                const uint64_t la = Hle ? index + (uint64_t)(31 - __clz(Hle)) * (uint64_t)N : start;
                                const uint32_t D = __ballot_sync(FULL, inr && (la + (uint64_t)c.post_len <= fi));
                                if (lane == 0) { S.pb_end[k] = D ? __ffs(D) - 1 : NOF; S.pb_hm[k] = H; }
                            }
                            __syncthreads();
                            n_pb += t_new; n_new += t_new; f_done = fs;
                            if (c.max_bursts > 0) {               // squelch (:593-631) is left to the replay
                                int cnt = 0;
                                for (int k = lane; k < n_pb; k += 32) cnt += (S.pb_create[k] <= fs && fs < S.pb_end[k]) ? 1 : 0;
                                cnt = __reduce_add_sync(FULL, cnt);
                                if (cnt > c.max_bursts) { abort_plan = true; break; }
                            }
                            if (n_pb > PB_MAX - 64) { abort_plan = true; break; }
                            IR_SUB(6);
                        }
                        if (!abort_plan) {
                            // ---- apply the plan: one thread per burst
                            planned = true;
                            const int C = q < Kb ? q + 1 : Kb;    // the batch is cut after the first quiet frame
                            commit = C;
                            push_normal = q < Kb ? 1 : 0;
                            ActBurst mine;
                            const bool have = tid < n_pb;
                            int endf = NOF;
                            bool ended = false;
                            if (have) {
                                endf = S.pb_end[tid];
                                const int cr = S.pb_create[tid];
                                if (cr < 0) {
                                    mine = S.act[tid];
                                } else {
                                    const int t = tid - n_old;
                                    mine.id = S.next_id + 10ull * (unsigned long long)t;
                                    mine.start = index + (uint64_t)cr * (uint64_t)N - (uint64_t)c.pre_len;
                                    mine.last_active = mine.start;
                                    mine.center_bin = S.pb_cb[tid];
                                    mine.peak_rel = S.nb_rel[t];
                                    mine.base_at_create = S.nb_base[t];
                                    mine.pad = 0;
                                }
                                ended = endf < C;
                                const int lastf = ended ? endf : C - 1;
                                const uint32_t hits = S.pb_hm[tid] & (lastf >= 31 ? FULL : ((1u << (lastf + 1)) - 1u));
                                if (hits) mine.last_active = index + (uint64_t)(31 - __clz(hits)) * (uint64_t)N;
                            }
                            S.pl_rank[tid] = have ? (ended ? endf : -1) : -2;
                            __syncthreads();
                            int pos = 0;
                            if (have) {
                                // gone records in the replay's order: by deletion frame, then list order;
                                // survivors keep their list order
                                for (int k = 0; k < n_pb; k++) {
                                    const int e2 = S.pl_rank[k];
                                    if (ended) pos += (e2 >= 0 && (e2 < endf || (e2 == endf && k < tid))) ? 1 : 0;
                                    else pos += (e2 == -1 && k < tid) ? 1 : 0;
                                }
                                if (ended) {
                                    const uint32_t slot = S.n_gone + (uint32_t)pos;
                                    if (slot < gone_cap) {
                                        GoneBurst g;
                                        g.id = mine.id; g.start = mine.start; g.stop = index + (uint64_t)endf * (uint64_t)N;
                                        g.last_active = mine.last_active; g.center_bin = mine.center_bin;
                                        g.peak_rel = mine.peak_rel; g.base_at_create = mine.base_at_create; g.pad = 0;
                                        gone[slot] = g;
                                    } else {
                                        S.overflow = 1;
                                    }
                                }
                            }
                            const int n_end = __syncthreads_count(have && ended);     // also: every S.act read is done
                            if (have && !ended) S.act[pos] = mine;
                            for (int w = tid; w < W; w += CT) S.free_mask[w] = FULL;
                            if (tid == 0) {
                                S.n_act = n_pb - n_end;
                                S.n_gone += (uint32_t)n_end;
                                S.next_id += 10ull * (unsigned long long)n_new;
                                S.squelch_count = max(S.squelch_count - C, 0);
                            }
                            __syncthreads();
                            if (have && !ended)
                                cclear(S.free_mask, max(mine.center_bin - c.half_bw, 0), min(mine.center_bin + c.half_bw, N - 1));
                            __syncthreads();
                        }
                    }
                }
                IR_SUB(7);
                IR_COUNT(sub[planned ? 10 : 11] += 1);
                if (!planned) {
                // ======== plain replay, frame by frame (every case)
                uint64_t fidx = index;
                for (int j = 0; j < Kb; j++, fidx += (uint64_t)N) {
                    // Fast-forward over the frames on which nothing happens (no candidate peak, no
                    // burst ending, some burst still active), applying their only effects (hysteresis
                    // refresh, squelch count-down).  The search is parallel: one thread per bitmap word
                    // looks for the first frame with an unmasked valid crossing, one warp per active
                    // burst replays its hysteresis to find the frame on which it ends; the earliest wins.
                    if (tid == 0) S.ff_frame = Kb;
                    __syncthreads();
                    const int na_ff = S.n_act;
                    // frames [j, Kb) of the batch as a bit range (KB == 32 == warp width)
                    const uint32_t range = (Kb >= 32 ? FULL : ((1u << Kb) - 1u)) & ~((1u << j) - 1u);
                    {
                        int mine = na_ff == 0 ? j : Kb;
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
                            const bool hit = inr && hyst_hit(Wf, cb, N);
                            const uint32_t H = __ballot_sync(FULL, hit);
                            const uint32_t Hle = H & (FULL >> (31 - lane));
                            const uint64_t fi = index + (uint64_t)lane * (uint64_t)N;
                            const uint64_t la = Hle ? index + (uint64_t)(31 - __clz(Hle)) * (uint64_t)N : b.last_active;
                            const bool too_long = c.max_burst_len > 0 && la - b.start > (uint64_t)c.max_burst_len;
                            const bool done = inr && ((la + (uint64_t)c.post_len <= fi) || too_long);
                            const uint32_t D = __ballot_sync(FULL, done);
                            if (D) mine = min(mine, __ffs(D) - 1);
                        }
                        mine = __reduce_min_sync(FULL, mine);
                        if (lane == 0 && mine < Kb) atomicMin(&S.ff_frame, mine);
                    }
                    __syncthreads();
                    {
                        const int jn = S.ff_frame;
                        // effects of the skipped frames [j, jn): hysteresis refresh, squelch count-down
                        const uint32_t skipped = range & (jn >= 32 ? FULL : ((1u << jn) - 1u));
                        for (int i = warp; i < na_ff; i += CT / 32) {
                            const int cb = S.act[i].center_bin;
                            const uint32_t *Wf = S.words[lane];
                            const bool hit = ((skipped >> lane) & 1u) && hyst_hit(Wf, cb, N);
                            const uint32_t H = __ballot_sync(FULL, hit);
                            if (lane == 0 && H) S.act[i].last_active = index + (uint64_t)(31 - __clz(H)) * (uint64_t)N;
                        }
                        if (tid == 0 && primed) {
                            const int d = jn - j;
                            S.squelch_count = S.squelch_count > d ? S.squelch_count - d : 0;
                        }
                        fidx += (uint64_t)(jn - j) * (uint64_t)N;
                        j = jn;
                    }
                    if (j >= Kb) break;
                    const uint32_t *Wd = S.words[j];
                    if (tid == 0) S.flags = 0;
                    __syncthreads();
                    int n_act = S.n_act;
                    int fl = 0;
                    // update_bursts (:458-469)
                    for (int i = tid; i < n_act; i += CT) {
                        ActBurst &b = S.act[i];
                        const int cb = b.center_bin;
                        if (hyst_hit(Wd, cb, N)) b.last_active = fidx;
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
                        for (int w = tid; w < W; w += CT) S.free_mask[w] = FULL;
                        __syncthreads();
                        for (int i = tid; i < n_act; i += CT)
                            cclear(S.free_mask, max(S.act[i].center_bin - c.half_bw, 0), min(S.act[i].center_bin + c.half_bw, N - 1));
                        __syncthreads();
                    }
                    // A forced update (too-long burst) changes the baseline between peak extraction
                    // and create_new_bursts: peaks keep their pre-update relative magnitude, the
                    // noise field reads the updated sum (:583).  The leader recomputes that one
                    // value per new burst itself; the owners apply the update after the batch.
                    bool created_fast = false;
                    if (flags & 1) {
                        // Gather the candidate peaks once (bin, relative magnitude, baseline the
                        // noise field will read) so that the greedy strongest-first selection runs
                        // out of shared memory in a single warp.
                        const float *row = rows + (size_t)j * N;
                        const float *hold = hist + (size_t)hist_idx * N;
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
                                    const unsigned owner = __ballot_sync(FULL, best.i == bin && bslot >= 0);
                                    const int src = __ffs(owner) - 1;
                                    const float bc = __shfl_sync(FULL, bslot >= 0 ? S.cbase[bslot] : 0.0f, src);
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
                            for (int w = tid; w < W; w += CT) S.free_mask[w] = FULL;
                            __syncthreads();
                            if (S.flags & 8) reset_noise = 1;
                        } else if (tid == 0 && S.squelch_count > 0) {
                            S.squelch_count--;
                        }
                    }
                    __syncthreads();
                    // does this frame end the batch?  any baseline update does.
                    const bool quiet_after = S.n_act == 0;
                    if (forced || quiet_after || reset_noise) {
                        commit = j + 1;
                        push_forced = forced ? 1 : 0;
                        push_normal = quiet_after ? 1 : 0;
                        break;
                    }
                }
                }
                IR_SUB(8);
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
    if (rank == 0 && tid == 0) for (int i = 0; i < 12; i++) gs->dbg[8 + i] += sub[i];
    if (rank == 3 && tid == 0) for (int i = 0; i < 4; i++) gs->dbg[20 + i] += p1s[i];
}

template <int BPT>
static cudaError_t launch_cluster_t(const DetConfig &c, DetState *state, float *base, float *hist,
                                    const float *mag, int64_t n_frames, GoneBurst *gone,
                                    uint32_t gone_cap, const int *run_if, const ScanSnapshot &snap, cudaStream_t st) {
    const size_t smem = ((sizeof(ClShared) + 127) / 128) * 128;
    cudaError_t e = cudaFuncSetAttribute(k_detect_scan_cluster<BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const char *env = getenv("IR_SCAN");
    const int force_dense = env && strcmp(env, "cluster_dense") == 0;      // cross-check path of the tests
    k_detect_scan_cluster<BPT><<<CL, CT, smem, st>>>(c, state, base, hist, mag, n_frames, gone, gone_cap, force_dense, run_if, snap);
    return cudaGetLastError();
}

cudaError_t launch_detect_scan_cluster_if(const DetConfig &c, DetState *state, float *base, float *hist,
                                          const float *mag, int64_t n_frames, GoneBurst *gone,
                                          uint32_t gone_cap, const int *run_if, const ScanSnapshot &snap,
                                          cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    switch (c.N / (CL * CT)) {
    case 1: return launch_cluster_t<1>(c, state, base, hist, mag, n_frames, gone, gone_cap, run_if, snap, st);
    case 2: return launch_cluster_t<2>(c, state, base, hist, mag, n_frames, gone, gone_cap, run_if, snap, st);
    case 4: return launch_cluster_t<4>(c, state, base, hist, mag, n_frames, gone, gone_cap, run_if, snap, st);
    case 8: return launch_cluster_t<8>(c, state, base, hist, mag, n_frames, gone, gone_cap, run_if, snap, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_detect_scan_cluster(const DetConfig &c, DetState *state, float *base, float *hist,
                                       const float *mag, int64_t n_frames, GoneBurst *gone,
                                       uint32_t gone_cap, cudaStream_t st) {
    return launch_detect_scan_cluster_if(c, state, base, hist, mag, n_frames, gone, gone_cap, nullptr, ScanSnapshot{}, st);
}

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
