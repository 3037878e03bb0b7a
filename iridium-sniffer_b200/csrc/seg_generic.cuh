// seg_generic.cuh -- one segment of the segmented state machine (k_detect_seg.cu) walked the plain way: no
// limit of 32 bursts, no lane tricks, one thread, the reference's steps in order (burst_detect.c:426-632).
// It takes over a segment the fast walker cannot hold -- more than 32 bursts alive at once: BASELINE config 4
// (170 at once), or the ~100 spurious detections a strong burst spawns at 12 MHz / int8 / 16384-pt frames --
// so that such a stretch costs a slow segment instead of handing the whole chunk to the cluster kernel.
//
// Same inputs and outputs as the fast walker (guard-banded bitmaps, exact baseline snapshots, burst list at the
// segment's first frame in, quiet flags / list at the last frame / gone records / creation count out), and the same
// decisions: the bitmaps decide what they prove, everything else is an IEEE divide on the frame's baseline B_v.
// Written once over a "lanes" policy: on the device the 32 lanes of the segment's warp share the word loops, the
// burst loops, the ordered compaction and the arg-max of the creation step (SegLanesWarp, k_detect_seg.cu); on the
// host one lane runs the same code (SegLanesOne): tests/seg_generic_host_shim.cpp compiles it for the host and
// tests/test_seg_scan_model.py runs it in place of the numpy walker against the CPU oracle.
#pragma once
#include <stdint.h>

#include "ir_internal.h"

namespace ir {

#ifndef __CUDACC__
#define IR_HD
#else
#define IR_HD __host__ __device__
#endif

struct SegGenArgs {
    int N, half_bw, max_bursts, pre_len, post_len, max_burst_len;
    float thr;
    int seg, f0, n_frames;              // frames [f0, f0 + n_frames) of the chunk
    long long index0;                   // sample index of the segment's first frame
    int sq_start;                       // squelch counter at the segment's first frame
    const uint32_t *xu;                 // chunk bitmaps: row f = [XU words][X words], 2 * N/32 words
    const float *mag;                   // chunk magnitudes [frames][N]
    const float *snap;                  // baseline snapshots [slots][N]
    const int *fslot;                   // snapshot slot per chunk frame (-1: none)
    const uint32_t *valid;              // peak search range minus the DC notch, N/32 words
    unsigned long long *keys;           // scratch: candidate peaks of one frame (pcap of them)
    int pcap;
};

struct SegGenOut {
    int n_end, n_gone, n_create;
    uint32_t qbits[(IR_SEG_LEN + 31) / 32];
};

constexpr int SEGG_NONE = -0x40000000;
constexpr unsigned long long SEGG_CODE = 1ull << 63;

// one lane: the host build, and what the policy's members mean
struct SegLanesOne {
    static constexpr int L = 1;
    IR_HD int lane() const { return 0; }
    IR_HD void sync() const {}
    IR_HD bool any(bool p) const { return p; }
    IR_HD int sum(int v) const { return v; }
    IR_HD int max(int v) const { return v; }
    IR_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    IR_HD void and_word(uint32_t *p, uint32_t m) const { *p &= m; }
    IR_HD int excl_scan(int v, int &total) const { total = v; return 0; }     // over the lanes
    IR_HD unsigned long long max64(unsigned long long v) const { return v; }
    // the bitmap row of the frame at hand, somewhere close (the device copies it into shared memory)
    IR_HD const uint32_t *stage(const uint32_t *row, int) const { return row; }
    IR_HD void tick(int) const {}     // phase boundary (cycle counters of -DIR_SEG_TIMING builds)
};

IR_HD inline int segg_ctz(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
IR_HD inline uint32_t segg_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    __builtin_memcpy(&u, &f, 4);
    return u;
#endif
}
IR_HD inline float segg_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    __builtin_memcpy(&f, &u, 4);
    return f;
#endif
}
IR_HD inline int segg_popc(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

// work: the burst list, n_start entries on entry (dl / lah / tl in frames of the CHUNK), capacity cap.
// fv: N/32 words of scratch the lanes share.  Every lane of the policy calls it with the same arguments; the
// return value is the same on every lane: 0 or the bail reason (3 too long, 4 peak list, 5 burst table, 6 squelch,
// 9 missing snapshot, 10 gone list).
template <class LP>
IR_HD inline int seg_walk_generic_t(const LP lp, const SegGenArgs &a, SegBurst *work, int n_start, int cap, GoneBurst *gl,
                                    int gl_cap, uint32_t *fv, SegGenOut &out) {
    constexpr int L = LP::L;
    const int lane = lp.lane();
    const uint32_t lt = L == 1 ? 0u : ((1u << lane) - 1u);
    const int N = a.N, W = N >> 5;
    const int PF = (a.post_len + N - 1) / N;
    const int pf0 = (a.post_len - a.pre_len + N - 1) / N;
    const int PF0 = pf0 > 1 ? pf0 : 1;
    const int TLF = a.max_burst_len <= 0 ? 0x20000000 : (a.max_burst_len >= a.pre_len ? (a.max_burst_len - a.pre_len) / N : -1);
    int n = n_start, sq = a.sq_start, n_gone = 0, n_create = 0;            // the same on every lane
    uint32_t qb[(IR_SEG_LEN + 31) / 32];
    for (int w = 0; w < (IR_SEG_LEN + 31) / 32; w++) qb[w] = 0;
    auto mask_range = [&](int cb) {                           // (bursts of different lanes may share a word)
        int lo = cb - a.half_bw, hi = cb + a.half_bw;
        if (lo < 0) lo = 0;
        if (hi > N - 1) hi = N - 1;
        for (int w = lo >> 5; w <= (hi >> 5); w++) {
            int s = lo > (w << 5) ? lo : (w << 5), e = hi < (w << 5) + 31 ? hi : (w << 5) + 31;
            s &= 31; e &= 31;
            const uint32_t m = (e == 31 ? 0xffffffffu : ((1u << (e + 1)) - 1u)) & ~((1u << s) - 1u);
            lp.and_word(&fv[w], ~m);
        }
    };
    auto rebuild_fv = [&]() {
        for (int w = lane; w < W; w += L) fv[w] = a.valid[w];
        lp.sync();
        for (int i = lane; i < n; i += L) mask_range(work[i].cb);
        lp.sync();
    };
    rebuild_fv();
    auto bits3 = [&](const uint32_t *bm, int cb) {          // bins cb-1 .. cb+1 (inside the spectrum) -> 3-bit mask
        uint32_t r = 0;
        for (int d = -1; d <= 1; d++) {
            const int b = cb + d;
            if (b >= 0 && b < N && ((bm[b >> 5] >> (b & 31)) & 1u)) r |= 1u << (d + 1);
        }
        return r;
    };
    auto live = [&](int bn) { return ((fv[bn >> 5] >> (bn & 31)) & 1u) != 0u; };
    // (burst i is lane i % L's between two compactions / creations: its fields need no hand-over in between)
    for (int fl = 0; fl < a.n_frames; fl++) {
        const int f = a.f0 + fl;                            // frame of the chunk
        lp.tick(-1);
        const uint32_t *XU = lp.stage(a.xu + (size_t)f * (2 * W), 2 * W), *X = XU + W;
        lp.tick(0);
        uint32_t acc = 0;
        for (int w = lane; w < W; w += L) acc |= XU[w] & fv[w];
        bool ev = acc != 0u, too_long = false;
        for (int i = lane; i < n; i += L) {
            const SegBurst &b = work[i];
            const bool x3 = bits3(X, b.cb) != 0u, u3 = bits3(XU, b.cb) != 0u;
            if ((!x3 && (u3 || f >= b.dl)) || f > b.tl) ev = true;
            if (f > b.tl) too_long = true;
        }
        ev = lp.any(ev);
        too_long = lp.any(too_long);
        lp.tick(1);
        if (!ev) {
            for (int i = lane; i < n; i += L)
                if (bits3(X, work[i].cb)) { work[i].dl = f + PF; work[i].lah = f; }
            sq = sq > 0 ? sq - 1 : 0;
            if (n == 0) qb[fl >> 5] |= 1u << (fl & 31);
            lp.tick(2);
            continue;
        }
        if (too_long) return 3;
        const float *row = a.mag + (size_t)f * N;
        const int slot = a.fslot[f];
        const float *B = slot >= 0 ? a.snap + (size_t)slot * N : nullptr;
        int err = 0;
        // update_bursts (:458-469) and who is gone (:490-518)
        int n_done = 0;
        for (int i = lane; i < n; i += L) {
            SegBurst &b = work[i];
            bool hit = bits3(X, b.cb) != 0u;
            if (!hit) {
                const uint32_t u3 = bits3(XU, b.cb);
                if (u3) {
                    if (!B) { err = 9; break; }
                    for (int d = -1; d <= 1; d++)
                        if ((u3 >> (d + 1)) & 1u) {
                            const int bn = b.cb + d;
                            if (B[bn] > 0.0f && row[bn] / B[bn] > a.thr) hit = true;
                        }
                }
            }
            if (hit) { b.dl = f + PF; b.lah = f; }
            if (!hit && f >= b.dl) n_done++;
        }
        lp.tick(3);
        // peaks: exact crossings & mask of the previous frame & search range (:522-548).  First the bins the bitmaps
        // cannot rule out, packed into the list 32 words at a time; then lane l takes entries l, l + L, ... (bins of
        // one burst sit in one word: this spreads them over the lanes), divides, and keeps what crosses as a key
        // (rel's bits, then the bin counted from the top: a larger key is a stronger peak, or the same at a lower bin)
        unsigned long long *keys = a.keys;
        int np = 0;
        for (int w0 = 0; w0 < W; w0 += L) {
            const int w = w0 + lane;
            uint32_t m = w < W ? (XU[w] & fv[w]) : 0u;
            int tot;
            int pos = np + lp.excl_scan(segg_popc(m), tot);
            if (np + tot > a.pcap) { err = 4; break; }
            while (m) {
                keys[pos++] = (unsigned long long)((w << 5) + segg_ctz(m));
                m &= m - 1;
            }
            np += tot;
        }
        if (np && !B) err = 9;
        err = lp.max(err);
        n_done = lp.sum(n_done);
        if (err) return err;
        lp.sync();
        int cl = np > lane ? (np - lane + L - 1) / L : 0;   // this lane's entries: lane + L * k, k < cl
        for (int k0 = cl - 1; k0 >= 0; k0 -= 4) {
            int bnv[4];
            float bv[4], rv[4];
            for (int u = 0; u < 4; u++) bnv[u] = k0 - u >= 0 ? (int)keys[lane + L * (k0 - u)] : -1;
            for (int u = 0; u < 4; u++)
                if (bnv[u] >= 0) { bv[u] = B[bnv[u]]; rv[u] = row[bnv[u]]; }
            for (int u = 0; u < 4; u++) {
                if (bnv[u] < 0) continue;
                const int k = k0 - u;
                float rel = 0.0f;
                bool keep = false;
                if (bv[u] > 0.0f) { rel = rv[u] / bv[u]; keep = rel > a.thr; }
                if (keep) {
                    keys[lane + L * k] = ((unsigned long long)segg_f2u(rel) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)bnv[u]);
                } else {                                    // the lane's last entry (examined already) takes the slot
                    cl--;
                    if (k != cl) keys[lane + L * k] = keys[lane + L * cl];
                }
            }
        }
        lp.tick(4);
        // delete_gone_bursts: list order = creation order = ascending id; compacted L entries at a time
        if (n_done) {
            int k = 0;
            for (int i0 = 0; i0 < n; i0 += L) {
                const int i = i0 + lane;
                const bool have = i < n;
                SegBurst b;
                if (have) b = work[i];
                const bool gone = have && !(b.lah == f) && f >= b.dl;
                const bool keep = have && !gone;
                const uint32_t gm = lp.ballot(gone), km = lp.ballot(keep);
                if (n_gone + segg_popc(gm) > gl_cap) return 10;
                lp.sync();                                  // everyone holds its entry before slots are rewritten
                if (gone) {
                    GoneBurst g;
                    g.id = b.id; g.start = b.start;
                    g.stop = (unsigned long long)(a.index0 + (long long)fl * N);
                    g.last_active = b.lah == SEGG_NONE ? b.last0
                                                       : (unsigned long long)(a.index0 + (long long)(b.lah - a.f0) * N);
                    g.center_bin = b.cb; g.peak_rel = b.rel; g.base_at_create = b.base; g.pad = 0;
                    gl[n_gone + segg_popc(gm & lt)] = g;
                }
                if (keep) work[k + segg_popc(km & lt)] = b;
                n_gone += segg_popc(gm);
                k += segg_popc(km);
            }
            n = k;
            lp.sync();
            rebuild_fv();                                   // update_burst_mask (:482-486)
        }
        lp.tick(5);
        // create_new_bursts (:556-591): strongest remaining peak first, ties by bin.  Every candidate was outside
        // every burst's range when it was listed; the only ranges that can cover it since are the ones created here,
        // so each pass drops what the newest burst covers (the reference skips those when their turn comes) and
        // finds the strongest of the rest.
        int klo = 1, khi = 0;
        for (;;) {
            unsigned long long best = 0ull;
            for (int k = cl - 1; k >= 0; k--) {
                const unsigned long long key = keys[lane + L * k];
                const int bn = (int)(0xffffffffu - (uint32_t)key);
                if (bn >= klo && bn <= khi) {
                    cl--;
                    if (k != cl) keys[lane + L * k] = keys[lane + L * cl];
                } else if (key > best) {
                    best = key;
                }
            }
            best = lp.max64(best);
            if (best == 0ull) break;
            const int bb = (int)(0xffffffffu - (uint32_t)best);
            const float br = segg_u2f((uint32_t)(best >> 32));
            if (n >= cap) return 5;
            if (lane == 0) {
                SegBurst nb;
                nb.id = SEGG_CODE | ((unsigned long long)a.seg << 32) | (unsigned long long)n_create;
                nb.start = (unsigned long long)(a.index0 + (long long)fl * N - (long long)a.pre_len);
                nb.last0 = nb.start;
                nb.cb = bb; nb.rel = br; nb.base = B[bb];
                nb.dl = f + PF0; nb.lah = SEGG_NONE; nb.tl = f + TLF;
                work[n] = nb;
                mask_range(bb);
            }
            n++;
            n_create++;
            klo = bb - a.half_bw; khi = bb + a.half_bw;
        }
        lp.sync();
        lp.tick(6);
        if (a.max_bursts > 0 && n > a.max_bursts) return 6;  // squelch (:593-631)
        if (sq > 0) sq--;
        if (n == 0) qb[fl >> 5] |= 1u << (fl & 31);
    }
    lp.sync();
    if (lane == 0) {
        out.n_end = n; out.n_gone = n_gone; out.n_create = n_create;
        for (int w = 0; w < (IR_SEG_LEN + 31) / 32; w++) out.qbits[w] = qb[w];
    }
    lp.sync();
    return 0;
}

// one lane (the host's entry point)
IR_HD inline int seg_walk_generic(const SegGenArgs &a, SegBurst *work, int n_start, int cap, GoneBurst *gl, int gl_cap,
                                  uint32_t *fv /* N/32 words of scratch */, SegGenOut &out) {
    return seg_walk_generic_t(SegLanesOne{}, a, work, n_start, cap, gl, gl_cap, fv, out);
}

}  // namespace ir
