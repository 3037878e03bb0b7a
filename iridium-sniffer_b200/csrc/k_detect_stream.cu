// k_detect_stream.cu -- the burst state machine (burst_detect.c:426-632) as two cooperating
// kernels that take the dense magnitude data off the serial path.
//
// The frame-to-frame dependency of the reference is carried by two things only: the list of
// active bursts (a handful of integers) and the noise baseline, which changes on "quiet" frames
// (no burst active, burst_detect.c:438-454) and is frozen otherwise.  Everything else is a
// per-(frame, bin) threshold test.  So:
//
//  k_detect_classify (all SMs): tests every magnitude of a run of frames against a REFERENCE
//      baseline `ref` (the baseline as it stood when the kernel ran) with a guard band:
//         X  bit: mag > thr * HI*ref            -> certainly above for any baseline in [LO, HI]*ref
//         XU bit: not (mag < thr * LO*ref)      -> above or "uncertain" (needs the exact test)
//      (LO, HI = IR_GUARD_LO, IR_GUARD_HI) and writes two bitmaps per frame (N/32 words each) --
//      1/16 of the magnitude bytes.
//
//  k_detect_scan_stream (one 8-CTA cluster, 1 + 64 warps):
//      * leader = one warp.  Walks the frames in order reading only the bitmaps (streamed into a
//        shared-memory ring by TMA bulk copies, eight rows per mbarrier, ~7 blocks ahead).  Up to
//        32 active bursts live in lane registers.  A frame on which no unmasked bit is set, no
//        burst ends and every hysteresis test is decided by an X bit costs ~170 cycles (two such
//        frames per trip).  Otherwise ("event") the leader runs the reference's steps for that
//        frame exactly: IEEE divide against the true baseline for the few bins concerned, peaks
//        strongest first, deletions in list (= id) order, mask, creations.
//      * workers = 8 x 256 threads, two to eight bins each, baseline in registers.  The leader
//        hands them ranges of quiet frames through a command ring in global memory; they apply
//        the reference's update (two roundings per bin, in frame order), keep the 512-row
//        history and an undo log of what they overwrite in it, publish the baseline, and check
//        that it stays inside the guard band the bitmaps were made for.
//      The leader waits for the workers only when it needs baseline values: at the first event
//      after a quiet run.
//
//  Anything unusual -- guard band violated, squelch, a burst that may exceed max_burst_len (forced
//  baseline update), the history priming in mid-launch, a 33rd concurrent burst -- makes the
//  leader bail: the state block is left untouched, the fallback launch (the cluster kernel of
//  k_detect_cluster.cu, which handles every case) replays the undo log and redoes the frames.
//  Both produce the reference's result bit for bit; tests compare all variants with the CPU oracle.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace ir {

namespace {

constexpr int SCL = IR_STREAM_CL;      // CTAs per cluster
constexpr int SWT = 256;               // worker threads per CTA (few threads = many registers for the leader)
constexpr int SNT = SWT + 32;          // + one more warp (the leader, in CTA 0)
constexpr int SGF = 8;                 // rows per ring block
constexpr int SPF = 4;                 // candidate words whose loads are in flight together
constexpr int SMAXW = 512;             // bitmap words per frame (N <= 16384)
constexpr int SMAXC = 1024;            // candidate peaks of one frame
constexpr int SQCH = 128;              // quiet frames per worker command
constexpr int SGMAX = 8;               // frames per worker register chunk (4 when a thread owns 4 bins)
constexpr uint32_t FULL = 0xffffffffu;

struct StShared {
    unsigned long long bar[8];         // one per ring block
    uint32_t valid[SMAXW];             // peak search range minus the DC notch
    uint32_t fvs[SMAXW];               // valid & not covered by an active burst (the lanes cache their words)
    int cw[SMAXW];                     // words of the frame with a possible unmasked crossing (beyond SPF)
    int cbin[SMAXC];                   // candidate peaks of one frame when they do not fit the registers
    float crel[SMAXC];
    float cbase[SMAXC];
    unsigned long long wcmd;           // workers: the command being executed
};

__device__ __forceinline__ unsigned long long ld_vol64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ unsigned ld_vol32(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }
__device__ __forceinline__ unsigned ld_acq32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// shared-memory loads by 32-bit address (the leader's hot loop keeps its row addresses in registers)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// mbarrier / bulk-copy helpers by 32-bit shared address (kept in registers: the generic-to-shared
// conversion would otherwise be rematerialised, S2R and all, at every use)
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d_a(uint32_t dst, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SWT) : "memory"); }

__device__ __forceinline__ uint32_t st_range_bits(int w, int lo, int hi) {
    int a = max(lo, w << 5), b = min(hi, (w << 5) + 31);
    if (a > b) return 0u;
    a &= 31; b &= 31;
    return (b == 31 ? FULL : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
}

// command word: bit 0 exit, bits 1..20 first frame, 21..41 end frame, 42..63 launch epoch
__device__ __forceinline__ unsigned long long pack_cmd(unsigned epoch, int s, int e, int ex) {
    return ((unsigned long long)(epoch & 0x3fffffu) << 42) | ((unsigned long long)e << 21) |
           ((unsigned long long)s << 1) | (unsigned long long)ex;
}

}  // namespace

// =========================================================================== classification
// One warp = 1024 consecutive bins (32 bitmap words) of a run of frames.
__global__ void __launch_bounds__(128)
k_detect_classify(const float *__restrict__ mag, const float *base_g, float thr, int N, int n_frames,
                  int frames_per_warp, uint32_t *__restrict__ xu, float *__restrict__ ref_out,
                  unsigned char *__restrict__ rowany) {
    const int lane = threadIdx.x & 31;
    const int gw = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int ncol = N >> 10;
    const int col = gw % ncol, part = gw / ncol;
    const int f0 = part * frames_per_warp;
    if (f0 >= n_frames) return;
    const int f1 = min(f0 + frames_per_warp, n_frames);
    const int W = N >> 5;
    const float INF = __int_as_float(0x7f800000);
    float thi[32], tlo[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int bin = (col << 10) + (j << 5) + lane;
        const float r = *reinterpret_cast<const volatile float *>(base_g + bin);
        if (part == 0 && ref_out != nullptr) ref_out[bin] = r;
        if (r > 0.0f) {
            thi[j] = thr * (r * IR_GUARD_HI) * 1.0001f;
            tlo[j] = thr * (r * IR_GUARD_LO) * 0.9999f;
        } else {
            thi[j] = INF; tlo[j] = INF;        // rel is 0 while the baseline is not positive
        }
    }
    for (int f = f0; f < f1; f++) {
        const float *row = mag + (size_t)f * N + (col << 10) + lane;
        float m[32];
#pragma unroll
        for (int j = 0; j < 32; j++) m[j] = __ldg(row + (j << 5));
        uint32_t xw = 0, uw = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const bool hi = m[j] > thi[j];
            const bool un = !(m[j] < tlo[j]);                  // certain or uncertain (NaN: uncertain)
            const uint32_t bx = __ballot_sync(FULL, hi), bu = __ballot_sync(FULL, un);
            if (lane == j) { xw = bx; uw = bu; }
        }
        uint32_t *o = xu + (size_t)f * (2 * W) + (col << 5) + lane;
        o[0] = uw;                                             // row = [XU][X]
        o[W] = xw;
        // (segmented scan) does the row have any bit at all?  cleared by the caller; everybody stores 1
        if (rowany != nullptr && __any_sync(FULL, uw != 0u) && lane == 0) rowany[f] = 1;
    }
}

// =========================================================================== state machine
// BPT = bins per worker thread = N / (SCL * SWT)
template <int BPT>
__global__ void __cluster_dims__(SCL, 1, 1) __launch_bounds__(SNT, 1)
k_detect_scan_stream(DetConfig c, DetState *__restrict__ gs, float *base_g, float *hist,
                     const float *__restrict__ mag, const uint32_t *__restrict__ xu,
                     const float *__restrict__ ref, int n_frames, GoneBurst *__restrict__ gone,
                     uint32_t gone_cap, StreamCtl *ctl, unsigned epoch, float *__restrict__ undo,
                     float *__restrict__ base_snap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StShared &S = *reinterpret_cast<StShared *>(smem_raw);
    const int rank = (int)blockIdx.x;
    const int N = c.N, W = N >> 5, H = c.hist_size;
    const bool classified = xu != nullptr;

    if (threadIdx.x < SWT) {
        // ------------------------------------------------------------------ workers
        const int wt = threadIdx.x;
        constexpr int NB = BPT * SWT;
        constexpr int SG = BPT >= 8 ? 4 : SGMAX;
        const int bin0 = rank * NB;
        float base[BPT], glo[BPT], ghi[BPT];
        int bad = 0;
#pragma unroll
        for (int u = 0; u < BPT; u++) {
            const int bin = bin0 + u * SWT + wt;
            base[u] = base_g[bin];
            base_snap[bin] = base[u];                         // what a bailed launch is undone to
            glo[u] = 0.0f; ghi[u] = 0.0f;
            if (classified) {
                const float r = ref[bin];
                const float a = r * IR_GUARD_LO, b = r * IR_GUARD_HI;
                glo[u] = fminf(a, b); ghi[u] = fmaxf(a, b);
                bad |= !(base[u] >= glo[u] && base[u] <= ghi[u]);
            }
        }
        int hist_idx = gs->hist_idx, primed = gs->primed;
        unsigned my = 0;
        float *ulog = undo + bin0 + wt;                       // undo log: the history value every update overwrites
        for (;;) {
            if (wt == 0) {
                unsigned long long v;
                for (;;) {
                    v = ld_vol64(&ctl->cmd[my]);
                    if ((unsigned)(v >> 42) == (epoch & 0x3fffffu) && v != 0ull) break;
                    __nanosleep(20);
                }
                S.wcmd = v;
            }
            worker_bar();
            const unsigned long long v = S.wcmd;
            if (v & 1ull) break;
            const int s = (int)((v >> 1) & 0xfffffu), e = (int)((v >> 21) & 0x1fffffu);
            for (int f = s; f < e; f += SG) {
                float mv[SG][BPT], ov[SG][BPT];
#pragma unroll
                for (int g = 0; g < SG; g++) {
                    const bool in = f + g < e;
                    const float *row = mag + (size_t)(f + g) * N + bin0 + wt;
                    int hrow = hist_idx + g;
                    if (hrow >= H) hrow -= H;
                    const float *h = hist + (size_t)hrow * N + bin0 + wt;
                    // a history row is live if the detector is primed or wraps before reaching it
                    const bool live = primed || (hist_idx + g >= H);
#pragma unroll
                    for (int u = 0; u < BPT; u++) {
                        mv[g][u] = in ? __ldg(row + u * SWT) : 0.0f;
                        ov[g][u] = (in && live) ? h[u * SWT] : 0.0f;
                    }
                }
#pragma unroll
                for (int g = 0; g < SG; g++) {
                    if (f + g < e) {
                        float *h = hist + (size_t)hist_idx * N + bin0 + wt;
#pragma unroll
                        for (int u = 0; u < BPT; u++) {
                            const float t = base[u] - ov[g][u];            // simd_avx2.c:221-236: two roundings
                            base[u] = t + mv[g][u];
                            h[u * SWT] = mv[g][u];
                            ulog[u * SWT] = ov[g][u];
                            if (classified) bad |= !(base[u] >= glo[u] && base[u] <= ghi[u]);
                        }
                        if (++hist_idx == H) { primed = 1; hist_idx = 0; }
                        ulog += N;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < BPT; u++) __stcg(base_g + bin0 + u * SWT + wt, base[u]);
            if (bad) { atomicOr(&ctl->guard_bad, 1u); bad = 0; }
            worker_bar();                                     // the CTA's stores are ordered before thread 0's release
            my++;
            if (wt == 0) st_rel32(&ctl->done[rank], my);
        }
        return;
    }
    if (rank != 0) return;

    // ---------------------------------------------------------------------- leader (one warp)
    // One frame per trip through a loop small enough to stay in the instruction cache.  A lone
    // warp pays ~6 cycles per instruction, so everything is organised to execute few of them: up
    // to 32 active bursts live entirely in registers, one per lane (the reference's list order is
    // creation order, i.e. ascending id, so no list needs to be kept: deletion frees a lane,
    // creation takes a free one); one vote per frame decides whether anything happens; only then
    // ("event") does the warp run the reference's steps for that frame.  More than 32 concurrent
    // bursts is the cluster kernel's business (bail).
    constexpr int WPL = 2 * BPT;                              // bitmap words per lane = W / 32
    constexpr int RB = BPT >= 8 ? 4 : 8;                      // ring blocks of SGF rows
    constexpr int RROWS = RB * SGF;
    const int lane = threadIdx.x & 31;
    const float thr = c.thr;
    uint64_t *bars = reinterpret_cast<uint64_t *>(S.bar);
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(StShared) + 127) / 128) * 128);
    const int RW = 2 * W;                                     // words per row: [XU][X]
    const uint32_t row_bytes = (uint32_t)RW * sizeof(uint32_t);
    int n_act = gs->n_act, sq = gs->squelch_count, hist_idx = gs->hist_idx, primed = gs->primed;
    unsigned long long next_id = gs->next_id;
    const uint64_t index0 = gs->index;
    uint32_t n_gone = gs->n_gone;
    uint32_t overflow = gs->overflow;
    int bail = 0;
    unsigned n_cmd = 0, n_waited = 0;
    unsigned long long st_events = 0, st_exact = 0, st_waits = 0, st_mini = 0;
    unsigned long long t_glob0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_glob0));
    // per-phase cycle counters of the leader, compiled in only with -DIR_SCAN_TIMING
    unsigned long long cy_scan = 0, cy_wait = 0, cy_e1 = 0, cy_e2 = 0, cy_e3 = 0, cy_e4 = 0, cy_ring = 0, n_trips = 0, cy_p1 = 0, cy_p2 = 0, cy_p3 = 0;
    long long tk = clock64();
#ifdef IR_SCAN_TIMING
#define ST_TICK(acc) do { const long long _t = clock64(); acc += (unsigned long long)(_t - tk); tk = _t; } while (0)
#else
#define ST_TICK(acc) do { (void)tk; } while (0)
#endif
    // frames a burst survives without a hit: idx - la >= post_len  <=>  frames >= PF (la = a frame's
    // index) resp. PF0 (la = start = creation frame's index - pre_len)
    const int PF = (c.post_len + N - 1) / N;
    const int PF0 = max(1, (c.post_len - c.pre_len + N - 1) / N);
    constexpr int NONE = -0x40000000;
    // la <= index of frame f', so la - start <= (f' - f)*N + pre_len <= max_burst_len while f' - f <= TLF
    const int TLF = c.max_burst_len <= 0 ? 0x20000000 : (c.max_burst_len >= c.pre_len ? (c.max_burst_len - c.pre_len) / N : -1);

    // ---- the burst of this lane (burst_detect.c:39-48) and its times in frames of this launch
    bool r_have = false;
    unsigned long long r_id = 0, r_start = 0, r_last0 = 0;    // r_last0: last_active unless a hit came in this launch
    int r_cb = 0;
    float r_rel = 0.0f, r_base = 0.0f;
    int b_dl = 0x3fffffff;        // first frame on which the burst is deleted unless a hit comes
    int b_lah = NONE;             // frame of the latest hit in this launch
    int b_tl = 0x3fffffff;        // last frame on which it cannot be "too long" yet
    uint32_t b_o0 = 0, b_o1 = 0, b_msk = 0;                   // hysteresis window: byte offsets of its two words, bit mask
    int b_sh = 0;
    auto set_window = [&]() {
        const int w0 = (r_cb - 1) >> 5;
        b_o0 = (uint32_t)w0 * 4u; b_o1 = (uint32_t)min(w0 + 1, W - 1) * 4u; b_sh = (r_cb - 1) & 31; b_msk = 7u;
    };
    if (n_act > 32) { bail = 7; n_act = 0; }
    if (lane < n_act) {
        const ActBurst b = gs->act[lane];
        r_have = true;
        r_id = b.id; r_start = b.start; r_last0 = b.last_active; r_cb = b.center_bin; r_rel = b.peak_rel; r_base = b.base_at_create;
        const long long d = (long long)(b.last_active + (unsigned long long)c.post_len) - (long long)index0;
        b_dl = d <= 0 ? 0 : (int)min((long long)0x3fffffff, (d + N - 1) / N);
        const long long t = (long long)(b.start + (unsigned long long)c.max_burst_len) - (long long)index0;
        b_tl = c.max_burst_len <= 0 ? 0x3fffffff : (t < 0 ? -1 : (int)min((long long)0x3fffffff, t / N));
        set_window();
    }
    uint32_t have_mask = __ballot_sync(FULL, r_have);
#pragma unroll 1
    for (int w = lane; w < W; w += 32) {
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) {
            const int bin = (w << 5) + b;
            const bool ok = bin >= c.half_bw && bin < N - c.half_bw && !(bin >= N / 2 - 3 && bin <= N / 2 + 3);
            v |= ok ? (1u << b) : 0u;
        }
        S.valid[w] = v;
        S.fvs[w] = v;
    }
    if (lane < SCL) *reinterpret_cast<volatile unsigned *>(&ctl->done[lane]) = 0u;
    if (lane == 0) {
        *reinterpret_cast<volatile unsigned *>(&ctl->guard_bad) = 0u;
        for (int d = 0; d < RB; d++) mbar_init(&bars[d], 1);
        fence_mbar_init();
    }
    __threadfence();
    __syncwarp();
    // free & valid mask (S.fvs): clear / set the bins lo..hi (at most three words), one lane per word
    auto mask_bins = [&](int lo, int hi, bool set) {
        const int w = (lo >> 5) + lane;
        if (lane < 3 && w <= (hi >> 5)) {
            const uint32_t m = st_range_bits(w, lo, hi);
            S.fvs[w] = set ? (S.fvs[w] | (m & S.valid[w])) : (S.fvs[w] & ~m);
        }
        __syncwarp();
    };
    auto mask_burst = [&](int cb, bool set) { mask_bins(max(cb - c.half_bw, 0), min(cb + c.half_bw, N - 1), set); };
    for (uint32_t hm = have_mask; hm; hm &= hm - 1) mask_burst(__shfl_sync(FULL, r_cb, __ffs(hm) - 1), false);
    uint32_t fv[WPL];                                         // this lane's words of S.fvs
    auto reload_fv = [&]() {
        __syncwarp();
#pragma unroll
        for (int k4 = 0; k4 < WPL / 4; k4++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(&S.fvs[lane * WPL + 4 * k4]);
            fv[4 * k4] = v.x; fv[4 * k4 + 1] = v.y; fv[4 * k4 + 2] = v.z; fv[4 * k4 + 3] = v.w;
        }
    };
    reload_fv();

    // ring of bitmap rows, filled a block (SGF rows, one bulk copy, one mbarrier) at a time
    const int n_blocks = classified ? (n_frames + SGF - 1) / SGF : 0;
    int blk_issued = 0, blk_landed = 0;
    // rows up to frame `upto` are in shared memory; blocks wholly before frame `fcur` may be refilled
    const uint32_t ring_a = opaque_u32(smem_u32(ring)), bars_a = opaque_u32(smem_u32(bars));
    const uint32_t ring_end_a = ring_a + (uint32_t)RROWS * row_bytes;
    auto ring_advance = [&](int fcur, int upto) {
        __syncwarp();
        while (blk_issued < n_blocks && (blk_issued < RB || (blk_issued - RB + 1) * SGF <= fcur)) {
            if (lane == 0) {
                const int r0 = blk_issued * SGF, nr = min(SGF, n_frames - r0);
                const uint32_t bar = bars_a + 8u * (uint32_t)(blk_issued % RB);
                mbar_expect_tx_a(bar, row_bytes * (uint32_t)nr);
                tma_load_1d_a(ring_a + (uint32_t)(r0 % RROWS) * row_bytes, xu + (size_t)r0 * RW, row_bytes * (uint32_t)nr, bar);
            }
            blk_issued++;
        }
        while (blk_landed <= upto / SGF) {
            mbar_wait_a(bars_a + 8u * (uint32_t)(blk_landed % RB), (uint32_t)((blk_landed / RB) & 1));
            blk_landed++;
        }
    };
    if (!classified && primed) bail = 1;                      // a priming launch on a primed detector

    int undo_frames = 0;                                      // quiet frames handed to the workers so far
    auto issue = [&](int s, int e, int ex) {
        if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(&ctl->cmd[n_cmd]) = pack_cmd(epoch, s, e, ex);
        n_cmd++;
        undo_frames += e - s;
    };
    // the baseline the workers hold is published and inside the guard band
    auto wait_all = [&]() -> bool {
        if (n_waited == n_cmd) return true;
        if (lane < SCL)
            while (ld_acq32(&ctl->done[lane]) < n_cmd) __nanosleep(20);
        __syncwarp();
        n_waited = n_cmd;
        st_waits++;
        return ld_vol32(&ctl->guard_bad) == 0u;
    };
    int qs = -1;                                              // first frame of the quiet range not yet handed out
    // frames [a, b) end with no burst active: update_filters_post (:438-454), by the workers
    auto quiet_frames = [&](int a, int b) {
        if (qs < 0) qs = a;
        if (b - qs >= SQCH) { issue(qs, b, 0); qs = -1; }
        hist_idx += b - a;
        if (hist_idx >= H) {                                  // (a run of unprimed frames never crosses the priming point)
            hist_idx -= H;
            if (!primed) {
                primed = 1;
                if (b < n_frames) bail = 8;                   // the bitmaps were not made for this baseline
            }
        }
    };

    // delete_gone_bursts (:490-518) for the lanes in dmask on frame fr: gone records in list order
    // (= ascending id), lanes freed, update_burst_mask (:482-486) by freeing the deleted ranges and
    // re-covering what the survivors next to them still mask.  (The caller reloads fv.)
    auto delete_done = [&](uint32_t dmask, bool done, int fr) {
        int rank_d = 0;
        if (dmask & (dmask - 1)) {                            // several at once: rank by id
            for (uint32_t m = dmask; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                const unsigned long long oid = ((unsigned long long)__shfl_sync(FULL, (unsigned)(r_id >> 32), src) << 32) |
                                               (unsigned long long)__shfl_sync(FULL, (unsigned)r_id, src);
                rank_d += oid < r_id ? 1 : 0;
            }
        }
        if (done) {
            const uint32_t slot_g = n_gone + (uint32_t)rank_d;
            if (slot_g < gone_cap) {
                GoneBurst g;
                g.id = r_id; g.start = r_start; g.stop = index0 + (uint64_t)fr * (uint64_t)N;
                g.last_active = b_lah == NONE ? r_last0 : index0 + (uint64_t)b_lah * (uint64_t)N;
                g.center_bin = r_cb; g.peak_rel = r_rel; g.base_at_create = r_base; g.pad = 0;
                gone[slot_g] = g;
            } else {
                overflow = 1;
            }
            r_have = false;
            b_dl = 0x3fffffff; b_lah = NONE; b_tl = 0x3fffffff; b_msk = 0; b_o0 = 0; b_o1 = 0; b_sh = 0;
        }
        overflow = __any_sync(FULL, overflow) ? 1u : 0u;
        n_gone += (uint32_t)__popc(dmask);
        n_act -= __popc(dmask);
        have_mask &= ~dmask;
        for (uint32_t m = dmask; m; m &= m - 1) mask_burst(__shfl_sync(FULL, r_cb, __ffs(m) - 1), true);
        for (uint32_t m = dmask; m; m &= m - 1) {
            const int cbd = __shfl_sync(FULL, r_cb, __ffs(m) - 1);
            for (uint32_t nm = __ballot_sync(FULL, r_have && abs(r_cb - cbd) <= 2 * c.half_bw); nm; nm &= nm - 1)
                mask_burst(__shfl_sync(FULL, r_cb, __ffs(nm) - 1), false);
        }
    };

    // bitmap words of one frame: this lane's XU words, and the 3-bin hysteresis window of its burst
    struct FrameRegs { uint32_t xu[WPL]; uint32_t bx0, bx1, bu0, bu1; };
    const uint32_t my_off = (uint32_t)(lane * WPL) * 4u, x_off = (uint32_t)W * 4u;
    auto load_frame = [&](FrameRegs &r, uint32_t ra) {        // ra = shared address of the frame's row
#pragma unroll
        for (int k4 = 0; k4 < WPL / 4; k4++) {
            const uint4 v = lds128(ra + my_off + 16u * k4);
            r.xu[4 * k4] = v.x; r.xu[4 * k4 + 1] = v.y; r.xu[4 * k4 + 2] = v.z; r.xu[4 * k4 + 3] = v.w;
        }
        r.bu0 = lds32(ra + b_o0); r.bu1 = lds32(ra + b_o1);
        r.bx0 = lds32(ra + x_off + b_o0); r.bx1 = lds32(ra + x_off + b_o1);
    };

    int f = 0;
    // ---- frames before the detector is primed: nothing is detected (:426-428), every frame is quiet
    while (f < n_frames && !bail && !primed) {
        const int G = min(min(SQCH, n_frames - f), H - hist_idx);
        quiet_frames(f, f + G);
        f += G;
    }
    FrameRegs cur;
    uint32_t ra = ring_a + (uint32_t)(f % RROWS) * row_bytes;   // shared address of frame f's row
    if (f < n_frames && !bail && (f & (SGF - 1)) != 0) ring_advance(f, f);
    // (no software pipelining of the loads: a single warp pays more for the register moves of a
    // double buffer than for one shared-memory latency per frame)
#pragma unroll 1
    while (f < n_frames && !bail) {
        n_trips++;
        // ---- (A) uneventful frames, two at a time, in a small loop of their own: their instruction
        // streams are independent (but for the deadline of frame f+1, which a hit on frame f
        // moves), so the warp overlaps them.  An event in the pair: commit frame f if it is
        // uneventful, leave the loop and take the one-frame path (B) at the event frame.
        ST_TICK(cy_scan);
#pragma unroll 1
        for (;;) {
            if ((f & (SGF - 1)) == 0) ring_advance(f, f);     // frame f's block has landed; older blocks are refilled
            if ((f & (SGF - 1)) == SGF - 1 || f + 1 >= n_frames) break;
            FrameRegs c1;
            load_frame(cur, ra);
            load_frame(c1, ra + row_bytes);                   // same block: no wrap, landed
            uint32_t a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < WPL; k++) { a0 |= cur.xu[k] & fv[k]; a1 |= c1.xu[k] & fv[k]; }
            const bool h0 = (__funnelshift_r(cur.bx0, cur.bx1, b_sh) & b_msk) != 0u;
            const bool h1 = (__funnelshift_r(c1.bx0, c1.bx1, b_sh) & b_msk) != 0u;
            const bool q0 = (__funnelshift_r(cur.bu0, cur.bu1, b_sh) & b_msk) != 0u;
            const bool q1 = (__funnelshift_r(c1.bu0, c1.bu1, b_sh) & b_msk) != 0u;
            const int dl1 = h0 ? f + PF : b_dl;
            const bool e0 = a0 != 0u || (!h0 && (q0 || f >= b_dl)) || f > b_tl;
            const bool e1 = a1 != 0u || (!h1 && (q1 || f + 1 >= dl1)) || f + 1 > b_tl;
            const uint32_t em = __ballot_sync(FULL, e0) ? 1u : (__ballot_sync(FULL, e1) ? 2u : 0u);
            if (em == 1u) {                                   // frame f is an event
                // the commonest one -- bursts reaching their deadline, nothing else -- is handled right
                // here: no baseline value is needed and no peak can appear
                const bool d0 = !h0 && f >= b_dl;
                if (__any_sync(FULL, a0 != 0u || (!h0 && q0) || f > b_tl)) break;
                st_events++;
                st_mini++;
                if (h0) { b_dl = f + PF; b_lah = f; }
                delete_done(__ballot_sync(FULL, d0), d0, f);
                reload_fv();
                if (sq > 0) sq--;                             // :628-631
                if (n_act == 0) quiet_frames(f, f + 1);
                f += 1;
                ra += row_bytes;
                if (ra == ring_end_a) ra = ring_a;
                if (f >= n_frames) break;
                continue;
            }
            if (h0) { b_dl = f + PF; b_lah = f; }
            const int adv = em == 0u ? 2 : 1;
            if (em == 0u && h1) { b_dl = f + 1 + PF; b_lah = f + 1; }
            sq = max(sq - adv, 0);                            // :628-631
            if (n_act == 0) quiet_frames(f, f + adv);
            f += adv;
            ra += row_bytes * (uint32_t)adv;
            if (ra == ring_end_a) ra = ring_a;
            if (f >= n_frames) break;                         // (an event on frame f+1 is met again as frame f of the next trip)
        }
        ST_TICK(cy_p1);
        if (f >= n_frames || bail) break;
        // ---- (B) one frame, any case
        load_frame(cur, ra);
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < WPL; k++) acc |= cur.xu[k] & fv[k];
        const uint32_t x3 = __funnelshift_r(cur.bx0, cur.bx1, b_sh) & b_msk;
        const uint32_t u3 = __funnelshift_r(cur.bu0, cur.bu1, b_sh) & b_msk;
        bool hit = x3 != 0u;                                  // update_bursts (:458-469), decided by an X bit
        const bool ev = acc != 0u || (!hit && (u3 != 0u || f >= b_dl)) || f > b_tl;
        if (!__any_sync(FULL, ev)) {
            // ---- nothing happens on this frame
            if (hit) { b_dl = f + PF; b_lah = f; }
            sq = max(sq - 1, 0);                              // :628-631
            if (n_act == 0) quiet_frames(f, f + 1);
        } else {
            // ================= event frame f: the reference's steps, exactly
            st_events++;
            ST_TICK(cy_scan);
            const bool cand_any = __any_sync(FULL, acc != 0u);
            const bool unc_any = __any_sync(FULL, !hit && u3 != 0u);
            if (__any_sync(FULL, f > b_tl)) { bail = 3; break; }   // a burst may exceed max_burst_len (:498-517)
            // words with a possible unmasked crossing; get their magnitudes (and, while no baseline
            // update is pending, baselines) moving before anything else
            const float *row = mag + (size_t)f * N;
            int n_cw = 0;
            float mvp[SPF], bsp[SPF];
            int wj[SPF];
#pragma unroll
            for (int j = 0; j < SPF; j++) { mvp[j] = 0.0f; bsp[j] = 0.0f; wj[j] = -1; }
            const bool base_ok = qs < 0 && n_waited == n_cmd;
            if (cand_any) {
                uint32_t wmask = 0;
#pragma unroll
                for (int k = 0; k < WPL; k++) wmask |= (cur.xu[k] & fv[k]) ? (1u << k) : 0u;
                // the (few) lanes with such words take turns; every lane learns the whole list
                uint32_t lm = __ballot_sync(FULL, wmask != 0u);
                while (lm) {
                    const int src = __ffs(lm) - 1;
                    lm &= lm - 1;
                    uint32_t wm = __shfl_sync(FULL, wmask, src);
                    while (wm) {
                        const int w = src * WPL + __ffs(wm) - 1;
                        wm &= wm - 1;
#pragma unroll
                        for (int j = 0; j < SPF; j++) if (j == n_cw) wj[j] = w;
                        if (n_cw >= SPF && lane == 0) S.cw[n_cw] = w;
                        n_cw++;
                    }
                }
                if (n_cw > SPF && lane == 0) {
#pragma unroll
                    for (int j = 0; j < SPF; j++) S.cw[j] = wj[j];
                }
#pragma unroll
                for (int j = 0; j < SPF; j++)
                    if (wj[j] >= 0) {
                        mvp[j] = row[(wj[j] << 5) + lane];
                        if (base_ok) bsp[j] = __ldcg(base_g + (wj[j] << 5) + lane);
                    }
            }
            ST_TICK(cy_e1);
            if (cand_any || unc_any) {                        // baseline values are about to be used
                if (qs >= 0) { issue(qs, f, 0); qs = -1; }
                if (!wait_all()) { bail = 2; break; }
            }
            ST_TICK(cy_wait);
            const uint64_t idx = index0 + (uint64_t)f * (uint64_t)N;
            __syncwarp();
            // update_bursts: the tests an X bit did not decide
            if (unc_any && !hit) {
                uint32_t q = u3;
                while (q) {
                    const int b = r_cb - 1 + __ffs(q) - 1;
                    q &= q - 1;
                    const float bs = __ldcg(base_g + b);
                    if (bs > 0.0f && row[b] / bs > thr) hit = true;
                }
            }
            if (hit) { b_dl = f + PF; b_lah = f; }
            const bool done = !hit && f >= b_dl;              // (a lane without a burst has no deadline)
            const uint32_t dmask = __ballot_sync(FULL, done);
            // peaks: exact crossings & mask of the previous frame & search range (:522-548).  Up to SPF
            // words stay in registers (lane = bin inside the word); more go through shared memory.
            const bool fast = n_cw <= SPF;
            float relj[SPF], bsj[SPF];
            bool exj[SPF];
            int n_cand = 0;
#pragma unroll 1
            for (int j0 = 0; j0 < n_cw; j0 += SPF) {
                float mv[SPF];
#pragma unroll
                for (int j = 0; j < SPF; j++) {
                    relj[j] = 0.0f; bsj[j] = 0.0f; exj[j] = false; mv[j] = 0.0f;
                    if (j0 > 0) wj[j] = j0 + j < n_cw ? S.cw[j0 + j] : -1;
                    if (wj[j] >= 0) {
                        const int bin = (wj[j] << 5) + lane;
                        mv[j] = j0 == 0 ? mvp[j] : row[bin];
                        bsj[j] = (j0 == 0 && base_ok) ? bsp[j] : __ldcg(base_g + bin);
                    }
                }
#pragma unroll
                for (int j = 0; j < SPF; j++) {
                    if (wj[j] >= 0) {
                        if (bsj[j] > 0.0f) { relj[j] = mv[j] / bsj[j]; exj[j] = relj[j] > thr; }   // simd_avx2.c:239-257
                        exj[j] = exj[j] && ((S.fvs[wj[j]] >> lane) & 1u);
                        const uint32_t bal = __ballot_sync(FULL, exj[j]);
                        if (!fast && exj[j]) {
                            const int pos = n_cand + __popc(bal & ((1u << lane) - 1u));
                            if (pos < SMAXC) { S.cbin[pos] = (wj[j] << 5) + lane; S.crel[pos] = relj[j]; S.cbase[pos] = bsj[j]; }
                        }
                        n_cand += __popc(bal);
                    }
                }
            }
            st_exact += (unsigned long long)n_cw;
            if (n_cand > SMAXC) { bail = 4; break; }
            __syncwarp();
            ST_TICK(cy_e2);
            bool mask_changed = false;
            if (dmask) {
                delete_done(dmask, done, f);
                mask_changed = true;
            }
            ST_TICK(cy_e3);
            if (n_cand > 0) {
                // create_new_bursts (:556-591): strongest remaining peak first, ties by bin.  Up to SPF
                // words: the candidates never leave the registers; else a list in shared memory.
                const int nc = n_cand;
#pragma unroll 1
                for (;;) {
                    int bin;
                    float rel_w, bc;
                    if (fast) {
                        uint32_t key = 0;
                        int kb = 0x7fffffff;
#pragma unroll
                        for (int j = 0; j < SPF; j++) {
                            const uint32_t kj = exj[j] ? __float_as_uint(relj[j]) : 0u;   // rel > thr > 0: bit order = value order
                            const int bj = (wj[j] << 5) + lane;
                            if (kj > key || (kj != 0u && kj == key && bj < kb)) { key = kj; kb = bj; }
                        }
                        const uint32_t m = __reduce_max_sync(FULL, key);
                        if (m == 0u) break;
                        bin = __reduce_min_sync(FULL, key == m ? kb : 0x7fffffff);
                        float bcl = 0.0f;
#pragma unroll
                        for (int j = 0; j < SPF; j++) bcl = wj[j] == (bin >> 5) ? bsj[j] : bcl;
                        bc = __shfl_sync(FULL, bcl, bin & 31);
                        rel_w = __uint_as_float(m);
#pragma unroll
                        for (int j = 0; j < SPF; j++) {
                            const int bj = (wj[j] << 5) + lane;
                            if (bj >= bin - c.half_bw && bj <= bin + c.half_bw) exj[j] = false;
                        }
                    } else {
                        ArgMax best{-1.0f, 0x7fffffff};
                        int bslot = -1;
                        for (int i = lane; i < nc; i += 32) {
                            const int cbn = S.cbin[i];
                            if (cbn >= 0) {
                                const ArgMax cur2{S.crel[i], cbn};
                                const ArgMax nb = argmax_pick(best, cur2);
                                if (nb.i != best.i) bslot = i;
                                best = nb;
                            }
                        }
                        const ArgMax wbest = warp_argmax(best);
                        if (wbest.v < 0.0f) break;
                        bin = wbest.i;
                        rel_w = wbest.v;
                        const unsigned owner = __ballot_sync(FULL, best.i == bin && bslot >= 0);
                        bc = __shfl_sync(FULL, bslot >= 0 ? S.cbase[bslot] : 0.0f, __ffs(owner) - 1);
                        for (int i = lane; i < nc; i += 32) {
                            const int bb = S.cbin[i];
                            if (bb >= bin - c.half_bw && bb <= bin + c.half_bw) S.cbin[i] = -1;
                        }
                        __syncwarp();
                    }
                    if (have_mask == FULL) { bail = 5; break; }          // a 33rd concurrent burst
                    const int slot = __ffs(~have_mask) - 1;
                    if (lane == slot) {
                        r_have = true;
                        r_id = next_id;
                        r_start = idx - (unsigned long long)c.pre_len;
                        r_last0 = r_start;
                        r_cb = bin; r_rel = rel_w; r_base = bc;
                        b_dl = f + PF0; b_lah = NONE; b_tl = f + TLF;
                        set_window();
                    }
                    have_mask |= 1u << slot;
                    n_act++;
                    next_id += 10ull;
                    mask_burst(bin, false);
                }
                if (bail) break;
                mask_changed = true;
            }
            if (c.max_bursts > 0 && n_act > c.max_bursts) { bail = 6; break; }      // squelch (:593-631)
            if (sq > 0) sq--;                                 // :628-631
            if (n_act == 0) quiet_frames(f, f + 1);
            if (mask_changed) reload_fv();
            ST_TICK(cy_e4);
        }
        f++;
        ra += row_bytes;
        if (ra == ring_end_a) ra = ring_a;
    }
    // ---- wrap up
    if (!bail) {
        if (qs >= 0) { issue(qs, n_frames, 0); qs = -1; }
        if (!wait_all()) bail = 2;
    }
    issue(0, 0, 1);
    if (lane < SCL)
        while (ld_vol32(&ctl->done[lane]) < n_cmd - 1) __nanosleep(32);     // every real command is finished
    __syncwarp();
    // bulk copies still in flight (a bailed launch) must land before the shared memory is released
    for (; blk_landed < blk_issued; blk_landed++) mbar_wait_a(bars_a + 8u * (uint32_t)(blk_landed % RB), (uint32_t)((blk_landed / RB) & 1));
    if (!bail) {
        // the list of active bursts, in the reference's order (creation order = ascending id)
        int rank_a = 0;
        for (uint32_t m = have_mask; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const unsigned long long oid = ((unsigned long long)__shfl_sync(FULL, (unsigned)(r_id >> 32), src) << 32) |
                                           (unsigned long long)__shfl_sync(FULL, (unsigned)r_id, src);
            rank_a += (r_have && oid < r_id) ? 1 : 0;
        }
        if (r_have) {
            ActBurst b;
            b.id = r_id; b.start = r_start;
            b.last_active = b_lah == NONE ? r_last0 : index0 + (uint64_t)b_lah * (uint64_t)N;
            b.center_bin = r_cb; b.peak_rel = r_rel; b.base_at_create = r_base; b.pad = 0;
            gs->act[rank_a] = b;
        }
        if (lane == 0) {
            gs->hist_idx = hist_idx; gs->primed = primed; gs->n_act = n_act; gs->squelch_count = sq;
            gs->next_id = next_id; gs->index = index0 + (uint64_t)n_frames * (uint64_t)N;
            gs->n_gone = n_gone; gs->overflow = overflow;
        }
    }
    if (lane == 0) {
        ctl->bailed = bail ? 1 : 0;
        ctl->undo_frames = undo_frames;
        if (bail) { ctl->reason = bail; ctl->stats[1] += 1; ctl->stats[6] = (unsigned long long)f; } else ctl->stats[0] += 1;
        ctl->stats[2] += n_cmd; ctl->stats[3] += st_events; ctl->stats[4] += st_exact; ctl->stats[5] += st_waits;
        unsigned long long t_glob1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_glob1));
        ctl->stats[7] += t_glob1 - t_glob0;
        ctl->stats[8] += cy_scan; ctl->stats[9] += cy_e1; ctl->stats[10] += cy_wait; ctl->stats[11] += cy_e2;
        ctl->stats[12] += cy_e3; ctl->stats[13] += cy_e4; ctl->stats[14] += cy_ring; ctl->stats[15] += n_trips;
        ctl->stats[16] += st_mini;
        ctl->stats[1 + 16] += cy_p1; ctl->stats[2 + 16] += cy_p2; ctl->stats[3 + 16] += cy_p3;
    }
}

size_t stream_ctl_bytes() { return sizeof(StreamCtl); }

cudaError_t launch_detect_classify(const float *mag, const float *base, float thr, int N, int n_frames,
                                   uint32_t *xu, float *ref_out, int sm_count, cudaStream_t st, unsigned char *rowany) {
    if (n_frames <= 0) return cudaSuccess;
    if (rowany != nullptr) {
        cudaError_t e = cudaMemsetAsync(rowany, 0, (size_t)n_frames, st);
        if (e != cudaSuccess) return e;
    }
    const int ncol = N >> 10;
    // enough warps for every SM, not so many that the per-warp threshold setup dominates
    int parts = (sm_count * 16 + ncol - 1) / ncol;
    int fpw = (n_frames + parts - 1) / parts;
    if (fpw < 8) fpw = 8;
    parts = (n_frames + fpw - 1) / fpw;
    const int blocks = (parts * ncol + 3) / 4;
    k_detect_classify<<<blocks, 128, 0, st>>>(mag, base, thr, N, n_frames, fpw, xu, ref_out, rowany);
    return cudaGetLastError();
}

template <int BPT>
static cudaError_t launch_stream_t(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                   const uint32_t *xu, const float *ref, int n_frames, GoneBurst *gone,
                                   uint32_t gone_cap, StreamCtl *ctl, unsigned epoch, float *undo, float *base_snap, cudaStream_t st) {
    constexpr int RB = BPT >= 8 ? 4 : 8;
    size_t smem = ((sizeof(StShared) + 127) / 128) * 128 + (size_t)RB * SGF * (size_t)(c.N / 16) * sizeof(uint32_t);
    // IR_SCAN_EXCLUSIVE_SM=1: ask for all of the SM's shared memory so that no other kernel's CTAs are
    // co-scheduled on the cluster's SMs (measured: no effect on the leader's pace, so off by default)
    {
        const char *env = getenv("IR_SCAN_EXCLUSIVE_SM");
        if (env && *env == '1') smem = std::max<size_t>(smem, (size_t)227 * 1024);
    }
    cudaError_t e = cudaFuncSetAttribute(k_detect_scan_stream<BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_detect_scan_stream<BPT><<<SCL, SNT, smem, st>>>(c, state, base, hist, mag, xu, ref, n_frames, gone, gone_cap, ctl, epoch, undo, base_snap);
    return cudaGetLastError();
}

bool stream_scan_supported(const DetConfig &c) {
    const int bpt = c.N / (SCL * SWT);
    return c.N % (SCL * SWT) == 0 && (bpt == 2 || bpt == 4 || bpt == 8) && c.hist_size >= 2 * SGMAX;
}

// One launch of the streaming state machine over frames [0, n_frames) of `mag` (n_frames <=
// IR_STREAM_MAX_FRAMES).  xu == nullptr: priming launch (detector not primed, every frame quiet).
// The snapshot / restore / fallback around it is the caller's (pipeline.cu).
cudaError_t launch_detect_scan_stream(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                      const uint32_t *xu, const float *ref, int n_frames, GoneBurst *gone,
                                      uint32_t gone_cap, StreamCtl *ctl, unsigned epoch, float *undo, float *base_snap,
                                      cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    if (n_frames > IR_STREAM_MAX_FRAMES) return cudaErrorInvalidValue;
    switch (c.N / (SCL * SWT)) {
    case 2: return launch_stream_t<2>(c, state, base, hist, mag, xu, ref, n_frames, gone, gone_cap, ctl, epoch, undo, base_snap, st);
    case 4: return launch_stream_t<4>(c, state, base, hist, mag, xu, ref, n_frames, gone, gone_cap, ctl, epoch, undo, base_snap, st);
    case 8: return launch_stream_t<8>(c, state, base, hist, mag, xu, ref, n_frames, gone, gone_cap, ctl, epoch, undo, base_snap, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ir
