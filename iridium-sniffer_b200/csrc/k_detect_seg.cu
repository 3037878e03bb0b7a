// k_detect_seg.cu -- the burst state machine (burst_detect.c:426-632) WITHOUT a serial spine.
//
// The reference's frame loop is serial through two things: the list of active bursts and the
// noise baseline, which changes on "quiet" frames only (no burst active at the end of the frame,
// burst_detect.c:438-454).  k_detect_stream.cu walks the frames with one warp; here a chunk of up to
// IR_SEG_MAX_FRAMES frames is cut every IR_SEG_LEN frames and every segment is walked by a warp of
// its own, all at once.  Both dependencies are broken by speculation and the result is PROVED by a
// fixed point (tests/model_seg_scan.py is the executable model of the algorithm):
//
//   round r, three kernels:
//     k_seg_index  (1 CTA)        quiet flags of round r-1 -> list of quiet frames, baseline "version" of
//                                 every frame (= number of quiet frames before it), and a snapshot slot
//                                 for every version a frame with a set bitmap bit can ask for;
//     k_seg_base   (1 thread/bin) the reference's baseline recurrence (two roundings per quiet frame,
//                                 simd_avx2.c:221-236) over the quiet list; what a quiet frame replaces in
//                                 the 512-row history is the carried-in history for the first 512, else
//                                 the magnitudes of the chunk's own quiet frame 512 earlier -- no history
//                                 array is written during the rounds; snapshots B_v are stored for the
//                                 slots; every B_v is checked against the bitmaps' guard band;
//     k_seg_walk   (1 warp/segment) the streaming kernel's leader (bursts in lane registers, guard-banded
//                                 bitmaps by TMA into a shared-memory ring, events exact), started from the
//                                 burst list round r-1 left at the segment's first frame (round 0: empty;
//                                 segment 0: the detector's real list), taking exact baseline values from
//                                 the snapshots, and writing quiet flags + the list at its last frame.
//   A round whose outputs equal the previous round's is exact: by induction over the frames every
//   decision was taken from the true burst list and the true baseline.  Ids are (segment, ordinal) codes
//   until k_seg_commit turns them into the reference's running counter by a prefix sum over the
//   segments' creation counts; gone records are gathered in segment order.
//
// Anything unusual (squelch, a burst that may exceed max_burst_len, guard band violated, a 33rd
// concurrent burst, no fixed point within IR_SEG_ROUNDS) leaves the detector state untouched and sets
// `bailed`: the cluster kernel (k_detect_cluster.cu) then redoes the chunk.  A bail caused only by a wrong
// speculative start disappears in the next round.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ir_device.cuh"
#include "ir_internal.h"
#include "seg_generic.cuh"

namespace ir {

namespace {

// (small on purpose: a walker has to fit beside a resident FIR CTA -- 201 KB of an SM's shared memory -- or every
// launch of a round waits for an SM to drain; measured, the wait was half of a chunk's scan time)
constexpr int SGF = 4;                 // rows per ring block
constexpr int RB = 2;                  // ring blocks
constexpr int SPF = 4;                 // candidate words whose loads are in flight together
constexpr int SMAXW = 512;             // bitmap words per frame (N <= 16384)
template <int WPL> __host__ __device__ constexpr int smaxc() { return WPL >= 16 ? 512 : 128; }   // candidate peaks of one frame
constexpr uint32_t FULL = 0xffffffffu;
constexpr int NONE = -0x40000000;
constexpr int BIGF = 0x3fffffff;
constexpr unsigned long long CODE = 1ull << 63;

struct WkShared {
    unsigned long long bar[8];
    uint32_t fvs[SMAXW];
    unsigned short cw[SMAXW];
    int fslot[IR_SEG_LEN];
    SegGenOut gout;                    // the generic walker's result and its shared candidate counter
};
// (after it in shared memory: cbin / crel / cbase [smaxc<WPL>()], then the ring)

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
__device__ __forceinline__ uint32_t range_bits(int w, int lo, int hi) {
    int a = max(lo, w << 5), b = min(hi, (w << 5) + 31);
    if (a > b) return 0u;
    a &= 31; b &= 31;
    return (b == 31 ? FULL : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
}

// exclusive prefix sum of arr[0..n) in place (one CTA of 1024 threads); returns the total in every thread
__device__ int block_excl_scan(int *arr, int n, int *sh /* >= 33 ints */) {
    const int t = threadIdx.x, nt = blockDim.x;
    const int ipt = (n + nt - 1) / nt;
    const int a = min(t * ipt, n), b = min(a + ipt, n);
    int sum = 0;
    for (int i = a; i < b; i++) sum += arr[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    __syncthreads();
    if ((t & 31) == 31) sh[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int w = t < (nt >> 5) ? sh[t] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, wi, o);
            if (t >= o) wi += v;
        }
        sh[t] = wi - w;
        if (t == 31) sh[32] = wi;
    }
    __syncthreads();
    int run = sh[t >> 5] + incl - sum;
    for (int i = a; i < b; i++) { const int v = arr[i]; arr[i] = run; run += v; }
    const int total = sh[32];
    __syncthreads();
    return total;
}

}  // namespace

// =========================================================================== chunk begin
__global__ void __launch_bounds__(256)
k_seg_begin(DetConfig c, const DetState *__restrict__ gs, SegCtl *ctl, SegState *stA, SegState *stB, uint32_t *qw,
            int *ncreate, int *ngone, int *segbail, int *stch, uint32_t *valid_g, const float *__restrict__ ref,
            float *glo_g, float *ghi_g, SegBurst *ovfA, SegBurst *ovfB, uint32_t *seggen, int F, int S) {
    const int t = threadIdx.x;
    const int N = c.N, W = N >> 5;
    if (t == 0) {
        ctl->finished = 0; ctl->converged = 0; ctl->changed = 1; ctl->round = 0; ctl->cur = 0;
        ctl->hard_bail = 0; ctl->guard_bad = 0; ctl->bailed = 0;
        ctl->qfc[0] = 0; ctl->qfc[1] = 0; ctl->skip_base = 0; ctl->reclass = 0; ctl->reclass_prev = 0;
        ctl->F = F; ctl->S = S; ctl->nq = 0; ctl->n_slots = 0;
        ctl->index0 = gs->index; ctl->next_id0 = gs->next_id; ctl->sq0 = gs->squelch_count;
        ctl->n_gone0 = gs->n_gone; ctl->hist_idx0 = gs->hist_idx;
        if (!gs->primed) ctl->hard_bail = 1;
        if (gs->n_act > IR_SEG_LIST) ctl->hard_bail = 7;
        const int na = gs->n_act > IR_SEG_LIST ? 0 : gs->n_act;
        stA[0].n_act = na; stB[0].n_act = na;
    }
    const int na = gs->n_act > IR_SEG_LIST ? 0 : gs->n_act;
    for (int i = t; i < na; i += blockDim.x) {
        const ActBurst b = gs->act[i];
        const unsigned long long index0 = gs->index;
        SegBurst sb;
        sb.id = b.id; sb.start = b.start; sb.last0 = b.last_active; sb.cb = b.center_bin; sb.rel = b.peak_rel; sb.base = b.base_at_create;
        const long long d = (long long)(b.last_active + (unsigned long long)c.post_len) - (long long)index0;
        sb.dl = d <= 0 ? 0 : (int)min((long long)BIGF, (d + N - 1) / N);
        sb.lah = NONE;
        const long long tl = (long long)(b.start + (unsigned long long)c.max_burst_len) - (long long)index0;
        sb.tl = c.max_burst_len <= 0 ? BIGF : (tl < 0 ? -1 : (int)min((long long)BIGF, tl / N));
        if (i < 32) { stA[0].b[i] = sb; stB[0].b[i] = sb; }
        else { ovfA[i - 32] = sb; ovfB[i - 32] = sb; }
    }
    for (int s = 1 + t; s <= S; s += blockDim.x) { stA[s].n_act = 0; stB[s].n_act = 0; }
    for (int s = t; s < S; s += blockDim.x) { ncreate[s] = 0; ngone[s] = 0; segbail[s] = 0; seggen[s] = 0u; }
    // [parity][s]: the list at segment s's first frame changed in the round of that parity (segment 0's never does)
    for (int s = t; s < 2 * (S + 1); s += blockDim.x) stch[s] = (s == 0 || s == S + 1) ? 0 : 1;
    for (int w = t; w < (F + 31) / 32; w += blockDim.x) qw[w] = 0u;
    // the band the chunk's bitmaps were made for (k_detect_classify: thresholds at LO and HI times the reference)
    for (int bin = t; bin < N; bin += blockDim.x) {
        const float r = ref[bin];
        const float ga = r * IR_GUARD_LO, gb = r * IR_GUARD_HI;
        glo_g[bin] = fminf(ga, gb); ghi_g[bin] = fmaxf(ga, gb);
    }
    for (int w = t; w < W; w += blockDim.x) {
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) {
            const int bin = (w << 5) + b;
            const bool ok = bin >= c.half_bw && bin < N - c.half_bw && !(bin >= N / 2 - 3 && bin <= N / 2 + 3);
            v |= ok ? (1u << b) : 0u;
        }
        valid_g[w] = v;
    }
}

// =========================================================================== round: index
// qw: quiet flags of the previous round (bit f%32 of word f/32).  Outputs: qlist[q] = frame of the q-th
// quiet frame, slotv[v] = snapshot slot of baseline version v (or -1), fslot[f] = slot frame f reads.
__global__ void __launch_bounds__(256)
k_seg_index(SegCtl *ctl, const uint32_t *__restrict__ qw, const unsigned char *__restrict__ rowany, int *wpre,
            int *qlist, int *slotv, int *fslot, int slot_cap, int max_rounds) {
    __shared__ int sh[40];
    const int t = threadIdx.x;
    const int fin = ctl->finished, r = ctl->round, hb = ctl->hard_bail, ch = ctl->changed;
    __syncthreads();                                        // (thread 0 rewrites these below)
    if (fin) return;
    if (hb || (r > 0 && !ch) || r >= max_rounds) {
        if (t == 0) {
            // outputs of round r-1 equal those of round r-2: round r-1 was exact
            ctl->converged = (!hb && r > 0 && !ch) ? 1 : 0;
            ctl->finished = 1;
        }
        return;
    }
    const int F = ctl->F, nW = (F + 31) / 32;
    // no quiet flag changed in the previous round: list, versions, slots, snapshots, final baseline and the
    // guard verdict of that round all stand
    const bool keep = r > 0 && ctl->qfc[(r - 1) & 1] == 0x7fffffff && !ctl->reclass;
    __syncthreads();
    if (t == 0) {
        ctl->round = r + 1; ctl->cur = r & 1; ctl->changed = 0;
        ctl->reclass_prev = ctl->reclass; ctl->reclass = 0;
        ctl->qfc[r & 1] = 0x7fffffff;
        ctl->skip_base = keep ? 1 : 0;
        if (!keep) ctl->guard_bad = 0;
    }
    if (keep) return;
    // rows of the packed quiet frames that stand: the quiet frames before the first flag that changed last round
    const int fc_prev = r > 0 ? ctl->qfc[(r - 1) & 1] : 0;
    for (int w = t; w < nW; w += blockDim.x) {
        uint32_t v = qw[w];
        if (w == nW - 1 && (F & 31)) v &= (1u << (F & 31)) - 1u;
        wpre[w] = __popc(v);
    }
    __syncthreads();
    const int nq = block_excl_scan(wpre, nW, sh);
    for (int w = t; w < nW; w += blockDim.x) {
        uint32_t v = qw[w];
        if (w == nW - 1 && (F & 31)) v &= (1u << (F & 31)) - 1u;
        int q = wpre[w];
        while (v) { const int b = __ffs(v) - 1; v &= v - 1; qlist[q++] = (w << 5) + b; }
    }
    for (int v = t; v <= nq + 1; v += blockDim.x) slotv[v] = 0;
    __syncthreads();
    for (int f = t; f < F; f += blockDim.x)
        if (rowany[f]) {
            const int ver = wpre[f >> 5] + __popc(qw[f >> 5] & ((1u << (f & 31)) - 1u));
            slotv[ver] = 1;                                  // benign race: everybody stores 1
        }
    __syncthreads();
    // slotv: flag -> exclusive prefix; keep the flags in fslot's tail?  (no: re-derive from the prefix)
    // a version v is needed iff prefix[v+1] > prefix[v]; store slot or -1 afterwards
    const int n_slots = block_excl_scan(slotv, nq + 2, sh);   // (element nq+1 is a zero sentinel, see below)
    (void)n_slots;
    __syncthreads();
    for (int f = t; f < F; f += blockDim.x) {
        int sl = -1;
        if (rowany[f]) {
            const int ver = wpre[f >> 5] + __popc(qw[f >> 5] & ((1u << (f & 31)) - 1u));
            sl = slotv[ver];
        }
        fslot[f] = sl;
    }
    if (t == 0) {
        ctl->nq = nq;
        ctl->q_keep = (r > 0 && fc_prev < F) ? min(nq, wpre[fc_prev >> 5] + __popc(qw[fc_prev >> 5] & ((1u << (fc_prev & 31)) - 1u))) : 0;
        const int ns = slotv[nq + 1];
        ctl->n_slots = ns;
        if (ns > slot_cap) ctl->hard_bail = 11;
    }
}

// =========================================================================== round: baseline
// One thread per bin.  slotp[v] (from k_seg_index) is the exclusive prefix of "version v needs a snapshot":
// version v is stored at slot slotp[v] iff slotp[v+1] > slotp[v].
// The quiet frames' magnitude rows, packed in list order.  The baseline recurrence walks them one after the other,
// all bins in step; read where they lie (rows 4*N bytes wide, a dozen frames apart) every step lands on a
// new page for all CTAs at once and the pass runs at the pace of the page walks (measured: 155 us for 1300 quiet
// frames whatever the prefetch depth).  Here every CTA takes different rows, so the misses overlap.
// qbuf rows: [0, H) the carried-in history in replacement order (row j = history row (h0+j)%H, copied once per
// chunk by k_seg_gather_hist), [H + q] the chunk's q-th quiet frame.  Quiet frame q then replaces row q: the
// "old" row of the recurrence is always H rows before the "new" one.
__global__ void __launch_bounds__(256)
k_seg_gather(DetConfig c, const SegCtl *ctl, const float *__restrict__ mag, const int *__restrict__ qlist, float *__restrict__ qbuf) {
    if (ctl->finished || ctl->hard_bail || ctl->skip_base) return;
    const int n4 = c.N >> 2, nq = ctl->nq;
    for (int q = ctl->q_keep + blockIdx.x; q < nq; q += gridDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(mag + (size_t)qlist[q] * c.N);
        float4 *dst = reinterpret_cast<float4 *>(qbuf + (size_t)(c.hist_size + q) * c.N);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}
__global__ void __launch_bounds__(256)
k_seg_gather_hist(DetConfig c, const DetState *__restrict__ gs, const float *__restrict__ hist, float *__restrict__ qbuf) {
    const int n4 = c.N >> 2, H = c.hist_size, h0 = gs->hist_idx;
    for (int j = blockIdx.x; j < H; j += gridDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(hist + (size_t)((h0 + j) % H) * c.N);
        float4 *dst = reinterpret_cast<float4 *>(qbuf + (size_t)j * c.N);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}

// One warp per 128 bins, four adjacent bins per thread: four independent dependency chains per thread (the
// recurrence is two dependent adds per bin and quiet frame -- with one or two warps per SM the pass runs at the
// pace of a single warp's instruction stream, so it has to be short and carry its own parallelism).  Rows arrive
// through a ring of 16-byte asynchronous copies; a thread reads back only what it copied: no barrier anywhere.
constexpr int BU = 4;                  // quiet frames per batch (one cp.async group)
constexpr int BD = 4;                  // batches in flight (ring: 16 KB, see the note on the walkers' ring)
constexpr int BT = 32;                 // threads per CTA
constexpr int BB = 4 * BT;             // bins per CTA
constexpr int BQ = 1024;               // snapshot flags staged in shared memory at a time
struct BaseShared {
    float4 ring[BD][BU][2][BT];        // [stage][frame of the batch][new | old][thread]
    int s_slot[BQ];
};
__global__ void __launch_bounds__(BT)
k_seg_base(DetConfig c, SegCtl *ctl, const float *__restrict__ base_g, const float *__restrict__ qbuf,
           float *__restrict__ glo_g, float *__restrict__ ghi_g,
           const int *__restrict__ slotp, float *__restrict__ snap, float *__restrict__ bfinal) {
    extern __shared__ __align__(16) unsigned char base_smem[];
    BaseShared &S = *reinterpret_cast<BaseShared *>(base_smem);
    if (ctl->finished || ctl->hard_bail || ctl->skip_base) return;
    const int N = c.N, H = c.hist_size;
    const int t = threadIdx.x;
    const int bin = blockIdx.x * BB + 4 * t;
    const int nq = ctl->nq;
    float base[4], glo[4], ghi[4], bmin[4], bmax[4];
    {
        const float4 b4 = *reinterpret_cast<const float4 *>(base_g + bin);
        const float4 l4 = *reinterpret_cast<const float4 *>(glo_g + bin), h4 = *reinterpret_cast<const float4 *>(ghi_g + bin);
        base[0] = b4.x; base[1] = b4.y; base[2] = b4.z; base[3] = b4.w;
        glo[0] = l4.x; glo[1] = l4.y; glo[2] = l4.z; glo[3] = l4.w;
        ghi[0] = h4.x; ghi[1] = h4.y; ghi[2] = h4.z; ghi[3] = h4.w;
#pragma unroll
        for (int j = 0; j < 4; j++) { bmin[j] = base[j]; bmax[j] = base[j]; }
        const int p0 = slotp[0], p1 = slotp[1];
        if (p1 > p0) *reinterpret_cast<float4 *>(snap + (size_t)p0 * N + bin) = b4;
    }
    // fetch cursor: the next quiet frame's row; the row it replaces lies H rows before (see k_seg_gather).  Past
    // the end the copies read slack rows of the buffer that nobody applies.
    const size_t back = (size_t)H * N;
    const float *pn = qbuf + back + bin;
    auto fetch = [&](int k) {                                  // batch k: quiet frames k*BU .. k*BU+BU
        const uint32_t sa = smem_u32(&S.ring[k % BD][0][0][t]);
#pragma unroll
        for (int u = 0; u < BU; u++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + (uint32_t)(u * 2 * BT * 16)), "l"(pn) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + (uint32_t)((u * 2 + 1) * BT * 16)), "l"(pn - back) : "memory");
            pn += N;
        }
        cp_async_commit();
    };
    const int nb = (nq + BU - 1) / BU;
#pragma unroll
    for (int d = 0; d < BD - 1; d++) fetch(d);
    for (int k = 0; k < nb; k++) {
        if ((k * BU) % BQ == 0) {                               // snapshot flags of the next BQ quiet frames
            __syncwarp();
            for (int i = t; i < BQ; i += BT) {
                const int q = k * BU + i;
                int sl = -1;
                if (q < nq) { const int pa = slotp[q + 1], pb = slotp[q + 2]; sl = pb > pa ? pa : -1; }
                S.s_slot[i] = sl;
            }
            __syncwarp();
        }
        fetch(k + BD - 1);
        cp_async_wait_group<BD - 1>();                          // batch k has landed (this thread's own copies)
        const int stage = k % BD;
#pragma unroll
        for (int u = 0; u < BU; u++) {
            const int q = k * BU + u;
            if (q < nq) {
                const float4 nv = S.ring[stage][u][0][t], ov = S.ring[stage][u][1][t];
                const float nn[4] = {nv.x, nv.y, nv.z, nv.w}, oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float tt = base[j] - oo[j];              // simd_avx2.c:221-236: two roundings
                    base[j] = tt + nn[j];
                    bmin[j] = fminf(bmin[j], base[j]); bmax[j] = fmaxf(bmax[j], base[j]);
                }
                const int sl = S.s_slot[q % BQ];
                if (sl >= 0) *reinterpret_cast<float4 *>(snap + (size_t)sl * N + bin) = make_float4(base[0], base[1], base[2], base[3]);
            }
        }
    }
    cp_async_wait_all();
    *reinterpret_cast<float4 *>(bfinal + bin) = make_float4(base[0], base[1], base[2], base[3]);
    // The bitmaps hold for baselines inside [glo, ghi] (per bin).  A bin whose baseline left its band -- typically a
    // channel that carried a burst while the detector was priming, whose inflated baseline collapses when those
    // rows rotate out of the history -- gets a band that covers what it really did (with a margin, so that the
    // small shifts of later rounds stay inside), and k_seg_reclass rebuilds the bitmaps before anybody walks them.
    int bad = 0, widen = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {                               // (a NaN is sticky under the adds: it reaches the final value)
        if (!(base[j] == base[j]) || !(bmin[j] > 0.0f) || !(bmax[j] < 3.0e38f)) {
            if (!(bmin[j] >= glo[j] && bmax[j] <= ghi[j])) bad = 1;          // nothing a positive finite band can cover
        } else if (!(bmin[j] >= glo[j] && bmax[j] <= ghi[j])) {
            glo[j] = fminf(glo[j], bmin[j] * 0.9f); ghi[j] = fmaxf(ghi[j], bmax[j] * 1.1f);
            widen = 1;
        }
    }
    if (widen) {
        *reinterpret_cast<float4 *>(glo_g + bin) = make_float4(glo[0], glo[1], glo[2], glo[3]);
        *reinterpret_cast<float4 *>(ghi_g + bin) = make_float4(ghi[0], ghi[1], ghi[2], ghi[3]);
        atomicOr(&ctl->reclass, 1u);
    }
    if (bad) atomicOr(&ctl->guard_bad, 1u);
}

// Bitmaps of the whole chunk again, against the per-bin bands k_seg_base widened (a no-op unless it did, this
// round).  The XU band only grows downwards, so "the row has a bit" flags are only ever added.
__global__ void __launch_bounds__(128)
k_seg_reclass(SegCtl *ctl, const float *__restrict__ mag, const float *__restrict__ glo_g, const float *__restrict__ ghi_g,
              float thr, int N, int frames_per_warp, uint32_t *__restrict__ xu, unsigned char *__restrict__ rowany) {
    if (ctl->finished || ctl->hard_bail || !ctl->reclass) return;
    const int n_frames = ctl->F;
    const int lane = threadIdx.x & 31;
    const int gw = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int ncol = N >> 10;
    const int col = gw % ncol, part = gw / ncol;
    const int f0 = part * frames_per_warp;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->changed = 1; ctl->stats[7] += 1; }
    if (f0 >= n_frames) return;
    const int f1 = min(f0 + frames_per_warp, n_frames);
    const int W = N >> 5;
    float thi[32], tlo[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int bin = (col << 10) + (j << 5) + lane;
        // certainly above needs a positive baseline all along; possibly above needs one at some point
        const float gl = glo_g[bin], gh = ghi_g[bin];
        thi[j] = gl > 0.0f ? thr * gh * 1.0001f : __int_as_float(0x7f800000);
        tlo[j] = gh > 0.0f ? thr * gl * 0.9999f : __int_as_float(0x7f800000);
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
            const bool un = !(m[j] < tlo[j]);
            const uint32_t bx = __ballot_sync(FULL, hi), bu = __ballot_sync(FULL, un);
            if (lane == j) { xw = bx; uw = bu; }
        }
        uint32_t *o = xu + (size_t)f * (2 * W) + (col << 5) + lane;
        o[0] = uw;
        o[W] = xw;
        if (__any_sync(FULL, uw != 0u) && lane == 0) rowany[f] = 1;
    }
}

#ifdef IR_SEG_TIMING
// cycles (lane 0): generic walker phases 0 stage 1 scan 2 quiet update 3 hits 4 peaks 5 delete 6 create; 7 generic walks whole,
// 8 the fast walker's attempt before them, 9 how many; 10 fast walks whole, 11 how many; 12 candidates, 13 creations, 14 event frames
__device__ unsigned long long g_seg_prof[16];
void seg_timing_dump() {
    unsigned long long h[16], z[16] = {0};
    cudaMemcpyFromSymbol(h, g_seg_prof, sizeof(h));
    cudaMemcpyToSymbol(g_seg_prof, z, sizeof(z));
    fprintf(stderr, "seg timing (Mcycles): generic walks %llu: stage %.2f scan %.2f quiet %.2f hits %.2f peaks %.2f delete %.2f create %.2f | whole %.2f "
                    "fast attempt before %.2f | fast walks %llu whole %.2f | generic start lists %llu bursts\n",
            h[9], h[0] * 1e-6, h[1] * 1e-6, h[2] * 1e-6, h[3] * 1e-6, h[4] * 1e-6, h[5] * 1e-6, h[6] * 1e-6, h[7] * 1e-6, h[8] * 1e-6,
            h[11], h[10] * 1e-6, h[12]);
}
#endif
// the 32 lanes of a segment's warp, for seg_generic.cuh
struct SegLanesWarp {
    static constexpr int L = 32;
    uint32_t *rowbuf;                                         // shared memory for one bitmap row
#ifdef IR_SEG_TIMING
    long long *tacc, *tlast;
    __device__ void tick(int k) const { const long long t = clock64(); if (k >= 0) tacc[k] += t - *tlast; *tlast = t; }
#else
    __device__ void tick(int) const {}
#endif
    __device__ const uint32_t *stage(const uint32_t *row, int words) const {
        __syncwarp();                                         // the previous row's readers are done
        for (int w = (int)(threadIdx.x & 31u) * 4; w < words; w += 128)
            *reinterpret_cast<uint4 *>(rowbuf + w) = __ldg(reinterpret_cast<const uint4 *>(row + w));
        __syncwarp();
        return rowbuf;
    }
    __device__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ void sync() const { __syncwarp(); }
    __device__ bool any(bool p) const { return __any_sync(FULL, p) != 0; }
    __device__ int sum(int v) const { return __reduce_add_sync(FULL, v); }
    __device__ int max(int v) const { return __reduce_max_sync(FULL, v); }
    __device__ uint32_t ballot(bool p) const { return __ballot_sync(FULL, p); }
    __device__ void and_word(uint32_t *p, uint32_t m) const { atomicAnd(p, m); }
    __device__ int excl_scan(int v, int &total) const {
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, x, o);
            if ((int)(threadIdx.x & 31u) >= o) x += y;
        }
        total = __shfl_sync(FULL, x, 31);
        return x - v;
    }
    __device__ unsigned long long max64(unsigned long long v) const {
        const uint32_t hi = __reduce_max_sync(FULL, (uint32_t)(v >> 32));
        const uint32_t lo = __reduce_max_sync(FULL, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0u);
        return ((unsigned long long)hi << 32) | lo;
    }
};

// =========================================================================== round: walk
template <int WPL>
__global__ void __launch_bounds__(32, 1)
k_seg_walk(DetConfig c, SegCtl *ctl, const float *__restrict__ mag_c, const uint32_t *__restrict__ xu_c,
           const float *__restrict__ snap, const int *__restrict__ fslot_g, const uint32_t *__restrict__ valid_g,
           SegState *stA, SegState *stB, uint32_t *qw_g, int *ncreate, int *ngone, int *segbail, int *stch,
           GoneBurst *glist, SegBurst *ovfA, SegBurst *ovfB, unsigned long long *gkeys, uint32_t *seggen) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WkShared &S = *reinterpret_cast<WkShared *>(smem_raw);
    if (ctl->finished || ctl->hard_bail) return;
    const int seg = (int)blockIdx.x;
    const int N = c.N, W = N >> 5;
    const int F = ctl->F;
    const int f0 = seg * IR_SEG_LEN;
    if (f0 >= F) return;
    const int n_frames = min(IR_SEG_LEN, F - f0);
    const int rnd = ctl->round - 1, par = rnd & 1;            // this round (k_seg_index counted it already)
    const SegState *cur = par ? stB : stA;
    SegState *nxt = par ? stA : stB;
    // bursts 32.. of a list (a segment cut with more than 32 alive: only the generic walker makes or reads those)
    const SegBurst *ovf_cur = par ? ovfB : ovfA;
    SegBurst *ovf_nxt = par ? ovfA : ovfB;
    auto list_get = [&](const SegState *st, const SegBurst *ov, int sgi, int i) -> SegBurst {
        return i < 32 ? st[sgi].b[i] : ov[(size_t)sgi * IR_SEG_OVF + (i - 32)];
    };
    auto list_put = [&](SegState *st, SegBurst *ov, int sgi, int i, const SegBurst &b) {
        if (i < 32) st[sgi].b[i] = b; else ov[(size_t)sgi * IR_SEG_OVF + (i - 32)] = b;
    };
    const int Sp1 = ctl->S + 1;
    const int *stch_prev = stch + (par ^ 1) * Sp1;            // [s]: the list at segment s's first frame changed last round
    int *stch_now = stch + par * Sp1;
    const int lane = threadIdx.x & 31;
    // Same inputs as last round -- the same burst list at the first frame, and no quiet flag changed before the
    // last frame (so every baseline version this segment reads is the same) -- give the same outputs: keep them.
    // (bitmaps rebuilt this round or the last: the snapshot slots moved too; a segment that gave up: its reason may be gone)
    if (rnd > 0 && !ctl->reclass && !ctl->reclass_prev && segbail[seg] == 0 && stch_prev[seg] == 0 &&
        f0 + n_frames <= ctl->qfc[par ^ 1]) {
        const int on = cur[seg + 1].n_act;
        for (int i = lane; i < on; i += 32) list_put(nxt, ovf_nxt, seg + 1, i, list_get(cur, ovf_cur, seg + 1, i));
        if (lane == 0) { nxt[seg + 1].n_act = on; stch_now[seg + 1] = 0; }
        return;
    }
#ifdef IR_SEG_TIMING
    const long long t_in = clock64();
    long long t_fast = 0, t_gen = 0, tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
#endif
    const float *mag = mag_c + (size_t)f0 * N;
    GoneBurst *gl = glist + (size_t)seg * IR_SEG_GONE;
    const float thr = c.thr;
    uint64_t *bars = reinterpret_cast<uint64_t *>(S.bar);
    constexpr int SMAXC = smaxc<WPL>();
    int *s_cbin = reinterpret_cast<int *>(smem_raw + ((sizeof(WkShared) + 15) / 16) * 16);
    float *s_crel = reinterpret_cast<float *>(s_cbin + SMAXC), *s_cbase = s_crel + SMAXC;
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + (((sizeof(WkShared) + 15) / 16) * 16 + 3 * SMAXC * 4 + 127) / 128 * 128);
    const int RW = 2 * W;                                     // words per row: [XU][X]
    const uint32_t row_bytes = (uint32_t)RW * sizeof(uint32_t);
    const uint32_t *xu = xu_c + (size_t)f0 * RW;
    constexpr int RROWS = RB * SGF;
    int n_act = cur[seg].n_act;
    int sq = max(ctl->sq0 - f0, 0);
    const long long index0 = (long long)ctl->index0 + (long long)f0 * N;     // sample index of the segment's first frame
    uint32_t ng_local = 0, nc_local = 0;
    int bail = 0;
    unsigned long long st_events = 0;
    const int PF = (c.post_len + N - 1) / N;
    const int PF0 = max(1, (c.post_len - c.pre_len + N - 1) / N);
    const int TLF = c.max_burst_len <= 0 ? 0x20000000 : (c.max_burst_len >= c.pre_len ? (c.max_burst_len - c.pre_len) / N : -1);

    // ---- the burst of this lane and its times in frames of this SEGMENT
    bool r_have = false;
    unsigned long long r_id = 0, r_start = 0, r_last0 = 0;
    int r_cb = 0;
    float r_rel = 0.0f, r_base = 0.0f;
    int b_dl = BIGF, b_lah = NONE, b_tl = BIGF;
    uint32_t b_o0 = 0, b_o1 = 0, b_msk = 0;
    int b_sh = 0;
    auto set_window = [&]() {
        const int w0 = (r_cb - 1) >> 5;
        b_o0 = (uint32_t)w0 * 4u; b_o1 = (uint32_t)min(w0 + 1, W - 1) * 4u; b_sh = (r_cb - 1) & 31; b_msk = 7u;
    };
    const int n_start = n_act;
    if (n_act > 32 || seggen[seg]) { bail = 7; n_act = 0; }   // more than the lanes hold (now, or inside the segment last round): the generic walker's
    if (lane < n_act) {
        const SegBurst b = cur[seg].b[lane];
        r_have = true;
        r_id = b.id; r_start = b.start; r_last0 = b.last0; r_cb = b.cb; r_rel = b.rel; r_base = b.base;
        b_dl = b.dl - f0; b_lah = b.lah == NONE ? NONE : b.lah - f0; b_tl = b.tl - f0;
        set_window();
    }
    uint32_t have_mask = __ballot_sync(FULL, r_have);
    for (int w = lane; w < W; w += 32) S.fvs[w] = valid_g[w];
    for (int i = lane; i < n_frames; i += 32) S.fslot[i] = fslot_g[f0 + i];
    if (lane == 0) {
        for (int d = 0; d < RB; d++) mbar_init(&bars[d], 1);
        fence_mbar_init();
    }
    __syncwarp();
    auto mask_bins = [&](int lo, int hi, bool set) {
        const int w = (lo >> 5) + lane;
        if (lane < 3 && w <= (hi >> 5)) {
            const uint32_t m = range_bits(w, lo, hi);
            S.fvs[w] = set ? (S.fvs[w] | (m & __ldg(valid_g + w))) : (S.fvs[w] & ~m);
        }
        __syncwarp();
    };
    auto mask_burst = [&](int cb, bool set) { mask_bins(max(cb - c.half_bw, 0), min(cb + c.half_bw, N - 1), set); };
    for (uint32_t hm = have_mask; hm; hm &= hm - 1) mask_burst(__shfl_sync(FULL, r_cb, __ffs(hm) - 1), false);
    uint32_t fv[WPL];
    auto reload_fv = [&]() {
        __syncwarp();
#pragma unroll
        for (int k4 = 0; k4 < WPL / 4; k4++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(&S.fvs[lane * WPL + 4 * k4]);
            fv[4 * k4] = v.x; fv[4 * k4 + 1] = v.y; fv[4 * k4 + 2] = v.z; fv[4 * k4 + 3] = v.w;
        }
    };
    reload_fv();

    // ring of bitmap rows, a block (SGF rows, one bulk copy, one mbarrier) at a time
    const int n_blocks = (n_frames + SGF - 1) / SGF;
    int blk_issued = 0, blk_landed = 0;
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

    // quiet flags of this segment: lane l holds frames [32l, 32l+32)
    uint32_t qbits = 0;
    int qs = -1, qe = -1;
    auto flush_quiet = [&]() {
        if (qs >= 0) {
            const int lo = max(qs, lane << 5), hi = min(qe, (lane << 5) + 32);
            if (lo < hi) {
                const int a = lo & 31, n = hi - lo;
                qbits |= (n == 32 ? FULL : ((1u << n) - 1u)) << a;
            }
            qs = -1;
        }
    };
    auto quiet_frames = [&](int a, int b) {
        if (qs >= 0 && a != qe) flush_quiet();
        if (qs < 0) qs = a;
        qe = b;
    };

    // delete_gone_bursts (:490-518) for the lanes in dmask on frame fr
    auto delete_done = [&](uint32_t dmask, bool done, int fr) {
        int rank_d = 0;
        if (dmask & (dmask - 1)) {
            for (uint32_t m = dmask; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                const unsigned long long oid = ((unsigned long long)__shfl_sync(FULL, (unsigned)(r_id >> 32), src) << 32) |
                                               (unsigned long long)__shfl_sync(FULL, (unsigned)r_id, src);
                rank_d += oid < r_id ? 1 : 0;
            }
        }
        if (done) {
            const uint32_t slot_g = ng_local + (uint32_t)rank_d;
            if (slot_g < (uint32_t)IR_SEG_GONE) {
                GoneBurst g;
                g.id = r_id; g.start = r_start; g.stop = (unsigned long long)(index0 + (long long)fr * N);
                g.last_active = b_lah == NONE ? r_last0 : (unsigned long long)(index0 + (long long)b_lah * N);
                g.center_bin = r_cb; g.peak_rel = r_rel; g.base_at_create = r_base; g.pad = 0;
                gl[slot_g] = g;
            } else {
                bail = 10;
            }
            r_have = false;
            b_dl = BIGF; b_lah = NONE; b_tl = BIGF; b_msk = 0; b_o0 = 0; b_o1 = 0; b_sh = 0;
        }
        bail = __any_sync(FULL, bail == 10) ? 10 : bail;
        ng_local += (uint32_t)__popc(dmask);
        n_act -= __popc(dmask);
        have_mask &= ~dmask;
        for (uint32_t m = dmask; m; m &= m - 1) mask_burst(__shfl_sync(FULL, r_cb, __ffs(m) - 1), true);
        for (uint32_t m = dmask; m; m &= m - 1) {
            const int cbd = __shfl_sync(FULL, r_cb, __ffs(m) - 1);
            for (uint32_t nm = __ballot_sync(FULL, r_have && abs(r_cb - cbd) <= 2 * c.half_bw); nm; nm &= nm - 1)
                mask_burst(__shfl_sync(FULL, r_cb, __ffs(nm) - 1), false);
        }
    };

    struct FrameRegs { uint32_t xu[WPL]; uint32_t bx0, bx1, bu0, bu1; };
    const uint32_t my_off = (uint32_t)(lane * WPL) * 4u, x_off = (uint32_t)W * 4u;
    auto load_frame = [&](FrameRegs &r, uint32_t ra) {
#pragma unroll
        for (int k4 = 0; k4 < WPL / 4; k4++) {
            const uint4 v = lds128(ra + my_off + 16u * k4);
            r.xu[4 * k4] = v.x; r.xu[4 * k4 + 1] = v.y; r.xu[4 * k4 + 2] = v.z; r.xu[4 * k4 + 3] = v.w;
        }
        r.bu0 = lds32(ra + b_o0); r.bu1 = lds32(ra + b_o1);
        r.bx0 = lds32(ra + x_off + b_o0); r.bx1 = lds32(ra + x_off + b_o1);
    };

    int f = 0;
    FrameRegs cur_r;
    uint32_t ra = ring_a;
#pragma unroll 1
    while (f < n_frames && !bail) {
        // ---- (A) uneventful frames, two at a time
#pragma unroll 1
        for (;;) {
            if ((f & (SGF - 1)) == 0) ring_advance(f, f);
            if ((f & (SGF - 1)) == SGF - 1 || f + 1 >= n_frames) break;
            FrameRegs c1;
            load_frame(cur_r, ra);
            load_frame(c1, ra + row_bytes);
            uint32_t a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < WPL; k++) { a0 |= cur_r.xu[k] & fv[k]; a1 |= c1.xu[k] & fv[k]; }
            const bool h0 = (__funnelshift_r(cur_r.bx0, cur_r.bx1, b_sh) & b_msk) != 0u;
            const bool h1 = (__funnelshift_r(c1.bx0, c1.bx1, b_sh) & b_msk) != 0u;
            const bool q0 = (__funnelshift_r(cur_r.bu0, cur_r.bu1, b_sh) & b_msk) != 0u;
            const bool q1 = (__funnelshift_r(c1.bu0, c1.bu1, b_sh) & b_msk) != 0u;
            const int dl1 = h0 ? f + PF : b_dl;
            const bool e0 = a0 != 0u || (!h0 && (q0 || f >= b_dl)) || f > b_tl;
            const bool e1 = a1 != 0u || (!h1 && (q1 || f + 1 >= dl1)) || f + 1 > b_tl;
            const uint32_t em = __ballot_sync(FULL, e0) ? 1u : (__ballot_sync(FULL, e1) ? 2u : 0u);
            if (em == 1u) {
                const bool d0 = !h0 && f >= b_dl;
                if (__any_sync(FULL, a0 != 0u || (!h0 && q0) || f > b_tl)) break;
                st_events++;
                if (h0) { b_dl = f + PF; b_lah = f; }
                delete_done(__ballot_sync(FULL, d0), d0, f);
                if (bail) break;
                reload_fv();
                if (sq > 0) sq--;
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
            sq = max(sq - adv, 0);
            if (n_act == 0) quiet_frames(f, f + adv);
            f += adv;
            ra += row_bytes * (uint32_t)adv;
            if (ra == ring_end_a) ra = ring_a;
            if (f >= n_frames) break;
        }
        if (f >= n_frames || bail) break;
        // ---- (B) one frame, any case
        load_frame(cur_r, ra);
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < WPL; k++) acc |= cur_r.xu[k] & fv[k];
        const uint32_t x3 = __funnelshift_r(cur_r.bx0, cur_r.bx1, b_sh) & b_msk;
        const uint32_t u3 = __funnelshift_r(cur_r.bu0, cur_r.bu1, b_sh) & b_msk;
        bool hit = x3 != 0u;
        const bool ev = acc != 0u || (!hit && (u3 != 0u || f >= b_dl)) || f > b_tl;
        if (!__any_sync(FULL, ev)) {
            if (hit) { b_dl = f + PF; b_lah = f; }
            sq = max(sq - 1, 0);
            if (n_act == 0) quiet_frames(f, f + 1);
        } else {
            // ================= event frame f: the reference's steps, exactly
            st_events++;
            const bool cand_any = __any_sync(FULL, acc != 0u);
            const bool unc_any = __any_sync(FULL, !hit && u3 != 0u);
            if (__any_sync(FULL, f > b_tl)) { bail = 3; break; }
            const float *row = mag + (size_t)f * N;
            const int slot = S.fslot[f];
            if ((cand_any || unc_any) && slot < 0) { bail = 9; break; }      // (cannot happen: the row has a set bit)
            const float *Bv = snap + (size_t)max(slot, 0) * N;                  // the exact baseline of this frame
            int n_cw = 0;
            float mvp[SPF], bsp[SPF];
            int wj[SPF];
#pragma unroll
            for (int j = 0; j < SPF; j++) { mvp[j] = 0.0f; bsp[j] = 0.0f; wj[j] = -1; }
            if (cand_any) {
                uint32_t wmask = 0;
#pragma unroll
                for (int k = 0; k < WPL; k++) wmask |= (cur_r.xu[k] & fv[k]) ? (1u << k) : 0u;
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
                        if (n_cw >= SPF && lane == 0) S.cw[n_cw] = (unsigned short)w;
                        n_cw++;
                    }
                }
                if (n_cw > SPF && lane == 0) {
#pragma unroll
                    for (int j = 0; j < SPF; j++) S.cw[j] = (unsigned short)wj[j];
                }
#pragma unroll
                for (int j = 0; j < SPF; j++)
                    if (wj[j] >= 0) {
                        mvp[j] = __ldg(row + (wj[j] << 5) + lane);
                        bsp[j] = __ldg(Bv + (wj[j] << 5) + lane);
                    }
            }
            __syncwarp();
            // update_bursts: the tests an X bit did not decide
            if (unc_any && !hit) {
                uint32_t q = u3;
                while (q) {
                    const int b = r_cb - 1 + __ffs(q) - 1;
                    q &= q - 1;
                    const float bs = __ldg(Bv + b);
                    if (bs > 0.0f && __ldg(row + b) / bs > thr) hit = true;
                }
            }
            if (hit) { b_dl = f + PF; b_lah = f; }
            const bool done = !hit && f >= b_dl;
            const uint32_t dmask = __ballot_sync(FULL, done);
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
                    if (j0 > 0) wj[j] = j0 + j < n_cw ? (int)S.cw[j0 + j] : -1;
                    if (wj[j] >= 0) {
                        const int bin = (wj[j] << 5) + lane;
                        mv[j] = j0 == 0 ? mvp[j] : __ldg(row + bin);
                        bsj[j] = j0 == 0 ? bsp[j] : __ldg(Bv + bin);
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
                            if (pos < SMAXC) { s_cbin[pos] = (wj[j] << 5) + lane; s_crel[pos] = relj[j]; s_cbase[pos] = bsj[j]; }
                        }
                        n_cand += __popc(bal);
                    }
                }
            }
            if (n_cand > SMAXC) { bail = 4; break; }
            __syncwarp();
            bool mask_changed = false;
            if (dmask) {
                delete_done(dmask, done, f);
                if (bail) break;
                mask_changed = true;
            }
            if (n_cand > 0) {
                const int nc = n_cand;
                const unsigned long long start_new = (unsigned long long)(index0 + (long long)f * N - (long long)c.pre_len);
#pragma unroll 1
                for (;;) {
                    int bin;
                    float rel_w, bc;
                    if (fast) {
                        uint32_t key = 0;
                        int kb = 0x7fffffff;
#pragma unroll
                        for (int j = 0; j < SPF; j++) {
                            const uint32_t kj = exj[j] ? __float_as_uint(relj[j]) : 0u;
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
                            const int cbn = s_cbin[i];
                            if (cbn >= 0) {
                                const ArgMax cur2{s_crel[i], cbn};
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
                        bc = __shfl_sync(FULL, bslot >= 0 ? s_cbase[bslot] : 0.0f, __ffs(owner) - 1);
                        for (int i = lane; i < nc; i += 32) {
                            const int bb = s_cbin[i];
                            if (bb >= bin - c.half_bw && bb <= bin + c.half_bw) s_cbin[i] = -1;
                        }
                        __syncwarp();
                    }
                    if (have_mask == FULL) { bail = 5; break; }
                    const int slot_l = __ffs(~have_mask) - 1;
                    if (lane == slot_l) {
                        r_have = true;
                        r_id = CODE | ((unsigned long long)seg << 32) | (unsigned long long)nc_local;
                        r_start = start_new;
                        r_last0 = r_start;
                        r_cb = bin; r_rel = rel_w; r_base = bc;
                        b_dl = f + PF0; b_lah = NONE; b_tl = f + TLF;
                        set_window();
                    }
                    have_mask |= 1u << slot_l;
                    n_act++;
                    nc_local++;
                    mask_burst(bin, false);
                }
                if (bail) break;
                mask_changed = true;
            }
            if (c.max_bursts > 0 && n_act > c.max_bursts) { bail = 6; break; }
            if (sq > 0) sq--;
            if (n_act == 0) quiet_frames(f, f + 1);
            if (mask_changed) reload_fv();
        }
        f++;
        ra += row_bytes;
        if (ra == ring_end_a) ra = ring_a;
    }
    // bulk copies still in flight must land before the shared memory is released
    for (; blk_landed < blk_issued; blk_landed++) mbar_wait_a(bars_a + 8u * (uint32_t)(blk_landed % RB), (uint32_t)((blk_landed / RB) & 1));
    // ---- more than 32 bursts alive somewhere in this segment: walk it again the plain way (seg_generic.cuh), from
    // the same start list, on lane 0; the other lanes wait.  Its outputs replace the fast walker's.
#ifdef IR_SEG_TIMING
    t_fast = clock64() - t_in;
#endif
    bool generic = false;
    SegBurst *work = reinterpret_cast<SegBurst *>(ring);      // (the ring is drained: its shared memory holds the list and a bitmap row)
    if (bail == 4 || bail == 5 || bail == 7) {
        generic = true;
        if (lane == 0) seggen[seg] = 1u;                     // next round: straight here
        for (int i = lane; i < n_start && i < IR_SEG_LIST; i += 32) work[i] = list_get(cur, ovf_cur, seg, i);
        __syncwarp();
        bail = 5;
        if (n_start <= IR_SEG_LIST) {
            SegGenArgs ga;
            ga.N = N; ga.half_bw = c.half_bw; ga.max_bursts = c.max_bursts; ga.pre_len = c.pre_len; ga.post_len = c.post_len;
            ga.max_burst_len = c.max_burst_len; ga.thr = c.thr; ga.seg = seg; ga.f0 = f0; ga.n_frames = n_frames;
            ga.index0 = index0; ga.sq_start = max(ctl->sq0 - f0, 0);
            ga.xu = xu_c; ga.mag = mag_c; ga.snap = snap; ga.fslot = fslot_g; ga.valid = valid_g;
            ga.keys = gkeys + (size_t)seg * IR_SEG_PCAP; ga.pcap = IR_SEG_PCAP;
            #ifdef IR_SEG_TIMING
            SegLanesWarp lw{ring + IR_SEG_LIST * (sizeof(SegBurst) / 4), tacc, &tlast};
#else
            SegLanesWarp lw{ring + IR_SEG_LIST * (sizeof(SegBurst) / 4)};
#endif
            bail = seg_walk_generic_t(lw, ga, work, n_start, IR_SEG_LIST, gl, IR_SEG_GONE, S.fvs, S.gout);
        }
        __syncwarp();
#ifdef IR_SEG_TIMING
        t_gen = clock64() - t_in - t_fast;
#endif
    }
#ifdef IR_SEG_TIMING
    if (lane == 0) {
        if (generic) {
            for (int k = 0; k < 7; k++) atomicAdd(&g_seg_prof[k], (unsigned long long)tacc[k]);
            atomicAdd(&g_seg_prof[7], (unsigned long long)t_gen); atomicAdd(&g_seg_prof[8], (unsigned long long)t_fast);
            atomicAdd(&g_seg_prof[9], 1ull); atomicAdd(&g_seg_prof[12], (unsigned long long)n_start);
        } else {
            atomicAdd(&g_seg_prof[10], (unsigned long long)t_fast); atomicAdd(&g_seg_prof[11], 1ull);
        }
    }
#endif
    // ---- outputs of this segment: quiet flags, burst list at its last frame, counts
    int changed = 0, st_changed = 0;
    const int nqw = (n_frames + 31) / 32, qw0 = f0 >> 5;
    if (!bail && generic) {
        const int n_end = S.gout.n_end;
        int fc = 0x7fffffff;
        if (lane < nqw) {
            const uint32_t was = qw_g[qw0 + lane], now = S.gout.qbits[lane];
            if (was != now) { changed = 1; qw_g[qw0 + lane] = now; fc = f0 + (lane << 5) + __ffs(was ^ now) - 1; }
        }
        fc = __reduce_min_sync(FULL, fc);
        if (lane == 0 && fc != 0x7fffffff) atomicMin(&ctl->qfc[par], fc);
        const int on = cur[seg + 1].n_act;
        if (on != n_end) st_changed = 1;
        for (int i = lane; i < n_end; i += 32) {
            const SegBurst b = work[i];
            if (i < on) {
                const SegBurst o = list_get(cur, ovf_cur, seg + 1, i);
                if (o.id != b.id || o.start != b.start || o.last0 != b.last0 || o.cb != b.cb ||
                    __float_as_uint(o.rel) != __float_as_uint(b.rel) || __float_as_uint(o.base) != __float_as_uint(b.base) ||
                    o.dl != b.dl || o.lah != b.lah || o.tl != b.tl)
                    st_changed = 1;
            }
            list_put(nxt, ovf_nxt, seg + 1, i, b);
        }
        if (lane == 0) {
            nxt[seg + 1].n_act = n_end; ncreate[seg] = S.gout.n_create; ngone[seg] = S.gout.n_gone;
            if (segbail[seg]) { segbail[seg] = 0; changed = 1; }
            atomicAdd(&ctl->stats[6], 1ull);
        }
    } else if (!bail) {
        flush_quiet();
        int fc = 0x7fffffff;                                  // first frame whose quiet flag changed
        if (lane < nqw) {
            const uint32_t was = qw_g[qw0 + lane];
            if (was != qbits) { changed = 1; qw_g[qw0 + lane] = qbits; fc = f0 + (lane << 5) + __ffs(was ^ qbits) - 1; }
        }
        fc = __reduce_min_sync(FULL, fc);
        if (lane == 0 && fc != 0x7fffffff) atomicMin(&ctl->qfc[par], fc);
        int rank_a = 0;
        for (uint32_t m = have_mask; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const unsigned long long oid = ((unsigned long long)__shfl_sync(FULL, (unsigned)(r_id >> 32), src) << 32) |
                                           (unsigned long long)__shfl_sync(FULL, (unsigned)r_id, src);
            rank_a += (r_have && oid < r_id) ? 1 : 0;
        }
        const SegState &old = cur[seg + 1];
        if (old.n_act != n_act) st_changed = 1;
        if (r_have) {
            SegBurst b;
            b.id = r_id; b.start = r_start; b.last0 = r_last0; b.cb = r_cb; b.rel = r_rel; b.base = r_base;
            b.dl = b_dl + f0; b.lah = b_lah == NONE ? NONE : b_lah + f0; b.tl = b_tl + f0;
            if (rank_a < old.n_act) {
                const SegBurst o = old.b[rank_a];
                if (o.id != b.id || o.start != b.start || o.last0 != b.last0 || o.cb != b.cb ||
                    __float_as_uint(o.rel) != __float_as_uint(b.rel) || __float_as_uint(o.base) != __float_as_uint(b.base) ||
                    o.dl != b.dl || o.lah != b.lah || o.tl != b.tl)
                    st_changed = 1;
            }
            nxt[seg + 1].b[rank_a] = b;
        }
        if (lane == 0) {
            nxt[seg + 1].n_act = n_act; ncreate[seg] = (int)nc_local; ngone[seg] = (int)ng_local;
            if (segbail[seg]) { segbail[seg] = 0; changed = 1; }
        }
    } else {
        // keep the previous round's view of this segment (a bail that only a wrong start produced goes away)
        const int on = cur[seg + 1].n_act;
        for (int i = lane; i < on; i += 32) list_put(nxt, ovf_nxt, seg + 1, i, list_get(cur, ovf_cur, seg + 1, i));
        if (lane == 0) {
            nxt[seg + 1].n_act = on; ncreate[seg] = 0; ngone[seg] = 0;
            if (segbail[seg] != bail) { segbail[seg] = bail; changed = 1; }
        }
    }
    st_changed = __any_sync(FULL, st_changed);
    changed = __any_sync(FULL, changed) | st_changed;
    if (lane == 0) {
        stch_now[seg + 1] = st_changed;
        if (changed) ctl->changed = 1;
        atomicAdd(&ctl->stats[3], st_events);
    }
}

// =========================================================================== commit
// One CTA: ids, gone list, detector header + active list.  Sets ctl->bailed for the fallback.
__global__ void __launch_bounds__(256)
k_seg_commit(DetConfig c, SegCtl *ctl, DetState *gs, const SegState *stA, const SegState *stB, const int *ncreate,
             const int *ngone, const int *segbail, int *cpre, int *gpre, const GoneBurst *__restrict__ glist,
             const SegBurst *ovfA, const SegBurst *ovfB, GoneBurst *__restrict__ gone, uint32_t gone_cap) {
    __shared__ int sh[40];
    const int t = threadIdx.x;
    const int S = ctl->S, F = ctl->F;
    int sb = 0;
    for (int s = t; s < S; s += blockDim.x) sb = max(sb, segbail[s]);
    sb = __syncthreads_or(sb);                                  // (which reason does not matter beyond != 0)
    const bool ok = ctl->finished && ctl->converged && !sb && !ctl->guard_bad && !ctl->hard_bail;
    __syncthreads();
    if (!ok) {
        if (t == 0) {
            int why = 12;
            if (ctl->hard_bail) why = ctl->hard_bail;
            else if (ctl->guard_bad) why = 2;
            else if (sb) { for (int s = 0; s < S; s++) if (segbail[s]) { why = segbail[s]; break; } }
            ctl->bailed = 1;
            ctl->reason = why;
            ctl->stats[1] += 1; ctl->stats[2] += (unsigned long long)ctl->round;
        }
        return;
    }
    const SegState *fin = (ctl->round - 1) & 1 ? stA : stB;    // written by the last executed round
    for (int s = t; s < S; s += blockDim.x) { cpre[s] = ncreate[s]; gpre[s] = ngone[s]; }
    __syncthreads();
    const int n_created = block_excl_scan(cpre, S, sh);
    const int n_gone_new = block_excl_scan(gpre, S, sh);
    const unsigned long long id0 = ctl->next_id0;
    auto real_id = [&](unsigned long long id) -> unsigned long long {
        if (!(id & CODE)) return id;
        const int sg = (int)((id >> 32) & 0x7fffffffu), ord = (int)(id & 0xffffffffu);
        return id0 + 10ull * (unsigned long long)(cpre[sg] + ord);
    };
    const uint32_t g0 = ctl->n_gone0;
    unsigned overflow = 0;
    for (int sg = t >> 5; sg < S; sg += (int)(blockDim.x >> 5)) {
        const int cnt = ngone[sg];
        for (int k = t & 31; k < cnt; k += 32) {
            const uint32_t pos = g0 + (uint32_t)(gpre[sg] + k);
            if (pos < gone_cap) {
                GoneBurst g = glist[(size_t)sg * IR_SEG_GONE + k];
                g.id = real_id(g.id);
                gone[pos] = g;
            } else {
                overflow = 1;
            }
        }
    }
    overflow = __syncthreads_or((int)overflow);
    const SegState &e = fin[S];
    const SegBurst *ovf_fin = (ctl->round - 1) & 1 ? ovfA : ovfB;
    for (int i = t; i < e.n_act; i += blockDim.x) {
        const SegBurst b = i < 32 ? e.b[i] : ovf_fin[(size_t)S * IR_SEG_OVF + (i - 32)];
        ActBurst a;
        a.id = real_id(b.id); a.start = b.start;
        a.last_active = b.lah == NONE ? b.last0 : (unsigned long long)((long long)ctl->index0 + (long long)b.lah * c.N);
        a.center_bin = b.cb; a.peak_rel = b.rel; a.base_at_create = b.base; a.pad = 0;
        gs->act[i] = a;
    }
    if (t == 0) {
        gs->hist_idx = (ctl->hist_idx0 + ctl->nq) % c.hist_size;
        gs->n_act = e.n_act;
        gs->squelch_count = max(ctl->sq0 - F, 0);
        gs->next_id = id0 + 10ull * (unsigned long long)n_created;
        gs->index = ctl->index0 + (unsigned long long)F * (unsigned long long)c.N;
        gs->n_gone = g0 + (uint32_t)n_gone_new;
        if (overflow) gs->overflow = 1;
        ctl->bailed = 0;
        ctl->stats[0] += 1; ctl->stats[2] += (unsigned long long)ctl->round;
        ctl->stats[4] += (unsigned long long)ctl->n_slots; ctl->stats[5] += (unsigned long long)ctl->nq;
    }
}

// Many CTAs: baseline <- final, history rows <- the chunk's last (<= hist_size) quiet frames.  Runs after
// k_seg_commit, only when it kept the chunk (ctl->bailed == 0).
__global__ void __launch_bounds__(256)
k_seg_commit_hist(DetConfig c, const SegCtl *ctl, float *base_g, float *hist, const float *__restrict__ mag,
                  const int *__restrict__ qlist, const float *__restrict__ bfinal) {
    if (ctl->bailed) return;
    const int N = c.N, H = c.hist_size, nq = ctl->nq, h0 = ctl->hist_idx0;
    const int qfirst = max(0, nq - H);
    const int n4 = N >> 2;
    for (int j = blockIdx.x; j < nq - qfirst + 1; j += gridDim.x) {
        if (j == nq - qfirst) {
            for (int i = threadIdx.x; i < n4; i += blockDim.x)
                reinterpret_cast<float4 *>(base_g)[i] = reinterpret_cast<const float4 *>(bfinal)[i];
        } else {
            const int q = qfirst + j;
            const float4 *src = reinterpret_cast<const float4 *>(mag + (size_t)qlist[q] * N);
            float4 *dst = reinterpret_cast<float4 *>(hist + (size_t)((h0 + q) % H) * N);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
        }
    }
}

// =========================================================================== priming
// Frames before the detector is primed: nothing is detected (burst_detect.c:426-428), every frame is quiet
// (:438-454) and the history rows it replaces still hold the zeros of calloc (:283-284).
constexpr int PT = 64, PU = 16;
__global__ void __launch_bounds__(PT)
k_seg_prime(DetConfig c, const DetState *__restrict__ gs, float *base_g, float *hist, const float *__restrict__ mag, int n_frames) {
    const int N = c.N, H = c.hist_size;
    const int bin = blockIdx.x * PT + threadIdx.x;
    const int h0 = gs->hist_idx;
    float base = base_g[bin];
    for (int f0 = 0; f0 < n_frames; f0 += PU) {
        float mv[PU];
#pragma unroll
        for (int u = 0; u < PU; u++) mv[u] = f0 + u < n_frames ? __ldg(mag + (size_t)(f0 + u) * N + bin) : 0.0f;
#pragma unroll
        for (int u = 0; u < PU; u++)
            if (f0 + u < n_frames) {
                const float t = base - 0.0f;
                base = t + mv[u];
                hist[(size_t)((h0 + f0 + u) % H) * N + bin] = mv[u];
            }
    }
    base_g[bin] = base;
}
__global__ void k_seg_prime_hdr(DetConfig c, DetState *gs, int n_frames) {
    int h = gs->hist_idx + n_frames;
    if (h >= c.hist_size) { h -= c.hist_size; gs->primed = 1; }
    gs->hist_idx = h;
    gs->index += (unsigned long long)n_frames * (unsigned long long)c.N;
}

// n_frames <= hist_size - hist_idx frames of an unprimed detector
cudaError_t launch_detect_seg_prime(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                    int n_frames, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    k_seg_prime<<<c.N / PT, PT, 0, st>>>(c, state, base, hist, mag, n_frames);
    k_seg_prime_hdr<<<1, 1, 0, st>>>(c, state, n_frames);
    return cudaGetLastError();
}

// =========================================================================== host side
size_t seg_ctl_bytes() { return sizeof(SegCtl); }

bool seg_scan_supported(const DetConfig &c) {
    const int wpl = c.N / 1024;
    return c.N % 1024 == 0 && (wpl == 4 || wpl == 8 || wpl == 16) && c.hist_size >= 32;
}

size_t seg_walk_smem(const DetConfig &c) {
    const size_t smaxc_ = c.N / 1024 >= 16 ? 512 : 128;
    const size_t row_bytes = (size_t)(c.N / 16) * sizeof(uint32_t);
    const size_t ring = (size_t)RB * SGF * row_bytes, gen = (size_t)IR_SEG_LIST * sizeof(SegBurst) + row_bytes;   // (one or the other)
    return (((sizeof(WkShared) + 15) / 16) * 16 + 3 * smaxc_ * 4 + 127) / 128 * 128 + (ring > gen ? ring : gen);
}

template <int WPL>
static cudaError_t launch_walk_t(const DetConfig &c, const SegBuffers &b, const float *mag, const uint32_t *xu, int S,
                                 cudaStream_t st) {
    const size_t smem = seg_walk_smem(c);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_seg_walk<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    k_seg_walk<WPL><<<S, 32, smem, st>>>(c, b.ctl, mag, xu, b.snap, b.fslot, b.valid, b.stA, b.stB, b.qw, b.ncreate, b.ngone,
                                         b.segbail, b.stch, b.glist, b.ovfA, b.ovfB, b.gkeys, b.seggen);
    return cudaGetLastError();
}

// One chunk of n_frames <= IR_SEG_MAX_FRAMES frames (detector primed).  xu / rowany / ref: bitmaps of the chunk
// from launch_detect_classify.  Enqueues begin + IR_SEG_ROUNDS x (index, base, walk) + commit; the caller
// enqueues the gated fallback (cluster kernel on b.ctl->bailed) behind it.  Returns the number of kernels.
cudaError_t launch_detect_scan_seg(const DetConfig &c, DetState *state, float *base, float *hist, const float *mag,
                                   uint32_t *xu, unsigned char *rowany, const float *ref, int n_frames,
                                   GoneBurst *gone, uint32_t gone_cap, const SegBuffers &b, int *n_launches, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    if (n_frames > IR_SEG_MAX_FRAMES || n_frames > b.frames_cap) return cudaErrorInvalidValue;
    const int S = (n_frames + IR_SEG_LEN - 1) / IR_SEG_LEN;
    {
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_seg_base, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BaseShared));
            if (e != cudaSuccess) return e;
            attr_set[dev] = true;
        }
    }
    k_seg_gather_hist<<<256, 256, 0, st>>>(c, state, hist, b.qmag);
    k_seg_begin<<<1, 256, 0, st>>>(c, state, b.ctl, b.stA, b.stB, b.qw, b.ncreate, b.ngone, b.segbail, b.stch, b.valid, ref, b.glo, b.ghi,
                                   b.ovfA, b.ovfB, b.seggen, n_frames, S);
    // rounds enqueued: a chunk that has no fixed point by then goes to the cluster kernel -- cheap for a short chunk,
    // so short chunks (the 4096-frame pieces of a host run) queue fewer no-op launches
    int rounds = n_frames <= 4096 ? 8 : IR_SEG_ROUNDS;
    if (const char *env = getenv("IR_SEG_ROUNDS")) { const int v = atoi(env); if (v >= 2 && v <= 64) rounds = v; }
    for (int r = 0; r <= rounds; r++) {
        k_seg_index<<<1, 256, 0, st>>>(b.ctl, b.qw, rowany, b.wpre, b.qlist, b.slotv, b.fslot, b.slot_cap, rounds);
        if (r == rounds) break;                                 // the last index launch only tests for the fixed point
        k_seg_gather<<<592, 256, 0, st>>>(c, b.ctl, mag, b.qlist, b.qmag);
        k_seg_base<<<c.N / BB, BT, sizeof(BaseShared), st>>>(c, b.ctl, base, b.qmag, b.glo, b.ghi, b.slotv, b.snap, b.bfinal);
        {
            const int ncol = c.N >> 10;
            int parts = (148 * 16 + ncol - 1) / ncol;
            int fpw = std::max((n_frames + parts - 1) / parts, 8);
            parts = (n_frames + fpw - 1) / fpw;
            k_seg_reclass<<<(parts * ncol + 3) / 4, 128, 0, st>>>(b.ctl, mag, b.glo, b.ghi, c.thr, c.N, fpw, xu, rowany);
        }
        cudaError_t e;
        switch (c.N / 1024) {
        case 4: e = launch_walk_t<4>(c, b, mag, xu, S, st); break;
        case 8: e = launch_walk_t<8>(c, b, mag, xu, S, st); break;
        case 16: e = launch_walk_t<16>(c, b, mag, xu, S, st); break;
        default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
    }
    k_seg_commit<<<1, 256, 0, st>>>(c, b.ctl, state, b.stA, b.stB, b.ncreate, b.ngone, b.segbail, b.cpre, b.gpre, b.glist, b.ovfA, b.ovfB,
                                     gone, gone_cap);
    k_seg_commit_hist<<<148, 256, 0, st>>>(c, b.ctl, base, hist, mag, b.qlist, b.bfinal);
    if (n_launches) *n_launches += 2 + 5 * rounds + 1 + 2;
    return cudaGetLastError();
}

}  // namespace ir
