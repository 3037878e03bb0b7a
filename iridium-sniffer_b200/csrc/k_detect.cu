// k_detect.cu -- burst detector on the device.
//
//  k_detect_fft : window -> N-point FFT -> fftshift -> |X|^2, one CTA per frame (persistent
//                 over frames), replaces process_fft_frame's arithmetic (burst_detect.c:679-687,
//                 simd_avx2.c:145-218) and the OpenCL window/VkFFT/fftshift_magnitude trio
//                 (opencl/burst_fft.c:53-80,333-370).
//  k_detect_scan: the burst state machine over the magnitude frames, strictly in frame order
//                 (burst_detect.c:426-632), one persistent CTA.  Magnitude rows stream into a
//                 shared-memory ring by TMA bulk copies (one 4N-byte cp.async.bulk per frame,
//                 mbarrier completion) several frames ahead of the consumer; the noise baseline
//                 of a thread's bins lives in registers for the whole launch.  Emits burst
//                 descriptors; IQ never moves.
#include "ir_device.cuh"
#include "ir_internal.h"

namespace ir {

#define IR_FFT_FPB 8             // frames per CTA of k_detect_fft

// =========================================================================== FFT + |X|^2
template <int L, int FMT>
__global__ void __launch_bounds__(fft_threads<L>(), (L == 12 || L == 13) ? 2 : 1)   // (two 101 KB CTAs an SM at N = 8192: 128 registers)
k_detect_fft(const void *__restrict__ iq, int64_t first_sample, const float *__restrict__ window,
             const float2 *__restrict__ tw_g, float *__restrict__ mag, int64_t n_frames) {
    constexpr int N = 1 << L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *data = reinterpret_cast<float2 *>(smem_raw);
    float2 *tw = data + fft_data_elems<L>();
    fft_load_twiddles<L>(tw, tw_g);
    __syncthreads();
    // A CTA takes IR_FFT_FPB consecutive frames and retires (the twiddle load above is ~1 % of that): CTAs that
    // lived for the whole launch kept every register of their SMs for 0.5 ms at a time, and the state machine's
    // kernels -- high priority, but nothing preempts a resident CTA -- queued behind them.
    const int64_t fa = (int64_t)blockIdx.x * IR_FFT_FPB, fb = fa + IR_FFT_FPB < n_frames ? fa + IR_FFT_FPB : n_frames;
    for (int64_t f = fa; f < fb; f++) {
        const int64_t s0 = first_sample + f * N;
        float *out = mag + f * N;
        fft_smem<L, false, true>(
            data, tw,
            [&](int p) {
                float2 v = load_sample<FMT>(iq, s0 + p);
                float w = __ldg(window + p);
                return make_float2(v.x * w, v.y * w);          // simd_avx2.c:145-160
            },
            [&](int k, float2 v) { out[k ^ (N >> 1)] = mag2_fma(v); });   // :177-218
    }
}

template <int L>
static cudaError_t launch_fft_L(int fmt, const void *iq, int64_t first_sample, const float *window,
                                const float2 *tw, float *mag, int64_t n_frames, int sm_count,
                                cudaStream_t st) {
    const size_t smem = sizeof(float2) * (fft_data_elems<L>() + fft_tw_elems<L>());
    const int threads = fft_threads<L>();
    (void)sm_count;
    const int64_t grid = (n_frames + IR_FFT_FPB - 1) / IR_FFT_FPB;
    if (grid < 1) return cudaSuccess;
#define IR_LAUNCH_FFT(F)                                                                          \
    do {                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(k_detect_fft<L, F>,                                  \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                           \
        k_detect_fft<L, F><<<(unsigned)grid, threads, smem, st>>>(iq, first_sample, window, tw,   \
                                                                  mag, n_frames);                 \
    } while (0)
    if (fmt == IR_FMT_CF32) IR_LAUNCH_FFT(IR_FMT_CF32);
    else if (fmt == IR_FMT_CI16) IR_LAUNCH_FFT(IR_FMT_CI16);
    else IR_LAUNCH_FFT(IR_FMT_CI8);
#undef IR_LAUNCH_FFT
    return cudaGetLastError();
}

cudaError_t launch_detect_fft(int L, int fmt, const void *iq, int64_t first_sample, const float *window,
                              const float2 *tw, float *mag, int64_t n_frames, int sm_count,
                              cudaStream_t st) {
    switch (L) {
    case 10: return launch_fft_L<10>(fmt, iq, first_sample, window, tw, mag, n_frames, sm_count, st);
    case 11: return launch_fft_L<11>(fmt, iq, first_sample, window, tw, mag, n_frames, sm_count, st);
    case 12: return launch_fft_L<12>(fmt, iq, first_sample, window, tw, mag, n_frames, sm_count, st);
    case 13: return launch_fft_L<13>(fmt, iq, first_sample, window, tw, mag, n_frames, sm_count, st);
    case 14: return launch_fft_L<14>(fmt, iq, first_sample, window, tw, mag, n_frames, sm_count, st);
    default: return cudaErrorInvalidValue;
    }
}

// =========================================================================== state machine
// Thread t owns bins t + 1024*u (u < BPT): one ballot per u yields 32 consecutive bins of the
// "above threshold" bitmap.  The 512-frame history stays in HBM/L2 and is touched only on quiet
// frames (burst_detect.c:438-454).
//
// Threshold test: rel = mag/base > thr is the reference's (IEEE divide, simd_avx2.c:239-257).
// A bin can only pass if mag > base*thr*(1-1e-5), so that cheap product comparison screens
// every bin and the exact divide runs only for the survivors.

struct ScanShared {
    unsigned long long bar[8];    // mbarriers of the magnitude ring
    uint32_t above[512];
    uint32_t free_mask[512];      // 1 = no active burst covers the bin (burst_mask != 0)
    uint32_t valid[512];          // peak search range minus the DC notch (burst_detect.c:537-542)
    ArgMax red[32];
    int n_act;
    int flags[2];
    int squelch_count;
    unsigned long long next_id;
    uint32_t n_gone, n_squelch, overflow;
    ActBurst act[IR_MAX_ACTIVE];
};

__device__ __forceinline__ bool bit_at(const uint32_t *bm, int bin) { return (bm[bin >> 5] >> (bin & 31)) & 1u; }

__device__ __forceinline__ void clear_range(uint32_t *bm, int lo, int hi) {   // inclusive
    for (int w = lo >> 5; w <= (hi >> 5); w++) {
        int a = max(lo, w << 5) & 31, b = min(hi, (w << 5) + 31) & 31;
        uint32_t m = (b == 31 ? 0xffffffffu : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
        atomicAnd(&bm[w], ~m);
    }
}

__device__ __forceinline__ void push_gone(ScanShared &S, GoneBurst *gone, uint32_t gone_cap,
                                          const ActBurst &b, uint64_t index) {
    if (S.n_gone < gone_cap) {
        GoneBurst g;
        g.id = b.id; g.start = b.start; g.stop = index; g.last_active = b.last_active;
        g.center_bin = b.center_bin; g.peak_rel = b.peak_rel; g.base_at_create = b.base_at_create;
        g.pad = 0;
        gone[S.n_gone] = g;
    } else {
        S.overflow = 1;
    }
    S.n_gone++;
}

template <int BPT, int NT, int DEPTH>
__global__ void __launch_bounds__(NT, 1)
k_detect_scan(DetConfig c, DetState *__restrict__ gs, float *__restrict__ base_g,
              float *__restrict__ hist, const float *__restrict__ mag, int64_t n_frames,
              GoneBurst *__restrict__ gone, uint32_t gone_cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                       // [DEPTH][N]
    ScanShared &S = *reinterpret_cast<ScanShared *>(smem_raw + (size_t)DEPTH * c.N * sizeof(float));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = c.N, W = N >> 5;
    const float thr = c.thr;
    const float thr_lo = thr * 0.99999f;
    const uint32_t row_bytes = (uint32_t)N * sizeof(float);
    uint64_t *bars = reinterpret_cast<uint64_t *>(S.bar);

    // ---- load state
    float base[BPT], lim[BPT];
#pragma unroll
    for (int u = 0; u < BPT; u++) {
        base[u] = base_g[u * NT + tid];
        lim[u] = base[u] > 0.0f ? base[u] * thr_lo : __int_as_float(0x7f800000);
    }
    int hist_idx = gs->hist_idx, primed = gs->primed;
    uint64_t index = gs->index;
    if (tid == 0) {
        S.n_act = gs->n_act;
        S.squelch_count = gs->squelch_count;
        S.next_id = gs->next_id;
        S.n_gone = gs->n_gone;
        S.n_squelch = gs->n_squelch;
        S.overflow = gs->overflow;
        S.flags[0] = 0; S.flags[1] = 0;
        for (int d = 0; d < DEPTH; d++) mbar_init(&bars[d], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < IR_MAX_ACTIVE; i += blockDim.x)
        if (i < gs->n_act) S.act[i] = gs->act[i];
    for (int w = tid; w < W; w += blockDim.x) {
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) {
            int bin = (w << 5) + b;
            bool ok = bin >= c.half_bw && bin < N - c.half_bw && !(bin >= N / 2 - 3 && bin <= N / 2 + 3);
            v |= ok ? (1u << b) : 0u;
        }
        S.valid[w] = v;
        S.free_mask[w] = 0xffffffffu;
        S.above[w] = 0;
    }
    __syncthreads();
    if (tid == 0) {                                        // prime the ring
        for (int d = 0; d < DEPTH && d < n_frames; d++) {
            mbar_expect_tx(&bars[d], row_bytes);
            tma_load_1d(ring + (size_t)d * N, mag + (size_t)d * N, row_bytes, &bars[d]);
        }
    }
    if (tid < S.n_act) clear_range(S.free_mask, max(S.act[tid].center_bin - c.half_bw, 0),
                                   min(S.act[tid].center_bin + c.half_bw, N - 1));
    __syncthreads();

    auto baseline_push = [&](const float (&m)[BPT], const float (&old)[BPT]) {   // burst_detect.c:438-454, simd_avx2.c:221-236
        float *h = hist + (size_t)hist_idx * N;
#pragma unroll
        for (int u = 0; u < BPT; u++) {
            const int bin = u * NT + tid;
            float v = base[u] - old[u];
            base[u] = v + m[u];
            lim[u] = base[u] > 0.0f ? base[u] * thr_lo : __int_as_float(0x7f800000);
            h[bin] = m[u];
        }
        if (++hist_idx == c.hist_size) { primed = 1; hist_idx = 0; }
    };
    auto load_old = [&](float (&old)[BPT]) {               // untouched history is zero (calloc / reset)
        const float *h = hist + (size_t)hist_idx * N;
#pragma unroll
        for (int u = 0; u < BPT; u++) old[u] = primed ? h[u * NT + tid] : 0.0f;
    };

    int slot = 0;
    uint32_t slot_phase = 0;
    bool warp_had_bits = true;          // this warp's words of S.above may be non-zero
    for (int64_t f = 0; f < n_frames; f++, index += (uint64_t)N) {
        const float *row = ring + (size_t)slot * N;
        // a frame that starts with no active burst will most likely end in a baseline update:
        // get the history row moving before waiting for the magnitudes
        const bool expect_quiet = S.n_act == 0;
        float old[BPT];
        if (expect_quiet) load_old(old);
        mbar_wait(&bars[slot], slot_phase);
        float m[BPT];
#pragma unroll
        for (int u = 0; u < BPT; u++) m[u] = row[u * NT + tid];
        bool old_valid = expect_quiet;

        if (primed) {
            // Screen all of the thread's bins with one predicate and the warp with one vote;
            // only warps that see a candidate build their 32-bin words bal[u] (bins
            // u*NT + warp*32 .. +31) with the reference's exact divide test.
            uint32_t bal[BPT];
            bool pass_any = false;
#pragma unroll
            for (int u = 0; u < BPT; u++) pass_any = pass_any || (m[u] > lim[u]);
            uint32_t any = 0;
            const bool warp_pass = __any_sync(0xffffffffu, pass_any);
            if (warp_pass) {
#pragma unroll
                for (int u = 0; u < BPT; u++) {
                    const bool ab = (m[u] > lim[u]) && (m[u] / base[u] > thr);   // base > 0 (lim is +inf otherwise)
                    bal[u] = __ballot_sync(0xffffffffu, ab);
                    any |= bal[u];
                }
            } else {
#pragma unroll
                for (int u = 0; u < BPT; u++) bal[u] = 0;
            }
            if (any || warp_had_bits) {
                if (lane < BPT) {
                    uint32_t w = 0;
#pragma unroll
                    for (int u = 0; u < BPT; u++) if (lane == u) w = bal[u];
                    S.above[lane * (NT / 32) + warp] = w;
                }
                warp_had_bits = any != 0;
            }
            const int par = (int)(f & 1);
            if (tid == 0) S.flags[par ^ 1] = 0;
            const int any_above = __syncthreads_or(any != 0);
            // every thread has consumed its part of `row`: refill the slot DEPTH frames ahead
            if (tid == 0 && f + DEPTH < n_frames) {
                mbar_expect_tx(&bars[slot], row_bytes);
                tma_load_1d(ring + (size_t)slot * N, mag + (size_t)(f + DEPTH) * N, row_bytes, &bars[slot]);
            }
            int n_act = S.n_act;
            if (any_above || n_act > 0) {
                int fl = 0;
                // update_bursts (:458-469)
                if (tid < n_act) {
                    ActBurst &b = S.act[tid];
                    const int cb = b.center_bin;
                    bool hit = (cb > 0 && bit_at(S.above, cb - 1)) || bit_at(S.above, cb) ||
                               (cb < N - 1 && bit_at(S.above, cb + 1));
                    if (hit) b.last_active = index;
                    bool too_long = c.max_burst_len > 0 &&
                                    b.last_active - b.start > (uint64_t)c.max_burst_len;
                    bool done = (b.last_active + (uint64_t)c.post_len <= index) || too_long;
                    if (done) fl |= 2;
                    if (too_long) fl |= 4;
                }
                // peaks after masking with the mask left by the previous frame (:522-548),
                // one 32-bin word at a time (warp-uniform)
                uint32_t cw[BPT];
                uint32_t anyc = 0;
                if (any) {
#pragma unroll
                    for (int u = 0; u < BPT; u++) {
                        cw[u] = bal[u] & S.free_mask[u * (NT / 32) + warp] & S.valid[u * (NT / 32) + warp];
                        anyc |= cw[u];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < BPT; u++) cw[u] = 0;
                }
                if (anyc && lane == 0) fl |= 1;
                if (fl) atomicOr(&S.flags[par], fl);
                __syncthreads();
                const int flags = S.flags[par];
                if (flags & 2) {
                    // delete_gone_bursts (:490-518): order-preserving, one thread
                    if (tid == 0) {
                        int k = 0;
                        for (int i = 0; i < n_act; i++) {
                            ActBurst b = S.act[i];
                            bool too_long = c.max_burst_len > 0 &&
                                            b.last_active - b.start > (uint64_t)c.max_burst_len;
                            if ((b.last_active + (uint64_t)c.post_len <= index) || too_long)
                                push_gone(S, gone, gone_cap, b, index);
                            else
                                S.act[k++] = b;
                        }
                        S.n_act = k;
                    }
                    __syncthreads();
                    n_act = S.n_act;
                    if (flags & 4) {                                    // update_filters_post(d, 1)
                        load_old(old);
                        baseline_push(m, old);
                        old_valid = false;
                    }
                    // update_burst_mask (:482-486)
                    for (int w = tid; w < W; w += blockDim.x) S.free_mask[w] = 0xffffffffu;
                    __syncthreads();
                    if (tid < n_act)
                        clear_range(S.free_mask, max(S.act[tid].center_bin - c.half_bw, 0),
                                    min(S.act[tid].center_bin + c.half_bw, N - 1));
                    __syncthreads();
                }
                if (flags & 1) {
                    // create_new_bursts (:556-591): strongest remaining peak first
                    bool cand[BPT];
#pragma unroll
                    for (int u = 0; u < BPT; u++) cand[u] = (cw[u] >> lane) & 1u;
                    for (;;) {
                        ArgMax best{-1.0f, 0x7fffffff};
#pragma unroll
                        for (int u = 0; u < BPT; u++)
                            if (cand[u]) best = argmax_pick(best, ArgMax{m[u] / base[u], u * NT + tid});
                        best = block_argmax(best, S.red);
                        if (best.v < 0.0f) break;
                        const int bin = best.i;
                        const int slot_b = S.n_act;
                        if (slot_b < IR_MAX_ACTIVE) {
                            if (tid == (bin & (NT - 1))) {
                                ActBurst nb;
                                nb.id = S.next_id;
                                nb.start = index - (uint64_t)c.pre_len;
                                nb.last_active = nb.start;
                                nb.center_bin = bin;
                                nb.peak_rel = best.v;
                                float bs = 0.0f;
#pragma unroll
                                for (int u = 0; u < BPT; u++) if (u == bin / NT) bs = base[u];
                                nb.base_at_create = bs;
                                nb.pad = 0;
                                S.act[slot_b] = nb;
                                clear_range(S.free_mask, max(bin - c.half_bw, 0), min(bin + c.half_bw, N - 1));
                            }
                        }
                        const int lo = bin - c.half_bw, hi = bin + c.half_bw;
#pragma unroll
                        for (int u = 0; u < BPT; u++) {
                            const int b = u * NT + tid;
                            if (b >= lo && b <= hi) cand[u] = false;
                        }
                        __syncthreads();
                        if (tid == 0) {
                            if (slot_b < IR_MAX_ACTIVE) S.n_act = slot_b + 1; else S.overflow = 1;
                            S.next_id += 10;
                        }
                        __syncthreads();
                    }
                }
                // squelch (:593-631)
                n_act = S.n_act;
                if (c.max_bursts > 0 && n_act > c.max_bursts) {
                    __syncthreads();
                    if (tid == 0) {
                        for (int i = 0; i < n_act; i++) {
                            const ActBurst &b = S.act[i];
                            if (b.start != index - (uint64_t)c.pre_len) push_gone(S, gone, gone_cap, b, index);
                        }
                        S.n_act = 0;
                        S.n_squelch++;
                        S.squelch_count += 3;
                    }
                    for (int w = tid; w < W; w += blockDim.x) S.free_mask[w] = 0xffffffffu;
                    __syncthreads();
                    const bool reset_noise = S.squelch_count >= 10;
                    __syncthreads();
                    if (reset_noise) {
                        hist_idx = 0; primed = 0;
#pragma unroll
                        for (int u = 0; u < BPT; u++) { base[u] = 0.0f; lim[u] = __int_as_float(0x7f800000); }
                        if (tid == 0) S.squelch_count = 0;
                        old_valid = false;
                    }
                    __syncthreads();
                } else if (flags != 0) {
                    if (tid == 0 && S.squelch_count > 0) S.squelch_count--;
                    __syncthreads();              // S.n_act changed in this frame: publish before the read below
                } else if (tid == 0 && S.squelch_count > 0) {
                    S.squelch_count--;
                }
            } else if (tid == 0 && S.squelch_count > 0) {
                S.squelch_count--;
            }
        } else {
            __syncthreads();                      // all threads done with `row`
            if (tid == 0 && f + DEPTH < n_frames) {
                mbar_expect_tx(&bars[slot], row_bytes);
                tma_load_1d(ring + (size_t)slot * N, mag + (size_t)(f + DEPTH) * N, row_bytes, &bars[slot]);
            }
        }
        // update_filters_post(d, 0) (:438-454)
        if (S.n_act == 0) {
            if (!old_valid) load_old(old);
            baseline_push(m, old);
        }
        if (++slot == DEPTH) { slot = 0; slot_phase ^= 1u; }
    }

    // ---- store state
    __syncthreads();
#pragma unroll
    for (int u = 0; u < BPT; u++) base_g[u * NT + tid] = base[u];
    for (int i = tid; i < S.n_act; i += blockDim.x) gs->act[i] = S.act[i];
    if (tid == 0) {
        gs->hist_idx = hist_idx; gs->primed = primed; gs->n_act = S.n_act;
        gs->squelch_count = S.squelch_count; gs->next_id = S.next_id; gs->index = index;
        gs->n_gone = S.n_gone; gs->n_squelch = S.n_squelch; gs->overflow = S.overflow;
    }
}

template <int BPT, int NT, int DEPTH>
static cudaError_t launch_scan_t(const DetConfig &c, DetState *state, float *base, float *hist,
                                 const float *mag, int64_t n_frames, GoneBurst *gone,
                                 uint32_t gone_cap, cudaStream_t st) {
    const size_t smem = (size_t)DEPTH * c.N * sizeof(float) + sizeof(ScanShared);
    cudaError_t e = cudaFuncSetAttribute(k_detect_scan<BPT, NT, DEPTH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_detect_scan<BPT, NT, DEPTH><<<1, NT, smem, st>>>(c, state, base, hist, mag, n_frames,
                                                       gone, gone_cap);
    return cudaGetLastError();
}

// Threads x bins-per-thread = N.  512 threads keep every per-bin value in registers (no
// spills at <=128 registers); the 16384-point detector needs 1024 threads.
cudaError_t launch_detect_scan(const DetConfig &c, DetState *state, float *base, float *hist,
                               const float *mag, int64_t n_frames, GoneBurst *gone,
                               uint32_t gone_cap, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    switch (c.N) {
    case 1024: return launch_scan_t<2, 512, 8>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 2048: return launch_scan_t<4, 512, 8>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 4096: return launch_scan_t<8, 512, 8>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 8192: return launch_scan_t<16, 512, 5>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    case 16384: return launch_scan_t<16, 1024, 2>(c, state, base, hist, mag, n_frames, gone, gone_cap, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace ir
