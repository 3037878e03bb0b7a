// k_downmix.cu -- per-burst downmix on the device (burst_downmix.c:643-797).
//
//  k_rot_tables : exact NCO phase checkpoints.  The reference rotates with the float
//                 recurrence phase *= incr (rotator.h:36-46); its rounding drift is visible
//                 in the RAW `level` field, so the recurrence is
//                 reproduced, not replaced: one thread runs it once per detector bin and
//                 stores every 16th phase; FIR CTAs restart from the checkpoints.
//  k_fir        : coarse rotate + 801-tap low-pass + decimate (steps 1-2).  Taps sit in
//                 constant memory and the tap loop is fully unrolled, so every FFMA takes its
//                 coefficient as a constant-bank operand; each loaded sample feeds 8 outputs
//                 from registers.  Summation order is the AVX2 kernel's (simd_avx2.c:62-110):
//                 four interleaved chains k = j mod 4, combined (c0+c2)+(c1+c3), then tail taps.
//  k_chain      : everything at 250 kHz (steps 2b-9), one CTA per burst.
#include <stdlib.h>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace ir {

__constant__ float c_in_taps[808];
__constant__ float c_noise[32];
__constant__ float c_box[24];
__constant__ float c_rrc[56];
__constant__ float c_cfo_win[IR_CFO_N];
__constant__ int c_ntaps[4];      // input, noise, box, rrc

cudaError_t upload_input_taps(const float *taps, int ntaps) {
    if (ntaps != IR_INPUT_NTAPS) return cudaErrorInvalidValue;
    float tmp[808] = {0};
    for (int i = 0; i < ntaps; i++) tmp[i] = taps[i];
    return cudaMemcpyToSymbol(c_in_taps, tmp, sizeof(tmp));
}

cudaError_t upload_chain_tables(const HostTables &t) {
    if (t.h_noise.size() > 32 || t.h_box.size() > 24 || t.h_rrc.size() > 56 ||
        t.cfo_window.size() != IR_CFO_N)
        return cudaErrorInvalidValue;
    float a[32] = {0}, b[24] = {0}, r[56] = {0};
    for (size_t i = 0; i < t.h_noise.size(); i++) a[i] = t.h_noise[i];
    for (size_t i = 0; i < t.h_box.size(); i++) b[i] = t.h_box[i];
    for (size_t i = 0; i < t.h_rrc.size(); i++) r[i] = t.h_rrc[i];
    int nt[4] = {(int)t.h_input.size(), (int)t.h_noise.size(), (int)t.h_box.size(), (int)t.h_rrc.size()};
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_noise, a, sizeof(a))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_box, b, sizeof(b))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_rrc, r, sizeof(r))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_cfo_win, t.cfo_window.data(), sizeof(float) * IR_CFO_N)) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_ntaps, nt, sizeof(nt));
}

// ============================================================== NCO checkpoints
__global__ void k_rot_tables(const float2 *__restrict__ incr, float2 *const *__restrict__ tables,
                             const int *__restrict__ lens, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float2 w = incr[t];
    float2 ph = make_float2(1.0f, 0.0f);
    float2 *T = tables[t];
    const int len = lens[t];
    for (int i = 0; i < len; i++) {
        if ((i & (IR_ROT_G - 1)) == 0) T[i / IR_ROT_G] = ph;
        ph = cmul(ph, w);                                  // rotator.h:41
    }
}

cudaError_t launch_rot_tables(const float2 *incr, float2 *const *tables, const int *lens, int n,
                              cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_rot_tables<<<(n + 31) / 32, 32, 0, st>>>(incr, tables, lens, n);
    return cudaGetLastError();
}

// ============================================================== rotate + FIR + decimate
// decimated outputs per tile: 256 (32 lanes x 8) at DEC = 40; 240 (30 lanes) at DEC = 48, where two 256-output sample
// buffers would not fit an SM's shared memory (240: 216.5 KB of the 227)
template <int DEC> __host__ __device__ constexpr int fir_tile() { return IR_FIR_TILE_OF(DEC); }
template <int DEC> __host__ __device__ constexpr int fir_in_max() { return (fir_tile<DEC>() - 1) * DEC + IR_INPUT_NTAPS; }
// Shared-memory index of burst sample e.  Two access patterns must both be conflict-free for
// 8-byte accesses: the rotate phase (lane stride 16 samples) and the FIR phase (lane stride
// R*DEC samples).  e + e/16 + e/(R*DEC) gives lane strides of 17 and R*DEC*17/16+1 (341 for
// DEC=40, 409 for DEC=48), both odd.
template <int DEC> __host__ __device__ constexpr int fir_pi_c(int e) { return e + (e >> 4) + e / (IR_FIR_R * DEC); }
template <int DEC> __device__ __forceinline__ int fir_pi(int e) { return e + (e >> 4) + e / (IR_FIR_R * DEC); }
// A chain walks whole blocks of DEC/4 taps: at DEC = 48 its last block holds taps 192..203 of which 200..203 are
// zero padding, and the samples those meet lie up to 15 past the tile's last input; the unpeeled middle blocks
// also run a handful of zero-tap FMAs on samples past an output's last input, which for the last outputs of a burst's
// last (partial) tile lie past what was staged.  0 * x must be 0: the staging writes zeros from the tile's last input
// to the end of the buffer (stale shared memory can hold anything, NaN patterns included; IR_FIR_POISON=1 fills the
// shared memory with NaNs at kernel start so that tests/test_zzz_gpu_fir_poison.py sees any such read).
template <int DEC> __host__ __device__ constexpr int fir_over() {
    return 4 * ((IR_INPUT_NTAPS / 4 + DEC / 4 - 1) / (DEC / 4)) * (DEC / 4) - (IR_INPUT_NTAPS / 4) * 4;
}
template <int DEC> __host__ __device__ constexpr int fir_stage_max() { return fir_in_max<DEC>() + fir_over<DEC>(); }
template <int DEC> __host__ __device__ constexpr int fir_pitch_elems() { return fir_pi_c<DEC>(fir_stage_max<DEC>()) + 4; }

// Zero-padded tap table in shared memory: hp[IR_FIR_HPAD + k] = taps[k] for 0 <= k < 800, else 0.
#define IR_FIR_HPAD 0
template <int DEC> __host__ __device__ constexpr int fir_nit() { return (IR_INPUT_NTAPS / 4 + DEC / 4 - 1) / (DEC / 4) + IR_FIR_R - 1; }
template <int DEC> __host__ __device__ constexpr int fir_hp_elems() { return IR_FIR_HPAD + DEC * ((fir_nit<DEC>() + 7) / 8 * 8) + 8; }

// One lane accumulates chain J (taps J, J+4, J+8, ... in ascending order, one FMA each --
// simd_avx2.c:76-86) of 8 consecutive outputs.  Output i lags output 0 by DEC/4 steps, so the
// tap a sample meets in chain i is the one chain i-1 used DEC/4 steps earlier: taps rotate
// through 8 register sets of DEC/4 values, each loaded sample feeds 16 FMAs, and the whole
// loop is ~1.5 K instructions shared by all warps (J is a run-time offset).  Before a chain
// starts and after it ends it sees zero taps: fma(0, x, acc) leaves acc unchanged.
template <int DEC>
__device__ __forceinline__ void fir_chains(const float2 *__restrict__ sp, const float *__restrict__ hp,
                                           float2 (&acc)[IR_FIR_R]) {
    constexpr int D4 = DEC / 4;
    constexpr int NIT = fir_nit<DEC>();
    constexpr int ROW = IR_FIR_R * DEC + IR_FIR_R * DEC / 16 + 1;      // fir_pi advance per 8 iterations
    float G[8][D4];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int r = 0; r < D4; r++) G[a][r] = 0.0f;
    for (int T = 0; T < (NIT + 7) / 8; T++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (8 * T + u < NIT) {                                     // warp-uniform
#pragma unroll
                for (int r = 0; r < D4; r++) G[u][r] = hp[DEC * u + 4 * r];
#pragma unroll
                for (int r = 0; r < D4; r++) {
                    const int ql = D4 * u + r;
                    const float2 x = sp[4 * ql + (ql >> 2)];
#pragma unroll
                    for (int i = 0; i < IR_FIR_R; i++) {
                        const float h = G[(u - i + 8) & 7][r];
                        acc[i].x = fmaf(h, x.x, acc[i].x);
                        acc[i].y = fmaf(h, x.y, acc[i].y);
                    }
                }
            }
        }
        sp += ROW;
        hp += 8 * DEC;
    }
}

// The same chains with packed FMAs (sm_100 fma.rn.f32x2, SASS FFMA2): one instruction updates the real and the
// imaginary accumulator of an output with the shared tap -- each lane of the pair is an IEEE fma with a single
// rounding, so the sums are the ones above, bit for bit.  (FFMA2 takes the tap as a scalar operand broadcast to
// both halves -- `FFMA2 Rd, Rh.F32, Rx.F32x2.HI_LO, Racc.F32x2.HI_LO` -- so the tap registers stay 80.)
// The first and the last block of iterations are separate bodies that leave out the FMAs on zero taps (21 % of them).
// One block of 8 iterations of the packed chains.  MODE 1: every (iteration, output) pair.  MODE 0 (the first block)
// and MODE 2 (the last): the pairs whose tap block lies inside the filter -- output i lags output 0 by i iterations, so
// in the first block it has not started before iteration i and in the last it has ended after iteration i + NB - 1;
// outside that range the rotating registers hold zero taps and fma(0, x, acc) == acc: the FMA is dropped, not the sum.
template <int DEC, int MODE, int T>
__device__ __forceinline__ void fir_block_f2(const float2 *__restrict__ sp, const float *__restrict__ hp,
                                             float (&G)[8][DEC / 4], float2 (&acc)[IR_FIR_R]) {
    constexpr int D4 = DEC / 4;
    constexpr int NIT = fir_nit<DEC>();
    constexpr int NB = (IR_INPUT_NTAPS / 4 + D4 - 1) / D4;            // tap blocks of a chain
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int it = 8 * T + u;
        if (MODE != 1 && it >= NIT) break;
        if (MODE == 1 && it >= NIT) break;
#pragma unroll
        for (int r = 0; r < D4; r++) G[u][r] = hp[DEC * u + 4 * r];
#pragma unroll
        for (int r = 0; r < D4; r++) {
            const int ql = D4 * u + r;
            const float2 x = sp[4 * ql + (ql >> 2)];
#pragma unroll
            for (int i = 0; i < IR_FIR_R; i++) {
                const bool live = MODE == 1 || (it - i >= 0 && it - i < NB);
                if (live) acc[i] = __ffma2_rn(make_float2(G[(u - i + 8) & 7][r], G[(u - i + 8) & 7][r]), x, acc[i]);
            }
        }
    }
}
template <int DEC>
__device__ __forceinline__ void fir_chains_f2(const float2 *__restrict__ sp, const float *__restrict__ hp,
                                              float2 (&acc)[IR_FIR_R]) {
    constexpr int D4 = DEC / 4;
    constexpr int NIT = fir_nit<DEC>();
    constexpr int NT = (NIT + 7) / 8;
    constexpr int ROW = IR_FIR_R * DEC + IR_FIR_R * DEC / 16 + 1;
    static_assert(NT == 3 || NT == 4, "block structure below");
    float G[8][D4];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int r = 0; r < D4; r++) G[a][r] = 0.0f;
    fir_block_f2<DEC, 0, 0>(sp, hp, G, acc);
    sp += ROW; hp += 8 * DEC;
#pragma unroll 1
    for (int T = 1; T < NT - 1; T++) {                                // (iterations 8..NIT-9+: every pair inside the filter,
        fir_block_f2<DEC, 1, 1>(sp, hp, G, acc);                      //  or a handful of zero-tap FMAs not worth a third body)
        sp += ROW; hp += 8 * DEC;
    }
    fir_block_f2<DEC, 2, NT - 1>(sp, hp, G, acc);
}

template <int FMT, int DEC>
__global__ void __launch_bounds__(128)
k_fir(const void *__restrict__ iq, int64_t n_total, uint64_t ring, const BurstParam *__restrict__ bp,
      const int *__restrict__ tile_start, int n_bursts, float2 *__restrict__ dec_out) {
    static_assert(DEC % 4 == 0, "register-tiled FIR needs dec % 4 == 0");
    static_assert((fir_tile<DEC>() * DEC) % IR_ROT_G == 0, "tiles must start on a phase checkpoint");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s = reinterpret_cast<float2 *>(smem_raw);
    float2 *part = s + fir_pitch_elems<DEC>();            // [4][fir_tile<DEC>()]
    float *hp = reinterpret_cast<float *>(part + 4 * fir_tile<DEC>());   // zero-padded body taps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < fir_hp_elems<DEC>(); k += blockDim.x) {
        const int kk = k - IR_FIR_HPAD;
        hp[k] = (kk >= 0 && kk < (IR_INPUT_NTAPS / 4) * 4) ? c_in_taps[kk] : 0.0f;
    }

    // which burst owns this tile
    int lo = 0, hi = n_bursts;
    const int tile = blockIdx.x;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_start[mid] <= tile) lo = mid; else hi = mid;
    }
    const BurstParam P = bp[lo];
    const int o0 = (tile - P.tile0) * fir_tile<DEC>();
    const int n_out = min(fir_tile<DEC>(), P.dec_len - o0);
    const int e0 = o0 * DEC;
    const int n_in = (n_out - 1) * DEC + IR_INPUT_NTAPS;

    // A: stage the raw samples.  Positions the detector had not yet received when it emitted
    // the burst read the ring slot's previous content: one lap earlier, or zero.
    // cf32 rows go global -> shared asynchronously (LDGSTS, 8 B each: burst starts are only
    // guaranteed 8-byte aligned inside the pitched layout), so all ~86 copies of a thread are in
    // flight at once instead of one L2 round trip per element.
#pragma unroll 4
    for (int e = tid; e < fir_stage_max<DEC>(); e += blockDim.x) {
        float2 *dst = &s[fir_pi<DEC>(e)];
        bool filled = false;
        if (e < n_in && e0 + e < P.n) {
            int64_t q = P.start + e0 + e;
            if (q >= P.emit_count) q -= (int64_t)ring;
            if (q >= 0 && q < n_total) {
                if (FMT == IR_FMT_CF32) cp_async_8(dst, reinterpret_cast<const float2 *>(iq) + q);
                else *dst = load_sample<FMT>(iq, q);
                filled = true;
            }
        }
        if (!filled) *dst = make_float2(0.0f, 0.0f);
    }
    if (FMT == IR_FMT_CF32) cp_async_wait_all();
    __syncthreads();
    // B: coarse frequency shift in place, 16 samples per checkpoint (rotator.h:36-42)
    {
        const int nseg = (n_in + IR_ROT_G - 1) / IR_ROT_G;
        for (int seg = tid; seg < nseg; seg += blockDim.x) {
            float2 ph = P.rot_table[(e0 >> 4) + seg];
#pragma unroll
            for (int i = 0; i < IR_ROT_G; i++) {
                const int e = seg * IR_ROT_G + i;
                if (e < n_in) {
                    float2 *p = &s[fir_pi<DEC>(e)];
                    *p = cmul(*p, ph);
                }
                ph = cmul(ph, P.incr_coarse);
            }
        }
    }
    __syncthreads();
    // C: four chains per output, 8 outputs per lane
    {
        float2 acc[IR_FIR_R];
#pragma unroll
        for (int i = 0; i < IR_FIR_R; i++) acc[i] = make_float2(0.0f, 0.0f);
        constexpr int ROW = IR_FIR_R * DEC + IR_FIR_R * DEC / 16 + 1;
        if (IR_FIR_R * lane < fir_tile<DEC>()) {
            fir_chains<DEC>(s + ROW * lane + warp, hp + IR_FIR_HPAD + warp, acc);     // chain J = warp
#pragma unroll
            for (int i = 0; i < IR_FIR_R; i++) part[warp * fir_tile<DEC>() + IR_FIR_R * lane + i] = acc[i];
        }
    }
    __syncthreads();
    // D: (c0+c2)+(c1+c3), leftover taps, store (simd_avx2.c:88-109)
    for (int o = tid; o < n_out; o += blockDim.x) {
        const float2 c0 = part[o], c1 = part[fir_tile<DEC>() + o], c2 = part[2 * fir_tile<DEC>() + o],
                     c3 = part[3 * fir_tile<DEC>() + o];
        float ar = (c0.x + c2.x) + (c1.x + c3.x);
        float ai = (c0.y + c2.y) + (c1.y + c3.y);
#pragma unroll
        for (int k = (IR_INPUT_NTAPS / 4) * 4; k < IR_INPUT_NTAPS; k++) {
            const float2 x = s[fir_pi<DEC>(o * DEC + k)];
            ar = fmaf(c_in_taps[k], x.x, ar);
            ai = fmaf(c_in_taps[k], x.y, ai);
        }
        dec_out[P.dec_off + o0 + o] = make_float2(ar, ai);
    }
}

// ---- warp-specialised version: the default.  One CTA of 8 warps works through a strip of consecutive tiles
// with two sample buffers.  Warps 4-7 ("producers") stage tile t+1 -- asynchronous copies, then the exact NCO
// in place -- while warps 0-3 ("consumers", one per chain) run the FIR of tile t, so the FMA pipe never waits
// for a copy or a rotate phase; hand-over by mbarriers (full / empty per buffer).  The arithmetic is k_fir's,
// instruction for instruction (same fir_chains, same combine, same tail taps): results are bit-identical.
#define IR_FIR_STRIP 16
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ int g_fir_poison = 0;       // IR_FIR_POISON=1 (tests): start from shared memory full of NaNs
constexpr int FWS_C = 128;             // consumer threads (one warp per chain)
// a sample buffer: the padded index of the last sample + 4, and -- the rotate phase turns whole 16-sample segments, so
// the last segment's unused tail is written too -- the end of the last segment (at DEC = 48 that tail is 15 samples
// and reached two elements past the old pitch, into the other buffer's first samples)
template <int DEC> __host__ __device__ constexpr int fws_seg_end() {
    constexpr int last = (fir_in_max<DEC>() + IR_ROT_G - 1) / IR_ROT_G - 1;
    return last * (IR_ROT_G + 1) + last / (IR_FIR_R * DEC / IR_ROT_G) + IR_ROT_G;
}
template <int DEC> __host__ __device__ constexpr int fws_pitch() {
    constexpr int a = fir_pitch_elems<DEC>() + 8, b = fws_seg_end<DEC>();
    return ((a > b ? a : b) + 1) & ~1;
}
static_assert(fws_pitch<40>() >= fws_seg_end<40>() && fws_pitch<48>() >= fws_seg_end<48>(), "rotate phase stays inside its buffer");
static_assert(fir_over<40>() == 0 && fir_over<48>() == 16, "zero-tap over-read of the last tap block");
// FWS_P producer threads; F2: packed FMAs in the chains
template <int FMT, int DEC, int FWS_P, bool F2>
__global__ void __launch_bounds__(FWS_C + FWS_P, 1)
k_fir_ws(const void *__restrict__ iq, int64_t n_total, uint64_t ring, const BurstParam *__restrict__ bp,
         const int *__restrict__ tile_burst, int n_tiles, float2 *__restrict__ dec_out) {
    static_assert(DEC % 4 == 0, "register-tiled FIR needs dec % 4 == 0");
    static_assert((fir_tile<DEC>() * DEC) % IR_ROT_G == 0, "tiles must start on a phase checkpoint");
    static_assert((IR_FIR_R * DEC) % IR_ROT_G == 0 && IR_ROT_G == 16, "segment layout below");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PITCH = fws_pitch<DEC>();
    constexpr int NSEG_MAX = (fir_in_max<DEC>() + IR_ROT_G - 1) / IR_ROT_G;
    constexpr int SPT = (NSEG_MAX + FWS_P - 1) / FWS_P;         // 16-sample segments per producer thread
    constexpr int CPT = (fir_in_max<DEC>() + FWS_P - 1) / FWS_P; // copies per producer thread
    constexpr int SEG_PER_ROW = IR_FIR_R * DEC / IR_ROT_G;       // segments between two "row" pads of fir_pi
    float2 *sb0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *part0 = sb0 + 2 * PITCH;                           // [4][fir_tile<DEC>()]
    float *hp = reinterpret_cast<float *>(part0 + 4 * fir_tile<DEC>());
    uint64_t *bars = reinterpret_cast<uint64_t *>(hp + ((fir_hp_elems<DEC>() + 1) & ~1));   // full[2], empty[2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (g_fir_poison) {
        for (int k = tid; k < 2 * (2 * PITCH + 4 * fir_tile<DEC>()); k += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[k] = 0xffffffffu;
        __syncthreads();
    }
    for (int k = tid; k < fir_hp_elems<DEC>(); k += blockDim.x) {
        const int kk = k - IR_FIR_HPAD;
        hp[k] = (kk >= 0 && kk < (IR_INPUT_NTAPS / 4) * 4) ? c_in_taps[kk] : 0.0f;
    }
    if (tid == 0) {
        mbar_init(&bars[0], FWS_P); mbar_init(&bars[1], FWS_P); mbar_init(&bars[2], FWS_C); mbar_init(&bars[3], FWS_C);
        fence_mbar_init();
    }
    __syncthreads();
    const int t0 = blockIdx.x * IR_FIR_STRIP, t1 = min(n_tiles, t0 + IR_FIR_STRIP);
    if (tid >= FWS_C) {
        // ------------------------------------------------------------------ producers
        const int ptid = tid - FWS_C;
        for (int t = t0; t < t1; t++) {
            const int b = (t - t0) & 1, use = (t - t0) >> 1;
            float2 *s = sb0 + b * PITCH;
            const BurstParam P = bp[tile_burst[t]];
            const int o0 = (t - P.tile0) * fir_tile<DEC>();
            const int n_out = min(fir_tile<DEC>(), P.dec_len - o0);
            const int e0 = o0 * DEC;
            const int n_in = (n_out - 1) * DEC + IR_INPUT_NTAPS;
            // the NCO checkpoints of this thread's segments: in flight while the samples are staged
            const int nseg = (n_in + IR_ROT_G - 1) / IR_ROT_G;
            const float2 *tab = P.rot_table + (e0 >> 4);
            float2 ph[SPT];
#pragma unroll
            for (int j = 0; j < SPT; j++) {
                const int seg = ptid + FWS_P * j;
                ph[j] = seg < nseg ? __ldg(tab + seg) : make_float2(0.0f, 0.0f);
            }
            if (use > 0) mbar_wait(&bars[2 + b], (uint32_t)((use - 1) & 1));   // the consumers are done with this buffer
            // A: stage the raw samples.  Positions the detector had not yet received when it emitted the burst
            // read the ring slot's previous content: one lap earlier, or zero.
            const int64_t q0 = P.start + e0;
            const bool plain = q0 >= 0 && e0 + n_in <= P.n &&
                               q0 + n_in <= (P.emit_count < n_total ? P.emit_count : n_total);
            if (plain && FMT == IR_FMT_CF32) {
                const float2 *src = reinterpret_cast<const float2 *>(iq) + q0;
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const int e = ptid + FWS_P * k;
                    if (e < n_in) cp_async_8(&s[fir_pi<DEC>(e)], src + e);
                }
                for (int e = n_in + ptid; e < fir_stage_max<DEC>(); e += FWS_P) s[fir_pi<DEC>(e)] = make_float2(0.0f, 0.0f);
            } else if (plain) {
                // integer samples: every load of the thread in flight, then the conversions (simd_avx2.c:264-294)
                uint32_t raw[CPT];
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const int e = ptid + FWS_P * k;
                    raw[k] = e < n_in ? load_raw<FMT>(iq, q0 + e) : 0u;
                }
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    const int e = ptid + FWS_P * k;
                    if (e < n_in) s[fir_pi<DEC>(e)] = conv_raw<FMT>(raw[k]);
                }
                for (int e = n_in + ptid; e < fir_stage_max<DEC>(); e += FWS_P) s[fir_pi<DEC>(e)] = make_float2(0.0f, 0.0f);
            } else {
#pragma unroll 4
                for (int e = ptid; e < fir_stage_max<DEC>(); e += FWS_P) {
                    float2 *dst = &s[fir_pi<DEC>(e)];
                    bool filled = false;
                    if (e < n_in && e0 + e < P.n) {
                        int64_t q = q0 + e;
                        if (q >= P.emit_count) q -= (int64_t)ring;
                        if (q >= 0 && q < n_total) {
                            if (FMT == IR_FMT_CF32) cp_async_8(dst, reinterpret_cast<const float2 *>(iq) + q);
                            else *dst = load_sample<FMT>(iq, q);
                            filled = true;
                        }
                    }
                    if (!filled) *dst = make_float2(0.0f, 0.0f);
                }
            }
            if (FMT == IR_FMT_CF32) cp_async_wait_all();
            asm volatile("bar.sync 2, %0;" ::"n"(FWS_P) : "memory");
            // B: coarse frequency shift in place, 16 samples per checkpoint (rotator.h:36-42).  A thread's
            // segments are independent recurrences: they advance together.  Inside a segment the padded index is
            // the segment's base + i, and whatever lies past n_in in the last segment is nobody's input.
            {
                float2 *sp[SPT];
#pragma unroll
                for (int j = 0; j < SPT; j++) {
                    const int seg = min(ptid + FWS_P * j, NSEG_MAX - 1);
                    sp[j] = s + seg * (IR_ROT_G + 1) + seg / SEG_PER_ROW;
                }
                const float2 w = P.incr_coarse;
#pragma unroll
                for (int i = 0; i < IR_ROT_G; i++) {
#pragma unroll
                    for (int j = 0; j < SPT; j++) {
                        if (ptid + FWS_P * j < nseg) sp[j][i] = cmul(sp[j][i], ph[j]);
                        ph[j] = cmul(ph[j], w);
                    }
                }
            }
            mbar_arrive(&bars[b]);                              // full (release: the rotated samples are visible)
        }
    } else {
        // ------------------------------------------------------------------ consumers: chain J = warp
        for (int t = t0; t < t1; t++) {
            const int b = (t - t0) & 1, use = (t - t0) >> 1;
            float2 *s = sb0 + b * PITCH;
            float2 *part = part0;
            const BurstParam P = bp[tile_burst[t]];
            const int o0 = (t - P.tile0) * fir_tile<DEC>();
            const int n_out = min(fir_tile<DEC>(), P.dec_len - o0);
            mbar_wait(&bars[b], (uint32_t)(use & 1));
            {
                float2 acc[IR_FIR_R];
#pragma unroll
                for (int i = 0; i < IR_FIR_R; i++) acc[i] = make_float2(0.0f, 0.0f);
                constexpr int ROW = IR_FIR_R * DEC + IR_FIR_R * DEC / 16 + 1;
                if (IR_FIR_R * lane < fir_tile<DEC>()) {
                    if (F2) fir_chains_f2<DEC>(s + ROW * lane + warp, hp + IR_FIR_HPAD + warp, acc);
                    else fir_chains<DEC>(s + ROW * lane + warp, hp + IR_FIR_HPAD + warp, acc);
                }
                // (one partial-sum buffer: the previous tile's combine must be over before it is rewritten -- 8 KB
                // less shared memory is what lets a state-machine walker sit beside this CTA)
                if (t > t0) asm volatile("bar.sync 1, %0;" ::"n"(FWS_C) : "memory");
                if (IR_FIR_R * lane < fir_tile<DEC>()) {
#pragma unroll
                    for (int i = 0; i < IR_FIR_R; i++) part[warp * fir_tile<DEC>() + IR_FIR_R * lane + i] = acc[i];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(FWS_C) : "memory");
            // D: (c0+c2)+(c1+c3), leftover taps, store (simd_avx2.c:88-109)
            for (int o = tid; o < n_out; o += FWS_C) {
                const float2 c0 = part[o], c1 = part[fir_tile<DEC>() + o], c2 = part[2 * fir_tile<DEC>() + o],
                             c3 = part[3 * fir_tile<DEC>() + o];
                float ar = (c0.x + c2.x) + (c1.x + c3.x);
                float ai = (c0.y + c2.y) + (c1.y + c3.y);
#pragma unroll
                for (int k = (IR_INPUT_NTAPS / 4) * 4; k < IR_INPUT_NTAPS; k++) {
                    const float2 x = s[fir_pi<DEC>(o * DEC + k)];
                    ar = fmaf(c_in_taps[k], x.x, ar);
                    ai = fmaf(c_in_taps[k], x.y, ai);
                }
                dec_out[P.dec_off + o0 + o] = make_float2(ar, ai);
            }
            mbar_arrive(&bars[2 + b]);                          // empty
        }
    }
}

template <int FMT, int DEC>
static cudaError_t launch_fir_t(const void *iq, int64_t n_total, uint64_t ring, const BurstParam *bp,
                                const int *tile_start, const int *tile_burst, int n_bursts, int n_tiles, float2 *dec_out,
                                cudaStream_t st) {
    static const bool legacy = getenv("IR_FIR_LEGACY") != nullptr;
    constexpr int PITCH = fws_pitch<DEC>();
    const size_t smem_ws = sizeof(float2) * (2 * PITCH + 4 * fir_tile<DEC>()) + sizeof(float) * ((fir_hp_elems<DEC>() + 1) & ~1) + 4 * sizeof(uint64_t);
    // (the one-tile kernel: for comparison, and should a tile size ever not fit two sample buffers)
    if (legacy || tile_burst == nullptr || smem_ws > (size_t)227 * 1024) {
        const size_t smem = sizeof(float2) * (fir_pitch_elems<DEC>() + 4 * fir_tile<DEC>()) + sizeof(float) * fir_hp_elems<DEC>();
        cudaError_t e = cudaFuncSetAttribute(k_fir<FMT, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_fir<FMT, DEC><<<n_tiles, 128, smem, st>>>(iq, n_total, ring, bp, tile_start, n_bursts, dec_out);
        return cudaGetLastError();
    }
    const size_t smem = smem_ws;
    static const bool poison = [] {
        const bool on = getenv("IR_FIR_POISON") != nullptr;
        const int v = on ? 1 : 0;
        if (on) cudaMemcpyToSymbol(g_fir_poison, &v, sizeof(int));
        return on;
    }();
    (void)poison;
    static const bool scalar = getenv("IR_FIR_SCALAR") != nullptr;    // the unpacked chains with eight producer warps
    const unsigned grid = (unsigned)((n_tiles + IR_FIR_STRIP - 1) / IR_FIR_STRIP);
    if (scalar) {
        cudaError_t e = cudaFuncSetAttribute(k_fir_ws<FMT, DEC, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_fir_ws<FMT, DEC, 256, false><<<grid, FWS_C + 256, smem, st>>>(iq, n_total, ring, bp, tile_burst, n_tiles, dec_out);
    } else if (getenv("IR_FIR_P128") != nullptr) {
        cudaError_t e = cudaFuncSetAttribute(k_fir_ws<FMT, DEC, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_fir_ws<FMT, DEC, 128, true><<<grid, FWS_C + 128, smem, st>>>(iq, n_total, ring, bp, tile_burst, n_tiles, dec_out);
    } else {
        cudaError_t e = cudaFuncSetAttribute(k_fir_ws<FMT, DEC, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_fir_ws<FMT, DEC, 256, true><<<grid, FWS_C + 256, smem, st>>>(iq, n_total, ring, bp, tile_burst, n_tiles, dec_out);
    }
    return cudaGetLastError();
}

cudaError_t launch_fir(int fmt, int dec, const void *iq, int64_t n_total, uint64_t ring,
                       const BurstParam *bp, const int *tile_start, const int *tile_burst, int n_bursts, int n_tiles,
                       float2 *dec_out, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
#define IR_FIR_CASE(F, D) \
    if (fmt == F && dec == D) return launch_fir_t<F, D>(iq, n_total, ring, bp, tile_start, tile_burst, n_bursts, n_tiles, dec_out, st)
    IR_FIR_CASE(IR_FMT_CF32, 40);
    IR_FIR_CASE(IR_FMT_CI16, 40);
    IR_FIR_CASE(IR_FMT_CI8, 40);
    IR_FIR_CASE(IR_FMT_CF32, 48);
    IR_FIR_CASE(IR_FMT_CI16, 48);
    IR_FIR_CASE(IR_FMT_CI8, 48);
#undef IR_FIR_CASE
    return cudaErrorInvalidValue;     // decimation ratios other than 40 / 48: not built yet
}

// ============================================================== 250 kHz chain
// Remainder-loop arithmetic of the reference's FIR kernels as compiled (see oracle/ir_oracle.c
// tail_mac): chunks of 8 then one chunk of 4 unfused in tap order, last <4 taps fused.
template <class GET>
__device__ __forceinline__ float tail_mac(const float *h, int nt, GET get) {
    float a = 0.0f;
    int k = 0;
    const int k8 = nt & ~7;
    for (; k < k8; k++) a = a + h[k] * get(k);
    if (nt - k >= 4)
        for (int e = k + 4; k < e; k++) a = a + h[k] * get(k);
    for (; k < nt; k++) a = fmaf(h[k], get(k), a);
    return a;
}

// centred ("same") complex FIR with zero padding: out[i] = sum_k h[k] * x[i + k - half]
__device__ __forceinline__ float2 fir_same(const float *h, int nt, const float2 *__restrict__ x, int n, int i) {
    const int half = (nt - 1) / 2;
    const int body = n & ~3;
    if (i < body) {                                         // simd_avx2.c:28-44
        float ar = 0.0f, ai = 0.0f;
        for (int k = 0; k < nt; k++) {
            const int j = i + k - half;
            float2 v = (j >= 0 && j < n) ? x[j] : make_float2(0.0f, 0.0f);
            ar = fmaf(h[k], v.x, ar);
            ai = fmaf(h[k], v.y, ai);
        }
        return make_float2(ar, ai);
    }
    float ar = tail_mac(h, nt, [&](int k) { int j = i + k - half; return (j >= 0 && j < n) ? x[j].x : 0.0f; });
    float ai = tail_mac(h, nt, [&](int k) { int j = i + k - half; return (j >= 0 && j < n) ? x[j].y : 0.0f; });
    return make_float2(ar, ai);
}

// smoothed |x|^2 (burst_downmix.c:449-458): mag via one FMA (simd_avx2.c:297-318), 20-tap box
__device__ __forceinline__ float box_mag(const float2 *__restrict__ a, int flen, int i, int ntb) {
    auto mg = [&](int k) { return mag2_fma(a[i + k]); };
    if (i < (flen & ~7)) {                                  // simd_avx2.c:117-128
        float acc = 0.0f;
        for (int k = 0; k < ntb; k++) acc = fmaf(c_box[k], mg(k), acc);
        return acc;
    }
    return tail_mac(c_box, ntb, mg);
}

__device__ __forceinline__ float quad_peak(float a, float b, float c) {
    float den = a - 2.0f * b + c;
    if (fabsf(den) > 1e-10f) return 0.5f * (a - c) / den;
    return 0.0f;
}

// cexpf(j*ph) with the libm-quality the host would deliver: evaluated in double, rounded once.
__device__ __forceinline__ float2 unit_phasor(float ph) {
    double s, c;
    sincos((double)ph, &s, &c);
    return make_float2((float)c, (float)s);
}

struct ChainShared {
    float2 data[3 * (IR_CORR_N + 16)];        // >= fft_data_elems<12>()
    float2 tw12[fft_tw_elems<12>()];
    float2 tw11c[fft_twc_elems<11>()];        // compact per-stage tables of the 2048-pt FFTs (the 8.7 KB full table of
                                              // their first pass is read from global memory: 70 KB per CTA = 3 CTAs/SM)
    ArgMax red[32];
    float redf[32];
    int redi[32];
    float f0;
    int i0, i1;
    float2 c0;
};

__device__ __forceinline__ float block_max(float v, float *scratch) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float r = l < nw ? scratch[l] : -1e30f;
    return warp_max(r);
}

__device__ __forceinline__ int block_min_int(int v, int *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    int r = l < nw ? scratch[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}

__global__ void __launch_bounds__(256)
k_chain(const BurstParam *__restrict__ bp, int n_bursts, const float2 *__restrict__ dec_all,
        float2 *__restrict__ scrA, float2 *__restrict__ scrB, const float2 *__restrict__ tw4096,
        const float2 *__restrict__ tw2048, const float2 *__restrict__ sync_dl,
        const float2 *__restrict__ sync_ul, ChainOut *__restrict__ out, float2 *__restrict__ frames) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChainShared &S = *reinterpret_cast<ChainShared *>(smem_raw);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int b = blockIdx.x;
    if (b >= n_bursts) return;
    const BurstParam P = bp[b];
    const int dlen = P.dec_len;
    const float2 *dec = dec_all + P.dec_off;
    float2 *A = scrA + P.dec_off, *B = scrB + P.dec_off;
    ChainOut co;
    co.status = 0; co.start = 0; co.center_offset = 0; co.cfo_peak_bin = 0; co.direction = 0;
    co.corr_offset = 0; co.uw_start = 0; co.uw_corr = 0; co.corr_re = 0; co.corr_im = 0;
    co.frame_len = 0; co.incr_fine = make_float2(1.0f, 0.0f);
    if (dlen < 100) {                                       // burst_downmix.c:677-680
        if (tid == 0) { co.status = 2; out[b] = co; }
        return;
    }
    for (int i = tid; i < fft_tw_elems<12>(); i += nth) S.tw12[i] = tw4096[i];
    for (int i = tid; i < fft_twc_elems<11>(); i += nth) S.tw11c[i] = tw2048[fft_twfull_elems<11>() + i];
    const int n_noise = c_ntaps[1], n_box = c_ntaps[2], n_rrc = c_ntaps[3];

    // 2b: noise-limiting low-pass, centred (burst_downmix.c:683-698)
    for (int i = tid; i < dlen; i += nth) A[i] = fir_same(c_noise, n_noise, dec, dlen, i);
    __syncthreads();

    // 3: burst start (burst_downmix.c:441-478)
    int start;
    {
        const int search = min(IR_OUT_RATE, dlen);
        int mlen = search + n_box - 1;
        if (mlen > dlen) mlen = dlen;
        int flen = mlen - n_box + 1;
        if (flen <= 0) {
            start = 0;
        } else {
            if (flen > search) flen = search;
            float mx = -1e30f;
            for (int i = tid; i < flen; i += nth) mx = fmaxf(mx, box_mag(A, flen, i, n_box));
            mx = block_max(mx, S.redf);
            const float th = 0.45f * mx;
            int first = 0x7fffffff;
            for (int i = tid; i < flen; i += nth)
                if (box_mag(A, flen, i, n_box) >= th) { first = i; break; }
            first = block_min_int(first, S.redi);
            start = first == 0x7fffffff ? flen : first;
            if (start > 0) {
                start = start + (n_box - 1) / 2 - 25;           // pre_start_samples = 25 (:241)
                if (start < 0) start = 0;
            }
        }
    }
    co.start = start;
    if (start >= dlen - 100) {                              // :702-705
        if (tid == 0) { co.status = 3; out[b] = co; }
        return;
    }
    const int flen = dlen - start;

    // 4: fine CFO from the squared signal (:482-535)
    {
        const int m = min(IR_CFO_N, flen);
        for (int p = tid; p < IR_CFO_TOTAL; p += nth) {
            float2 v = make_float2(0.0f, 0.0f);
            if (p < m) {                                     // simd_avx2.c:345-388 as compiled
                const float2 x = A[start + p];
                const float sr = fmaf(x.x, x.x, -(x.y * x.y));
                const float si = 2.0f * (x.x * x.y);
                v = make_float2(sr * c_cfo_win[p], si * c_cfo_win[p]);
            }
            S.data[fft_pad<12>(p)] = v;
        }
        __syncthreads();
        fft_smem<12, false, false>(S.data, S.tw12, [&](int p) { return S.data[fft_pad<12>(p)]; },
                                   [](int, float2) {});
        ArgMax best{-1.0f, 0x7fffffff};
        for (int k = tid; k < IR_CFO_TOTAL; k += nth)
            best = argmax_pick(best, ArgMax{mag2_plain(fft_result<12>(S.data, k)), k});
        best = block_argmax(best, S.red);
        if (tid == 0) {
            const int T = IR_CFO_TOTAL;
            int bi = best.i;
            float bv = best.v;
            if (!(bv > 0.0f)) { bi = 0; bv = 0.0f; }          // strict '>' scan starting from 0
            const int ui = bi >= T / 2 ? bi - T : bi;
            float corr = 0.0f;
            if (bi > 0 && bi < T - 1) {
                const int im1 = ui - 1 < 0 ? ui - 1 + T : ui - 1;
                const int ip1 = ui + 1 < 0 ? ui + 1 + T : ui + 1;
                corr = quad_peak(mag2_plain(fft_result<12>(S.data, im1)), bv,
                                 mag2_plain(fft_result<12>(S.data, ip1)));
            }
            S.f0 = (ui + corr) / T / 2.0f;
            S.i0 = bi;
        }
        __syncthreads();
    }
    const float coff = S.f0;
    co.center_offset = coff;
    co.cfo_peak_bin = S.i0;

    // What later steps read of the shifted, matched-filtered signal: the first IR_SYNC_SEARCH samples (correlation)
    // and the frame, which starts at uw_start <= IR_SYNC_SEARCH - IR_SYNC_LEN + 1 + 320 and is at most 1910 / 4440
    // samples long.  The burst's extract carries post_len (16 ms = 4000 samples here) of tail behind that: nobody
    // reads it, so the shift and the matched filter stop where the frame can end (+ the filter's half width).
    const double cfreq = P.cfreq_coarse + (double)(coff * (float)IR_OUT_RATE);
    const bool simplex = cfreq > 1626000000.0;
    const int maxl = simplex ? 4440 : 1910, minl = simplex ? 800 : 1310;
    const int need = min(flen, IR_SYNC_SEARCH - IR_SYNC_LEN + 1 + 320 + maxl);
    const int n_rrc_half = c_ntaps[3] / 2;
    const int need_b = min(flen, need + n_rrc_half + 1);
    // 5: fine shift (:713-720): the phase recurrence is serial; one thread lays the phases down,
    // everybody applies them.
    {
        const float ph = -2.0f * (float)M_PI * coff;
        const float2 w = unit_phasor(ph);
        co.incr_fine = w;
        if (tid == 0) {
            float2 p = make_float2(1.0f, 0.0f);
            for (int i = 0; i < need_b; i++) { B[i] = p; p = cmul(p, w); }
        }
        __syncthreads();
        for (int i = tid; i < need_b; i += nth) B[i] = cmul(A[start + i], B[i]);
        __syncthreads();
    }
    // 6: matched filter, centred (:723-734) -> A[0..need).  The true length goes in (it decides which outputs are
    // the SIMD body and which the scalar tail of the reference's kernel, and where its zero padding begins); the
    // taps of output i < need reach B[i + half] < need_b at most, or past flen
    for (int i = tid; i < need; i += nth) A[i] = fir_same(c_rrc, n_rrc, B, flen, i);
    __syncthreads();

    // 7: sync-word correlation (:539-639)
    const int sl = min(IR_SYNC_SEARCH, flen);
    float2 *F = S.data, *PD = S.data + (IR_CORR_N + 16), *PU = S.data + 2 * (IR_CORR_N + 16);
    for (int p = tid; p < IR_CORR_N; p += nth)
        F[fft_pad<11>(p)] = p < sl ? A[p] : make_float2(0.0f, 0.0f);
    __syncthreads();
    fft_smem2<11, false, false, true>(F, tw2048, S.tw11c, [&](int p) { return F[fft_pad<11>(p)]; }, [](int, float2) {});
    for (int k = tid; k < IR_CORR_N; k += nth) {
        const float2 f = fft_result<11>(F, k);
        PD[fft_pad<11>(k)] = cmul(f, sync_dl[k]);
        PU[fft_pad<11>(k)] = cmul(f, sync_ul[k]);
    }
    __syncthreads();
    fft_smem2<11, true, false, true>(PD, tw2048, S.tw11c, [&](int p) { return PD[fft_pad<11>(p)]; }, [](int, float2) {});
    fft_smem2<11, true, false, true>(PU, tw2048, S.tw11c, [&](int p) { return PU[fft_pad<11>(p)]; }, [](int, float2) {});
    ArgMax bd{-1.0f, 0x7fffffff}, bu{-1.0f, 0x7fffffff};
    for (int i = tid; i < sl; i += nth) {
        bd = argmax_pick(bd, ArgMax{mag2_plain(fft_result<11>(PD, i)), i});
        bu = argmax_pick(bu, ArgMax{mag2_plain(fft_result<11>(PU, i)), i});
    }
    bd = block_argmax(bd, S.red);
    bu = block_argmax(bu, S.red);
    if (tid == 0) {
        float md = bd.v, mu = bu.v;
        int od = bd.i, ou = bu.i;
        if (!(md > 0.0f)) { md = 0.0f; od = 0; }
        if (!(mu > 0.0f)) { mu = 0.0f; ou = 0; }
        const float2 *R;
        int cofs, dir;
        if (md >= mu) { dir = 1; cofs = od; R = PD; } else { dir = 2; cofs = ou; R = PU; }
        const float2 cr = fft_result<11>(R, cofs);
        float uwc = 0.0f;
        if (cofs > 0 && cofs < sl - 1)
            uwc = quad_peak(mag2_plain(fft_result<11>(R, cofs - 1)), mag2_plain(cr),
                            mag2_plain(fft_result<11>(R, cofs + 1)));
        const int pre_syms = dir == 1 ? 16 : 32;               // :633-634 (quirk kept)
        S.i0 = dir;
        S.i1 = cofs;
        S.c0 = cr;
        S.f0 = uwc;
        (void)pre_syms;
    }
    __syncthreads();
    const int dir = S.i0, cofs = S.i1;
    const float2 cres = S.c0;
    const int uw_start = cofs - IR_SYNC_LEN + 1 + (dir == 1 ? 16 : 32) * 10;
    co.direction = dir;
    co.corr_offset = cofs;
    co.uw_start = uw_start;
    co.uw_corr = S.f0;
    co.corr_re = cres.x;
    co.corr_im = cres.y;
    if (uw_start < 0 || uw_start >= flen) {                 // :744-747
        if (tid == 0) { co.status = 7; out[b] = co; }
        return;
    }
    // 8: phase alignment (:750-760).  cabsf == (float)sqrt((double)re*re + (double)im*im) in glibc.
    float2 pc;
    {
        const float mg = (float)sqrt((double)cres.x * (double)cres.x + (double)cres.y * (double)cres.y);
        pc = mg > 0.0f ? make_float2(cres.x / mg, -(cres.y / mg)) : make_float2(1.0f, 0.0f);
    }
    // 9: extraction (:763-793)
    // center_frequency after both shifts decides the frame-length limits (:671,:719,:764-770)
    const int avail = flen - uw_start;
    if (avail < minl) {
        if (tid == 0) { co.status = 9; out[b] = co; }
        return;
    }
    const int xl = avail < maxl ? avail : maxl;
    co.frame_len = xl;
    float2 *fr = frames + (size_t)b * IR_MAX_FRAME;
    for (int i = tid; i < xl; i += nth) fr[i] = cmul(A[uw_start + i], pc);
    if (tid == 0) out[b] = co;
}

cudaError_t launch_chain(const BurstParam *bp, int n_bursts, const float2 *dec, float2 *scrA,
                         float2 *scrB, const float2 *tw4096, const float2 *tw2048,
                         const float2 *sync_dl, const float2 *sync_ul, ChainOut *out,
                         float2 *frames, cudaStream_t st) {
    if (n_bursts <= 0) return cudaSuccess;
    const size_t smem = sizeof(ChainShared);
    cudaError_t e = cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_chain<<<n_bursts, 256, smem, st>>>(bp, n_bursts, dec, scrA, scrB, tw4096, tw2048, sync_dl,
                                         sync_ul, out, frames);
    return cudaGetLastError();
}

}  // namespace ir
