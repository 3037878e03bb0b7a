// k_demod.cu -- DQPSK demodulation of the extracted frames (qpsk_demod.c:393-535).
//
// Timing recovery and the phase loop are serial recurrences over <= 445 symbols, so the
// parallelism is across frames: one thread per frame, 32 frames per CTA, every frame of a run
// in one launch.  All arithmetic is the reference's scalar C (no fused multiply-adds; the
// translation unit is built with -fmad=false).  libm calls (atan2f, sinf, cosf, cabsf) are
// evaluated in double and rounded once, which reproduces glibc's float results except in
// rare last-bit cases (DESIGN.md "Exactness").
#include <stdlib.h>
#include <string.h>

#include "ir_device.cuh"
#include "ir_internal.h"

namespace ir {

__device__ __forceinline__ float2 cscale(float a, float2 v) { return make_float2(a * v.x, a * v.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ float f_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float f_hypot(float x, float y) {
    return (float)sqrt((double)x * (double)x + (double)y * (double)y);
}

// Catmull-Rom interpolation at fractional position (qpsk_demod.c:56-81)
__device__ __forceinline__ float2 catmull(const float2 *__restrict__ in, int n, float pos) {
    int idx = (int)pos;
    const float mu = pos - idx;
    if (idx < 1) idx = 1;
    if (idx >= n - 2) idx = n - 3;
    const float2 s0 = in[idx - 1], s1 = in[idx], s2 = in[idx + 1], s3 = in[idx + 2];
    const float mu2 = mu * mu, mu3 = mu2 * mu;
    // a = -0.5 s0 + 1.5 s1 - 1.5 s2 + 0.5 s3, evaluated left to right per component
    const float2 a = cadd(csub(cadd(cscale(-0.5f, s0), cscale(1.5f, s1)), cscale(1.5f, s2)), cscale(0.5f, s3));
    const float2 b = csub(cadd(csub(s0, cscale(2.5f, s1)), cscale(2.0f, s2)), cscale(0.5f, s3));
    const float2 c = cadd(cscale(-0.5f, s0), cscale(0.5f, s2));
    // a*mu3 + b*mu2 + c*mu + d
    return cadd(cadd(cadd(cscale(mu3, a), cscale(mu2, b)), cscale(mu, c)), s1);
}

__constant__ int c_uw_dl[12] = {0, 2, 2, 2, 2, 0, 0, 0, 2, 0, 0, 2};   // iridium.h:30
__constant__ int c_uw_ul[12] = {2, 2, 0, 0, 0, 2, 0, 0, 2, 0, 2, 2};   // iridium.h:31

__global__ void __launch_bounds__(32)
k_demod(const ChainOut *__restrict__ co, int n_bursts, const float2 *__restrict__ frames,
        int use_gardner, DemodOut *__restrict__ out, uint8_t *__restrict__ bits_all,
        float *__restrict__ llr_all) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bursts) return;
    DemodOut d;
    d.ok = 0; d.direction = co[b].direction; d.confidence = 0; d.level = 0; d.n_symbols = 0;
    d.n_raw_symbols = 0; d.total_phase = 0; d.pad = 0;
    if (co[b].status != 0) { out[b] = d; return; }
    const float2 *in = frames + (size_t)b * IR_MAX_FRAME;
    const int n = co[b].frame_len;
    uint8_t *bits = bits_all + (size_t)b * 2 * IR_MAX_SYMS;
    float2 *pl = reinterpret_cast<float2 *>(llr_all + (size_t)b * 2 * IR_MAX_SYMS);   // PLL output, later LLR
    const float sps = 10.0f;
    const float PI_F = (float)M_PI;

    // ---- symbol timing (qpsk_demod.c:85-141) fused with the phase loop (:145-195) and the
    //      hard decisions (:199-243): each symbol is consumed as soon as it is interpolated.
    int k = 0;                    // symbols produced
    float pos = 0.0f, integ = 0.0f;
    float2 prev = make_float2(0.0f, 0.0f);
    float2 phi = make_float2(1.0f, 0.0f);
    float total = 0.0f;
    const float r = 0.70710678118654752f;
    // decision bookkeeping
    float mx = 0.0f;
    int low = 0, nv = 0;
    bool ended = false;
    float sum_hist[4] = {0, 0, 0, 0};   // running sums of |x| after symbol i, i-1, i-2, i-3
    int ok_hist[4] = {0, 0, 0, 0};
    float sum_run = 0.0f;
    int ok_run = 0;
    int isamp = 0;                // --no-gardner sample cursor

    for (;;) {
        float2 now;
        if (use_gardner) {
            if (!(pos < (float)(n - 3))) break;
            now = catmull(in, n, pos);
        } else {
            if (isamp >= n) break;
            now = in[isamp];
        }
        // ---- PLL step on symbol k
        const float2 y = cmul(now, phi);
        pl[k] = y;
        {
            float2 ideal;
            if (y.x >= 0 && y.y >= 0) ideal = make_float2(r, r);
            else if (y.x >= 0) ideal = make_float2(r, -r);
            else if (y.y < 0) ideal = make_float2(-r, -r);
            else ideal = make_float2(-r, r);
            const float2 er = cmul(make_float2(ideal.x, -ideal.y), y);
            const float em = f_hypot(er.x, er.y);
            if (!(em < 1e-10f)) {
                const float2 unit = make_float2(er.x / em, er.y / em);
                const float ang = f_atan2(unit.y, unit.x);
                const float sa = 0.2f * ang;
                double sn, cs;
                sincos((double)sa, &sn, &cs);
                const float2 corr = make_float2((float)cs, (float)sn);
                total += sa;
                phi = cmul(make_float2(corr.x, -corr.y), phi);
                const float pm = f_hypot(phi.x, phi.y);
                if (pm > 0) phi = make_float2(phi.x / pm, phi.y / pm);
            }
        }
        // ---- hard decision on symbol k (until the end-of-frame rule fires)
        if (!ended) {
            const float mg = sqrtf(y.x * y.x + y.y * y.y);
            if (mg > mx) mx = mg;
            int sy;
            if (y.x >= 0 && y.y >= 0) sy = 0;
            else if (y.x < 0 && y.y >= 0) sy = 1;
            else if (y.x < 0) sy = 2;
            else sy = 3;
            bits[2 * k] = (uint8_t)sy;
            const float phs = (f_atan2(y.y, y.x) + PI_F) * 180.0f / PI_F;
            const float off = 45.0f - fmodf(phs, 90.0f);
            sum_run += mg;
            if (fabsf(off) <= 22.0f) ok_run++;
            sum_hist[3] = sum_hist[2]; sum_hist[2] = sum_hist[1]; sum_hist[1] = sum_hist[0]; sum_hist[0] = sum_run;
            ok_hist[3] = ok_hist[2]; ok_hist[2] = ok_hist[1]; ok_hist[1] = ok_hist[0]; ok_hist[0] = ok_run;
            nv++;
            if (mg < mx / 8.0f) {
                if (++low >= 3) { nv -= 3; ended = true; }
            } else {
                low = 0;
            }
        }
        // ---- Gardner loop update (uses the un-rotated samples)
        if (use_gardner) {
            if (k > 0) {
                const float mp = pos - sps * 0.5f;
                if (mp >= 1.0f) {
                    const float2 mid = catmull(in, n, mp);
                    const float2 df = csub(prev, now);
                    float e = df.x * mid.x - df.y * (-mid.y);      // Re{(prev-now) * conj(mid)}
                    if (e > 1.0f) e = 1.0f;
                    if (e < -1.0f) e = -1.0f;
                    integ += 0.0002f * e;
                    float adj = 0.02f * e + integ;
                    if (adj > 0.5f) adj = 0.5f;
                    if (adj < -0.5f) adj = -0.5f;
                    pos += adj;
                }
            }
            prev = now;
            pos += sps;
        } else {
            isamp += 10;
        }
        k++;
        if (k >= IR_MAX_SYMS) break;
    }
    d.n_raw_symbols = k;
    d.total_phase = total;
    // confidence / level over the first nv symbols (:245-255)
    float level;
    int conf;
    {
        float sum; int okc;
        if (ended) { sum = sum_hist[3]; okc = ok_hist[3]; }     // drop the three weak symbols
        else { sum = sum_run; okc = ok_run; }
        level = nv > 0 ? sum / nv : 0.0f;
        conf = nv > 0 ? (100 * okc) / nv : 0;
    }
    // ---- unique word (:277-325, :429-465)
    bool accept = true;
    {
        bool okd = false, oku = false;
        if (nv >= 12) {
            int dd = 0, du = 0;
            for (int i = 0; i < 12; i++) {
                int s = bits[2 * i];
                int a = abs(s - c_uw_dl[i]); if (a == 3) a = 1; dd += a;
                int c = abs(s - c_uw_ul[i]); if (c == 3) c = 1; du += c;
            }
            okd = dd <= 2; oku = du <= 2;
        }
        if (!okd && !oku) {
            float ed = 999.0f, eu = 999.0f;
            if (nv >= 12) {
                ed = 0.0f; eu = 0.0f;
                for (int i = 0; i < 12; i++) {
                    float act = f_atan2(pl[i].y, pl[i].x);
                    if (act < 0) act += 2.0f * PI_F;
                    const float xd = PI_F * 0.25f + c_uw_dl[i] * PI_F * 0.5f;
                    float d1 = act - xd;
                    if (d1 > PI_F) d1 -= 2.0f * PI_F;
                    if (d1 < -PI_F) d1 += 2.0f * PI_F;
                    ed += fabsf(d1) * (float)(2.0 / M_PI);
                    const float xu = PI_F * 0.25f + c_uw_ul[i] * PI_F * 0.5f;
                    float d2 = act - xu;
                    if (d2 > PI_F) d2 -= 2.0f * PI_F;
                    if (d2 < -PI_F) d2 += 2.0f * PI_F;
                    eu += fabsf(d2) * (float)(2.0 / M_PI);
                }
            }
            const float em = ed < eu ? ed : eu;
            if (em > 3.0f) accept = false;
            else d.direction = eu < ed ? 2 : 1;
        } else if (oku && !okd) {
            d.direction = 2;
        } else if (okd && !oku) {
            d.direction = 1;
        }
    }
    if (accept) {
        // LLR (:489-503) -- needs the mean |pl| first, then overwrites pl in place
        float sm = 0.0f;
        for (int i = 0; i < nv; i++) sm += f_hypot(pl[i].x, pl[i].y);
        const float sc = (nv > 0 && sm > 0) ? (0.70710678118654752f / (sm / nv)) : 1.0f;
        float *llr = reinterpret_cast<float *>(pl);
        for (int i = 0; i < nv; i++) {
            const float2 v = pl[i];
            llr[2 * i] = fabsf(v.x) * sc;
            llr[2 * i + 1] = fabsf(v.y) * sc;
        }
        // differential decode + bit mapping (:264-273, :329-335)
        int old = 0;
        for (int i = 0; i < nv; i++) {
            const int s = bits[2 * i];
            const int df = (s - old + 4) % 4;
            old = s;
            const int v = (df == 0) ? 0 : (df == 1) ? 2 : (df == 2) ? 3 : 1;     // {0,2,3,1}
            bits[2 * i] = (uint8_t)((v >> 1) & 1);
            bits[2 * i + 1] = (uint8_t)(v & 1);
        }
        d.ok = 1;
        d.confidence = conf;
        d.level = level;
        d.n_symbols = nv;
    }
    out[b] = d;
}

// ---- FPW frames per warp, staged in shared memory: the default.
// Symbol timing (Gardner + Catmull-Rom) and the decision-directed phase loop are serial recurrences per frame; a
// warp that gave all its lanes to ONE frame would spend 32 issue slots per step of a recurrence only one lane can
// advance (measured: 1400 such warps per wave cost more SM time than the 250 kHz chain), and the one-thread-per-frame
// kernel above pays an L2 round trip per Catmull-Rom tap (0.93 ms per launch).  Here lanes 0..FPW-1 of a warp each run
// the recurrences of one frame, in step (SIMT), on samples that sit in shared memory; everything per symbol that does
// not feed a recurrence (magnitudes, quadrants, the atan2 / fmodf of the confidence test, unique-word checks, hypot for
// the LLR scale, LLRs, differential decoding, bit mapping) is done by all 32 lanes, frame after frame, with coalesced
// stores.  Sums the reference accumulates in symbol order (level, LLR scale, soft unique-word distances) are added in
// that order by the frame's lane.  Arithmetic is k_demod's expression for expression: identical bytes.
struct DemodFrameShared {
    float2 now[IR_MAX_SYMS];           // interpolated symbols (decimate_gardner / decimate_simple)
    float2 y[IR_MAX_SYMS];             // after the phase loop
    float mg[IR_MAX_SYMS];             // |y| (float), later |y| by hypot for the LLR scale
    unsigned char sy[IR_MAX_SYMS];     // quadrant
    unsigned char okf[IR_MAX_SYMS];    // |offset| <= 22 degrees
    float uw[2][12];
};
template <int FPW>
__global__ void __launch_bounds__(32)
k_demod_g(const ChainOut *__restrict__ co, int n_bursts, const float2 *__restrict__ frames,
          int use_gardner, DemodOut *__restrict__ out, uint8_t *__restrict__ bits_all,
          float *__restrict__ llr_all, int max_frame) {
    extern __shared__ __align__(16) unsigned char dsm[];
    DemodFrameShared *SF = reinterpret_cast<DemodFrameShared *>(dsm);
    float2 *sin_all = reinterpret_cast<float2 *>(dsm + ((FPW * sizeof(DemodFrameShared) + 15) / 16) * 16);
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * FPW;
    const float PI_F = (float)M_PI;
    // this lane's frame (serial phases): lanes >= FPW mirror nothing
    const int fb = b0 + lane;
    const bool mine = lane < FPW && fb < n_bursts;
    DemodOut d;
    d.ok = 0; d.direction = mine ? co[fb].direction : 0; d.confidence = 0; d.level = 0; d.n_symbols = 0;
    d.n_raw_symbols = 0; d.total_phase = 0; d.pad = 0;
    const bool live = mine && co[fb].status == 0;
    const int n = live ? min(co[fb].frame_len, max_frame) : 0;
    // ---- stage the frames
    int nf[FPW];
#pragma unroll
    for (int f = 0; f < FPW; f++) {
        nf[f] = __shfl_sync(0xffffffffu, n, f);
        const float2 *in = frames + (size_t)(b0 + f) * IR_MAX_FRAME;
        float2 *dst = sin_all + (size_t)f * max_frame;
        for (int i = lane; i < nf[f]; i += 32) dst[i] = __ldg(in + i);
    }
    __syncwarp();
    DemodFrameShared &S = SF[lane < FPW ? lane : 0];
    const float2 *sin = sin_all + (size_t)(lane < FPW ? lane : 0) * max_frame;
    // ---- symbol timing (qpsk_demod.c:85-141)
    int K = 0;
    if (live) {
        if (use_gardner) {
            const float sps = 10.0f;
            int k = 0;
            float pos = 0.0f, integ = 0.0f;
            float2 prev = make_float2(0.0f, 0.0f);
            for (;;) {
                if (!(pos < (float)(n - 3))) break;
                const float2 now = catmull(sin, n, pos);
                S.now[k] = now;
                if (k > 0) {
                    const float mp = pos - sps * 0.5f;
                    if (mp >= 1.0f) {
                        const float2 mid = catmull(sin, n, mp);
                        const float2 df = csub(prev, now);
                        float e = df.x * mid.x - df.y * (-mid.y);      // Re{(prev-now) * conj(mid)}
                        if (e > 1.0f) e = 1.0f;
                        if (e < -1.0f) e = -1.0f;
                        integ += 0.0002f * e;
                        float adj = 0.02f * e + integ;
                        if (adj > 0.5f) adj = 0.5f;
                        if (adj < -0.5f) adj = -0.5f;
                        pos += adj;
                    }
                }
                prev = now;
                pos += sps;
                k++;
                if (k >= IR_MAX_SYMS) break;
            }
            K = k;
        } else {
            K = min((n + 9) / 10, IR_MAX_SYMS);
            for (int k = 0; k < K; k++) S.now[k] = sin[10 * k];
        }
    }
    // ---- phase loop (:145-195)
    float total = 0.0f;
    if (live) {
        const float r = 0.70710678118654752f;
        float2 phi = make_float2(1.0f, 0.0f);
        for (int k = 0; k < K; k++) {
            const float2 y = cmul(S.now[k], phi);
            S.y[k] = y;
            float2 ideal;
            if (y.x >= 0 && y.y >= 0) ideal = make_float2(r, r);
            else if (y.x >= 0) ideal = make_float2(r, -r);
            else if (y.y < 0) ideal = make_float2(-r, -r);
            else ideal = make_float2(-r, r);
            const float2 er = cmul(make_float2(ideal.x, -ideal.y), y);
            const float em = f_hypot(er.x, er.y);
            if (!(em < 1e-10f)) {
                const float2 unit = make_float2(er.x / em, er.y / em);
                const float ang = f_atan2(unit.y, unit.x);
                const float sa = 0.2f * ang;
                double sn, cs;
                sincos((double)sa, &sn, &cs);
                const float2 corr = make_float2((float)cs, (float)sn);
                total += sa;
                phi = cmul(make_float2(corr.x, -corr.y), phi);
                const float pm = f_hypot(phi.x, phi.y);
                if (pm > 0) phi = make_float2(phi.x / pm, phi.y / pm);
            }
        }
    }
    d.n_raw_symbols = K;
    d.total_phase = total;
    __syncwarp();
    // ---- hard decisions (:199-243): per symbol, all lanes, frame after frame
    int Kf[FPW];
#pragma unroll
    for (int f = 0; f < FPW; f++) {
        Kf[f] = __shfl_sync(0xffffffffu, K, f);
        DemodFrameShared &T = SF[f];
        for (int k = lane; k < Kf[f]; k += 32) {
            const float2 y = T.y[k];
            T.mg[k] = sqrtf(y.x * y.x + y.y * y.y);
            int sy;
            if (y.x >= 0 && y.y >= 0) sy = 0;
            else if (y.x < 0 && y.y >= 0) sy = 1;
            else if (y.x < 0) sy = 2;
            else sy = 3;
            T.sy[k] = (unsigned char)sy;
            const float phs = (f_atan2(y.y, y.x) + PI_F) * 180.0f / PI_F;
            const float off = 45.0f - fmodf(phs, 90.0f);
            T.okf[k] = fabsf(off) <= 22.0f ? 1 : 0;
        }
    }
    __syncwarp();
    // end-of-frame rule, level and confidence: running max / running sums in symbol order
    int nv = 0, conf = 0;
    float level = 0.0f;
    if (live) {
        float mx = 0.0f, sum_run = 0.0f;
        int low = 0, ok_run = 0;
        bool ended = false;
        float sum_hist[4] = {0, 0, 0, 0};
        int ok_hist[4] = {0, 0, 0, 0};
        for (int k = 0; k < K && !ended; k++) {
            const float mg = S.mg[k];
            if (mg > mx) mx = mg;
            sum_run += mg;
            if (S.okf[k]) ok_run++;
            sum_hist[3] = sum_hist[2]; sum_hist[2] = sum_hist[1]; sum_hist[1] = sum_hist[0]; sum_hist[0] = sum_run;
            ok_hist[3] = ok_hist[2]; ok_hist[2] = ok_hist[1]; ok_hist[1] = ok_hist[0]; ok_hist[0] = ok_run;
            nv++;
            if (mg < mx / 8.0f) {
                if (++low >= 3) { nv -= 3; ended = true; }
            } else {
                low = 0;
            }
        }
        float sum; int okc;
        if (ended) { sum = sum_hist[3]; okc = ok_hist[3]; }     // drop the three weak symbols
        else { sum = sum_run; okc = ok_run; }
        level = nv > 0 ? sum / nv : 0.0f;
        conf = nv > 0 ? (100 * okc) / nv : 0;
    }
    // ---- unique word (:277-325, :429-465), on the frame's lane
    bool accept = live;
    if (live) {
        bool okd = false, oku = false;
        if (nv >= 12) {
            int dd = 0, du = 0;
            for (int i = 0; i < 12; i++) {
                const int sv = S.sy[i];
                int a = abs(sv - c_uw_dl[i]); if (a == 3) a = 1; dd += a;
                int c = abs(sv - c_uw_ul[i]); if (c == 3) c = 1; du += c;
            }
            okd = dd <= 2; oku = du <= 2;
        }
        if (!okd && !oku) {
            float ed = 999.0f, eu = 999.0f;
            if (nv >= 12) {
                ed = 0.0f; eu = 0.0f;
                for (int i = 0; i < 12; i++) {
                    float act = f_atan2(S.y[i].y, S.y[i].x);
                    if (act < 0) act += 2.0f * PI_F;
                    const float xd = PI_F * 0.25f + c_uw_dl[i] * PI_F * 0.5f;
                    float d1 = act - xd;
                    if (d1 > PI_F) d1 -= 2.0f * PI_F;
                    if (d1 < -PI_F) d1 += 2.0f * PI_F;
                    ed += fabsf(d1) * (float)(2.0 / M_PI);
                    const float xu = PI_F * 0.25f + c_uw_ul[i] * PI_F * 0.5f;
                    float d2 = act - xu;
                    if (d2 > PI_F) d2 -= 2.0f * PI_F;
                    if (d2 < -PI_F) d2 += 2.0f * PI_F;
                    eu += fabsf(d2) * (float)(2.0 / M_PI);
                }
            }
            const float em = ed < eu ? ed : eu;
            if (em > 3.0f) accept = false;
            else d.direction = eu < ed ? 2 : 1;
        } else if (oku && !okd) {
            d.direction = 2;
        } else if (okd && !oku) {
            d.direction = 1;
        }
    }
    // ---- LLR (:489-503): the scale is the mean of |y| by hypot, summed in symbol order
    int nvf[FPW];
#pragma unroll
    for (int f = 0; f < FPW; f++) {
        nvf[f] = __shfl_sync(0xffffffffu, accept ? nv : 0, f);
        DemodFrameShared &T = SF[f];
        for (int i = lane; i < nvf[f]; i += 32) T.mg[i] = f_hypot(T.y[i].x, T.y[i].y);
    }
    __syncwarp();
    float sc = 1.0f;
    if (accept) {
        float sm = 0.0f;
        for (int i = 0; i < nv; i++) sm += S.mg[i];
        sc = (nv > 0 && sm > 0) ? (0.70710678118654752f / (sm / nv)) : 1.0f;
    }
#pragma unroll
    for (int f = 0; f < FPW; f++) {
        const float scf = __shfl_sync(0xffffffffu, sc, f);
        if (nvf[f] <= 0) continue;
        DemodFrameShared &T = SF[f];
        uint8_t *bits = bits_all + (size_t)(b0 + f) * 2 * IR_MAX_SYMS;
        float *llr = llr_all + (size_t)(b0 + f) * 2 * IR_MAX_SYMS;
        const float *yv = reinterpret_cast<const float *>(T.y);
        for (int i = lane; i < 2 * nvf[f]; i += 32) {
            llr[i] = fabsf(yv[i]) * scf;
            // differential decode + bit mapping (:264-273, :329-335)
            const int sidx = i >> 1;
            const int sv = T.sy[sidx], old = sidx > 0 ? T.sy[sidx - 1] : 0;
            const int df = (sv - old + 4) % 4;
            const int v = (df == 0) ? 0 : (df == 1) ? 2 : (df == 2) ? 3 : 1;     // {0,2,3,1}
            bits[i] = (uint8_t)((i & 1) ? (v & 1) : ((v >> 1) & 1));
        }
    }
    if (accept) {
        d.ok = 1;
        d.confidence = conf;
        d.level = level;
        d.n_symbols = nv;
    }
    if (mine) out[fb] = d;
}

cudaError_t launch_demod(const ChainOut *co, int n_bursts, const float2 *frames, int use_gardner,
                         DemodOut *out, uint8_t *bits, float *llr, int max_frame, cudaStream_t st) {
    if (n_bursts <= 0) return cudaSuccess;
    // Which slicer: IR_DEMOD=thread | group forces one; by default the frames-in-shared-memory kernel takes the small
    // launches (<= 512 frames: a tail wave, a single frame through qpsk_demod(): latency is what counts, a few CTAs hold a few SMs)
    // and the one-thread-per-frame kernel the big waves (SIMT across 32 frames; no shared memory, so it never keeps
    // the FIR / chain CTAs of later waves off an SM -- measured in-run).
    static const char *mode = getenv("IR_DEMOD");
    const bool legacy = mode ? strcmp(mode, "thread") == 0 : n_bursts > 512;
    if (legacy) {
        k_demod<<<(n_bursts + 31) / 32, 32, 0, st>>>(co, n_bursts, frames, use_gardner, out, bits, llr);
        return cudaGetLastError();
    }
    if (max_frame <= 0 || max_frame > IR_MAX_FRAME) max_frame = IR_MAX_FRAME;
    max_frame = (max_frame + 1) & ~1;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_demod_g<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_demod_g<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    auto smem_of = [&](int fpw) { return ((fpw * sizeof(DemodFrameShared) + 15) / 16) * 16 + (size_t)fpw * max_frame * sizeof(float2); };
    if (smem_of(8) <= (size_t)226 * 1024)
        k_demod_g<8><<<(n_bursts + 7) / 8, 32, smem_of(8), st>>>(co, n_bursts, frames, use_gardner, out, bits, llr, max_frame);
    else
        k_demod_g<4><<<(n_bursts + 3) / 4, 32, smem_of(4), st>>>(co, n_bursts, frames, use_gardner, out, bits, llr, max_frame);
    return cudaGetLastError();
}

}  // namespace ir
