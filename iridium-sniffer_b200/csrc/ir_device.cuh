// ir_device.cuh -- device-side building blocks shared by every kernel of the path.
//
// Arithmetic contract (DESIGN.md "Exactness"): every translation unit is compiled with
// -fmad=false, so a*b+c is fused ONLY where fmaf() is written.  That lets each kernel
// mirror the reference's AVX2 build operation for operation (explicit FMA intrinsics in
// simd_avx2.c, unfused scalar C elsewhere) and makes GPU results bit-identical to the CPU
// restatement wherever no libm call is involved.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IR_FMT_CF32 0
#define IR_FMT_CI16 1
#define IR_FMT_CI8 2

namespace ir {

// ---------------------------------------------------------------- sample formats
// cf32 passes through (burst_detect.c:862-863); ci8 -> x/128 (simd_avx2.c:264-294);
// ci16 keeps the upper byte first (main.c:245-246) and is then treated as ci8.
template <int FMT>
__device__ __forceinline__ float2 load_sample(const void *__restrict__ base, int64_t i) {
    if (FMT == IR_FMT_CF32) {
        return __ldg(reinterpret_cast<const float2 *>(base) + i);
    } else if (FMT == IR_FMT_CI16) {
        short2 v = __ldg(reinterpret_cast<const short2 *>(base) + i);
        int a = (int)(signed char)(v.x >> 8), b = (int)(signed char)(v.y >> 8);
        return make_float2((float)a / 128.0f, (float)b / 128.0f);
    } else {
        char2 v = __ldg(reinterpret_cast<const char2 *>(base) + i);
        return make_float2((float)v.x / 128.0f, (float)v.y / 128.0f);
    }
}

// the same in two steps for the integer formats: the raw sample (4 or 2 bytes), so that many loads can be in flight
// before the first conversion
template <int FMT>
__device__ __forceinline__ uint32_t load_raw(const void *__restrict__ base, int64_t i) {
    if (FMT == IR_FMT_CI16) return __ldg(reinterpret_cast<const uint32_t *>(base) + i);
    return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(base) + i);
}
template <int FMT>
__device__ __forceinline__ float2 conv_raw(uint32_t r) {
    if (FMT == IR_FMT_CI16) {
        const int a = (int)(signed char)(r >> 8), b = (int)(signed char)(r >> 24);
        return make_float2((float)a / 128.0f, (float)b / 128.0f);
    }
    const int a = (int)(signed char)(r & 0xffu), b = (int)(signed char)((r >> 8) & 0xffu);
    return make_float2((float)a / 128.0f, (float)b / 128.0f);
}

__host__ __device__ inline int fmt_bytes(int fmt) {
    return fmt == IR_FMT_CF32 ? 8 : (fmt == IR_FMT_CI16 ? 4 : 2);
}

// Unfused complex product exactly as gcc expands `float complex * float complex` without
// FMA hardware flags (rotator.h:40-41, burst_downmix.c:554-555, qpsk_demod.c:153,170,186).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float mag2_fma(float2 v) {      // simd_avx2.c:198-199,311
    return fmaf(v.x, v.x, v.y * v.y);
}
__device__ __forceinline__ float mag2_plain(float2 v) {    // burst_downmix.c:498-500
    return v.x * v.x + v.y * v.y;
}

// ---------------------------------------------------------------- FFT engine
// Radix-2 decimation-in-frequency, executed 3..5 stages at a time in registers.  The
// butterfly and the twiddle table are the ones oracle/ir_oracle.c::orc_fft uses, so the
// grouping into passes changes where operands live, never their values.
//
//   stage s (0-based), block size B = N>>s, half = B/2:
//     a' = a + b ; d = a - b ; b' = ( fma(d.re,w.re,-(d.im*w.im)), fma(d.re,w.im, d.im*w.re) )
//     with w = W_N^(j<<s), j = position inside the half block.
//   After all stages position p holds X[bitrev(p)].
//
// Shared-memory layout: element p lives at p + (p >> (L-4))  (conflict-free for every pass
// incl. the bit-reversed last one; verified by simulation).
// Twiddles: full table W[k], k < N/2 at k + (k>>4) for the stages of the first pass;
// per-stage compact tables Wc_s[j] = W[j<<s] for all later stages.

template <int L> struct FftPlan;
template <> struct FftPlan<9>  { static constexpr int P = 2; static constexpr int Q0 = 5, Q1 = 4, Q2 = 0; };
template <> struct FftPlan<10> { static constexpr int P = 2; static constexpr int Q0 = 5, Q1 = 5, Q2 = 0; };
template <> struct FftPlan<11> { static constexpr int P = 3; static constexpr int Q0 = 4, Q1 = 4, Q2 = 3; };
template <> struct FftPlan<12> { static constexpr int P = 3; static constexpr int Q0 = 4, Q1 = 4, Q2 = 4; };
template <> struct FftPlan<13> { static constexpr int P = 3; static constexpr int Q0 = 5, Q1 = 4, Q2 = 4; };
template <> struct FftPlan<14> { static constexpr int P = 3; static constexpr int Q0 = 5, Q1 = 5, Q2 = 4; };

template <int L> __host__ __device__ constexpr int fft_data_elems() { return (1 << L) + 16; }
template <int L> __host__ __device__ constexpr int fft_twfull_elems() { return (1 << (L - 1)) + (1 << (L - 5)) + 1; }
template <int L> __host__ __device__ constexpr int fft_twc_elems() { return (1 << (L - FftPlan<L>::Q0)); }
template <int L> __host__ __device__ constexpr int fft_tw_elems() { return fft_twfull_elems<L>() + fft_twc_elems<L>(); }
template <int L> __host__ __device__ constexpr int fft_threads() {
    return 1 << (L - (FftPlan<L>::Q0 > FftPlan<L>::Q1 ? FftPlan<L>::Q0 : FftPlan<L>::Q1));
}

template <int L> __device__ __forceinline__ int fft_pad(int p) { return p + (p >> (L - 4)); }
__host__ __device__ __forceinline__ int tw_pad(int k) { return k + (k >> 4); }
// offset of the compact table of stage s (s >= Q0) inside the compact area
template <int L> __host__ __device__ constexpr int twc_off(int s) {
    // sum_{t=Q0}^{s-1} N>>(t+1)  =  (N>>Q0) - (N>>s)
    return (1 << (L - FftPlan<L>::Q0)) - (1 << (L - s));
}

template <bool INV>
__device__ __forceinline__ void bfly(float2 &a, float2 &b, float2 w) {
    if (INV) w.y = -w.y;
    float dr = a.x - b.x, di = a.y - b.y;
    a.x = a.x + b.x;
    a.y = a.y + b.y;
    float p = di * w.y;
    float q = di * w.x;
    b.x = fmaf(dr, w.x, -p);
    b.y = fmaf(dr, w.y, q);
}
// w = 1 : (d.re, d.im)      w = -j (forward) / +j (inverse) at k = N/4
template <bool INV>
__device__ __forceinline__ void bfly_w1(float2 &a, float2 &b) {
    float dr = a.x - b.x, di = a.y - b.y;
    a.x = a.x + b.x;
    a.y = a.y + b.y;
    b.x = dr;
    b.y = di;
}
template <bool INV>
__device__ __forceinline__ void bfly_wq(float2 &a, float2 &b) {
    float dr = a.x - b.x, di = a.y - b.y;
    a.x = a.x + b.x;
    a.y = a.y + b.y;
    if (INV) { b.x = -di; b.y = dr; } else { b.x = di; b.y = -dr; }
}

// One pass over one group: load 2^Q points, run Q stages, hand the results to `st`.
// LD(pos) -> float2, ST(pos, m, value).  g = group index in [0, N>>Q).
// twf: the full table (stages of the first pass), twc: the compact per-stage tables.  TWG: twf points to global
// memory (read through the read-only path; the detector's FFT keeps only the 2 KB of compact tables in shared
// memory so that three frames fit an SM) -- same values either way.
template <int L, int S0, int Q, bool INV, bool TWG, class LD, class ST>
__device__ __forceinline__ void fft_group(const float2 *__restrict__ twf, const float2 *__restrict__ twc, int g, LD ld, ST st) {
    constexpr int N = 1 << L;
    constexpr int B0 = N >> S0;
    constexpr int STRIDE = B0 >> Q;
    constexpr int M = 1 << Q;
    constexpr int Q0 = FftPlan<L>::Q0;
    const int blk = g / STRIDE, r = g % STRIDE;
    const int base = blk * B0 + r;
    float2 v[M];
#pragma unroll
    for (int m = 0; m < M; m++) v[m] = ld(base + m * STRIDE);
#pragma unroll
    for (int i = 0; i < Q; i++) {
        const int span = 1 << (Q - 1 - i);
        const int s = S0 + i;
#pragma unroll
        for (int mm = 0; mm < span; mm++) {
            const int j = r + mm * STRIDE;
            if (s == L - 1) {
#pragma unroll
                for (int hi = 0; hi < M / (2 * span); hi++) {
                    int m = hi * 2 * span + mm;
                    bfly_w1<INV>(v[m], v[m + span]);
                }
            } else if (s == L - 2 && STRIDE == 1) {
                // j is 0 or 1 and known at compile time here (r == 0, mm in {0,1})
#pragma unroll
                for (int hi = 0; hi < M / (2 * span); hi++) {
                    int m = hi * 2 * span + mm;
                    if (mm == 0) bfly_w1<INV>(v[m], v[m + span]);
                    else bfly_wq<INV>(v[m], v[m + span]);
                }
            } else {
                float2 w = (s < Q0) ? (TWG ? __ldg(twf + tw_pad(j << s)) : twf[tw_pad(j << s)]) : twc[twc_off<L>(s) + j];
#pragma unroll
                for (int hi = 0; hi < M / (2 * span); hi++) {
                    int m = hi * 2 * span + mm;
                    bfly<INV>(v[m], v[m + span], w);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < M; m++) st(base + m * STRIDE, m, v[m]);
}

__device__ __forceinline__ int bitrev_n(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }

// Copy the host-built twiddle image (fft_tw_elems<L>() float2) into shared memory.
template <int L>
__device__ __forceinline__ void fft_load_twiddles(float2 *tw_s, const float2 *__restrict__ tw_g) {
    for (int i = threadIdx.x; i < fft_tw_elems<L>(); i += blockDim.x) tw_s[i] = tw_g[i];
}

// In-shared-memory transform of `data` (padded layout).  All threads of the CTA must call.
// Natural-order input at data[fft_pad(p)]; on return data[fft_pad(p)] = X[bitrev(p)].
// The first pass reads through `ld0(p)` so callers can fuse their own producer; the last
// pass hands (k = natural frequency index, value) to `out(k, value)` instead of storing when
// FUSE_OUT is set.  The last pass maps thread t to group bitrev(t) so that, for fixed m,
// consecutive threads own consecutive k.
template <int L, bool INV, bool FUSE_OUT, bool TWG, class LD0, class OUT>
__device__ __forceinline__ void fft_smem2(float2 *data, const float2 *twf, const float2 *twc, LD0 ld0, OUT out) {
    using PL = FftPlan<L>;
    constexpr int N = 1 << L;
    auto lds = [&](int p) { return data[fft_pad<L>(p)]; };
    auto sts = [&](int p, int, float2 v) { data[fft_pad<L>(p)] = v; };
    // pass 0
    for (int g = threadIdx.x; g < (N >> PL::Q0); g += blockDim.x)
        fft_group<L, 0, PL::Q0, INV, TWG>(twf, twc, g, ld0, sts);
    __syncthreads();
    if constexpr (PL::P == 2) {
        constexpr int Q = PL::Q1, S0 = PL::Q0;
        for (int t = threadIdx.x; t < (N >> Q); t += blockDim.x) {
            int g = bitrev_n(t, L - Q);
            if (FUSE_OUT) {
                fft_group<L, S0, Q, INV, TWG>(twf, twc, g, lds, [&](int, int m, float2 v) {
                    out((bitrev_n(m, Q) << (L - Q)) | t, v);
                });
            } else {
                fft_group<L, S0, Q, INV, TWG>(twf, twc, g, lds, sts);
            }
        }
    } else {
        for (int g = threadIdx.x; g < (N >> PL::Q1); g += blockDim.x)
            fft_group<L, PL::Q0, PL::Q1, INV, TWG>(twf, twc, g, lds, sts);
        __syncthreads();
        constexpr int Q = PL::Q2 > 0 ? PL::Q2 : 1, S0 = PL::Q0 + PL::Q1;
        for (int t = threadIdx.x; t < (N >> Q); t += blockDim.x) {
            int g = bitrev_n(t, L - Q);
            if (FUSE_OUT) {
                fft_group<L, S0, Q, INV, TWG>(twf, twc, g, lds, [&](int, int m, float2 v) {
                    out((bitrev_n(m, Q) << (L - Q)) | t, v);
                });
            } else {
                fft_group<L, S0, Q, INV, TWG>(twf, twc, g, lds, sts);
            }
        }
    }
    __syncthreads();
}
// the whole twiddle image (fft_load_twiddles) in shared memory
template <int L, bool INV, bool FUSE_OUT, class LD0, class OUT>
__device__ __forceinline__ void fft_smem(float2 *data, const float2 *tw, LD0 ld0, OUT out) {
    fft_smem2<L, INV, FUSE_OUT, false>(data, tw, tw + fft_twfull_elems<L>(), ld0, out);
}

// Read X[k] after a non-fused fft_smem.
template <int L>
__device__ __forceinline__ float2 fft_result(const float2 *data, int k) {
    return data[fft_pad<L>(bitrev_n(k, L))];
}

// ---------------------------------------------------------------- block reductions
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// argmax with "first index wins on ties" (strict > scan order of the reference loops)
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax argmax_pick(ArgMax a, ArgMax b) {
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
        a = argmax_pick(a, b);
    }
    return a;
}
// scratch: >= 32 ArgMax in shared memory.  Result valid in every thread.
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, ArgMax *scratch) {
    a = warp_argmax(a);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) scratch[w] = a;
    __syncthreads();
    ArgMax r = (l < nw) ? scratch[l] : ArgMax{-1.0f, 0x7fffffff};
    r = warp_argmax(r);
    return r;
}

}  // namespace ir

// ---------------------------------------------------------------- TMA / mbarrier (PTX)
// 1-D bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion, used to stream the
// magnitude rows into the detector's state machine and IQ rows into the FIR CTAs.
namespace ir {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared, `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Ampere-style 8-byte asynchronous copy (LDGSTS) for sources that are only 8-byte aligned
__device__ __forceinline__ void cp_async_8(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace ir
