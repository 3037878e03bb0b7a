// k_classify.cu -- frame classification on the device: frame_decode() (frame_decode.c:414-598) and
// ida_decode() (ida_decode.c:543-662) for every demodulated frame of a run in one launch, reading the
// bits and LLRs where k_demod left them.  The arithmetic lives in frame_classify.cuh (shared with the
// host-compiled parity test); this file is the mapping onto the machine.
//
// Mapping: one warp per frame.  A frame is at most 2*IR_MAX_SYMS one-byte bits plus as many float LLRs
// (<= 4.8 KB); its warp copies them into shared memory with coalesced loads, then ALL 32 lanes run the
// classifiers in lock step out of shared memory -- same data, same branches, so SIMT issues each instruction
// once, exactly as for a lone lane -- and part ways only in the Chase search of a code word that does not
// decode outright: lane L tries subset L of the five least reliable bits, `VOTE` + find-first-set picks
// the reference's "first that decodes in counting order" (frame_classify.cuh: fc_decode31).  The rest of a
// frame is a few hundred dependent integer operations on <= 16 code words, too little to be worth spreading
// further; thousands of frames per run keep all SMs busy with one warp each.  The syndrome tables (25 KB)
// stay in global memory: every warp reads a handful of entries, L2-resident after the first frames.  The
// result is assembled in the lanes' (identical) local copies and stored cooperatively, one word per lane.
// Bound: latency of the dependent chain; bytes moved per frame = 5 * n_bits in, 504 out.
#include <mutex>

#include "frame_classify.cuh"
#include "ir_internal.h"

namespace ir {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxBits = 2 * IR_MAX_SYMS;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_classify_frames(const FcTables *__restrict__ T, const FrameSrc *__restrict__ src, int n_frames,
                  ir_frame_class_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t s_bits[kWarpsPerBlock][kMaxBits];
    __shared__ __align__(16) float s_llr[kWarpsPerBlock][kMaxBits];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kWarpsPerBlock + warp;
    if (f >= n_frames) return;
    const FrameSrc fs = src[f];
    int n = fs.n_bits;
    if (n < 0) n = 0;
    const uint8_t *bits = fs.bits;
    const float *llr = fs.llr;
    if (n <= kMaxBits) {
        for (int i = lane; i < n; i += 32) s_bits[warp][i] = fs.bits[i];
        if (fs.llr)
            for (int i = lane; i < n; i += 32) s_llr[warp][i] = fs.llr[i];
        __syncwarp();
        bits = s_bits[warp];
        if (fs.llr) llr = s_llr[warp];
    }
    // (longer than anything k_demod writes -- a caller's own frame through ir_classify_frames: the same
    // arithmetic straight from global memory)
    ir_frame_class_t r;
    fc_classify(*T, bits, llr, n, fs.direction, &r);
    static_assert(sizeof(ir_frame_class_t) % 4 == 0, "result is stored word by word");
    const uint32_t *rw = reinterpret_cast<const uint32_t *>(&r);
    uint32_t *ow = reinterpret_cast<uint32_t *>(&out[f]);
    for (int i = lane; i < (int)(sizeof(ir_frame_class_t) / 4); i += 32) ow[i] = rw[i];
}

namespace {

std::mutex g_tab_mu;
FcTables *g_host_tab = nullptr;
void *g_dev_tab[64] = {nullptr};

}  // namespace

cudaError_t classify_tables(int device, const void **d_tables) {
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (!g_host_tab) {
        g_host_tab = new FcTables();
        fc_build_tables(*g_host_tab);
    }
    if (!g_dev_tab[device]) {
        void *d = nullptr;
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&d, sizeof(FcTables));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, g_host_tab, sizeof(FcTables), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d); return e; }
        g_dev_tab[device] = d;
    }
    *d_tables = g_dev_tab[device];
    return cudaSuccess;
}

cudaError_t launch_classify(const void *d_tables, const FrameSrc *src, int n_frames, ir_frame_class_t *out,
                            cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    const int blocks = (n_frames + kWarpsPerBlock - 1) / kWarpsPerBlock;
    k_classify_frames<<<blocks, kWarpsPerBlock * 32, 0, st>>>((const FcTables *)d_tables, src, n_frames, out);
    return cudaGetLastError();
}

}  // namespace ir
