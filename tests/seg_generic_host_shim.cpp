// seg_generic_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's generic segment walker
// (iridium-sniffer_b200/csrc/seg_generic.cuh: the code k_seg_walk hands a segment with more than 32 bursts to) for the
// HOST, so that tests/test_seg_scan_model.py can run it in place of the numpy walker against the CPU oracle.
#include <vector>

#include "../iridium-sniffer_b200/csrc/seg_generic.cuh"

extern "C" int segg_sizeof_burst(void) { return (int)sizeof(ir::SegBurst); }
extern "C" int segg_sizeof_gone(void) { return (int)sizeof(ir::GoneBurst); }
extern "C" int segg_seg_len(void) { return IR_SEG_LEN; }

extern "C" int segg_walk(int N, int half_bw, int max_bursts, int pre_len, int post_len, int max_burst_len, float thr,
                         int seg, int f0, int n_frames, long long index0, int sq_start, const uint32_t *xu, const float *mag,
                         const float *snap, const int *fslot, const uint32_t *valid, ir::SegBurst *work, int n_start, int cap,
                         ir::GoneBurst *gl, int gl_cap, int *counts /* n_end, n_gone, n_create */, uint32_t *qbits) {
    ir::SegGenArgs a;
    a.N = N; a.half_bw = half_bw; a.max_bursts = max_bursts; a.pre_len = pre_len; a.post_len = post_len;
    a.max_burst_len = max_burst_len; a.thr = thr; a.seg = seg; a.f0 = f0; a.n_frames = n_frames; a.index0 = index0;
    a.sq_start = sq_start; a.xu = xu; a.mag = mag; a.snap = snap; a.fslot = fslot; a.valid = valid;
    std::vector<float> prel(8192);
    std::vector<int> pbin(8192);
    a.prel = prel.data(); a.pbin = pbin.data(); a.pcap = 8192;
    std::vector<uint32_t> fv((size_t)N / 32);
    ir::SegGenOut out;
    const int rc = ir::seg_walk_generic(a, work, n_start, cap, gl, gl_cap, fv.data(), out);
    if (rc == 0) {
        counts[0] = out.n_end; counts[1] = out.n_gone; counts[2] = out.n_create;
        for (int w = 0; w < (IR_SEG_LEN + 31) / 32; w++) qbits[w] = out.qbits[w];
    }
    return rc;
}
