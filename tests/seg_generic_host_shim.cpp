// seg_generic_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's generic segment walker
// (iridium-sniffer_b200/csrc/seg_generic.cuh: the code k_seg_walk hands a segment with more than 32 bursts to) for the
// HOST, so that tests/test_seg_scan_model.py can run it in place of the numpy walker against the CPU oracle:
// with one lane, and with several lanes played by threads that meet at a barrier wherever the warp's lanes
// exchange something (the device's policy is SegLanesWarp in k_detect_seg.cu; same walker code).
#include <atomic>
#include <thread>
#include <vector>

#include "../iridium-sniffer_b200/csrc/seg_generic.cuh"

extern "C" int segg_sizeof_burst(void) { return (int)sizeof(ir::SegBurst); }
extern "C" int segg_sizeof_gone(void) { return (int)sizeof(ir::GoneBurst); }
extern "C" int segg_seg_len(void) { return IR_SEG_LEN; }

namespace {

template <int LN>
struct LanesShared {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    unsigned long long slot[LN];
};

template <int LN>
struct SegLanesThreads {
    static constexpr int L = LN;
    LanesShared<LN> *sh;
    int me;
    int lane() const { return me; }
    void sync() const {
        const int ph = sh->phase.load(std::memory_order_acquire);
        if (sh->arrived.fetch_add(1, std::memory_order_acq_rel) == LN - 1) {
            sh->arrived.store(0, std::memory_order_relaxed);
            sh->phase.store(ph + 1, std::memory_order_release);
        } else {
            while (sh->phase.load(std::memory_order_acquire) == ph) std::this_thread::yield();
        }
    }
    template <class F>
    unsigned long long all(unsigned long long v, F f) const {       // f folds every lane's value
        sh->slot[me] = v;
        sync();
        unsigned long long r = sh->slot[0];
        for (int i = 1; i < LN; i++) r = f(r, sh->slot[i]);
        sync();
        return r;
    }
    bool any(bool p) const { return all(p, [](unsigned long long a, unsigned long long b) { return a | b; }) != 0; }
    int sum(int v) const { return (int)all((unsigned long long)v, [](unsigned long long a, unsigned long long b) { return a + b; }); }
    int max(int v) const { return (int)all((unsigned long long)v, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }); }
    unsigned long long max64(unsigned long long v) const { return all(v, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }); }
    uint32_t ballot(bool p) const {
        return (uint32_t)all(p ? 1ull << me : 0ull, [](unsigned long long a, unsigned long long b) { return a | b; });
    }
    int excl_scan(int v, int &total) const {
        sh->slot[me] = (unsigned long long)v;
        sync();
        int before = 0, t = 0;
        for (int i = 0; i < LN; i++) { if (i < me) before += (int)sh->slot[i]; t += (int)sh->slot[i]; }
        sync();
        total = t;
        return before;
    }
    void and_word(uint32_t *p, uint32_t m) const { __atomic_fetch_and(p, m, __ATOMIC_RELAXED); }
    const uint32_t *stage(const uint32_t *row, int) const { return row; }
    void tick(int) const {}
};

template <int LN>
int walk_threads(const ir::SegGenArgs &a, ir::SegBurst *work, int n_start, int cap, ir::GoneBurst *gl, int gl_cap, uint32_t *fv,
                 ir::SegGenOut &out) {
    LanesShared<LN> sh;
    int rc[LN];
    std::vector<std::thread> th;
    for (int i = 0; i < LN; i++)
        th.emplace_back([&, i]() { rc[i] = ir::seg_walk_generic_t(SegLanesThreads<LN>{&sh, i}, a, work, n_start, cap, gl, gl_cap, fv, out); });
    for (auto &t : th) t.join();
    for (int i = 1; i < LN; i++)
        if (rc[i] != rc[0]) return -1000 - i;                        // every lane must come back with the same answer
    return rc[0];
}

}  // namespace

extern "C" int segg_walk(int N, int half_bw, int max_bursts, int pre_len, int post_len, int max_burst_len, float thr,
                         int seg, int f0, int n_frames, long long index0, int sq_start, const uint32_t *xu, const float *mag,
                         const float *snap, const int *fslot, const uint32_t *valid, ir::SegBurst *work, int n_start, int cap,
                         ir::GoneBurst *gl, int gl_cap, int *counts /* n_end, n_gone, n_create */, uint32_t *qbits, int lanes) {
    ir::SegGenArgs a;
    a.N = N; a.half_bw = half_bw; a.max_bursts = max_bursts; a.pre_len = pre_len; a.post_len = post_len;
    a.max_burst_len = max_burst_len; a.thr = thr; a.seg = seg; a.f0 = f0; a.n_frames = n_frames; a.index0 = index0;
    a.sq_start = sq_start; a.xu = xu; a.mag = mag; a.snap = snap; a.fslot = fslot; a.valid = valid;
    std::vector<unsigned long long> keys(8192);
    a.keys = keys.data(); a.pcap = 8192;
    std::vector<uint32_t> fv((size_t)N / 32);
    ir::SegGenOut out;
    int rc;
    if (lanes == 4) rc = walk_threads<4>(a, work, n_start, cap, gl, gl_cap, fv.data(), out);
    else if (lanes == 3) rc = walk_threads<3>(a, work, n_start, cap, gl, gl_cap, fv.data(), out);
    else rc = ir::seg_walk_generic(a, work, n_start, cap, gl, gl_cap, fv.data(), out);
    if (rc == 0) {
        counts[0] = out.n_end; counts[1] = out.n_gone; counts[2] = out.n_create;
        for (int w = 0; w < (IR_SEG_LEN + 31) / 32; w++) qbits[w] = out.qbits[w];
    }
    return rc;
}
