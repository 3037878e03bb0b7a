// blocks_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's multi-GPU host driver
// (iridium-sniffer_b200/csrc/blocks.cu: ir_multi_* -- a pipeline and a host thread per device, time blocks dealt
// round-robin, merge, RAW text) for a machine without a GPU by putting stand-ins under it: the ir_pipeline_* calls
// it makes are answered here by the CPU oracle (oracle/libir_oracle.so) on "devices" that are just numbers.  What is
// exercised is the driver's own logic -- planning, threads, origins, result copies, error propagation, merge, text;
// ir_plan_blocks / ir_merge_blocks / ir_format_raw / the detector parameters are the library's own code either way.
// The library never does this: libiridium_b200.so has no CPU path, and its ir_pipeline_* are the CUDA ones.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../iridium-sniffer_b200/csrc/blocks.cu"
#include "../iridium-sniffer_b200/csrc/frame_classify.cuh"
#include "../oracle/ir_oracle.h"

static_assert(sizeof(orc_result) == sizeof(ir_frame_t), "orc_result mirrors ir_frame_t field for field");

struct ir_pipeline {
    ir_config_t cfg;
    uint64_t origin = 0;
    orc_run run;
    bool has_run = false;
    std::vector<float> llr;
    int runs = 0;
};

static int g_fake_devices = 4;
static int g_fail_device = -1;        // this "device" fails its second run (error propagation out of a worker thread)

extern "C" void shim_set_devices(int n, int fail_device) { g_fake_devices = n; g_fail_device = fail_device; }

extern "C" ir_pipeline_t *ir_pipeline_create(const ir_config_t *cfg) {
    if (!cfg || cfg->device < 0 || cfg->device >= g_fake_devices) { ir::set_last_error("shim: no such device"); return nullptr; }
    ir_pipeline *p = new ir_pipeline();
    p->cfg = *cfg;
    memset(&p->run, 0, sizeof(p->run));
    return p;
}
extern "C" void ir_pipeline_destroy(ir_pipeline_t *p) {
    if (!p) return;
    if (p->has_run) orc_run_free(&p->run);
    delete p;
}
extern "C" int ir_pipeline_set_origin(ir_pipeline_t *p, uint64_t o) { p->origin = o; return 0; }
extern "C" int ir_pipeline_set_start_time(ir_pipeline_t *p, uint64_t t) { p->cfg.start_time_ns = t; return 0; }
extern "C" int ir_pipeline_run_host(ir_pipeline_t *p, const void *iq, size_t n, int fmt) {
    if (fmt != IR_FMT_CF32) { ir::set_last_error("shim: cf32 only"); return -1; }
    if (p->cfg.device == g_fail_device && ++p->runs >= 2) { ir::set_last_error("shim: device fell off the bus"); return -1; }
    if (p->has_run) orc_run_free(&p->run);
    const uint64_t t = p->cfg.start_time_ns + (uint64_t)((double)p->origin / (double)p->cfg.sample_rate * 1e9);
    const int rc = orc_run_recording((const orc_cf32 *)iq, n, p->cfg.center_frequency, p->cfg.sample_rate,
                                     p->cfg.threshold_db > 0 ? p->cfg.threshold_db : 16.0f,
                                     p->cfg.feed_block > 0 ? (size_t)p->cfg.feed_block : 32768, t, p->cfg.use_gardner, &p->run);
    p->has_run = true;
    p->llr.assign(p->run.bits_len + 1, 0.0f);
    return rc;
}
extern "C" int ir_pipeline_results(ir_pipeline_t *p, ir_results_t *out) {
    memset(out, 0, sizeof(*out));
    out->n_frames = p->run.n_results;
    out->frames = (const ir_frame_t *)p->run.results;
    out->bits = p->run.bits;
    out->llr = p->llr.data();
    out->n_bits_total = p->run.bits_len;
    out->kernel_launches = 1;
    return 0;
}

// classification stand-in: the product's own arithmetic (csrc/frame_classify.cuh, pinned to the reference by
// tests/test_frame_classify_host.py) compiled for the host, over the oracle's bits (no LLRs: no Chase search)
extern "C" long ir_pipeline_classify(ir_pipeline_t *p, ir_frame_class_t *out, size_t cap) {
    static const ir::FcTables &tab = *[] { auto *t = new ir::FcTables(); ir::fc_build_tables(*t); return t; }();   // thread-safe
    if (p->run.n_results > cap) { ir::set_last_error("shim: class array too small"); return -1; }
    for (size_t i = 0; i < p->run.n_results; i++) {
        const orc_result &r = p->run.results[i];
        ir::fc_classify(tab, p->run.bits + r.bits_offset, nullptr, r.n_bits, r.direction, &out[i]);
    }
    return (long)p->run.n_results;
}

// ir_pipeline_format_parsed_all over the stand-in (same rule as the library's: IDA line where ida_decode() accepted,
// RAW line otherwise; main.c:328-331) -- only so that the GPU cases' own code can be dry-run on the CPU
extern "C" long ir_pipeline_format_parsed_all(ir_pipeline_t *p, const char *file_info, uint64_t t0, const ir_frame_class_t *cls,
                                              size_t n_cls, char *dst, size_t cap) {
    const size_t n = p->run.n_results;
    if (!dst) return (long)(n * 1024 + p->run.bits_len + 64);
    std::vector<ir_frame_class_t> own;
    if (!cls) { own.resize(n + 1); if (ir_pipeline_classify(p, own.data(), n) < 0) return -1; cls = own.data(); n_cls = n; }
    if (n_cls != n) return -1;
    const ir_frame_t *fr = (const ir_frame_t *)p->run.results;
    if (n && t0 == 0) t0 = (fr[0].timestamp / 1000000000ULL) * 1000000000ULL;
    size_t pos = 0;
    for (size_t i = 0; i < n; i++) {
        const int k = cls[i].ida_ok ? ir_format_ida(dst + pos, cap - pos, t0, &fr[i], &cls[i])
                                    : ir_format_raw(dst + pos, cap - pos, file_info, t0, &fr[i], p->run.bits + fr[i].bits_offset);
        if (k < 0) return -1;
        pos += (size_t)k;
    }
    return (long)pos;
}
