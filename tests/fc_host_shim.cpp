// fc_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the product's frame classification arithmetic
// (iridium-sniffer_b200/csrc/frame_classify.cuh, the very code k_classify.cu runs one warp per frame) for the
// HOST, so that tests without a GPU can hold it against the reference's frame_decode() / ida_decode() and the
// oracle.  The library never does this: libiridium_b200.so only runs it inside k_classify_frames.
#include "../iridium-sniffer_b200/csrc/frame_classify.cuh"

static ir::FcTables g_tab;
static bool g_ready;

extern "C" int fc_host_classify(const uint8_t *bits, const float *llr, int n_bits, int direction, ir_frame_class_t *out) {
    if (!g_ready) { ir::fc_build_tables(g_tab); g_ready = true; }
    ir::fc_classify(g_tab, bits, llr, n_bits, direction, out);
    return 0;
}
extern "C" int fc_host_sizeof(void) { return (int)sizeof(ir_frame_class_t); }
