// refapi_frames_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the library's reference-named frame entry points
// (iridium-sniffer_b200/csrc/refapi_frames.cu: frame_decode() / ida_decode() per frame) for a machine without a GPU:
// their one device call, ir_classify_frames, is answered here by the same arithmetic compiled for the host
// (csrc/frame_classify.cuh).  Lets tests/test_gpu_case_dry_runs.py run the GPU case's own code.  The library never
// does this: its ir_classify_frames launches k_classify_frames and fails without a device.
#include <stddef.h>
#include <stdint.h>
static inline int cudaGetDevice(int *d) { *d = 0; return 0; }
#include "../iridium-sniffer_b200/csrc/refapi_frames.cu"

extern "C" int ir_classify_frames(int, const ir_frame_t *frames, size_t n_frames, const uint8_t *bits, const float *llr,
                                  size_t, ir_frame_class_t *out) {
    static const ir::FcTables &tab = *[] { auto *t = new ir::FcTables(); ir::fc_build_tables(*t); return t; }();
    for (size_t i = 0; i < n_frames; i++)
        ir::fc_classify(tab, bits + frames[i].bits_offset, llr ? llr + frames[i].bits_offset : nullptr, frames[i].n_bits,
                        frames[i].direction, &out[i]);
    return 0;
}
