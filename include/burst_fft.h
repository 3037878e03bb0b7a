/*
 * burst_fft.h -- the reference's existing accelerator plug-in ABI, implemented by
 * libiridium_b200.so with the hand-written sm_100a window -> FFT -> fftshift -> |X|^2 kernel.
 *
 * Identical to opencl/burst_fft.h:35-47 of the reference (which is what burst_detect.c:309,659
 * call under -DUSE_GPU), so the unmodified reference detector can be built against this
 * library instead of opencl/burst_fft.c or vulkan/burst_fft.c:
 *
 *   gpu_burst_fft_create (fft_size, batch_size, window)  -> context or NULL
 *   gpu_burst_fft_process(g, input, output, batch_count) -> 0 / -1
 *        input : batch_count * fft_size interleaved (re, im) floats   (host memory)
 *        output: batch_count * fft_size floats, fftshifted |X|^2      (host memory)
 *   gpu_burst_fft_destroy(g)
 *
 * fft_size must be a power of two in 1024..16384.  NULL / -1 on any CUDA error; the caller's
 * own fallback (burst_detect.c:316-318) then decides -- this library itself never computes
 * on the CPU.
 */
#ifndef IR_BURST_FFT_H
#define IR_BURST_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpu_burst_fft gpu_burst_fft_t;

gpu_burst_fft_t *gpu_burst_fft_create(int fft_size, int batch_size, const float *window);
void gpu_burst_fft_destroy(gpu_burst_fft_t *g);
int gpu_burst_fft_process(gpu_burst_fft_t *g, const float *input, float *output, int batch_count);

#ifdef __cplusplus
}
#endif
#endif
