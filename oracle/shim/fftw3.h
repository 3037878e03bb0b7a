/*
 * Minimal stand-in for <fftw3.h> -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * libfftw3f is not installed in this image, so the
 * reference sources are compiled unmodified against this 7-symbol contract
 *  and linked with oracle/shim/fftw_shim.c.
 * Nothing under iridium-sniffer_b200/ includes or links this.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H

#include <complex.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float _Complex fftwf_complex;
typedef struct shim_plan_s *fftwf_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void      *fftwf_alloc_complex(size_t n);
void       fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
                             int sign, unsigned flags);
void       fftwf_execute(const fftwf_plan p);
void       fftwf_destroy_plan(fftwf_plan p);
int        fftwf_import_wisdom_from_filename(const char *filename);
int        fftwf_export_wisdom_to_filename(const char *filename);

#ifdef __cplusplus
}
#endif
#endif
