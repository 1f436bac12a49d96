/* oracle/ref_shim/fftw3.h -- the five FFTW3f entry points the reference's fft.cpp uses
 * (src/various/fft.cpp:36-41, 54, 91-97), implemented by ref_provider.cpp on top of the labelled
 * stand-in FFT (oracle/fft_standin.c).  FFTW itself is not available in this image. */
#ifndef ORC_SHIM_FFTW3_H
#define ORC_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct orc_fftw_plan_s *fftwf_plan;
#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
void      *fftwf_malloc (size_t n);
void       fftwf_free (void *p);
fftwf_plan fftwf_plan_dft_1d (int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void       fftwf_execute (const fftwf_plan p);
void       fftwf_destroy_plan (fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
