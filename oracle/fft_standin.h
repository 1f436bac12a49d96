/*
 * fft_standin.h -- labelled STAND-IN for FFTW3f (test infrastructure, see dab_oracle.h).
 *
 * The reference's only FFT is libfftw3f via fftwf_plan_dft_1d(..., FFTW_ESTIMATE), in place, sizes
 * 256/512/1024/2048 (src/various/fft.cpp:36-41, 91-97).  FFTW is not vendored, not pinned and not installed
 * here, so the oracle (both the port and the compiled-reference build) uses this float32 Stockham radix-4
 * FFT with double-precision-generated twiddles.  It is unnormalised in both directions like FFTW.
 */
#ifndef FFT_STANDIN_H
#define FFT_STANDIN_H
#ifdef __cplusplus
extern "C" {
#endif
/* in-place transform of n (power of two, 2..4096) interleaved complex floats; sign = -1 forward, +1 backward */
int fft_standin_exec (float *v, int n, int sign);
#ifdef __cplusplus
}
#endif
#endif
