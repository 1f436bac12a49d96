/*
 * fft_standin.c -- float32 Stockham autosort FFT (radix 4 with one radix-2 pass for odd log2 n).
 * Stand-in for FFTW3f; see fft_standin.h.  Test infrastructure only.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "fft_standin.h"

#define FS_MAXN 4096

typedef struct { float re, im; } cf;

/* one twiddle table per size, built on first use: tw[k] = exp(-2*pi*i*k/n), computed in double */
static cf *tw_tab[13];

static const cf *twiddles (int n, int lg) {
	if (tw_tab[lg] == NULL) {
		cf *t = (cf *) malloc (sizeof (cf) * n);
		for (int k = 0; k < n; k++) {
			double a = -2.0 * M_PI * (double) k / (double) n;
			t[k].re = (float) cos (a);
			t[k].im = (float) sin (a);
		}
		if (!__sync_bool_compare_and_swap (&tw_tab[lg], NULL, t))
			free (t);     /* another thread published the same table first */
	}
	return tw_tab[lg];
}

/* forward transform, x -> result in x, y is scratch */
static void stockham_forward (int N, cf *x0, cf *y0, const cf *tw) {
	cf *x = x0, *y = y0;
	int n = N, s = 1;
	while (n >= 4) {
		const int m = n / 4;
		const int tstep = N / n;
		for (int p = 0; p < m; p++) {
			const cf w1 = tw[p * tstep], w2 = tw[2 * p * tstep], w3 = tw[3 * p * tstep];
			for (int q = 0; q < s; q++) {
				const cf a = x[q + s * p], b = x[q + s * (p + m)];
				const cf c = x[q + s * (p + 2 * m)], d = x[q + s * (p + 3 * m)];
				const cf apc = { a.re + c.re, a.im + c.im }, amc = { a.re - c.re, a.im - c.im };
				const cf bpd = { b.re + d.re, b.im + d.im };
				/* j*(b-d) */
				const cf jbmd = { -(b.im - d.im), b.re - d.re };
				cf t;
				y[q + s * (4 * p)].re = apc.re + bpd.re;
				y[q + s * (4 * p)].im = apc.im + bpd.im;
				t.re = amc.re - jbmd.re; t.im = amc.im - jbmd.im;
				y[q + s * (4 * p + 1)].re = t.re * w1.re - t.im * w1.im;
				y[q + s * (4 * p + 1)].im = t.re * w1.im + t.im * w1.re;
				t.re = apc.re - bpd.re; t.im = apc.im - bpd.im;
				y[q + s * (4 * p + 2)].re = t.re * w2.re - t.im * w2.im;
				y[q + s * (4 * p + 2)].im = t.re * w2.im + t.im * w2.re;
				t.re = amc.re + jbmd.re; t.im = amc.im + jbmd.im;
				y[q + s * (4 * p + 3)].re = t.re * w3.re - t.im * w3.im;
				y[q + s * (4 * p + 3)].im = t.re * w3.im + t.im * w3.re;
			}
		}
		n /= 4; s *= 4;
		cf *t = x; x = y; y = t;
	}
	if (n == 2) {
		for (int q = 0; q < s; q++) {
			const cf a = x[q], b = x[q + s];
			y[q].re = a.re + b.re;      y[q].im = a.im + b.im;
			y[q + s].re = a.re - b.re;  y[q + s].im = a.im - b.im;
		}
		cf *t = x; x = y; y = t;
	}
	if (x != x0)
		memcpy (x0, x, sizeof (cf) * N);
}

int fft_standin_exec (float *v, int n, int sign) {
	int lg = 0;
	while ((1 << lg) < n) lg++;
	if (n < 2 || n > FS_MAXN || (1 << lg) != n)
		return -1;
	cf scratch[FS_MAXN];
	cf *x = (cf *) v;
	const cf *tw = twiddles (n, lg);
	if (sign > 0)        /* backward = conj (forward (conj x)); conjugation is exact in float */
		for (int i = 0; i < n; i++) x[i].im = -x[i].im;
	stockham_forward (n, x, scratch, tw);
	if (sign > 0)
		for (int i = 0; i < n; i++) x[i].im = -x[i].im;
	return 0;
}
