/*
 * orc_port.c -- plain-C restatement of the reference's Tier-A/B primitives (provider "port" of
 * orc_kernels.h -> oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY (see dab_oracle.h).
 * Every function cites the reference file:line (relative to /root/reference) it follows; the restatement
 * is pinned against the compiled reference (oracle/_ref/libdabref.so) by tests/test_oracle_vs_ref.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "orc_kernels.h"
#include "fft_standin.h"

const char *k_kind (void) { return "port"; }

/* ------------------------------------------------------------------------------------------------
 * frequency (de)interleaver table -- src/ofdm/mapper.cpp:33-55 (createMapper), :57-106 (per-mode V1/lwb)
 * ---------------------------------------------------------------------------------------------- */
int k_perm_table (const orc_params *p, int16_t *out) {
	int V1, lwb;
	switch (p -> dabMode) {             /* mapper.cpp:82-105; "default" = Mode I */
	   case 2:  V1 = 127; lwb = 64;  break;
	   case 3:  V1 = 63;  lwb = 32;  break;
	   case 4:  V1 = 255; lwb = 128; break;
	   default: V1 = 511; lwb = 256; break;
	}
	const int T_u = p -> T_u, upb = lwb + p -> K;
	int n = 0;
	int16_t pi = 0;                      /* mapper.cpp:39-41: Pi(0)=0, Pi(i)=(13*Pi(i-1)+V1) mod T_u */
	for (int i = 0; i < T_u; i ++) {
		if (i > 0) pi = (int16_t) ((13 * pi + V1) % T_u);
		if (pi == T_u / 2) continue;      /* :43-44 the DC carrier is skipped   */
		if (pi < lwb || pi > upb) continue; /* :45-47                          */
		out [n ++] = (int16_t) (pi - T_u / 2); /* :51                            */
	}
	return n == p -> K ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * PRS phases -- src/ofdm/phasetable.cpp:26-121 (tables), :145-156 (h), :172-185 (get_Phi)
 * The tables are rows {kmin, kmax, i, n} with kmax = kmin+31 and contiguous kmin, so only (i,n) per
 * 32-carrier block is stored: first the negative blocks from -K/2 upward, then the positive from +1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int8_t i, n; } in_t;
static const in_t m1_neg [24] = { {0,1},{1,2},{2,0},{3,1},{0,3},{1,2},{2,2},{3,3},{0,2},{1,1},{2,2},{3,3},
                                  {0,1},{1,2},{2,3},{3,3},{0,2},{1,2},{2,2},{3,1},{0,1},{1,3},{2,1},{3,2} };
static const in_t m1_pos [24] = { {0,3},{3,1},{2,1},{1,1},{0,2},{3,2},{2,1},{1,0},{0,2},{3,2},{2,3},{1,3},
                                  {0,0},{3,2},{2,1},{1,3},{0,3},{3,3},{2,3},{1,0},{0,3},{3,0},{2,1},{1,1} };
static const in_t m2_neg [6]  = { {0,2},{1,3},{2,2},{3,2},{0,1},{1,2} };
static const in_t m2_pos [6]  = { {2,0},{1,2},{0,2},{3,1},{2,0},{1,3} };
static const in_t m4_neg [12] = { {0,0},{1,1},{2,1},{3,2},{0,2},{1,2},{2,0},{3,3},{0,3},{1,1},{2,3},{3,2} };
static const in_t m4_pos [12] = { {0,0},{3,1},{2,0},{1,2},{0,0},{3,1},{2,2},{1,2},{0,2},{3,1},{2,3},{1,0} };
static const int8_t h_rows [4][16] = {            /* phasetable.cpp:145-156 (second halves repeat) */
	{0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1},
	{0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0},
	{0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3},
	{0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2} };

float k_phi (int mode, int k) {
	const in_t *neg, *pos; int nblk;
	switch (mode) {                      /* phasetable.cpp:123-139: Mode III falls through to Mode I */
	   case 2:  neg = m2_neg; pos = m2_pos; nblk = 6;  break;
	   case 4:  neg = m4_neg; pos = m4_pos; nblk = 12; break;
	   default: neg = m1_neg; pos = m1_pos; nblk = 24; break;
	}
	int blk, off; const in_t *e;
	if (k < 0) { int d = k + 32 * nblk; if (d < 0) goto help; blk = d / 32; off = d % 32; e = &neg [blk]; }
	else if (k > 0) { int d = k - 1; blk = d / 32; off = d % 32; if (blk >= nblk) goto help; e = &pos [blk]; }
	else goto help;
	/* phasetable.cpp:180: return M_PI / 2 * (h_table (i, k - k_prime) + n) -- double product, float result */
	return (float) (M_PI / 2 * (h_rows [e -> i][off & 15] + e -> n));
help:
	return 0;                            /* phasetable.cpp:183-184 */
}

/* ------------------------------------------------------------------------------------------------
 * puncturing vectors -- src/backend/protTables.cpp:27-54.  PI_n (n = 1..24) has 8+n ones in 32: every
 * 4-bit group starts from 1000; groups receive a second / third / fourth one in the order below.
 * ---------------------------------------------------------------------------------------------- */
int k_pcode (int n, int8_t *out) {
	static const int8_t fill_order [8] = { 0, 4, 2, 6, 1, 5, 3, 7 };
	if (n < 1 || n > 24) return -1;
	int ones [8];
	for (int g = 0; g < 8; g ++) ones [g] = 1;
	for (int e = 0; e < n; e ++) ones [fill_order [e % 8]] ++;
	for (int g = 0; g < 8; g ++)
		for (int b = 0; b < 4; b ++) out [4 * g + b] = b < ones [g];
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Viterbi -- src/backend/viterbi.cpp:130-182 (tables), :225-242 (deconvolve), :245-271 (butterfly),
 * :333-357 (chain-back), :360-371 (initial metrics); spiral-no-sse.c:209-225 (strict '>' tie-break).
 * Runs exactly frameBits+6 trellis steps (the reference's SPIRAL loop runs 2x that, the surplus is never
 * read by the chain-back: SURVEY.md Appendix B-1).
 * ---------------------------------------------------------------------------------------------- */
static int parity32 (unsigned x) { x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1; }

int k_viterbi (int frameBits, const int16_t *soft, uint8_t *out) {
	static const int polys [4] = { 0155, 0117, 0123, 0155 };       /* viterbi.cpp:63 */
	const int nsteps = frameBits + 6;
	uint32_t branch [4][32];
	for (int s = 0; s < 32; s ++)                                   /* viterbi.cpp:159-164 */
		for (int j = 0; j < 4; j ++) branch [j][s] = parity32 ((2 * s) & polys [j]) ? 255 : 0;
	uint64_t *dec = (uint64_t *) calloc (nsteps, sizeof (uint64_t));
	if (!dec) return -1;
	uint32_t ma [64], mb [64], *old = ma, *nw = mb;
	for (int i = 0; i < 64; i ++) old [i] = 63;                     /* viterbi.cpp:364-370 */
	old [0] = 0;
	for (int t = 0; t < nsteps; t ++) {
		uint32_t sym [4];
		for (int j = 0; j < 4; j ++) {                                /* viterbi.cpp:229-235 */
			int v = soft [4 * t + j] + 127;
			sym [j] = v < 0 ? 0 : v > 255 ? 255 : v;
		}
		uint64_t d = 0;
		for (int i = 0; i < 32; i ++) {                               /* viterbi.cpp:250-270 */
			uint32_t m = 0;
			for (int j = 0; j < 4; j ++) m += branch [j][i] ^ sym [j];
			const uint32_t m0 = old [i] + m,          m1 = old [i + 32] + (1020 - m);
			const uint32_t m2 = old [i] + (1020 - m), m3 = old [i + 32] + m;
			const int d0 = (int32_t) (m0 - m1) > 0, d1 = (int32_t) (m2 - m3) > 0;
			nw [2 * i]     = d0 ? m1 : m0;
			nw [2 * i + 1] = d1 ? m3 : m2;
			d |= ((uint64_t) d0 << (2 * i)) | ((uint64_t) d1 << (2 * i + 1));
		}
		dec [t] = d;
		uint32_t *tmp = old; old = nw; nw = tmp;
	}
	/* chain-back from state 0 (viterbi.cpp:238, 343-356): the decision of step i+6 along the survivor
	 * is information bit i; the predecessor state is (state >> 1) | (bit << 5) */
	unsigned state = 0;
	for (int i = frameBits - 1; i >= 0; i --) {
		const unsigned k = (unsigned) (dec [i + 6] >> state) & 1;
		state = (state >> 1) | (k << 5);
		out [i] = (uint8_t) k;
	}
	free (dec);
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * protection classes -- src/backend/deconvolve.cpp:142-237 (uep), :244-366 (eep)
 * ---------------------------------------------------------------------------------------------- */
struct k_prot { int bitRate, nseg; int16_t L [4], PI [4]; int16_t *block; };

k_prot *k_prot_new (int bitRate, int uepFlag, int protLevel) {
	k_prot *p = (k_prot *) calloc (1, sizeof (k_prot));
	p -> bitRate = bitRate;
	if (uepFlag == 0) {                  /* dab-concurrent.cpp:78-83: uepFlag == 0 selects UEP */
		if (orc_uep_profile (bitRate, protLevel, p -> L, p -> PI) != 0) {
			fprintf (stderr, "%d (%d) has a problem\n", bitRate, protLevel);
			/* deconvolve.cpp:148-151 falls back to table row 1 = {32, 4} */
			orc_uep_profile (32, 4, p -> L, p -> PI);
		}
		p -> nseg = 4;
	} else {
		if (orc_eep_profile (bitRate, protLevel, p -> L, p -> PI) != 0) { free (p); return NULL; }
		p -> nseg = 2;
	}
	p -> block = (int16_t *) malloc (sizeof (int16_t) * (24 * bitRate * 4 + 24));
	return p;
}
void k_prot_free (k_prot *p) { if (p) { free (p -> block); free (p); } }

int k_prot_deconvolve (k_prot *p, const int16_t *v, int32_t size, uint8_t *out) {
	(void) size;                         /* deconvolve.cpp:184, 334: unused */
	const int outSize = 24 * p -> bitRate;
	int32_t in = 0, vc = 0;
	memset (p -> block, 0, sizeof (int16_t) * (outSize * 4 + 24));   /* :185 */
	for (int s = 0; s < p -> nseg; s ++) {                          /* :186-223 / :338-355 */
		int8_t pi [32];
		if (p -> L [s] <= 0) continue;
		k_pcode (p -> PI [s], pi);
		for (int i = 0; i < p -> L [s]; i ++)
			for (int j = 0; j < 128; j ++) {
				if (pi [j % 32] == 1) p -> block [vc] = v [in ++];
				vc ++;
			}
	}
	for (int i = 0; i < 24; i ++) {                                 /* :226-231 tail, PI_X = 1100 x 6 */
		if ((i & 3) < 2) p -> block [vc] = v [in ++];
		vc ++;
	}
	return k_viterbi (outSize, p -> block, out);
}

/* ------------------------------------------------------------------------------------------------
 * FFT wrapper -- src/various/fft.cpp:53-55 (forward, unnormalised), :109-121 (backward, scaled by 1/N
 * where Factor = 1.0 / float(N) is a float)
 * ---------------------------------------------------------------------------------------------- */
int k_fft (float *v, int n, int inverse) {
	if (fft_standin_exec (v, n, inverse ? +1 : -1) != 0) return -1;
	if (inverse) {
		const float f = (float) (1.0 / (float) n);
		for (int i = 0; i < 2 * n; i ++) v [i] *= f;
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * phaseReference (src/ofdm/phasereference.cpp) + ofdmDecoder (src/ofdm/ofdm-decoder.cpp)
 * ---------------------------------------------------------------------------------------------- */
#define SEARCH_RANGE       72            /* ofdm-decoder.cpp:36 */
#define CORRELATION_LENGTH 18            /* ofdm-decoder.cpp:37 */
struct k_ofdm {
	orc_params p; int level, method;
	float *refTable, *phaseRef, *buf, *res; int16_t *perm;
	float refArg [CORRELATION_LENGTH];
};

static float jan_abs_f (float re, float im) {       /* dab-constants.h:127-134 */
	if (re < 0) re = -re;
	if (im < 0) im = -im;
	return re + im;
}
static float arg_mulconj (const float *a, const float *b) {      /* arg (a * conj (b)) in float */
	const float re = a [0] * b [0] + a [1] * b [1];
	const float im = a [1] * b [0] - a [0] * b [1];
	return atan2f (im, re);
}

k_ofdm *k_ofdm_new (const orc_params *p, int threshold, int freqSyncMethod) {
	k_ofdm *o = (k_ofdm *) calloc (1, sizeof (k_ofdm));
	const int T_u = p -> T_u;
	o -> p = *p; o -> level = threshold; o -> method = freqSyncMethod;
	o -> refTable = (float *) calloc (2 * T_u, sizeof (float));
	o -> phaseRef = (float *) calloc (2 * T_u, sizeof (float));
	o -> buf      = (float *) calloc (2 * T_u, sizeof (float));
	o -> res      = (float *) calloc (2 * T_u, sizeof (float));
	o -> perm     = (int16_t *) calloc (p -> K, sizeof (int16_t));
	k_perm_table (p, o -> perm);
	for (int i = 1; i <= p -> K / 2; i ++) {          /* phasereference.cpp:41-47 (float cos/sin) */
		float phi = k_phi (p -> dabMode, i);
		o -> refTable [2 * i] = cosf (phi);           o -> refTable [2 * i + 1] = sinf (phi);
		phi = k_phi (p -> dabMode, -i);
		o -> refTable [2 * (T_u - i)] = cosf (phi);   o -> refTable [2 * (T_u - i) + 1] = sinf (phi);
	}
	for (int i = 0; i < CORRELATION_LENGTH; i ++)     /* ofdm-decoder.cpp:73-76 */
		o -> refArg [i] = arg_mulconj (&o -> refTable [2 * ((T_u + i) % T_u)],
		                               &o -> refTable [2 * ((T_u + i + 1) % T_u)]);
	return o;
}
void k_ofdm_free (k_ofdm *o) {
	if (!o) return;
	free (o -> refTable); free (o -> phaseRef); free (o -> buf); free (o -> res); free (o -> perm); free (o);
}

int32_t k_find_index (k_ofdm *o, const float *v) {     /* phasereference.cpp:60-88 */
	const int Tu = o -> p. T_u;
	memcpy (o -> buf, v, sizeof (float) * 2 * Tu);
	k_fft (o -> buf, Tu, 0);
	for (int i = 0; i < Tu; i ++) {                    /* :70-71 fft * conj (refTable) */
		const float a = o -> buf [2 * i], b = o -> buf [2 * i + 1];
		const float c = o -> refTable [2 * i], d = o -> refTable [2 * i + 1];
		o -> res [2 * i]     = a * c + b * d;
		o -> res [2 * i + 1] = b * c - a * d;
	}
	k_fft (o -> res, Tu, 1);
	float sum = 0, Max = -10000;
	int32_t maxIndex = -1;
	for (int i = 0; i < Tu; i ++) sum += hypotf (o -> res [2 * i], o -> res [2 * i + 1]);
	for (int i = 0; i < Tu; i ++) {
		const float a = hypotf (o -> res [2 * i], o -> res [2 * i + 1]);
		if (a > Max) { maxIndex = i; Max = a; }
	}
	if (Max < o -> level * sum / Tu)                   /* :84-85 */
		return (int32_t) (- fabsf (Max / (sum / Tu)) - 1);
	return maxIndex;
}

static int16_t get_middle (k_ofdm *o, const float *v) {   /* ofdm-decoder.cpp:233-258 (incl. its bug) */
	const int T_u = o -> p. T_u;
	float sum = 0, oldMax = 0;
	int16_t maxIndex = 0;
	for (int i = 40; i < 1536 + 40; i ++) {
		const int k = (T_u / 2 + i) % T_u;
		sum += hypotf (v [2 * k], v [2 * k + 1]);
	}
	for (int i = 40; i < T_u - (1536 - 40); i ++) {
		int k = (T_u / 2 + i) % T_u;
		sum -= hypotf (v [2 * k], v [2 * k + 1]);
		k = (T_u / 2 + i + 1536) % T_u;
		sum += hypotf (v [2 * k], v [2 * k + 1]);
		if (sum > oldMax) { sum = oldMax; maxIndex = i; }        /* sic, :252-255 */
	}
	return maxIndex - (T_u - 1536) / 2;
}

int k_block0 (k_ofdm *o, const float *vi, int flag) {     /* ofdm-decoder.cpp:85-162 */
	const int T_u = o -> p. T_u;
	int16_t index_1 = 100;
	float *f = o -> buf;
	memcpy (f, vi, sizeof (float) * 2 * T_u);
	k_fft (f, T_u, 0);
	memcpy (o -> phaseRef, f, sizeof (float) * 2 * T_u);
	/* :93-97 SNR display value: no influence on any output, not restated */
	if (!flag) return 0;
	if (o -> method == 0) return get_middle (o, f);
	if (o -> method == 1) {                            /* :106-127 */
		float cv [SEARCH_RANGE + CORRELATION_LENGTH];
		for (int i = 0; i < SEARCH_RANGE + CORRELATION_LENGTH; i ++) {
			const int16_t base = (int16_t) (T_u - SEARCH_RANGE / 2 + i);
			cv [i] = arg_mulconj (&f [2 * (base % T_u)], &f [2 * ((base + 1) % T_u)]);
		}
		float MMax = 0;
		for (int i = 0; i < SEARCH_RANGE; i ++) {
			float sum = 0;
			for (int j = 1; j < CORRELATION_LENGTH; j ++) sum += fabsf (o -> refArg [j] * cv [i + j]);
			if (sum > MMax) { MMax = sum; index_1 = (int16_t) i; }
		}
		return T_u - SEARCH_RANGE / 2 + index_1 - T_u;
	}
	/* :132-161 zero-pattern match; the a1/a2/b1 terms are float(abs(double(abs(arg))/M_PI - 1)) */
	float Mmin = 1000;
	index_1 = 100;
#define ARGP(x, y) arg_mulconj (&f [2 * ((x) % T_u)], &f [2 * ((y) % T_u)])
	for (int i = T_u - SEARCH_RANGE / 2; i < T_u + SEARCH_RANGE / 2; i ++) {
		const float a1 = (float) fabs (fabsf (ARGP (i + 1, i + 2)) / M_PI - 1);
		const float a2 = (float) fabs (fabsf (ARGP (i + 2, i + 3)) / M_PI - 1);
		const float a3 = fabsf (ARGP (i + 3, i + 4));
		const float a4 = fabsf (ARGP (i + 4, i + 5));
		const float a5 = fabsf (ARGP (i + 5, i + 6));
		const float b1 = (float) fabs (fabsf (ARGP (i + 16 + 1, i + 16 + 3)) / M_PI - 1);
		const float b2 = fabsf (ARGP (i + 16 + 3, i + 16 + 4));
		const float b3 = fabsf (ARGP (i + 16 + 4, i + 16 + 5));
		const float b4 = fabsf (ARGP (i + 16 + 5, i + 16 + 6));
		const float sum = a1 + a2 + a3 + a4 + a5 + b1 + b2 + b3 + b4;
		if (sum < Mmin) { Mmin = sum; index_1 = (int16_t) i; }
	}
#undef ARGP
	return index_1 - T_u;
}

void k_token (k_ofdm *o, const float *inv, int16_t *ibits) {   /* ofdm-decoder.cpp:167-190 */
	const int T_u = o -> p. T_u, K = o -> p. K, delta = o -> p. T_s - o -> p. T_u;
	float *f = o -> buf;
	memcpy (f, &inv [2 * delta], sizeof (float) * 2 * T_u);
	k_fft (f, T_u, 0);
	for (int i = 0; i < K; i ++) {
		int index = o -> perm [i];
		if (index < 0) index += T_u;
		const float a = f [2 * index], b = f [2 * index + 1];
		const float c = o -> phaseRef [2 * index], d = o -> phaseRef [2 * index + 1];
		const float re = a * c + b * d, im = b * c - a * d;    /* fft * conj (phaseReference) */
		o -> phaseRef [2 * index] = a; o -> phaseRef [2 * index + 1] = b;
		const float ab1 = jan_abs_f (re, im);
		/* :188-189 float negate and divide, double multiply, truncation; NaN (ab1 == 0) converts like
		 * x86 cvttsd2si (0x80000000 -> low 16 bits 0), SURVEY.md Appendix B-5 */
		const double x = (double) (- re / ab1) * 127.0, y = (double) (- im / ab1) * 127.0;
		ibits [i]     = isnan (x) ? 0 : (int16_t) x;
		ibits [K + i] = isnan (y) ? 0 : (int16_t) y;
	}
}

void k_get_phase_reference (k_ofdm *o, float *out) { memcpy (out, o -> phaseRef, sizeof (float) * 2 * o -> p. T_u); }
void k_get_ref_table (k_ofdm *o, float *out) { memcpy (out, o -> refTable, sizeof (float) * 2 * o -> p. T_u); }

/* ====================================================================================================
 * DAB+ super-frame layer primitives: Fire code check and RS(120,110) over GF(2^8).
 * ==================================================================================================== */
/* ---- firecode_checker (firecode-checker.cpp:33-95): g(x) = (x^11+1)(x^5+x^3+x^2+x+1), byte-wise by table ---- */
static uint16_t fire_tab [256];
static int fire_ready;
static uint16_t fire_run8 (uint8_t regs [16]) {                         /* :59-75 */
	static const uint8_t g [16] = { 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0 };
	for (int i = 0; i < 8; i ++) {
		const uint8_t z = regs [15];
		for (int j = 15; j > 0; j --) regs [j] = regs [j - 1] ^ (z & g [j]);
		regs [0] = z;
	}
	uint16_t v = 0;
	for (int i = 15; i >= 0; i --) v = (uint16_t) ((v << 1) | regs [i]);
	return v;
}
static void fire_init (void) {                                           /* :36-54 */
	uint16_t itab [8];
	for (int i = 0; i < 8; i ++) {
		uint8_t regs [16];
		memset (regs, 0, 16);
		regs [8 + i] = 1;
		itab [i] = fire_run8 (regs);
	}
	for (int i = 0; i < 256; i ++) {
		fire_tab [i] = 0;
		for (int j = 0; j < 8; j ++) if (i & (1 << j)) fire_tab [i] ^= itab [j];
	}
	fire_ready = 1;
}
void k_firecode_table (uint16_t *tab256) { if (!fire_ready) fire_init (); memcpy (tab256, fire_tab, sizeof (fire_tab)); }
int k_firecode_check (const uint8_t *x) {                                /* :77-95 */
	if (!fire_ready) fire_init ();
	uint16_t state = (uint16_t) ((x [2] << 8) | x [3]), istate;
	for (int i = 4; i < 11; i ++) {
		istate = fire_tab [state >> 8];
		state = (uint16_t) (((istate & 0x00ff) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00));
	}
	for (int i = 0; i < 2; i ++) {
		istate = fire_tab [state >> 8];
		state = (uint16_t) (((istate & 0x00ff) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00));
	}
	return state == 0;
}

/* ---- galois (galois.cpp:39-127) for symsize 8, gfpoly 0435; reedSolomon (8, 0435, 0, 1, 10) (mp4processor.cpp:75) ---- */
#define RS_NN 255
#define RS_NROOTS 10
static uint16_t gf_alpha_to [256], gf_index_of [256];
static uint8_t rs_generator [RS_NROOTS + 1];
static int gf_ready;
static int gf_modnn (int x) { while (x >= RS_NN) { x -= RS_NN; x = (x >> 8) + (x & RS_NN); } return x; }     /* :64-70 */
static int gf_mul_power (int a, int b) { return gf_modnn (a + b); }
static int gf_mul_poly (int a, int b) { return (a == 0 || b == 0) ? 0 : gf_alpha_to [gf_mul_power (gf_index_of [a], gf_index_of [b])]; }
static int gf_div_power (int a, int b) { return gf_modnn (256 - 1 + a - b); }
static int gf_div_poly (int a, int b) { return a == 0 ? 0 : gf_alpha_to [gf_div_power (gf_index_of [a], gf_index_of [b])]; }
static int gf_pow_power (int a, int n) { return a == 0 ? 0 : (a * n) % 255; }
static void gf_init (void) {
	int sr = 1;
	gf_index_of [0] = RS_NN; gf_alpha_to [RS_NN] = 0;
	for (int i = 0; i < RS_NN; i ++) {
		gf_index_of [sr] = (uint16_t) i; gf_alpha_to [i] = (uint16_t) sr;
		sr <<= 1;
		if (sr & 256) sr ^= 0435;
		sr &= RS_NN;
	}
	/* generator polynomial, reed-solomon.cpp:47-76 (fcr 0, prim 1) */
	memset (rs_generator, 0, sizeof (rs_generator));
	rs_generator [0] = 1;
	for (int i = 0, root = 0; i < RS_NROOTS; i ++, root ++) {
		rs_generator [i + 1] = 1;
		for (int j = i; j > 0; j --) {
			if (rs_generator [j] != 0)
				rs_generator [j] = (uint8_t) (rs_generator [j - 1] ^ gf_alpha_to [gf_mul_power (gf_index_of [rs_generator [j]], root)]);
			else
				rs_generator [j] = rs_generator [j - 1];
		}
		rs_generator [0] = (uint8_t) gf_alpha_to [gf_mul_power (root, gf_index_of [rs_generator [0]])];
	}
	for (int i = 0; i <= RS_NROOTS; i ++) rs_generator [i] = (uint8_t) gf_index_of [rs_generator [i]];
	gf_ready = 1;
}
void k_gf_tables (uint16_t *alpha_to, uint16_t *index_of) {
	if (!gf_ready) gf_init ();
	memcpy (alpha_to, gf_alpha_to, sizeof (gf_alpha_to)); memcpy (index_of, gf_index_of, sizeof (gf_index_of));
}

/* reedSolomon::decode_rs on a full 255-symbol word (reed-solomon.cpp:145-232 with :236-399) */
static int rs_decode (uint8_t *data) {
	uint8_t syn [RS_NROOTS + 1], Lambda [RS_NROOTS + 1], Corrector [RS_NROOTS + 1], omega [RS_NROOTS + 1];
	uint8_t rootTable [RS_NROOTS], locTable [RS_NROOTS];
	/* syndromes by Horner, :236-270 (uu1 = root for fcr 0, prim 1) */
	int any = 0;
	for (int i = 0; i < RS_NROOTS; i ++) {
		int s = data [0];
		for (int j = 1; j < RS_NN; j ++)
			s = s == 0 ? data [j] : data [j] ^ gf_alpha_to [gf_mul_power (gf_index_of [s], gf_pow_power (gf_mul_power (0, i), 1))];
		syn [i] = (uint8_t) s; any |= s;
	}
	syn [RS_NROOTS] = 0;                                     /* the reference reads one element past its array here; the value is never used */
	if (!any) return 0;
	/* Berlekamp-Massey, :275-325 */
	int K = 1, L = 0, deg_lambda = 0;
	memset (Lambda, 0, sizeof (Lambda)); memset (Corrector, 0, sizeof (Corrector));
	int error = syn [0];
	Lambda [0] = 1; Corrector [1] = 1;
	while (K <= RS_NROOTS) {
		uint8_t old [RS_NROOTS + 1];
		memcpy (old, Lambda, sizeof (old));
		for (int i = 0; i < RS_NROOTS + 1; i ++) Lambda [i] ^= (uint8_t) gf_mul_poly (error, Corrector [i]);
		if (2 * L < K && error != 0) {
			L = K - L;
			for (int i = 0; i < RS_NROOTS + 1; i ++) Corrector [i] = (uint8_t) gf_div_poly (old [i], error);
		}
		for (int i = RS_NROOTS; i >= 1; i --) Corrector [i] = Corrector [i - 1];
		Corrector [0] = 0;
		error = syn [K];
		for (int i = 1; i <= K; i ++) error ^= gf_mul_poly (syn [K - i], Lambda [i]);
		K ++;
	}
	for (int i = 0; i < RS_NROOTS + 1; i ++) {
		if (Lambda [i] != 0) deg_lambda = i;
		Lambda [i] = (uint8_t) gf_index_of [Lambda [i]];
	}
	/* Chien search, :330-362 (iprim = 1) */
	int rootCount = 0;
	{
		uint8_t reg [RS_NROOTS + 1];
		memcpy (reg, Lambda, sizeof (reg));
		for (int i = 1, k = 0; i <= RS_NN; i ++, k ++) {
			int result = 1;
			for (int j = deg_lambda; j > 0; j --)
				if (reg [j] != RS_NN) { reg [j] = (uint8_t) gf_mul_power (reg [j], j); result ^= gf_alpha_to [reg [j]]; }
			if (result != 0) continue;
			if (rootCount < RS_NROOTS) { rootTable [rootCount] = (uint8_t) i; locTable [rootCount] = (uint8_t) k; }
			rootCount ++;
		}
		if (rootCount != deg_lambda) return -1;
	}
	/* omega = s * lambda mod x^nroots, :372-399 */
	int deg_omega = 0;
	for (int i = 0; i < RS_NROOTS; i ++) {
		int tmp = 0;
		for (int j = deg_lambda < i ? deg_lambda : i; j >= 0; j --)
			if (gf_index_of [syn [i - j]] != RS_NN && Lambda [j] != RS_NN)
				tmp ^= gf_alpha_to [gf_mul_power (gf_index_of [syn [i - j]], Lambda [j])];
		if (tmp != 0) deg_omega = i;
		omega [i] = (uint8_t) gf_index_of [tmp];
	}
	omega [RS_NROOTS] = RS_NN;
	/* Forney, :171-231 */
	for (int j = rootCount - 1; j >= 0; j --) {
		int num1 = 0, den = 0;
		for (int i = deg_omega; i >= 0; i --)
			if (omega [i] != RS_NN) num1 ^= gf_alpha_to [gf_mul_power (omega [i], gf_pow_power (i, rootTable [j]))];
		const int num2 = gf_alpha_to [gf_mul_power (gf_pow_power (rootTable [j], gf_div_power (0, 1)), RS_NN)];
		for (int i = (deg_lambda < RS_NROOTS - 1 ? deg_lambda : RS_NROOTS - 1) & ~1; i >= 0; i -= 2)
			if (Lambda [i + 1] != RS_NN) den ^= gf_alpha_to [gf_mul_power (Lambda [i + 1], gf_pow_power (i, rootTable [j]))];
		if (den == 0) return -1;
		if (num1 != 0) {
			if (locTable [j] >= (uint8_t) (RS_NN - RS_NROOTS)) rootCount --;
			else {
				int y = gf_mul_power (gf_index_of [num1], gf_index_of [num2]);
				y = gf_mul_power (y, RS_NN - gf_index_of [den]);
				data [locTable [j]] ^= (uint8_t) gf_alpha_to [y];
			}
		}
	}
	return rootCount;
}

int k_rs_dec (const uint8_t *r120, uint8_t *d110) {          /* reedSolomon::dec (r, d, 135), reed-solomon.cpp:129-143 */
	if (!gf_ready) gf_init ();
	uint8_t rf [RS_NN];
	memset (rf, 0, 135);
	memcpy (rf + 135, r120, 120);
	const int ret = rs_decode (rf);
	memcpy (d110, rf + 135, 110);
	return ret;
}

void k_rs_enc (const uint8_t *d110, uint8_t *r120) {         /* reedSolomon::enc (r, d, 135), :79-127 */
	if (!gf_ready) gf_init ();
	uint8_t rf [RS_NN], bb [RS_NROOTS];
	memset (rf, 0, 135);
	memcpy (rf + 135, d110, 110);
	memset (bb, 0, sizeof (bb));
	for (int i = 0; i < RS_NN - RS_NROOTS; i ++) {
		const int feedback = gf_index_of [rf [i] ^ bb [0]];
		if (feedback != RS_NN)
			for (int j = 1; j < RS_NROOTS; j ++) bb [j] ^= (uint8_t) gf_alpha_to [gf_mul_power (feedback, rs_generator [RS_NROOTS - j])];
		memmove (&bb [0], &bb [1], RS_NROOTS - 1);
		bb [RS_NROOTS - 1] = feedback != RS_NN ? (uint8_t) gf_alpha_to [gf_mul_power (feedback, rs_generator [0])] : 0;
	}
	memcpy (r120, rf + 135, 110);
	memcpy (r120 + 110, bb, RS_NROOTS);
}
