/*
 * dab_oracle.c -- the oracle's public API (dab_oracle.h): shared constant tables plus the Tier-C loops
 * (frame/sync/AFC loop, FIC regroup, CIF assembly, time de-interleaver, energy dispersal, CRC) restated
 * in plain C from the reference, on top of the Tier-A/B primitives of orc_kernels.h.  The same file is
 * linked into liboracle.so (primitives = orc_port.c) and into _ref/libdabref.so (primitives = the
 * reference's own classes, ref_shim/ref_provider.cpp).
 * TEST INFRASTRUCTURE ONLY -- nothing in the product path may call this (see dab_oracle.h).
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "dab_oracle.h"
#include "orc_kernels.h"

#define INPUT_RATE 2048000              /* includes/dab-constants.h:79 */

const char *orc_build_kind (void) { return k_kind (); }

/* ---- mode parameters: gui.cpp:1328-1372; CIF geometry: msc-handler.cpp:61-71; FIC: fic-handler.cpp:192-230 */
int orc_mode_params (int mode, orc_params *p) {
	memset (p, 0, sizeof (*p));
	switch (mode) {
	   case 1: p -> L = 76;  p -> K = 1536; p -> T_null = 2656; p -> T_F = 196608; p -> T_s = 2552;
	           p -> T_u = 2048; p -> carrierDiff = 1000; p -> blocksPerCIF = 18; break;
	   case 2: p -> L = 76;  p -> K = 384;  p -> T_null = 664;  p -> T_F = 49152;  p -> T_s = 638;
	           p -> T_u = 512;  p -> carrierDiff = 4000; p -> blocksPerCIF = 72; break;
	   case 3: p -> L = 153; p -> K = 192;  p -> T_null = 345;  p -> T_F = 49152;  p -> T_s = 319;
	           p -> T_u = 256;  p -> carrierDiff = 2000; p -> blocksPerCIF = 18; break;   /* sic: ref values */
	   case 4: p -> L = 76;  p -> K = 768;  p -> T_null = 1328; p -> T_F = 98304;  p -> T_s = 1276;
	           p -> T_u = 1024; p -> carrierDiff = 2000; p -> blocksPerCIF = 36; break;
	   default: return -1;
	}
	p -> dabMode = mode;
	p -> T_g = p -> T_s - p -> T_u;
	p -> ficSymbols = 3;                               /* ofdm-processor.cpp:421-422 */
	p -> ficGroups = 3 * 2 * p -> K / 2304;            /* fic-handler.cpp:211-217 */
	p -> cifsPerFrame = (p -> L - 4) / p -> blocksPerCIF;
	return 0;
}

int   orc_perm_table (int mode, int16_t *out) { orc_params p; if (orc_mode_params (mode, &p)) return -1; return k_perm_table (&p, out); }
float orc_phi (int mode, int k) { return k_phi (mode, k); }
int   orc_pcode (int n, int8_t *out) { return k_pcode (n, out); }
int   orc_ref_table (int mode, float *out) {
	orc_params p; if (orc_mode_params (mode, &p)) return -1;
	k_ofdm *o = k_ofdm_new (&p, 3, 1);
	k_get_ref_table (o, out);
	k_ofdm_free (o);
	return 0;
}

/* ---- UEP profiles: deconvolve.cpp:28-114 {bitRate, protLevel, L1..L4, PI1..PI4}; PI4 == -1 -> unused.
 * Row {80,1} keeps the reference's PI2 = 7. */
static const int16_t uep_rows [][10] = {
	{32,5, 3,4,17,0, 5,3,2,-1},     {32,4, 3,3,18,0, 11,6,5,-1},    {32,3, 3,4,14,3, 15,9,6,8},
	{32,2, 3,4,14,3, 22,13,8,13},   {32,1, 3,5,13,3, 24,17,12,17},
	{48,5, 4,3,26,3, 5,4,2,3},      {48,4, 3,4,26,3, 9,6,4,6},      {48,3, 3,4,26,3, 15,10,6,9},
	{48,2, 3,4,26,3, 24,14,8,15},   {48,1, 3,5,25,3, 24,18,13,18},
	{64,5, 6,9,31,2, 5,3,2,3},      {64,4, 6,9,33,0, 11,6,6,-1},    {64,3, 6,12,27,3, 16,8,6,9},
	{64,2, 6,10,29,3, 23,13,8,13},  {64,1, 6,11,28,3, 24,18,12,18},
	{80,5, 6,10,41,3, 6,3,2,3},     {80,4, 6,10,41,3, 11,6,5,6},    {80,3, 6,11,40,3, 16,8,6,7},
	{80,2, 6,10,41,3, 23,13,8,13},  {80,1, 6,10,41,3, 24,7,12,18},
	{96,5, 7,9,53,3, 5,4,2,4},      {96,4, 7,10,52,3, 9,6,4,6},     {96,3, 6,12,51,3, 16,9,6,10},
	{96,2, 6,10,53,3, 22,12,9,12},  {96,1, 6,13,50,3, 24,18,13,19},
	{112,5, 14,17,50,3, 5,4,2,5},   {112,4, 11,21,49,3, 9,6,4,8},   {112,3, 11,23,47,3, 16,8,6,9},
	{112,2, 11,21,49,3, 23,12,9,14},
	{128,5, 12,19,62,3, 5,3,2,4},   {128,4, 11,21,61,3, 11,6,5,7},  {128,3, 11,22,60,3, 16,9,6,10},
	{128,2, 11,21,61,3, 22,12,9,14}, {128,1, 11,20,62,3, 24,17,13,19},
	{160,5, 11,19,87,3, 5,4,2,4},   {160,4, 11,23,83,3, 11,6,5,9},  {160,3, 11,24,82,3, 16,8,6,11},
	{160,2, 11,21,85,3, 22,11,9,13}, {160,1, 11,22,84,3, 24,18,12,19},
	{192,5, 11,20,110,3, 6,4,2,5},  {192,4, 11,22,108,3, 10,6,4,9}, {192,3, 11,24,106,3, 16,10,6,11},
	{192,2, 11,20,110,3, 22,13,9,13}, {192,1, 11,21,109,3, 24,20,13,24},
	{224,5, 12,22,131,3, 8,6,2,6},  {224,4, 12,26,127,3, 12,8,4,11}, {224,3, 11,20,134,3, 16,10,7,9},
	{224,2, 11,22,132,3, 24,16,10,15}, {224,1, 11,24,130,3, 24,20,12,20},
	{256,5, 11,24,154,3, 6,5,2,5},  {256,4, 11,24,154,3, 12,9,5,10}, {256,3, 11,27,151,3, 16,10,7,10},
	{256,2, 11,22,156,3, 24,14,10,13}, {256,1, 11,26,152,3, 24,19,14,18},
	{320,5, 11,26,200,3, 8,5,2,6},  {320,4, 11,25,201,3, 13,9,5,10}, {320,2, 11,26,200,3, 24,17,9,17},
	{384,5, 11,27,247,3, 8,6,2,7},  {384,3, 11,24,250,3, 16,9,7,10}, {384,1, 12,28,245,3, 24,20,14,23},
};

int orc_uep_profile (int bitRate, int protLevel, int16_t L [4], int16_t PI [4]) {
	for (unsigned r = 0; r < sizeof (uep_rows) / sizeof (uep_rows [0]); r ++)
		if (uep_rows [r][0] == bitRate && uep_rows [r][1] == protLevel) {        /* deconvolve.cpp:123-133 */
			for (int j = 0; j < 4; j ++) { L [j] = uep_rows [r][2 + j]; PI [j] = uep_rows [r][6 + j]; }
			if (PI [3] < 0) PI [3] = 0;                                           /* :162-165, never read: L4 == 0 */
			return 0;
		}
	return -1;
}

/* EEP profiles: deconvolve.cpp:249-318.  protLevel is 0100+level (A) or 0200+level (B). */
int orc_eep_profile (int bitRate, int protLevel, int16_t L [2], int16_t PI [2]) {
	const int lvl = protLevel & 07;
	if (protLevel & 0100) {
		switch (lvl) {
		   case 1: L [0] = 6 * bitRate / 8 - 3; L [1] = 3; PI [0] = 24; PI [1] = 23; return 0;
		   case 2: if (bitRate == 8) { L [0] = 5; L [1] = 1; PI [0] = 13; PI [1] = 12; }
		           else { L [0] = 2 * bitRate / 8 - 3; L [1] = 4 * bitRate / 8 + 3; PI [0] = 14; PI [1] = 13; }
		           return 0;
		   case 3: L [0] = 6 * bitRate / 8 - 3; L [1] = 3; PI [0] = 8; PI [1] = 7; return 0;
		   case 4: L [0] = 4 * bitRate / 8 - 3; L [1] = 2 * bitRate / 8 + 3; PI [0] = 3; PI [1] = 2; return 0;
		}
	} else if (protLevel & 0200) {
		L [0] = 24 * bitRate / 32 - 3; L [1] = 3;
		switch (lvl) {
		   case 1: PI [0] = 10; PI [1] = 9; return 0;
		   case 2: PI [0] = 6;  PI [1] = 5; return 0;
		   case 3: PI [0] = 4;  PI [1] = 3; return 0;
		   case 4: PI [0] = 2;  PI [1] = 1; return 0;
		}
	}
	return -1;          /* the reference leaves L1/L2/PI uninitialised here (UB); the oracle reports it */
}

/* ---- energy-dispersal PRBS: fic-handler.cpp:100-108 (== dab-concurrent.cpp:183-190) ---- */
void orc_prbs (uint8_t *out, int n) {
	uint8_t sr [9];
	memset (sr, 1, 9);
	for (int i = 0; i < n; i ++) {
		const uint8_t b = sr [8] ^ sr [4];
		for (int j = 8; j > 0; j --) sr [j] = sr [j - 1];
		sr [0] = b;
		out [i] = b;
	}
}

/* ---- FIB CRC: dab-constants.h:310-340, on a private copy ---- */
int orc_check_crc (const uint8_t *src, int size) {
	static const uint8_t poly [15] = { 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0 };
	uint8_t b [16], in [4096];
	if (size < 16 || size > 4096) return 0;
	memcpy (in, src, size);
	memset (b, 1, 16);
	for (int i = size - 16; i < size; i ++) in [i] ^= 1;
	for (int i = 0; i < size; i ++) {
		if ((b [0] ^ in [i]) == 1) {
			for (int f = 0; f < 15; f ++) b [f] = poly [f] ^ b [f + 1];
			b [15] = 1;
		} else {
			memmove (&b [0], &b [1], 15);
			b [15] = 0;
		}
	}
	int sum = 0;
	for (int i = 0; i < 16; i ++) sum += b [i];
	return sum == 0;
}

int orc_viterbi (int frameBits, const int16_t *soft, uint8_t *out) { return k_viterbi (frameBits, soft, out); }

static int prot_once (int bitRate, int uepFlag, int protLevel, const int16_t *v, int32_t size, uint8_t *out) {
	k_prot *p = k_prot_new (bitRate, uepFlag, protLevel);
	if (!p) return -1;
	k_prot_deconvolve (p, v, size, out);
	k_prot_free (p);
	return 0;
}
int orc_eep_deconvolve (int bitRate, int protLevel, const int16_t *v, int32_t size, uint8_t *out) {
	int16_t L [2], PI [2];
	if (orc_eep_profile (bitRate, protLevel, L, PI)) return -1;
	return prot_once (bitRate, 1, protLevel, v, size, out);
}
int orc_uep_deconvolve (int bitRate, int protLevel, const int16_t *v, int32_t size, uint8_t *out) {
	int16_t L [4], PI [4];
	if (orc_uep_profile (bitRate, protLevel, L, PI)) return -1;
	return prot_once (bitRate, 0, protLevel, v, size, out);
}

/* ---- FIC code word: fic-handler.cpp:241-321 (without the fib_processor hand-over) ---- */
int orc_fic_decode (const int16_t *fic, uint8_t *bits768, uint8_t *crc_ok) {
	int16_t block [3072 + 24];
	int8_t pi16 [32], pi15 [32];
	uint8_t prbs [768];
	int in = 0, local = 0;
	k_pcode (16, pi16); k_pcode (15, pi15);
	for (int i = 0; i < 21; i ++)                       /* :254-262 */
		for (int k = 0; k < 128; k ++) block [local ++] = pi16 [k % 32] == 1 ? fic [in ++] : 0;
	for (int i = 0; i < 3; i ++)                        /* :269-276 */
		for (int k = 0; k < 128; k ++) block [local ++] = pi15 [k % 32] == 1 ? fic [in ++] : 0;
	for (int k = 0; k < 24; k ++)                       /* :282-288, PI_X = 1100 x 6 */
		block [local ++] = (k & 3) < 2 ? fic [in ++] : 0;
	if (k_viterbi (768, block, bits768)) return -1;     /* :293 */
	orc_prbs (prbs, 768);
	for (int i = 0; i < 768; i ++) bits768 [i] ^= prbs [i];   /* :301-302 */
	for (int i = 0; i < 3; i ++) crc_ok [i] = (uint8_t) orc_check_crc (&bits768 [256 * i], 256);  /* :310-312 */
	return 0;
}

/* ---- time de-interleaver: dab-concurrent.cpp:41-43, 162-169 ---- */
static const int8_t interleave_delays [16] = { 15, 7, 11, 3, 13, 5, 9, 1, 14, 6, 10, 2, 12, 4, 8, 0 };

typedef struct { int fragmentSize; int16_t *hist; } deint_t;      /* hist[i][16] shift registers */
static deint_t *deint_new (int fragmentSize) {
	deint_t *d = (deint_t *) malloc (sizeof (deint_t));
	d -> fragmentSize = fragmentSize;
	d -> hist = (int16_t *) calloc ((size_t) fragmentSize * 16, sizeof (int16_t));
	return d;
}
static void deint_free (deint_t *d) { free (d -> hist); free (d); }
static void deint_step (deint_t *d, int16_t *Data) {                /* in place, like the reference */
	for (int i = 0; i < d -> fragmentSize; i ++) {
		int16_t *h = &d -> hist [16 * (size_t) i];
		const int dl = interleave_delays [i & 017];
		h [dl] = Data [i];
		Data [i] = h [0];
		memmove (&h [0], &h [1], dl * sizeof (int16_t));
	}
}

void orc_time_deinterleave (const int16_t *in, int ncif, int fragmentSize, int16_t *out) {
	deint_t *d = deint_new (fragmentSize);
	for (int t = 0; t < ncif; t ++) {
		memcpy (&out [(size_t) t * fragmentSize], &in [(size_t) t * fragmentSize], sizeof (int16_t) * fragmentSize);
		deint_step (d, &out [(size_t) t * fragmentSize]);
	}
	deint_free (d);
}

/* ---- MSC back end: dab-concurrent.cpp:144-193 ---- */
int orc_msc_backend (const int16_t *in, int ncif, int fragmentSize, int bitRate, int uepFlag,
                     int protLevel, uint8_t *out) {
	k_prot *prot = k_prot_new (bitRate, uepFlag, protLevel);
	if (!prot) return -1;
	deint_t *d = deint_new (fragmentSize);
	int16_t *Data = (int16_t *) malloc (sizeof (int16_t) * fragmentSize);
	uint8_t *prbs = (uint8_t *) malloc (24 * bitRate);
	orc_prbs (prbs, 24 * bitRate);
	int countforInterleaver = 0, nout = 0;
	for (int t = 0; t < ncif; t ++) {
		memcpy (Data, &in [(size_t) t * fragmentSize], sizeof (int16_t) * fragmentSize);   /* :161 */
		deint_step (d, Data);                                                                /* :162-169 */
		if (countforInterleaver <= 15) { countforInterleaver ++; continue; }                 /* :172-175 */
		uint8_t *outV = &out [(size_t) nout * 24 * bitRate];
		k_prot_deconvolve (prot, Data, fragmentSize, outV);                                  /* :177-180 */
		for (int i = 0; i < 24 * bitRate; i ++) outV [i] ^= prbs [i];                        /* :183-190 */
		nout ++;
	}
	free (prbs); free (Data); deint_free (d); k_prot_free (prot);
	return nout;
}

/* ---- FIC regroup over frames: fic-handler.cpp:192-230 ---- */
int orc_fic_frames (int mode, const int16_t *sym, int nframes, uint8_t *bits, uint8_t *crc_ok) {
	orc_params p;
	if (orc_mode_params (mode, &p)) return -1;
	const int bpb = 2 * p. K;
	int16_t ofdm_input [2304];
	int index = 0, ngroups = 0;
	for (int f = 0; f < nframes; f ++)
		for (int blkno = 1; blkno < 4; blkno ++) {
			const int16_t *data = &sym [((size_t) f * (p. L - 1) + (blkno - 1)) * bpb];
			if (blkno == 1) index = 0;                                    /* :206-209 */
			for (int i = 0; i < bpb; i ++) {
				ofdm_input [index ++] = data [i];
				if (index >= 2304) {                                      /* :213-217 */
					orc_fic_decode (ofdm_input, &bits [(size_t) ngroups * 768], &crc_ok [(size_t) ngroups * 3]);
					index = 0;
					ngroups ++;
				}
			}
		}
	return ngroups;
}

/* ---- FIG 0/1 sub-channel organisation: fib_processor::process_FIB -> process_FIG0 -> FIG0Extension1 ->
 * HandleFIG0Extension1 (fib-processor.cpp:123-158, 163-177, 278-347), called for every FIB whose CRC holds
 * (fic-handler.cpp:309-319).  Only the ficList fields are kept (fib-processor.h:78-87).  The reference reads bits
 * through unchecked pointers; a FIG whose length field runs past the FIB reads on into the rest of the 768-bit
 * group (defined) or past it (undefined): here bits past the group read as 0. ---- */
static const int16_t fig_prot_level [64][3] = {                              /* fib-processor.cpp:32-95 */
	{16,5,32}, {21,4,32}, {24,3,32}, {29,2,32}, {35,1,32}, {24,5,48}, {29,4,48}, {35,3,48}, {42,2,48}, {52,1,48},
	{29,5,56}, {35,4,56}, {42,3,56}, {52,2,56}, {32,5,64}, {42,4,64}, {48,3,64}, {58,2,64}, {70,1,64}, {40,5,80},
	{52,4,80}, {58,3,80}, {70,2,80}, {84,1,80}, {48,5,96}, {58,4,96}, {70,3,96}, {84,2,96}, {104,1,96}, {58,5,112},
	{70,4,112}, {84,3,112}, {104,2,112}, {64,5,128}, {84,4,128}, {96,3,128}, {116,2,128}, {140,1,128}, {80,5,160},
	{104,4,160}, {116,3,160}, {140,2,160}, {168,1,160}, {96,5,192}, {116,4,192}, {140,3,192}, {168,2,192}, {208,1,192},
	{116,5,224}, {140,4,224}, {168,3,224}, {208,2,224}, {232,1,224}, {128,5,256}, {168,4,256}, {192,3,256}, {232,2,256},
	{280,1,256}, {160,5,320}, {208,4,320}, {280,2,320}, {192,5,384}, {280,3,384}, {416,1,384} };

typedef struct { const uint8_t *g; int lim; } fig_bits;                      /* g = the 768-bit group, reads relative to it */
static int fig_get (const fig_bits *b, int off, int n) {                     /* getBits, dab-constants.h:182-191 */
	int r = 0;
	for (int i = 0; i < n; i ++) r = (r << 1) | (off + i < b -> lim ? (b -> g [off + i] & 1) : 0);
	return r;
}

static int fig0_ext1_entry (const fig_bits *b, int d, int offset, orc_subch_info *list) {       /* :289-347 */
	int bitOffset = offset * 8;
	const int SubChId = fig_get (b, d + bitOffset, 6), StartAdr = fig_get (b, d + bitOffset + 6, 10);
	orc_subch_info *e = &list [SubChId];
	e -> defined = 1;
	e -> startAddr = StartAdr;
	if (fig_get (b, d + bitOffset + 16, 1) == 0) {                           /* short form */
		const int ti = fig_get (b, d + bitOffset + 18, 6);
		e -> length = fig_prot_level [ti][0]; e -> uepFlag = 0;
		e -> protLevel = fig_prot_level [ti][1]; e -> bitRate = fig_prot_level [ti][2];
		bitOffset += 24;
	} else {                                                                 /* EEP long form */
		e -> uepFlag = 1;
		const int option = fig_get (b, d + bitOffset + 17, 3);
		if (option == 0 || option == 1) {
			static const int divA [5] = { 0, 12, 8, 6, 4 }, divB [5] = { 0, 27, 21, 18, 15 };
			const int protLevel = fig_get (b, d + bitOffset + 20, 2) + 1, size = fig_get (b, d + bitOffset + 22, 10);
			e -> protLevel = protLevel + (option == 0 ? 0100 : 0200);
			e -> length = size;
			e -> bitRate = option == 0 ? size / divA [protLevel] * 8 : size / divB [protLevel] * 32;
		}
		bitOffset += 32;
	}
	return bitOffset / 8;
}

int orc_fig01_scan (const uint8_t *bits, const uint8_t *crc_ok, int ngroups, orc_subch_info *list /* [64], updated */) {
	for (int gidx = 0; gidx < ngroups; gidx ++)
		for (int j = 0; j < 3; j ++) {
			if (!crc_ok [3 * gidx + j]) continue;                            /* fic-handler.cpp:311-314 */
			const fig_bits b = { bits + (size_t) gidx * 768, 768 };
			const int p = 256 * j;
			int processedBytes = 0, d = p;                                   /* process_FIB, :123-158 */
			while (processedBytes < 30) {
				const int FIGtype = fig_get (&b, d, 3);
				if (FIGtype == 7) break;
				if (FIGtype == 0 && fig_get (&b, d + 8 + 3, 5) == 1) {       /* process_FIG0 -> FIG0Extension1, :278-287 */
					const int Length = fig_get (&b, d + 3, 5);
					int used = 2;
					while (used < Length - 1) used = fig0_ext1_entry (&b, d, used, list);
				}
				processedBytes += fig_get (&b, d + 3, 5) + 1;
				d = p + processedBytes * 8;
			}
		}
	return 0;
}

int  orc_firecode_check (const uint8_t *x11) { return k_firecode_check (x11); }
int  orc_rs_dec (const uint8_t *r120, uint8_t *d110) { return k_rs_dec (r120, d110); }
void orc_rs_enc (const uint8_t *d110, uint8_t *r120) { k_rs_enc (d110, r120); }

/* ---- DAB+ super-frame layer: mp4Processor::addtoFrame + the Fire code / Reed-Solomon / AU-table part of
 * processSuperframe (mp4processor.cpp:107-150, 152-275), dabPlus_crc (:40-61).  The AAC decoder behind it is out of
 * scope; what leaves here is the corrected super frame, the access-unit table and the per-AU CRC verdicts. ---- */
struct orc_dabplus {
	int bitRate, RSDims;
	uint8_t *frameBytes, *outVector;
	int blockFillIndex, blocksInBuffer;
	int64_t cifs;
};
orc_dabplus *orc_dabplus_new (int bitRate) {
	orc_dabplus *d = (orc_dabplus *) calloc (1, sizeof (*d));
	d -> bitRate = bitRate; d -> RSDims = bitRate / 8;
	d -> frameBytes = (uint8_t *) calloc ((size_t) d -> RSDims * 120, 1);
	d -> outVector = (uint8_t *) calloc ((size_t) d -> RSDims * 110, 1);
	return d;
}
void orc_dabplus_free (orc_dabplus *d) { if (d) { free (d -> frameBytes); free (d -> outVector); free (d); } }

static int dabplus_crc (const uint8_t *msg, int len) {                      /* :40-61 */
	unsigned acc = 0xFFFF;
	for (int i = 0; i < len; i ++) {
		unsigned data = (unsigned) msg [i] << 8;
		for (int j = 8; j > 0; j --) {
			acc = ((data ^ acc) & 0x8000) ? ((acc << 1) ^ 0x1021) & 0xFFFF : (acc << 1) & 0xFFFF;
			data = (data << 1) & 0xFFFF;
		}
	}
	const unsigned crc = ~(((unsigned) msg [len] << 8) | msg [len + 1]) & 0xFFFF;
	return (crc ^ acc) == 0;
}

static int dabplus_superframe (orc_dabplus *d, int base, orc_superframe_info *fi) {     /* processSuperframe, :152-275 */
	const int R = d -> RSDims;
	uint8_t rsIn [120], rsOut [110];
	int nErrors = 0;
	for (int j = 0; j < R; j ++) {
		for (int k = 0; k < 120; k ++) rsIn [k] = d -> frameBytes [(base + j + k * R) % (R * 120)];
		const int ler = k_rs_dec (rsIn, rsOut);
		if (ler > 0) nErrors += ler;
		if (ler < 0) return 0;
		for (int k = 0; k < 110; k ++) d -> outVector [j + k * R] = rsOut [k];
	}
	const uint8_t *o = d -> outVector;
	const int dacRate = (o [2] >> 6) & 1, sbrFlag = (o [2] >> 5) & 1;
	int au [7], n;
	switch (2 * dacRate + sbrFlag) {
	   default:
	   case 0: n = 4; au [0] = 8;  au [1] = o [3] * 16 + (o [4] >> 4); au [2] = (o [4] & 0xf) * 256 + o [5];
	           au [3] = o [6] * 16 + (o [7] >> 4); au [4] = 110 * R; break;
	   case 1: n = 2; au [0] = 5;  au [1] = o [3] * 16 + (o [4] >> 4); au [2] = 110 * R; break;
	   case 2: n = 6; au [0] = 11; au [1] = o [3] * 16 + (o [4] >> 4); au [2] = (o [4] & 0xf) * 256 + o [5];
	           au [3] = o [6] * 16 + (o [7] >> 4); au [4] = (o [7] & 0xf) * 256 + o [8];
	           au [5] = o [9] * 16 + (o [10] >> 4); au [6] = 110 * R; break;
	   case 3: n = 3; au [0] = 6;  au [1] = o [3] * 16 + (o [4] >> 4); au [2] = (o [4] & 0xf) * 256 + o [5]; au [3] = 110 * R; break;
	}
	int crcmask = 0;
	for (int i = 0; i < n; i ++) {
		if (au [i + 1] < au [i]) return 0;
		const int len = au [i + 1] - au [i] - 2;
		if (len >= 960 || len < 0) return 0;
		if (dabplus_crc (&o [au [i]], len)) crcmask |= 1 << i;
	}
	fi -> corrected = nErrors; fi -> num_aus = n; fi -> au_crc = crcmask;
	for (int i = 0; i < 7; i ++) fi -> au_start [i] = i <= n ? au [i] : 0;
	return 1;
}

int orc_dabplus_process (orc_dabplus *d, const uint8_t *bits, int ncif, uint8_t *superframes, orc_superframe_info *info, int max_sf) {
	const int nbits = 24 * d -> bitRate, nb = nbits / 8;
	int nsf = 0;
	for (int c = 0; c < ncif; c ++) {                                        /* addtoFrame, :107-150 */
		const uint8_t *V = bits + (size_t) c * nbits;
		for (int i = 0; i < nb; i ++) {
			uint8_t temp = 0;
			for (int j = 0; j < 8; j ++) temp = (uint8_t) ((temp << 1) | (V [i * 8 + j] & 1));
			d -> frameBytes [d -> blockFillIndex * nb + i] = temp;
		}
		d -> blocksInBuffer ++;
		d -> blockFillIndex = (d -> blockFillIndex + 1) % 5;
		d -> cifs ++;
		if (d -> blocksInBuffer >= 5) {
			orc_superframe_info fi;
			if (k_firecode_check (&d -> frameBytes [d -> blockFillIndex * nb]) && dabplus_superframe (d, d -> blockFillIndex * nb, &fi)) {
				d -> blocksInBuffer = 0;
				if (nsf < max_sf) {
					fi. first_cif = d -> cifs - 5;
					info [nsf] = fi;
					memcpy (superframes + (size_t) nsf * d -> RSDims * 110, d -> outVector, (size_t) d -> RSDims * 110);
				}
				nsf ++;
			} else d -> blocksInBuffer = 4;
		}
	}
	return nsf;
}

/* ---- airspyHandler: sample-rate conversion to 2.048 MS/s by linear interpolation in 1 ms blocks
 * (airspy-handler.cpp:138-148 tables, 342-370 data_available).  Streaming over the whole input; returns the number
 * of output samples. ---- */
int64_t orc_resample_i16 (const int16_t *sbuf, int64_t nSamples, int selectedRate, float *out) {
	const int convBufferSize = selectedRate / 1000;
	int16_t mapTable_int [2048]; float mapTable_float [2048];
	for (int i = 0; i < 2048; i ++) {
		const float inVal = (float) (selectedRate / 1000);
		mapTable_int [i] = (int16_t) (int) floor (i * (inVal / 2048.0));
		mapTable_float [i] = (float) (i * (inVal / 2048.0) - mapTable_int [i]);
	}
	float *convBuffer = (float *) malloc (sizeof (float) * 2 * (convBufferSize + 1));
	int convIndex = 0;
	int64_t nout = 0;
	for (int64_t i = 0; i < nSamples; i ++) {
		convBuffer [2 * convIndex] = sbuf [2 * i] / (float) 2048;
		convBuffer [2 * convIndex + 1] = sbuf [2 * i + 1] / (float) 2048;
		convIndex ++;
		if (convIndex > convBufferSize) {
			for (int j = 0; j < 2048; j ++) {
				const int b = mapTable_int [j];
				const float r = mapTable_float [j], q = 1 - r;
				/* cmul (complex, float) = component-wise product (dab-constants.h) */
				out [2 * nout]     = convBuffer [2 * (b + 1)] * r + convBuffer [2 * b] * q;
				out [2 * nout + 1] = convBuffer [2 * (b + 1) + 1] * r + convBuffer [2 * b + 1] * q;
				nout ++;
			}
			convBuffer [0] = convBuffer [2 * convBufferSize]; convBuffer [1] = convBuffer [2 * convBufferSize + 1];
			convIndex = 1;
		}
	}
	free (convBuffer);
	return nout;
}

/* ---- CIF assembly + sub-channel slice: msc-handler.cpp:125-193 ---- */
int orc_msc_slice (int mode, const int16_t *sym, int nframes, int startAddr, int Length, int16_t *frag) {
	orc_params p;
	if (orc_mode_params (mode, &p) || mode == 3) return -1;     /* Mode III: "cannot happen" branch :70-71 */
	const int bpb = 2 * p. K;
	int16_t *cif = (int16_t *) calloc (55296, sizeof (int16_t));
	int ncif = 0;
	for (int f = 0; f < nframes; f ++)
		for (int blkno = 4; blkno < p. L; blkno ++) {
			const int cur = (blkno - 4) % p. blocksPerCIF;                              /* :133 */
			memcpy (&cif [cur * bpb], &sym [((size_t) f * (p. L - 1) + (blkno - 1)) * bpb], sizeof (int16_t) * bpb);
			if (cur < p. blocksPerCIF - 1) continue;                                    /* :181-182 */
			memcpy (&frag [(size_t) ncif * Length * 64], &cif [startAddr * 64], sizeof (int16_t) * Length * 64);
			ncif ++;
		}
	free (cif);
	return ncif;
}

/* ---- OFDM front end ---- */
struct orc_ofdm { orc_params p; k_ofdm *k; };
orc_ofdm *orc_ofdm_new (int mode, int threshold, int freqSyncMethod) {
	orc_ofdm *o = (orc_ofdm *) calloc (1, sizeof (orc_ofdm));
	if (orc_mode_params (mode, &o -> p)) { free (o); return NULL; }
	o -> k = k_ofdm_new (&o -> p, threshold, freqSyncMethod);
	return o;
}
void orc_ofdm_free (orc_ofdm *o) { if (o) { k_ofdm_free (o -> k); free (o); } }
int  orc_fft (float *v, int n, int inverse) { return k_fft (v, n, inverse); }
int32_t orc_find_index (orc_ofdm *o, const float *v) { return k_find_index (o -> k, v); }
int  orc_block0 (orc_ofdm *o, const float *v, int flag) { return k_block0 (o -> k, v, flag); }
void orc_token (orc_ofdm *o, const float *inv, int16_t *ibits) { k_token (o -> k, inv, ibits); }
void orc_get_phase_reference (orc_ofdm *o, float *out) { k_get_phase_reference (o -> k, out); }

/* oscillator table: ofdm-processor.cpp:76-81 (double cos/sin rounded to float), built once */
static float *osc_table;
static const float *oscillator (void) {
	if (!osc_table) {
		float *t = (float *) malloc (sizeof (float) * 2 * INPUT_RATE);
		for (int i = 0; i < INPUT_RATE; i ++) {
			t [2 * i]     = (float) cos (2.0 * M_PI * i / INPUT_RATE);
			t [2 * i + 1] = (float) sin (2.0 * M_PI * i / INPUT_RATE);
		}
		if (!__sync_bool_compare_and_swap (&osc_table, NULL, t)) free (t);
	}
	return osc_table;
}

/* the ofdmProcessor's sample pump (ofdm-processor.cpp:133-183, 186-240) over an in-memory u8 "rawfile"
 * (rawfiles.cpp:113-116) */
typedef struct {
	const uint8_t *iq; int64_t n, pos;
	const float *osc; int32_t localPhase; float sLevel;
	const float *fq;                                  /* non-NULL: complex float samples as virtualInput::getSamples delivers them */
} pump_t;

static inline int pump_sample (pump_t *s, int32_t phase, float *re, float *im) {
	if (s -> pos >= s -> n) return 0;                 /* the reference blocks/throws here (:135-145) */
	/* rawfiles.cpp:113-116 for u8 files; wavfiles.cpp:168-180 and every device handler hand over floats directly */
	const float a = s -> fq ? s -> fq [2 * s -> pos]     : (float) (s -> iq [2 * s -> pos] - 128) / 128.0f;
	const float b = s -> fq ? s -> fq [2 * s -> pos + 1] : (float) (s -> iq [2 * s -> pos + 1] - 128) / 128.0f;
	s -> pos ++;
	s -> localPhase -= phase;                                           /* :165 */
	s -> localPhase = (s -> localPhase + INPUT_RATE) % INPUT_RATE;      /* :166 */
	const float c = s -> osc [2 * s -> localPhase], d = s -> osc [2 * s -> localPhase + 1];
	*re = a * c - b * d;                                                /* :167 */
	*im = a * d + b * c;
	float ar = *re < 0 ? - *re : *re, ai = *im < 0 ? - *im : *im;
	s -> sLevel = (float) (0.00001 * (ar + ai) + (1 - 0.00001) * s -> sLevel);   /* :168 */
	return 1;
}
static inline int pump_samples (pump_t *s, float *v, int n, int32_t phase) {
	if (s -> pos + n > s -> n) return 0;
	for (int i = 0; i < n; i ++) pump_sample (s, phase, &v [2 * i], &v [2 * i + 1]);
	return 1;
}

static int ofdm_run_impl (int mode, int threshold, int freqSyncMethod, const uint8_t *iq, const float *fq, int64_t nsamples,
                          int max_frames, int16_t *sym, orc_frame_info *info);
int orc_ofdm_run (int mode, int threshold, int freqSyncMethod, const uint8_t *iq, int64_t nsamples,
                  int max_frames, int16_t *sym, orc_frame_info *info) {
	return ofdm_run_impl (mode, threshold, freqSyncMethod, iq, NULL, nsamples, max_frames, sym, info);
}
int orc_ofdm_run_cf32 (int mode, int threshold, int freqSyncMethod, const float *iq, int64_t nsamples,
                       int max_frames, int16_t *sym, orc_frame_info *info) {
	return ofdm_run_impl (mode, threshold, freqSyncMethod, NULL, iq, nsamples, max_frames, sym, info);
}
static int ofdm_run_impl (int mode, int threshold, int freqSyncMethod, const uint8_t *iq, const float *fq, int64_t nsamples,
                          int max_frames, int16_t *sym, orc_frame_info *info) {
	orc_params p;
	if (orc_mode_params (mode, &p)) return -1;
	const int T_u = p. T_u, T_s = p. T_s, T_null = p. T_null, T_F = p. T_F, K = p. K;
	k_ofdm *dec = k_ofdm_new (&p, threshold, freqSyncMethod);
	pump_t s = { iq, nsamples, 0, oscillator (), 0, 0, fq };
	float *ofdmBuffer = (float *) malloc (sizeof (float) * 2 * 76 * T_s);
	float *envBuffer = (float *) calloc (32768, sizeof (float));
	const int syncBufferMask = 32768 - 1;
	int syncBufferIndex = 0;
	float currentStrength = 0;
	int16_t fineCorrector = 0, previous_1 = 1000, previous_2 = 999;
	int32_t coarseCorrector = 0, counter, startIndex, i;
	int f2Correction = 1, nframes = 0;
	float re, im;

	s. sLevel = 0;                                                     /* :271 */
notSynced:                                                              /* :275-293 */
	syncBufferIndex = 0; currentStrength = 0; s. sLevel = 0;
	for (i = 0; i < 20 * T_s; i ++)
		if (!pump_sample (&s, 0, &re, &im)) goto done;
	syncBufferIndex = 0; currentStrength = 0;
	for (i = 0; i < 50; i ++) {
		if (!pump_sample (&s, 0, &re, &im)) goto done;
		envBuffer [syncBufferIndex] = (re < 0 ? -re : re) + (im < 0 ? -im : im);
		currentStrength += envBuffer [syncBufferIndex];
		syncBufferIndex ++;
	}
/* SyncOnNull: :298-317 */
	counter = 0;
	while (currentStrength / 50 > 0.40 * s. sLevel) {
		if (!pump_sample (&s, coarseCorrector + fineCorrector, &re, &im)) goto done;
		envBuffer [syncBufferIndex] = (re < 0 ? -re : re) + (im < 0 ? -im : im);
		currentStrength += envBuffer [syncBufferIndex] - envBuffer [(syncBufferIndex - 50) & syncBufferMask];
		syncBufferIndex = (syncBufferIndex + 1) & syncBufferMask;
		counter ++;
		if (counter > T_F) goto notSynced;
	}
	counter = 0;
/* SyncOnEndNull: :322-338 (true magnitude here) */
	while (currentStrength / 50 < 0.75 * s. sLevel) {
		if (!pump_sample (&s, coarseCorrector + fineCorrector, &re, &im)) goto done;
		envBuffer [syncBufferIndex] = hypotf (re, im);
		currentStrength += envBuffer [syncBufferIndex] - envBuffer [(syncBufferIndex - 50) & syncBufferMask];
		syncBufferIndex = (syncBufferIndex + 1) & syncBufferMask;
		counter ++;
		if (counter > T_null + 50) goto notSynced;
	}
SyncOnPhase:                                                            /* :344-381 */
	if (nframes >= max_frames) goto done;
	{
		orc_frame_info fi;
		memset (&fi, 0, sizeof (fi));
		fi. pos = s. pos; fi. phase0 = s. localPhase;
		if (!pump_samples (&s, ofdmBuffer, T_u, coarseCorrector + fineCorrector)) goto done;
		startIndex = k_find_index (dec, ofdmBuffer);
		if (startIndex < 0) goto notSynced;
		memmove (ofdmBuffer, &ofdmBuffer [2 * startIndex], sizeof (float) * 2 * (T_u - startIndex));
		const int ofdmBufferIndex = T_u - startIndex;
/* OFDM_PRS: :383-406 */
		if (!pump_samples (&s, &ofdmBuffer [2 * ofdmBufferIndex], T_u - ofdmBufferIndex,
		                   coarseCorrector + fineCorrector)) goto done;
		const int16_t correction = (int16_t) k_block0 (dec, ofdmBuffer, f2Correction);
		if (f2Correction) {
			if (correction == 0 && previous_1 == 0 && previous_2 == 0)
				f2Correction = 0;
			else if (correction != 100) {
				coarseCorrector += correction * p. carrierDiff;
				if (abs (coarseCorrector) > 35 * 1000) coarseCorrector = 0;
				previous_2 = previous_1;
				previous_1 = correction;
			}
		}
		fi. startIndex = startIndex; fi. correction = correction;
		fi. coarse = coarseCorrector; fi. fine = fineCorrector;
/* OFDM_SYMBOLS: :414-442.  Symbols 1..3 go to the FIC handler, 4..L-1 to the MSC handler; here both
 * land in sym[frame][symbol-1][2K]. */
		float fcRe = 0, fcIm = 0;
		int16_t *fsym = &sym [(size_t) nframes * (p. L - 1) * 2 * K];
		for (int n = 1; n < p. L; n ++) {
			if (!pump_samples (&s, ofdmBuffer, T_s, coarseCorrector + fineCorrector)) goto done;
			for (i = T_u; i < T_s; i ++) {              /* :424-425: FreqCorr += x[i] * conj (x[i - T_u]) */
				const float a = ofdmBuffer [2 * i], b = ofdmBuffer [2 * i + 1];
				const float c = ofdmBuffer [2 * (i - T_u)], d = ofdmBuffer [2 * (i - T_u) + 1];
				fcRe += a * c + b * d;
				fcIm += b * c - a * d;
			}
			k_token (dec, ofdmBuffer, &fsym [(size_t) (n - 1) * 2 * K]);
		}
		/* :445-446 fineCorrector (int16) += 0.1 * arg (FreqCorr) / M_PI * (carrierDiff / 2) */
		fineCorrector = (int16_t) ((double) fineCorrector +
		                           0.1 * atan2f (fcIm, fcRe) / M_PI * (p. carrierDiff / 2));
		fi. freqCorrRe = fcRe; fi. freqCorrIm = fcIm;
		info [nframes] = fi;
		nframes ++;
		syncBufferIndex = 0; currentStrength = 0;
		/* :453: skip the null symbol (the frame is complete even if the stream ends inside it) */
		if (!pump_samples (&s, ofdmBuffer, T_null, coarseCorrector + fineCorrector)) goto done;
		if (fineCorrector > p. carrierDiff / 2) {        /* :458-465 */
			coarseCorrector += p. carrierDiff; fineCorrector -= p. carrierDiff;
		} else if (fineCorrector < - p. carrierDiff / 2) {
			coarseCorrector -= p. carrierDiff; fineCorrector += p. carrierDiff;
		}
	}
	goto SyncOnPhase;
done:
	free (envBuffer); free (ofdmBuffer); k_ofdm_free (dec);
	return nframes;
}
