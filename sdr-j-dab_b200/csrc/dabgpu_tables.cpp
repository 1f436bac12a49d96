// dabgpu_tables.cpp -- host-side constant tables of the decode path: mode parameters, puncturing vectors,
// UEP/EEP protection profiles turned into depuncture LUTs, the energy-dispersal sequence.
// Citations are into the reference tree (AlbrechtL/sdr-j-dab 0.997).
#include "dabgpu_internal.h"

int dab_mode_params (int mode, DabParams *p) {          // gui.cpp:1328-1372; msc-handler.cpp:61-71
	memset (p, 0, sizeof (*p));
	struct row { int L, K, T_null, T_F, T_s, T_u, cd, bpc; };
	static const row rows [5] = { {0, 0, 0, 0, 0, 0, 0, 0},
		{76, 1536, 2656, 196608, 2552, 2048, 1000, 18}, {76, 384, 664, 49152, 638, 512, 4000, 72},
		{153, 192, 345, 49152, 319, 256, 2000, 18},     {76, 768, 1328, 98304, 1276, 1024, 2000, 36} };
	if (mode < 1 || mode > 4) return -1;
	const row &r = rows [mode];
	p -> dabMode = mode; p -> L = r. L; p -> K = r. K; p -> T_null = r. T_null; p -> T_F = r. T_F;
	p -> T_s = r. T_s; p -> T_u = r. T_u; p -> T_g = r. T_s - r. T_u; p -> carrierDiff = r. cd;
	p -> blocksPerCIF = r. bpc;
	p -> ficGroups = 3 * 2 * r. K / 2304;                // three FIC symbols (ofdm-processor.cpp:421-422)
	p -> cifsPerFrame = (r. L - 4) / r. bpc;
	return 0;
}

// PI_n, n = 1..24: 8+n ones per 32 bits (protTables.cpp:27-54).  Each 4-bit group starts as 1000; the n
// extra ones are dealt to the groups in the order 0,4,2,6,1,5,3,7, a second / third / fourth one per round.
static void puncture_vector (int n, uint8_t v [32]) {
	static const int order [8] = { 0, 4, 2, 6, 1, 5, 3, 7 };
	int ones [8] = { 1, 1, 1, 1, 1, 1, 1, 1 };
	for (int e = 0; e < n; e ++) ones [order [e & 7]] ++;
	for (int g = 0; g < 8; g ++)
		for (int b = 0; b < 4; b ++) v [4 * g + b] = b < ones [g];
}

// UEP profiles (deconvolve.cpp:39-113): {bitRate, level, L1..L4, PI1..PI4}; PI4 = 0 marks "unused".
static const int16_t uep_table [][10] = {
	{32,5,3,4,17,0,5,3,2,0},       {32,4,3,3,18,0,11,6,5,0},      {32,3,3,4,14,3,15,9,6,8},
	{32,2,3,4,14,3,22,13,8,13},    {32,1,3,5,13,3,24,17,12,17},   {48,5,4,3,26,3,5,4,2,3},
	{48,4,3,4,26,3,9,6,4,6},       {48,3,3,4,26,3,15,10,6,9},     {48,2,3,4,26,3,24,14,8,15},
	{48,1,3,5,25,3,24,18,13,18},   {64,5,6,9,31,2,5,3,2,3},       {64,4,6,9,33,0,11,6,6,0},
	{64,3,6,12,27,3,16,8,6,9},     {64,2,6,10,29,3,23,13,8,13},   {64,1,6,11,28,3,24,18,12,18},
	{80,5,6,10,41,3,6,3,2,3},      {80,4,6,10,41,3,11,6,5,6},     {80,3,6,11,40,3,16,8,6,7},
	{80,2,6,10,41,3,23,13,8,13},   {80,1,6,10,41,3,24,7,12,18},   // sic: PI2 = 7 as in the reference
	{96,5,7,9,53,3,5,4,2,4},       {96,4,7,10,52,3,9,6,4,6},      {96,3,6,12,51,3,16,9,6,10},
	{96,2,6,10,53,3,22,12,9,12},   {96,1,6,13,50,3,24,18,13,19},  {112,5,14,17,50,3,5,4,2,5},
	{112,4,11,21,49,3,9,6,4,8},    {112,3,11,23,47,3,16,8,6,9},   {112,2,11,21,49,3,23,12,9,14},
	{128,5,12,19,62,3,5,3,2,4},    {128,4,11,21,61,3,11,6,5,7},   {128,3,11,22,60,3,16,9,6,10},
	{128,2,11,21,61,3,22,12,9,14}, {128,1,11,20,62,3,24,17,13,19}, {160,5,11,19,87,3,5,4,2,4},
	{160,4,11,23,83,3,11,6,5,9},   {160,3,11,24,82,3,16,8,6,11},  {160,2,11,21,85,3,22,11,9,13},
	{160,1,11,22,84,3,24,18,12,19}, {192,5,11,20,110,3,6,4,2,5},  {192,4,11,22,108,3,10,6,4,9},
	{192,3,11,24,106,3,16,10,6,11}, {192,2,11,20,110,3,22,13,9,13}, {192,1,11,21,109,3,24,20,13,24},
	{224,5,12,22,131,3,8,6,2,6},   {224,4,12,26,127,3,12,8,4,11}, {224,3,11,20,134,3,16,10,7,9},
	{224,2,11,22,132,3,24,16,10,15}, {224,1,11,24,130,3,24,20,12,20}, {256,5,11,24,154,3,6,5,2,5},
	{256,4,11,24,154,3,12,9,5,10}, {256,3,11,27,151,3,16,10,7,10}, {256,2,11,22,156,3,24,14,10,13},
	{256,1,11,26,152,3,24,19,14,18}, {320,5,11,26,200,3,8,5,2,6}, {320,4,11,25,201,3,13,9,5,10},
	{320,2,11,26,200,3,24,17,9,17}, {384,5,11,27,247,3,8,6,2,7},  {384,3,11,24,250,3,16,9,7,10},
	{384,1,12,28,245,3,24,20,14,23},
};

static int segments_for (int bitRate, int uepFlag, int protLevel, int L [4], int PI [4]) {
	for (int i = 0; i < 4; i ++) L [i] = PI [i] = 0;
	if (uepFlag == 0) {                                 // dab-concurrent.cpp:78-83: 0 selects the UEP table
		for (size_t r = 0; r < sizeof (uep_table) / sizeof (uep_table [0]); r ++)
			if (uep_table [r][0] == bitRate && uep_table [r][1] == protLevel) {
				for (int i = 0; i < 4; i ++) { L [i] = uep_table [r][2 + i]; PI [i] = uep_table [r][6 + i]; }
				return 0;
			}
		return -1;
	}
	const int n = bitRate / 8, lvl = protLevel & 07;    // deconvolve.cpp:249-318
	if (protLevel & 0100) {
		switch (lvl) {
		   case 1: L [0] = 6 * n - 3; L [1] = 3; PI [0] = 24; PI [1] = 23; return 0;
		   case 2: if (bitRate == 8) { L [0] = 5; L [1] = 1; PI [0] = 13; PI [1] = 12; }
		           else { L [0] = 2 * n - 3; L [1] = 4 * n + 3; PI [0] = 14; PI [1] = 13; }
		           return 0;
		   case 3: L [0] = 6 * n - 3; L [1] = 3; PI [0] = 8; PI [1] = 7; return 0;
		   case 4: L [0] = 4 * n - 3; L [1] = 2 * n + 3; PI [0] = 3; PI [1] = 2; return 0;
		}
		return -1;
	}
	if (protLevel & 0200) {
		static const int pi1 [5] = { 0, 10, 6, 4, 2 };
		if (lvl < 1 || lvl > 4) return -1;
		L [0] = 24 * bitRate / 32 - 3; L [1] = 3; PI [0] = pi1 [lvl]; PI [1] = pi1 [lvl] - 1;
		return 0;
	}
	return -1;
}

// runs of (L_j blocks of 128 mother bits, PI_j) followed by the 24-bit tail punctured by 1100 x 6, placed
// at the running position exactly as the reference's counters do (deconvolve.cpp:186-231)
static int build_lut (int frameBits, const int L [4], const int PI [4], ProtProfile *pp) {
	const int total = 4 * (frameBits + 6);
	pp -> frameBits = frameBits;
	pp -> lut. assign (total, (uint16_t) 0xFFFF);
	int in = 0, pos = 0;
	for (int s = 0; s < 4; s ++) {
		if (L [s] <= 0) continue;
		uint8_t v [32];
		puncture_vector (PI [s], v);
		for (int i = 0; i < L [s] * 128; i ++, pos ++) {
			if (pos >= total) return -1;
			if (v [i & 31]) pp -> lut [pos] = (uint16_t) in ++;
		}
	}
	for (int i = 0; i < 24; i ++, pos ++) {
		if (pos >= total) return -1;
		if ((i & 3) < 2) pp -> lut [pos] = (uint16_t) in ++;
	}
	if (in >= 0xFFFF) return -1;
	pp -> nPunctured = in;
	prot_build_gather (pp);
	return 0;
}

// Tables of the throughput Viterbi (dabgpu_vit_simd.cu).  The trellis is processed in chunks of VS_CHUNK steps; for
// chunk k the input columns [a_k, a_k + 8 G_k) of every source row are staged in shared memory (a_k = the chunk's
// first input index rounded down to 8), and gather entry (step, j) is the byte offset of symbol j of that step
// relative to the code word's own tile row: (15 - D[i & 15]) rows up/down for the time de-interleaver
// (dab-concurrent.cpp:41-43, 162-169), column i - a_k; a punctured position points at the pad column, which holds
// the erasure value 127 (deconvolve.cpp:185).
void prot_build_gather (ProtProfile *pp) {
	static const int D [16] = { 15, 7, 11, 3, 13, 5, 9, 1, 14, 6, 10, 2, 12, 4, 8, 0 };
	const int nsteps = pp -> frameBits + 6, total = 4 * nsteps, in = pp -> nPunctured;
	const int nchunks = (nsteps + VS_CHUNK - 1) / VS_CHUNK;
	std::vector<int> first (nchunks + 1, in);           // first input index whose mother position is in chunk >= k
	for (int m = total - 1; m >= 0; m --)
		if (pp -> lut [m] != 0xFFFF) first [m / (4 * VS_CHUNK)] = pp -> lut [m];
	for (int k = nchunks - 1; k >= 0; k --) if (first [k] > first [k + 1]) first [k] = first [k + 1];
	pp -> chunk. assign (2 * nchunks, 0);
	for (int d = 0; d < 2; d ++) pp -> gather [d]. assign ((size_t) 4 * nchunks * VS_CHUNK + 8, 0);      // + two entries: the kernel's table prefetch runs two steps ahead
	for (int k = 0; k < nchunks; k ++) {
		const int a = first [k] & ~7, g = (first [k + 1] - a + 7) / 8;
		pp -> chunk [2 * k] = a; pp -> chunk [2 * k + 1] = first [k + 1] > first [k] ? g : 0;
		for (int s = 0; s < VS_CHUNK; s ++)
			for (int j = 0; j < 4; j ++) {
				const int m = 4 * (VS_CHUNK * k + s) + j;
				const int i = m < total && pp -> lut [m] != 0xFFFF ? pp -> lut [m] : -1;
				for (int d = 0; d < 2; d ++) {
					const int dmax = d ? 15 : 0;
					pp -> gather [d][(size_t) m] = i < 0 ? dmax * VS_PITCH + VS_PITCH - 1
					                                     : (dmax - (d ? D [i & 15] : 0)) * VS_PITCH + (i - a);
				}
			}
	}
}

int prot_build_identity (int frameBits, ProtProfile *pp) {  // viterbi::deconvolve on unpunctured input (viterbi.cpp:225-242)
	if (frameBits <= 0 || 4 * (frameBits + 6) >= 0xFFFF) return -1;
	const int total = 4 * (frameBits + 6);
	pp -> frameBits = frameBits; pp -> nPunctured = total;
	pp -> lut. resize (total);
	for (int m = 0; m < total; m ++) pp -> lut [m] = (uint16_t) m;
	prot_build_gather (pp);
	return 0;
}

int prot_build_fic (ProtProfile *pp) {                  // fic-handler.cpp:254-288
	const int L [4] = { 21, 3, 0, 0 }, PI [4] = { 16, 15, 0, 0 };
	return build_lut (768, L, PI, pp);
}

int prot_build_msc (int bitRate, int uepFlag, int protLevel, ProtProfile *pp) {
	int L [4], PI [4];
	if (bitRate <= 0 || bitRate > 384 || (bitRate % 8) != 0) return -1;
	if (segments_for (bitRate, uepFlag, protLevel, L, PI)) return -1;
	for (int s = 0; s < 4; s ++) if (L [s] < 0) return -1;
	return build_lut (24 * bitRate, L, PI, pp);
}

void prbs_packed (int nbits, std::vector<uint32_t> *words) {   // x^9 + x^5 + 1, all ones (fic-handler.cpp:100-108)
	words -> assign ((nbits + 31) / 32, 0u);
	unsigned sr = 0x1FF;
	for (int i = 0; i < nbits; i ++) {
		const unsigned b = ((sr >> 8) ^ (sr >> 4)) & 1;
		sr = ((sr << 1) | b) & 0x1FF;
		(*words) [i >> 5] |= b << (i & 31);
	}
}

// ".sdr" = RIFF/WAVE (wavfiles.cpp:44-75).  Header walk: "RIFF" size "WAVE", then chunks {id, size, payload padded to even};
// "fmt " gives the sample type, "data" the samples.  The reference's acceptance test is libsndfile's view of the file:
// 2 channels at 2 048 000 samples/s (wavfiles.cpp:66-71).
static uint32_t rd32 (const uint8_t *p) { return (uint32_t) p [0] | ((uint32_t) p [1] << 8) | ((uint32_t) p [2] << 16) | ((uint32_t) p [3] << 24); }
static uint32_t rd16 (const uint8_t *p) { return (uint32_t) p [0] | ((uint32_t) p [1] << 8); }

extern "C" int dabgpu_host_wav_parse (const void *file_image, size_t nbytes, dabgpu_wav_info *info) {
	if (!file_image || !info) return DABGPU_ERR_ARG;
	const uint8_t *b = (const uint8_t *) file_image;
	memset (info, 0, sizeof (*info));
	if (nbytes < 12 || memcmp (b, "RIFF", 4) != 0 || memcmp (b + 8, "WAVE", 4) != 0) return DABGPU_ERR_ARG;
	size_t pos = 12;
	bool have_fmt = false;
	while (pos + 8 <= nbytes) {
		const uint32_t size = rd32 (b + pos + 4);
		const uint8_t *body = b + pos + 8;
		if (memcmp (b + pos, "fmt ", 4) == 0) {
			if (size < 16 || pos + 8 + 16 > nbytes) return DABGPU_ERR_ARG;
			info -> format_tag = (int32_t) rd16 (body);
			info -> channels = (int32_t) rd16 (body + 2);
			info -> samplerate = (int32_t) rd32 (body + 4);
			info -> bits = (int32_t) rd16 (body + 14);
			if (info -> format_tag == 0xFFFE) {              // WAVE_FORMAT_EXTENSIBLE: the first two bytes of the sub-format GUID are the tag
				if (size < 40 || pos + 8 + 26 > nbytes) return DABGPU_ERR_ARG;
				info -> format_tag = (int32_t) rd16 (body + 24);
			}
			have_fmt = true;
		} else if (memcmp (b + pos, "data", 4) == 0) {
			if (!have_fmt) return DABGPU_ERR_ARG;
			if (info -> channels != 2 || info -> samplerate != DAB_INPUT_RATE) return DABGPU_ERR_ARG;   // wavfiles.cpp:66-71
			if (info -> format_tag == 3 && info -> bits == 32) info -> sample_format = 1;
			else if (info -> format_tag == 1 && info -> bits == 16) info -> sample_format = 2;
			else return DABGPU_ERR_ARG;
			const int64_t bps = info -> bits / 8 * 2;
			info -> data_offset = (int64_t) (pos + 8);
			int64_t announced = size;
			const int64_t present = (int64_t) nbytes - info -> data_offset;
			if (size == 0xFFFFFFFFu || size == 0) announced = present;        // streamed recordings leave the size open
			info -> nsamples_total = announced / bps;
			info -> nsamples = (announced < present ? announced : present) / bps;
			return DABGPU_OK;
		}
		pos += 8 + (size_t) size + (size & 1);
	}
	return DABGPU_ERR_ARG;
}
