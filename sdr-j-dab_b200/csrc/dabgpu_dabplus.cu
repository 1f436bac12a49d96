// dabgpu_dabplus.cu -- DAB+ super-frame layer on the device, the stage right behind the MSC Viterbi for
// every DAB+ service: mp4Processor::addtoFrame + the Fire-code / Reed-Solomon / access-unit-table part of
// processSuperframe (mp4processor.cpp:107-150, 152-275), firecode_checker::check (firecode-checker.cpp:77-95),
// reedSolomon (8, 0435, 0, 1, 10)::dec with cutlen 135 (reed-solomon.cpp:129-399, galois.cpp), dabPlus_crc
// (mp4processor.cpp:40-61).  The AAC decoder is out of scope; what leaves here is the repaired super frame, its
// access-unit table and the per-AU CRC verdicts.
//
// The reference slides a 5-CIF window one CIF at a time until the Fire code AND the repair succeed, then jumps five
// CIFs.  Which windows are looked at depends on earlier verdicts, but a verdict itself depends only on its window,
// so: (1) every CIF is packed to bytes and tested as a potential super-frame start in parallel (Fire code over its
// first 11 bytes); (2) every candidate that passes is repaired speculatively, one thread per RS code word
// (bitRate/8 per super frame, column interleaved); (3) one thread per candidate builds the verdict (all columns
// decodable, access-unit table sane) and the AU CRCs; (4) a single thread replays the reference's counter logic over
// the verdicts and picks the windows the reference would have processed; (5) those are gathered in order.
// All arithmetic is integer: bit-exact.
#include "dabgpu_internal.h"

struct DabPlusTables { uint8_t alpha_to [256], index_of [256]; uint16_t fire [256]; };

static void dabplus_host_tables (DabPlusTables *t) {
	// galois (8, 0435): galois.cpp:39-62
	int sr = 1;
	t -> index_of [0] = 255; t -> alpha_to [255] = 0;
	for (int i = 0; i < 255; i ++) {
		t -> index_of [sr] = (uint8_t) i; t -> alpha_to [i] = (uint8_t) sr;
		sr <<= 1;
		if (sr & 256) sr ^= 0435;
		sr &= 255;
	}
	// firecode_checker::firecode_checker, firecode-checker.cpp:33-75: g(x) = (x^11+1)(x^5+x^3+x^2+x+1)
	static const uint8_t g [16] = { 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0 };
	uint16_t itab [8];
	for (int b = 0; b < 8; b ++) {
		uint8_t regs [16];
		memset (regs, 0, 16);
		regs [8 + b] = 1;
		for (int i = 0; i < 8; i ++) {
			const uint8_t z = regs [15];
			for (int j = 15; j > 0; j --) regs [j] = regs [j - 1] ^ (z & g [j]);
			regs [0] = z;
		}
		uint16_t v = 0;
		for (int i = 15; i >= 0; i --) v = (uint16_t) ((v << 1) | regs [i]);
		itab [b] = v;
	}
	for (int i = 0; i < 256; i ++) {
		t -> fire [i] = 0;
		for (int j = 0; j < 8; j ++) if (i & (1 << j)) t -> fire [i] ^= itab [j];
	}
}

// ---- (1) pack one CIF (24*bitRate one-bit bytes, MSB first, mp4processor.cpp:114-119) and test it as a super-frame start
__global__ void __launch_bounds__ (128) dabp_pack_kernel (const uint8_t *bits, int nbytes, uint8_t *rows, uint8_t *fire, const DabPlusTables *T) {
	const uint8_t *V = bits + (size_t) blockIdx. x * nbytes * 8;
	uint8_t *row = rows + (size_t) blockIdx. x * nbytes;
	for (int i = threadIdx. x; i < nbytes; i += blockDim. x) {
		const uint2 w = *reinterpret_cast<const uint2 *> (V + 8 * i);                 // 8 bits = 8 bytes, 8-byte aligned (nbits is a multiple of 8)
		const uint32_t hi = ((w. x & 0x01010101u) * 0x08040201u) >> 24, lo = ((w. y & 0x01010101u) * 0x08040201u) >> 24;
		row [i] = (uint8_t) ((hi << 4) | (lo & 15u));
	}
	__syncthreads ();
	if (threadIdx. x == 0) {                                                       // firecode_checker::check
		const uint8_t *x = row;
		unsigned state = ((unsigned) x [2] << 8) | x [3], istate;
		for (int i = 4; i < 11; i ++) {
			istate = T -> fire [state >> 8];
			state = (((istate & 0x00ffu) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00u)) & 0xffffu;
		}
		for (int i = 0; i < 2; i ++) {
			istate = T -> fire [state >> 8];
			state = (((istate & 0x00ffu) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00u)) & 0xffffu;
		}
		fire [blockIdx. x] = state == 0;
	}
}

// Fire-code flags of already packed rows (the history rows at the head of a call)
__global__ void dabp_fire_rows_kernel (const uint8_t *rows, int nrows, int rowbytes, uint8_t *fire, const DabPlusTables *T) {
	const int r = threadIdx. x;
	if (r >= nrows) return;
	const uint8_t *x = rows + (size_t) r * rowbytes;
	unsigned state = ((unsigned) x [2] << 8) | x [3], istate;
	for (int i = 4; i < 11; i ++) {
		istate = T -> fire [state >> 8];
		state = (((istate & 0x00ffu) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00u)) & 0xffffu;
	}
	for (int i = 0; i < 2; i ++) {
		istate = T -> fire [state >> 8];
		state = (((istate & 0x00ffu) ^ x [i]) | ((istate ^ (state << 8)) & 0xff00u)) & 0xffffu;
	}
	fire [r] = state == 0;
}

// ---- (2) reedSolomon::dec (rsIn, rsOut, 135) for column j of the super frame starting at row s ----
#define RS_NN 255
#define RS_NROOTS 10
#define RS_PAD 135
__device__ __forceinline__ int rs_modnn (int x) { while (x >= RS_NN) { x -= RS_NN; x = (x >> 8) + (x & RS_NN); } return x; }   // galois.cpp:64-70

__global__ void __launch_bounds__ (64) dabp_rs_kernel (const uint8_t *rows, int nrows, int R, const uint8_t *fire, const DabPlusTables *T,
                                                      uint8_t *sfbuf, int8_t *ler) {
	__shared__ uint8_t A [256], I [256];                                          // alpha_to, index_of
	for (int i = threadIdx. x; i < 256; i += blockDim. x) { A [i] = T -> alpha_to [i]; I [i] = T -> index_of [i]; }
	__syncthreads ();
	const int idx = blockIdx. x * blockDim. x + threadIdx. x;
	const int s = idx / R, j = idx % R;
	if (s + 5 > nrows || !fire [s]) return;
	const uint8_t *in = rows + (size_t) s * 24 * R + j;                            // byte k of the code word: in [k * R]
	uint8_t r [120];
	for (int k = 0; k < 120; k ++) r [k] = in [(size_t) k * R];
	// syndromes by Horner at alpha^0..alpha^9 (reed-solomon.cpp:236-270); the 135 leading zeros of the shortened word change nothing
	uint8_t syn [RS_NROOTS + 1];
	{
		int sy [RS_NROOTS];
		for (int i = 0; i < RS_NROOTS; i ++) sy [i] = r [0];
		for (int k = 1; k < 120; k ++)
			for (int i = 0; i < RS_NROOTS; i ++) sy [i] = sy [i] == 0 ? r [k] : r [k] ^ A [rs_modnn (I [sy [i]] + i)];
		int any = 0;
		for (int i = 0; i < RS_NROOTS; i ++) { syn [i] = (uint8_t) sy [i]; any |= sy [i]; }
		syn [RS_NROOTS] = 0;
		uint8_t *out = sfbuf + (size_t) s * 110 * R + j;
		if (!any) {                                                                // no error
			for (int k = 0; k < 110; k ++) out [(size_t) k * R] = r [k];
			ler [idx] = 0;
			return;
		}
	}
	int ret;
	do {
		uint8_t Lambda [RS_NROOTS + 1], Corr [RS_NROOTS + 1], omega [RS_NROOTS + 1], rootTable [RS_NROOTS], locTable [RS_NROOTS];
#define GMUL(a, b) (((a) == 0 || (b) == 0) ? 0 : (int) A [rs_modnn (I [a] + I [b])])
#define GDIV(a, b) ((a) == 0 ? 0 : (int) A [rs_modnn (255 + I [a] - I [b])])
		// Berlekamp-Massey (:275-325)
		int K = 1, L = 0, deg_lambda = 0, error = syn [0];
		for (int i = 0; i <= RS_NROOTS; i ++) Lambda [i] = Corr [i] = 0;
		Lambda [0] = 1; Corr [1] = 1;
		while (K <= RS_NROOTS) {
			uint8_t old [RS_NROOTS + 1];
			for (int i = 0; i <= RS_NROOTS; i ++) old [i] = Lambda [i];
			for (int i = 0; i <= RS_NROOTS; i ++) Lambda [i] ^= (uint8_t) GMUL (error, Corr [i]);
			if (2 * L < K && error != 0) {
				L = K - L;
				for (int i = 0; i <= RS_NROOTS; i ++) Corr [i] = (uint8_t) GDIV (old [i], error);
			}
			for (int i = RS_NROOTS; i >= 1; i --) Corr [i] = Corr [i - 1];
			Corr [0] = 0;
			error = syn [K];
			for (int i = 1; i <= K; i ++) error ^= GMUL (syn [K - i], Lambda [i]);
			K ++;
		}
		for (int i = 0; i <= RS_NROOTS; i ++) {
			if (Lambda [i] != 0) deg_lambda = i;
			Lambda [i] = I [Lambda [i]];
		}
		// Chien search (:330-362), iprim = 1: location k = i - 1
		int rootCount = 0;
		{
			uint8_t reg [RS_NROOTS + 1];
			for (int i = 0; i <= RS_NROOTS; i ++) reg [i] = Lambda [i];
			for (int i = 1; i <= RS_NN; i ++) {
				int result = 1;
				for (int q = deg_lambda; q > 0; q --)
					if (reg [q] != RS_NN) { reg [q] = (uint8_t) rs_modnn (reg [q] + q); result ^= A [reg [q]]; }
				if (result != 0) continue;
				if (rootCount < RS_NROOTS) { rootTable [rootCount] = (uint8_t) i; locTable [rootCount] = (uint8_t) (i - 1); }
				rootCount ++;
			}
		}
		if (rootCount != deg_lambda) { ret = -1; break; }
		// omega = syndromes * lambda mod x^nroots (:372-399)
		int deg_omega = 0;
		for (int i = 0; i < RS_NROOTS; i ++) {
			int tmp = 0;
			for (int q = deg_lambda < i ? deg_lambda : i; q >= 0; q --)
				if (I [syn [i - q]] != RS_NN && Lambda [q] != RS_NN) tmp ^= A [rs_modnn (I [syn [i - q]] + Lambda [q])];
			if (tmp != 0) deg_omega = i;
			omega [i] = I [tmp];
		}
		// Forney (:171-231); pow_power (a, n) = a == 0 ? 0 : a n mod 255 (galois.cpp:124-126)
		ret = rootCount;
		for (int q = rootCount - 1; q >= 0; q --) {
			const int root = rootTable [q];
			int num1 = 0, den = 0;
			for (int i = deg_omega; i >= 0; i --)
				if (omega [i] != RS_NN) num1 ^= A [rs_modnn (omega [i] + (i == 0 ? 0 : (i * root) % 255))];
			const int num2 = A [rs_modnn ((root * 254) % 255 + RS_NN)];
			for (int i = (deg_lambda < RS_NROOTS - 1 ? deg_lambda : RS_NROOTS - 1) & ~1; i >= 0; i -= 2)
				if (Lambda [i + 1] != RS_NN) den ^= A [rs_modnn (Lambda [i + 1] + (i == 0 ? 0 : (i * root) % 255))];
			if (den == 0) { ret = -1; break; }
			if (num1 != 0) {
				if (locTable [q] >= RS_NN - RS_NROOTS) ret --;
				else {
					int y = rs_modnn (I [num1] + I [num2]);
					y = rs_modnn (y + RS_NN - I [den]);
					if (locTable [q] >= RS_PAD) r [locTable [q] - RS_PAD] ^= A [y];     // (a "repair" inside the zero padding is lost, as in the reference)
				}
			}
		}
#undef GMUL
#undef GDIV
	} while (0);
	if (ret >= 0) {
		uint8_t *out = sfbuf + (size_t) s * 110 * R + j;
		for (int k = 0; k < 110; k ++) out [(size_t) k * R] = r [k];
	}
	ler [idx] = (int8_t) ret;
}

// ---- (3) verdict of processSuperframe for the candidate starting at row s (mp4processor.cpp:170-275) ----
__global__ void dabp_verdict_kernel (const uint8_t *sfbuf, const int8_t *ler, const uint8_t *fire, int nrows, int R,
                                     uint8_t *good, dabgpu_superframe_info *info) {
	const int s = blockIdx. x * blockDim. x + threadIdx. x;
	if (s >= nrows) return;
	good [s] = 0;
	if (s + 5 > nrows || !fire [s]) return;
	int nErrors = 0;
	for (int j = 0; j < R; j ++) {                                                 // the reference stops at the first column that fails
		const int l = ler [s * R + j];
		if (l < 0) return;
		nErrors += l;
	}
	const uint8_t *o = sfbuf + (size_t) s * 110 * R;
	const int dacRate = (o [2] >> 6) & 1, sbrFlag = (o [2] >> 5) & 1;
	int au [7] = { 0, 0, 0, 0, 0, 0, 0 }, n;
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
		if (au [i + 1] < au [i]) return;
		const int len = au [i + 1] - au [i] - 2;
		if (len >= 960 || len < 0) return;
		unsigned acc = 0xFFFFu;                                                    // dabPlus_crc, :40-61
		const uint8_t *msg = o + au [i];
		for (int b = 0; b < len; b ++) {
			unsigned data = (unsigned) msg [b] << 8;
			for (int q = 8; q > 0; q --) {
				acc = ((data ^ acc) & 0x8000u) ? ((acc << 1) ^ 0x1021u) & 0xFFFFu : (acc << 1) & 0xFFFFu;
				data = (data << 1) & 0xFFFFu;
			}
		}
		const unsigned crc = ~(((unsigned) msg [len] << 8) | msg [len + 1]) & 0xFFFFu;
		if ((crc ^ acc) == 0) crcmask |= 1 << i;
	}
	dabgpu_superframe_info fi;
	fi. first_cif = 0; fi. corrected = nErrors; fi. num_aus = n; fi. au_crc = crcmask;
	for (int i = 0; i < 7; i ++) fi. au_start [i] = i <= n ? au [i] : 0;
	info [s] = fi;
	good [s] = 1;
}

// ---- (4) the counter logic of addtoFrame (:121-149) replayed over the verdicts; rows [H, nrows) are this call's CIFs ----
struct DabPlusCtl { int cnt, nsf; };
__global__ void dabp_scan_kernel (const uint8_t *good, int H, int nrows, DabPlusCtl *ctl, int *accepted) {
	int cnt = ctl -> cnt, nsf = 0;
	for (int k = H; k < nrows; k ++) {
		cnt ++;
		if (cnt >= 5) {
			if (good [k - 4]) { accepted [nsf ++] = k - 4; cnt = 0; }
			else cnt = 4;
		}
	}
	ctl -> cnt = cnt; ctl -> nsf = nsf;
}

// ---- (5) the accepted super frames, in order ----
__global__ void __launch_bounds__ (128) dabp_gather_kernel (const uint8_t *sfbuf, const dabgpu_superframe_info *info, const DabPlusCtl *ctl, const int *accepted,
                                                          int R, long long cif0, uint8_t *out, dabgpu_superframe_info *out_info) {
	const int i = blockIdx. x;
	if (i >= ctl -> nsf) return;
	const int s = accepted [i];
	const uint8_t *src = sfbuf + (size_t) s * 110 * R;
	uint8_t *dst = out + (size_t) i * 110 * R;
	for (int b = threadIdx. x; b < 110 * R; b += blockDim. x) dst [b] = src [b];
	if (threadIdx. x == 0) { dabgpu_superframe_info fi = info [s]; fi. first_cif = cif0 + s; out_info [i] = fi; }
}

struct dabgpu_dabplus {
	dabgpu *h;
	int bitRate, R, rowbytes;
	int hist_rows, cnt;                       // rows carried over from earlier calls (<= 4), blocksInBuffer
	long long cifs;                           // CIFs seen so far
	DevBuf rows, hist, fire, ler, sfbuf, good, info, ctl, accepted, out, out_info, tables, bits;
	PinBuf h_ctl;
};

extern "C" int dabgpu_dabplus_create (dabgpu_t *h, int32_t bitRate, dabgpu_dabplus_t **out) {
	if (!h || !out || bitRate < 8 || bitRate > 384 || (bitRate % 8) != 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_dabplus_create: bad bit rate %d", bitRate);
	*out = nullptr;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	dabgpu_dabplus *d = new dabgpu_dabplus ();
	d -> h = h; d -> bitRate = bitRate; d -> R = bitRate / 8; d -> rowbytes = 3 * bitRate;
	d -> hist_rows = 0; d -> cnt = 0; d -> cifs = 0;
	DabPlusTables t;
	dabplus_host_tables (&t);
	cudaError_t e = d -> tables. ensure (sizeof (t));
	if (e == cudaSuccess) e = d -> hist. ensure ((size_t) 4 * d -> rowbytes);
	if (e == cudaSuccess) e = d -> ctl. ensure (sizeof (DabPlusCtl));
	if (e == cudaSuccess) e = d -> h_ctl. ensure (sizeof (DabPlusCtl));
	if (e == cudaSuccess) e = cudaMemcpy (d -> tables. p, &t, sizeof (t), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { dabgpu_dabplus_destroy (d); return dab_fail (h, DABGPU_ERR_CUDA, "dabplus alloc: %s", cudaGetErrorString (e)); }
	*out = d;
	return DABGPU_OK;
}

extern "C" void dabgpu_dabplus_destroy (dabgpu_dabplus_t *d) {
	if (!d) return;
	cudaSetDevice (d -> h -> device);
	cudaStreamSynchronize (d -> h -> stream);
	for (DevBuf *b : { &d -> rows, &d -> hist, &d -> fire, &d -> ler, &d -> sfbuf, &d -> good, &d -> info, &d -> ctl, &d -> accepted, &d -> out, &d -> out_info, &d -> tables, &d -> bits })
		b -> release ();
	d -> h_ctl. release ();
	delete d;
}

extern "C" int dabgpu_dabplus_process_dev (dabgpu_dabplus_t *d, const uint8_t *d_bits, int32_t ncif, uint8_t *superframes,
                                           dabgpu_superframe_info *info, int32_t max_sf, int32_t *nsf) {
	if (!d || !nsf || ncif < 0 || max_sf < 0 || (ncif > 0 && !d_bits)) return dab_fail (d ? d -> h : nullptr, DABGPU_ERR_ARG, "dabgpu_dabplus_process: bad argument");
	dabgpu *h = d -> h;
	*nsf = 0;
	if (ncif == 0) return DABGPU_OK;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	cudaStream_t st = h -> stream;
	const int H = d -> hist_rows, n = H + ncif, R = d -> R, rb = d -> rowbytes;
	const int maxsf = n / 5 + 1;
	CUDA_TRY (h, d -> rows. ensure ((size_t) n * rb));
	CUDA_TRY (h, d -> fire. ensure (n)); CUDA_TRY (h, d -> good. ensure (n));
	CUDA_TRY (h, d -> ler. ensure ((size_t) n * R));
	CUDA_TRY (h, d -> sfbuf. ensure ((size_t) n * 110 * R));
	CUDA_TRY (h, d -> info. ensure ((size_t) n * sizeof (dabgpu_superframe_info)));
	CUDA_TRY (h, d -> accepted. ensure ((size_t) maxsf * sizeof (int)));
	CUDA_TRY (h, d -> out. ensure ((size_t) maxsf * 110 * R));
	CUDA_TRY (h, d -> out_info. ensure ((size_t) maxsf * sizeof (dabgpu_superframe_info)));
	uint8_t *rows = (uint8_t *) d -> rows. p, *fire = (uint8_t *) d -> fire. p;
	const DabPlusTables *T = (const DabPlusTables *) d -> tables. p;
	if (H) CUDA_TRY (h, cudaMemcpyAsync (rows, d -> hist. p, (size_t) H * rb, cudaMemcpyDeviceToDevice, st));
	dabp_pack_kernel<<<ncif, 128, 0, st>>> (d_bits, rb, rows + (size_t) H * rb, fire + H, T);
	DabPlusCtl *hc = (DabPlusCtl *) d -> h_ctl. p;
	hc -> cnt = d -> cnt; hc -> nsf = 0;
	CUDA_TRY (h, cudaMemcpyAsync (d -> ctl. p, hc, sizeof (DabPlusCtl), cudaMemcpyHostToDevice, st));
	if (H) dabp_fire_rows_kernel<<<1, 32, 0, st>>> (rows, H, rb, fire, T);   // the history rows are candidates again: their windows are complete now
	const int nthreads = (n - 4 > 0 ? n - 4 : 0) * R;
	if (nthreads > 0) dabp_rs_kernel<<<(nthreads + 63) / 64, 64, 0, st>>> (rows, n, R, fire, T, (uint8_t *) d -> sfbuf. p, (int8_t *) d -> ler. p);
	dabp_verdict_kernel<<<(n + 63) / 64, 64, 0, st>>> ((const uint8_t *) d -> sfbuf. p, (const int8_t *) d -> ler. p, fire, n, R, (uint8_t *) d -> good. p,
	                                                  (dabgpu_superframe_info *) d -> info. p);
	dabp_scan_kernel<<<1, 1, 0, st>>> ((const uint8_t *) d -> good. p, H, n, (DabPlusCtl *) d -> ctl. p, (int *) d -> accepted. p);
	dabp_gather_kernel<<<maxsf, 128, 0, st>>> ((const uint8_t *) d -> sfbuf. p, (const dabgpu_superframe_info *) d -> info. p, (const DabPlusCtl *) d -> ctl. p,
	                                          (const int *) d -> accepted. p, R, d -> cifs - H, (uint8_t *) d -> out. p, (dabgpu_superframe_info *) d -> out_info. p);
	h -> launches += 5 + (H ? 1 : 0);
	CUDA_TRY (h, cudaGetLastError ());
	CUDA_TRY (h, cudaMemcpyAsync (hc, d -> ctl. p, sizeof (DabPlusCtl), cudaMemcpyDeviceToHost, st));
	// history for the next call: the last (up to) 4 rows
	const int keep = n < 4 ? n : 4;
	CUDA_TRY (h, cudaMemcpyAsync (d -> hist. p, rows + (size_t) (n - keep) * rb, (size_t) keep * rb, cudaMemcpyDeviceToDevice, st));
	CUDA_TRY (h, cudaStreamSynchronize (st));
	d -> cnt = hc -> cnt; d -> hist_rows = keep; d -> cifs += ncif;
	const int got = hc -> nsf;
	*nsf = got;
	const int take = got < max_sf ? got : max_sf;
	if (take > 0) {
		if (superframes) CUDA_TRY (h, cudaMemcpyAsync (superframes, d -> out. p, (size_t) take * 110 * R, cudaMemcpyDeviceToHost, st));
		if (info) CUDA_TRY (h, cudaMemcpyAsync (info, d -> out_info. p, (size_t) take * sizeof (dabgpu_superframe_info), cudaMemcpyDeviceToHost, st));
		CUDA_TRY (h, cudaStreamSynchronize (st));
	}
	return DABGPU_OK;
}

extern "C" int dabgpu_dabplus_process (dabgpu_dabplus_t *d, const uint8_t *bits, int32_t ncif, uint8_t *superframes,
                                       dabgpu_superframe_info *info, int32_t max_sf, int32_t *nsf) {
	if (!d || !nsf || ncif < 0 || (ncif > 0 && !bits)) return dab_fail (d ? d -> h : nullptr, DABGPU_ERR_ARG, "dabgpu_dabplus_process: bad argument");
	dabgpu *h = d -> h;
	if (ncif == 0) { *nsf = 0; return DABGPU_OK; }
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const size_t nb = (size_t) ncif * 24 * d -> bitRate;
	CUDA_TRY (h, d -> bits. ensure (nb));
	CUDA_TRY (h, h -> h_in. ensure (nb));
	memcpy (h -> h_in. p, bits, nb);
	CUDA_TRY (h, cudaMemcpyAsync (d -> bits. p, h -> h_in. p, nb, cudaMemcpyHostToDevice, h -> stream));
	return dabgpu_dabplus_process_dev (d, (const uint8_t *) d -> bits. p, ncif, superframes, info, max_sf, nsf);
}
