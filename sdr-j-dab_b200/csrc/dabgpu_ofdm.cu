// dabgpu_ofdm.cu -- FFT + demod group, per-call parity entry points: common_fft / common_ifft (fft.cpp),
// phaseReference::findIndex (phasereference.cpp:60-88), ofdmDecoder::processBlock_0 / processToken
// (ofdm-decoder.cpp:85-207) and the constant tables they use (mapper.cpp, phasetable.cpp).
#include <math.h>
#include "dabgpu_ofdm.cuh"
#include "dabgpu_engine.h"
#include "dabgpu_fftp.cuh"

// ---------------------------------------------------------------------------------------------------
// host-side tables
// ---------------------------------------------------------------------------------------------------
// frequency de-interleaver (mapper.cpp:33-55, per-mode constants :82-105), wrapped like ofdm-decoder.cpp:179-181
static int build_perm (const DabParams &p, std::vector<int16_t> *perm) {
	int V1, lwb;
	switch (p. dabMode) {
	   case 2:  V1 = 127; lwb = 64;  break;
	   case 3:  V1 = 63;  lwb = 32;  break;
	   case 4:  V1 = 255; lwb = 128; break;
	   default: V1 = 511; lwb = 256; break;
	}
	const int upb = lwb + p. K;
	int pi = 0;
	perm -> clear ();
	for (int i = 0; i < p. T_u; i ++) {
		if (i > 0) pi = (13 * pi + V1) % p. T_u;
		if (pi == p. T_u / 2 || pi < lwb || pi > upb) continue;
		const int k = pi - p. T_u / 2;              // carrier number in [-K/2, K/2] \ {0}
		perm -> push_back ((int16_t) (k < 0 ? k + p. T_u : k));
	}
	return (int) perm -> size () == p. K ? 0 : -1;
}

// PRS phases (phasetable.cpp): per 32-carrier block the pair (i, n); phi = pi/2 * (h_i[k - kmin] + n).
// Blocks are listed from -K/2 upward, then from +1 upward.  Mode III has no table of its own in the
// reference and falls through to Mode I's (phasetable.cpp:123-139).
static const char *prs_blocks (int mode, int *nblk) {
	switch (mode) {
	   case 2: *nblk = 6;  return "021322320112" "201202312013";
	   case 4: *nblk = 12; return "001121320212203303112332" "003120120031221202312310";
	   default: *nblk = 24;
	      return "011220310312223302112233011223330212223101132132"
	             "033121110232211002322313003221130333231003302111";
	}
}
static float prs_phi (int mode, int k) {
	static const int8_t h [4][16] = { {0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1}, {0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0},
	                                  {0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3}, {0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2} };
	int nblk;
	const char *t = prs_blocks (mode, &nblk);
	int blk, off;
	if (k < 0) { const int d = k + 32 * nblk; if (d < 0) return 0; blk = d / 32; off = d % 32; }
	else if (k > 0) { const int d = k - 1; blk = nblk + d / 32; off = d % 32; if (d / 32 >= nblk) return 0; }
	else return 0;
	const int i = t [2 * blk] - '0', n = t [2 * blk + 1] - '0';
	return (float) (M_PI / 2 * (h [i][off & 15] + n));      // phasetable.cpp:180, DSPFLOAT result
}

int ofdm_tables_init (dabgpu *h, OfdmTables *T) {
	const DabParams &p = h -> p;
	const int N = p. T_u;
	std::vector<float2> tw (N), ref (N, make_float2 (0.f, 0.f)), hi (1000), lo (2048);
	for (int k = 0; k < N; k ++) {
		const double a = -2.0 * M_PI * (double) k / (double) N;
		tw [k] = make_float2 ((float) cos (a), (float) sin (a));
	}
	for (int i = 1; i <= p. K / 2; i ++) {                  // phasereference.cpp:41-47 (float cos/sin)
		float phi = prs_phi (p. dabMode, i);
		ref [i] = make_float2 (cosf (phi), sinf (phi));
		phi = prs_phi (p. dabMode, -i);
		ref [N - i] = make_float2 (cosf (phi), sinf (phi));
	}
	float refArg [18];
	for (int i = 0; i < 18; i ++) {                         // ofdm-decoder.cpp:73-76
		const float2 a = ref [i % N], b = ref [(i + 1) % N];
		refArg [i] = atan2f (a. y * b. x - a. x * b. y, a. x * b. x + a. y * b. y);
	}
	for (int i = 0; i < 1000; i ++) {                       // oscillator (ofdm-processor.cpp:76-81), factored
		const double a = 2.0 * M_PI * (double) i / 1000.0;
		hi [i] = make_float2 ((float) cos (a), (float) sin (a));
	}
	for (int i = 0; i < 2048; i ++) {
		const double a = 2.0 * M_PI * (double) i / (double) DAB_INPUT_RATE;
		lo [i] = make_float2 ((float) cos (a), (float) sin (a));
	}
	std::vector<int16_t> perm;
	if (build_perm (p, &perm)) return dab_fail (h, DABGPU_ERR_ARG, "frequency interleaver table has the wrong size");
	void *d;
	int rc;
	const long long base = 4ll << 40;
	if ((rc = dab_device_table (h, base + 1, tw. data (), N * sizeof (float2), &d))) return rc;   T -> tw = (const float2 *) d;
	if ((rc = dab_device_table (h, base + 2, ref. data (), N * sizeof (float2), &d))) return rc;  T -> ref = (const float2 *) d;
	if ((rc = dab_device_table (h, base + 3, perm. data (), p. K * sizeof (int16_t), &d))) return rc; T -> perm = (const int16_t *) d;
	if ((rc = dab_device_table (h, base + 4, refArg, sizeof (refArg), &d))) return rc;            T -> refArg = (const float *) d;
	if ((rc = dab_device_table (h, base + 5, hi. data (), 1000 * sizeof (float2), &d))) return rc; T -> osc_hi = (const float2 *) d;
	if ((rc = dab_device_table (h, base + 6, lo. data (), 2048 * sizeof (float2), &d))) return rc; T -> osc_lo = (const float2 *) d;
	T -> permpos = nullptr;
	if (p. dabMode == 1 || p. dabMode == 2 || p. dabMode == 4) {      // packed register-FFT kernels: carrier -> position in the transform's shared layout
		std::vector<uint16_t> pp (p. K);
		for (int i = 0; i < p. K; i ++) {
			const int pos = p. dabMode == 1 ? p_pos<1> (0, perm [i]) : p. dabMode == 4 ? p_pos<2> (0, perm [i]) : p_pos<4> (0, perm [i]);
			pp [i] = (uint16_t) r8_swz (pos);
		}
		if ((rc = dab_device_table (h, base + 7, pp. data (), p. K * sizeof (uint16_t), &d))) return rc;
		T -> permpos = (const uint16_t *) d;
	}
	T -> T_u = N; T -> T_s = p. T_s; T -> T_g = p. T_g; T -> K = p. K; T -> L = p. L;
	T -> log2n = 0; while ((1 << T -> log2n) < N) T -> log2n ++;
	T -> level = h -> cfg. threshold; T -> method = h -> cfg. freqSyncMethod;
	return DABGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-only table exports: let the CPU test-suite check the host logic without a GPU
// ---------------------------------------------------------------------------------------------------
extern "C" int dabgpu_host_mode_params (int32_t mode, int32_t out [12]) {
	DabParams p;
	if (!out || dab_mode_params (mode, &p)) return DABGPU_ERR_ARG;
	const int v [12] = { p. dabMode, p. L, p. K, p. T_null, p. T_F, p. T_s, p. T_u, p. T_g, p. carrierDiff, p. ficGroups, p. cifsPerFrame, p. blocksPerCIF };
	for (int i = 0; i < 12; i ++) out [i] = v [i];
	return DABGPU_OK;
}

extern "C" int dabgpu_host_perm_table (int32_t mode, int16_t *out) {
	DabParams p;
	std::vector<int16_t> perm;
	if (!out || dab_mode_params (mode, &p) || build_perm (p, &perm)) return DABGPU_ERR_ARG;
	memcpy (out, perm. data (), perm. size () * sizeof (int16_t));
	return DABGPU_OK;
}

extern "C" int dabgpu_host_ref_table (int32_t mode, float *out) {
	DabParams p;
	if (!out || dab_mode_params (mode, &p)) return DABGPU_ERR_ARG;
	memset (out, 0, sizeof (float) * 2 * p. T_u);
	for (int i = 1; i <= p. K / 2; i ++) {
		float phi = prs_phi (mode, i);
		out [2 * i] = cosf (phi); out [2 * i + 1] = sinf (phi);
		phi = prs_phi (mode, -i);
		out [2 * (p. T_u - i)] = cosf (phi); out [2 * (p. T_u - i) + 1] = sinf (phi);
	}
	return DABGPU_OK;
}

extern "C" int dabgpu_host_depuncture_lut (int32_t fic, int32_t bitRate, int32_t uepFlag, int32_t protLevel,
                                           int32_t *lut, int32_t lut_capacity, int32_t *lut_len, int32_t *n_punctured) {
	ProtProfile pp;
	if (!lut_len || !n_punctured) return DABGPU_ERR_ARG;
	if (fic ? prot_build_fic (&pp) : prot_build_msc (bitRate, uepFlag, protLevel, &pp)) return DABGPU_ERR_PROFILE;
	*lut_len = (int32_t) pp. lut. size (); *n_punctured = pp. nPunctured;
	if (lut) {
		if (lut_capacity < (int32_t) pp. lut. size ()) return DABGPU_ERR_ARG;
		for (size_t i = 0; i < pp. lut. size (); i ++) lut [i] = pp. lut [i] == 0xFFFF ? -1 : (int32_t) pp. lut [i];
	}
	return DABGPU_OK;
}

extern "C" int dabgpu_host_prbs (int32_t nbits, uint8_t *out) {
	if (!out || nbits < 0) return DABGPU_ERR_ARG;
	std::vector<uint32_t> w;
	prbs_packed (nbits, &w);
	for (int i = 0; i < nbits; i ++) out [i] = (w [i >> 5] >> (i & 31)) & 1;
	return DABGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-call kernels (float input, as the reference's class interfaces take it)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (OFDM_THREADS) fft_kernel (float2 *v, OfdmTables T, int inverse) {
	extern __shared__ __align__ (1024) float2 sm [];                   // 512-byte alignment: fft2048_r8
	const int N = T. T_u;
	float2 *g = v + (size_t) blockIdx. x * N;
	const float factor = (float) (1.0 / (float) N);
	if (N == R8_N) {                                     // the register FFT of the Mode I symbol kernel
		float2 x [8], tw1 [6];
		float2 *tw2 = sm + R8_SMEM, *tw3 = tw2 + R8_TW2;
		r8_fill_tables (tw2, tw3, T. tw);
		r8_load_tw1 (tw1, T. tw);
#pragma unroll
		for (int k = 0; k < 8; k ++) { x [k] = g [threadIdx. x + 256 * k]; if (inverse) x [k]. y = - x [k]. y; }
		__syncthreads ();
		fft2048_r8 (x, sm, tw1, tw2, tw3);
		for (int k = threadIdx. x; k < N; k += OFDM_THREADS) {
			float2 r = sm [r8_swz (r8_pos (k))];
			if (inverse) r = make_float2 (r. x * factor, (- r. y) * factor);
			g [k] = r;
		}
		return;
	}
	float2 *a = sm, *b = sm + N;
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) { float2 x = g [i]; if (inverse) x. y = - x. y; a [i] = x; }
	float2 *r = block_fft (a, b, N, T. tw);
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) {
		float2 x = r [i];
		if (inverse) x = make_float2 (x. x * factor, (- x. y) * factor);
		g [i] = x;
	}
}

__global__ void __launch_bounds__ (OFDM_THREADS) find_index_kernel (const float2 *v, OfdmTables T, int *idx) {
	extern __shared__ float2 sm [];
	const int N = T. T_u;
	float2 *a = sm, *b = sm + N;
	const float2 *g = v + (size_t) blockIdx. x * N;
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) a [i] = g [i];
	const int r = find_index_block (a, b, T);
	if (threadIdx. x == 0) idx [blockIdx. x] = r;
}

__global__ void __launch_bounds__ (OFDM_THREADS) block0_kernel (const float2 *v, OfdmTables T, float2 *phaseRef, int flag, int *correction) {
	extern __shared__ float2 sm [];
	__shared__ float cv [96];
	const int N = T. T_u;
	float2 *a = sm, *b = sm + N;
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) a [i] = v [i];
	float2 *f = block_fft (a, b, N, T. tw);
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) phaseRef [i] = f [i];       // ofdm-decoder.cpp:91
	int c = 0;
	if (flag) c = coarse_offset_warp0 (f, T, cv);
	if (threadIdx. x == 0) *correction = c;
}

__global__ void __launch_bounds__ (OFDM_THREADS) token_kernel (const float2 *inv, int nsym, OfdmTables T, float2 *phaseRef, int16_t *ibits) {
	extern __shared__ float2 sm [];
	const int N = T. T_u;
	float2 *a = sm, *b = sm + N, *prev = sm + 2 * N;
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) prev [i] = phaseRef [i];
	for (int s = 0; s < nsym; s ++) {
		const float2 *g = inv + (size_t) s * T. T_s + T. T_g;                          // ofdm-decoder.cpp:172
		__syncthreads ();
		for (int i = threadIdx. x; i < N; i += OFDM_THREADS) a [i] = g [i];
		float2 *f = block_fft (a, b, N, T. tw);
		demod_symbol (f, prev, T, ibits + (size_t) s * 2 * T. K);
	}
	__syncthreads ();
	for (int i = threadIdx. x; i < N; i += OFDM_THREADS) phaseRef [i] = prev [i];
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
static int upload (dabgpu *h, const void *host, size_t bytes) {
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, h -> d_in. ensure (bytes));
	CUDA_TRY (h, h -> h_in. ensure (bytes));
	memcpy (h -> h_in. p, host, bytes);
	CUDA_TRY (h, cudaMemcpyAsync (h -> d_in. p, h -> h_in. p, bytes, cudaMemcpyHostToDevice, h -> stream));
	return DABGPU_OK;
}
static int download (dabgpu *h, void *host, const void *dev, size_t bytes) {
	CUDA_TRY (h, h -> h_out. ensure (bytes));
	CUDA_TRY (h, cudaMemcpyAsync (h -> h_out. p, dev, bytes, cudaMemcpyDeviceToHost, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	memcpy (host, h -> h_out. p, bytes);
	return DABGPU_OK;
}

extern "C" int dabgpu_fft (dabgpu_t *h, float *v, int32_t n, int32_t inverse) {
	if (!h || !v || n < 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_fft: bad argument");
	if (n == 0) return DABGPU_OK;
	Engine *E = h -> engine;
	const size_t bytes = (size_t) n * h -> p. T_u * sizeof (float2);
	int rc = upload (h, v, bytes);
	if (rc) return rc;
	fft_kernel<<<n, OFDM_THREADS, (2 * h -> p. T_u + 512) * sizeof (float2), h -> stream>>> ((float2 *) h -> d_in. p, E -> T, inverse);
	h -> launches ++;
	CUDA_TRY (h, cudaGetLastError ());
	return download (h, v, h -> d_in. p, bytes);
}

extern "C" int dabgpu_find_index (dabgpu_t *h, const float *v, int32_t n, int32_t *idx) {
	if (!h || !v || !idx || n < 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_find_index: bad argument");
	if (n == 0) return DABGPU_OK;
	Engine *E = h -> engine;
	int rc = upload (h, v, (size_t) n * h -> p. T_u * sizeof (float2));
	if (rc) return rc;
	CUDA_TRY (h, h -> d_out. ensure ((size_t) n * sizeof (int)));
	find_index_kernel<<<n, OFDM_THREADS, 2 * h -> p. T_u * sizeof (float2), h -> stream>>> ((const float2 *) h -> d_in. p, E -> T, (int *) h -> d_out. p);
	h -> launches ++;
	CUDA_TRY (h, cudaGetLastError ());
	return download (h, idx, h -> d_out. p, (size_t) n * sizeof (int));
}

extern "C" int dabgpu_block0 (dabgpu_t *h, const float *v, int32_t flag, int16_t *correction) {
	if (!h || !v || !correction) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_block0: bad argument");
	Engine *E = h -> engine;
	int rc = upload (h, v, (size_t) h -> p. T_u * sizeof (float2));
	if (rc) return rc;
	CUDA_TRY (h, h -> d_out. ensure (sizeof (int)));
	block0_kernel<<<1, OFDM_THREADS, 2 * h -> p. T_u * sizeof (float2), h -> stream>>> ((const float2 *) h -> d_in. p, E -> T, E -> d_phaseRef, flag, (int *) h -> d_out. p);
	h -> launches ++;
	CUDA_TRY (h, cudaGetLastError ());
	int c = 0;
	if ((rc = download (h, &c, h -> d_out. p, sizeof (int)))) return rc;
	*correction = (int16_t) c;
	E -> have_phase_ref = true;
	return DABGPU_OK;
}

extern "C" int dabgpu_token (dabgpu_t *h, const float *inv, int32_t nsym, int16_t *ibits) {
	if (!h || !inv || !ibits || nsym < 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_token: bad argument");
	Engine *E = h -> engine;
	if (!E -> have_phase_ref) return dab_fail (h, DABGPU_ERR_STATE, "dabgpu_token before dabgpu_block0: no phase reference yet");
	if (nsym == 0) return DABGPU_OK;
	int rc = upload (h, inv, (size_t) nsym * h -> p. T_s * sizeof (float2));
	if (rc) return rc;
	const size_t obytes = (size_t) nsym * 2 * h -> p. K * sizeof (int16_t);
	CUDA_TRY (h, h -> d_out. ensure (obytes));
	token_kernel<<<1, OFDM_THREADS, 3 * h -> p. T_u * sizeof (float2), h -> stream>>> ((const float2 *) h -> d_in. p, nsym, E -> T, E -> d_phaseRef, (int16_t *) h -> d_out. p);
	h -> launches ++;
	CUDA_TRY (h, cudaGetLastError ());
	return download (h, ibits, h -> d_out. p, obytes);
}

extern "C" int dabgpu_get_phase_reference (dabgpu_t *h, float *out) {
	if (!h || !out) return DABGPU_ERR_ARG;
	return download (h, out, h -> engine -> d_phaseRef, (size_t) h -> p. T_u * sizeof (float2));
}
