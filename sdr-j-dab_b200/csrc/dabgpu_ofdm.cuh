// dabgpu_ofdm.cuh -- device building blocks of the FFT + demod group (shared by the per-call parity entry
// points in dabgpu_ofdm.cu and the stream engine in dabgpu_engine.cu).  sm_100a.
#pragma once
#include "dabgpu_internal.h"

#define OFDM_THREADS 256

struct OfdmTables {                   // device pointers, built once per handle (dab_engine_init)
	const float2  *tw;                // [T_u]  exp (-2 pi i k / T_u), double -> float
	const float2  *ref;               // [T_u]  PRS spectrum (phasereference.cpp:41-47)
	const int16_t *perm;              // [K]    frequency de-interleaver, wrapped to [0, T_u) (mapper.cpp, ofdm-decoder.cpp:179-181)
	const uint16_t *permpos;          // [K]    Mode I only: position of that carrier in the register FFT's shared layout
	const float   *refArg;            // [18]   ofdm-decoder.cpp:73-76
	const float2  *osc_hi;            // [1000] exp (2 pi i 2048 h / 2048000)   NCO = osc_hi[lp >> 11] * osc_lo[lp & 2047]
	const float2  *osc_lo;            // [2048] exp (2 pi i l / 2048000)        (ofdm-processor.cpp:76-81, 165-167)
	int T_u, T_s, T_g, K, L, log2n;
	int level, method;
};

__device__ __forceinline__ float2 cmul (float2 a, float2 b) { return make_float2 (a. x * b. x - a. y * b. y, a. x * b. y + a. y * b. x); }
__device__ __forceinline__ float2 cmulc (float2 a, float2 b) { return make_float2 (a. x * b. x + a. y * b. y, a. y * b. x - a. x * b. y); }   // a * conj (b)

// In-shared-memory Stockham autosort FFT, radix 4 with a final radix-2 pass when log2 N is odd; forward,
// unnormalised (fft.cpp:53-55).  x holds the input, y is scratch; returns the buffer holding the result.
// All OFDM_THREADS threads of the block must call it; ends with a __syncthreads.
__device__ __forceinline__ float2 *block_fft (float2 *x, float2 *y, const int N, const float2 *__restrict__ tw) {
	const int tid = threadIdx. x;
	int n = N, s = 1, ls = 0;
	__syncthreads ();
	while (n >= 4) {
		const int quarter = N >> 2, tstep = N / n;
		for (int j = tid; j < quarter; j += OFDM_THREADS) {
			const int q = j & (s - 1), p = j >> ls;
			const float2 a = x [j], b = x [j + quarter], c = x [j + 2 * quarter], d = x [j + 3 * quarter];
			const float2 apc = make_float2 (a. x + c. x, a. y + c. y), amc = make_float2 (a. x - c. x, a. y - c. y);
			const float2 bpd = make_float2 (b. x + d. x, b. y + d. y);
			const float2 jbmd = make_float2 (-(b. y - d. y), b. x - d. x);            // j * (b - d)
			const int o = q + s * 4 * p;
			y [o] = make_float2 (apc. x + bpd. x, apc. y + bpd. y);
			if (p == 0) {                                                               // twiddles are 1
				y [o + s]     = make_float2 (amc. x - jbmd. x, amc. y - jbmd. y);
				y [o + 2 * s] = make_float2 (apc. x - bpd. x, apc. y - bpd. y);
				y [o + 3 * s] = make_float2 (amc. x + jbmd. x, amc. y + jbmd. y);
			} else {
				const float2 w1 = __ldg (&tw [p * tstep]), w2 = __ldg (&tw [2 * p * tstep]), w3 = __ldg (&tw [3 * p * tstep]);
				y [o + s]     = cmul (make_float2 (amc. x - jbmd. x, amc. y - jbmd. y), w1);
				y [o + 2 * s] = cmul (make_float2 (apc. x - bpd. x, apc. y - bpd. y), w2);
				y [o + 3 * s] = cmul (make_float2 (amc. x + jbmd. x, amc. y + jbmd. y), w3);
			}
		}
		n >>= 2; s <<= 2; ls += 2;
		float2 *t = x; x = y; y = t;
		__syncthreads ();
	}
	if (n == 2) {
		for (int j = tid; j < s; j += OFDM_THREADS) {
			const float2 a = x [j], b = x [j + s];
			y [j]     = make_float2 (a. x + b. x, a. y + b. y);
			y [j + s] = make_float2 (a. x - b. x, a. y - b. y);
		}
		float2 *t = x; x = y; y = t;
		__syncthreads ();
	}
	return x;
}

// NCO phasor for localPhase index lp in [0, 2048000)
__device__ __forceinline__ float2 nco (const OfdmTables &T, int lp) {
	return cmul (__ldg (&T. osc_hi [lp >> 11]), __ldg (&T. osc_lo [lp & 2047]));
}

__device__ __forceinline__ int mod_rate (long long v) {          // v mod 2048000 into [0, 2048000)
	int r = (int) (v % DAB_INPUT_RATE);
	return r < 0 ? r + DAB_INPUT_RATE : r;
}

// dst[i] = ((iq[first+i] - 128) / 128) * osc[(lp_before - (i+1) * phase) mod RATE], i < n
// (rawfiles.cpp:113-116 + ofdm-processor.cpp:217-226).  iq = interleaved u8 I,Q.
__device__ __forceinline__ void load_u8_nco (float2 *dst, const uchar2 *__restrict__ iq, long long first, int n,
                                             int lp_before, int phase, const OfdmTables &T) {
	const int tid = threadIdx. x;
	const int ph = mod_rate (phase);
	int lp = mod_rate ((long long) lp_before - (long long) (tid + 1) * ph);
	const int step = mod_rate ((long long) OFDM_THREADS * ph);
	for (int i = tid; i < n; i += OFDM_THREADS) {
		const uchar2 s = __ldg (&iq [first + i]);
		const float2 v = make_float2 ((float) ((int) s. x - 128) * (1.0f / 128.0f), (float) ((int) s. y - 128) * (1.0f / 128.0f));
		dst [i] = cmul (v, nco (T, lp));
		lp -= step;
		if (lp < 0) lp += DAB_INPUT_RATE;
	}
}

// soft-bit quantisation of ofdm-decoder.cpp:183-189: float negate + divide, double multiply, truncation
__device__ __forceinline__ short quant127 (float num, float ab1) {
	// the reference's double product q * 127.0 is exact (24 + 7 bits); truncating it equals truncating the float product
	// rounded toward zero, because that rounding never crosses an integer (every |integer| <= 127 is a float)
	return (short) __float2int_rz (__fmul_rz (- num / ab1, 127.0f));   // NaN (ab1 == 0) -> 0, as x86's conversion ends up (App. B-5)
}

// DQPSK demod + frequency de-interleave of one symbol (ofdm-decoder.cpp:178-190): cur = FFT of the symbol,
// prev = running phase reference (updated in place on the K used bins), ibits[2K]
__device__ __forceinline__ void demod_symbol (const float2 *cur, float2 *prev, const OfdmTables &T, int16_t *__restrict__ ibits,
                                              uint8_t *__restrict__ sym8 = nullptr) {
	for (int i = threadIdx. x; i < T. K; i += OFDM_THREADS) {
		const int idx = __ldg (&T. perm [i]);
		const float2 c = cur [idx];
		const float2 r1 = cmulc (c, prev [idx]);
		prev [idx] = c;
		const float ab1 = fabsf (r1. x) + fabsf (r1. y);                 // jan_abs, dab-constants.h:127-134
		const short re = quant127 (r1. x, ab1), im = quant127 (r1. y, ab1);
		ibits [i]        = re;
		ibits [T. K + i] = im;
		if (sym8) { sym8 [i] = (uint8_t) (re + 127); sym8 [T. K + i] = (uint8_t) (im + 127); }   // the Viterbi's symbol (viterbi.cpp:229-235)
	}
}

// ---------------------------------------------------------------------------------------------------
// device routines shared with the engine
// ---------------------------------------------------------------------------------------------------
// phaseReference::findIndex on the T_u samples in buf (destroyed); scratch = second T_u buffer.
// Returns (to every thread) the reference's result: peak index, or -|Max/mean| - 1 when below threshold.
static __device__ int find_index_block (float2 *buf, float2 *scratch, const OfdmTables &T) {
	__shared__ float s_red [OFDM_THREADS / 32];
	__shared__ float s_max [OFDM_THREADS / 32];
	__shared__ int   s_idx [OFDM_THREADS / 32];
	__shared__ int   s_result;
	const int N = T. T_u, tid = threadIdx. x;
	float2 *spec = block_fft (buf, scratch, N, T. tw);
	float2 *other = spec == buf ? scratch : buf;
	// res = conj (fft * conj (ref)): the backward transform is conj (forward (conj (x)))
	for (int i = tid; i < N; i += OFDM_THREADS) {
		const float2 r = cmulc (spec [i], __ldg (&T. ref [i]));
		spec [i] = make_float2 (r. x, - r. y);
	}
	float2 *res = block_fft (spec, other, N, T. tw);
	const float factor = (float) (1.0 / (float) N);          // fft.cpp:114-121
	float sum = 0.f, mx = -10000.f;
	int mi = -1;
	for (int i = tid; i < N; i += OFDM_THREADS) {
		const float2 v = res [i];
		const float a = hypotf (v. x * factor, (- v. y) * factor);
		sum += a;
		if (a > mx) { mx = a; mi = i; }
	}
	for (int o = 16; o > 0; o >>= 1) {
		sum += __shfl_xor_sync (0xffffffffu, sum, o);
		const float om = __shfl_xor_sync (0xffffffffu, mx, o);
		const int   oi = __shfl_xor_sync (0xffffffffu, mi, o);
		if (om > mx || (om == mx && oi >= 0 && (mi < 0 || oi < mi))) { mx = om; mi = oi; }   // first maximum wins
	}
	if ((tid & 31) == 0) { s_red [tid >> 5] = sum; s_max [tid >> 5] = mx; s_idx [tid >> 5] = mi; }
	__syncthreads ();
	if (tid == 0) {
		float tsum = 0.f, tmx = -10000.f;
		int tmi = -1;
		for (int w = 0; w < OFDM_THREADS / 32; w ++) {
			tsum += s_red [w];
			if (s_max [w] > tmx || (s_max [w] == tmx && s_idx [w] >= 0 && (tmi < 0 || s_idx [w] < tmi))) { tmx = s_max [w]; tmi = s_idx [w]; }
		}
		if (tmx < (float) T. level * tsum / (float) N)                       // phasereference.cpp:84-85
			s_result = (int) (- fabsf (tmx / (tsum / (float) N)) - 1.0f);
		else
			s_result = tmi;
	}
	__syncthreads ();
	return s_result;
}

static __device__ __forceinline__ float arg_mulconj (float2 a, float2 b) {
	const float2 r = cmulc (a, b);
	return atan2f (r. y, r. x);
}

// coarse frequency offset from the PRS spectrum (ofdm-decoder.cpp:99-161), executed by warp 0; the result
// is returned to thread 0 only.
static __device__ int coarse_offset_warp0 (const float2 *f, const OfdmTables &T, float *cv /* smem, >= 90 floats */) {
	const int N = T. T_u, tid = threadIdx. x;
	int result = 100;
	if (T. method == 1) {                                    // :106-127
		if (tid < 90) {
			const int base = N - 36 + tid;
			cv [tid] = arg_mulconj (f [base % N], f [(base + 1) % N]);
		}
		__syncthreads ();
		if (tid < 32) {
			float best = 0.f; int bi = 100;                  // index_1 stays 100 when no sum exceeds 0
			for (int i = tid; i < 72; i += 32) {
				float sum = 0.f;
				for (int j = 1; j < 18; j ++) sum += fabsf (__ldg (&T. refArg [j]) * cv [i + j]);
				if (sum > best) { best = sum; bi = i; }
			}
			for (int o = 16; o > 0; o >>= 1) {
				const float ob = __shfl_xor_sync (0xffffffffu, best, o);
				const int   oi = __shfl_xor_sync (0xffffffffu, bi, o);
				if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
			}
			result = N - 36 + bi - N;
		}
	} else if (T. method == 2) {                             // :132-161
		__syncthreads ();
		if (tid < 32) {
			float best = 1000.f; int bi = 100;
			for (int k = tid; k < 72; k += 32) {
				const int i = N - 36 + k;
#define ARGP(x, y) arg_mulconj (f [(x) % N], f [(y) % N])
				const float a1 = (float) fabs ((double) fabsf (ARGP (i + 1, i + 2)) / M_PI - 1);
				const float a2 = (float) fabs ((double) fabsf (ARGP (i + 2, i + 3)) / M_PI - 1);
				const float a3 = fabsf (ARGP (i + 3, i + 4)), a4 = fabsf (ARGP (i + 4, i + 5)), a5 = fabsf (ARGP (i + 5, i + 6));
				const float b1 = (float) fabs ((double) fabsf (ARGP (i + 17, i + 19)) / M_PI - 1);
				const float b2 = fabsf (ARGP (i + 19, i + 20)), b3 = fabsf (ARGP (i + 20, i + 21)), b4 = fabsf (ARGP (i + 21, i + 22));
#undef ARGP
				const float sum = a1 + a2 + a3 + a4 + a5 + b1 + b2 + b3 + b4;
				if (sum < best) { best = sum; bi = i; }
			}
			for (int o = 16; o > 0; o >>= 1) {
				const float ob = __shfl_xor_sync (0xffffffffu, best, o);
				const int   oi = __shfl_xor_sync (0xffffffffu, bi, o);
				if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
			}
			result = bi - N;
		}
	} else {                                                 // method 0: getMiddle, :233-258, bug included
		__syncthreads ();
		if (tid == 0) {
			float sum = 0.f; const float oldMax = 0.f; int maxIndex = 0;
			for (int i = 40; i < 1536 + 40; i ++) { const float2 v = f [(N / 2 + i) % N]; sum += hypotf (v. x, v. y); }
			for (int i = 40; i < N - (1536 - 40); i ++) {
				float2 v = f [(N / 2 + i) % N];        sum -= hypotf (v. x, v. y);
				v = f [(N / 2 + i + 1536) % N];        sum += hypotf (v. x, v. y);
				if (sum > oldMax) { sum = oldMax; maxIndex = i; }
			}
			result = (int16_t) (maxIndex - (N - 1536) / 2);
		}
	}
	return result;
}


// ---------------------------------------------------------------------------------------------------
// 2048-point forward FFT for 256 threads, 8 points per thread in registers, decimation in frequency with
// radices 4 x 8 x 8 x 8: three shared-memory exchanges instead of the six passes of the generic routine.
// Thread t enters with x[t + 256 k], k = 0..7 (so the first butterflies need no shared memory at all).
// The result is left in shared memory in digit-reversed order: X[k] sits at r8_pos (k) = 512 (k & 3) +
// 64 ((k >> 2) & 7) + 8 ((k >> 5) & 7) + (k >> 8); consumers index through that map (the demodulator's carrier
// table is pre-permuted on the host), so no reordering pass exists.
// Shared-memory layout: element i lives at r8_swz (i) = i ^ h, h = (i4 i5 i6) into bits 0..2 and i6 again into
// bit 3 (iN = bit N of i).  With 8-byte elements a half-warp is conflict-free iff its 16 addresses differ in bits
// 0..3; this XOR makes that true for all four access patterns of the transform (16 consecutive elements; two runs
// of 8, 64 apart; stride 8) and, unlike padding, keeps every address of a thread of the form `base ^ constant`
// or `base + constant`, so the per-symbol address arithmetic all but disappears.
// ---------------------------------------------------------------------------------------------------
#define R8_N      2048
#define R8_SMEM   R8_N                         // float2 elements
__host__ __device__ __forceinline__ int r8_swz (int i) { return i ^ (((i >> 4) & 7) | (((i >> 6) & 1) << 3)); }
__host__ __device__ __forceinline__ int r8_pos (int k) { return 512 * (k & 3) + 64 * ((k >> 2) & 7) + 8 * ((k >> 5) & 7) + (k >> 8); }

__device__ __forceinline__ float2 cadd (float2 a, float2 b) { return make_float2 (a. x + b. x, a. y + b. y); }
__device__ __forceinline__ float2 csub (float2 a, float2 b) { return make_float2 (a. x - b. x, a. y - b. y); }
__device__ __forceinline__ float2 cmulmj (float2 a) { return make_float2 (a. y, - a. x); }          // a * (-j)

// 4-point DFT (forward): y_q = sum_m c_m exp (-2 pi i m q / 4)
__device__ __forceinline__ void dft4 (float2 c0, float2 c1, float2 c2, float2 c3, float2 &y0, float2 &y1, float2 &y2, float2 &y3) {
	const float2 d0 = cadd (c0, c2), d2 = csub (c0, c2), d1 = cadd (c1, c3), d3 = cmulmj (csub (c1, c3));
	y0 = cadd (d0, d1); y2 = csub (d0, d1); y1 = cadd (d2, d3); y3 = csub (d2, d3);
}
// 8-point DFT (forward), in place: a[q] <- sum_m a[m] exp (-2 pi i m q / 8)
__device__ __forceinline__ void dft8 (float2 (&a) [8]) {
	const float h = 0.70710678118654752440f;
	const float2 b0 = cadd (a [0], a [4]), b4 = csub (a [0], a [4]);
	const float2 b1 = cadd (a [1], a [5]), t5 = csub (a [1], a [5]);
	const float2 b2 = cadd (a [2], a [6]), t6 = csub (a [2], a [6]);
	const float2 b3 = cadd (a [3], a [7]), t7 = csub (a [3], a [7]);
	const float2 b5 = make_float2 (h * (t5. x + t5. y), h * (t5. y - t5. x));      // * (1 - j) / sqrt 2
	const float2 b6 = cmulmj (t6);                                                 // * (-j)
	const float2 b7 = make_float2 (h * (t7. y - t7. x), - h * (t7. x + t7. y));    // * (-1 - j) / sqrt 2
	dft4 (b0, b1, b2, b3, a [0], a [2], a [4], a [6]);
	dft4 (b4, b5, b6, b7, a [1], a [3], a [5], a [7]);
}

// x[k] = sample t + 256 k on entry; A = R8_SMEM float2 of shared memory.  Twiddles: tw1[2 g + ...] = the thread's own six
// stage-1 factors W_2048^(n q), n = t + 256 g, q = 1..3 (registers, loaded once per CTA by r8_load_tw1); tw2[j] = W_512^j
// and tw3[j] = W_64^j, stored per output index q (r8_fill_tables), in shared memory -- with most of the SM's on-chip
// memory carved out as shared memory the L1 is too small to keep a global twiddle table resident.
#ifndef R8_TWPOW
#define R8_TWPOW 0                     // 1: stage-2/3 twiddles as powers of one loaded value (A/B switch)
#endif
#define R8_TW2 512
#define R8_TW3 64
__device__ __forceinline__ void r8_fill_tables (float2 *tw2, float2 *tw3, const float2 *__restrict__ tw) {
	// one run per output index q, so that the lanes of a warp read consecutive entries (tw2 [n q] for even q is a
	// 2- or 4-way bank conflict): tw2 [64 (q - 1) + n] = W_512^(n q), tw3 [8 (q - 1) + n] = W_64^(n q), q = 1..7
	for (int j = threadIdx. x; j < 7 * 64; j += 256) tw2 [j] = __ldg (&tw [4 * (((j & 63) * ((j >> 6) + 1)) & 511)]);
	if (threadIdx. x < 7 * 8) tw3 [threadIdx. x] = __ldg (&tw [32 * ((((int) threadIdx. x & 7) * (((int) threadIdx. x >> 3) + 1)) & 63)]);
}
__device__ __forceinline__ void r8_load_tw1 (float2 (&tw1) [6], const float2 *__restrict__ tw) {
#pragma unroll
	for (int g = 0; g < 2; g ++)
#pragma unroll
		for (int q = 1; q < 4; q ++) tw1 [3 * g + q - 1] = __ldg (&tw [(((int) threadIdx. x + 256 * g) * q) & 2047]);
}

__device__ __forceinline__ float2 r8_lds (uint32_t a) {
	float2 v;
	asm volatile ("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f" (v. x), "=f" (v. y) : "r" (a));
	return v;
}
__device__ __forceinline__ void r8_sts (uint32_t a, float2 v) {
	asm volatile ("st.shared.v2.f32 [%0], {%1, %2};" :: "r" (a), "f" (v. x), "f" (v. y) : "memory");
}

__device__ __forceinline__ void fft2048_r8 (float2 (&x) [8], float2 *A, const float2 (&tw1) [6], const float2 *tw2, const float2 *tw3) {
	const int t = threadIdx. x;
	// 32-bit shared-window addresses throughout.  A must be 512-byte aligned: the thread bases below already contain
	// A's address, and the per-element constants are XORed into address bits 3..8
	const uint32_t Ab = (uint32_t) __cvta_generic_to_shared (A);
#define R8_LD(addr)      r8_lds ((uint32_t) (addr))
#define R8_ST(addr, v)   r8_sts ((uint32_t) (addr), (v))
	// stage 1: two radix-4 butterflies, n = t (k even) and n = t + 256 (k odd); twiddle W_2048^(n q).
	// element t + 256 g + 512 q: the swizzle only sees bits of t
	{
		const uint32_t base = Ab + r8_swz (t) * 8;
		float2 y [4];
#pragma unroll
		for (int g = 0; g < 2; g ++) {
			dft4 (x [g], x [g + 2], x [g + 4], x [g + 6], y [0], y [1], y [2], y [3]);
			R8_ST (base + 8 * (256 * g), y [0]);
#pragma unroll
			for (int q = 1; q < 4; q ++) R8_ST (base + 8 * (256 * g + 512 * q), cmul (y [q], tw1 [3 * g + q - 1]));
		}
	}
	__syncthreads ();
	// stage 2: radix 8 inside blocks of 512: n = t & 63, twiddle W_512^(n q).  element 512 b + n + 64 m: bit 6 = m & 1
	{
		const int n = t & 63;
		const uint32_t be = Ab + r8_swz (512 * (t >> 6) + n) * 8, bo = be ^ (0xC * 8);
		float2 a [8];
#pragma unroll
		for (int m = 0; m < 8; m ++) a [m] = R8_LD (((m & 1) ? bo : be) + 8 * 64 * m);
		dft8 (a);
		R8_ST (be, a [0]);
#if R8_TWPOW
		{	// powers of one loaded twiddle instead of seven loads: shared memory is the stressed resource, the FP pipe has room
			const float2 w1 = tw2 [n], w2 = cmul (w1, w1), w3 = cmul (w2, w1), w4 = cmul (w2, w2);
			const float2 w5 = cmul (w4, w1), w6 = cmul (w3, w3), w7 = cmul (w4, w3);
			const float2 wq [8] = { w1, w1, w2, w3, w4, w5, w6, w7 };
#pragma unroll
			for (int q = 1; q < 8; q ++) R8_ST (((q & 1) ? bo : be) + 8 * 64 * q, cmul (a [q], wq [q]));
		}
#else
#pragma unroll
		for (int q = 1; q < 8; q ++) R8_ST (((q & 1) ? bo : be) + 8 * 64 * q, cmul (a [q], tw2 [64 * (q - 1) + n]));
#endif
	}
	// the 512-point block of stage 2 is produced and consumed by the same 64 threads (two warps): a named barrier for
	// them instead of the whole CTA
	switch (t >> 6) {                                        // (literal ids: a register id makes ptxas reserve all 16 barriers)
	case 0:  asm volatile ("bar.sync 1, 64;" ::: "memory"); break;
	case 1:  asm volatile ("bar.sync 2, 64;" ::: "memory"); break;
	case 2:  asm volatile ("bar.sync 3, 64;" ::: "memory"); break;
	default: asm volatile ("bar.sync 4, 64;" ::: "memory"); break;
	}
	// stage 3: radix 8 inside blocks of 64: n = t & 7, twiddle W_64^(n q).  element 64 b + n + 8 m: bits 3 | 4,5 = m,
	// bit 6 = b & 1, so the address is base ^ K(m) with K(m) = 8 ((m >> 1) & 3) + 64 (m & 1) + 128 (m >> 1)
	{
		const int n = t & 7, bb = t >> 3;
		const uint32_t base = Ab + ((512 * bb + 8 * (n ^ (4 * (bb & 1)))) ^ (64 * (bb & 1)));
		float2 a [8];
#pragma unroll
		for (int m = 0; m < 8; m ++) a [m] = R8_LD (base ^ (8 * ((m >> 1) & 3) + 64 * (m & 1) + 128 * (m >> 1)));
		dft8 (a);
		R8_ST (base, a [0]);
#if R8_TWPOW
		{
			const float2 w1 = tw3 [n], w2 = cmul (w1, w1), w3 = cmul (w2, w1), w4 = cmul (w2, w2);
			const float2 w5 = cmul (w4, w1), w6 = cmul (w3, w3), w7 = cmul (w4, w3);
			const float2 wq [8] = { w1, w1, w2, w3, w4, w5, w6, w7 };
#pragma unroll
			for (int q = 1; q < 8; q ++) R8_ST (base ^ (8 * ((q >> 1) & 3) + 64 * (q & 1) + 128 * (q >> 1)), cmul (a [q], wq [q]));
		}
#else
#pragma unroll
		for (int q = 1; q < 8; q ++) R8_ST (base ^ (8 * ((q >> 1) & 3) + 64 * (q & 1) + 128 * (q >> 1)), cmul (a [q], tw3 [8 * (q - 1) + n]));
#endif
	}
	__syncwarp ();                                           // a 64-point block lives in 8 consecutive threads
	// stage 4: radix 8 on 8 consecutive points, no twiddles.  element 8 t + m: address = base ^ 8 m
	{
		const uint32_t base = Ab + r8_swz (8 * t) * 8;
		float2 a [8];
#pragma unroll
		for (int m = 0; m < 8; m ++) a [m] = R8_LD (base ^ (8 * m));
		dft8 (a);
#pragma unroll
		for (int q = 0; q < 8; q ++) R8_ST (base ^ (8 * q), a [q]);
	}
	__syncthreads ();
#undef R8_LD
#undef R8_ST
}
