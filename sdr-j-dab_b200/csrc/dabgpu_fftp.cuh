// dabgpu_fftp.cuh -- the register FFT of dabgpu_ofdm.cuh (fft2048_r8) re-written on sm_100a's packed FP32 pipe:
// a complex value is one 64-bit register pair and every complex add / subtract / scale is ONE instruction
// (FADD2 / FFMA2 / FMUL2; per-lane IEEE rounding, so the results are those of the scalar code bit for bit wherever the
// same operations are performed).  The SASS forms used (checked with cuobjdump, see profiles/r02_sass_*.txt):
//   * whole-operand negation  (-R.F32x2)        : a - b is FADD2 (a, -b);
//   * half swap               (R.F32x2.LO_HI)   : multiplication by +-j is a swap plus a (1, -1) / (-1, 1) factor folded
//                                                 into an FFMA2;
//   * scalar broadcast        (R.F32)           : a real factor applied to both halves.
// A complex product a * w costs three instructions: j a (FMUL2 on the swapped operand), a * w.x (FMUL2, broadcast),
// (j a) * w.y + that (FFMA2, broadcast) -- against four scalar ones -- and needs no re-arranged copy of the twiddle.
//
// Generic over the transform length through the number of symbols a 256-thread CTA transforms side by side:
//   NSYM = 1: one 2048-point transform  (Mode I):   radix 4 . 8 . 8 . 8, all 256 threads on one symbol
//   NSYM = 2: two 1024-point transforms (Mode IV):  radix 2 . 8 . 8 . 8, 128 threads per symbol
//   NSYM = 4: four 512-point transforms (Mode II):  radix     8 . 8 . 8,  64 threads per symbol
// Stages 2-4 of the 2048-point transform ARE four independent 512-point transforms on 64 threads each, so the three
// variants share them; only the head differs.  In every variant thread t enters with the eight samples
// u + G k (k = 0..7) of its symbol, G = 256 / NSYM threads per symbol, u = t mod G.
#pragma once
#include "dabgpu_ofdm.cuh"

typedef float2 c32;
__device__ __forceinline__ c32 p_add (c32 a, c32 b) { return __fadd2_rn (a, b); }
__device__ __forceinline__ c32 p_sub (c32 a, c32 b) { return __fadd2_rn (a, make_float2 (- b. x, - b. y)); }
__device__ __forceinline__ c32 p_swap (c32 a) { return make_float2 (a. y, a. x); }
// a + (-j) e  and  a + (+j) e
__device__ __forceinline__ c32 p_add_mj (c32 a, c32 e) { return __ffma2_rn (p_swap (e), make_float2 (1.f, -1.f), a); }
__device__ __forceinline__ c32 p_add_pj (c32 a, c32 e) { return __ffma2_rn (p_swap (e), make_float2 (-1.f, 1.f), a); }
__device__ __forceinline__ c32 p_j (c32 a) { return __fmul2_rn (p_swap (a), make_float2 (-1.f, 1.f)); }          // j a
__device__ __forceinline__ c32 p_scale (c32 a, float s) { return __fmul2_rn (a, make_float2 (s, s)); }
// a * w  and  a * conj (w)
__device__ __forceinline__ c32 p_cmul (c32 a, c32 w) {
	return __ffma2_rn (p_j (a), make_float2 (w. y, w. y), __fmul2_rn (a, make_float2 (w. x, w. x)));
}
__device__ __forceinline__ c32 p_cmulc (c32 a, c32 w) {
	return __ffma2_rn (p_j (a), make_float2 (- w. y, - w. y), __fmul2_rn (a, make_float2 (w. x, w. x)));
}

// a * w with w given as its real part and the prepared pair (-w.y, w.y): two instructions (for factors that are reused)
__device__ __forceinline__ c32 p_cmul_pre (c32 a, float wx, c32 wyy) { return __ffma2_rn (p_swap (a), wyy, __fmul2_rn (a, make_float2 (wx, wx))); }

// 4-point DFT (forward): y_q = sum_m c_m exp (-2 pi i m q / 4); 8 packed instructions
__device__ __forceinline__ void p_dft4 (c32 c0, c32 c1, c32 c2, c32 c3, c32 &y0, c32 &y1, c32 &y2, c32 &y3) {
	const c32 d0 = p_add (c0, c2), d2 = p_sub (c0, c2), d1 = p_add (c1, c3), e = p_sub (c1, c3);
	y0 = p_add (d0, d1); y2 = p_sub (d0, d1); y1 = p_add_mj (d2, e); y3 = p_add_pj (d2, e);
}
// 8-point DFT (forward), in place; 26 packed instructions.  Odd half: b5 = h (t5 - j t5), b6 = -j t6, b7 = -h (t7 + j t7)
// (h = 1 / sqrt 2); the factor h and the multiplication of b6 by -j are folded into the closing butterflies.
__device__ __forceinline__ void p_dft8 (c32 (&a) [8]) {
	const float h = 0.70710678118654752440f;
	const c32 b0 = p_add (a [0], a [4]), b4 = p_sub (a [0], a [4]);
	const c32 b1 = p_add (a [1], a [5]), t5 = p_sub (a [1], a [5]);
	const c32 b2 = p_add (a [2], a [6]), t6 = p_sub (a [2], a [6]);
	const c32 b3 = p_add (a [3], a [7]), t7 = p_sub (a [3], a [7]);
	p_dft4 (b0, b1, b2, b3, a [0], a [2], a [4], a [6]);
	const c32 u5 = p_add_mj (t5, t5), u7 = p_add_pj (t7, t7);            // b5 = h u5, b7 = -h u7
	const c32 s = p_add (u5, u7), r = p_sub (u5, u7);                    // b5 - b7 = h s, b5 + b7 = h r
	const c32 d0 = p_add_mj (b4, t6), d2 = p_add_pj (b4, t6);            // b4 + b6, b4 - b6
	a [1] = __ffma2_rn (r, make_float2 (h, h), d0);                      // d0 + (b5 + b7)
	a [5] = __ffma2_rn (r, make_float2 (- h, - h), d0);                  // d0 - (b5 + b7)
	a [3] = __ffma2_rn (p_swap (s), make_float2 (h, - h), d2);           // d2 + (-j) (b5 - b7)
	a [7] = __ffma2_rn (p_swap (s), make_float2 (- h, h), d2);           // d2 + (+j) (b5 - b7)
}

// ---- stages 2-4: 512-point transforms inside blocks of 512 consecutive elements, 64 threads (two warps) each ----
// Same shared-memory layout, swizzle and barriers as fft2048_r8 (dabgpu_ofdm.cuh).  REGS_IN: the thread already holds
// its stage-2 inputs a[m] = element n + 64 m of its block (NSYM = 4: they are the samples themselves).
template <bool REGS_IN>
__device__ __forceinline__ void p_fft512_tail (c32 (&a) [8], const uint32_t Ab, const float2 *tw2, const float2 *tw3) {
	const int t = threadIdx. x;
	{	// stage 2: radix 8, n = t & 63, twiddle W_512^(n q).  element 512 b + n + 64 m: bit 6 = m & 1
		const int n = t & 63;
		const uint32_t be = Ab + r8_swz (512 * (t >> 6) + n) * 8, bo = be ^ (0xC * 8);
		if (!REGS_IN) {
#pragma unroll
			for (int m = 0; m < 8; m ++) a [m] = r8_lds (((m & 1) ? bo : be) + 8 * 64 * m);
		}
		// (the stores below are compiler barriers: the seven twiddles are fetched up front, next to the data, instead of one
		// by one between the stores -- a chain of seven shared-memory latencies otherwise)
		float2 w [7];
#pragma unroll
		for (int q = 1; q < 8; q ++) w [q - 1] = tw2 [64 * (q - 1) + n];
		p_dft8 (a);
		r8_sts (be, a [0]);
#pragma unroll
		for (int q = 1; q < 8; q ++) r8_sts (((q & 1) ? bo : be) + 8 * 64 * q, p_cmul (a [q], w [q - 1]));
	}
	switch (t >> 6) {                                        // the block's two warps only (literal ids, see fft2048_r8)
	case 0:  asm volatile ("bar.sync 1, 64;" ::: "memory"); break;
	case 1:  asm volatile ("bar.sync 2, 64;" ::: "memory"); break;
	case 2:  asm volatile ("bar.sync 3, 64;" ::: "memory"); break;
	default: asm volatile ("bar.sync 4, 64;" ::: "memory"); break;
	}
	{	// stage 3: radix 8 inside blocks of 64: n = t & 7, twiddle W_64^(n q)
		const int n = t & 7, bb = t >> 3;
		const uint32_t base = Ab + ((512 * bb + 8 * (n ^ (4 * (bb & 1)))) ^ (64 * (bb & 1)));
#pragma unroll
		for (int m = 0; m < 8; m ++) a [m] = r8_lds (base ^ (8 * ((m >> 1) & 3) + 64 * (m & 1) + 128 * (m >> 1)));
		float2 w [7];
#pragma unroll
		for (int q = 1; q < 8; q ++) w [q - 1] = tw3 [8 * (q - 1) + n];
		p_dft8 (a);
		r8_sts (base, a [0]);
#pragma unroll
		for (int q = 1; q < 8; q ++) r8_sts (base ^ (8 * ((q >> 1) & 3) + 64 * (q & 1) + 128 * (q >> 1)), p_cmul (a [q], w [q - 1]));
	}
	__syncwarp ();
	{	// stage 4: radix 8 on 8 consecutive points, no twiddles
		const uint32_t base = Ab + r8_swz (8 * t) * 8;
#pragma unroll
		for (int m = 0; m < 8; m ++) a [m] = r8_lds (base ^ (8 * m));
		p_dft8 (a);
#pragma unroll
		for (int q = 0; q < 8; q ++) r8_sts (base ^ (8 * q), a [q]);
	}
}

// Where X[k] of symbol j (j < NSYM) ends up in the NSYM x (2048 / NSYM) element buffer, before the swizzle.
// 512-point blocks leave X[k'] at 64 (k' & 7) + 8 ((k' >> 3) & 7) + (k' >> 6); the head of the longer transforms sends
// X[k] to block k mod HEAD with k' = k / HEAD (decimation in frequency).
template <int NSYM> __host__ __device__ __forceinline__ int p_pos (int j, int k) {
	const int head = 4 / NSYM;                               // 4, 2, 1
	const int b = k % head, kk = k / head;
	return 512 * ((NSYM == 1 ? 0 : NSYM == 2 ? 2 * j : j) + b) + 64 * (kk & 7) + 8 * ((kk >> 3) & 7) + (kk >> 6);
}

// head twiddles of a thread (registers, loaded once per CTA): NSYM = 1: W_2048^(n q), n = t + 256 g, q = 1..3 (six);
// NSYM = 2: W_1024^(u + 128 g), g = 0..3 (four)
template <int NSYM> __device__ __forceinline__ void p_load_tw1 (float2 (&tw1) [6], const float2 *__restrict__ tw) {
	if (NSYM == 1) {
#pragma unroll
		for (int g = 0; g < 2; g ++)
#pragma unroll
			for (int q = 1; q < 4; q ++) tw1 [3 * g + q - 1] = __ldg (&tw [(((int) threadIdx. x + 256 * g) * q) & 2047]);
	} else if (NSYM == 2) {
#pragma unroll
		for (int g = 0; g < 4; g ++) tw1 [g] = __ldg (&tw [((int) threadIdx. x & 127) + 128 * g]);   // tw = W_1024 table
		tw1 [4] = tw1 [5] = make_float2 (0.f, 0.f);
	}
}
// stage-2/3 twiddle tables in shared memory from the handle's W_N table (N = 2048 / NSYM): tw2 [64 (q - 1) + n] = W_512^(n q),
// tw3 [8 (q - 1) + n] = W_64^(n q), q = 1..7
template <int NSYM> __device__ __forceinline__ void p_fill_tables (float2 *tw2, float2 *tw3, const float2 *__restrict__ tw) {
	const int s2 = 4 / NSYM, s3 = 32 / NSYM;                 // W_512 = W_N^(N / 512), W_64 = W_N^(N / 64)
	for (int j = threadIdx. x; j < 7 * 64; j += 256) tw2 [j] = __ldg (&tw [s2 * (((j & 63) * ((j >> 6) + 1)) & 511)]);
	if (threadIdx. x < 7 * 8) tw3 [threadIdx. x] = __ldg (&tw [s3 * ((((int) threadIdx. x & 7) * (((int) threadIdx. x >> 3) + 1)) & 63)]);
}

// The transform.  x[k] = sample u + G k of the thread's symbol on entry; the spectra are left in A (2048 float2, 512-byte
// aligned) at r8_swz (p_pos<NSYM> (symbol, k)).  Ends with a CTA barrier.
template <int NSYM>
__device__ __forceinline__ void p_fft (c32 (&x) [8], float2 *A, const float2 (&tw1) [6], const float2 *tw2, const float2 *tw3) {
	const int t = threadIdx. x;
	const uint32_t Ab = (uint32_t) __cvta_generic_to_shared (A);
	if (NSYM == 1) {
		// two radix-4 butterflies, n = t (k even) and n = t + 256 (k odd); element t + 256 g + 512 q
		const uint32_t base = Ab + r8_swz (t) * 8;
		c32 y [4];
#pragma unroll
		for (int g = 0; g < 2; g ++) {
			p_dft4 (x [g], x [g + 2], x [g + 4], x [g + 6], y [0], y [1], y [2], y [3]);
			r8_sts (base + 8 * (256 * g), y [0]);
#pragma unroll
			for (int q = 1; q < 4; q ++) r8_sts (base + 8 * (256 * g + 512 * q), p_cmul (y [q], tw1 [3 * g + q - 1]));
		}
		__syncthreads ();
		p_fft512_tail<false> (x, Ab, tw2, tw3);
	} else if (NSYM == 2) {
		// four radix-2 butterflies per thread: samples j = u + 128 g and j + 512 -> a + b into block 2 s, (a - b) W_1024^j
		// into block 2 s + 1, both at offset j.  The swizzle sees bits 4..6 of the offset only: u + 128 g -> bits of u.
		const int u = t & 127, sy = t >> 7;
		const uint32_t base = Ab + (1024 * sy + r8_swz (u)) * 8;
#pragma unroll
		for (int g = 0; g < 4; g ++) {
			r8_sts (base + 8 * (128 * g), p_add (x [g], x [g + 4]));
			r8_sts (base + 8 * (128 * g + 512), p_cmul (p_sub (x [g], x [g + 4]), tw1 [g]));
		}
		// a symbol's 1024 elements are written and read by the same 128 threads: named barrier per symbol
		if (sy == 0) asm volatile ("bar.sync 5, 128;" ::: "memory"); else asm volatile ("bar.sync 6, 128;" ::: "memory");
		p_fft512_tail<false> (x, Ab, tw2, tw3);
	} else {
		p_fft512_tail<true> (x, Ab, tw2, tw3);
	}
	__syncthreads ();
}
