// dabgpu_vit_simd.cu -- throughput Viterbi: one terminated code word per THREAD, all 64 path metrics of the
// word in 32 registers as packed 16-bit pairs, add-compare-select with the sm_100a packed integer instructions
// (VIADD.16x2, VIMNMX.U16x2 with its two predicate outputs = the two decisions), decisions streamed to HBM as
// one coalesced uint2 per step, chain-back in a second kernel.  Used when a launch has enough code words to
// fill the GPU (the stream engine, large API batches); small batches use the warp-cooperative kernel of
// dabgpu_viterbi.cu.  Same arithmetic as the reference (viterbi.cpp:225-357, spiral-*.c), bit-exact:
//   * metrics start 63 / state 0 = 0; branch metric sum_j (Branchtab_j ^ sym_j), complement 1020 - m;
//   * VIMNMX's predicate is (upper <= lower): a tie keeps predecessor i, exactly the reference's strict '>';
//   * the common minimum is subtracted every 40 steps (differences, hence decisions, unchanged) so that the
//     16-bit lanes never wrap: spread <= 6*1020+63 and 40 steps add <= 40800  ->  < 2^16.
//
// Register layout.  State s = (s5..s0).  At the start of step k (k = t mod 5) "half position" p = k: register
// r holds the two states whose index is r with a 0 / 1 inserted at bit p (low / high half).  The butterfly pairs
// states differing in s5, i.e. registers r and r+16, same halves, so one packed op does two butterflies; the
// results 2i and 2i+1 go to registers 2r and 2r+1 and the half position becomes p+1.  After five steps the half
// position is 5 = the butterfly bit itself, and 32 PRMTs bring it back to 0.  All indices are compile-time, the
// code is unrolled over the five-step cycle.
#include "dabgpu_internal.h"

#define VS_THREADS 64          // code words per CTA
#define VS_CHUNK   40          // trellis steps staged per round (8 layout cycles, one renormalisation)
#define VS_ROWW   41          // tile row stride in words (odd: conflict-free)

__constant__ int8_t c_vs_delay [16] = { 15, 7, 11, 3, 13, 5, 9, 1, 14, 6, 10, 2, 12, 4, 8, 0 };   // dab-concurrent.cpp:41-43

__host__ __device__ constexpr int vs_parity (unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1; }
// branch pattern of butterfly i: bit0 = polys 0155 (used twice: j = 0 and 3), bit1 = 0117, bit2 = 0123 (viterbi.cpp:63, 159-164)
__host__ __device__ constexpr int vs_pat (int i) {
	return vs_parity ((2u * i) & 0155u) | (vs_parity ((2u * i) & 0117u) << 1) | (vs_parity ((2u * i) & 0123u) << 2);
}
__host__ __device__ constexpr int vs_insert0 (int r, int p) { return ((r >> p) << (p + 1)) | (r & ((1 << p) - 1)); }

// packed compare-select: min (upper, lower) per 16-bit lane; `bit` is ORed into wlo / whi where the lower candidate
// won (decision = upper > lower; a tie keeps the upper one).  __vibmin_u16x2 is one VIMNMX.U16x2 with two predicate
// outputs (predicate = upper <= lower) on sm_100a.
__device__ __forceinline__ uint32_t vs_acs (uint32_t upper, uint32_t lower, uint32_t &wlo, uint32_t &whi, const uint32_t bit, const uint32_t one) {
	bool ph, pl;
	const uint32_t r = __vibmin_u16x2 (upper, lower, &ph, &pl);
	// decision bits are accumulated with predicated multiply-adds (acc += one * bit): `one` is 1 at run time but opaque
	// to the compiler, which keeps these on the FMA pipe; the ALU pipe, busy with the packed min and everything else,
	// is the scarce one here (ncu: ALU 75 % vs FMA 22 % before this change)
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p mad.lo.u32 %0, %3, %2, %0; }" : "+r" (wlo) : "r" ((uint32_t) pl), "r" (bit), "r" (one));
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p mad.lo.u32 %0, %3, %2, %0; }" : "+r" (whi) : "r" ((uint32_t) ph), "r" (bit), "r" (one));
	return r;
}

template <int P>
__device__ __forceinline__ void vs_step (const uint32_t (&R) [32], uint32_t (&Q) [32], const uint32_t sym, uint2 &dec, const uint32_t one) {
	// the 8 branch metrics B[x], x = pattern, and their packed pairs PM[x] = B[x] | B[x ^ delta] << 16
	const uint32_t s0 = sym & 255u, s1 = (sym >> 8) & 255u, s2 = (sym >> 16) & 255u, s3 = sym >> 24;
	const uint32_t a0 = s0 + s3, a1 = 510u - a0, b1 = 255u - s1, c1 = 255u - s2;
	uint32_t B [8];
	B [0] = a0 + s1 + s2; B [1] = a1 + s1 + s2; B [2] = a0 + b1 + s2; B [3] = a1 + b1 + s2;
	B [4] = a0 + s1 + c1; B [5] = a1 + s1 + c1; B [6] = a0 + b1 + c1; B [7] = a1 + b1 + c1;
	constexpr int delta = vs_pat (1 << P);
	uint32_t PM [8];
#pragma unroll
	for (int x = 0; x < 8; x ++) PM [x] = B [x] + (B [x ^ delta] << 16);
	uint32_t wlo = 0, whi = 0;
#pragma unroll
	for (int r = 0; r < 16; r ++) {
		constexpr int dummy = 0; (void) dummy;
		const int x = vs_pat (vs_insert0 (r, P));
		const uint32_t a = R [r], b = R [r + 16];
		// plain 32-bit adds: every 16-bit lane stays below 2^16 (see the renormalisation bound), so no carry
		// crosses the halves and the full-rate IADD replaces the half-rate VIADD.16x2
		const uint32_t m0 = a + PM [x],     m1 = b + PM [7 - x];
		const uint32_t m2 = a + PM [7 - x], m3 = b + PM [x];
		Q [2 * r]     = vs_acs (m0, m1, wlo, whi, 1u << (2 * r), one);
		Q [2 * r + 1] = vs_acs (m2, m3, wlo, whi, 1u << (2 * r + 1), one);
	}
	dec = make_uint2 (wlo, whi);
}

// half position 5 -> 0
__device__ __forceinline__ void vs_repack (const uint32_t (&Q) [32], uint32_t (&R) [32]) {
#pragma unroll
	for (int q = 0; q < 32; q ++)
		R [q] = q < 16 ? __byte_perm (Q [(2 * q) & 31], Q [(2 * q + 1) & 31], 0x5410)
		               : __byte_perm (Q [(2 * q) & 31], Q [(2 * q + 1) & 31], 0x7632);
}

__device__ __forceinline__ void vs_renorm (uint32_t (&R) [32]) {
	uint32_t m = R [0];
#pragma unroll
	for (int q = 1; q < 32; q ++) m = __vminu2 (m, R [q]);
	const uint32_t mm = min (m & 0xffffu, m >> 16);
	const uint32_t neg = ((0u - mm) & 0xffffu) * 0x10001u;
#pragma unroll
	for (int q = 0; q < 32; q ++) R [q] = __vadd2 (R [q], neg);     // wraps per lane on purpose (subtraction)
}

// Stages the symbols of trellis steps [40k, 40k+40) of the CTA's 64 code words: tile[c][step] = the four 0..255
// symbols of the step (erasure = 127).  Source rows are read coalesced (a warp reads consecutive soft bits of one row)
// and scattered through the inverse puncturing map; with time de-interleaving the soft bit of code word c, position i
// comes from row c - D[i & 15] (dab-concurrent.cpp:162-169).
template <int NT, int VS_RB, int CW>
__device__ __forceinline__ void vs_stage (uint32_t *tile, const VitSimdJob &j, int k, int c0, int nrows, long long g_lo, int dmax) {
	const int tid = threadIdx. x, lane = tid & 31, warp = tid >> 5;
	__syncthreads ();
	for (int w = tid; w < CW * VS_CHUNK; w += NT)
		tile [(w / VS_CHUNK) * VS_ROWW + (w % VS_CHUNK)] = 0x7f7f7f7fu;          // erasure = 127 (deconvolve.cpp:185)
	__syncthreads ();
	const int m0 = 4 * VS_CHUNK * k;
	int i0, i1;
	if (j. inv) { i0 = j. chunk_i0 [k]; i1 = j. chunk_i0 [k + 1]; }
	else { i0 = m0; i1 = min (m0 + 4 * VS_CHUNK, 4 * j. nsteps); }
	uint8_t *tb = reinterpret_cast<uint8_t *> (tile);
	// VS_RB rows x up to 5 column groups of independent loads in flight per thread (the chunk has <= 160 columns):
	// the staging is latency bound, so memory-level parallelism is what matters here
	for (int row0 = warp * VS_RB; row0 < nrows; row0 += (NT / 32) * VS_RB) {
		int v [VS_RB][5];
#pragma unroll
		for (int rr = 0; rr < VS_RB; rr ++) {
			const int16_t *src = j. in + (g_lo + row0 + rr) * j. in_stride;
#pragma unroll
			for (int cc = 0; cc < 5; cc ++) {
				const int col = i0 + lane + 32 * cc;
				v [rr][cc] = (row0 + rr < nrows && col < i1) ? (int) __ldg (&src [col]) : 0;
			}
		}
#pragma unroll
		for (int cc = 0; cc < 5; cc ++) {
			const int col = i0 + lane + 32 * cc;
			if (col >= i1) continue;
			const int m = (j. inv ? (int) __ldg (&j. inv [col]) : col) - m0;
			const int dl = j. deint ? (int) c_vs_delay [col & 15] : 0;
#pragma unroll
			for (int rr = 0; rr < VS_RB; rr ++) {
				const int c = row0 + rr - dmax + dl;
				if (row0 + rr < nrows && c >= 0 && c < CW && c0 + c < j. ncw)
					tb [c * (VS_ROWW * 4) + m] = (uint8_t) min (max (v [rr][cc] + 127, 0), 255);   // viterbi.cpp:229-235
			}
		}
	}
	__syncthreads ();
}

__global__ void __launch_bounds__ (VS_THREADS) vit_simd_forward (const VitSimdJob *jobs, int njobs) {
	__shared__ uint32_t tile [VS_THREADS * VS_ROWW];
	// which job does this CTA belong to?
	int jb = 0;
	while (jb + 1 < njobs && (int) blockIdx. x >= jobs [jb + 1]. cta_first) jb ++;
	const VitSimdJob j = jobs [jb];
	const int c0 = ((int) blockIdx. x - j. cta_first) * VS_THREADS;       // first code word of this CTA
	const int tid = threadIdx. x;
	const int cw = c0 + tid;
	const bool live = cw < j. ncw;
	const int dmax = j. deint ? 15 : 0;
	const int nrows = min (VS_THREADS, j. ncw - c0) + dmax;
	const long long g_lo = (long long) c0 + j. first_row - dmax;          // first source row this CTA touches

	uint32_t R [32], Q [32];
	R [0] = 0u | (63u << 16);                                             // viterbi.cpp:364-370
#pragma unroll
	for (int q = 1; q < 32; q ++) R [q] = 63u | (63u << 16);

	uint2 *dec = j. dec + cw;
	const int nchunks = (j. nsteps + VS_CHUNK - 1) / VS_CHUNK;
	for (int k = 0; k < nchunks; k ++) {
		vs_stage<VS_THREADS, 16, VS_THREADS> (tile, j, k, c0, nrows, g_lo, dmax);
		// ---- 8 layout cycles of 5 steps ----
		const uint32_t *my = &tile [tid * VS_ROWW];
		uint2 *d = dec + (size_t) (VS_CHUNK * k) * j. ncw;
#pragma unroll 1
		for (int u = 0; u < VS_CHUNK / 5; u ++) {
			uint2 d0, d1, d2, d3, d4;
			vs_step<0> (R, Q, my [5 * u + 0], d0, j. one);
			vs_step<1> (Q, R, my [5 * u + 1], d1, j. one);
			vs_step<2> (R, Q, my [5 * u + 2], d2, j. one);
			vs_step<3> (Q, R, my [5 * u + 3], d3, j. one);
			vs_step<4> (R, Q, my [5 * u + 4], d4, j. one);
			vs_repack (Q, R);
			if (live) {
				d [(size_t) (5 * u + 0) * j. ncw] = d0; d [(size_t) (5 * u + 1) * j. ncw] = d1;
				d [(size_t) (5 * u + 2) * j. ncw] = d2; d [(size_t) (5 * u + 3) * j. ncw] = d3;
				d [(size_t) (5 * u + 4) * j. ncw] = d4;
			}
		}
		vs_renorm (R);
	}
}


// ---------------------------------------------------------------------------------------------------------------
// Two threads per code word.  One code word per thread leaves a 1024-frame batch with only ~8 warps per SM and every
// warp latency bound; splitting the 64 states over a lane pair doubles the warps in flight and halves the work per
// thread.  Layout: at step k of a 4-step cycle the half bit is state bit k and the lane bit is state bit k+1; register
// q (16 per thread) holds the remaining four bits.  Butterflies pair registers q and q+8 and write 2q, 2q+1, exactly
// as above.  After four steps the lane bit has reached s5 and the half bit s4; one round of 8 lane-pair shuffles and
// 16 PRMTs (plus a compile-time renaming) moves them back to s1 / s0.
// ---------------------------------------------------------------------------------------------------------------
#define VS2_CW      64          // code words per CTA (two threads each)
#define VS2_THREADS (2 * VS2_CW)
__host__ __device__ constexpr int vs_insert00 (int r, int p) { return ((r >> p) << (p + 2)) | (r & ((1 << p) - 1)); }

template <int K>
__device__ __forceinline__ void vs2_step (const uint32_t (&R) [16], uint32_t (&Q) [16], const uint32_t sym, const bool lanebit, uint32_t &dec, const uint32_t one) {
	const uint32_t s0 = sym & 255u, s1 = (sym >> 8) & 255u, s2 = (sym >> 16) & 255u, s3 = sym >> 24;
	const uint32_t a0 = s0 + s3, a1 = 510u - a0, b1 = 255u - s1, c1 = 255u - s2;
	uint32_t B [8];
	B [0] = a0 + s1 + s2; B [1] = a1 + s1 + s2; B [2] = a0 + b1 + s2; B [3] = a1 + b1 + s2;
	B [4] = a0 + s1 + c1; B [5] = a1 + s1 + c1; B [6] = a0 + b1 + c1; B [7] = a1 + b1 + c1;
	constexpr int dh = vs_pat (1 << K), dl = vs_pat (1 << (K + 1));
	uint32_t PM [8], PL [8];
#pragma unroll
	for (int x = 0; x < 8; x ++) PM [x] = B [x] + (B [x ^ dh] << 16);
#pragma unroll
	for (int x = 0; x < 8; x ++) PL [x] = lanebit ? PM [x ^ dl] : PM [x];
	uint32_t wlo = 0, whi = 0;
#pragma unroll
	for (int r = 0; r < 8; r ++) {
		const int x = vs_pat (vs_insert00 (r, K));
		const uint32_t a = R [r], b = R [r + 8];
		const uint32_t m0 = a + PL [x],     m1 = b + PL [7 - x];
		const uint32_t m2 = a + PL [7 - x], m3 = b + PL [x];
		Q [2 * r]     = vs_acs (m0, m1, wlo, whi, 1u << (2 * r), one);
		Q [2 * r + 1] = vs_acs (m2, m3, wlo, whi, 1u << (2 * r + 1), one);
	}
	dec = wlo | (whi << 16);
}

// lane = s5, half = s4, q = (s3 s2 s1 s0)  ->  lane = s1, half = s0, q = (s5 s4 s3 s2)
__device__ __forceinline__ void vs2_fixup (uint32_t (&Q) [16], uint32_t (&R) [16], const bool lanebit) {
#pragma unroll
	for (int qa = 0; qa < 16; qa ++) {
		if (qa & 2) continue;
		const int qb = qa | 2;                                  // lane (s5) <-> register bit 1 (s1)
		const uint32_t send = lanebit ? Q [qa] : Q [qb];
		const uint32_t recv = __shfl_xor_sync (0xffffffffu, send, 1);
		if (lanebit) Q [qa] = recv; else Q [qb] = recv;
	}
	uint32_t T [16];
#pragma unroll
	for (int qa = 0; qa < 16; qa += 2) {                         // half (s4) <-> register bit 0 (s0)
		T [qa]     = __byte_perm (Q [qa], Q [qa + 1], 0x5410);
		T [qa + 1] = __byte_perm (Q [qa], Q [qa + 1], 0x7632);
	}
	// T index bits (3,2,1,0) = (s3, s2, s5, s4); wanted (s5, s4, s3, s2)
#pragma unroll
	for (int q = 0; q < 16; q ++) R [q] = T [((q & 3) << 2) | (q >> 2)];
}

__global__ void __launch_bounds__ (VS2_THREADS) vit_simd2_forward (const VitSimdJob *jobs, int njobs) {
	__shared__ uint32_t tile [VS2_CW * VS_ROWW];
	int jb = 0;
	while (jb + 1 < njobs && (int) blockIdx. x >= jobs [jb + 1]. cta_first2) jb ++;
	const VitSimdJob j = jobs [jb];
	const int c0 = ((int) blockIdx. x - j. cta_first2) * VS2_CW;
	const int tid = threadIdx. x;
	const int cl = tid >> 1;                                   // code word within the CTA
	const bool lanebit = tid & 1;
	const int cw = c0 + cl;
	const bool live = cw < j. ncw;
	const int dmax = j. deint ? 15 : 0;
	const int nrows = min (VS2_CW, j. ncw - c0) + dmax;
	const long long g_lo = (long long) c0 + j. first_row - dmax;

	uint32_t R [16], Q [16];
#pragma unroll
	for (int q = 0; q < 16; q ++) R [q] = 63u | (63u << 16);
	if (!lanebit) R [0] = 0u | (63u << 16);                    // state 0 = lane 0, register 0, low half (viterbi.cpp:364-370)

	uint32_t *dec = reinterpret_cast<uint32_t *> (j. dec) + 2 * (size_t) cw + (lanebit ? 1 : 0);
	const size_t dstride = 2 * (size_t) j. ncw;
	const int nchunks = (j. nsteps + VS_CHUNK - 1) / VS_CHUNK;
	for (int k = 0; k < nchunks; k ++) {
		vs_stage<VS2_THREADS, 8, VS2_CW> (tile, j, k, c0, nrows, g_lo, dmax);
		const uint32_t *my = &tile [cl * VS_ROWW];
		uint32_t *d = dec + (size_t) (VS_CHUNK * k) * dstride;
#pragma unroll 1
		for (int u = 0; u < VS_CHUNK / 4; u ++) {
			uint32_t d0, d1, d2, d3;
			vs2_step<0> (R, Q, my [4 * u + 0], lanebit, d0, j. one);
			vs2_step<1> (Q, R, my [4 * u + 1], lanebit, d1, j. one);
			vs2_step<2> (R, Q, my [4 * u + 2], lanebit, d2, j. one);
			vs2_step<3> (Q, R, my [4 * u + 3], lanebit, d3, j. one);
			vs2_fixup (R, Q, lanebit);
#pragma unroll
			for (int q = 0; q < 16; q ++) R [q] = Q [q];
			if (live) {
				d [(size_t) (4 * u + 0) * dstride] = d0; d [(size_t) (4 * u + 1) * dstride] = d1;
				d [(size_t) (4 * u + 2) * dstride] = d2; d [(size_t) (4 * u + 3) * dstride] = d3;
			}
		}
		// renormalise: common minimum over the 64 states = both lanes of the pair
		uint32_t m = R [0];
#pragma unroll
		for (int q = 1; q < 16; q ++) m = __vminu2 (m, R [q]);
		uint32_t mm = min (m & 0xffffu, m >> 16);
		mm = min (mm, __shfl_xor_sync (0xffffffffu, mm, 1));
		const uint32_t neg = ((0u - mm) & 0xffffu) * 0x10001u;
#pragma unroll
		for (int q = 0; q < 16; q ++) R [q] = __vadd2 (R [q], neg);
	}
}

// chain-back (viterbi.cpp:333-357) + energy dispersal + unpack, one thread per code word.  The decision of new
// state n at step t sits in word h = bit p' of n, bit position = n with bit p' removed, p' = (t mod 5) + 1.
#define TB_THREADS 64
template <int LANES>
__global__ void __launch_bounds__ (TB_THREADS) vit_simd_traceback (const VitSimdJob *jobs, int njobs) {
	__shared__ uint32_t bits [TB_THREADS * 5];               // 128 decoded bits per code word per round, row stride 5 words
	int jb = 0;
	while (jb + 1 < njobs && (int) blockIdx. x >= jobs [jb + 1]. cta_first) jb ++;
	const VitSimdJob j = jobs [jb];
	const int c0 = ((int) blockIdx. x - j. cta_first) * TB_THREADS;
	const int tid = threadIdx. x, lane = tid & 31, warp = tid >> 5;
	const int cw = c0 + tid;
	const bool live = cw < j. ncw;
	const uint2 *dec = j. dec + (live ? cw : 0);
	unsigned state = 0;
	const int nrounds = (j. frameBits + 127) / 128;
	for (int rd = nrounds - 1; rd >= 0; rd --) {
		const int base = 128 * rd, top = min (j. frameBits, base + 128);
		uint32_t w [4] = { 0, 0, 0, 0 };
		if (live) {
			// the loads do not depend on the state: fetch 32 decision words (one output word), then walk them
#pragma unroll
			for (int wd = 3; wd >= 0; wd --) {
				const int hi = base + 32 * wd + 31;              // highest information bit of this output word
				if (base + 32 * wd >= top) continue;
				uint2 dd [32];
#pragma unroll
				for (int u = 0; u < 32; u ++) {
					const int i = hi - u;
					dd [u] = i < top ? dec [(size_t) (i + 6) * j. ncw] : make_uint2 (0u, 0u);
				}
				uint32_t acc = 0;
#pragma unroll
				for (int u = 0; u < 32; u ++) {
					const int i = hi - u;
					if (i >= top) continue;
					const int t = i + 6;                        // decision of step i+6 = information bit i
					unsigned bit;
					if (LANES == 1) {                            // one thread per code word: word = half, bit = state without bit pp
						const int pp = t % 5 + 1;
						const unsigned h = (state >> pp) & 1u;
						const unsigned q = ((state >> (pp + 1)) << pp) | (state & ((1u << pp) - 1u));
						bit = ((h ? dd [u]. y : dd [u]. x) >> q) & 1u;
					} else {                                     // lane pair: word = lane bit (state bit pp+1), bit = 16 half + state without bits pp, pp+1
						const int pp = (t & 3) + 1;
						const unsigned h = (state >> pp) & 1u, l = (state >> (pp + 1)) & 1u;
						const unsigned q = ((state >> (pp + 2)) << pp) | (state & ((1u << pp) - 1u));
						bit = ((l ? dd [u]. y : dd [u]. x) >> (16u * h + q)) & 1u;
					}
					state = (state >> 1) | (bit << 5);
					acc |= bit << (31 - u);
				}
				w [wd] = acc;
			}
		}
		__syncthreads ();
#pragma unroll
		for (int q = 0; q < 4; q ++) bits [tid * 5 + q] = w [q];
		__syncthreads ();
		// cooperative write-out: a warp writes the 128 bits of one code word as 128 bytes
		const int ncw_here = min (TB_THREADS, j. ncw - c0);
		for (int c = warp; c < ncw_here; c += TB_THREADS / 32) {
			const int i = base + 4 * lane;
			if (i < top) {
				uint32_t v = bits [c * 5 + (lane >> 3)];
				if (j. prbs) v ^= __ldg (&j. prbs [i >> 5]);
				v >>= (i & 31);
				uint8_t *o = j. out + (size_t) (c0 + c) * j. frameBits + i;
				if (i + 4 <= top && (j. frameBits & 3) == 0)
					*reinterpret_cast<uchar4 *> (o) = make_uchar4 (v & 1, (v >> 1) & 1, (v >> 2) & 1, (v >> 3) & 1);
				else
					for (int b = 0; b < 4 && i + b < top; b ++) o [b] = (v >> b) & 1;
			}
		}
	}
}

cudaError_t vit_simd_launch (dabgpu *h, const VitSimdJob *d_jobs, int njobs, int total_ctas, int total_ctas2) {
	if (njobs <= 0 || total_ctas <= 0) return cudaSuccess;
	cudaStream_t st = h -> vst ();
	const bool pair = h -> cfg. reserved [0] != 1;           // reserved[0] = 1: one thread per code word (kept for comparison)
	{ ProfScope prof (h, KC_VITERBI_MSC, st);
	  if (pair) vit_simd2_forward<<<total_ctas2, VS2_THREADS, 0, st>>> (d_jobs, njobs);
	  else      vit_simd_forward<<<total_ctas, VS_THREADS, 0, st>>> (d_jobs, njobs); }
	{ ProfScope prof (h, KC_VITERBI_TB, st);
	  if (pair) vit_simd_traceback<2><<<total_ctas, TB_THREADS, 0, st>>> (d_jobs, njobs);
	  else      vit_simd_traceback<1><<<total_ctas, TB_THREADS, 0, st>>> (d_jobs, njobs); }
	h -> launches += 2;
	return cudaGetLastError ();
}
