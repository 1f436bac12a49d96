// dabgpu_vit_simd.cu -- throughput Viterbi: TWO threads per terminated code word, the 64 path metrics of the word
// in 2 x 16 registers as packed 16-bit pairs, add-compare-select with the sm_100a packed integer instructions
// (32-bit IADD on carry-free pairs, VIMNMX.U16x2 whose two predicate outputs are the two decisions), decisions
// streamed to HBM as one coalesced word per thread and step, chain-back in a second kernel.  Used when a launch has
// enough code words to fill the GPU (the stream engine, large API batches); small batches use the warp-cooperative
// kernel of dabgpu_viterbi.cu.  Same arithmetic as the reference (viterbi.cpp:225-357, spiral-*.c), bit-exact:
//   * metrics start 63 / state 0 = 0; branch metric sum_j (Branchtab_j ^ sym_j), complement 1020 - m;
//   * VIMNMX's predicate is (upper <= lower): a tie keeps predecessor i, exactly the reference's strict '>';
//   * the common minimum is subtracted every 40 steps (differences, hence decisions, unchanged) so that the
//     16-bit lanes never wrap: spread <= 6*1020+63 and 40 steps add <= 40800  ->  < 2^16.
//
// Input path.  A streaming pre-pass turns the int16 soft bits of the rows a job reads into 0..255 symbols
// (viterbi.cpp:229-235) in a byte plane.  The forward kernel stages, per chunk of 40 trellis steps, the byte columns
// of that chunk for its CW + 15 source rows with 8-byte cp.async copies into a double-buffered shared-memory tile --
// no arithmetic at all -- and every thread then gathers the four symbols of a step through a per-profile table that
// already contains depuncturing (deconvolve.cpp:186-231) AND the time de-interleaver's row offset
// (dab-concurrent.cpp:162-169); a punctured position points at a pad column holding the erasure value 127.
//
// Register layout.  State s = (s5..s0).  At step k of a 4-step cycle the half bit is state bit k and the lane bit
// is state bit k+1; register q (16 per thread) holds the remaining four bits.  The butterfly pairs states differing
// in s5, i.e. registers q and q+8, same halves / lanes, so one packed op does two butterflies; the results 2i and
// 2i+1 go to registers 2q and 2q+1 and both special bits move up by one.  After four steps the lane bit has reached
// s5 and the half bit s4; one round of 8 lane-pair shuffles and 16 PRMTs (plus a compile-time renaming) moves them
// back to s1 / s0.  All indices are compile-time, the code is unrolled over the four-step cycle.
#include "dabgpu_internal.h"

#ifndef VS_CW
#define VS_CW      32          // code words per forward CTA (two threads each)
#endif
#ifndef VS_MINB
#define VS_MINB    (320 / VS_CW)   // resident CTAs per SM the register budget is cut for
#endif
#define VS_THREADS (2 * VS_CW)
#define VS_ROWS    (VS_CW + 15)
#define VS_TILE    (VS_ROWS * VS_PITCH)

__host__ __device__ constexpr int vs_parity (unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1; }
// branch pattern of butterfly i: bit0 = polys 0155 (used twice: j = 0 and 3), bit1 = 0117, bit2 = 0123 (viterbi.cpp:63, 159-164)
__host__ __device__ constexpr int vs_pat (int i) {
	return vs_parity ((2u * i) & 0155u) | (vs_parity ((2u * i) & 0117u) << 1) | (vs_parity ((2u * i) & 0123u) << 2);
}
__host__ __device__ constexpr int vs_insert00 (int r, int p) { return ((r >> p) << (p + 2)) | (r & ((1 << p) - 1)); }

// packed compare-select: min (upper, lower) per 16-bit lane; `bit` is ORed into wlo / whi where the lower candidate
// won (decision = upper > lower; a tie keeps the upper one).  __vibmin_u16x2 is one VIMNMX.U16x2 with two predicate
// outputs (predicate = upper <= lower) on sm_100a.
__device__ __forceinline__ uint32_t vs_acs (uint32_t upper, uint32_t lower, uint32_t &acc, const uint32_t bit_lo, const uint32_t bit_hi, const uint32_t one) {
	bool ph, pl;
	const uint32_t r = __vibmin_u16x2 (upper, lower, &ph, &pl);
	// decision bits are accumulated with predicated multiply-adds (acc += one * bit): `one` is 1 at run time but opaque
	// to the compiler, which keeps these on the FMA pipe; the ALU pipe, busy with the packed min and everything else,
	// is the scarce one here
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p mad.lo.u32 %0, %3, %2, %0; }" : "+r" (acc) : "r" ((uint32_t) pl), "r" (bit_lo), "r" (one));
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p mad.lo.u32 %0, %3, %2, %0; }" : "+r" (acc) : "r" ((uint32_t) ph), "r" (bit_hi), "r" (one));
	return r;
}

// Where the decision of a new state goes.  At step K the decision word of a thread (lane bit = new-state bit K + 2) holds
// the 32 new states that share this lane bit; the bit index is the new state number n with bit K + 2 deleted -- the form
// the chain-back can rebuild with one shift and one select-by-mask.  Here n = (r without/with the half and lane bits) :
// n0 = e (even / odd result), bits 1..K = low K bits of r, bit K + 1 = half, bits above = the rest of r.
__host__ __device__ constexpr uint32_t vs_decbit (int K, int r, int e, int half) {
	return 1u << (e | ((r & ((1 << K) - 1)) << 1) | (half << (K + 1)) | ((r >> K) << (K + 2)));
}

__host__ __device__ constexpr int vs_decidx (int K, int r, int e, int half) {
	return e | ((r & ((1 << K) - 1)) << 1) | (half << (K + 1)) | ((r >> K) << (K + 2));
}
// The same decisions collected in FLOAT accumulators: acc += 2^(index mod 16) under the predicate, one accumulator group
// for the bit positions below 16 and one for those above; every accumulator receives four distinct powers of two inside a
// 16-bit window, so the sums are exact, and two float-to-int conversions per step turn them back into the decision word.
// Unlike the integer multiply-adds, which ptxas partly rewrites into select + three-input add on the (half-rate, busy)
// ALU pipe, predicated FFMAs stay on the FMA pipe.
#ifndef VS_FDEC
#define VS_FDEC 0
#endif
#ifndef VS_PL2
#define VS_PL2 0
#endif
__device__ __forceinline__ uint32_t vs_acs_f (uint32_t upper, uint32_t lower, float &acc_l, float &acc_h, const float c_l, const float c_h, const float onef) {
	bool ph, pl;
	const uint32_t r = __vibmin_u16x2 (upper, lower, &ph, &pl);
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p fma.rn.f32 %0, %3, %2, %0; }" : "+f" (acc_l) : "r" ((uint32_t) pl), "f" (c_l), "f" (onef));
	asm ("{ .reg .pred p; setp.eq.u32 p, %1, 0; @p fma.rn.f32 %0, %3, %2, %0; }" : "+f" (acc_h) : "r" ((uint32_t) ph), "f" (c_h), "f" (onef));
	return r;
}

__device__ __forceinline__ uint32_t vs_sel (uint32_t a, uint32_t b, uint32_t mask) { return (a & ~mask) | (b & mask); }   // one LOP3

// One trellis step.  a0 = sym0 + sym3 (both use polynomial 0155), s1, s2 = the other two symbols.  The branch metric
// of pattern x is B[x] = sum_j (x_j ? C_j - v_j : v_j) with (v, C) = (a0, 510), (s1, 255), (s2, 255); a register needs
// it for the low-half state and, 16 bits up, for the high-half state (pattern x ^ dh); the odd lane of a pair sees
// pattern x ^ dl (lm = all ones in the odd lane, 0 in the even one).  Built per component, so that PL[x] is a sum of
// three terms.  The 32 decisions of the step are disjoint bits spread over VS_NACC accumulators (the predicates of the
// packed min are consumed about as fast as they are produced; see VS_NACC for the count).
template <int K>
__device__ __forceinline__ void vs_step (const uint32_t (&R) [16], uint32_t (&Q) [16], const uint32_t a0, const uint32_t s1, const uint32_t s2,
                                         const uint32_t lm, uint32_t &dec, const uint32_t one) {
	constexpr int dh = vs_pat (1 << K), dl = vs_pat (1 << (K + 1));
	uint32_t c [3][2];
#pragma unroll
	for (int j = 0; j < 3; j ++) {
		const uint32_t v = j == 0 ? a0 : j == 1 ? s1 : s2;
		const uint32_t C = j == 0 ? 510u : 255u;
		const uint32_t c0 = ((dh >> j) & 1) ? ((C - v) << 16) + v : v * 0x00010001u;     // low: v, high: C - v or v
		const uint32_t c1 = C * 0x00010001u - c0;
		if ((dl >> j) & 1) { c [j][0] = vs_sel (c0, c1, lm); c [j][1] = vs_sel (c1, c0, lm); }
		else               { c [j][0] = c0; c [j][1] = c1; }
	}
	uint32_t PL [8];
#if VS_PL2
	// two levels of two-input adds (12 instead of 8 three-input ones): two-input adds can go to the FMA pipe as IMAD.IADD
	uint32_t c01 [4];
#pragma unroll
	for (int x = 0; x < 4; x ++) c01 [x] = c [0][x & 1] + c [1][x >> 1];
#pragma unroll
	for (int x = 0; x < 8; x ++) PL [x] = c01 [x & 3] + c [2][x >> 2];
#else
#pragma unroll
	for (int x = 0; x < 8; x ++)                                // one three-input add each (kept from being split into shared partial sums)
		asm ("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r" (PL [x]) : "r" (c [0][x & 1]), "r" (c [1][(x >> 1) & 1]), "r" (c [2][(x >> 2) & 1]));
#endif
#if VS_FDEC
	const float onef = __uint_as_float (one * 0x3f800000u);     // 1.0f, opaque like `one`
	float g [8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };      // [0..3]: bit positions 0..15, [4..7]: positions 16..31
#pragma unroll
	for (int r = 0; r < 8; r ++) {
		const int x = vs_pat (vs_insert00 (r, K));
		const uint32_t a = R [r], b = R [r + 8];
		const uint32_t m0 = a + PL [x],     m1 = b + PL [7 - x];
		const uint32_t m2 = a + PL [7 - x], m3 = b + PL [x];
		const int slot = K < 3 ? (r & 3) : (r >> 1);                // every accumulator ends up with four decisions
#pragma unroll
		for (int e = 0; e < 2; e ++) {
			const int il = vs_decidx (K, r, e, 0), ih = vs_decidx (K, r, e, 1);
			Q [2 * r + e] = vs_acs_f (e ? m2 : m0, e ? m3 : m1, g [4 * (il >> 4) + slot], g [4 * (ih >> 4) + slot],
			                          (float) (1u << (il & 15)), (float) (1u << (ih & 15)), onef);
		}
	}
	dec = __float2uint_rz ((g [0] + g [1]) + (g [2] + g [3])) + (__float2uint_rz ((g [4] + g [5]) + (g [6] + g [7])) << 16);
#else
#ifndef VS_NACC
#define VS_NACC 2                               // accumulators of decision bits per step (A/B on B200: 8 -> 1.428, 4 -> 1.413, 2 -> 1.386 ms per 1024 frames; 1 makes ptxas spill predicates)
#endif
	uint32_t acc [8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
	for (int r = 0; r < 8; r ++) {
		const int x = vs_pat (vs_insert00 (r, K));
		const uint32_t a = R [r], b = R [r + 8];
		// plain 32-bit adds: every 16-bit lane stays below 2^16 (see the renormalisation bound), so no carry crosses
		// the halves and the full-rate IADD replaces the half-rate VIADD.16x2
		const uint32_t m0 = a + PL [x],     m1 = b + PL [7 - x];
		const uint32_t m2 = a + PL [7 - x], m3 = b + PL [x];
		Q [2 * r]     = vs_acs (m0, m1, acc [r % VS_NACC], vs_decbit (K, r, 0, 0), vs_decbit (K, r, 0, 1), one);
		Q [2 * r + 1] = vs_acs (m2, m3, acc [r % VS_NACC], vs_decbit (K, r, 1, 0), vs_decbit (K, r, 1, 1), one);
	}
	dec = ((acc [0] + acc [1]) + (acc [2] + acc [3])) + ((acc [4] + acc [5]) + (acc [6] + acc [7]));
#endif
}

// lane = s5, half = s4, q = (s3 s2 s1 s0)  ->  lane = s1, half = s0, q = (s5 s4 s3 s2)
__device__ __forceinline__ void vs_fixup (uint32_t (&Q) [16], uint32_t (&R) [16], const uint32_t lm) {
#pragma unroll
	for (int qa = 0; qa < 16; qa ++) {
		if (qa & 2) continue;
		const int qb = qa | 2;                                  // lane (s5) <-> register bit 1 (s1)
		const uint32_t recv = __shfl_xor_sync (0xffffffffu, vs_sel (Q [qb], Q [qa], lm), 1);
		Q [qa] = vs_sel (Q [qa], recv, lm); Q [qb] = vs_sel (recv, Q [qb], lm);
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

// int16 soft bits -> 0..255 symbols (viterbi.cpp:229-235) for the rows and columns the jobs of a launch read.
// blockIdx.y = job; sym8 row r = source row first_row - dmax + r.
__global__ void __launch_bounds__ (256) vit_sym8_kernel (const VitSimdJob *jobs) {
	const VitSimdJob j = jobs [blockIdx. y];
	if (j. sym8_ready) return;
	const int dmax = j. deint ? 15 : 0, nrows = j. ncw + dmax;
	const int wpr = (int) (j. stride8 >> 2);                      // 4-byte words per row
	const long long total = (long long) nrows * wpr;
	const int16_t *src0 = j. in + (long long) (j. first_row - dmax) * j. in_stride;
	for (long long w = (long long) blockIdx. x * blockDim. x + threadIdx. x; w < total; w += (long long) gridDim. x * blockDim. x) {
		const int r = (int) (w / wpr), c = 4 * (int) (w % wpr);
		const int16_t *src = src0 + (long long) r * j. in_stride + c;
		uint32_t o = 0;
#pragma unroll
		for (int b = 0; b < 4; b ++) {
			const int v = c + b < j. ncols ? (int) __ldg (&src [b]) : 0;
			o |= (uint32_t) min (max (v + 127, 0), 255) << (8 * b);
		}
		*reinterpret_cast<uint32_t *> (j. sym8 + (long long) r * j. stride8 + c) = o;
	}
}

__device__ __forceinline__ void vs_cp_async8 (void *smem, const void *gmem) {
	asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r" ((uint32_t) __cvta_generic_to_shared (smem)), "l" (gmem));
}

__device__ __forceinline__ void vs_cp_async16 (void *smem, const void *gmem) {
	asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r" ((uint32_t) __cvta_generic_to_shared (smem)), "l" (gmem));
}

__global__ void __launch_bounds__ (VS_THREADS, VS_MINB) vit_simd_forward (const VitSimdJob *jobs, int njobs) {
	__shared__ __align__ (16) uint8_t tile [2 * VS_TILE];
	int jb = 0;                                                // the job this CTA belongs to: last one with cta_first2 <= blockIdx (hundreds of jobs with many streams: bisection)
	for (int hi = njobs; hi - jb > 1; ) { const int mid = (jb + hi) >> 1; if ((int) blockIdx. x >= __ldg (&jobs [mid]. cta_first2)) jb = mid; else hi = mid; }
	const VitSimdJob j = jobs [jb];
	const int c0 = ((int) blockIdx. x - j. cta_first2) * VS_CW;   // first code word of this CTA
	const int tid = threadIdx. x;
	const int cl = tid >> 1;                                   // code word within the CTA
	const bool lanebit = tid & 1;
	uint32_t lm = 0u - (uint32_t) (tid & 1);
	asm volatile ("" : "+r" (lm));                               // keep the lane select a bit mask (LOP3), not a predicate: predicates are scarce here
	const int cw = c0 + cl;
	const bool live = cw < j. ncw;
	const int dmax = j. deint ? 15 : 0;
	const int nrows = min (VS_CW, j. ncw - c0) + dmax;
	const uint8_t *src = j. sym8 + (long long) c0 * j. stride8;    // tile row 0 = sym8 row c0 = source row of code word c0 minus dmax

	// chunk k -> tile buffer k & 1: the 8-byte granules g < G_k of every row, one (row, granule) per thread and round
	auto stage = [&] (int k) {
		const int2 ch = __ldg (&j. chunk [k]);
		uint8_t *dst = tile + (k & 1) * VS_TILE;
		for (int idx = tid; idx < nrows * 32; idx += VS_THREADS) {
			const int r = idx >> 5, g = idx & 31;
			if (g < ch. y) vs_cp_async8 (dst + r * VS_PITCH + 8 * g, src + (long long) r * j. stride8 + ch. x + 8 * g);
		}
		asm volatile ("cp.async.commit_group;");
	};
	for (int r = tid; r < 2 * VS_ROWS; r += VS_THREADS) tile [r * VS_PITCH + VS_PITCH - 1] = 127;   // erasure (deconvolve.cpp:185)
	stage (0);

	uint32_t R [16], Q [16];
#pragma unroll
	for (int q = 0; q < 16; q ++) R [q] = 63u | (63u << 16);
	if (!lanebit) R [0] = 0u | (63u << 16);                    // state 0 = lane 0, register 0, low half (viterbi.cpp:364-370)

	const int nchunks = (j. nsteps + VS_CHUNK - 1) / VS_CHUNK;
	// decision words: [block of 32 code words][step pair][32 code words] uint4 = {even step lane 0, odd step lane 0,
	// even step lane 1, odd step lane 1}: a CTA writes one sequential stream (a thread stores the words of two steps as
	// 8 bytes), and the chain-back fetches 16 bytes per thread and step pair, 16 KB contiguous per warp and output word
	uint32_t *dec = reinterpret_cast<uint32_t *> (j. dec) + 4 * ((size_t) (cw >> 5) * (size_t) (vs_npad (j. nsteps) / 2) * 32 + (cw & 31)) + (lanebit ? 2 : 0);
	for (int k = 0; k < nchunks; k ++) {
		if (k + 1 < nchunks) stage (k + 1); else asm volatile ("cp.async.commit_group;");
		asm volatile ("cp.async.wait_group 1;");
		__syncthreads ();
		const uint8_t *my = tile + (k & 1) * VS_TILE + cl * VS_PITCH;
		const uint4 *ga = j. gather + VS_CHUNK * k;
		uint32_t *d = dec + (size_t) (VS_CHUNK / 2 * k) * 128;
		// two-deep software pipeline: while step s is processed, the symbols of step s + 1 are read from the tile and the
		// gather entry of step s + 2 is on its way from the table
		uint32_t sa, sb, sc;
		uint4 en;
		{ const uint4 e = __ldg (&ga [0]); sa = (uint32_t) my [e. x] + my [e. w]; sb = my [e. y]; sc = my [e. z]; en = __ldg (&ga [1]); }
#pragma unroll 1
		for (int u = 0; u < VS_CHUNK / 4; u ++) {
			uint32_t d0, d1, d2, d3, na, nb, nc;
			uint4 enn;
#define VS_NEXT(i) enn = __ldg (&ga [(i) + 1]); na = (uint32_t) my [en. x] + my [en. w]; nb = my [en. y]; nc = my [en. z];
#define VS_ROLL    sa = na; sb = nb; sc = nc; en = enn;
			VS_NEXT (4 * u + 1); vs_step<0> (R, Q, sa, sb, sc, lm, d0, j. one); VS_ROLL
			VS_NEXT (4 * u + 2); vs_step<1> (Q, R, sa, sb, sc, lm, d1, j. one); VS_ROLL
			VS_NEXT (4 * u + 3); vs_step<2> (R, Q, sa, sb, sc, lm, d2, j. one); VS_ROLL
			VS_NEXT (4 * u + 4); vs_step<3> (Q, R, sa, sb, sc, lm, d3, j. one); VS_ROLL     // (up to two entries past the chunk are read, not used)
#undef VS_NEXT
#undef VS_ROLL
			vs_fixup (R, Q, lm);
#pragma unroll
			for (int q = 0; q < 16; q ++) R [q] = Q [q];
			if (live) {
				// (four 32-bit stores on purpose: pairing them into 64-bit stores changes the schedule enough to make ptxas spill
				// predicates -- the decisions -- into a register, +25 % instructions)
				d [(size_t) (2 * u) * 128] = d0;     d [(size_t) (2 * u) * 128 + 1] = d1;
				d [(size_t) (2 * u + 1) * 128] = d2; d [(size_t) (2 * u + 1) * 128 + 1] = d3;
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
		__syncthreads ();                                        // everyone is done with buffer k & 1 before chunk k + 2 lands in it
	}
}

// chain-back (viterbi.cpp:333-357) + energy dispersal + unpack, one thread per code word.  The decision of new
// state n at step t sits in the word of lane l = bit p+1 of n, at bit (n with bit p+1 deleted), p = (t mod 4) + 1
// (vs_decbit).  The walk keeps the decoded bits in a 32-bit shift register S whose top six bits ARE the state
// (newest bit on top, viterbi.cpp:347-352): one step is  n = S >> 26;  word = lane-select;  bit = word >> index (n);
// S = (bit : S) >> 1  -- a funnel shift -- and after 32 steps S is the output word, bit-reversed.
#ifndef TB_THREADS
#define TB_THREADS 64
#endif
#ifndef TB_NBUF
#define TB_NBUF 2                      // output words whose decision words are resident or in flight per thread (A/B: 3 and 4 are slower, as are 32-thread CTAs: 0.188 / 0.195 / 0.201 ms)
#endif
#if TB_NBUF == 2
#define TB_WAIT "1"
#elif TB_NBUF == 3
#define TB_WAIT "2"
#else
#define TB_WAIT "3"
#endif
template <bool FULL>
__device__ __forceinline__ uint32_t tb_walk_word (uint32_t &S, const uint4 (*dq) [TB_THREADS], const int tid, const int nvalid) {
#pragma unroll
	for (int m = 0; m < 16; m ++) {                         // slot m = step pair (32 wi + 36 - 2 m, + 1): steps u = 2 m (odd step) and 2 m + 1 (even step)
		if (!FULL && 2 * m + 1 < 32 - nvalid) continue;     // a partial top word: slots of bits beyond frameBits are not there
		const uint4 dd = dq [m][tid];
#pragma unroll
		for (int e = 0; e < 2; e ++) {
			const int u = 2 * m + e;
			if (!FULL && u < 32 - nvalid) continue;
			const int pp = ((37 - u) & 3) + 1;              // step t = 32 wi + 37 - u
			const uint32_t n = S >> 26, n1 = S >> 27;
			const uint32_t lowmask = (1u << (pp + 1)) - 1u;
			const uint32_t idx = (n & lowmask) | (n1 & ~lowmask);
			const bool lane1 = (S & (1u << (26 + pp + 1))) != 0;
			const uint32_t word = e == 0 ? (lane1 ? dd. w : dd. y) : (lane1 ? dd. z : dd. x);
			S = __funnelshift_r (S, word >> (idx & 31u), 1);
		}
	}
	return __brev (S);
}

__global__ void __launch_bounds__ (TB_THREADS) vit_simd_traceback (const VitSimdJob *jobs, int njobs) {
	__shared__ uint32_t bits [TB_THREADS * 5];               // 128 decoded bits per code word per round, row stride 5 words
	__shared__ __align__ (16) uint4 dq [TB_NBUF][16][TB_THREADS];  // the decision words of TB_NBUF output words (16 step pairs each) per thread
	int jb = 0;
	for (int hi = njobs; hi - jb > 1; ) { const int mid = (jb + hi) >> 1; if ((int) blockIdx. x >= __ldg (&jobs [mid]. cta_first)) jb = mid; else hi = mid; }
	const VitSimdJob j = jobs [jb];
	const int c0 = ((int) blockIdx. x - j. cta_first) * TB_THREADS;
	const int tid = threadIdx. x, lane = tid & 31, warp = tid >> 5;
	const int cw = c0 + tid;
	const bool live = cw < j. ncw;
	const uint4 *dec = reinterpret_cast<const uint4 *> (j. dec) + (live ? (size_t) (cw >> 5) * (size_t) (vs_npad (j. nsteps) / 2) * 32 + (cw & 31) : 0);
	uint32_t S = 0;                                          // chain-back starts in state 0 (viterbi.cpp:340)
	const int nrounds = (j. frameBits + 127) / 128, nwords = (j. frameBits + 31) / 32;
	// The loads do not depend on the state: while output word wi is walked, the 32 decision words of word wi - 1 are
	// already on their way into shared memory (cp.async; every thread consumes only what it copied itself, so no
	// barrier is involved).  Slot u of a word holds information bit 32 wi + 31 - u = trellis step 32 wi + 37 - u.
	auto fetch = [&] (int wi) {
		if (wi >= 0 && live) {
			const uint4 *src = dec + (size_t) (16 * wi + 18) * 32;      // pair of steps 32 wi + 36, 37
			if (32 * wi + 32 <= j. frameBits) {
#pragma unroll
				for (int m = 0; m < 16; m ++) vs_cp_async16 (&dq [wi % TB_NBUF][m][tid], src - 32 * m);
			} else {
#pragma unroll
				for (int m = 0; m < 16; m ++)
					if (32 * wi + 30 - 2 * m < j. frameBits) vs_cp_async16 (&dq [wi % TB_NBUF][m][tid], src - 32 * m);
			}
		}
		asm volatile ("cp.async.commit_group;");
	};
#pragma unroll
	for (int k = 1; k < TB_NBUF; k ++) fetch (nwords - k);
	for (int rd = nrounds - 1; rd >= 0; rd --) {
		const int base = 128 * rd, top = min (j. frameBits, base + 128);
		uint32_t w [4] = { 0, 0, 0, 0 };
#pragma unroll
		for (int wd = 3; wd >= 0; wd --) {
			const int wi = 4 * rd + wd;
			if (wi >= nwords) continue;
			fetch (wi - (TB_NBUF - 1));
			asm volatile ("cp.async.wait_group " TB_WAIT ";" ::: "memory");
			if (!live) continue;
			const int nvalid = min (32, j. frameBits - 32 * wi);
			// (a partial word can only be the first one walked: S is still 0 below the bits it shifts in)
			w [wd] = nvalid == 32 ? tb_walk_word<true> (S, dq [wi % TB_NBUF], tid, 32) : tb_walk_word<false> (S, dq [wi % TB_NBUF], tid, nvalid);
		}
		if (j. packed) {
			// packed output (dabgpu_set_msc_output): 8 bits per byte, first bit in the byte's top bit -- the thread writes the four
			// words it has just decoded itself (frameBits is a multiple of 32 here)
			if (live) {
				uint32_t *o = reinterpret_cast<uint32_t *> (j. out + (size_t) cw * (size_t) (j. frameBits >> 3)) + 4 * rd;
#pragma unroll
				for (int q = 0; q < 4; q ++)
					if (4 * rd + q < nwords) {
						uint32_t v = w [q];
						if (j. prbs) v ^= __ldg (&j. prbs [4 * rd + q]);
						o [q] = __byte_perm (__brev (v), 0, 0x0123);
					}
			}
			continue;
		}
		__syncthreads ();
#pragma unroll
		for (int q = 0; q < 4; q ++) bits [tid * 5 + q] = w [q];
		__syncthreads ();
		// cooperative write-out, one bit per byte (viterbi.cpp:240-241) with the dispersal sequence applied
		// (dab-concurrent.cpp:183-190).  Fast path: 8 lanes x 16 bytes cover the 128 bits of a code word, a warp does four
		// code words per pass; 4 bits -> 4 bytes is one multiply and one mask.
		const int ncw_here = min (TB_THREADS, j. ncw - c0);
		if ((j. frameBits & 15) == 0 && (reinterpret_cast<size_t> (j. out) & 15) == 0) {
			const int sub = lane & 7, i = base + 16 * sub;
			uint32_t pr = 0;
			if (j. prbs && i < top) pr = (__ldg (&j. prbs [i >> 5]) >> (i & 31)) & 0xffffu;
			for (int c = 4 * warp + (lane >> 3); c < ncw_here; c += 4 * (TB_THREADS / 32)) {
				if (i >= top) continue;
				const uint32_t v = ((bits [c * 5 + (sub >> 1)] >> (16 * (sub & 1))) & 0xffffu) ^ pr;
				uint4 o;
				o. x = ((v & 15u) * 0x00204081u) & 0x01010101u;         o. y = (((v >> 4) & 15u) * 0x00204081u) & 0x01010101u;
				o. z = (((v >> 8) & 15u) * 0x00204081u) & 0x01010101u;  o. w = (((v >> 12) & 15u) * 0x00204081u) & 0x01010101u;
				*reinterpret_cast<uint4 *> (j. out + (size_t) (c0 + c) * j. frameBits + i) = o;
			}
		} else {
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
}

int vit_simd_cw_per_cta () { return VS_CW; }
int vit_simd_tb_cw_per_cta () { return TB_THREADS; }

cudaError_t vit_simd_launch (dabgpu *h, const VitSimdJob *d_jobs, int njobs, int total_ctas, int total_ctas2, bool convert) {
	if (njobs <= 0 || total_ctas <= 0) return cudaSuccess;
	cudaStream_t st = h -> vst ();
	if (convert) {
		ProfScope prof (h, KC_VITERBI_SYM, st);
		vit_sym8_kernel<<<dim3 (296, njobs), 256, 0, st>>> (d_jobs);
		h -> launches ++;
	}
	{ ProfScope prof (h, KC_VITERBI_MSC, st);
	  vit_simd_forward<<<total_ctas2, VS_THREADS, 0, st>>> (d_jobs, njobs); }
	{ ProfScope prof (h, KC_VITERBI_TB, st);
	  vit_simd_traceback<<<total_ctas, TB_THREADS, 0, st>>> (d_jobs, njobs); }
	h -> launches += 2;
	return cudaGetLastError ();
}
