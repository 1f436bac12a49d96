// dabgpu_viterbi.cu -- Viterbi group: (time de-interleave gather) + depuncture + K=7 rate-1/4 64-state
// soft-decision Viterbi + energy dispersal, one terminated code word per warp.  sm_100a.
//
// Replaces viterbi::deconvolve (viterbi.cpp:225-242), the SPIRAL add-compare-select loops
// (spiral-no-sse.c / spiral-sse.c), chainback_viterbi (viterbi.cpp:333-357), the depuncturing loops of
// eep_/uep_deconvolve (deconvolve.cpp:186-231, 338-362) and ficHandler::process_ficInput
// (fic-handler.cpp:254-302), the time de-interleaver and dispersal of dabConcurrent::run
// (dab-concurrent.cpp:162-190).
//
// Bit-exactness notes: metrics start at 63 (state 0: 0); branch metric
// = sum_j (Branchtab_j ^ sym_j), complement 1020 - m; decision = strict "upper > lower" so ties keep
// predecessor i; exactly frameBits+6 steps; chain-back from state 0.  Metrics are renormalised by
// subtracting the common minimum every 32 steps, which leaves every metric DIFFERENCE (hence every
// decision) unchanged while keeping them below 2^16 so two of them travel in one shuffle.
#include "dabgpu_internal.h"

__constant__ int8_t c_deint_delay [16] = { 15, 7, 11, 3, 13, 5, 9, 1, 14, 6, 10, 2, 12, 4, 8, 0 };  // dab-concurrent.cpp:41-43

// soft value -> 0..255 symbol (viterbi.cpp:229-235)
__device__ __forceinline__ unsigned soft_to_sym (int v) {
	return (unsigned) min (max (v + 127, 0), 255);
}

template <bool DEINT>
__device__ __forceinline__ unsigned load_step_symbols (const VitJob &j, const int16_t *src, const uint8_t *src8, int blk, int t) {
	if (t >= j. nsteps) return 0x7f7f7f7fu;
	int idx [4];
	if (j. lut) {
		const ushort4 l = reinterpret_cast<const ushort4 *> (j. lut) [t];
		idx [0] = l. x == 0xFFFF ? -1 : l. x; idx [1] = l. y == 0xFFFF ? -1 : l. y;
		idx [2] = l. z == 0xFFFF ? -1 : l. z; idx [3] = l. w == 0xFFFF ? -1 : l. w;
	} else {
		idx [0] = 4 * t; idx [1] = 4 * t + 1; idx [2] = 4 * t + 2; idx [3] = 4 * t + 3;
	}
	unsigned s = 0;
#pragma unroll
	for (int k = 0; k < 4; k ++) {
		unsigned sym = 127;                           // punctured position = erasure (deconvolve.cpp:185)
		if (idx [k] >= 0) {
			const long long o = (long long) idx [k] - (DEINT ? (long long) c_deint_delay [idx [k] & 15] * j. in_stride : 0ll);
			sym = src8 ? (unsigned) __ldg (src8 + o) : soft_to_sym (__ldg (src + o));     // in8: already 0..255 symbols
		}
		s |= sym << (8 * k);
	}
	return s;
}

template <bool DEINT>
__global__ void __launch_bounds__ (128) vit_warp_kernel (const VitJob j, const int smem_words_per_warp) {
	extern __shared__ uint32_t smem [];
	const int lane = threadIdx. x & 31, warp = threadIdx. x >> 5;
	const int blk = blockIdx. x * (blockDim. x >> 5) + warp;
	if (blk >= j. nblocks) return;
	uint2    *dec   = reinterpret_cast<uint2 *> (smem + (size_t) warp * smem_words_per_warp);
	uint32_t *bitsw = reinterpret_cast<uint32_t *> (dec + j. nsteps);

	const long long row = (long long) (blk + (DEINT ? j. first_row : 0)) * j. in_stride;
	const int16_t *src = j. in ? j. in + row : nullptr;
	const uint8_t *src8 = j. in8 ? j. in8 + row : nullptr;

	// lane = butterfly i: Branchtab_j[i] = parity ((2i) & poly_j) ? 255 : 0 (viterbi.cpp:159-164)
	unsigned xmask = 0;
	{
		const unsigned polys [4] = { 0155, 0117, 0123, 0155 };
#pragma unroll
		for (int k = 0; k < 4; k ++)
			if (__popc ((2u * lane) & polys [k]) & 1) xmask |= 0xffu << (8 * k);
	}
	unsigned a = lane == 0 ? 0u : 63u, b = 63u;       // old[lane], old[lane + 32] (viterbi.cpp:364-370)
	const int srcA = lane >> 1, srcB = (lane >> 1) + 16, sh = (lane & 1) * 16;

	unsigned nxt = load_step_symbols<DEINT> (j, src, src8, blk, lane);
	for (int t0 = 0; t0 < j. nsteps; t0 += 32) {
		const unsigned cur = nxt;
		nxt = load_step_symbols<DEINT> (j, src, src8, blk, t0 + 32 + lane);
		const int n = min (32, j. nsteps - t0);
		for (int k = 0; k < n; k ++) {
			const unsigned S = __shfl_sync (0xffffffffu, cur, k);
			const unsigned m  = __dp4a (S ^ xmask, 0x01010101u, 0u);
			const unsigned mc = 1020u - m;
			const unsigned m0 = a + m,  m1 = b + mc;      // -> state 2i   (viterbi.cpp:258-268)
			const unsigned m2 = a + mc, m3 = b + m;       // -> state 2i+1
			const bool d0 = m0 > m1, d1 = m2 > m3;
			const unsigned n0 = d0 ? m1 : m0, n1 = d1 ? m3 : m2;
			const unsigned e = __ballot_sync (0xffffffffu, d0);   // bit i = decision of state 2i
			const unsigned o = __ballot_sync (0xffffffffu, d1);   // bit i = decision of state 2i+1
			if (lane == 0) dec [t0 + k] = make_uint2 (e, o);
			const unsigned packed = n0 | (n1 << 16);
			a = (__shfl_sync (0xffffffffu, packed, srcA) >> sh) & 0xffffu;
			b = (__shfl_sync (0xffffffffu, packed, srcB) >> sh) & 0xffffu;
		}
		const unsigned mn = __reduce_min_sync (0xffffffffu, min (a, b));
		a -= mn; b -= mn;
	}
	__syncwarp ();

	// chain-back from state 0 (viterbi.cpp:343-356): decision of step i+6 on the survivor = info bit i,
	// predecessor = (state >> 1) | (bit << 5).  Loads do not depend on the state, so they pipeline.
	if (lane == 0) {
		unsigned state = 0, word = 0;
#pragma unroll 8
		for (int i = j. frameBits - 1; i >= 0; i --) {
			const uint2 d = dec [i + 6];
			const unsigned w = (state & 1) ? d. y : d. x;
			const unsigned k = (w >> (state >> 1)) & 1u;
			state = (state >> 1) | (k << 5);
			word |= k << (i & 31);
			if ((i & 31) == 0) { bitsw [i >> 5] = word; word = 0; }
		}
	}
	__syncwarp ();

	// energy dispersal (dab-concurrent.cpp:183-190) + unpack to one bit per byte (viterbi.cpp:240-241)
	uint8_t *out = j. out + (long long) blk * j. frameBits;
	if ((j. frameBits & 3) == 0) {
		for (int i = lane * 4; i < j. frameBits; i += 128) {
			unsigned w = bitsw [i >> 5];
			if (j. prbs) w ^= __ldg (&j. prbs [i >> 5]);
			w >>= (i & 31);
			uchar4 v = make_uchar4 (w & 1, (w >> 1) & 1, (w >> 2) & 1, (w >> 3) & 1);
			*reinterpret_cast<uchar4 *> (out + i) = v;
		}
	} else {
		for (int i = lane; i < j. frameBits; i += 32) {
			unsigned w = bitsw [i >> 5];
			if (j. prbs) w ^= __ldg (&j. prbs [i >> 5]);
			out [i] = (w >> (i & 31)) & 1;
		}
	}
}

cudaError_t vit_launch (dabgpu *h, int cls, const VitJob &job) {
	if (job. nblocks <= 0) return cudaSuccess;
	cudaStream_t st = h -> vst ();
	ProfScope prof (h, cls, st);
	const int words_per_warp = (2 * job. nsteps + (job. frameBits + 31) / 32 + 2) & ~1;   // keeps uint2 alignment
	const size_t per_warp = (size_t) words_per_warp * 4;
	int wpc = (int) ((110 * 1024) / per_warp);
	wpc = wpc < 1 ? 1 : wpc > 4 ? 4 : wpc;
	const size_t smem = per_warp * wpc;
	if (smem > 227 * 1024) return cudaErrorInvalidValue;
	const int grid = (job. nblocks + wpc - 1) / wpc;
	cudaError_t e;
	if (job. deint) {
		e = cudaFuncSetAttribute (vit_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		vit_warp_kernel<true><<<grid, 32 * wpc, smem, st>>> (job, words_per_warp);
	} else {
		e = cudaFuncSetAttribute (vit_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if (e != cudaSuccess) return e;
		vit_warp_kernel<false><<<grid, 32 * wpc, smem, st>>> (job, words_per_warp);
	}
	h -> launches ++;
	return cudaGetLastError ();
}

// FIB CRC (dab-constants.h:310-340): CRC-16-CCITT, all-ones start, last 16 bits complemented, over 256
// one-bit-per-byte values.  Eight lanes per FIB: lane j runs the shift register over its own 32 bits (lane 0 starts from
// all ones, the others from zero) and the result is advanced over the 32 (7 - j) bits that follow by a precomputed linear
// map (the register is linear over GF(2): crc (A | B) = advance (crc (A), |B|) ^ crc_0 (B)); an XOR across the eight
// lanes gives the reference's register.
__host__ __device__ constexpr unsigned crc_advance_bits (unsigned reg, int n) {
	for (int i = 0; i < n; i ++) { const unsigned top = (reg >> 15) & 1u; reg = (reg << 1) & 0xffffu; if (top) reg ^= 0x1021u; }
	return reg;
}
struct CrcAdvance { unsigned short m [8][16]; };
constexpr CrcAdvance make_crc_advance () {
	CrcAdvance t {};
	for (int k = 0; k < 8; k ++) for (int b = 0; b < 16; b ++) t. m [k][b] = (unsigned short) crc_advance_bits (1u << b, 32 * k);
	return t;
}
__constant__ CrcAdvance c_crc_adv = make_crc_advance ();

__global__ void __launch_bounds__ (128) fib_crc_kernel (const uint8_t *bits, int nfibs, uint8_t *ok) {
	const int g = blockIdx. x * blockDim. x + threadIdx. x, f = g >> 3, j = g & 7;
	unsigned reg = 0;
	if (f < nfibs) {
		const uint4 *p = reinterpret_cast<const uint4 *> (bits + (size_t) f * 256 + 32 * j);
		const uint4 v0 = __ldg (p), v1 = __ldg (p + 1);
		const uint32_t w [8] = { v0. x, v0. y, v0. z, v0. w, v1. x, v1. y, v1. z, v1. w };
		reg = j == 0 ? 0xffffu : 0u;
#pragma unroll
		for (int i = 0; i < 32; i ++) {
			unsigned in = (w [i >> 2] >> (8 * (i & 3))) & 1u;
			if (j == 7 && i >= 16) in ^= 1u;                               // bits 240..255
			const unsigned top = (reg >> 15) & 1u;
			reg = (reg << 1) & 0xffffu;
			if (top ^ in) reg ^= 0x1021u;
		}
		unsigned adv = 0;
#pragma unroll
		for (int b = 0; b < 16; b ++) if ((reg >> b) & 1u) adv ^= c_crc_adv. m [7 - j][b];
		reg = adv;
	}
	reg ^= __shfl_xor_sync (0xffffffffu, reg, 1);
	reg ^= __shfl_xor_sync (0xffffffffu, reg, 2);
	reg ^= __shfl_xor_sync (0xffffffffu, reg, 4);
	if (f < nfibs && j == 0) ok [f] = reg == 0;
}

cudaError_t fib_crc_launch (dabgpu *h, const uint8_t *bits, int nfibs, uint8_t *ok) {
	if (nfibs <= 0) return cudaSuccess;
	ProfScope prof (h, KC_CRC, h -> vst ());
	fib_crc_kernel<<<(8 * nfibs + 127) / 128, 128, 0, h -> vst ()>>> (bits, nfibs, ok);
	h -> launches ++;
	return cudaGetLastError ();
}
