// dabgpu_symbol.cu -- the FFT + demod group of the stream engine: per frame the PRS correlator and PRS block
// (phaseReference::findIndex, ofdmDecoder::processBlock_0; ofdm-processor.cpp:344-406) and the data symbols
// (ofdmDecoder::processToken with the cyclic-prefix correlation fused; ofdm-processor.cpp:414-442, ofdm-decoder.cpp:167-207).
//
// Two families: the generic kernels (in-shared-memory Stockham FFT, any mode, any sample format) and the packed
// register-FFT kernels (dabgpu_fftp.cuh) for Modes I, II and IV.  Soft bits leave as the Viterbi's byte symbols
// (0..255 = soft + 127, viterbi.cpp:229-235) only: the value range [-127, 127] makes that lossless, and the int16 form
// the reference hands to process_ficBlock / process_mscBlock is produced on demand (soft16_kernel).
#include <math.h>
#include "dabgpu_engine.h"
#include "dabgpu_fftp.cuh"

__device__ __forceinline__ uchar2 win_fetch (const SampleWin &w, long long i) {      // u8 windows only
	return i < w. len0 ? __ldg (&w. seg0 [i]) : __ldg (&w. seg1 [i - w. len0]);
}
// sample i as the complex float the reference's getSample sees before the NCO (ofdm-processor.cpp:133-183)
__device__ __forceinline__ float2 win_sample (const SampleWin &w, long long i) {
	if (w. cf32 == 1)
		return i < w. len0 ? __ldg (reinterpret_cast<const float2 *> (w. seg0) + i) : __ldg (reinterpret_cast<const float2 *> (w. seg1) + (i - w. len0));
	if (w. cf32 == 2) {                                      // 16-bit PCM as sf_readf_float delivers it (wavfiles.cpp:190): x / 32768, exact in float
		const short2 v = i < w. len0 ? __ldg (reinterpret_cast<const short2 *> (w. seg0) + i) : __ldg (reinterpret_cast<const short2 *> (w. seg1) + (i - w. len0));
		return make_float2 ((float) v. x * (1.0f / 32768.0f), (float) v. y * (1.0f / 32768.0f));
	}
	const uchar2 s = win_fetch (w, i);
	return make_float2 ((float) ((int) s. x - 128) * (1.0f / 128.0f), (float) ((int) s. y - 128) * (1.0f / 128.0f));
}

// dst[i] = sample (first + i) after conversion and NCO, i < n (rawfiles.cpp:113-116; ofdm-processor.cpp:217-226)
__device__ __forceinline__ void load_win_nco (float2 *dst, const SampleWin &w, long long first, int n,
                                              int lp_before, int phase, const OfdmTables &T) {
	const int tid = threadIdx. x;
	const int ph = mod_rate (phase);
	int lp = mod_rate ((long long) lp_before - (long long) (tid + 1) * ph);
	const int step = mod_rate ((long long) OFDM_THREADS * ph);
	for (int i = tid; i < n; i += OFDM_THREADS) {
		dst [i] = cmul (win_sample (w, first + i), nco (T, lp));
		lp -= step;
		if (lp < 0) lp += DAB_INPUT_RATE;
	}
}

// where the soft bits of symbol l (1 .. L-1) of output frame slot `slot` go: FIC symbols 1..3, MSC symbols 4.. as CIF rows
// (fic-handler.cpp:143-153; msc-handler.cpp:125-193: row = CIF, 15 history rows in front)
// A CIF is blocksPerCIF symbols of 2K soft bits = exactly CIF_BITS in every mode (18 x 3072, 72 x 768, 36 x 1536), so the MSC
// symbols of a frame are contiguous: symbol l >= 4 sits (l - 4) * 2K behind the frame's first CIF row.
struct SymGeom { int K2, blocksPerCIF, cifsPerFrame; };
__device__ __forceinline__ uint8_t *sym_out (const StreamDev &S, const SymGeom &g, int slot, int l) {
	if (l < 4) return S. fic8 + ((size_t) slot * 3 + (l - 1)) * g. K2;
	return S. msc8 + ((size_t) 15 + (size_t) slot * g. cifsPerFrame) * CIF_BITS + (size_t) (l - 4) * g. K2;
}

// ---------------------------------------------------------------------------------------------------
// generic front kernel, one CTA per chunk slot: SyncOnPhase + OFDM_PRS (ofdm-processor.cpp:344-406)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (OFDM_THREADS) front_kernel (const StreamDev *sd, OfdmTables T, const FrameIn *fin, FrameOut *fo, float2 *spec0) {
	extern __shared__ float2 sm [];
	__shared__ float cv [96];
	const int N = T. T_u, c = blockIdx. x;
	const FrameIn in = fin [c];
	if (!in. active) return;
	const SampleWin w = sd [in. stream]. w;
	float2 *a = sm, *b = sm + N;
	load_win_nco (a, w, in. P, N, in. lp, in. phiA, T);                // :347-348
	const int s = find_index_block (a, b, T);                          // :352
	int corr = 0;
	if (s >= 0) {
		// block 0 = the T_u samples from P + s on (:362-388), same NCO run
		const int lp0 = mod_rate ((long long) in. lp - (long long) s * mod_rate (in. phiA));
		__syncthreads ();
		load_win_nco (a, w, in. P + s, N, lp0, in. phiA, T);
		float2 *f = block_fft (a, b, N, T. tw);
		float2 *g = spec0 + (size_t) c * N;
		for (int i = threadIdx. x; i < N; i += OFDM_THREADS) g [i] = f [i];   // phaseReference (ofdm-decoder.cpp:91)
		corr = coarse_offset_warp0 (f, T, cv);                         // always computed; the scan applies the flag
	}
	if (threadIdx. x == 0) { fo [c]. startIndex = s; fo [c]. correction = corr; }
}

// ---------------------------------------------------------------------------------------------------
// generic symbol kernel, CTA (c, g) = chunk slot c, symbol group g: OFDM_SYMBOLS (ofdm-processor.cpp:414-442)
// with processToken (ofdm-decoder.cpp:167-190) and the cyclic-prefix correlation (:424-425) fused.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void demod_symbol8 (const float2 *cur, float2 *prev, const OfdmTables &T, uint8_t *__restrict__ sym8) {
	for (int i = threadIdx. x; i < T. K; i += OFDM_THREADS) {
		const int idx = __ldg (&T. perm [i]);
		const float2 c = cur [idx];
		const float2 r1 = cmulc (c, prev [idx]);
		prev [idx] = c;
		const float ab1 = fabsf (r1. x) + fabsf (r1. y);                 // jan_abs, dab-constants.h:127-134
		sym8 [i] = (uint8_t) (quant127 (r1. x, ab1) + 127);
		sym8 [T. K + i] = (uint8_t) (quant127 (r1. y, ab1) + 127);
	}
}

__global__ void __launch_bounds__ (OFDM_THREADS) symbol_kernel (const StreamDev *sd, OfdmTables T, const FrameIn *fin, int groups, SymGeom geo,
                                                                const FrameOut *fo, const float2 *spec0, float2 *fcpart) {
	extern __shared__ float2 sm [];
	__shared__ float2 s_fc [OFDM_THREADS / 32];
	const int N = T. T_u, Ts = T. T_s, Tg = T. T_g, c = blockIdx. x / groups, g = blockIdx. x % groups;
	const FrameIn in = fin [c];
	if (!in. active) return;
	const int s = fo [c]. startIndex;
	if (s < 0) { if (threadIdx. x == 0) fcpart [c * MAX_GROUPS + g] = make_float2 (0.f, 0.f); return; }
	const StreamDev &S = sd [in. stream];
	const SampleWin w = S. w;
	float2 *symbuf = sm, *scratch = sm + Ts, *prev = sm + Ts + N;
	const int nsym = T. L - 1, per = (nsym + groups - 1) / groups;
	const int l0 = 1 + g * per, l1 = min (nsym + 1, l0 + per);             // symbols [l0, l1)
	const int phA = mod_rate (in. phiA), phiB = in. phiB, phB = mod_rate (phiB);
	const long long F = in. P + s;                                         // first sample of the PRS
	const int lpD = mod_rate ((long long) in. lp - (long long) (s + N) * phA);  // localPhase after the PRS
	// symbol l (>= 1) occupies samples [F + N + (l-1) Ts, + Ts): guard first, then the useful part
	if (l0 == 1) {
		const float2 *p0 = spec0 + (size_t) c * N;
		for (int i = threadIdx. x; i < N; i += OFDM_THREADS) prev [i] = p0 [i];
	} else {
		const long long first = F + N + (long long) (l0 - 2) * Ts + Tg;
		const int lpb = mod_rate ((long long) lpD - ((long long) (l0 - 2) * Ts + Tg) % DAB_INPUT_RATE * phB);
		load_win_nco (symbuf, w, first, N, lpb, phiB, T);
		float2 *f = block_fft (symbuf, scratch, N, T. tw);
		for (int i = threadIdx. x; i < N; i += OFDM_THREADS) prev [i] = f [i];
	}
	float2 acc = make_float2 (0.f, 0.f);
	for (int l = l0; l < l1; l ++) {
		const long long first = F + N + (long long) (l - 1) * Ts;
		const int lpb = mod_rate ((long long) lpD - ((long long) (l - 1) * Ts) % DAB_INPUT_RATE * phB);
		__syncthreads ();
		load_win_nco (symbuf, w, first, Ts, lpb, phiB, T);
		__syncthreads ();
		for (int i = N + threadIdx. x; i < Ts; i += OFDM_THREADS) {        // FreqCorr += x[i] * conj (x[i - T_u])
			const float2 r = cmulc (symbuf [i], symbuf [i - N]);
			acc. x += r. x; acc. y += r. y;
		}
		float2 *f = block_fft (symbuf + Tg, scratch, N, T. tw);
		demod_symbol8 (f, prev, T, sym_out (S, geo, in. slot, l));
	}
	for (int o = 16; o > 0; o >>= 1) {
		acc. x += __shfl_xor_sync (0xffffffffu, acc. x, o);
		acc. y += __shfl_xor_sync (0xffffffffu, acc. y, o);
	}
	if ((threadIdx. x & 31) == 0) s_fc [threadIdx. x >> 5] = acc;
	__syncthreads ();
	if (threadIdx. x == 0) {
		float2 t = make_float2 (0.f, 0.f);
		for (int k = 0; k < OFDM_THREADS / 32; k ++) { t. x += s_fc [k]. x; t. y += s_fc [k]. y; }
		fcpart [c * MAX_GROUPS + g] = t;
	}
}

// ---------------------------------------------------------------------------------------------------
// Packed register-FFT symbol kernel, Modes I / IV / II = NSYM 1 / 2 / 4 symbols per pass, FMT = sample format.
// A pass covers NSYM consecutive symbols = 2552 samples in every mode (T_s = 2552 / NSYM):
//   * the raw samples of pass p+1 stream into shared memory (one TMA bulk copy + mbarrier, double buffered) while pass p
//     is transformed, so no thread ever waits on HBM;
//   * thread t = (symbol j = t / G, u = t mod G), G = 256 / NSYM, takes the useful samples u + G k (k = 0..7) of its
//     symbol from the raw buffer, converts them (u8: PRMT + one packed add, the 2^23 trick; int16: x / 32768) and mixes
//     them with the NCO by a per-thread phasor recurrence -- the integer localPhase of the reference is reproduced
//     exactly at the start of every pass, the 1/128 of rawfiles.cpp:113-116 is folded into the phasor;
//   * p_fft<NSYM> leaves the spectra in shared memory in the FFT's own digit-reversed order (the carrier table is
//     pre-permuted on the host); two spectrum buffers alternate, so the previous pass's last symbol is still there
//     as phase reference of this pass's first one (NSYM = 1 keeps the reference values in registers instead: the very
//     same thread read them one symbol earlier);
//   * a thread demodulates carrier PAIRS of its symbol: two byte symbols per 16-bit store.
// ---------------------------------------------------------------------------------------------------
#define P_TW2 512
#define P_TW3 64
template <int FMT> struct RawFmt;
template <> struct RawFmt<0> { static constexpr int B = 2; };      // bytes per complex sample
template <> struct RawFmt<1> { static constexpr int B = 8; };
template <> struct RawFmt<2> { static constexpr int B = 4; };
#define P_RAW(FMT) (2552 * RawFmt<FMT>::B + 48)                   // bytes per raw buffer: a pass + alignment slack, multiple of 16
#ifndef P_NRAW
#define P_NRAW 2                                                 // raw-sample buffers: the copy of pass p + P_NRAW - 1 is issued at the top of pass p
#endif
#define P_DYN_SMEM(FMT) ((2 * R8_SMEM + P_TW2 + P_TW3) * (int) sizeof (float2) + P_NRAW * P_RAW (FMT))

// (b - 128) as float, exactly, b = byte `which` of s: 0x4B0000bb is 2^23 + b
__device__ __forceinline__ float u8_bits (uint32_t s, int which) {
	return __uint_as_float (__byte_perm (s, 0x4B000000u, which ? 0x7441 : 0x7440));
}
template <int FMT> __device__ __forceinline__ c32 raw_sample (const unsigned char *raw, int i) {
	if (FMT == 0) {                                                         // (I - 128, Q - 128); the 1 / 128 sits in the phasor
		const uint32_t v = reinterpret_cast<const unsigned short *> (raw) [i];
		return __fadd2_rn (make_float2 (u8_bits (v, 0), u8_bits (v, 1)), make_float2 (-8388736.0f, -8388736.0f));
	}
	if (FMT == 2) {                                                         // 16-bit PCM: x / 32768 (wavfiles.cpp:186-197), exact
		const short2 v = reinterpret_cast<const short2 *> (raw) [i];
		return __fmul2_rn (make_float2 ((float) v. x, (float) v. y), make_float2 (1.0f / 32768.0f, 1.0f / 32768.0f));
	}
	return reinterpret_cast<const float2 *> (raw) [i];
}
// soft-bit quantisation of ofdm-decoder.cpp:183-189 for both components of r1 at once: q = -r1 / (|re| + |im|) through the
// reciprocal unit (2 ulp; the soft bits' stated tolerance is +-1 step and comes from the FFT, whose rounding differs from
// the reference's FFTW anyway), times 127 rounded toward zero (the reference's double product is exact; truncating it
// equals truncating the round-toward-zero float product, because that rounding never crosses an integer)
__device__ __forceinline__ void quant_pair (c32 r1, int &re, int &im) {
	const float ab1 = fabsf (r1. x) + fabsf (r1. y);
	float rc;
	asm ("rcp.approx.ftz.f32 %0, %1;" : "=f" (rc) : "f" (ab1));
	const c32 q = __fmul2_rz (__fmul2_rn (r1, make_float2 (- rc, - rc)), make_float2 (127.0f, 127.0f));
	re = __float2int_rz (q. x); im = __float2int_rz (q. y);               // NaN (ab1 == 0) -> 0
}

struct SymPArgs {
	const StreamDev *sd; OfdmTables T; const FrameIn *fin; const FrameOut *fo; const float2 *spec0; float2 *fcpart;
	int groups, per; SymGeom geo;
};

#ifndef P_MINB
#define P_MINB 3
#endif
#ifndef P_TOPSYNC
#define P_TOPSYNC 0
#endif
#ifndef P_PHASOR_LATE
#define P_PHASOR_LATE 0
#endif
#ifndef P_HOIST
#define P_HOIST 1
#endif
template <int NSYM, int FMT>
__global__ void __launch_bounds__ (256, FMT == 1 ? 2 : P_MINB) symbol_kernel_p (const SymPArgs a) {
	extern __shared__ __align__ (1024) unsigned char p_dyn [];          // the FFT buffers must be 512-byte aligned (p_fft)
	float2 *buf0 = reinterpret_cast<float2 *> (p_dyn), *buf1 = buf0 + R8_SMEM, *tw2 = buf1 + R8_SMEM, *tw3 = tw2 + P_TW2;
	unsigned char *raw = reinterpret_cast<unsigned char *> (tw3 + P_TW3);
	__shared__ float2 s_fc [8];
	__shared__ int s_fail;
	__shared__ int s_off [P_NRAW];
	__shared__ __align__ (8) unsigned long long s_mbar [P_NRAW];
	constexpr int G = 256 / NSYM, N = 2048 / NSYM, Ts = 2552 / NSYM, Tg = 504 / NSYM, K = 1536 / NSYM, B = RawFmt<FMT>::B;
	constexpr int RAW = P_RAW (FMT), GSH = 8 / NSYM;                    // N - Tg = 6 G + GSH
	const OfdmTables &T = a. T;
	const int t = threadIdx. x, j = t / G, u = t % G;
	const int c = blockIdx. x / a. groups, g = blockIdx. x % a. groups;
	const FrameIn in = a. fin [c];
	if (!in. active) return;
	const int s = a. fo [c]. startIndex;
	if (s < 0) { if (t == 0) a. fcpart [c * MAX_GROUPS + g] = make_float2 (0.f, 0.f); return; }
	const StreamDev &S = a. sd [in. stream];
	const SampleWin w = S. w;
	const int nsym = T. L - 1;
	const int l0 = 1 + g * a. per, l1 = min (nsym + 1, l0 + a. per);      // symbols [l0, l1); a.per is a multiple of NSYM
	if (l0 >= l1) { if (t == 0) a. fcpart [c * MAX_GROUPS + g] = make_float2 (0.f, 0.f); return; }
	const int phA = mod_rate (in. phiA), phB = mod_rate (in. phiB);
	const long long F = in. P + s;                                         // first sample of the PRS
	const int lpD = mod_rate ((long long) in. lp - (long long) (s + N) * phA);  // localPhase after the PRS
	const c32 rotG = nco (T, mod_rate (- (long long) G * phB));            // G samples further: phase index - G f
	// pass-relative index of the thread's first useful sample, and of the guard sample its useful sample 6 G + u pairs with
	const int iu = j * Ts + Tg + u, ig = j * Ts + u - GSH;                 // (ig < j Ts: useful sample 6 G + u has no guard partner)
	const int offU = mod_rate ((long long) (iu + 1) * phB), offG = mod_rate ((long long) (ig + 1) * phB);
	const int dPass = mod_rate ((long long) NSYM * Ts * phB);
	const c32 rotGn = make_float2 (- rotG. y, rotG. y);                    // (the rotation is applied nine times per pass: its sign-prepared pair is kept)
	c32 x [8];
	float2 tw1 [6];
	const uint32_t mbar0 = (uint32_t) __cvta_generic_to_shared (&s_mbar [0]);
	if (t == 0) {
		s_fail = ((uint32_t) __cvta_generic_to_shared (buf0) & 511u) != 0 ? 1 : 0;   // layout contract of p_fft
		for (int k = 0; k < P_NRAW; k ++) asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r" (mbar0 + 8 * k));
		asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	p_fill_tables<NSYM> (tw2, tw3, T. tw);
	p_load_tw1<NSYM> (tw1, T. tw);
	__syncthreads ();
	if (s_fail) { if (t == 0) atomicOr (const_cast<int *> (&S. ctl. fault), 1); return; }

	// raw samples of the pass starting with symbol l (nv symbols of it are wanted) -> raw buffer b.  Thread 0 alone works out
	// where they are and issues ONE bulk copy by the TMA engine from the 16-byte-aligned address below the first sample; it
	// leaves the byte offset of the pass's first sample inside the buffer in s_off [b].  A pass that straddles the seam between
	// the kept tail and the new input, or touches a buffer end, cannot be copied that way: s_off [b] = -1, and all threads copy
	// it by hand when the pass is consumed (rare: at most one pass per seam).
	auto stage = [&] (int l, int b) {
		if (t != 0) return;
		const int ns = min (NSYM, l1 - l) * Ts;
		const long long first = F + N + (long long) (l - 1) * Ts;
		const unsigned char *seg = nullptr; long long rel = 0, seglen = 0;
		if (first + ns <= w. len0) { seg = reinterpret_cast<const unsigned char *> (w. seg0); rel = first; seglen = w. len0; }
		else if (first >= w. len0) { seg = reinterpret_cast<const unsigned char *> (w. seg1); rel = first - w. len0; seglen = w. len1; }
		int off = -1;
		const uint32_t mb = mbar0 + 8u * (uint32_t) b;
		if (seg != nullptr) {
			const unsigned long long p = (unsigned long long) (seg + rel * B), pa = p & ~15ull;
			const int n16 = ((int) (p - pa) + B * ns + 15) >> 4;
			if (pa >= (unsigned long long) seg && pa + 16ull * n16 <= (unsigned long long) (seg + seglen * B)) {
				off = (int) (p - pa);
				const uint32_t bytes = 16u * (uint32_t) n16;
				asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's last readers (generic proxy) are past a CTA barrier
				asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r" (mb), "r" (bytes) : "memory");
				asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				              :: "r" ((uint32_t) __cvta_generic_to_shared (raw + b * RAW)), "l" (pa), "r" (bytes), "r" (mb) : "memory");
			}
		}
		if (off < 0) asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r" (mb) : "memory");   // nothing in flight: the phase completes at once
		s_off [b] = off;
	};
	// the by-hand copy of a pass (all threads), then a CTA barrier
	auto stage_by_hand = [&] (int l, int b) {
		const int ns = min (NSYM, l1 - l) * Ts;
		const long long first = F + N + (long long) (l - 1) * Ts;
		unsigned char *dst = raw + b * RAW;
		if (FMT == 0) for (int i = t; i < ns; i += 256) reinterpret_cast<uchar2 *> (dst) [i] = win_fetch (w, first + i);
		else if (FMT == 2) {
			for (int i = t; i < ns; i += 256) {
				const long long q = first + i;
				reinterpret_cast<short2 *> (dst) [i] = q < w. len0 ? __ldg (reinterpret_cast<const short2 *> (w. seg0) + q) : __ldg (reinterpret_cast<const short2 *> (w. seg1) + (q - w. len0));
			}
		} else for (int i = t; i < ns; i += 256) reinterpret_cast<float2 *> (dst) [i] = win_sample (w, first + i);
		__syncthreads ();
	};
	// wait for raw buffer b (use = how often it has been waited for before).  A copy that never arrives must not hang the
	// GPU: after 2 s (global timer) the CTA gives up together and reports through StreamCtl::fault.
	auto wait_raw = [&] (int b, int use) {
		const uint32_t mb = mbar0 + 8u * (uint32_t) b, parity = (uint32_t) use & 1u;
		uint32_t ok = 0;
		unsigned long long t0 = 0;
		for (int spin = 0; !ok; spin ++) {
			asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r" (ok) : "r" (mb), "r" (parity) : "memory");
			if (!ok && (spin & 1023) == 1023) {
				unsigned long long now;
				asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (now));
				if (t0 == 0) t0 = now; else if (now - t0 > 2000000000ull) *(volatile int *) &s_fail = 1;
				if (*(volatile int *) &s_fail) break;
			}
		}
	};

	for (int k = 0; k < P_NRAW - 1; k ++) if (l0 + k * NSYM < l1) stage (l0 + k * NSYM, k);
	// ---- phase reference of the first symbol: spectrum of symbol l0 - 1 -> block NSYM - 1 of buf1 ----
	if (l0 == 1) {
		const float2 *p0 = a. spec0 + (size_t) c * N;
		for (int k = t; k < N; k += 256) buf1 [r8_swz (p_pos<NSYM> (NSYM - 1, k))] = p0 [k];
		__syncthreads ();
	} else {
		// the thread group of symbol slot NSYM - 1 loads symbol l0 - 1 (straight from the window); the other groups transform zeros
		const long long first = F + N + (long long) (l0 - 2) * Ts + Tg;
		const int lpb = mod_rate ((long long) lpD - ((long long) (l0 - 2) * Ts + Tg) % DAB_INPUT_RATE * phB);
		c32 ph = nco (T, mod_rate ((long long) lpb - (long long) (u + 1) * phB));
#pragma unroll
		for (int k = 0; k < 8; k ++) {
			x [k] = j == NSYM - 1 ? p_cmul (win_sample (w, first + u + G * k), ph) : make_float2 (0.f, 0.f);
			ph = p_cmul (ph, rotG);
		}
		p_fft<NSYM> (x, buf1, tw1, tw2, tw3);
	}
	c32 acc = make_float2 (0.f, 0.f);
	// this thread's six carriers of its symbol: the pairs 2 u + 2 G m (+ 1), m = 0..2; positions relative to the symbol's block
	uint32_t pidx [3];
#pragma unroll
	for (int m = 0; m < 3; m ++) pidx [m] = (uint32_t) __ldg (&T. permpos [2 * u + 2 * G * m]) | ((uint32_t) __ldg (&T. permpos [2 * u + 1 + 2 * G * m]) << 16);
	c32 pv [6];                                                            // NSYM == 1: the phase reference lives in registers
	if (NSYM == 1) {
#pragma unroll
		for (int m = 0; m < 3; m ++) { pv [2 * m] = buf1 [pidx [m] & 0xffffu]; pv [2 * m + 1] = buf1 [pidx [m] >> 16]; }
	}
	int lpb = mod_rate ((long long) lpD - ((long long) (l0 - 1) * Ts) % DAB_INPUT_RATE * phB);   // localPhase before the pass's first sample
	const float sc = FMT == 0 ? 1.0f / 128.0f : 1.0f;
	auto phasors = [&] (int lp, c32 &pg, c32 &pu) {                        // the NCO phasors of guard sample ig and useful sample iu
		int xg = lp - offG; if (xg < 0) xg += DAB_INPUT_RATE;
		int xu = lp - offU; if (xu < 0) xu += DAB_INPUT_RATE;
		pg = p_scale (nco (T, xg), sc); pu = p_scale (nco (T, xu), sc);
	};
	auto phasor_prefetch = [&] (int lp) {
		int xg = lp - offG; if (xg < 0) xg += DAB_INPUT_RATE;
		int xu = lp - offU; if (xu < 0) xu += DAB_INPUT_RATE;
		asm volatile ("prefetch.global.L1 [%0];" :: "l" (&T. osc_hi [xg >> 11]));
		asm volatile ("prefetch.global.L1 [%0];" :: "l" (&T. osc_lo [xg & 2047]));
		asm volatile ("prefetch.global.L1 [%0];" :: "l" (&T. osc_hi [xu >> 11]));
		asm volatile ("prefetch.global.L1 [%0];" :: "l" (&T. osc_lo [xu & 2047]));
	};
	c32 phg_n, ph_n;
	phasors (lpb, phg_n, ph_n);
	const c32 rotG6 = NSYM == 1 ? rotG : nco (T, mod_rate (- 6ll * G * phB));   // guard partner of useful sample 7 G + u is G further on than that of 6 G + u
	// output rows of the frame: FIC symbols 1..3, then the MSC symbols contiguously (sym_out), fetched from the stream table once
	uint8_t *const fic_base = S. fic8 + (size_t) in. slot * 3 * a. geo. K2 - a. geo. K2;
	uint8_t *const msc_base = S. msc8 + ((size_t) 15 + (size_t) in. slot * a. geo. cifsPerFrame) * CIF_BITS - (size_t) 4 * a. geo. K2;
	int pass = 0;
	for (int l = l0; l < l1; l += NSYM, pass ++) {
		const int b = pass & 1, rb = pass % P_NRAW;                        // spectrum buffer, raw buffer of this pass
		if (l + (P_NRAW - 1) * NSYM < l1) stage (l + (P_NRAW - 1) * NSYM, (pass + P_NRAW - 1) % P_NRAW);
		c32 phg = phg_n, ph = ph_n;
		lpb -= dPass; if (lpb < 0) lpb += DAB_INPUT_RATE;
#if P_PHASOR_LATE
		phasor_prefetch (lpb);                                             // the next pass's NCO table entries: into L1 now, looked up after the transform
#else
		phasors (lpb, phg_n, ph_n);                                        // looked up one pass ahead
#endif
		// raw buffer b complete (every thread waits on the copy's mbarrier itself).  NSYM = 1 needs no CTA barrier here: the spectrum
		// buffers alternate and the phase reference lives in registers, so this pass's transform writes the buffer the demodulation
		// of the pass BEFORE the last one read -- and every thread has been through the barriers inside the last pass's transform
		// since then; s_off [b] was written before those too.  NSYM > 1: the last pass's demodulation read BOTH buffers.
		wait_raw (rb, pass / P_NRAW);
		if (NSYM > 1 || P_TOPSYNC) __syncthreads ();
		int off_cur = s_off [rb];
		if (off_cur < 0) { stage_by_hand (l, rb); off_cur = 0; }
		const unsigned char *rs = raw + rb * RAW + off_cur;
		const bool live = l + j < l1;                                      // a partial last pass: the trailing thread groups idle
		// guard samples ig and ig + G (the ones x[6] and x[7] are correlated with), mixed like every other sample
		c32 g6 = make_float2 (0.f, 0.f), g7;
		g7 = p_cmul (raw_sample<FMT> (rs, ig + G), p_cmul_pre (phg, rotG. x, rotGn));
		if (u >= GSH) g6 = p_cmul (raw_sample<FMT> (rs, ig), phg);
#pragma unroll
		for (int k = 0; k < 8; k ++) {
			x [k] = p_cmul (raw_sample<FMT> (rs, iu + G * k), ph);
			ph = p_cmul_pre (ph, rotG. x, rotGn);
		}
		// FreqCorr += x[i] * conj (x[i - T_u]), i in [T_u, T_s): useful sample e pairs with guard sample e - (T_u - T_g)
		if (live) { acc = p_add (acc, p_cmulc (x [7], g7)); if (u >= GSH) acc = p_add (acc, p_cmulc (x [6], g6)); }
		float2 *cur = b ? buf1 : buf0, *oth = b ? buf0 : buf1;             // (NSYM = 1: the reference itself lives in registers, buf1 only takes turns)
		p_fft<NSYM> (x, cur, tw1, tw2, tw3);
		// (checked behind the transform's barriers, where every thread sees the same value: a thread whose wait gave up has gone
		// through the transform with whatever was in the buffer -- nothing of it leaves the CTA)
		if (*(volatile int *) &s_fail) { if (t == 0) atomicOr (const_cast<int *> (&S. ctl. fault), 2); return; }
#if P_PHASOR_LATE
		phasors (lpb, phg_n, ph_n);                                        // (L1 hits by now)
#endif
		if (live) {
#if P_HOIST
			uint8_t *out8 = (l + j < 4 ? fic_base : msc_base) + (size_t) (l + j) * a. geo. K2;
#else
			uint8_t *out8 = sym_out (S, a. geo, in. slot, l + j);
#endif
			const float2 *cb = cur + j * N, *pb = j > 0 ? cur + (j - 1) * N : oth + (NSYM - 1) * N;
			// all six carriers (and their references) are fetched before the first result is stored: the stores go through a
			// generic pointer the compiler cannot tell apart from the spectrum buffers, so it would not move a later load above one
			c32 cc [6], pr [6];
#pragma unroll
			for (int m = 0; m < 3; m ++)
#pragma unroll
				for (int q = 0; q < 2; q ++) {
					const int idx = (int) (q ? pidx [m] >> 16 : pidx [m] & 0xffffu);
					cc [2 * m + q] = cb [idx];
					if (NSYM == 1) { pr [2 * m + q] = pv [2 * m + q]; pv [2 * m + q] = cc [2 * m + q]; } else pr [2 * m + q] = pb [idx];
				}
#pragma unroll
			for (int m = 0; m < 3; m ++) {
				const int i = 2 * u + 2 * G * m;
				int re [2], im [2];
#pragma unroll
				for (int q = 0; q < 2; q ++) quant_pair (p_cmulc (cc [2 * m + q], pr [2 * m + q]), re [q], im [q]);
				*reinterpret_cast<unsigned short *> (out8 + i)     = (unsigned short) (re [0] + 256 * re [1] + 0x7f7f);
				*reinterpret_cast<unsigned short *> (out8 + K + i) = (unsigned short) (im [0] + 256 * im [1] + 0x7f7f);
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		acc. x += __shfl_xor_sync (0xffffffffu, acc. x, o);
		acc. y += __shfl_xor_sync (0xffffffffu, acc. y, o);
	}
	if ((t & 31) == 0) s_fc [t >> 5] = acc;
	__syncthreads ();
	if (t == 0) {
		float2 sum = make_float2 (0.f, 0.f);
		for (int k = 0; k < 8; k ++) { sum. x += s_fc [k]. x; sum. y += s_fc [k]. y; }
		a. fcpart [c * MAX_GROUPS + g] = sum;
	}
}

// ---------------------------------------------------------------------------------------------------
// Mode I front kernel on the packed register FFT: same work as front_kernel (SyncOnPhase + OFDM_PRS,
// ofdm-processor.cpp:344-406).  The spectrum of a transform stays in the FFT's digit-reversed shared layout; the
// correlator's product with the PRS reference reads it through the position map, the magnitude scan walks the storage
// order and maps every slot back to its bin (first maximum = smallest bin among equals, as the reference's ascending
// loop finds it).
// ---------------------------------------------------------------------------------------------------
#define FRONT_P_SMEM ((2 * R8_SMEM + P_TW2 + P_TW3) * (int) sizeof (float2))
__host__ __device__ __forceinline__ int p_pos_inv1 (int p) { return (p >> 9) | (((p >> 6) & 7) << 2) | (((p >> 3) & 7) << 5) | ((p & 7) << 8); }

__global__ void __launch_bounds__ (256) front_kernel_p (const StreamDev *sd, OfdmTables T, const FrameIn *fin, FrameOut *fo, float2 *spec0) {
	extern __shared__ __align__ (1024) unsigned char fp_dyn [];
	float2 *A = reinterpret_cast<float2 *> (fp_dyn), *Bf = A + R8_SMEM, *tw2 = Bf + R8_SMEM, *tw3 = tw2 + P_TW2;
	__shared__ float cv [96];
	__shared__ float s_red [8], s_max [8];
	__shared__ int s_idx [8], s_result;
	const int N = R8_N, c = blockIdx. x, t = threadIdx. x;
	const FrameIn in = fin [c];
	if (!in. active) return;
	const SampleWin w = sd [in. stream]. w;
	c32 x [8];
	float2 tw1 [6];
	p_fill_tables<1> (tw2, tw3, T. tw);
	p_load_tw1<1> (tw1, T. tw);
	const int phA = mod_rate (in. phiA), step256 = mod_rate (256ll * phA);
	// x [k] = sample (first + t + 256 k) after conversion and NCO (rawfiles.cpp:113-116; ofdm-processor.cpp:217-226)
	auto load = [&] (long long first, int lp_before) {
		int lp = mod_rate ((long long) lp_before - (long long) (t + 1) * phA);
#pragma unroll
		for (int k = 0; k < 8; k ++) {
			x [k] = p_cmul (win_sample (w, first + t + 256 * k), nco (T, lp));
			lp -= step256; if (lp < 0) lp += DAB_INPUT_RATE;
		}
	};
	load (in. P, in. lp);                                              // :347-348
	__syncthreads ();                                                  // twiddle tables in place
	p_fft<1> (x, A, tw1, tw2, tw3);
	// res = conj (fft * conj (ref)): the backward transform is conj (forward (conj (x))) (phasereference.cpp:66-73)
#pragma unroll
	for (int k = 0; k < 8; k ++) {
		const int kk = t + 256 * k;
		const c32 r = p_cmulc (A [r8_swz (p_pos<1> (0, kk))], __ldg (&T. ref [kk]));
		x [k] = make_float2 (r. x, - r. y);
	}
	__syncthreads ();                                                  // everybody has read A
	p_fft<1> (x, A, tw1, tw2, tw3);
	const float factor = (float) (1.0 / (float) N);                    // fft.cpp:114-121
	float sum = 0.f, mx = -10000.f;
	int mi = -1;
#pragma unroll
	for (int m = 0; m < 8; m ++) {
		const int slot = t + 256 * m, k = p_pos_inv1 (r8_swz (slot));  // the swizzle is an involution
		const float2 v = A [slot];
		const float av = hypotf (v. x * factor, (- v. y) * factor);
		sum += av;
		if (av > mx || (av == mx && k < mi)) { mx = av; mi = k; }
	}
	for (int o = 16; o > 0; o >>= 1) {
		sum += __shfl_xor_sync (0xffffffffu, sum, o);
		const float om = __shfl_xor_sync (0xffffffffu, mx, o);
		const int   oi = __shfl_xor_sync (0xffffffffu, mi, o);
		if (om > mx || (om == mx && oi >= 0 && (mi < 0 || oi < mi))) { mx = om; mi = oi; }
	}
	if ((t & 31) == 0) { s_red [t >> 5] = sum; s_max [t >> 5] = mx; s_idx [t >> 5] = mi; }
	__syncthreads ();
	if (t == 0) {
		float tsum = 0.f, tmx = -10000.f;
		int tmi = -1;
		for (int q = 0; q < 8; q ++) {
			tsum += s_red [q];
			if (s_max [q] > tmx || (s_max [q] == tmx && s_idx [q] >= 0 && (tmi < 0 || s_idx [q] < tmi))) { tmx = s_max [q]; tmi = s_idx [q]; }
		}
		if (tmx < (float) T. level * tsum / (float) N)                 // phasereference.cpp:84-85
			s_result = (int) (- fabsf (tmx / (tsum / (float) N)) - 1.0f);
		else
			s_result = tmi;
	}
	__syncthreads ();
	const int s = s_result;                                            // :352
	int corr = 0;
	if (s >= 0) {
		// block 0 = the T_u samples from P + s on (:362-388), same NCO run
		load (in. P + s, mod_rate ((long long) in. lp - (long long) s * phA));
		p_fft<1> (x, A, tw1, tw2, tw3);
		float2 *g = spec0 + (size_t) c * N;
#pragma unroll
		for (int m = 0; m < 8; m ++) {                                 // phaseReference (ofdm-decoder.cpp:91), natural order
			const int k = t + 256 * m;
			const float2 v = A [r8_swz (p_pos<1> (0, k))];
			g [k] = v; Bf [k] = v;
		}
		__syncthreads ();
		corr = coarse_offset_warp0 (Bf, T, cv);                        // always computed; the scan applies the flag
	}
	if (t == 0) { fo [c]. startIndex = s; fo [c]. correction = corr; }
}

// ---------------------------------------------------------------------------------------------------
// byte symbols -> the int16 soft bits the reference hands to process_ficBlock / process_mscBlock (value = symbol - 127),
// in dabgpu_result.soft's layout [frame][L - 1][2K]; only run when a caller asks for them
// ---------------------------------------------------------------------------------------------------
__global__ void soft16_kernel (const uint8_t *fic8, const uint8_t *msc8, int16_t *out, int nframes, int ficw, int mscw) {
	const long long per = (long long) ficw + mscw, total = (long long) nframes * per;
	for (long long i = (long long) blockIdx. x * blockDim. x + threadIdx. x; i < total; i += (long long) gridDim. x * blockDim. x) {
		const long long f = i / per; const int r = (int) (i % per);
		const int v = r < ficw ? fic8 [f * ficw + r] : msc8 [(size_t) 15 * CIF_BITS + f * mscw + (r - ficw)];
		out [i] = (int16_t) (v - 127);
	}
}
void soft16_launch (dabgpu *h, const uint8_t *fic8, const uint8_t *msc8, int16_t *out, int nframes, cudaStream_t st) {
	const DabParams &p = h -> p;
	soft16_kernel<<<592, 256, 0, st>>> (fic8, msc8, out, nframes, 3 * 2 * p. K, p. cifsPerFrame * CIF_BITS);
	h -> launches ++;
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static int nsym_of_mode (int mode) { return mode == 1 ? 1 : mode == 4 ? 2 : mode == 2 ? 4 : 0; }

bool symbol_has_packed (const dabgpu *h, int fmt) {
	return nsym_of_mode (h -> p. dabMode) != 0 && h -> engine -> T. permpos != nullptr && fmt >= 0 && fmt <= 2;
}

template <int NSYM, int FMT> static cudaError_t sym_attr () {
	return cudaFuncSetAttribute (symbol_kernel_p<NSYM, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_DYN_SMEM (FMT));
}
int symbol_init (dabgpu *h) {
	const int big = 100 * 1024;
	CUDA_TRY (h, cudaFuncSetAttribute (front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
	CUDA_TRY (h, cudaFuncSetAttribute (symbol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
	CUDA_TRY (h, cudaFuncSetAttribute (front_kernel_p, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_P_SMEM));
	CUDA_TRY (h, (sym_attr<1, 0> ())); CUDA_TRY (h, (sym_attr<1, 1> ())); CUDA_TRY (h, (sym_attr<1, 2> ()));
	CUDA_TRY (h, (sym_attr<2, 0> ())); CUDA_TRY (h, (sym_attr<2, 1> ())); CUDA_TRY (h, (sym_attr<2, 2> ()));
	CUDA_TRY (h, (sym_attr<4, 0> ())); CUDA_TRY (h, (sym_attr<4, 1> ())); CUDA_TRY (h, (sym_attr<4, 2> ()));
	return DABGPU_OK;
}

void front_launch (dabgpu *h, const StreamDev *sd, int nslots, const ChunkBufs &cb, int which, cudaStream_t st) {
	ProfScope prof (h, KC_FRONT, st);
	const DabParams &p = h -> p;
	if (which == 0 && p. dabMode == 1 && h -> engine -> T. permpos)
		front_kernel_p<<<nslots, 256, FRONT_P_SMEM, st>>> (sd, h -> engine -> T, cb. fin, cb. fo, cb. spec0);
	else
		front_kernel<<<nslots, OFDM_THREADS, 2 * (size_t) p. T_u * sizeof (float2), st>>> (sd, h -> engine -> T, cb. fin, cb. fo, cb. spec0);
	h -> launches ++;
}

template <int NSYM> static void sym_p_go (int fmt, int grid, const SymPArgs &a, cudaStream_t st) {
	if (fmt == 0) symbol_kernel_p<NSYM, 0><<<grid, 256, P_DYN_SMEM (0), st>>> (a);
	else if (fmt == 1) symbol_kernel_p<NSYM, 1><<<grid, 256, P_DYN_SMEM (1), st>>> (a);
	else symbol_kernel_p<NSYM, 2><<<grid, 256, P_DYN_SMEM (2), st>>> (a);
}

void symbol_launch (dabgpu *h, const StreamDev *sd, int nslots, int groups, const ChunkBufs &cb, int which, int fmt, cudaStream_t st) {
	ProfScope prof (h, KC_SYMBOL, st);
	const DabParams &p = h -> p;
	Engine *E = h -> engine;
	const SymGeom geo { 2 * p. K, p. blocksPerCIF, p. cifsPerFrame };
	const int nsym = nsym_of_mode (p. dabMode);
	if (which == 0 && symbol_has_packed (h, fmt)) {
		int per = (p. L - 1 + groups - 1) / groups;
		per = (per + nsym - 1) / nsym * nsym;
		const SymPArgs a { sd, E -> T, cb. fin, cb. fo, cb. spec0, cb. fcpart, groups, per, geo };
		if (nsym == 1) sym_p_go<1> (fmt, nslots * groups, a, st);
		else if (nsym == 2) sym_p_go<2> (fmt, nslots * groups, a, st);
		else sym_p_go<4> (fmt, nslots * groups, a, st);
	} else
		symbol_kernel<<<nslots * groups, OFDM_THREADS, ((size_t) p. T_s + 2 * p. T_u) * sizeof (float2), st>>> (sd, E -> T, cb. fin, groups, geo, cb. fo, cb. spec0, cb. fcpart);
	h -> launches ++;
}
