// dabgpu_engine.cu -- the stream engine: ofdmProcessor::run (ofdm-processor.cpp:247-474) re-designed for a
// frame-parallel GPU, feeding the FIC and MSC decoders (fic-handler.cpp, msc-handler.cpp, dab-concurrent.cpp).
//
// The reference loop is sequential through three couplings: the frame position (findIndex of frame n fixes
// where frame n+1 is read), the NCO phase/frequency (coarse/fine correctors updated once per frame) and the
// 16-CIF time de-interleaver.  The engine speculates: a pass decodes a chunk of frames in parallel assuming
// the tracking state stays what it is at the chunk start (start index T_g, correctors unchanged -- the steady
// state of a locked receiver), then a single-thread scan kernel replays the reference's scalar state machine
// over the per-frame results (findIndex, coarse correction, cyclic-prefix correlation) and accepts frames up
// to the first one whose assumed inputs differ from the replayed truth; the next pass restarts there with the
// true state.  Accepted frames are therefore computed from exactly the inputs the reference would have used.
// Acquisition (null-symbol search) is a sample-serial scan done once per (re)synchronisation.
#include <math.h>
#include <stdlib.h>
#include "dabgpu_engine.h"

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

// dst[i] = sample (first + i) after u8 conversion and NCO, i < n (rawfiles.cpp:113-116; ofdm-processor.cpp:217-226)
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

// ---------------------------------------------------------------------------------------------------
// acquisition: notSynced -> SyncOnNull -> SyncOnEndNull (ofdm-processor.cpp:275-338), one warp.
// The reference walks the samples one by one through two recurrences -- the signal level IIR
// (sLevel = 0.00001 * jan_abs (v) + (1 - 0.00001) * sLevel, in double, rounded to float, :168) and the running sum of a
// 50-sample envelope window -- and tests a threshold before every sample.  Only the two recurrences are serial.  Per chunk:
//   1. all lanes convert / mix the samples and prepare everything that does not depend on the recurrences: the
//      envelope value e_i (|re| + |im|, or the true magnitude in SyncOnEndNull), the double product 0.00001 * jan_abs,
//      the window difference e_i - e_(i-50);
//   2. lane 0 runs the two recurrences (one double multiply-add-round chain, one float add chain) and records the
//      values BEFORE every sample;
//   3. all lanes evaluate the reference's threshold tests on the recorded values and the first sample that leaves the
//      state is found by a warp reduction; the state is committed up to there.
// Same operations in the same order on every value as the reference's loop, hence the same result bit for bit.
// ---------------------------------------------------------------------------------------------------
#define ACQ_CHUNK 512
__global__ void __launch_bounds__ (32) acquire_kernel (SampleWin w, OfdmTables T, int T_F, int T_null, StreamCtl *ctl) {
	__shared__ float s_e [ACQ_CHUNK], s_d [ACQ_CHUNK], s_sl [ACQ_CHUNK], s_csb [ACQ_CHUNK], s_ring [64];
	__shared__ double s_ax [ACQ_CHUNK];
	const int lane = threadIdx. x;
	const long long total = w. len0 + w. len1;
	// scalar state, identical in every lane (updated from lane 0's results by shuffles)
	int stage = 0, cnt = 0, counter = 0, idx = 0, done = 0;
	float sLevel = 0.f, cs = 0.f;
	long long pos = ctl -> pos, attempt_pos = pos;
	int lp = ctl -> lp, attempt_lp = lp;
	const int phi = ctl -> coarse + ctl -> fine;
	while (true) {
		int n;                                                       // samples until the stage can change by COUNT
		if (stage == 0) n = 20 * T. T_s - cnt;                       // :278-280
		else if (stage == 1) n = 50 - cnt;                           // :284-290
		else if (stage == 2) n = T_F + 1 - counter;                  // :314-315: the (T_F + 1)-th sample is still consumed
		else n = T_null + 51 - counter;                              // :336-337
		if (n > ACQ_CHUNK) n = ACQ_CHUNK;
		if (pos + n > total) { done = 2; break; }                    // out of data: rewind to the attempt start
		const int ph = stage < 2 ? 0 : mod_rate (phi);               // getSample (0) while looking for a signal at all (:279, 285)
		{	// 1. per-sample values
			int l = mod_rate ((long long) lp - (long long) (lane + 1) * ph);
			const int step = mod_rate (32ll * ph);
			for (int i = lane; i < n; i += 32) {
				const float2 v = cmul (win_sample (w, pos + i), nco (T, l));
				const float ja = fabsf (v. x) + fabsf (v. y);            // jan_abs
				s_ax [i] = __dmul_rn (0.00001, (double) ja);
				s_e [i] = stage == 3 ? hypotf (v. x, v. y) : ja;           // abs () in SyncOnEndNull (:329)
				l -= step; if (l < 0) l += DAB_INPUT_RATE;
			}
			__syncwarp ();
			if (stage >= 1)
				for (int i = lane; i < n; i += 32)
					s_d [i] = stage == 1 ? s_e [i] : __fsub_rn (s_e [i], i >= 50 ? s_e [i - 50] : s_ring [(idx + i - 50) & 63]);
			__syncwarp ();
		}
		float sl_end = sLevel, cs_end = cs;
		if (lane == 0) {                                             // 2. the two recurrences
			float a = sLevel, c = cs;
			if (stage == 0) {
#pragma unroll 4
				for (int i = 0; i < n; i ++)
					a = __double2float_rn (__dadd_rn (s_ax [i], __dmul_rn (1 - 0.00001, (double) a)));
			} else {
#pragma unroll 4
				for (int i = 0; i < n; i ++) {
					s_sl [i] = a; s_csb [i] = c;
					a = __double2float_rn (__dadd_rn (s_ax [i], __dmul_rn (1 - 0.00001, (double) a)));
					c = __fadd_rn (c, s_d [i]);
				}
			}
			sl_end = a; cs_end = c;
		}
		__syncwarp ();
		int used = n;
		if (stage >= 2) {                                            // 3. the threshold tests (:301, :323), in parallel
			int first = n;
			for (int i = lane; i < n && first == n; i += 32) {
				const double lhs = (double) (s_csb [i] / 50.0f), lv = (double) s_sl [i];
				const bool leave = stage == 2 ? !(lhs > 0.40 * lv) : !(lhs < 0.75 * lv);
				if (leave) first = i;
			}
			for (int o = 16; o > 0; o >>= 1) first = min (first, __shfl_xor_sync (0xffffffffu, first, o));
			used = first;
		}
		// commit the state after `used` samples
		if (used < n) { sLevel = s_sl [used]; cs = s_csb [used]; }
		else { sLevel = __shfl_sync (0xffffffffu, sl_end, 0); cs = __shfl_sync (0xffffffffu, cs_end, 0); }
		if (stage >= 1) {
			for (int i = lane; i < used; i += 32) if (i >= used - 64) s_ring [(idx + i) & 63] = s_e [i];
			idx += used;
		}
		__syncwarp ();
		pos += used;
		lp = mod_rate ((long long) lp - (long long) used * ph);
		bool restart = false;
		if (stage == 0) { cnt += used; if (cnt == 20 * T. T_s) { stage = 1; cnt = 0; idx = 0; cs = 0.f; } }
		else if (stage == 1) { cnt += used; if (cnt == 50) { stage = 2; counter = 0; } }
		else if (stage == 2) {
			counter += used;
			if (used < n) { stage = 3; counter = 0; }                // :301 fails before sample `used`: on to SyncOnEndNull
			else if (counter > T_F) restart = true;                  // :314-315
		} else {
			counter += used;
			if (used < n) { done = 1; break; }                       // :323 fails: the null symbol has ended, sample `used` is the next to read
			if (counter > T_null + 50) restart = true;               // :336-337
		}
		if (restart) {                                               // goto notSynced
			stage = 0; cnt = 0; counter = 0; idx = 0; cs = 0.f; sLevel = 0.f;
			attempt_pos = pos; attempt_lp = lp;
		}
	}
	if (lane == 0) {
		if (done == 1) { ctl -> synced = 1; ctl -> pos = pos; ctl -> lp = lp; ctl -> acq_done = 1; }
		else           { ctl -> synced = 0; ctl -> pos = attempt_pos; ctl -> lp = attempt_lp; ctl -> acq_done = 0; }
	}
}

// ---------------------------------------------------------------------------------------------------
// predict kernel: frame c of the chunk is assumed to start T_F after frame c-1 with unchanged correctors
// ---------------------------------------------------------------------------------------------------
__global__ void predict_kernel (const StreamCtl *ctl, FrameIn *fin, int T_F, int nframes) {
	const int c = blockIdx. x * blockDim. x + threadIdx. x;
	if (c >= nframes) return;
	const int phi = ctl -> coarse + ctl -> fine;
	FrameIn f;
	f. P = ctl -> pos + (long long) c * T_F;
	f. lp = mod_rate ((long long) ctl -> lp - (long long) c * T_F % DAB_INPUT_RATE * mod_rate (phi));
	f. phiA = f. phiB = phi; f. active = 1; f. pad = 0;
	fin [c] = f;
}

// ---------------------------------------------------------------------------------------------------
// front kernel, one CTA per frame of the chunk: SyncOnPhase + OFDM_PRS (ofdm-processor.cpp:344-406)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (OFDM_THREADS) front_kernel (SampleWin w, OfdmTables T, const FrameIn *fin,
                                                               FrameOut *fo, float2 *spec0) {
	extern __shared__ float2 sm [];
	__shared__ float cv [96];
	const int N = T. T_u, c = blockIdx. x;
	const FrameIn in = fin [c];
	if (!in. active) return;
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
// symbol kernel, CTA (c, g) = frame c of the chunk, symbol group g: OFDM_SYMBOLS (ofdm-processor.cpp:414-442)
// with processToken (ofdm-decoder.cpp:167-190) and the cyclic-prefix correlation (:424-425) fused.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (OFDM_THREADS) symbol_kernel (SampleWin w, OfdmTables T, const FrameIn *fin, int slot0, int groups,
                                                                int blocksPerCIF, int cifsPerFrame, const FrameOut *fo,
                                                                const float2 *spec0, float2 *fcpart, int16_t *fic, int16_t *msc,
                                                                uint8_t *fic8, uint8_t *msc8) {
	extern __shared__ float2 sm [];
	__shared__ float2 s_fc [OFDM_THREADS / 32];
	const int N = T. T_u, Ts = T. T_s, Tg = T. T_g, c = blockIdx. x / groups, g = blockIdx. x % groups;
	const FrameIn in = fin [c];
	if (!in. active) return;
	const int s = fo [c]. startIndex;
	if (s < 0) { if (threadIdx. x == 0) fcpart [c * MAX_GROUPS + g] = make_float2 (0.f, 0.f); return; }
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
	const int slot = slot0 + c;
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
		size_t o; int16_t *out; uint8_t *out8;
		if (l < 4) { o = ((size_t) slot * 3 + (l - 1)) * 2 * T. K; out = fic + o; out8 = fic8 + o; }
		else {
			const int m = l - 4;
			o = ((size_t) 15 + (size_t) slot * cifsPerFrame + m / blocksPerCIF) * CIF_BITS + (size_t) (m % blocksPerCIF) * 2 * T. K;
			out = msc + o; out8 = msc8 + o;
		}
		demod_symbol (f, prev, T, out, out8);
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
// Mode I symbol kernel: same work as symbol_kernel, built on the 8-points-per-thread register FFT.
//   * the raw u8 IQ of symbol l+1 streams into shared memory (one TMA bulk copy + mbarrier, double buffered) while symbol l
//     is transformed, so no thread ever waits on HBM;
//   * samples go shared (raw) -> registers (u8 convert + NCO by a per-thread phasor recurrence, the 1/128 of
//     rawfiles.cpp:113-116 folded into the phasor: an exact power-of-two scaling) -> first butterflies;
//   * the spectrum stays in shared memory in the FFT's own digit-reversed order (the carrier table is
//     pre-permuted), the previous symbol's spectrum is simply the other buffer (pointer swap instead of a copy);
//   * NCO phase indices advance by 32-bit adds (one 64-bit modulo per thread and CTA instead of three per symbol);
//   * a thread demodulates carrier PAIRS, so soft bits leave as 32-bit stores and the Viterbi's byte symbols
//     (viterbi.cpp:229-235) as 16-bit stores.
// ---------------------------------------------------------------------------------------------------
#define R8_RAW 5152                                                        // bytes per raw buffer: 2 T_s + alignment slack, multiple of 16
#define R8_DYN_SMEM ((2 * R8_SMEM + R8_TW2 + R8_TW3) * (int) sizeof (float2) + 2 * R8_RAW)

__device__ __forceinline__ float2 u8_to_c (uchar2 s) {
	return make_float2 ((float) ((int) s. x - 128) * (1.0f / 128.0f), (float) ((int) s. y - 128) * (1.0f / 128.0f));
}
// (b - 128) as float, exactly, b = byte `which` of s: 0x4B0000bb is 2^23 + b
__device__ __forceinline__ float u8_to_f (uint32_t s, int which) {
	return __uint_as_float (__byte_perm (s, 0x4B000000u, which ? 0x7441 : 0x7440)) - 8388736.0f;
}
// soft-bit quantisation of ofdm-decoder.cpp:183-189 with the quotient from the reciprocal unit (2 ulp; the soft bits'
// stated tolerance is +-1 step and comes from the FFT, whose rounding differs from the reference's FFTW anyway)
__device__ __forceinline__ int quant127_fast (float num, float ab1) {
	// (double) q * 127.0 is exact (24 + 7 bits); its truncation equals the truncation of the round-toward-zero float
	// product, because rounding toward zero never crosses an integer (all |integers| <= 127 are floats)
	return __float2int_rz (__fmul_rz (__fdividef (- num, ab1), 127.0f));      // NaN (ab1 == 0) -> 0 (App. B-5)
}

#ifndef R8_MINB
#define R8_MINB 4
#endif
#ifndef R8_TMA
#define R8_TMA 1                                                           // raw symbols staged by TMA bulk copies + mbarrier (0: per-thread cp.async)
#endif
__global__ void __launch_bounds__ (256, R8_MINB) symbol_kernel_r8 (SampleWin w, OfdmTables T, const FrameIn *fin, int slot0, int groups,
                                                          int blocksPerCIF, int cifsPerFrame, const FrameOut *fo,
                                                          const float2 *spec0, float2 *fcpart, int16_t *fic, int16_t *msc,
                                                          uint8_t *fic8, uint8_t *msc8) {
	extern __shared__ __align__ (1024) unsigned char r8_dyn [];     // the FFT buffers must be 512-byte aligned (fft2048_r8)
	float2 *bufA = reinterpret_cast<float2 *> (r8_dyn), *bufB = bufA + R8_SMEM, *tw2 = bufB + R8_SMEM, *tw3 = tw2 + R8_TW2;
	unsigned char *raw = reinterpret_cast<unsigned char *> (tw3 + R8_TW3);
	__shared__ float2 s_fc [8];
	const int N = R8_N, Ts = T. T_s, Tg = T. T_g, t = threadIdx. x;
	const int c = blockIdx. x / groups, g = blockIdx. x % groups;
	const FrameIn in = fin [c];
	if (!in. active) return;
	const int s = fo [c]. startIndex;
	if (s < 0) { if (t == 0) fcpart [c * MAX_GROUPS + g] = make_float2 (0.f, 0.f); return; }
	const int nsym = T. L - 1, per = (nsym + groups - 1) / groups;
	const int l0 = 1 + g * per, l1 = min (nsym + 1, l0 + per);             // symbols [l0, l1)
	const int phA = mod_rate (in. phiA), phB = mod_rate (in. phiB);
	const long long F = in. P + s;                                         // first sample of the PRS
	const int lpD = mod_rate ((long long) in. lp - (long long) (s + N) * phA);  // localPhase after the PRS
	const float2 rot256 = nco (T, mod_rate (- 256ll * phB));               // 256 samples further: phase index - 256 f
	// guard samples this thread correlates with its own useful samples 1536 + t and 1792 + t: gs and gs + 256
	const int gs = t + 1536 - (N - Tg);                                    // = t - 8 (T_u - T_g = 1544); negative: x[6] has no partner
	const int offT = mod_rate ((long long) (gs + 1) * phB), offU = mod_rate ((long long) (Tg + t + 1) * phB);
	const int dTs = mod_rate ((long long) Ts * phB);
	float2 *cur = bufA, *prev = bufB;
	float2 x [8], tw1 [6];
	if (((uint32_t) __cvta_generic_to_shared (bufA) & 511u) != 0) __trap ();   // layout contract of fft2048_r8
	r8_fill_tables (tw2, tw3, T. tw);
	r8_load_tw1 (tw1, T. tw);
#if R8_TMA
	// one mbarrier per raw buffer: the TMA engine's bulk copy of a symbol signals it (complete_tx), every thread waits on it
	__shared__ __align__ (8) unsigned long long s_mbar [2];
	const uint32_t mbar0 = (uint32_t) __cvta_generic_to_shared (&s_mbar [0]);
	if (t == 0) {
		asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r" (mbar0));
		asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r" (mbar0 + 8));
		asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
#endif
	__syncthreads ();

	// raw u8 IQ of symbol l (guard + useful part, T_s samples from `first`) -> raw buffer b; returns the byte offset of
	// sample 0 inside the buffer.  Fast path: 16-byte cp.async from the 16-byte-aligned address below the first sample.
	auto stage = [&] (int l, int b) -> int {
		const long long first = F + N + (long long) (l - 1) * Ts;
		unsigned char *dst = raw + b * R8_RAW;
		const uchar2 *seg = nullptr; long long rel = 0, seglen = 0;
		if (first + Ts <= w. len0) { seg = w. seg0; rel = first; seglen = w. len0; }
		else if (first >= w. len0) { seg = w. seg1; rel = first - w. len0; seglen = w. len1; }
		int off = 0;
		bool fast = seg != nullptr;
		if (fast) {
			const unsigned long long p = (unsigned long long) (seg + rel), pa = p & ~15ull;
			off = (int) (p - pa);
			const int n16 = (off + 2 * Ts + 15) >> 4;
			fast = pa >= (unsigned long long) seg && pa + 16ull * n16 <= (unsigned long long) (seg + seglen);
#if R8_TMA
			if (fast && t == 0) {                                          // one bulk copy by the TMA engine, 16-byte aligned on both sides
				const uint32_t bytes = 16u * (uint32_t) n16, mb = mbar0 + 8u * (uint32_t) b;
				asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's last readers (generic proxy) are past a CTA barrier
				asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r" (mb), "r" (bytes) : "memory");
				asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				              :: "r" ((uint32_t) __cvta_generic_to_shared (dst)), "l" (pa), "r" (bytes), "r" (mb) : "memory");
			}
#else
			if (fast)
				for (int i = t; i < n16; i += 256)
					asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r" ((uint32_t) __cvta_generic_to_shared (dst + 16 * i)), "l" (pa + 16ull * i));
#endif
		}
		if (!fast) {                                                       // symbol straddles the tail | input seam or touches a buffer end
			off = 0;
			for (int i = t; i < Ts; i += 256) reinterpret_cast<uchar2 *> (dst) [i] = win_fetch (w, first + i);
#if R8_TMA
			if (t == 0) asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r" (mbar0 + 8u * (uint32_t) b) : "memory");   // nothing in flight: the CTA barrier below orders the stores
#endif
		}
#if !R8_TMA
		asm volatile ("cp.async.commit_group;");
#endif
		return off;
	};
#if R8_TMA
	auto wait_raw = [&] (int b, int use) {                                 // use = how often buffer b has been waited for before
		const uint32_t mb = mbar0 + 8u * (uint32_t) b, parity = (uint32_t) use & 1u;
		uint32_t ok = 0;
		unsigned long long t0 = 0;
		for (int spin = 0; !ok; spin ++) {
			asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r" (ok) : "r" (mb), "r" (parity) : "memory");
			if (!ok && (spin & 1023) == 1023) {                            // a lost copy must not hang the GPU: give up after 2 s of waiting
				unsigned long long now;
				asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (now));
				if (t0 == 0) t0 = now; else if (now - t0 > 2000000000ull) __trap ();
			}
		}
	};
#endif

	int off_cur = stage (l0, 0);
	if (l0 == 1) {
		const float2 *p0 = spec0 + (size_t) c * N;
		for (int k = t; k < N; k += 256) prev [r8_swz (r8_pos (k))] = p0 [k];
	} else {                                                               // spectrum of symbol l0-1 as reference
		const long long first = F + N + (long long) (l0 - 2) * Ts + Tg;
		const int lpb = mod_rate ((long long) lpD - ((long long) (l0 - 2) * Ts + Tg) % DAB_INPUT_RATE * phB);
		float2 ph = nco (T, mod_rate ((long long) lpb - (long long) (t + 1) * phB));
#pragma unroll
		for (int k = 0; k < 8; k ++) { x [k] = cmul (u8_to_c (win_fetch (w, first + t + 256 * k)), ph); ph = cmul (ph, rot256); }
		fft2048_r8 (x, prev, tw1, tw2, tw3);
	}
	float2 acc = make_float2 (0.f, 0.f);
	const int slot = slot0 + c;
	uint32_t pidx [3];                                                     // this thread's six carriers: the pairs 2t, 2t+1 (+ 512 m), packed
#pragma unroll
	for (int m = 0; m < 3; m ++) pidx [m] = (uint32_t) __ldg (&T. permpos [2 * t + 512 * m]) | ((uint32_t) __ldg (&T. permpos [2 * t + 1 + 512 * m]) << 16);
	// the phase reference of a carrier is the value this very thread read for it one symbol earlier: it stays in registers,
	// and only the first symbol of the group fetches it from the (pseudo-randomly scattered, bank-conflicting) spectrum
	float2 pv [6];
	__syncthreads ();
#pragma unroll
	for (int m = 0; m < 3; m ++) { pv [2 * m] = prev [pidx [m] & 0xffffu]; pv [2 * m + 1] = prev [pidx [m] >> 16]; }
	int lpb = mod_rate ((long long) lpD - ((long long) (l0 - 1) * Ts) % DAB_INPUT_RATE * phB);   // localPhase before the symbol's first sample
	const float sc = 1.0f / 128.0f;
	// the two NCO phasors of a symbol (guard sample gs, useful sample t) are looked up one symbol ahead
	auto phasors = [&] (int lp, float2 &pg, float2 &pu) {
		int ig = lp - offT; if (ig < 0) ig += DAB_INPUT_RATE;
		int iu = lp - offU; if (iu < 0) iu += DAB_INPUT_RATE;
		pg = nco (T, ig); pu = nco (T, iu);
	};
	float2 phg_n, ph_n;
	phasors (lpb, phg_n, ph_n);
	for (int l = l0; l < l1; l ++) {
		const int b = (l - l0) & 1;
		int off_next = 0;
#if R8_TMA
		if (l + 1 < l1) off_next = stage (l + 1, b ^ 1);
#else
		if (l + 1 < l1) off_next = stage (l + 1, b ^ 1); else asm volatile ("cp.async.commit_group;");
#endif
		float2 phg = make_float2 (phg_n. x * sc, phg_n. y * sc), ph = make_float2 (ph_n. x * sc, ph_n. y * sc);
		lpb -= dTs; if (lpb < 0) lpb += DAB_INPUT_RATE;
		phasors (lpb, phg_n, ph_n);
#if R8_TMA
		wait_raw (b, (l - l0) >> 1);
#else
		asm volatile ("cp.async.wait_group 1;" ::: "memory");
#endif
		__syncthreads ();                                                  // raw buffer b complete; last symbol's demod reads done
		const unsigned short *rs = reinterpret_cast<const unsigned short *> (raw + b * R8_RAW + off_cur);
		// guard samples gs and gs + 256 (the ones x[6] and x[7] are correlated with), mixed like every other sample
		float2 g6 = make_float2 (0.f, 0.f), g7;
		{
			const uint32_t r7 = rs [gs + 256];
			g7 = cmul (make_float2 (u8_to_f (r7, 0), u8_to_f (r7, 1)), cmul (phg, rot256));
			if (gs >= 0) { const uint32_t r6 = rs [gs]; g6 = cmul (make_float2 (u8_to_f (r6, 0), u8_to_f (r6, 1)), phg); }
		}
#pragma unroll
		for (int k = 0; k < 8; k ++) {
			const uint32_t v = rs [Tg + t + 256 * k];
			x [k] = cmul (make_float2 (u8_to_f (v, 0), u8_to_f (v, 1)), ph);
			ph = cmul (ph, rot256);
		}
		// FreqCorr += x[i] * conj (x[i - T_u]), i in [T_u, T_s): useful sample e = i - T_g pairs with guard sample e - (T_u - T_g)
		if (gs >= 0) { const float2 r = cmulc (x [6], g6); acc. x += r. x; acc. y += r. y; }
		{ const float2 r = cmulc (x [7], g7); acc. x += r. x; acc. y += r. y; }
		fft2048_r8 (x, cur, tw1, tw2, tw3);
		size_t o;
		if (l < 4) o = ((size_t) slot * 3 + (l - 1)) * 2 * T. K;
		else {
			const int m = l - 4;
			o = ((size_t) 15 + (size_t) slot * cifsPerFrame + m / blocksPerCIF) * CIF_BITS + (size_t) (m % blocksPerCIF) * 2 * T. K;
		}
		int16_t *out = (l < 4 ? fic : msc) + o;
		uint8_t *out8 = (l < 4 ? fic8 : msc8) + o;
#pragma unroll
		for (int m = 0; m < 3; m ++) {                                      // K = 1536 = 3 x 256 carrier pairs
			const int i = 2 * t + 512 * m;
			int re [2], im [2];
#pragma unroll
			for (int q = 0; q < 2; q ++) {
				const int idx = (int) (q ? pidx [m] >> 16 : pidx [m] & 0xffffu);
				const float2 cc = cur [idx];
				const float2 r1 = cmulc (cc, pv [2 * m + q]);
				pv [2 * m + q] = cc;
				const float ab1 = fabsf (r1. x) + fabsf (r1. y);
				re [q] = quant127_fast (r1. x, ab1); im [q] = quant127_fast (r1. y, ab1);
			}
			*reinterpret_cast<uint32_t *> (out + i)        = (uint32_t) (re [0] & 0xffff) | ((uint32_t) re [1] << 16);
			*reinterpret_cast<uint32_t *> (out + T. K + i) = (uint32_t) (im [0] & 0xffff) | ((uint32_t) im [1] << 16);
			*reinterpret_cast<unsigned short *> (out8 + i)        = (unsigned short) ((re [0] + 127) | ((re [1] + 127) << 8));
			*reinterpret_cast<unsigned short *> (out8 + T. K + i) = (unsigned short) ((im [0] + 127) | ((im [1] + 127) << 8));
		}
		off_cur = off_next;
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
		fcpart [c * MAX_GROUPS + g] = sum;
	}
}

// ---------------------------------------------------------------------------------------------------
// Mode I front kernel on the register FFT: same work as front_kernel (SyncOnPhase + OFDM_PRS, ofdm-processor.cpp:344-406).
// The spectrum of a transform stays in the FFT's digit-reversed shared layout; the correlator's product with the PRS
// reference reads it through the position map, the magnitude scan walks the storage order and maps every slot back to
// its bin (first maximum = smallest bin among equals, as the reference's ascending loop finds it).
// ---------------------------------------------------------------------------------------------------
#define FRONT_R8_SMEM ((2 * R8_SMEM + R8_TW2 + R8_TW3) * (int) sizeof (float2))
__host__ __device__ __forceinline__ int r8_pos_inv (int p) { return (p >> 9) | (((p >> 6) & 7) << 2) | (((p >> 3) & 7) << 5) | ((p & 7) << 8); }

__global__ void __launch_bounds__ (256) front_kernel_r8 (SampleWin w, OfdmTables T, const FrameIn *fin, FrameOut *fo, float2 *spec0) {
	extern __shared__ __align__ (1024) unsigned char fr8_dyn [];
	float2 *A = reinterpret_cast<float2 *> (fr8_dyn), *B = A + R8_SMEM, *tw2 = B + R8_SMEM, *tw3 = tw2 + R8_TW2;
	__shared__ float cv [96];
	__shared__ float s_red [8], s_max [8];
	__shared__ int s_idx [8], s_result;
	const int N = R8_N, c = blockIdx. x, t = threadIdx. x;
	const FrameIn in = fin [c];
	if (!in. active) return;
	float2 x [8], tw1 [6];
	r8_fill_tables (tw2, tw3, T. tw);
	r8_load_tw1 (tw1, T. tw);
	const int phA = mod_rate (in. phiA), step256 = mod_rate (256ll * phA);
	// x [k] = sample (first + t + 256 k) after u8 conversion and NCO (rawfiles.cpp:113-116; ofdm-processor.cpp:217-226)
	auto load = [&] (long long first, int lp_before) {
		int lp = mod_rate ((long long) lp_before - (long long) (t + 1) * phA);
#pragma unroll
		for (int k = 0; k < 8; k ++) {
			x [k] = cmul (u8_to_c (win_fetch (w, first + t + 256 * k)), nco (T, lp));
			lp -= step256; if (lp < 0) lp += DAB_INPUT_RATE;
		}
	};
	load (in. P, in. lp);                                              // :347-348
	__syncthreads ();                                                  // twiddle tables in place
	fft2048_r8 (x, A, tw1, tw2, tw3);
	// res = conj (fft * conj (ref)): the backward transform is conj (forward (conj (x))) (phasereference.cpp:66-73)
#pragma unroll
	for (int k = 0; k < 8; k ++) {
		const int kk = t + 256 * k;
		const float2 r = cmulc (A [r8_swz (r8_pos (kk))], __ldg (&T. ref [kk]));
		x [k] = make_float2 (r. x, - r. y);
	}
	__syncthreads ();                                                  // everybody has read A
	fft2048_r8 (x, A, tw1, tw2, tw3);
	const float factor = (float) (1.0 / (float) N);                    // fft.cpp:114-121
	float sum = 0.f, mx = -10000.f;
	int mi = -1;
#pragma unroll
	for (int m = 0; m < 8; m ++) {
		const int slot = t + 256 * m, k = r8_pos_inv (r8_swz (slot));  // the swizzle is an involution
		const float2 v = A [slot];
		const float a = hypotf (v. x * factor, (- v. y) * factor);
		sum += a;
		if (a > mx || (a == mx && k < mi)) { mx = a; mi = k; }
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
		fft2048_r8 (x, A, tw1, tw2, tw3);
		float2 *g = spec0 + (size_t) c * N;
#pragma unroll
		for (int m = 0; m < 8; m ++) {                                 // phaseReference (ofdm-decoder.cpp:91), natural order
			const int k = t + 256 * m;
			const float2 v = A [r8_swz (r8_pos (k))];
			g [k] = v; B [k] = v;
		}
		__syncthreads ();
		corr = coarse_offset_warp0 (B, T, cv);                         // always computed; the scan applies the flag
	}
	if (t == 0) { fo [c]. startIndex = s; fo [c]. correction = corr; }
}

// int16 soft bits -> byte symbols for n elements (the 15 history rows of the time de-interleaver at the start of a call)
__global__ void soft_to_sym8_kernel (const int16_t *in, uint8_t *out, long long n) {
	for (long long i = (long long) blockIdx. x * blockDim. x + threadIdx. x; i < n; i += (long long) gridDim. x * blockDim. x)
		out [i] = (uint8_t) min (max ((int) in [i] + 127, 0), 255);
}

// ---------------------------------------------------------------------------------------------------
// scan kernel (one thread): the scalar state machine of ofdmProcessor::run replayed over the chunk.
//   derive = 1: optimistic pass.  Walks ALL frames, replacing fin[c] by the inputs the replayed state asks
//     for and marking the frames whose inputs changed for recomputation.  Where the data symbols were mixed
//     with another frequency than the replay wants, the cyclic-prefix correlation is corrected by the exact
//     identity  FreqCorr(f') = FreqCorr(f) * exp (-j 2 pi (f' - f) T_u / 2048000).  Nothing is committed.
//   derive = 0: verification pass.  Accepts frames only while the inputs they were actually computed from
//     equal the replayed state, commits the stream state and the per-frame records.
// ---------------------------------------------------------------------------------------------------
#define SCAN_MAX 1024
struct ScanSmem { FrameIn in [SCAN_MAX]; FrameOut fo [SCAN_MAX]; float2 fc [SCAN_MAX]; double inc [SCAN_MAX]; };
#define SCAN_THREADS 256
__global__ void __launch_bounds__ (SCAN_THREADS) scan_kernel (StreamCtl *ctl, FrameIn *fin, int nframes, int slot0, int groups, DabParams dp, const FrameOut *fo,
                             const float2 *fcpart, dabgpu_frame_info *info, long long abs_base, int derive) {
	extern __shared__ unsigned char scan_raw [];
	ScanSmem &S = *reinterpret_cast<ScanSmem *> (scan_raw);
	__shared__ int s_first;
	const int lane = threadIdx. x;                           // (thread index; thread 0 runs the serial replay)
	if (!derive && ctl -> n_redo == 0) return;               // the derive pass found nothing to redo and committed already
	const StreamCtl s0 = *ctl;
	const int cd = dp. carrierDiff;
	if (lane == 0) s_first = nframes;
	// parallel preload of the per-frame records; the serial replay then runs out of shared memory
	for (int c = lane; c < nframes; c += SCAN_THREADS) {
		S. in [c] = fin [c]; S. fo [c] = fo [c];
		float2 fc = make_float2 (0.f, 0.f);
		for (int g = 0; g < groups; g ++) { fc. x += fcpart [c * MAX_GROUPS + g]. x; fc. y += fcpart [c * MAX_GROUPS + g]. y; }
		S. fc [c] = fc;
		S. inc [c] = __dmul_rn (__ddiv_rn (__dmul_rn (0.1, (double) atan2f (fc. y, fc. x)), 3.14159265358979323846), (double) (cd / 2));
	}
	__syncthreads ();
	// Fast path for a locked receiver.  A frame leaves the tracking state as it found it (apart from advancing by one
	// frame length) when its window was placed where the replay wants it, the coarse search is off, findIndex returned
	// T_g and the fine integrator's truncated sum is unchanged.  The leading run of such frames is accepted in parallel;
	// the serial replay starts at the first frame that is not of this kind.
	const int phi0 = s0. coarse + s0. fine;
	const int phm = mod_rate (phi0);
	int first_slow = 0;
	if (!s0. f2 && s0. fine <= cd / 2 && s0. fine >= - cd / 2) {
		for (int c = lane; c < nframes; c += SCAN_THREADS) {
			const FrameIn in = S. in [c];
			const long long P = s0. pos + (long long) c * dp. T_F;
			const int lp = mod_rate ((long long) s0. lp - (long long) c * dp. T_F % DAB_INPUT_RATE * phm);
			const bool ok = in. P == P && in. lp == lp && in. phiA == phi0 && in. phiB == phi0 && S. fo [c]. startIndex == dp. T_s - dp. T_u &&
			                (int) (short) __double2int_rz (__dadd_rn ((double) s0. fine, S. inc [c])) == s0. fine;
			if (!ok) atomicMin (&s_first, c);
		}
		__syncthreads ();
		first_slow = s_first;
		for (int c = lane; c < first_slow; c += SCAN_THREADS) {
			dabgpu_frame_info fi;
			fi. pos = abs_base + s0. pos + (long long) c * dp. T_F; fi. startIndex = dp. T_s - dp. T_u; fi. coarse = s0. coarse; fi. fine = s0. fine;
			fi. phase0 = mod_rate ((long long) s0. lp - (long long) c * dp. T_F % DAB_INPUT_RATE * phm);
			fi. correction = 0; fi. freqCorrRe = S. fc [c]. x; fi. freqCorrIm = S. fc [c]. y;
			info [slot0 + c] = fi;                           // harmless if the chunk is redone: rewritten then
			if (derive) S. in [c]. active = 0;
		}
	}
	__syncthreads ();
	if (lane == 0) {
		StreamCtl s = s0;
		int n_redo = 0;
		int k_a = 0x7fffffff, k_b = 0, k_c = 0, k_si = -1, k_delta = 0;
		s. n_valid = first_slow; s. lost = 0;
		s. pos = s0. pos + (long long) first_slow * dp. T_F;
		s. lp = mod_rate ((long long) s0. lp - (long long) first_slow * dp. T_F % DAB_INPUT_RATE * phm);
		for (int c = first_slow; c < nframes; c ++) {
			FrameIn in = S. in [c];
			const int phiA = s. coarse + s. fine;
			bool changed = in. P != s. pos || in. lp != s. lp || in. phiA != phiA;
			if (derive) { in. P = s. pos; in. lp = s. lp; in. phiA = phiA; }
			else if (changed) break;                             // computed from other inputs than the replay wants
			const int si = S. fo [c]. startIndex;                // (derive: from the old window if `changed`; verified later)
			if (si < 0) {                                        // :353-356 -> notSynced; T_u samples were consumed
				if (derive && changed) {
					in. phiB = phiA; in. active = 1; S. in [c] = in; n_redo ++;
					for (int k = c + 1; k < nframes; k ++) S. in [k]. active = 0;
					nframes = c + 1;
					break;
				}
				if (derive) for (int k = c; k < nframes; k ++) S. in [k]. active = 0;
				s. pos += dp. T_u;
				s. lp = mod_rate ((long long) s. lp - (long long) dp. T_u * mod_rate (phiA));
				s. synced = 0; s. lost = 1;
				break;
			}
			int correction = 0;
			const StreamCtl before = s;
			if (s. f2) {                                         // :390-405
				correction = S. fo [c]. correction;
				if (correction == 0 && s. prev1 == 0 && s. prev2 == 0) s. f2 = 0;
				else if (correction != 100) {
					s. coarse += correction * cd;
					if (abs (s. coarse) > 35000) s. coarse = 0;
					s. prev2 = s. prev1; s. prev1 = correction;
				}
			}
			const int phiB = s. coarse + s. fine;
			const int usedB = in. phiB;
			if (derive) {
				changed = changed || usedB != phiB;
				in. phiB = phiB; in. active = changed; S. in [c] = in; n_redo += changed;
			} else if (usedB != phiB) { s = before; break; }
			const float2 fc = S. fc [c];
			double inc = S. inc [c];                             // :445-446
			if (derive && usedB != phiB) {                       // the symbols were mixed with another frequency: rotate
				double ang = (double) atan2f (fc. y, fc. x);
				ang -= 2.0 * 3.14159265358979323846 * (double) (phiB - usedB) * (double) dp. T_u / (double) DAB_INPUT_RATE;
				ang = remainder (ang, 2.0 * 3.14159265358979323846);
				inc = __dmul_rn (__ddiv_rn (__dmul_rn (0.1, ang), 3.14159265358979323846), (double) (cd / 2));
			}
			dabgpu_frame_info fi;
			fi. pos = abs_base + s. pos; fi. startIndex = si; fi. coarse = s. coarse; fi. fine = s. fine;
			fi. phase0 = s. lp; fi. correction = correction; fi. freqCorrRe = fc. x; fi. freqCorrIm = fc. y;
			info [slot0 + c] = fi;
			// fineCorrector (int16) += 0.1 * arg (FreqCorr) / M_PI * (carrierDiff / 2)
			s. fine = (int) (short) __double2int_rz (__dadd_rn ((double) s. fine, inc));
			const int phiC = s. coarse + s. fine;
			// localPhase after the whole frame: -(si + T_u) fA - (L-1) T_s fB - T_null fC (mod rate); cached while nothing moves
			if (phiA != k_a || phiB != k_b || phiC != k_c || si != k_si) {
				long long d = (long long) (si + dp. T_u) * mod_rate (phiA);
				d += ((long long) (dp. L - 1) * dp. T_s) % DAB_INPUT_RATE * mod_rate (phiB);
				d += (long long) dp. T_null * mod_rate (phiC);                // :453
				k_delta = mod_rate (d); k_a = phiA; k_b = phiB; k_c = phiC; k_si = si;
			}
			s. lp -= k_delta;
			if (s. lp < 0) s. lp += DAB_INPUT_RATE;
			s. pos += si + dp. T_u + (long long) (dp. L - 1) * dp. T_s + dp. T_null;
			if (s. fine > cd / 2) { s. coarse += cd; s. fine -= cd; }      // :458-465
			else if (s. fine < - cd / 2) { s. coarse -= cd; s. fine += cd; }
			s. n_valid = c + 1;
		}
		if (derive) {
			ctl -> n_redo = n_redo;
			if (n_redo == 0) { s. n_redo = 0; *ctl = s; }    // nothing changes: this replay IS the verification
		} else { s. n_redo = ctl -> n_redo; *ctl = s; }
	}
	__syncthreads ();
	if (derive) for (int c = lane; c < nframes; c += SCAN_THREADS) fin [c] = S. in [c];
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int dab_engine_init (dabgpu *h) {
	Engine *E = new Engine ();
	h -> engine = E;
	int rc = ofdm_tables_init (h, &E -> T);
	if (rc) return rc;
	CUDA_TRY (h, cudaMalloc ((void **) &E -> d_phaseRef, (size_t) h -> p. T_u * sizeof (float2)));
	CUDA_TRY (h, E -> d_ctl. ensure (sizeof (StreamCtl)));
	CUDA_TRY (h, E -> h_ctl. ensure (sizeof (StreamCtl)));
	memset (&E -> ctl, 0, sizeof (StreamCtl));
	E -> ctl. f2 = 1; E -> ctl. prev1 = 1000; E -> ctl. prev2 = 999;     // ofdm-processor.cpp:258-259, 73
	E -> groups = h -> p. L > 100 ? 8 : 5;
	if (const char *g = getenv ("DABGPU_GROUPS")) { const int v = atoi (g); if (v >= 1 && v <= MAX_GROUPS) { E -> groups = v; E -> groups_fixed = true; } }   // tuning knob (A/B runs)
	CUDA_TRY (h, E -> d_figkeys. ensure (128 * sizeof (unsigned long long)));
	CUDA_TRY (h, cudaMemsetAsync (E -> d_figkeys. p, 0, 128 * sizeof (unsigned long long), h -> stream));
	CUDA_TRY (h, cudaStreamCreateWithFlags (&E -> copy_st, cudaStreamNonBlocking));
	const int big = 100 * 1024;
	CUDA_TRY (h, cudaFuncSetAttribute (front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
	CUDA_TRY (h, cudaFuncSetAttribute (symbol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
	CUDA_TRY (h, cudaFuncSetAttribute (scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof (ScanSmem)));
	CUDA_TRY (h, cudaFuncSetAttribute (symbol_kernel_r8, cudaFuncAttributeMaxDynamicSharedMemorySize, R8_DYN_SMEM));
	CUDA_TRY (h, cudaFuncSetAttribute (front_kernel_r8, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_R8_SMEM));
	return DABGPU_OK;
}

void dab_engine_free (dabgpu *h) {
	Engine *E = h -> engine;
	if (!E) return;
	for (auto *b : E -> backends) dabgpu_backend_destroy (b);
	if (E -> d_phaseRef) cudaFree (E -> d_phaseRef);
	if (E -> copy_st) { cudaStreamSynchronize (E -> copy_st); cudaStreamDestroy (E -> copy_st); }
	for (auto e : E -> copy_events) cudaEventDestroy (e);
	E -> tail. release (); E -> tail_spare. release (); E -> d_ctl. release (); E -> h_ctl. release ();
	E -> d_frameout. release (); E -> d_framein. release (); E -> d_fcpart. release (); E -> d_spec0. release (); E -> d_info. release ();
	E -> d_fic. release (); E -> d_msc. release (); E -> d_histtmp. release (); E -> d_fic8. release (); E -> d_msc8. release ();
	E -> d_ficbits. release (); E -> d_ficcrc. release (); E -> d_figkeys. release ();
	for (auto &b : E -> d_mscbits) b. release ();
	delete E;
	h -> engine = nullptr;
}

extern "C" int dabgpu_reset (dabgpu_t *h) {                 // ofdmProcessor::reset (ofdm-processor.cpp:476-479)
	if (!h) return DABGPU_ERR_ARG;
	h -> engine -> ctl. fine = h -> engine -> ctl. coarse = 0;
	h -> engine -> ctl. f2 = 1;
	return DABGPU_OK;
}

extern "C" int dabgpu_coarse_corrector (dabgpu_t *h, int32_t on) {   // ofdm-processor.cpp:499-506
	if (!h) return DABGPU_ERR_ARG;
	if (on) { h -> engine -> ctl. f2 = 1; h -> engine -> ctl. coarse = 0; }
	else h -> engine -> ctl. f2 = 0;
	return DABGPU_OK;
}

extern "C" int dabgpu_set_subchannels (dabgpu_t *h, const dabgpu_subch *sc, int32_t nsub) {
	if (!h || nsub < 0 || (nsub > 0 && !sc)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_set_subchannels: bad argument");
	Engine *E = h -> engine;
	std::vector<dabgpu_backend *> nb;
	for (int i = 0; i < nsub; i ++) {
		dabgpu_backend *b = nullptr;
		int rc = dabgpu_backend_create (h, &sc [i], &b);
		if (rc) { for (auto *x : nb) dabgpu_backend_destroy (x); return rc; }
		nb. push_back (b);
	}
	for (auto *b : E -> backends) dabgpu_backend_destroy (b);
	E -> backends = nb;
	E -> subch. assign (sc, sc + nsub);
	for (auto &b : E -> d_mscbits) b. release ();
	E -> d_mscbits. assign (nsub, DevBuf ());
	return DABGPU_OK;
}

extern "C" int dabgpu_get_subch_table (dabgpu_t *h, dabgpu_subch_info *table) {
	if (!h || !table) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_get_subch_table: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	return fig01_table (h, (const unsigned long long *) h -> engine -> d_figkeys. p, table, h -> stream);
}

extern "C" int dabgpu_state_get (dabgpu_t *h, dabgpu_stream_state *s) {
	if (!h || !s) return DABGPU_ERR_ARG;
	const Engine *E = h -> engine;
	s -> synced = E -> ctl. synced; s -> coarse = E -> ctl. coarse; s -> fine = E -> ctl. fine;
	s -> f2Correction = E -> ctl. f2; s -> previous_1 = E -> ctl. prev1; s -> previous_2 = E -> ctl. prev2;
	s -> localPhase = E -> ctl. lp; s -> abs_pos = E -> abs_base + E -> ctl. pos;
	s -> frames = E -> frames_total; s -> cifs = E -> cifs_total;
	return DABGPU_OK;
}

extern "C" int dabgpu_state_set (dabgpu_t *h, const dabgpu_stream_state *s) {
	if (!h || !s) return DABGPU_ERR_ARG;
	Engine *E = h -> engine;
	E -> ctl. synced = s -> synced; E -> ctl. coarse = s -> coarse; E -> ctl. fine = s -> fine;
	E -> ctl. f2 = s -> f2Correction; E -> ctl. prev1 = s -> previous_1; E -> ctl. prev2 = s -> previous_2;
	E -> ctl. lp = s -> localPhase;
	E -> tail_len = 0; E -> ctl. pos = 0; E -> abs_base = s -> abs_pos;      // the next input starts at abs_pos
	E -> frames_total = s -> frames; E -> cifs_total = s -> cifs;
	return DABGPU_OK;
}

extern "C" int dabgpu_host_state_predict (int32_t mode, const dabgpu_stream_state *in, int64_t nframes, dabgpu_stream_state *out) {
	DabParams p;
	if (!in || !out || nframes < 0 || dab_mode_params (mode, &p)) return DABGPU_ERR_ARG;
	if (!in -> synced || in -> f2Correction) return DABGPU_ERR_STATE;
	*out = *in;
	long long phi = ((long long) in -> coarse + in -> fine) % DAB_INPUT_RATE;
	if (phi < 0) phi += DAB_INPUT_RATE;
	const long long adv = (long long) ((__int128) nframes * p. T_F % DAB_INPUT_RATE);
	long long lp = ((long long) in -> localPhase - adv * phi % DAB_INPUT_RATE) % DAB_INPUT_RATE;
	if (lp < 0) lp += DAB_INPUT_RATE;
	out -> localPhase = (int32_t) lp;
	out -> abs_pos = in -> abs_pos + nframes * p. T_F;
	out -> frames = in -> frames + nframes;
	out -> cifs = in -> cifs + nframes * p. cifsPerFrame;
	return DABGPU_OK;
}

static int ensure_frame_capacity (dabgpu *h, long long frames);

// ---- whole stream state as one blob (the multi-GPU hand-over: sync/AFC state + unconsumed samples + the
// 15-CIF soft-bit halo of the time de-interleaver + per-sub-channel warm-up counters) ----
struct StateBlobHeader {
	uint32_t magic; int32_t mode, nsub, hist_valid, cf32, pad;
	StreamCtl ctl; long long abs_base, frames_total, cifs_total, tail_len;
};
#define STATE_MAGIC 0x44414247u

extern "C" int dabgpu_state_export (dabgpu_t *h, void *buf, size_t capacity, size_t *used) {
	if (!h || !used) return DABGPU_ERR_ARG;
	Engine *E = h -> engine;
	const size_t rowb = (size_t) CIF_BITS * sizeof (int16_t);
	const size_t need = sizeof (StateBlobHeader) + E -> backends. size () * sizeof (int64_t) + (size_t) E -> tail_len * E -> sample_bytes () + 15 * rowb;
	*used = need;
	if (!buf) return DABGPU_OK;
	if (capacity < need) return dab_fail (h, DABGPU_ERR_ARG, "state blob needs %zu bytes", need);
	CUDA_TRY (h, cudaSetDevice (h -> device));
	StateBlobHeader hd {};
	hd. cf32 = E -> cf32; hd. pad = 0;
	hd. magic = STATE_MAGIC; hd. mode = h -> p. dabMode; hd. nsub = (int32_t) E -> backends. size (); hd. hist_valid = E -> hist_init;
	hd. ctl = E -> ctl; hd. abs_base = E -> abs_base; hd. frames_total = E -> frames_total; hd. cifs_total = E -> cifs_total; hd. tail_len = E -> tail_len;
	char *q = (char *) buf;
	memcpy (q, &hd, sizeof (hd)); q += sizeof (hd);
	for (auto *b : E -> backends) { int64_t c = dab_backend_cifs_seen (b); memcpy (q, &c, sizeof (c)); q += sizeof (c); }
	if (E -> tail_len) CUDA_TRY (h, cudaMemcpyAsync (q, E -> tail. p, (size_t) E -> tail_len * E -> sample_bytes (), cudaMemcpyDeviceToHost, h -> stream));
	q += (size_t) E -> tail_len * E -> sample_bytes ();
	if (E -> hist_init) CUDA_TRY (h, cudaMemcpyAsync (q, E -> d_msc. p, 15 * rowb, cudaMemcpyDeviceToHost, h -> stream));
	else memset (q, 0, 15 * rowb);
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	return DABGPU_OK;
}

extern "C" int dabgpu_state_import (dabgpu_t *h, const void *buf, size_t n) {
	if (!h || !buf || n < sizeof (StateBlobHeader)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_state_import: bad argument");
	Engine *E = h -> engine;
	StateBlobHeader hd;
	memcpy (&hd, buf, sizeof (hd));
	const size_t rowb = (size_t) CIF_BITS * sizeof (int16_t);
	if (hd. magic != STATE_MAGIC || hd. mode != h -> p. dabMode || hd. nsub != (int32_t) E -> backends. size () || hd. tail_len < 0 ||
	    n != sizeof (hd) + (size_t) hd. nsub * sizeof (int64_t) + (size_t) hd. tail_len * dab_sample_bytes (hd. cf32) + 15 * rowb || hd. cf32 < 0 || hd. cf32 > 2)
		return dab_fail (h, DABGPU_ERR_ARG, "state blob does not match this handle (mode / sub-channel count / size)");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int rc = ensure_frame_capacity (h, 1);
	if (rc) return rc;
	const char *q = (const char *) buf + sizeof (hd);
	for (auto *b : E -> backends) { int64_t c; memcpy (&c, q, sizeof (c)); q += sizeof (c); dab_backend_set_cifs_seen (b, c); }
	E -> ctl = hd. ctl; E -> abs_base = hd. abs_base; E -> frames_total = hd. frames_total; E -> cifs_total = hd. cifs_total;
	E -> tail_len = hd. tail_len;
	E -> cf32 = hd. cf32;
	if (hd. tail_len) {
		CUDA_TRY (h, E -> tail. ensure ((size_t) hd. tail_len * E -> sample_bytes ()));
		CUDA_TRY (h, cudaMemcpyAsync (E -> tail. p, q, (size_t) hd. tail_len * E -> sample_bytes (), cudaMemcpyHostToDevice, h -> stream));
	}
	q += (size_t) hd. tail_len * E -> sample_bytes ();
	CUDA_TRY (h, cudaMemcpyAsync (E -> d_msc. p, q, 15 * rowb, cudaMemcpyHostToDevice, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	E -> hist_init = true;
	return DABGPU_OK;
}

static int ensure_frame_capacity (dabgpu *h, long long frames) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	if (frames <= E -> cap_frames) return DABGPU_OK;
	const long long cap = frames + frames / 8 + 4;
	// the MSC row buffer carries 15 CIFs of history in front: preserve them across a re-allocation
	DevBuf nmsc;
	const size_t rowb = (size_t) CIF_BITS * sizeof (int16_t);
	CUDA_TRY (h, nmsc. ensure ((15 + (size_t) cap * p. cifsPerFrame) * rowb));
	if (E -> d_msc. p && E -> hist_init)
		CUDA_TRY (h, cudaMemcpyAsync (nmsc. p, E -> d_msc. p, 15 * rowb, cudaMemcpyDeviceToDevice, h -> stream));
	else
		CUDA_TRY (h, cudaMemsetAsync (nmsc. p, 0, 15 * rowb, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	E -> d_msc. release ();
	E -> d_msc = nmsc;
	E -> hist_init = true;
	CUDA_TRY (h, E -> d_fic. ensure ((size_t) cap * 3 * 2 * p. K * sizeof (int16_t)));
	CUDA_TRY (h, E -> d_fic8. ensure ((size_t) cap * 3 * 2 * p. K + 16));
	CUDA_TRY (h, E -> d_msc8. ensure ((15 + (size_t) cap * p. cifsPerFrame) * CIF_BITS + 16));
	CUDA_TRY (h, E -> d_info. ensure ((size_t) cap * sizeof (dabgpu_frame_info)));
	CUDA_TRY (h, E -> d_histtmp. ensure (15 * rowb));
	E -> cap_frames = cap;
	return DABGPU_OK;
}

// channel decoding (FIC + every configured sub-channel) of the frames [f0, f0 + nv) just accepted by the OFDM part,
// queued on a side stream so that it overlaps the OFDM work of the next chunk; results are copied to the caller's
// buffers on the same stream
static int channel_chunk (dabgpu *h, int f0, int nv, dabgpu_result *out, std::vector<int> &nblk) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	const int ngroups = nv * p. ficGroups, g0 = f0 * p. ficGroups, ncif = nv * p. cifsPerFrame, r0 = f0 * p. cifsPerFrame;
	if (nv <= 0) return DABGPU_OK;
	h -> cur = 1 + (E -> vrr ++ % 3);
	cudaStream_t st = h -> vst ();
	int rc = DABGPU_OK;
	// the throughput kernels take every job of the chunk in ONE launch pair; the warp-cooperative kernel is launched once per
	// sub-channel, 0.3 ms each however few code words there are -- so with two or more sub-channels the throughput path wins
	// even for a single frame (measured: 9 sub-channels, 1-32 frames per call: 2.9 ms against 0.6 ms)
	const bool simd = dab_use_simd (h, (long long) ngroups + (long long) ncif * (long long) E -> backends. size ()) ||
	                  (h -> cfg. viterbi_path == 0 && E -> backends. size () >= 2 && ncif > 0);
	std::vector<VitSimdJob> jobs;
	uint8_t *ficbits = (uint8_t *) E -> d_ficbits. p + (size_t) g0 * 768, *ficcrc = (uint8_t *) E -> d_ficcrc. p + (size_t) g0 * 3;
	do {
		if (ngroups > 0) {
			const int16_t *soft = (const int16_t *) E -> d_fic. p + (size_t) f0 * 3 * 2 * p. K;
			if (simd) {
				jobs. emplace_back ();
				if ((rc = dab_fic_simd_job (h, soft, 2304, ngroups, ficbits, &jobs. back ()))) break;
				jobs. back (). sym8 = (uint8_t *) E -> d_fic8. p + (size_t) f0 * 3 * 2 * p. K;   // written by the symbol kernel
				jobs. back (). stride8 = 2304;
			} else if ((rc = dab_fic_decode_dev (h, soft, 2304, ngroups, ficbits, ficcrc))) break;
		}
		std::vector<int> n_here (E -> backends. size (), 0);
		for (size_t i = 0; i < E -> backends. size () && ncif > 0; i ++) {
			const dabgpu_subch &sc = E -> subch [i];
			VitSimdJob job;
			uint8_t *dst = (uint8_t *) E -> d_mscbits [i]. p + (size_t) nblk [i] * 24 * sc. bitRate;
			if ((rc = dab_backend_run_dev (E -> backends [i], (const int16_t *) E -> d_msc. p + (size_t) sc. startAddr * 64, CIF_BITS, r0, ncif,
			                               dst, &n_here [i], simd ? &job : nullptr))) break;
			if (simd && n_here [i] > 0) {
				job. sym8 = (uint8_t *) E -> d_msc8. p + (size_t) (job. first_row - 15) * CIF_BITS + (size_t) sc. startAddr * 64;
				job. stride8 = CIF_BITS;
				jobs. push_back (job);
			}
			dab_backend_note_cifs (E -> backends [i], ncif);
		}
		if (rc) break;
		if (simd) {
			if ((rc = dab_vit_simd_run (h, jobs))) break;
			if (ngroups > 0) { cudaError_t e = fib_crc_launch (h, ficbits, 3 * ngroups, ficcrc); if (e != cudaSuccess) { rc = dab_fail (h, DABGPU_ERR_CUDA, "crc launch: %s", cudaGetErrorString (e)); break; } }
		}
		if (ngroups > 0) {                                   // FIG 0/1 of the FIBs just checked (fib-processor.cpp:278-347)
			cudaError_t fe = fig01_launch (h, ficbits, ficcrc, 3 * ngroups, (unsigned long long) (E -> frames_total * p. ficGroups + g0) * 3ull,
			                               (unsigned long long *) E -> d_figkeys. p, st);
			if (fe != cudaSuccess) { rc = dab_fail (h, DABGPU_ERR_CUDA, "fig scan launch: %s", cudaGetErrorString (fe)); break; }
		}
		cudaError_t e = cudaSuccess;
		if (ngroups > 0 && out -> fic_bits) e = cudaMemcpyAsync (out -> fic_bits + (size_t) g0 * 768, ficbits, (size_t) ngroups * 768, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && ngroups > 0 && out -> fic_crc) e = cudaMemcpyAsync (out -> fic_crc + (size_t) g0 * 3, ficcrc, (size_t) ngroups * 3, cudaMemcpyDeviceToHost, st);
		for (size_t i = 0; i < E -> backends. size () && e == cudaSuccess; i ++) {
			const size_t fb = (size_t) 24 * E -> subch [i]. bitRate;
			if (out -> msc_bits && out -> msc_bits [i] && n_here [i] > 0)
				e = cudaMemcpyAsync (out -> msc_bits [i] + (size_t) nblk [i] * fb, (uint8_t *) E -> d_mscbits [i]. p + (size_t) nblk [i] * fb,
				                     (size_t) n_here [i] * fb, cudaMemcpyDeviceToHost, st);
			nblk [i] += n_here [i];
		}
		if (e != cudaSuccess) rc = dab_fail (h, DABGPU_ERR_CUDA, "result copy: %s", cudaGetErrorString (e));
	} while (0);
	h -> cur = 0;
	return rc;
}

// decode core on a device-resident input segment.  `ready` (optional): events of the piecewise host-to-device copy
// of the input; piece k covers new-segment samples [k * piece, (k+1) * piece).
static int decode_core (dabgpu *h, const void *d_new_v, long long nnew, dabgpu_result *out,
                        const std::vector<cudaEvent_t> *ready = nullptr, long long piece = 0, int vit_batch_frames = 0x7fffffff) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	if (p. dabMode == 3)
		return dab_fail (h, DABGPU_ERR_ARG, "stream decode is not available for Mode III (the reference has no Mode III framing either)");
	const uchar2 *d_new = (const uchar2 *) d_new_v;
	const size_t sb = E -> sample_bytes ();
	SampleWin w { (const uchar2 *) E -> tail. p, E -> tail_len, d_new, nnew, E -> cf32 };
	const long long total = E -> tail_len + nnew;
	const long long frame_need = 2ll * p. T_u + (long long) (p. L - 1) * p. T_s + p. T_null;   // worst case from P
	const long long max_frames_possible = total / p. T_F + 2;
	long long want = out -> max_frames < max_frames_possible ? out -> max_frames : max_frames_possible;
	if (want < 0) want = 0;
	int rc = ensure_frame_capacity (h, want);
	if (rc) return rc;
	CUDA_TRY (h, E -> d_frameout. ensure ((size_t) E -> max_chunk * sizeof (FrameOut)));
	CUDA_TRY (h, E -> d_framein. ensure ((size_t) E -> max_chunk * sizeof (FrameIn)));
	CUDA_TRY (h, E -> d_fcpart. ensure ((size_t) E -> max_chunk * MAX_GROUPS * sizeof (float2)));
	CUDA_TRY (h, E -> d_spec0. ensure ((size_t) E -> max_chunk * p. T_u * sizeof (float2)));
	// result buffers on the device for the whole call (side streams write into them chunk by chunk)
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	CUDA_TRY (h, E -> d_ficbits. ensure ((size_t) (want * p. ficGroups + 1) * 768));
	CUDA_TRY (h, E -> d_ficcrc. ensure ((size_t) (want * p. ficGroups + 1) * 3));
	for (size_t i = 0; i < E -> backends. size (); i ++)
		CUDA_TRY (h, E -> d_mscbits [i]. ensure ((size_t) (want * p. cifsPerFrame + 1) * 24 * E -> subch [i]. bitRate));
	// the 15 history rows of the time de-interleaver as byte symbols (whatever put them there: the previous call, a state
	// import, a re-allocation)
	soft_to_sym8_kernel<<<296, 256, 0, h -> stream>>> ((const int16_t *) E -> d_msc. p, (uint8_t *) E -> d_msc8. p, 15ll * CIF_BITS);
	h -> launches ++;
	std::vector<int> nblk (E -> backends. size (), 0);
	StreamCtl *hctl = (StreamCtl *) E -> h_ctl. p;
	int nframes = 0, decoded_upto = 0;
	// host input arrives piecewise: smaller chunks let the first frames start before the last samples are up
	// device-resident input with a channel-decoding batch given (dabgpu_config.reserved[1]): chunks no larger than the batch, so
	// that the OFDM part of the next chunk overlaps the channel decoding of this one on the side streams
	const int dev_cap = !ready && vit_batch_frames < E -> max_chunk ? (vit_batch_frames > 16 ? vit_batch_frames : 16) : E -> max_chunk;
	const int chunk_cap = ready && !ready -> empty () ? (E -> max_chunk < 128 ? E -> max_chunk : 128) : dev_cap;
	long long waited = -1;                                   // input pieces [0, waited] are known to have arrived on the main stream
	auto need_input = [&] (long long upto_window_pos) -> cudaError_t {      // samples before this window position must be resident
		if (!ready || ready -> empty ()) return cudaSuccess;
		long long rel = upto_window_pos - E -> tail_len;
		if (rel <= 0) return cudaSuccess;
		long long k = (rel - 1) / piece;
		if (k >= (long long) ready -> size ()) k = (long long) ready -> size () - 1;
		cudaError_t e = cudaSuccess;
		if (k > waited) { e = cudaStreamWaitEvent (h -> stream, (*ready) [k], 0); waited = k; }    // copies are in order on one stream
		return e;
	};
	const size_t sm_front = 2 * (size_t) p. T_u * sizeof (float2), sm_sym = ((size_t) p. T_s + 2 * p. T_u) * sizeof (float2);
	while (nframes < want) {
		if (!E -> ctl. synced) {
			CUDA_TRY (h, need_input (total));                // the null search reads until it finds one
			*hctl = E -> ctl;
			CUDA_TRY (h, cudaMemcpyAsync (E -> d_ctl. p, hctl, sizeof (StreamCtl), cudaMemcpyHostToDevice, h -> stream));
			{ ProfScope prof (h, KC_ACQUIRE);
			acquire_kernel<<<1, 32, 0, h -> stream>>> (w, E -> T, p. T_F, p. T_null, (StreamCtl *) E -> d_ctl. p); }
			h -> launches ++;
			CUDA_TRY (h, cudaGetLastError ());
			CUDA_TRY (h, cudaMemcpyAsync (hctl, E -> d_ctl. p, sizeof (StreamCtl), cudaMemcpyDeviceToHost, h -> stream));
			CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
			E -> ctl = *hctl;
			if (!E -> ctl. synced) break;                   // ran out of samples inside the attempt
			E -> chunk = 1;
		}
		long long avail = (total - E -> ctl. pos - frame_need) / p. T_F + 1;
		if (total - E -> ctl. pos < frame_need) avail = 0;
		if (avail <= 0) break;
		long long C = E -> chunk < chunk_cap ? E -> chunk : chunk_cap;
		if (C > avail) C = avail;
		if (C > want - nframes) C = want - nframes;
		CUDA_TRY (h, need_input (E -> ctl. pos + (C - 1) * p. T_F + frame_need));
		*hctl = E -> ctl;
		CUDA_TRY (h, cudaMemcpyAsync (E -> d_ctl. p, hctl, sizeof (StreamCtl), cudaMemcpyHostToDevice, h -> stream));
		StreamCtl *dctl = (StreamCtl *) E -> d_ctl. p;
		FrameIn *fin = (FrameIn *) E -> d_framein. p;
		FrameOut *fo = (FrameOut *) E -> d_frameout. p;
		{ ProfScope prof (h, KC_SCAN);
		predict_kernel<<<((int) C + 127) / 128, 128, 0, h -> stream>>> (dctl, fin, p. T_F, (int) C); }
		// symbol groups (CTAs) per frame: every group first recomputes the spectrum of the symbol before its own as phase
		// reference, so few groups mean less redundant work and many groups more CTAs.  A big chunk fills the GPU anyway
		// (A/B on B200, 1024 frames: 5 groups 0.849 ms, 3 groups 0.828 ms, 1-2 groups the same)
		const int groups = !E -> groups_fixed && C >= 128 && E -> groups > 3 ? 3 : E -> groups;
		for (int pass = 0; pass < 2; pass ++) {
			// pass 0: speculative inputs, then the optimistic replay (derive); pass 1: recompute what changed, then verify
			{ ProfScope prof (h, KC_FRONT);
			if (p. T_u == R8_N && p. K == 1536 && !E -> cf32 && !h -> cfg. reserved [0])
				front_kernel_r8<<<(int) C, 256, FRONT_R8_SMEM, h -> stream>>> (w, E -> T, fin, fo, (float2 *) E -> d_spec0. p);
			else
				front_kernel<<<(int) C, OFDM_THREADS, sm_front, h -> stream>>> (w, E -> T, fin, fo, (float2 *) E -> d_spec0. p); }
			{ ProfScope prof (h, KC_SYMBOL);
			if (p. T_u == R8_N && p. K == 1536 && !E -> cf32 && !h -> cfg. reserved [0])     // reserved[0] = 1: generic kernel (A/B testing)
				symbol_kernel_r8<<<(int) C * groups, 256, R8_DYN_SMEM, h -> stream>>> (w, E -> T, fin, nframes, groups, p. blocksPerCIF, p. cifsPerFrame,
					fo, (const float2 *) E -> d_spec0. p, (float2 *) E -> d_fcpart. p, (int16_t *) E -> d_fic. p, (int16_t *) E -> d_msc. p,
					(uint8_t *) E -> d_fic8. p, (uint8_t *) E -> d_msc8. p);
			else
				symbol_kernel<<<(int) C * groups, OFDM_THREADS, sm_sym, h -> stream>>> (w, E -> T, fin, nframes, groups, p. blocksPerCIF, p. cifsPerFrame,
					fo, (const float2 *) E -> d_spec0. p, (float2 *) E -> d_fcpart. p, (int16_t *) E -> d_fic. p, (int16_t *) E -> d_msc. p,
					(uint8_t *) E -> d_fic8. p, (uint8_t *) E -> d_msc8. p); }
			{ ProfScope prof (h, KC_SCAN);
			scan_kernel<<<1, SCAN_THREADS, sizeof (ScanSmem), h -> stream>>> (dctl, fin, (int) C, nframes, groups, p, fo, (const float2 *) E -> d_fcpart. p,
				(dabgpu_frame_info *) E -> d_info. p, E -> abs_base, pass == 0 ? 1 : 0); }
		}
		h -> launches += 7;
		CUDA_TRY (h, cudaGetLastError ());
		CUDA_TRY (h, cudaMemcpyAsync (hctl, E -> d_ctl. p, sizeof (StreamCtl), cudaMemcpyDeviceToHost, h -> stream));
		CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
		E -> ctl = *hctl;
		nframes += E -> ctl. n_valid;
		// the accepted frames' soft bits are final: once enough of them have piled up to fill the GPU, decode them on a
		// side stream while the OFDM part (and the input copy) of the following frames goes on
		if (nframes - decoded_upto >= vit_batch_frames) {
			if ((rc = channel_chunk (h, decoded_upto, nframes - decoded_upto, out, nblk))) return rc;
			decoded_upto = nframes;
		}
		// chunk policy: grow while the speculation holds outright; a chunk that needed recomputation (correctors still
		// moving) or was cut short shrinks, so a converging loop costs little and a locked one runs in one pass
		if (E -> ctl. n_valid == C && E -> ctl. n_redo == 0) { E -> chunk *= 2; if (E -> chunk > chunk_cap) E -> chunk = chunk_cap; }
		else if (E -> ctl. n_valid == C) { E -> chunk = E -> chunk > 16 ? E -> chunk / 2 : (E -> chunk < 8 ? E -> chunk * 2 : E -> chunk); }
		else { E -> chunk = E -> chunk > 2 ? E -> chunk / 2 : 1; }
	}
	out -> nframes = nframes;
	const int ncif = nframes * p. cifsPerFrame;
	if ((rc = channel_chunk (h, decoded_upto, nframes - decoded_upto, out, nblk))) return rc;
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	for (size_t i = 0; i < E -> backends. size (); i ++)
		if (out -> msc_nblocks) out -> msc_nblocks [i] = nblk [i];
	if (nframes > 0) {
		if (out -> info) CUDA_TRY (h, cudaMemcpyAsync (out -> info, E -> d_info. p, (size_t) nframes * sizeof (dabgpu_frame_info), cudaMemcpyDeviceToHost, h -> stream));
		if (out -> soft) {
			const size_t dpitch = (size_t) (p. L - 1) * 2 * p. K * sizeof (int16_t), fw = (size_t) 3 * 2 * p. K * sizeof (int16_t);
			const size_t mw = (size_t) p. cifsPerFrame * CIF_BITS * sizeof (int16_t);
			CUDA_TRY (h, cudaMemcpy2DAsync (out -> soft, dpitch, E -> d_fic. p, fw, fw, nframes, cudaMemcpyDeviceToHost, h -> stream));
			CUDA_TRY (h, cudaMemcpy2DAsync ((char *) out -> soft + fw, dpitch, (const int16_t *) E -> d_msc. p + (size_t) 15 * CIF_BITS, mw, mw, nframes,
			                                cudaMemcpyDeviceToHost, h -> stream));
		}
		// time de-interleaver history for the next call: the last 15 CIF rows move to the front
		const size_t rowb = (size_t) CIF_BITS * sizeof (int16_t);
		CUDA_TRY (h, cudaMemcpyAsync (E -> d_histtmp. p, (const char *) E -> d_msc. p + (size_t) ncif * rowb, 15 * rowb, cudaMemcpyDeviceToDevice, h -> stream));
		CUDA_TRY (h, cudaMemcpyAsync (E -> d_msc. p, E -> d_histtmp. p, 15 * rowb, cudaMemcpyDeviceToDevice, h -> stream));
	}
	// ---- keep the unconsumed samples for the next call ----
	CUDA_TRY (h, need_input (total));
	const long long consumed = E -> ctl. pos;               // everything before `pos` is done with
	const long long keep = total - consumed;
	out -> consumed = consumed - E -> tail_len;               // relative to this call's input (may be negative: none of it)
	if (keep > 0) {
		DevBuf &nt = E -> tail_spare;                       // ping-pong: no allocation once both buffers are big enough
		CUDA_TRY (h, nt. ensure ((size_t) keep * sb));
		long long from0 = consumed < E -> tail_len ? E -> tail_len - consumed : 0;     // part still in the old tail
		if (from0 > 0)
			CUDA_TRY (h, cudaMemcpyAsync (nt. p, (const char *) E -> tail. p + (size_t) consumed * sb, (size_t) from0 * sb, cudaMemcpyDeviceToDevice, h -> stream));
		const long long off1 = consumed > E -> tail_len ? consumed - E -> tail_len : 0;
		CUDA_TRY (h, cudaMemcpyAsync ((char *) nt. p + (size_t) from0 * sb, (const char *) d_new + (size_t) off1 * sb, (size_t) (nnew - off1) * sb, cudaMemcpyDeviceToDevice, h -> stream));
		CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
		DevBuf t = E -> tail; E -> tail = E -> tail_spare; E -> tail_spare = t;
	} else
		CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	E -> tail_len = keep > 0 ? keep : 0;
	E -> abs_base += consumed;
	E -> ctl. pos = 0;
	E -> frames_total += nframes; E -> cifs_total += ncif;
	return DABGPU_OK;
}

// the sample format of a stream may only change while no unconsumed samples are pending
static int set_format (dabgpu *h, int cf32) {
	Engine *E = h -> engine;
	if (E -> cf32 != cf32 && E -> tail_len > 0)
		return dab_fail (h, DABGPU_ERR_STATE, "sample format changed while %lld samples of the other format are pending", E -> tail_len);
	E -> cf32 = cf32;
	return DABGPU_OK;
}

static int decode_host (dabgpu *h, const void *iq, size_t nsamples, int cf32, dabgpu_result *out);

extern "C" int dabgpu_decode_dev (dabgpu_t *h, const uint8_t *d_iq_u8, size_t nsamples, dabgpu_result *out) {
	if (!h || !out || (nsamples > 0 && !d_iq_u8)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int rc = set_format (h, 0);
	if (rc) return rc;
	return decode_core (h, d_iq_u8, (long long) nsamples, out, nullptr, 0, h -> cfg. reserved [1] > 0 ? h -> cfg. reserved [1] : 0x7fffffff);
}

extern "C" int dabgpu_decode_cf32_dev (dabgpu_t *h, const float *d_iq, size_t nsamples, dabgpu_result *out) {
	if (!h || !out || (nsamples > 0 && !d_iq)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode_cf32: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int rc = set_format (h, 1);
	if (rc) return rc;
	return decode_core (h, d_iq, (long long) nsamples, out);
}

extern "C" int dabgpu_decode_i16_dev (dabgpu_t *h, const int16_t *d_iq, size_t nsamples, dabgpu_result *out) {
	if (!h || !out || (nsamples > 0 && !d_iq)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode_i16: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int rc = set_format (h, 2);
	if (rc) return rc;
	return decode_core (h, d_iq, (long long) nsamples, out);
}

extern "C" int dabgpu_decode (dabgpu_t *h, const uint8_t *iq_u8, size_t nsamples, dabgpu_result *out) { return decode_host (h, iq_u8, nsamples, 0, out); }
extern "C" int dabgpu_decode_i16 (dabgpu_t *h, const int16_t *iq, size_t nsamples, dabgpu_result *out) { return decode_host (h, iq, nsamples, 2, out); }
extern "C" int dabgpu_decode_cf32 (dabgpu_t *h, const float *iq, size_t nsamples, dabgpu_result *out) { return decode_host (h, iq, nsamples, 1, out); }

static int decode_host (dabgpu *h, const void *iq_v, size_t nsamples, int cf32, dabgpu_result *out) {
	const uint8_t *iq_u8 = (const uint8_t *) iq_v;
	if (!h || !out || (nsamples > 0 && !iq_u8)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	Engine *E = h -> engine;
	int frc = set_format (h, cf32);
	if (frc) return frc;
	const size_t sb = E -> sample_bytes ();
	const size_t bytes = nsamples * sb;
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	CUDA_TRY (h, h -> d_in. ensure (bytes + 16));
	// the input goes up in pieces on its own stream; every chunk of frames waits only for the pieces it reads, so
	// the host-to-device copy overlaps the decode of the frames already there
	const long long piece = 8ll << 20;                       // samples per piece (16 MB)
	std::vector<cudaEvent_t> ready;
	if (bytes) {
		cudaPointerAttributes attr;
		const bool pinned = cudaPointerGetAttributes (&attr, iq_u8) == cudaSuccess && attr. type == cudaMemoryTypeHost;
		cudaGetLastError ();
		const uint8_t *src = iq_u8;
		if (!pinned) {
			CUDA_TRY (h, h -> h_in. ensure (bytes));
			memcpy (h -> h_in. p, iq_u8, bytes);
			src = (const uint8_t *) h -> h_in. p;
		}
		const size_t npieces = (nsamples + piece - 1) / piece;
		while (E -> copy_events. size () < npieces) {
			cudaEvent_t e;
			CUDA_TRY (h, cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
			E -> copy_events. push_back (e);
		}
		for (size_t k = 0; k < npieces; k ++) {
			const size_t off = k * (size_t) piece * sb, len = (k + 1 == npieces ? bytes - off : (size_t) piece * sb);
			CUDA_TRY (h, cudaMemcpyAsync ((char *) h -> d_in. p + off, src + off, len, cudaMemcpyHostToDevice, E -> copy_st));
			CUDA_TRY (h, cudaEventRecord (E -> copy_events [k], E -> copy_st));
			ready. push_back (E -> copy_events [k]);
		}
	}
	// host input: the PCIe copy paces the call and the GPU idles most of the time, so channel decoding follows the
	// OFDM part in small batches: what is left to do once the last sample has arrived is then short
	int rc = decode_core (h, h -> d_in. p, (long long) nsamples, out, &ready, piece,
	                      h -> cfg. host_batch_frames > 0 ? h -> cfg. host_batch_frames : E -> vit_batch_frames);
	cudaStreamSynchronize (E -> copy_st);
	return rc;
}
