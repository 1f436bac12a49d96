// dabgpu_sync.cu -- the sequential part of ofdmProcessor::run (ofdm-processor.cpp:247-474) on the device: null-symbol
// acquisition, the per-chunk prediction of the tracking state and the scan that replays the reference's scalar state
// machine over the per-frame results.  Every kernel takes an array of streams (StreamDev): dabgpu_decode runs one,
// dabgpu_decode_multi many -- one warp / CTA per stream, so n acquisitions or replays cost the time of one.
#include <math.h>
#include "dabgpu_engine.h"

__device__ __forceinline__ uchar2 sy_fetch (const SampleWin &w, long long i) {      // u8 windows only
	return i < w. len0 ? __ldg (&w. seg0 [i]) : __ldg (&w. seg1 [i - w. len0]);
}
// sample i as the complex float the reference's getSample sees before the NCO (ofdm-processor.cpp:133-183)
__device__ __forceinline__ float2 sy_sample (const SampleWin &w, long long i) {
	if (w. cf32 == 1)
		return i < w. len0 ? __ldg (reinterpret_cast<const float2 *> (w. seg0) + i) : __ldg (reinterpret_cast<const float2 *> (w. seg1) + (i - w. len0));
	if (w. cf32 == 2) {                                      // 16-bit PCM as sf_readf_float delivers it (wavfiles.cpp:190): x / 32768, exact in float
		const short2 v = i < w. len0 ? __ldg (reinterpret_cast<const short2 *> (w. seg0) + i) : __ldg (reinterpret_cast<const short2 *> (w. seg1) + (i - w. len0));
		return make_float2 ((float) v. x * (1.0f / 32768.0f), (float) v. y * (1.0f / 32768.0f));
	}
	const uchar2 s = sy_fetch (w, i);
	return make_float2 ((float) ((int) s. x - 128) * (1.0f / 128.0f), (float) ((int) s. y - 128) * (1.0f / 128.0f));
}

// the same in two steps -- the bare load (nothing waits for it) and the conversion -- for loads issued long before their use
__device__ __forceinline__ uint2 sy_raw (const SampleWin &w, long long i) {
	const bool first = i < w. len0;
	const long long k = first ? i : i - w. len0;
	const uchar2 *seg = first ? w. seg0 : w. seg1;
	if (w. cf32 == 1) { const float2 v = __ldg (reinterpret_cast<const float2 *> (seg) + k); return make_uint2 (__float_as_uint (v. x), __float_as_uint (v. y)); }
	if (w. cf32 == 2) return make_uint2 (__ldg (reinterpret_cast<const unsigned int *> (seg) + k), 0u);
	return make_uint2 ((unsigned int) __ldg (reinterpret_cast<const unsigned short *> (seg) + k), 0u);
}
// the thread's ACQ_CHUNK / ACQ_THREADS samples of the chunk starting at window position p (those below `end`): the chunk
// normally lies inside one segment, so the segment and the format are picked once, not per sample
template <int NPER, int STRIDE>
__device__ __forceinline__ void sy_raw_chunk (const SampleWin &w, long long p, long long end, int tid, uint2 (&r) [NPER]) {
	const bool in0 = p + (long long) NPER * STRIDE <= w. len0, in1 = p >= w. len0;
	if (in0 || in1) {
		const uchar2 *seg = in0 ? w. seg0 : w. seg1;
		const long long k0 = (in0 ? p : p - w. len0) + tid;
		const int lim = (int) (end - p < (long long) NPER * STRIDE ? end - p : (long long) NPER * STRIDE) - tid;     // samples k STRIDE < lim exist
		if (w. cf32 == 1) {
#pragma unroll
			for (int k = 0; k < NPER; k ++) if (k * STRIDE < lim) { const float2 v = __ldg (reinterpret_cast<const float2 *> (seg) + k0 + k * STRIDE); r [k] = make_uint2 (__float_as_uint (v. x), __float_as_uint (v. y)); }
		} else if (w. cf32 == 2) {
#pragma unroll
			for (int k = 0; k < NPER; k ++) if (k * STRIDE < lim) r [k] = make_uint2 (__ldg (reinterpret_cast<const unsigned int *> (seg) + k0 + k * STRIDE), 0u);
		} else {
#pragma unroll
			for (int k = 0; k < NPER; k ++) if (k * STRIDE < lim) r [k] = make_uint2 ((unsigned int) __ldg (reinterpret_cast<const unsigned short *> (seg) + k0 + k * STRIDE), 0u);
		}
	} else {
#pragma unroll
		for (int k = 0; k < NPER; k ++) if (p + tid + k * STRIDE < end) r [k] = sy_raw (w, p + tid + k * STRIDE);
	}
}
__device__ __forceinline__ float2 sy_conv (const SampleWin &w, uint2 r) {
	if (w. cf32 == 1) return make_float2 (__uint_as_float (r. x), __uint_as_float (r. y));
	if (w. cf32 == 2) return make_float2 ((float) (short) (r. x & 0xffffu) * (1.0f / 32768.0f), (float) (short) (r. x >> 16) * (1.0f / 32768.0f));
	return make_float2 ((float) ((int) (r. x & 0xffu) - 128) * (1.0f / 128.0f), (float) ((int) ((r. x >> 8) & 0xffu) - 128) * (1.0f / 128.0f));
}

// ---------------------------------------------------------------------------------------------------
// acquisition: notSynced -> SyncOnNull -> SyncOnEndNull (ofdm-processor.cpp:275-338), one CTA of eight warps per stream.
// The reference walks the samples one by one through two recurrences -- the signal level IIR
// (sLevel = 0.00001 * jan_abs (v) + (1 - 0.00001) * sLevel, in double, rounded to float, :168) and the running sum of a
// 50-sample envelope window -- and tests a threshold before every sample.  Only the two recurrences are serial.  Per chunk:
//   1. all threads convert / mix the samples and prepare everything that does not depend on the recurrences: the
//      envelope value e_i (|re| + |im|, or the true magnitude in SyncOnEndNull), jan_abs, the window difference e_i - e_(i-50);
//   2. thread 0 runs the two recurrences and records the values BEFORE every sample.  The window sum is a float add chain
//      (4 cycles a step).  The level IIR as the reference writes it is a chain of four dependent double-precision operations
//      (widen, multiply, add, round to float: ~85 cycles a step on this part, 12-20 ms per search), so thread 0 runs a float
//      SURROGATE instead, a' = fma (1e-5f, |v| - a, a) (two dependent float operations), and
//   2b. all threads VERIFY it: for every step the reference's own double expression is evaluated on the recorded a_i and
//      compared with the recorded a_(i+1).  The surrogate's error is ~1e-12 a against a float spacing of ~1e-7 a, so the
//      two differ only when the exact value sits within 1e-12 a of a rounding tie (~3e-5 of the steps); the first such
//      step is repaired with the exact value and the chain re-run from there.  By induction the recorded sequence is the
//      reference's, bit for bit, whatever the surrogate did;
//   3. all threads evaluate the reference's threshold tests on the recorded values and the first sample that leaves the
//      state is found by a reduction; the state is committed up to there.
// ---------------------------------------------------------------------------------------------------
#define ACQ_CHUNK 1024
#ifdef ACQ_PROF
#define ACQ_T(k) { const long long now_ = clock64 (); prof_t [k] += now_ - prof_last; prof_last = now_; }
#else
#define ACQ_T(k)
#endif
#define ACQ_THREADS 256
__device__ __forceinline__ float acq_level_exact (float a, float ja) {          // ofdm-processor.cpp:168, operation by operation
	return __double2float_rn (__dadd_rn (__dmul_rn (0.00001, (double) ja), __dmul_rn (1 - 0.00001, (double) a)));
}
// wave / flags: the asynchronous form used by dabgpu_decode_multi.  sd and flags then lie in pinned host memory (read once at the
// start, written once at the end), only the streams with do_acquire == wave take part, and a stream's CTA publishes its result by
// writing flags [stream] = wave after a system-wide fence: the host picks up every stream the moment ITS search ends, not when the
// slowest search of the launch does.
__global__ void __launch_bounds__ (ACQ_THREADS) acquire_kernel (StreamDev *sd, OfdmTables T, int T_F, int T_null, int wave, volatile int *flags) {
	__shared__ __align__ (16) float s_e [ACQ_CHUNK], s_d [ACQ_CHUNK + 64], s_sl_raw [ACQ_CHUNK + 64 + 4], s_csb [ACQ_CHUNK + 64], s_ja [ACQ_CHUNK], s_t [ACQ_CHUNK + 64], s_ring [64];
	float *const s_sl = s_sl_raw + 3;                            // &s_sl [1 + 4 j] is 16-byte aligned: the chain records four steps per store
	__shared__ int s_first;
	StreamDev &S = sd [blockIdx. x];
	if (flags ? S. do_acquire != wave : !S. do_acquire) return;
	const SampleWin w = S. w;
	StreamCtl *ctl = &S. ctl;
	const int tid = threadIdx. x;
	const long long total = w. len0 + w. len1 < S. limit ? w. len0 + w. len1 : S. limit;
	// scalar state, identical in every thread
	int stage = 0, cnt = 0, counter = 0, idx = 0, done = 0;
	float sLevel = 0.f, cs = 0.f;
	long long pos = ctl -> pos, attempt_pos = pos;
	int lp = ctl -> lp, attempt_lp = lp;
	const int phi = ctl -> coarse + ctl -> fine;
	uint2 rawn [ACQ_CHUNK / ACQ_THREADS];                       // the next chunk's samples as loaded (converted when used)
#ifdef ACQ_PROF
	long long prof_t [8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = clock64 (), prof_chunks = 0;
#endif
	long long next_pos = -1;
	while (true) {
		int n;                                                       // samples until the stage can change by COUNT
		if (stage == 0) n = 20 * T. T_s - cnt;                       // :278-280
		else if (stage == 1) n = 50 - cnt;                           // :284-290
		else if (stage == 2) n = T_F + 1 - counter;                  // :314-315: the (T_F + 1)-th sample is still consumed
		else n = T_null + 51 - counter;                              // :336-337
		if (n > ACQ_CHUNK) n = ACQ_CHUNK;
		if (pos + n > total) { done = 2; break; }                    // out of data: rewind to the attempt start
		const int ph = stage < 2 ? 0 : mod_rate (phi);               // getSample (0) while looking for a signal at all (:279, 285)
		ACQ_T (0)
		{	// 1. per-sample values
			int l = mod_rate ((long long) lp - (long long) (tid + 1) * ph);
			const int step = mod_rate ((long long) ACQ_THREADS * ph);
			if (next_pos != pos)                                     // (normally the samples were requested while the previous chunk's chains ran)
				sy_raw_chunk<ACQ_CHUNK / ACQ_THREADS, ACQ_THREADS> (w, pos, pos + n, tid, rawn);
			float2 oh [ACQ_CHUNK / ACQ_THREADS], ol [ACQ_CHUNK / ACQ_THREADS];     // all table lookups first (independent loads), then the arithmetic
#pragma unroll
			for (int k = 0; k < ACQ_CHUNK / ACQ_THREADS; k ++) {
				oh [k] = __ldg (&T. osc_hi [l >> 11]); ol [k] = __ldg (&T. osc_lo [l & 2047]);
				l -= step; if (l < 0) l += DAB_INPUT_RATE;
			}
#pragma unroll
			for (int k = 0; k < ACQ_CHUNK / ACQ_THREADS; k ++) {
				const int i = tid + k * ACQ_THREADS;
				if (i < n) {
					const float2 v = cmul (sy_conv (w, rawn [k]), cmul (oh [k], ol [k]));
					const float ja = fabsf (v. x) + fabsf (v. y);        // jan_abs
					s_ja [i] = ja; s_t [i] = 0.00001f * ja;
					s_e [i] = stage == 3 ? hypotf (v. x, v. y) : ja;       // abs () in SyncOnEndNull (:329)
				}
			}
			ACQ_T (1)
			__syncthreads ();
			ACQ_T (2)
			if (stage >= 1)
				for (int i = tid; i < n; i += ACQ_THREADS)
					s_d [i] = stage == 1 ? s_e [i] : __fsub_rn (s_e [i], i >= 50 ? s_e [i - 50] : s_ring [(idx + i - 50) & 63]);
			__syncthreads ();
		}
		ACQ_T (3)
		// the samples of the NEXT chunk (if this one runs to its end) are requested now: their HBM latency passes behind the chains
		next_pos = pos + n;
		sy_raw_chunk<ACQ_CHUNK / ACQ_THREADS, ACQ_THREADS> (w, next_pos, total, tid, rawn);
		// 2. the two recurrences: s_sl [i] / s_csb [i] = value before sample i, [n] = value after the chunk.  Warp 0 runs the level
		// surrogate, warp 1 (concurrently) the window sum.  The inputs of 32 steps come in with eight 128-bit broadcast loads,
		// requested a block ahead, so only the dependent arithmetic is on the chain -- ONE FFMA / FADD a step; all lanes of the
		// warp compute the same chain and lane 0 records it (a predicated store per step, nothing waits for it).
		// Surrogate: a' = fma (a, c_hi, u), u = fma (a_before, c_lo, k |v|) with c_hi + c_lo = 1 - 1e-5 to 48 bits; u takes the level
		// of TWO steps earlier (it differs by 2e-5 a, times c_lo ~ 1e-8: invisible), which keeps it off the chain.  (Measured: the
		// chain costs ~7.5 cycles a step with two FFMAs or with one -- a fully prepared additive term -- so it is the dependent-issue
		// latency of this loop, not the instruction count, that bounds it.)
		const int lane = tid & 31;
		if (tid >= 32 && tid < 64) {
			float c = cs;
			if (stage >= 1) {
				float4 dn [8];
#pragma unroll
				for (int k = 0; k < 8; k ++) dn [k] = reinterpret_cast<const float4 *> (s_d) [k];
				for (int i0 = 0; i0 < n; i0 += 32) {
					float dc [32];
#pragma unroll
					for (int k = 0; k < 8; k ++) { dc [4 * k] = dn [k]. x; dc [4 * k + 1] = dn [k]. y; dc [4 * k + 2] = dn [k]. z; dc [4 * k + 3] = dn [k]. w; dn [k] = reinterpret_cast<const float4 *> (s_d + i0 + 32) [k]; }
#pragma unroll
					for (int k = 0; k < 32; k += 4) {                   // (one 128-bit record per four steps: a store per step cost more than the chain)
						const float c1 = __fadd_rn (c, dc [k]), c2 = __fadd_rn (c1, dc [k + 1]), c3 = __fadd_rn (c2, dc [k + 2]);
						if (lane == 0) *reinterpret_cast<float4 *> (&s_csb [i0 + k]) = make_float4 (c, c1, c2, c3);
						c = __fadd_rn (c3, dc [k + 3]);
					}
				}
				if ((n & 31) == 0 && lane == 0) s_csb [n] = c;       // (n inside a block: the record of sample n there is the value after n - 1 already)
			} else if (lane == 0) s_csb [n] = c;
		}
		ACQ_T (7)
		int from = 0;                                                // s_sl [from] is known to be exact
		if (tid == 0) s_sl [0] = sLevel;
		__syncwarp ();
		while (true) {
			if (tid < 32) {                                          // surrogate chain from `from` on
				const float c_hi = 0.99999f, c_lo = (float) ((1 - 0.00001) - (double) 0.99999f);
				int i0 = from;
				float a = s_sl [from], ap = a, app = a;                  // the level now, one and two steps ago
				for (; (i0 & 31) != 0 && i0 < n; i0 ++) {                // (after a repair: up to the next block boundary one by one)
					const float u = __fmaf_rn (app, c_lo, s_t [i0]); app = ap; ap = a; a = __fmaf_rn (a, c_hi, u);
					if (lane == 0) s_sl [i0 + 1] = a;
				}
				float4 tn [8];
				if (i0 < n) {                                        // (i0 is a multiple of 32 here)
#pragma unroll
					for (int k = 0; k < 8; k ++) tn [k] = reinterpret_cast<const float4 *> (s_t + i0) [k];
				}
				for (; i0 < n; i0 += 32) {
					float tc [32];
#pragma unroll
					for (int k = 0; k < 8; k ++) { tc [4 * k] = tn [k]. x; tc [4 * k + 1] = tn [k]. y; tc [4 * k + 2] = tn [k]. z; tc [4 * k + 3] = tn [k]. w; tn [k] = reinterpret_cast<const float4 *> (s_t + i0 + 32) [k]; }
#pragma unroll
					for (int k = 0; k < 32; k += 4) {
						float r [4];
#pragma unroll
						for (int j = 0; j < 4; j ++) { const float u = __fmaf_rn (app, c_lo, tc [k + j]); app = ap; ap = a; a = __fmaf_rn (a, c_hi, u); r [j] = a; }
						if (lane == 0) *reinterpret_cast<float4 *> (&s_sl [i0 + 1 + k]) = make_float4 (r [0], r [1], r [2], r [3]);
					}
				}
				if (lane == 0) s_first = n;
			}
			ACQ_T (4)
			__syncthreads ();
			int bad = n;                                             // 2b. first step whose recorded successor is not the exact one
#pragma unroll
			for (int k = ACQ_CHUNK / ACQ_THREADS - 1; k >= 0; k --) {    // (independent iterations, highest first: the lowest bad index wins)
				const int i = from + tid + k * ACQ_THREADS;
				if (i < n && acq_level_exact (s_sl [i], s_ja [i]) != s_sl [i + 1]) bad = i;
			}
			if (bad < n) atomicMin (&s_first, bad);
			__syncthreads ();
			const int first = s_first;
			if (first == n) break;
			if (tid == 0) s_sl [first + 1] = acq_level_exact (s_sl [first], s_ja [first]);    // repair, then re-run from there
			from = first + 1;
			__syncthreads ();                                        // (everybody has read s_first before thread 0 resets it)
		}
		ACQ_T (5)
		int used = n;
		if (stage >= 2) {                                            // 3. the threshold tests (:301, :323), in parallel
			if (tid == 0) s_first = n;
			__syncthreads ();
			int first = n;
#pragma unroll
			for (int k = ACQ_CHUNK / ACQ_THREADS - 1; k >= 0; k --) {
				const int i = tid + k * ACQ_THREADS;
				const double lhs = (double) (s_csb [i] / 50.0f), lv = (double) s_sl [i];
				const bool leave = stage == 2 ? !(lhs > 0.40 * lv) : !(lhs < 0.75 * lv);
				if (i < n && leave) first = i;
			}
			if (first < n) atomicMin (&s_first, first);
			__syncthreads ();
			used = s_first;
		}
		// commit the state after `used` samples
		sLevel = s_sl [used]; cs = s_csb [used];
		__syncthreads ();                                            // everybody has read the records before the ring / next chunk overwrite anything
		if (stage >= 1) {
			for (int i = tid; i < used; i += ACQ_THREADS) if (i >= used - 64) s_ring [(idx + i) & 63] = s_e [i];
			idx += used;
		}
		__syncthreads ();
		ACQ_T (6)
#ifdef ACQ_PROF
		prof_chunks ++;
#endif
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
#ifdef ACQ_PROF
	if (tid == 0) printf ("acquire: stream %d, %lld samples from %lld, done %d; %lld chunks; cycles per chunk: head %lld prep %lld sync %lld diff %lld prefetch+cs %lld chain %lld verify %lld tests+commit %lld\n", (int) blockIdx. x, pos - ctl -> pos, ctl -> pos, done, prof_chunks, prof_t [0] / prof_chunks, prof_t [1] / prof_chunks,
	                      prof_t [2] / prof_chunks, prof_t [3] / prof_chunks, prof_t [7] / prof_chunks, prof_t [4] / prof_chunks, prof_t [5] / prof_chunks, prof_t [6] / prof_chunks);
#endif
	if (tid == 0) {
		if (done == 1) { ctl -> synced = 1; ctl -> pos = pos; ctl -> lp = lp; ctl -> acq_done = 1; }
		else           { ctl -> synced = 0; ctl -> pos = attempt_pos; ctl -> lp = attempt_lp; ctl -> acq_done = 0; }
		if (flags) { __threadfence_system (); flags [blockIdx. x] = wave; }
	}
}

// ---------------------------------------------------------------------------------------------------
// predict kernel (one CTA per stream): frame c of the chunk is assumed to start T_F after frame c-1 with unchanged correctors.
// The stream's budget of chunk slots is clamped here by what is resident from the (possibly just acquired) position on.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (128) predict_kernel (StreamDev *sd, FrameIn *fin, int T_F, long long frame_need) {
	StreamDev &S = sd [blockIdx. x];
	const StreamCtl ctl = S. ctl;
	int n = 0;
	if (ctl. synced && S. budget > 0 && S. limit - ctl. pos >= frame_need) {
		const long long avail = (S. limit - ctl. pos - frame_need) / T_F + 1;
		n = avail < S. budget ? (int) avail : S. budget;
	}
	const int phi = ctl. coarse + ctl. fine;
	for (int c = threadIdx. x; c < S. budget; c += 128) {
		FrameIn f;
		f. P = ctl. pos + (long long) c * T_F;
		f. lp = mod_rate ((long long) ctl. lp - (long long) c * T_F % DAB_INPUT_RATE * mod_rate (phi));
		f. phiA = f. phiB = phi; f. active = c < n; f. stream = (int) blockIdx. x; f. slot = S. slot0 + c;
		fin [S. first + c] = f;
	}
	if (threadIdx. x == 0) {
		S. nframes = n;
		if (n == 0) { S. ctl. n_valid = 0; S. ctl. n_redo = 0; S. ctl. lost = 0; }
	}
}

// ---------------------------------------------------------------------------------------------------
// scan kernel (one CTA per stream, thread 0 runs the serial part): the scalar state machine of ofdmProcessor::run
// replayed over the chunk.
//   derive = 1: optimistic pass.  Walks ALL frames, replacing fin[c] by the inputs the replayed state asks
//     for and marking the frames whose inputs changed for recomputation.  Where the data symbols were mixed
//     with another frequency than the replay wants, the cyclic-prefix correlation is corrected by the exact
//     identity  FreqCorr(f') = FreqCorr(f) * exp (-j 2 pi (f' - f) T_u / 2048000).  Nothing is committed.
//     A frame whose replayed window would reach past the samples known to be resident (StreamDev::limit: the end of the
//     input, or of the pieces of a host-to-device copy that have arrived) is not recomputed: the chunk is cut there and
//     the frame is left for the next round.
//   derive = 0: verification pass.  Accepts frames only while the inputs they were actually computed from
//     equal the replayed state, commits the stream state and the per-frame records.
// ---------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
__global__ void __launch_bounds__ (SCAN_THREADS) scan_kernel (StreamDev *sd, FrameIn *fin_all, int cap, int groups, DabParams dp, const FrameOut *fo_all,
                             const float2 *fcpart_all, int derive) {
	extern __shared__ __align__ (16) unsigned char scan_raw [];
	StreamDev &SD = sd [blockIdx. x];
	int nframes = SD. nframes;
	if (nframes <= 0) return;
	struct { FrameIn *in; FrameOut *fo; float2 *fc; double *inc; } S;
	S. in = reinterpret_cast<FrameIn *> (scan_raw);
	S. inc = reinterpret_cast<double *> (S. in + cap);
	S. fc = reinterpret_cast<float2 *> (S. inc + cap);
	S. fo = reinterpret_cast<FrameOut *> (S. fc + cap);
	__shared__ int s_first;
	StreamCtl *ctl = &SD. ctl;
	FrameIn *fin = fin_all + SD. first;
	const FrameOut *fo = fo_all + SD. first;
	const float2 *fcpart = fcpart_all + (size_t) SD. first * MAX_GROUPS;
	dabgpu_frame_info *info = SD. info;
	const int slot0 = SD. slot0;
	const long long abs_base = SD. abs_base, limit = SD. limit;
	const long long frame_need = 2ll * dp. T_u + (long long) (dp. L - 1) * dp. T_s + dp. T_null;   // worst case from P
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
			if (derive && changed && s. pos + frame_need > limit) {      // its samples are not (known to be) there yet: next round
				for (int k = c; k < nframes; k ++) S. in [k]. active = 0;
				break;
			}
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
			if (n_redo == 0) { s. n_redo = 0; s. fault = ctl -> fault; *ctl = s; }    // nothing changes: this replay IS the verification
		} else { s. n_redo = ctl -> n_redo; s. fault = ctl -> fault; *ctl = s; }
	}
	__syncthreads ();
	if (derive) for (int c = lane; c < SD. nframes; c += SCAN_THREADS) fin [c] = S. in [c];
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static size_t scan_smem (int cap) { return (size_t) cap * (sizeof (FrameIn) + sizeof (double) + sizeof (float2) + sizeof (FrameOut)); }

int sync_init (dabgpu *h) {
	CUDA_TRY (h, cudaFuncSetAttribute (scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) scan_smem (1024)));
	return DABGPU_OK;
}

void acquire_launch (dabgpu *h, StreamDev *sd, int nstreams, cudaStream_t st, int wave, int *flags) {
	ProfScope prof (h, KC_ACQUIRE, st);
	acquire_kernel<<<nstreams, ACQ_THREADS, 0, st>>> (sd, h -> engine -> T, h -> p. T_F, h -> p. T_null, wave, flags);
	h -> launches ++;
}

void predict_launch (dabgpu *h, StreamDev *sd, int nstreams, const ChunkBufs &cb, cudaStream_t st) {
	ProfScope prof (h, KC_SCAN, st);
	const DabParams &p = h -> p;
	predict_kernel<<<nstreams, 128, 0, st>>> (sd, cb. fin, p. T_F, 2ll * p. T_u + (long long) (p. L - 1) * p. T_s + p. T_null);
	h -> launches ++;
}

void scan_launch (dabgpu *h, StreamDev *sd, int nstreams, int max_frames, int groups, const ChunkBufs &cb, int derive, cudaStream_t st) {
	ProfScope prof (h, KC_SCAN, st);
	int cap = 16;
	while (cap < max_frames) cap *= 2;
	scan_kernel<<<nstreams, SCAN_THREADS, scan_smem (cap), st>>> (sd, cb. fin, cap, groups, h -> p, cb. fo, cb. fcpart, derive);
	h -> launches ++;
}
