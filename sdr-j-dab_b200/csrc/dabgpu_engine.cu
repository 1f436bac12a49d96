// dabgpu_engine.cu -- host side of the stream engine: ofdmProcessor::run (ofdm-processor.cpp:247-474) re-designed for a
// frame-parallel GPU, feeding the FIC and MSC decoders (fic-handler.cpp, msc-handler.cpp, dab-concurrent.cpp).
//
// The reference loop is sequential through three couplings: the frame position (findIndex of frame n fixes
// where frame n+1 is read), the NCO phase/frequency (coarse/fine correctors updated once per frame) and the
// 16-CIF time de-interleaver.  The engine speculates: a round decodes a chunk of frames in parallel assuming
// the tracking state stays what it is at the chunk start (start index T_g, correctors unchanged -- the steady
// state of a locked receiver), then a scan kernel replays the reference's scalar state machine over the per-frame
// results (findIndex, coarse correction, cyclic-prefix correlation), recomputes what the replay wants differently and
// accepts frames up to the first one whose inputs differ from the replayed truth; the next round restarts there with
// the true state.  Accepted frames are therefore computed from exactly the inputs the reference would have used.
// Acquisition (null-symbol search) is a sample-serial scan done once per (re)synchronisation.
//
// A round works on an ARRAY of streams (StreamDev): dabgpu_decode runs one stream per handle, dabgpu_decode_multi
// any number of independent streams in lockstep through the same kernels (one warp / CTA per stream in the
// sequential kernels, one CTA per frame in the parallel ones), so n acquisitions or AFC convergences cost the time
// of one.  Kernels: dabgpu_sync.cu (acquire, predict, scan), dabgpu_symbol.cu (front, symbol).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "dabgpu_engine.h"
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
static double now_ms () { return std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now (). time_since_epoch ()). count (); }

// ---------------------------------------------------------------------------------------------------
// handle life cycle, control members
// ---------------------------------------------------------------------------------------------------
static void ctl_fresh (StreamCtl *c) {                      // ofdm-processor.cpp:258-259, 73
	memset (c, 0, sizeof (*c));
	c -> f2 = 1; c -> prev1 = 1000; c -> prev2 = 999;
}

int dab_engine_init (dabgpu *h) {
	Engine *E = new Engine ();
	h -> engine = E;
	int rc = ofdm_tables_init (h, &E -> T);
	if (rc) return rc;
	CUDA_TRY (h, cudaMalloc ((void **) &E -> d_phaseRef, (size_t) h -> p. T_u * sizeof (float2)));
	ctl_fresh (&E -> ctl);
	E -> groups = h -> p. L > 100 ? 8 : 5;
	if (const char *g = getenv ("DABGPU_GROUPS")) { const int v = atoi (g); if (v >= 1 && v <= MAX_GROUPS) { E -> groups = v; E -> groups_fixed = true; } }   // tuning knob (A/B runs)
	CUDA_TRY (h, E -> d_figkeys. ensure (128 * sizeof (unsigned long long)));
	CUDA_TRY (h, cudaMemsetAsync (E -> d_figkeys. p, 0, 128 * sizeof (unsigned long long), h -> stream));
	CUDA_TRY (h, cudaStreamCreateWithFlags (&E -> copy_st, cudaStreamNonBlocking));
	if ((rc = sync_init (h))) return rc;
	if ((rc = symbol_init (h))) return rc;
	return DABGPU_OK;
}

void dab_engine_free (dabgpu *h) {
	Engine *E = h -> engine;
	if (!E) return;
	for (auto *b : E -> backends) dabgpu_backend_destroy (b);
	if (E -> d_phaseRef) cudaFree (E -> d_phaseRef);
	if (E -> copy_st) { cudaStreamSynchronize (E -> copy_st); cudaStreamDestroy (E -> copy_st); }
	for (int k = 0; k < 4; k ++) if (E -> acq_st [k]) { cudaStreamSynchronize (E -> acq_st [k]); cudaStreamDestroy (E -> acq_st [k]); }
	for (auto &v : E -> copy_events) for (auto e : v) cudaEventDestroy (e);
	E -> tail. release (); E -> tail_spare. release (); E -> d_sd. release (); E -> h_sd. release ();
	E -> d_frameout. release (); E -> d_framein. release (); E -> d_fcpart. release (); E -> d_spec0. release (); E -> d_info. release ();
	E -> d_soft16. release (); E -> d_hist8. release (); E -> d_fic8. release (); E -> d_msc8. release ();
	E -> d_ficbits. release (); E -> d_ficcrc. release (); E -> d_figkeys. release ();
	for (auto &b : E -> d_mscbits) b. release ();
	E -> m_in. release (); E -> m_fic8. release (); E -> m_msc8. release (); E -> m_info. release ();
	E -> m_ficbits. release (); E -> m_ficcrc. release (); E -> m_mscbits. release ();
	E -> mh_in. release (); E -> mh_out. release (); E -> mh_acq. release (); E -> m_offs. release ();
	for (int i = 0; i < 2; i ++) { E -> d_inbuf [i]. release (); E -> h_stage [i]. release (); }
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
	E -> ctl. lp = s -> localPhase; E -> ctl. fault = 0;
	E -> tail_len = 0; E -> ctl. pos = 0; E -> abs_base = s -> abs_pos;      // the next input starts at abs_pos
	E -> frames_total = s -> frames; E -> cifs_total = s -> cifs;
	E -> needs_reset = false;
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

// the time de-interleaver's 15 history rows (dab-concurrent.cpp:70-74: zero soft bits = erasures = byte symbol 127)
static int ensure_history (dabgpu *h) {
	Engine *E = h -> engine;
	if (E -> hist_init && E -> d_hist8. p) return DABGPU_OK;
	CUDA_TRY (h, E -> d_hist8. ensure ((size_t) 15 * CIF_BITS));
	CUDA_TRY (h, cudaMemsetAsync (E -> d_hist8. p, 127, (size_t) 15 * CIF_BITS, h -> stream));
	E -> hist_init = true;
	return DABGPU_OK;
}

// ---- whole stream state as one blob (the multi-GPU hand-over: sync/AFC state + unconsumed samples + the
// 15-CIF soft-bit halo of the time de-interleaver + per-sub-channel warm-up counters) ----
struct StateBlobHeader {
	uint32_t magic; int32_t mode, nsub, hist_valid, cf32, pad;
	StreamCtl ctl; long long abs_base, frames_total, cifs_total, tail_len;
};
#define STATE_MAGIC 0x44414238u                              // "DAB8": the halo travels as byte symbols

extern "C" int dabgpu_state_export (dabgpu_t *h, void *buf, size_t capacity, size_t *used) {
	if (!h || !used) return DABGPU_ERR_ARG;
	Engine *E = h -> engine;
	const size_t rowb = (size_t) CIF_BITS;
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
	if (E -> hist_init) CUDA_TRY (h, cudaMemcpyAsync (q, E -> d_hist8. p, 15 * rowb, cudaMemcpyDeviceToHost, h -> stream));
	else memset (q, 127, 15 * rowb);
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	return DABGPU_OK;
}

extern "C" int dabgpu_state_import (dabgpu_t *h, const void *buf, size_t n) {
	if (!h || !buf || n < sizeof (StateBlobHeader)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_state_import: bad argument");
	Engine *E = h -> engine;
	StateBlobHeader hd;
	memcpy (&hd, buf, sizeof (hd));
	const size_t rowb = (size_t) CIF_BITS;
	if (hd. magic != STATE_MAGIC || hd. mode != h -> p. dabMode || hd. nsub != (int32_t) E -> backends. size () || hd. tail_len < 0 || hd. cf32 < 0 || hd. cf32 > 2 ||
	    n != sizeof (hd) + (size_t) hd. nsub * sizeof (int64_t) + (size_t) hd. tail_len * dab_sample_bytes (hd. cf32) + 15 * rowb)
		return dab_fail (h, DABGPU_ERR_ARG, "state blob does not match this handle (mode / sub-channel count / size)");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const char *q = (const char *) buf + sizeof (hd);
	for (auto *b : E -> backends) { int64_t c; memcpy (&c, q, sizeof (c)); q += sizeof (c); dab_backend_set_cifs_seen (b, c); }
	E -> ctl = hd. ctl; E -> ctl. fault = 0;
	E -> abs_base = hd. abs_base; E -> frames_total = hd. frames_total; E -> cifs_total = hd. cifs_total;
	E -> tail_len = hd. tail_len;
	E -> cf32 = hd. cf32;
	if (hd. tail_len) {
		CUDA_TRY (h, E -> tail. ensure ((size_t) hd. tail_len * E -> sample_bytes ()));
		CUDA_TRY (h, cudaMemcpyAsync (E -> tail. p, q, (size_t) hd. tail_len * E -> sample_bytes (), cudaMemcpyHostToDevice, h -> stream));
	}
	q += (size_t) hd. tail_len * E -> sample_bytes ();
	CUDA_TRY (h, E -> d_hist8. ensure (15 * rowb));
	CUDA_TRY (h, cudaMemcpyAsync (E -> d_hist8. p, q, 15 * rowb, cudaMemcpyHostToDevice, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	E -> hist_init = true;
	E -> needs_reset = false;
	return DABGPU_OK;
}

static int ensure_frame_capacity (dabgpu *h, long long frames) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	if (frames < 1) frames = 1;
	if (frames <= E -> cap_frames) return DABGPU_OK;
	const long long cap = frames + frames / 8 + 4;
	CUDA_TRY (h, E -> d_fic8. ensure ((size_t) cap * 3 * 2 * p. K + 16));
	CUDA_TRY (h, E -> d_msc8. ensure ((15 + (size_t) cap * p. cifsPerFrame) * CIF_BITS + 16));
	CUDA_TRY (h, E -> d_info. ensure ((size_t) cap * sizeof (dabgpu_frame_info)));
	E -> cap_frames = cap;
	return DABGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// one round of the OFDM part over an array of streams
// ---------------------------------------------------------------------------------------------------
static int ensure_round_bufs (dabgpu *h, int nstreams, int nslots, ChunkBufs *cb) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	if (nslots < 1) nslots = 1;
	CUDA_TRY (h, E -> d_frameout. ensure ((size_t) nslots * sizeof (FrameOut)));
	CUDA_TRY (h, E -> d_framein. ensure ((size_t) nslots * sizeof (FrameIn)));
	CUDA_TRY (h, E -> d_fcpart. ensure ((size_t) nslots * MAX_GROUPS * sizeof (float2)));
	CUDA_TRY (h, E -> d_spec0. ensure ((size_t) nslots * p. T_u * sizeof (float2)));
	CUDA_TRY (h, E -> d_sd. ensure ((size_t) nstreams * sizeof (StreamDev)));
	CUDA_TRY (h, E -> h_sd. ensure ((size_t) nstreams * sizeof (StreamDev)));
	cb -> fin = (FrameIn *) E -> d_framein. p; cb -> fo = (FrameOut *) E -> d_frameout. p;
	cb -> spec0 = (float2 *) E -> d_spec0. p; cb -> fcpart = (float2 *) E -> d_fcpart. p;
	return DABGPU_OK;
}

// Runs the streams in the pinned table h_sd through one round on the handle's main stream and brings the table back:
// [acquire] -> predict -> front, symbol, scan (derive) -> front, symbol, scan (verify).  On return every StreamDev::ctl holds
// the committed state (ctl.n_valid frames accepted from the stream's chunk, ctl.n_redo of them after a recomputation) and
// StreamDev::nframes the frames the round actually attempted.
static int run_round (dabgpu *h, int nstreams, int nslots, int max_budget, bool any_acquire, int fmt, const ChunkBufs &cb, const std::function<int ()> *while_running = nullptr) {
	Engine *E = h -> engine;
	cudaStream_t st = h -> stream;
	StreamDev *hsd = (StreamDev *) E -> h_sd. p, *dsd = (StreamDev *) E -> d_sd. p;
	static const bool trace = getenv ("DABGPU_TRACE") != nullptr && atoi (getenv ("DABGPU_TRACE")) >= 2;     // (level 2: with a synchronisation after every step)
	const double t0 = trace ? now_ms () : 0;
	if (trace) { cudaStreamSynchronize (st); fprintf (stderr, "[round] stream idle after %.3f ms\n", now_ms () - t0); }
	CUDA_TRY (h, cudaMemcpyAsync (dsd, hsd, (size_t) nstreams * sizeof (StreamDev), cudaMemcpyHostToDevice, st));
	if (any_acquire) acquire_launch (h, dsd, nstreams, st);
	if (trace) { cudaStreamSynchronize (st); fprintf (stderr, "[round] table up + acquire: at %.3f ms\n", now_ms () - t0); }
	predict_launch (h, dsd, nstreams, cb, st);
	// symbol groups (CTAs) per frame: every group first recomputes the spectrum of the symbol before its own as phase
	// reference, so few groups mean less redundant work and many groups more CTAs.  A big round fills the GPU anyway
	// (A/B on B200, 1024 frames: 5 groups 0.849 ms, 3 groups 0.828 ms, 1-2 groups the same)
	const int groups = !E -> groups_fixed && nslots >= 128 && E -> groups > 3 ? 3 : E -> groups;
	const int which = h -> cfg. reserved [0] ? 1 : 0;        // reserved[0] = 1: generic kernels (A/B testing)
	for (int pass = 0; pass < 2; pass ++) {
		// pass 0: speculative inputs, then the optimistic replay (derive); pass 1: recompute what changed, then verify
		front_launch (h, dsd, nslots, cb, which, st);
		symbol_launch (h, dsd, nslots, groups, cb, which, fmt, st);
		scan_launch (h, dsd, nstreams, max_budget, groups, cb, pass == 0 ? 1 : 0, st);
		if (trace) { cudaStreamSynchronize (st); fprintf (stderr, "[round] pass %d done at %.3f ms\n", pass, now_ms () - t0); }
	}
	CUDA_TRY (h, cudaGetLastError ());
	CUDA_TRY (h, cudaMemcpyAsync (hsd, dsd, (size_t) nstreams * sizeof (StreamDev), cudaMemcpyDeviceToHost, st));
	if (while_running) { const int rc = (*while_running) (); if (rc) return rc; }      // host work that fits behind the round's kernels
	CUDA_TRY (h, cudaStreamSynchronize (st));
	for (int i = 0; i < nstreams; i ++)
		if (hsd [i]. ctl. fault)
			return dab_fail (h, DABGPU_ERR_CUDA, "symbol kernel gave up on stream %d (%s)", i,
			                 hsd [i]. ctl. fault & 2 ? "a bulk copy of raw samples never arrived" : "shared-memory layout contract violated");
	return DABGPU_OK;
}

// chunk policy: grow while the speculation holds outright; a chunk that needed recomputation (correctors still
// moving) or was cut short shrinks, so a converging loop costs little and a locked one runs in one pass
// eager (many streams in lockstep): a round costs its latency whatever it carries, so the chunk also grows when the derive pass
// predicted the whole chunk right and only the recomputation was needed
static int next_chunk (int chunk, int attempted, const StreamCtl &c, int cap, bool eager = false) {
	if (c. n_valid == attempted && eager) chunk *= 4;
	else if (c. n_valid == attempted && c. n_redo == 0) chunk *= 2;
	else if (c. n_valid == attempted) chunk = chunk > 16 ? chunk / 2 : (chunk < 8 ? chunk * 2 : chunk);
	else chunk = chunk > 2 ? chunk / 2 : 1;
	return chunk > cap ? cap : (chunk < 1 ? 1 : chunk);
}

// ---------------------------------------------------------------------------------------------------
// channel decoding (FIC + every configured sub-channel) of frames just accepted by the OFDM part
// ---------------------------------------------------------------------------------------------------
// bytes of one decoded MSC block in the output buffers: 24 * bitRate bits, one per byte (viterbi.cpp:240-241) or packed
static inline size_t msc_block_bytes (const Engine *E, const dabgpu_subch &sc) { return E -> msc_packed ? (size_t) 3 * sc. bitRate : (size_t) 24 * sc. bitRate; }

extern "C" int dabgpu_set_msc_output (dabgpu_t *h, int32_t packed) {
	if (!h) return DABGPU_ERR_ARG;
	h -> engine -> msc_packed = packed != 0;
	return DABGPU_OK;
}

struct ChanPart {                                           // frames [f0, f0 + nv) of one stream
	const uint8_t *fic8, *msc8;                             // the stream's soft-bit planes (frame slot 0 / history row 0)
	int f0, nv;
	int64_t cifs_before;                                    // CIFs of the stream decoded before frame f0 (warm-up, dab-concurrent.cpp:172-175)
	uint8_t *ficbits, *ficcrc;                              // device outputs of the stream: [groups][768], [groups][3]
	uint8_t **mscbits;                                      // [nsub] device outputs of the stream
	int *nblk;                                              // [nsub] blocks written so far per sub-channel (updated)
	dabgpu_result *out;                                     // host result buffers of the stream
	unsigned long long fib0;                                // global number of the stream's first FIB in this part (FIG scan order), ~0 = no scan
	int out_skip = 0;                                       // FIC results of frames below this slot are not delivered; frame f goes to output index f - out_skip
};

// queued on a side stream so that it overlaps the OFDM work of the next chunk; results are copied to the callers'
// buffers on the same stream.  All parts go through ONE launch pair of the throughput kernels when that path is taken.
// defer_crc: no CRC launch and no delivery of CRC flags here (dabgpu_decode_multi checks every group of the call in one launch at
// its end, so that its channel-decoding batches are independent of each other and may run on different side streams)
// later: the device-to-host result copies of the parts are not issued here but appended to this list -- with hundreds of them per
// batch (streams x sub-channels) issuing them costs the host half a millisecond, which dabgpu_decode_multi spends while a round's
// kernels are running instead of between two rounds
struct LaterCopy { void *dst; const void *src; size_t bytes; cudaStream_t st; };
static int channel_parts (dabgpu *h, std::vector<ChanPart> &parts, bool defer_crc = false, std::vector<LaterCopy> *later = nullptr) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	const size_t nsub = E -> backends. size ();
	long long ncw = 0, ncif_all = 0;
	for (auto &q : parts) { ncw += (long long) q. nv * p. ficGroups + (long long) q. nv * p. cifsPerFrame * (long long) nsub; ncif_all += (long long) q. nv * p. cifsPerFrame; }
	if (ncw == 0) return DABGPU_OK;
	h -> cur = 1 + (E -> vrr ++ % 3);
	cudaStream_t st = h -> vst ();
	int rc = DABGPU_OK;
	// the throughput kernels take every job of the chunk in ONE launch pair; the warp-cooperative kernel is launched once per
	// sub-channel, 0.3 ms each however few code words there are -- so with two or more sub-channels the throughput path wins
	// even for a single frame (measured: 9 sub-channels, 1-32 frames per call: 2.9 ms against 0.6 ms)
	const bool simd = dab_use_simd (h, ncw) || (h -> cfg. viterbi_path == 0 && (nsub >= 2 || parts. size () >= 2) && ncif_all > 0) ||
	                  (E -> msc_packed && ncif_all > 0 && nsub > 0);      // (packing is done by the throughput kernels' chain-back)
	// job order = CTA order: the long MSC code words first, the FIC's short ones (774 steps against up to 9000) last, so that
	// the hardware hands the cheap CTAs out as the LAST ones of every SM's share instead of letting them occupy slots from the start
	std::vector<VitSimdJob> jobs, fic_jobs;
	std::vector<std::vector<int>> n_here (parts. size (), std::vector<int> (nsub, 0));
	do {
		for (size_t k = 0; k < parts. size () && !rc; k ++) {
			ChanPart &q = parts [k];
			const int ngroups = q. nv * p. ficGroups, g0 = q. f0 * p. ficGroups, ncif = q. nv * p. cifsPerFrame, r0 = q. f0 * p. cifsPerFrame;
			if (ngroups > 0) {
				const uint8_t *soft8 = q. fic8 + (size_t) q. f0 * 3 * 2 * p. K;
				if (simd) {
					fic_jobs. emplace_back ();
					if ((rc = dab_fic_simd_job (h, nullptr, 2304, ngroups, q. ficbits + (size_t) g0 * 768, &fic_jobs. back ()))) break;
					fic_jobs. back (). sym8 = const_cast<uint8_t *> (soft8);      // written by the symbol kernel
					fic_jobs. back (). stride8 = 2304;
				} else if ((rc = dab_fic_decode_dev (h, nullptr, soft8, 2304, ngroups, q. ficbits + (size_t) g0 * 768, q. ficcrc + (size_t) g0 * 3))) break;
			}
			for (size_t i = 0; i < nsub && ncif > 0; i ++) {
				const dabgpu_subch &sc = E -> subch [i];
				VitSimdJob job;
				uint8_t *dst = q. mscbits [i] + (size_t) q. nblk [i] * msc_block_bytes (E, sc);
				if ((rc = dab_backend_run_dev (E -> backends [i], nullptr, q. msc8 + (size_t) sc. startAddr * 64, CIF_BITS, r0, ncif,
				                               dst, &n_here [k] [i], simd ? &job : nullptr, q. cifs_before))) break;
				if (simd && n_here [k] [i] > 0) {
					job. sym8 = const_cast<uint8_t *> (q. msc8) + (size_t) (job. first_row - 15) * CIF_BITS + (size_t) sc. startAddr * 64;
					job. stride8 = CIF_BITS;
					job. packed = E -> msc_packed ? 1 : 0;
					jobs. push_back (job);
				}
			}
		}
		if (rc) break;
		if (simd) {
			jobs. insert (jobs. end (), fic_jobs. begin (), fic_jobs. end ());
			if ((rc = dab_vit_simd_run (h, jobs))) break;
			if (!defer_crc) for (auto &q : parts) {
				const int ngroups = q. nv * p. ficGroups, g0 = q. f0 * p. ficGroups;
				if (ngroups <= 0) continue;
				cudaError_t e = fib_crc_launch (h, q. ficbits + (size_t) g0 * 768, 3 * ngroups, q. ficcrc + (size_t) g0 * 3);
				if (e != cudaSuccess) { rc = dab_fail (h, DABGPU_ERR_CUDA, "crc launch: %s", cudaGetErrorString (e)); break; }
			}
			if (rc) break;
		}
		cudaError_t e = cudaSuccess;
		for (size_t k = 0; k < parts. size () && e == cudaSuccess; k ++) {
			ChanPart &q = parts [k];
			const int ngroups = q. nv * p. ficGroups, g0 = q. f0 * p. ficGroups;
			uint8_t *fb = q. ficbits + (size_t) g0 * 768, *fc = q. ficcrc + (size_t) g0 * 3;
			if (ngroups > 0 && q. fib0 != ~0ull) {                  // FIG 0/1 of the FIBs just checked (fib-processor.cpp:278-347)
				e = fig01_launch (h, fb, fc, 3 * ngroups, q. fib0, (unsigned long long *) E -> d_figkeys. p, st);
				if (e != cudaSuccess) break;
			}
			const int c0 = q. f0 > q. out_skip ? q. f0 : q. out_skip, cg = (q. f0 + q. nv - c0) * p. ficGroups;      // frames [c0, f0 + nv) are delivered
			const size_t src_g = (size_t) c0 * p. ficGroups, dst_g = (size_t) (c0 - q. out_skip) * p. ficGroups;
			auto d2h = [&] (void *dst, const void *src, size_t bytes) -> cudaError_t {
				if (later) { later -> push_back (LaterCopy { dst, src, bytes, st }); return cudaSuccess; }
				return cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToHost, st);
			};
			if (cg > 0 && q. out -> fic_bits) e = d2h (q. out -> fic_bits + dst_g * 768, q. ficbits + src_g * 768, (size_t) cg * 768);
			if (e == cudaSuccess && cg > 0 && q. out -> fic_crc && !defer_crc) e = d2h (q. out -> fic_crc + dst_g * 3, q. ficcrc + src_g * 3, (size_t) cg * 3);
			for (size_t i = 0; i < nsub && e == cudaSuccess; i ++) {
				const size_t fbytes = msc_block_bytes (E, E -> subch [i]);
				if (q. out -> msc_bits && q. out -> msc_bits [i] && n_here [k] [i] > 0)
					e = d2h (q. out -> msc_bits [i] + (size_t) q. nblk [i] * fbytes, q. mscbits [i] + (size_t) q. nblk [i] * fbytes, (size_t) n_here [k] [i] * fbytes);
				q. nblk [i] += n_here [k] [i];
			}
		}
		if (e != cudaSuccess) rc = dab_fail (h, DABGPU_ERR_CUDA, "result copy: %s", cudaGetErrorString (e));
	} while (0);
	h -> cur = 0;
	return rc;
}

// ---------------------------------------------------------------------------------------------------
// decode core on a device-resident input segment.  `ready` (optional): events of the piecewise host-to-device copy
// of the input; piece k covers new-segment samples [k * piece, (k+1) * piece).
// ---------------------------------------------------------------------------------------------------
static int decode_core_inner (dabgpu *h, const void *d_new_v, long long nnew, dabgpu_result *out,
                              const std::vector<cudaEvent_t> *ready, long long piece, int vit_batch_frames) {
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	const uchar2 *d_new = (const uchar2 *) d_new_v;
	const size_t sb = E -> sample_bytes ();
	const SampleWin w { (const uchar2 *) E -> tail. p, E -> tail_len, d_new, nnew, E -> cf32 };
	const long long total = E -> tail_len + nnew;
	const long long frame_need = 2ll * p. T_u + (long long) (p. L - 1) * p. T_s + p. T_null;   // worst case from P
	const long long max_frames_possible = total / p. T_F + 2;
	const int skip = E -> out_skip;                          // (multi-GPU shards: leading overlap frames decoded for the de-interleaver only)
	E -> out_skip = 0;
	long long want = (long long) out -> max_frames + skip < max_frames_possible ? (long long) out -> max_frames + skip : max_frames_possible;
	if (want < 0) want = 0;
	int rc = ensure_frame_capacity (h, want);
	if (rc) return rc;
	if ((rc = ensure_history (h))) return rc;
	ChunkBufs cb;
	if ((rc = ensure_round_bufs (h, 1, E -> max_chunk, &cb))) return rc;
	// result buffers on the device for the whole call (side streams write into them chunk by chunk)
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	CUDA_TRY (h, E -> d_ficbits. ensure ((size_t) (want * p. ficGroups + 1) * 768));
	CUDA_TRY (h, E -> d_ficcrc. ensure ((size_t) (want * p. ficGroups + 1) * 3));
	std::vector<uint8_t *> mscbits (E -> backends. size ());
	for (size_t i = 0; i < E -> backends. size (); i ++) {
		CUDA_TRY (h, E -> d_mscbits [i]. ensure ((size_t) (want * p. cifsPerFrame + 1) * 24 * E -> subch [i]. bitRate));
		mscbits [i] = (uint8_t *) E -> d_mscbits [i]. p;
	}
	// the 15 history rows of the time de-interleaver in front of this call's CIF rows
	CUDA_TRY (h, cudaMemcpyAsync (E -> d_msc8. p, E -> d_hist8. p, (size_t) 15 * CIF_BITS, cudaMemcpyDeviceToDevice, h -> stream));
	std::vector<int> nblk (E -> backends. size (), 0);
	StreamDev *hsd = (StreamDev *) E -> h_sd. p;
	int nframes = 0, decoded_upto = 0;
	// host input arrives piecewise: smaller chunks let the first frames start before the last samples are up
	// device-resident input with a channel-decoding batch given (dabgpu_config.reserved[1]): chunks no larger than the batch, so
	// that the OFDM part of the next chunk overlaps the channel decoding of this one on the side streams
	const bool piecewise = ready && !ready -> empty ();
	const int dev_cap = !ready && vit_batch_frames < E -> max_chunk ? (vit_batch_frames > 16 ? vit_batch_frames : 16) : E -> max_chunk;
	const int chunk_cap = piecewise ? (E -> max_chunk < 128 ? E -> max_chunk : 128) : dev_cap;
	long long waited = -1;                                   // input pieces [0, waited] are known to have arrived on the main stream
	long long limit = piecewise ? E -> tail_len : total;     // window samples known to be resident
	auto need_input = [&] (long long upto_window_pos) -> cudaError_t {      // samples before this window position must be resident
		if (!piecewise) return cudaSuccess;
		long long rel = upto_window_pos - E -> tail_len;
		if (rel <= 0) return cudaSuccess;
		long long k = (rel - 1) / piece;
		if (k >= (long long) ready -> size ()) k = (long long) ready -> size () - 1;
		cudaError_t e = cudaSuccess;
		if (k > waited) { e = cudaStreamWaitEvent (h -> stream, (*ready) [k], 0); waited = k; }    // copies are in order on one stream
		limit = std::min (total, E -> tail_len + (waited + 1) * piece);
		return e;
	};
	const int64_t cifs_seen0 = E -> backends. empty () ? 0 : dab_backend_cifs_seen (E -> backends [0]);
	auto channel = [&] (int f0, int nv) -> int {
		if (nv <= 0) return DABGPU_OK;
		std::vector<ChanPart> parts (1);
		ChanPart &q = parts [0];
		q. fic8 = (const uint8_t *) E -> d_fic8. p; q. msc8 = (const uint8_t *) E -> d_msc8. p;
		q. f0 = f0; q. nv = nv; q. cifs_before = cifs_seen0 + (int64_t) f0 * p. cifsPerFrame;
		q. ficbits = (uint8_t *) E -> d_ficbits. p; q. ficcrc = (uint8_t *) E -> d_ficcrc. p;
		q. mscbits = mscbits. data (); q. nblk = nblk. data (); q. out = out;
		q. fib0 = (unsigned long long) ((E -> frames_total + f0) * p. ficGroups) * 3ull;
		q. out_skip = skip;
		return channel_parts (h, parts);
	};
	while (nframes < want) {
		const bool acquiring = !E -> ctl. synced;
		long long C;
		if (acquiring) {
			CUDA_TRY (h, need_input (total));                // the null search reads until it finds one
			E -> chunk = 1; C = 1;
		} else {
			if (total - E -> ctl. pos < frame_need) break;
			const long long avail = (total - E -> ctl. pos - frame_need) / p. T_F + 1;
			C = E -> chunk < chunk_cap ? E -> chunk : chunk_cap;
			if (C > avail) C = avail;
			if (C > want - nframes) C = want - nframes;
			CUDA_TRY (h, need_input (E -> ctl. pos + (C - 1) * p. T_F + frame_need));
		}
		StreamDev &S = hsd [0];
		memset (&S, 0, sizeof (S));
		S. w = w; S. ctl = E -> ctl; S. ctl. fault = 0;
		S. fic8 = (uint8_t *) E -> d_fic8. p; S. msc8 = (uint8_t *) E -> d_msc8. p; S. info = (dabgpu_frame_info *) E -> d_info. p;
		S. abs_base = E -> abs_base; S. limit = limit;
		S. first = 0; S. budget = (int) C; S. nframes = 0; S. slot0 = nframes; S. do_acquire = acquiring;
		if ((rc = run_round (h, 1, (int) C, (int) C, acquiring, E -> cf32, cb))) return rc;
		E -> ctl = S. ctl;
		if (acquiring && !E -> ctl. synced) break;           // ran out of samples inside the attempt
		nframes += E -> ctl. n_valid;
		// the accepted frames' soft bits are final: once enough of them have piled up to fill the GPU, decode them on a
		// side stream while the OFDM part (and the input copy) of the following frames goes on
		if (nframes - decoded_upto >= vit_batch_frames) {
			if ((rc = channel (decoded_upto, nframes - decoded_upto))) return rc;
			decoded_upto = nframes;
		}
		if (S. nframes > 0) E -> chunk = next_chunk (E -> chunk, S. nframes, E -> ctl, chunk_cap);
		else if (!acquiring) break;                          // nothing attempted: not enough resident samples for another frame
	}
	const int ndeliver = nframes > skip ? nframes - skip : 0;
	out -> nframes = ndeliver;
	const int ncif = nframes * p. cifsPerFrame;
	if ((rc = channel (decoded_upto, nframes - decoded_upto))) return rc;
	for (auto *b : E -> backends) dab_backend_note_cifs (b, ncif);
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	for (size_t i = 0; i < E -> backends. size (); i ++)
		if (out -> msc_nblocks) out -> msc_nblocks [i] = nblk [i];
	if (nframes > 0) {
		if (out -> info && ndeliver > 0)
			CUDA_TRY (h, cudaMemcpyAsync (out -> info, (const dabgpu_frame_info *) E -> d_info. p + skip, (size_t) ndeliver * sizeof (dabgpu_frame_info), cudaMemcpyDeviceToHost, h -> stream));
		if (out -> soft && ndeliver > 0) {                   // the int16 form of process_ficBlock / process_mscBlock, on demand
			const size_t bytes = (size_t) ndeliver * (p. L - 1) * 2 * p. K * sizeof (int16_t);
			CUDA_TRY (h, E -> d_soft16. ensure (bytes));
			soft16_launch (h, (const uint8_t *) E -> d_fic8. p + (size_t) skip * 3 * 2 * p. K, (const uint8_t *) E -> d_msc8. p + (size_t) skip * p. cifsPerFrame * CIF_BITS,
			               (int16_t *) E -> d_soft16. p, ndeliver, h -> stream);
			CUDA_TRY (h, cudaMemcpyAsync (out -> soft, E -> d_soft16. p, bytes, cudaMemcpyDeviceToHost, h -> stream));
		}
		// time de-interleaver history for the next call: the last 15 CIF rows
		CUDA_TRY (h, cudaMemcpyAsync (E -> d_hist8. p, (const char *) E -> d_msc8. p + (size_t) ncif * CIF_BITS, (size_t) 15 * CIF_BITS, cudaMemcpyDeviceToDevice, h -> stream));
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

static int decode_core (dabgpu *h, const void *d_new, long long nnew, dabgpu_result *out,
                        const std::vector<cudaEvent_t> *ready = nullptr, long long piece = 0, int vit_batch_frames = 0x7fffffff) {
	Engine *E = h -> engine;
	if (h -> p. dabMode == 3)
		return dab_fail (h, DABGPU_ERR_ARG, "stream decode is not available for Mode III (the reference has no Mode III framing either)");
	if (E -> needs_reset)
		return dab_fail (h, DABGPU_ERR_STATE, "an earlier call on this handle failed half way: set the stream state (dabgpu_state_set / dabgpu_state_import) before decoding on");
	const int rc = decode_core_inner (h, d_new, nnew, out, ready, piece, vit_batch_frames);
	if (rc) {
		// a call that fails half way leaves positions, tail and counters inconsistent with each other: quiesce every stream
		// (nothing of this call may still be reading the caller's buffers) and refuse further decoding until the caller sets a state
		const std::string msg = h -> err;
		cudaStreamSynchronize (h -> stream);
		for (int i = 1; i < 4; i ++) cudaStreamSynchronize (h -> vctx [i]. st);
		if (E -> copy_st) cudaStreamSynchronize (E -> copy_st);
		cudaGetLastError ();
		E -> tail_len = 0; E -> ctl. pos = 0; E -> ctl. synced = 0; E -> ctl. fault = 0;
		E -> needs_reset = true;
		h -> cur = 0;
		h -> err = msg;
	}
	return rc;
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

// The input of a host call goes up in pieces on its own stream, into one of two device buffers; every chunk of frames waits only
// for the pieces it reads, so the host-to-device copy overlaps the decode of the frames already there.  dabgpu_prefetch starts
// the upload of the NEXT block while the current call is still busy with its last frames (the link never idles between calls).
static long long upload_piece_init () {                      // samples per piece: 16 M (32 MB of u8 IQ; measured 4 / 8 / 16 / 32 / 64 M: 8.08 / 7.86 / 7.75 / 7.77 / 7.80 ms per 1024-frame step); DABGPU_PIECE_MSAMPLES overrides (A/B runs)
	if (const char *e = getenv ("DABGPU_PIECE_MSAMPLES")) { const long long v = atoll (e); if (v >= 1 && v <= 1024) return v << 20; }
	return 16ll << 20;
}
static const long long UPLOAD_PIECE = upload_piece_init ();

static int issue_upload (dabgpu *h, const uint8_t *iq_u8, size_t nsamples, int fmt, int b, std::vector<cudaEvent_t> *ready) {
	Engine *E = h -> engine;
	const size_t sb = dab_sample_bytes (fmt), bytes = nsamples * sb;
	ready -> clear ();
	CUDA_TRY (h, E -> d_inbuf [b]. ensure (bytes + 16));
	if (!bytes) return DABGPU_OK;
	cudaPointerAttributes attr;
	const bool pinned = cudaPointerGetAttributes (&attr, iq_u8) == cudaSuccess && attr. type == cudaMemoryTypeHost;
	cudaGetLastError ();
	const uint8_t *src = iq_u8;
	if (!pinned) {                                           // pageable memory: through the handle's pinned staging buffer of this slot
		CUDA_TRY (h, E -> h_stage [b]. ensure (bytes));
		memcpy (E -> h_stage [b]. p, iq_u8, bytes);
		src = (const uint8_t *) E -> h_stage [b]. p;
	}
	const size_t npieces = (nsamples + UPLOAD_PIECE - 1) / UPLOAD_PIECE;
	std::vector<cudaEvent_t> &ev = E -> copy_events [b];
	while (ev. size () < npieces) {
		cudaEvent_t e;
		CUDA_TRY (h, cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
		ev. push_back (e);
	}
	for (size_t k = 0; k < npieces; k ++) {
		const size_t off = k * (size_t) UPLOAD_PIECE * sb, len = (k + 1 == npieces ? bytes - off : (size_t) UPLOAD_PIECE * sb);
		CUDA_TRY (h, cudaMemcpyAsync ((char *) E -> d_inbuf [b]. p + off, src + off, len, cudaMemcpyHostToDevice, E -> copy_st));
		CUDA_TRY (h, cudaEventRecord (ev [k], E -> copy_st));
		ready -> push_back (ev [k]);
	}
	return DABGPU_OK;
}

extern "C" int dabgpu_prefetch (dabgpu_t *h, const void *iq, size_t nsamples, int32_t sample_format) {
	if (!h || (nsamples > 0 && !iq) || sample_format < 0 || sample_format > 2) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_prefetch: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	Engine *E = h -> engine;
	int b = !E -> pending [0]. valid ? 0 : !E -> pending [1]. valid ? 1 : -1;      // no call is running now: a buffer without an announcement is free
	if (b < 0) return dab_fail (h, DABGPU_ERR_STATE, "dabgpu_prefetch: two blocks are announced already (decode one first)");
	int rc = issue_upload (h, (const uint8_t *) iq, nsamples, sample_format, b, &E -> pending [b]. ready);
	if (rc) { cudaStreamSynchronize (E -> copy_st); return rc; }
	Engine::Pending &P = E -> pending [b];
	P. host = iq; P. nsamples = nsamples; P. fmt = sample_format; P. seq = ++ E -> pending_seq; P. valid = true;
	return DABGPU_OK;
}

static int decode_host_inner (dabgpu *h, const uint8_t *iq_u8, size_t nsamples, dabgpu_result *out, cudaEvent_t *last) {
	Engine *E = h -> engine;
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	std::vector<cudaEvent_t> ready;
	int b = -1;
	for (int i = 0; i < 2; i ++) {                            // announced by dabgpu_prefetch (the older one if both match): already on its way
		const Engine::Pending &P = E -> pending [i];
		if (P. valid && P. host == (const void *) iq_u8 && P. nsamples == nsamples && P. fmt == E -> cf32 && (b < 0 || P. seq < E -> pending [b]. seq)) b = i;
	}
	if (b >= 0) { ready = E -> pending [b]. ready; E -> pending [b]. valid = false; }
	else {
		b = !E -> pending [0]. valid ? 0 : !E -> pending [1]. valid ? 1 : -1;
		if (b < 0) {                                         // both buffers hold announcements for other blocks: drop them
			cudaStreamSynchronize (E -> copy_st);
			E -> pending [0]. valid = E -> pending [1]. valid = false;
			b = 0;
		}
		int rc = issue_upload (h, iq_u8, nsamples, E -> cf32, b, &ready);
		if (rc) return rc;
	}
	if (!ready. empty ()) *last = ready. back ();
	// host input: the PCIe copy paces the call and the GPU idles most of the time, so channel decoding follows the
	// OFDM part in small batches: what is left to do once the last sample has arrived is then short
	return decode_core (h, E -> d_inbuf [b]. p, (long long) nsamples, out, &ready, UPLOAD_PIECE,
	                    h -> cfg. host_batch_frames > 0 ? h -> cfg. host_batch_frames : E -> vit_batch_frames);
}

static int decode_host (dabgpu *h, const void *iq_v, size_t nsamples, int cf32, dabgpu_result *out) {
	if (!h || !out || (nsamples > 0 && !iq_v)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	int rc = set_format (h, cf32);
	if (rc) return rc;
	cudaEvent_t last = nullptr;
	rc = decode_host_inner (h, (const uint8_t *) iq_v, nsamples, out, &last);
	// on every path the caller's buffer is free again when the call returns: this call's own last piece has landed (a block
	// announced by dabgpu_prefetch may still be on its way -- that is the point of it)
	if (rc) { cudaStreamSynchronize (h -> engine -> copy_st); h -> engine -> pending [0]. valid = h -> engine -> pending [1]. valid = false; }
	else if (last) cudaEventSynchronize (last);
	return rc;
}

// ---------------------------------------------------------------------------------------------------
// dabgpu_decode_multi: n independent streams, each decoded as a fresh handle with this handle's configuration and
// sub-channels would decode it in one call (acquisition from the first sample, coarse search on), all of them in
// lockstep through the same kernel launches.  The handle's own stream state is not touched.
// ---------------------------------------------------------------------------------------------------
#define MULTI_SLOT_CAP 4096                                  // chunk slots of one round, all streams together

// rows 0..14 of every stream's MSC plane = byte symbol 127 (dab-concurrent.cpp:70-74: zero soft bits = erasures); planes are 256-byte aligned
__global__ void __launch_bounds__ (256) hist_fill_kernel (uint8_t *base, const unsigned long long *offs) {
	uint4 *row = reinterpret_cast<uint4 *> (base + offs [blockIdx. y] + (size_t) blockIdx. x * CIF_BITS);
	const uint4 v = make_uint4 (0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
	for (int i = threadIdx. x; i < CIF_BITS / 16; i += 256) row [i] = v;
}

struct MultiStream {
	long long nsamples, in_off;                              // samples, byte offset of the stream's input in m_in
	long long want;                                          // frames wanted at most
	size_t fic_off, msc_off, info_off, ficbits_off, ficcrc_off;   // offsets of the stream's planes / outputs in the handle's multi buffers
	std::vector<size_t> mscbits_off;
	StreamCtl ctl;
	int chunk, nframes, decoded_upto;
	bool done;
	bool direct;                                             // every result buffer of the stream is pinned host memory: results are copied straight into them
	long long limit;                                         // samples known to be resident (host input goes up in pieces)
	long long acq_tried = -1;                                // an acquisition ran out of samples at this limit
	long long acq_limit = 0;                                 // limit the search in flight was started with
	int acq_wave = 0;                                        // launch number of the stream's search in flight (0 = none)
	std::vector<int> nblk;
	std::vector<uint8_t *> mscbits;
};

static int decode_multi_core (dabgpu *h, const dabgpu_stream_job *jobs, int nstreams, int fmt, bool dev_input) {
	const double t_entry = now_ms ();
	Engine *E = h -> engine;
	const DabParams &p = h -> p;
	if (p. dabMode == 3)
		return dab_fail (h, DABGPU_ERR_ARG, "stream decode is not available for Mode III (the reference has no Mode III framing either)");
	const size_t sb = dab_sample_bytes (fmt), nsub = E -> backends. size ();
	const long long frame_need = 2ll * p. T_u + (long long) (p. L - 1) * p. T_s + p. T_null;
	const size_t ficw = (size_t) 3 * 2 * p. K, cifw = (size_t) CIF_BITS;
	std::vector<MultiStream> ms (nstreams);
	size_t in_bytes = 0, fic_bytes = 0, msc_bytes = 0, info_n = 0, ngroups_all = 0, mscbits_bytes = 0;
	auto up = [] (size_t v, size_t a) { return (v + a - 1) / a * a; };
	for (int i = 0; i < nstreams; i ++) {
		const dabgpu_stream_job &J = jobs [i];
		if (!J. out || (J. nsamples > 0 && !J. iq)) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode_multi: stream %d: bad argument", i);
		MultiStream &m = ms [i];
		m. nsamples = (long long) J. nsamples;
		const long long possible = m. nsamples / p. T_F + 2;
		m. want = J. out -> max_frames < possible ? J. out -> max_frames : possible;
		if (m. want < 0) m. want = 0;
		const size_t cap = (size_t) (m. want < 1 ? 1 : m. want);
		m. in_off = (long long) in_bytes;           in_bytes += up ((size_t) m. nsamples * sb + 16, 256);
		m. fic_off = fic_bytes;                     fic_bytes += up (cap * ficw + 16, 256);
		m. msc_off = msc_bytes;                     msc_bytes += up ((15 + cap * p. cifsPerFrame) * cifw + 16, 256);
		m. info_off = info_n;                       info_n += cap;
		m. ficbits_off = ngroups_all * 768;         m. ficcrc_off = ngroups_all * 3;      // one group index space for all streams (one CRC launch)
		ngroups_all += cap * p. ficGroups;
		m. mscbits_off. resize (nsub);
		for (size_t s = 0; s < nsub; s ++) { m. mscbits_off [s] = mscbits_bytes; mscbits_bytes += up ((cap * p. cifsPerFrame + 1) * 24 * (size_t) E -> subch [s]. bitRate, 256); }
		ctl_fresh (&m. ctl);
		m. chunk = 1; m. nframes = 0; m. decoded_upto = 0; m. done = m. want == 0;
		m. nblk. assign (nsub, 0);
		J. out -> nframes = 0; J. out -> consumed = 0;
		// results go straight from the side stream into the caller's buffers when all of them are pinned (asynchronous copies that
		// overlap the following rounds); otherwise through the bulk staging copies at the end of the call
		auto is_pinned = [] (const void *q) { cudaPointerAttributes a; const bool ok = cudaPointerGetAttributes (&a, q) == cudaSuccess && a. type == cudaMemoryTypeHost; cudaGetLastError (); return ok; };
		m. direct = !J. out -> soft && (!J. out -> fic_bits || is_pinned (J. out -> fic_bits)) && (!J. out -> fic_crc || is_pinned (J. out -> fic_crc));
		for (size_t s = 0; s < nsub && m. direct && J. out -> msc_bits; s ++) m. direct = !J. out -> msc_bits [s] || is_pinned (J. out -> msc_bits [s]);
	}
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	if (!dev_input) CUDA_TRY (h, E -> m_in. ensure (in_bytes + 256));
	CUDA_TRY (h, E -> m_fic8. ensure (fic_bytes + 256));
	CUDA_TRY (h, E -> m_msc8. ensure (msc_bytes + 256));
	CUDA_TRY (h, E -> m_info. ensure ((info_n + 1) * sizeof (dabgpu_frame_info)));
	CUDA_TRY (h, E -> m_ficbits. ensure ((ngroups_all + 1) * 768));
	CUDA_TRY (h, E -> m_ficcrc. ensure ((ngroups_all + 1) * 3));
	if (ngroups_all > 0x7fffffff / 3) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode_multi: too many frames in one call");
	CUDA_TRY (h, E -> m_mscbits. ensure (mscbits_bytes + 256));
	ChunkBufs cb;
	int rc = ensure_round_bufs (h, nstreams, MULTI_SLOT_CAP, &cb);
	if (rc) return rc;
	// inputs: one copy per stream on the copy stream (pinned or pageable, as the caller has them); history rows = erasures
	// Host input goes up on the copy stream in two steps: first the head of EVERY stream (what a null search normally reads: 6.6
	// frames, so the searches start after a fraction of the upload), then the rest of the streams one after the other in GROUPS
	// (large copies: many small ones did not reach the link rate).  An event follows the heads and every group; the rounds follow
	// the arriving samples (StreamDev::limit; the scan cuts a chunk at the first frame whose window is not resident).
	static const int NGROUP = [] { const char *e = getenv ("DABGPU_MULTI_PIECES"); const int v = e ? atoi (e) : 0; return v >= 1 && v <= 64 ? v : 4; } ();     // (A/B knobs)
	static const double first_frames = [] { const char *e = getenv ("DABGPU_MULTI_FIRST"); const double v = e ? atof (e) : 0; return v > 0 ? v : 6.6; } ();
	const long long first_piece = ((long long) (first_frames * p. T_F) + 4095) / 4096 * 4096;       // (copies start on 8 KB boundaries of the stream)
	const int ngroups = nstreams < NGROUP ? nstreams : NGROUP;
	const int NPIECE = 1 + ngroups;                          // events: heads, then group 0 .. ngroups - 1
	auto group_of = [&] (int i) { return (int) ((long long) i * ngroups / nstreams); };
	auto piece_end = [&] (int i, int k) -> long long {       // samples of stream i resident once events 0 .. k have been seen
		const MultiStream &m = ms [i];
		if (m. nsamples <= first_piece || k >= 1 + group_of (i)) return m. nsamples;
		return first_piece;
	};
	std::vector<cudaEvent_t> &pev = E -> copy_events [0];
	int pieces_waited = dev_input ? NPIECE : 0;
	std::vector<const void *> d_in (nstreams);
	if (!dev_input) {
		while ((int) pev. size () < NPIECE) { cudaEvent_t e; CUDA_TRY (h, cudaEventCreateWithFlags (&e, cudaEventDisableTiming)); pev. push_back (e); }
		for (int k = 0; k < NPIECE; k ++) {
			for (int i = 0; i < nstreams; i ++) {
				const MultiStream &m = ms [i];
				const long long a = k ? piece_end (i, k - 1) : 0, b = piece_end (i, k);
				if (b > a)
					CUDA_TRY (h, cudaMemcpyAsync ((char *) E -> m_in. p + m. in_off + (size_t) a * sb, (const char *) jobs [i]. iq + (size_t) a * sb, (size_t) (b - a) * sb, cudaMemcpyHostToDevice, E -> copy_st));
			}
			CUDA_TRY (h, cudaEventRecord (pev [k], E -> copy_st));
		}
	}
	// events [0, upto) must have been passed before the main stream goes on (the host does not wait)
	auto wait_pieces = [&] (int upto) -> cudaError_t {
		cudaError_t e = cudaSuccess;
		if (upto > NPIECE) upto = NPIECE;
		if (upto > pieces_waited) { e = cudaStreamWaitEvent (h -> stream, pev [upto - 1], 0); pieces_waited = upto; }
		for (int i = 0; i < nstreams; i ++) ms [i]. limit = piece_end (i, pieces_waited - 1);
		return e;
	};
	for (int i = 0; i < nstreams; i ++) {
		MultiStream &m = ms [i];
		m. limit = dev_input ? m. nsamples : 0;
		d_in [i] = dev_input ? jobs [i]. iq : (const void *) ((const char *) E -> m_in. p + m. in_off);
		m. mscbits. resize (nsub);
		for (size_t s = 0; s < nsub; s ++) m. mscbits [s] = (uint8_t *) E -> m_mscbits. p + m. mscbits_off [s];
	}
	{	// the 15 history rows of every stream = erasures: ONE launch (a cudaMemsetAsync per stream cost 0.35 ms each on B200)
		CUDA_TRY (h, E -> mh_in. ensure ((size_t) nstreams * sizeof (unsigned long long)));
		CUDA_TRY (h, E -> m_offs. ensure ((size_t) nstreams * sizeof (unsigned long long)));
		unsigned long long *ho = (unsigned long long *) E -> mh_in. p;
		for (int i = 0; i < nstreams; i ++) ho [i] = (unsigned long long) ms [i]. msc_off;
		CUDA_TRY (h, cudaMemcpyAsync (E -> m_offs. p, ho, (size_t) nstreams * sizeof (unsigned long long), cudaMemcpyHostToDevice, h -> stream));
		hist_fill_kernel<<<dim3 (15, nstreams), 256, 0, h -> stream>>> ((uint8_t *) E -> m_msc8. p, (const unsigned long long *) E -> m_offs. p);
		h -> launches ++;
	}
	StreamDev *hsd = (StreamDev *) E -> h_sd. p;
	// channel decoding of everything accepted since the last call, all streams in one launch pair
	dabgpu_result no_copy {};
	std::vector<LaterCopy> later;
	auto channel_all = [&] (bool final) -> int {
		std::vector<ChanPart> parts;
		long long pending = 0;
		for (auto &m : ms) pending += m. nframes - m. decoded_upto;
		// a batch of the throughput Viterbi kernels takes the time of ONE code word's walk (1.4 ms for 3078 steps) however small it
		// is, as long as it stays below a wave: batches start early and overlap each other on the three side streams
		if (pending == 0 || (!final && pending < 128)) return DABGPU_OK;
		for (int i = 0; i < nstreams; i ++) {
			MultiStream &m = ms [i];
			if (m. nframes == m. decoded_upto) continue;
			ChanPart q;
			q. fic8 = (const uint8_t *) E -> m_fic8. p + m. fic_off; q. msc8 = (const uint8_t *) E -> m_msc8. p + m. msc_off;
			q. f0 = m. decoded_upto; q. nv = m. nframes - m. decoded_upto; q. cifs_before = (int64_t) m. decoded_upto * p. cifsPerFrame;
			q. ficbits = (uint8_t *) E -> m_ficbits. p + m. ficbits_off; q. ficcrc = (uint8_t *) E -> m_ficcrc. p + m. ficcrc_off;
			q. mscbits = m. mscbits. data (); q. nblk = m. nblk. data (); q. fib0 = ~0ull;
			q. out = m. direct ? jobs [i]. out : &no_copy;      // (pageable result buffers: in bulk at the end)
			parts. push_back (q);
			m. decoded_upto = m. nframes;
		}
		return channel_parts (h, parts, true, &later);      // (CRC flags: one launch over every group of the call at its end)
	};
	const std::function<int ()> flush_later = [&] () -> int {    // the result copies queued by the batches so far
		for (const LaterCopy &c : later)
			CUDA_TRY (h, cudaMemcpyAsync (c. dst, c. src, c. bytes, cudaMemcpyDeviceToHost, c. st));
		later. clear ();
		return DABGPU_OK;
	};
	const bool trace = getenv ("DABGPU_TRACE") != nullptr;
	const double t_call0 = now_ms ();
	if (trace) fprintf (stderr, "[multi] set-up took %.3f ms of host time\n", t_call0 - t_entry);
	int round_no = 0;
	// Acquisitions run ASYNCHRONOUSLY beside the rounds: the streams that need one are enrolled into a launch of the acquisition
	// kernel on a side stream; the kernel reads a stream's entry from pinned host memory and publishes the outcome there, stream
	// by stream, the moment that stream's search ends.  The rounds go on with whatever streams are in sync, so neither the slowest
	// null search nor a stream that lost sync holds the others up.
	CUDA_TRY (h, E -> mh_acq. ensure (up ((size_t) nstreams * sizeof (StreamDev), 64) + (size_t) nstreams * sizeof (int)));
	StreamDev *acq_sd = (StreamDev *) E -> mh_acq. p;
	volatile int *acq_flag = (volatile int *) ((char *) E -> mh_acq. p + up ((size_t) nstreams * sizeof (StreamDev), 64));
	for (int i = 0; i < nstreams; i ++) acq_flag [i] = 0;
	for (int k = 0; k < 4; k ++) if (!E -> acq_st [k]) CUDA_TRY (h, cudaStreamCreateWithFlags (&E -> acq_st [k], cudaStreamNonBlocking));
	int wave = 0, inflight = 0;
	auto collect = [&] () -> int {                           // outcomes published since the last look
		int got = 0;
		for (int i = 0; i < nstreams && inflight > 0; i ++) {
			MultiStream &m = ms [i];
			if (!m. acq_wave || acq_flag [i] != m. acq_wave) continue;
			std::atomic_thread_fence (std::memory_order_acquire);
			const StreamCtl &c = acq_sd [i]. ctl;
			m. ctl. synced = c. synced; m. ctl. pos = c. pos; m. ctl. lp = c. lp; m. ctl. acq_done = c. acq_done;
			m. acq_wave = 0; inflight --; got ++;
			if (m. ctl. synced) m. chunk = 4;
			else if (m. acq_limit < m. nsamples) m. acq_tried = m. acq_limit;      // out of the samples that had arrived: once more when there are more
			else m. done = true;                                                   // out of samples inside the attempt
		}
		return got;
	};
	while (true) {
		if (!dev_input) {                                        // pieces that have landed by now (at least the first): no waiting for more than that
			int k = pieces_waited;
			while (k < NPIECE && cudaEventQuery (pev [k]) == cudaSuccess) k ++;
			cudaGetLastError ();
			CUDA_TRY (h, wait_pieces (k < 1 ? 1 : k));
		}
		const bool all_in = pieces_waited >= NPIECE;
		collect ();
		int enrolled = 0;
		bool starved = false;                                    // some stream can do nothing until more of its samples have arrived
		for (int i = 0; i < nstreams; i ++) {
			MultiStream &m = ms [i];
			if (m. done || m. ctl. synced || m. acq_wave) continue;
			if (!all_in && m. limit <= m. acq_tried) { starved = true; continue; }     // (the search ran out of resident samples at this limit already)
			StreamDev &A = acq_sd [i];
			memset (&A, 0, sizeof (A));
			A. w = SampleWin { nullptr, 0, (const uchar2 *) d_in [i], m. nsamples, fmt };
			A. ctl = m. ctl; A. limit = m. limit; A. do_acquire = wave + 1;
			m. acq_wave = wave + 1; m. acq_limit = m. limit;
			enrolled ++;
		}
		if (enrolled) {
			wave ++;
			cudaStream_t ast = E -> acq_st [wave & 3];
			if (!dev_input) CUDA_TRY (h, cudaStreamWaitEvent (ast, pev [pieces_waited - 1], 0));
			acquire_launch (h, acq_sd, nstreams, ast, wave, (int *) acq_flag);
			CUDA_TRY (h, cudaGetLastError ());
			inflight += enrolled;
		}
		int nslots = 0, max_budget = 0;
		bool any = false;
		for (int i = 0; i < nstreams; i ++) {
			MultiStream &m = ms [i];
			StreamDev &S = hsd [i];
			memset (&S, 0, sizeof (S));
			S. w = SampleWin { nullptr, 0, (const uchar2 *) d_in [i], m. nsamples, fmt };
			S. ctl = m. ctl; S. ctl. fault = 0;
			S. fic8 = (uint8_t *) E -> m_fic8. p + m. fic_off; S. msc8 = (uint8_t *) E -> m_msc8. p + m. msc_off;
			S. info = (dabgpu_frame_info *) E -> m_info. p + m. info_off;
			S. abs_base = 0; S. limit = m. limit; S. first = nslots; S. slot0 = m. nframes;
			int C = 0;
			if (!m. done && m. ctl. synced && !m. acq_wave) {
				bool wait_data = false;
				if (m. limit - m. ctl. pos >= frame_need) {
					const long long avail = (m. limit - m. ctl. pos - frame_need) / p. T_F + 1;
					C = m. chunk < E -> max_chunk ? m. chunk : E -> max_chunk;
					if (C > avail) C = (int) avail;
					if (C > m. want - m. nframes) C = (int) (m. want - m. nframes);
				} else wait_data = m. limit < m. nsamples;
				if (C > MULTI_SLOT_CAP - nslots) C = MULTI_SLOT_CAP - nslots;    // (a stream left without slots waits a round)
				if (C <= 0 && !wait_data && !(m. nsamples - m. ctl. pos >= frame_need && m. nframes < m. want)) m. done = true;
				starved = starved || (C <= 0 && wait_data);
			}
			S. budget = C > 0 ? C : 0;
			nslots += S. budget;
			if (S. budget > max_budget) max_budget = S. budget;
			any = any || S. budget > 0;
		}
		if (!any) {
			if (inflight > 0) {                                  // nothing but searches going on: decode what is pending, then wait for the next outcome
				if ((rc = channel_all (true)) || (rc = flush_later ())) return rc;
				for (unsigned spin = 0; ; spin ++) {
					if (collect () > 0) break;
					std::this_thread::yield ();
					if ((spin & 0xfff) != 0xfff) continue;
					// no outcome for a while: are the searches still running?  (a launch that failed, or a fault, must not hang the caller)
					bool busy = false;
					for (int k = 0; k < 4; k ++) {
						const cudaError_t q = cudaStreamQuery (E -> acq_st [k]);
						if (q == cudaErrorNotReady) busy = true;
						else if (q != cudaSuccess) return dab_fail (h, DABGPU_ERR_CUDA, "acquisition kernel: %s", cudaGetErrorString (q));
					}
					cudaGetLastError ();
					if (busy) continue;
					if (collect () > 0) break;                       // (finished between the two looks)
					return dab_fail (h, DABGPU_ERR_CUDA, "acquisition kernel ended without publishing %d outcome(s)", inflight);
				}
				continue;
			}
			if (starved && !all_in) { CUDA_TRY (h, wait_pieces (pieces_waited + 1)); continue; }   // let the main stream wait for the next piece
			break;
		}
		const double t_r0 = trace ? now_ms () : 0;
		if ((rc = run_round (h, nstreams, nslots, max_budget, false, fmt, cb, &flush_later))) return rc;
		if (trace) {
			int acc = 0, redo = 0, att = 0;
			for (int i = 0; i < nstreams; i ++) { acc += hsd [i]. ctl. n_valid * (hsd [i]. budget > 0); redo += hsd [i]. ctl. n_redo * (hsd [i]. budget > 0); att += hsd [i]. nframes; }
			fprintf (stderr, "[multi] round %d: pieces %d slots %d max %d searches in flight %d attempted %d accepted %d redo %d  %.3f ms (at %.3f)\n", ++ round_no, pieces_waited, nslots, max_budget, inflight, att, acc, redo, now_ms () - t_r0, now_ms () - t_call0);
		}
		for (int i = 0; i < nstreams; i ++) {
			MultiStream &m = ms [i];
			const StreamDev &S = hsd [i];
			if (S. budget == 0) continue;
			m. ctl = S. ctl;                                     // (a failed findIndex leaves synced = 0: the stream is enrolled for a new search)
			m. nframes += m. ctl. n_valid;
			if (S. nframes > 0) m. chunk = next_chunk (m. chunk, S. nframes, m. ctl, E -> max_chunk, true);
			else if (S. limit >= m. nsamples) m. done = true;
			if (m. nframes >= m. want) m. done = true;
		}
		if ((rc = channel_all (false))) return rc;
	}
	if (trace) fprintf (stderr, "[multi] rounds done at %.3f ms\n", now_ms () - t_call0);
	if ((rc = channel_all (true)) || (rc = flush_later ())) return rc;
	for (int i = 1; i < 4; i ++) CUDA_TRY (h, cudaStreamSynchronize (h -> vctx [i]. st));
	if (trace) fprintf (stderr, "[multi] channel decoding done at %.3f ms\n", now_ms () - t_call0);
	if (ngroups_all > 0) {                                   // FIB CRCs of every group of every stream (groups never decoded give flags nobody reads)
		cudaError_t e = fib_crc_launch (h, (const uint8_t *) E -> m_ficbits. p, 3 * (int) ngroups_all, (uint8_t *) E -> m_ficcrc. p);
		if (e != cudaSuccess) return dab_fail (h, DABGPU_ERR_CUDA, "crc launch: %s", cudaGetErrorString (e));
	}
	// results: the streams' outputs lie side by side in four device buffers -- four device-to-host copies into pinned staging,
	// then plain host copies to the callers' buffers (hundreds of small copies to pageable memory cost more than the decode)
	const size_t b_fic = ngroups_all * 768, b_crc = ngroups_all * 3, b_info = info_n * sizeof (dabgpu_frame_info);
	const size_t o_crc = up (b_fic, 256), o_info = o_crc + up (b_crc, 256), o_msc = o_info + up (b_info, 256);
	CUDA_TRY (h, E -> mh_out. ensure (o_msc + mscbits_bytes + 256));
	char *ho = (char *) E -> mh_out. p;
	bool all_direct = true;
	for (auto &m : ms) all_direct = all_direct && m. direct;
	if (b_fic && !all_direct) CUDA_TRY (h, cudaMemcpyAsync (ho, E -> m_ficbits. p, b_fic, cudaMemcpyDeviceToHost, h -> stream));
	if (b_crc) CUDA_TRY (h, cudaMemcpyAsync (ho + o_crc, E -> m_ficcrc. p, b_crc, cudaMemcpyDeviceToHost, h -> stream));
	if (b_info) CUDA_TRY (h, cudaMemcpyAsync (ho + o_info, E -> m_info. p, b_info, cudaMemcpyDeviceToHost, h -> stream));
	if (mscbits_bytes && !all_direct) CUDA_TRY (h, cudaMemcpyAsync (ho + o_msc, E -> m_mscbits. p, mscbits_bytes, cudaMemcpyDeviceToHost, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	for (int i = 0; i < nstreams; i ++) {
		MultiStream &m = ms [i];
		dabgpu_result *out = jobs [i]. out;
		out -> nframes = m. nframes;
		out -> consumed = m. ctl. pos;
		const size_t ng = (size_t) m. nframes * p. ficGroups;
		if (out -> fic_bits && ng && !m. direct) memcpy (out -> fic_bits, ho + m. ficbits_off, ng * 768);
		if (out -> fic_crc && ng) memcpy (out -> fic_crc, ho + o_crc + m. ficcrc_off, ng * 3);
		if (out -> info && m. nframes) memcpy (out -> info, ho + o_info + m. info_off * sizeof (dabgpu_frame_info), (size_t) m. nframes * sizeof (dabgpu_frame_info));
		for (size_t s = 0; s < nsub; s ++) {
			if (out -> msc_nblocks) out -> msc_nblocks [s] = m. nblk [s];
			if (out -> msc_bits && out -> msc_bits [s] && m. nblk [s] > 0 && !m. direct)
				memcpy (out -> msc_bits [s], ho + o_msc + m. mscbits_off [s], (size_t) m. nblk [s] * msc_block_bytes (E, E -> subch [s]));
		}
		if (m. nframes > 0 && out -> soft) {
			const size_t bytes = (size_t) m. nframes * (p. L - 1) * 2 * p. K * sizeof (int16_t);
			CUDA_TRY (h, cudaStreamSynchronize (h -> stream));                  // d_soft16 is reused stream after stream
			CUDA_TRY (h, E -> d_soft16. ensure (bytes));
			soft16_launch (h, (const uint8_t *) E -> m_fic8. p + m. fic_off, (const uint8_t *) E -> m_msc8. p + m. msc_off, (int16_t *) E -> d_soft16. p, m. nframes, h -> stream);
			CUDA_TRY (h, cudaMemcpyAsync (out -> soft, E -> d_soft16. p, bytes, cudaMemcpyDeviceToHost, h -> stream));
		}
	}
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	if (trace) fprintf (stderr, "[multi] results delivered at %.3f ms\n", now_ms () - t_call0);
	return DABGPU_OK;
}

static int decode_multi (dabgpu *h, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t fmt, bool dev_input) {
	if (!h || nstreams < 0 || (nstreams > 0 && !jobs) || fmt < 0 || fmt > 2) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_decode_multi: bad argument");
	if (nstreams == 0) return DABGPU_OK;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const bool trace = getenv ("DABGPU_TRACE") != nullptr;
	const double t0 = now_ms ();
	const int rc = decode_multi_core (h, jobs, nstreams, fmt, dev_input);
	const double t1 = now_ms ();
	if (!dev_input) cudaStreamSynchronize (h -> engine -> copy_st);      // the callers' input buffers are free again on every path
	if (trace) fprintf (stderr, "[multi] call: core %.3f ms, copy stream drained after %.3f ms more\n", t1 - t0, now_ms () - t1);
	if (rc) {                                                // nothing of a failed call may still be running when it returns
		const std::string msg = h -> err;
		cudaStreamSynchronize (h -> stream);
		for (int k = 0; k < 4; k ++) if (h -> engine -> acq_st [k]) cudaStreamSynchronize (h -> engine -> acq_st [k]);
		for (int i = 1; i < 4; i ++) cudaStreamSynchronize (h -> vctx [i]. st);
		cudaGetLastError ();
		h -> cur = 0;
		h -> err = msg;
	}
	return rc;
}

extern "C" int dabgpu_decode_multi (dabgpu_t *h, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format) {
	return decode_multi (h, jobs, nstreams, sample_format, false);
}
extern "C" int dabgpu_decode_multi_dev (dabgpu_t *h, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format) {
	return decode_multi (h, jobs, nstreams, sample_format, true);
}
