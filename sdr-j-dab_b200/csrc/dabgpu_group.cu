// dabgpu_group.cu -- several GPUs of one box behind the C ABI (BASELINE.json configs[3] and [4]):
// one process, one engine handle and one host thread per device.
//
//   * dabgpu_group_decode_multi: independent streams spread over the GPUs, no communication at all;
//   * dabgpu_group_decode: ONE long recording split into contiguous frame ranges.  The couplings between the shards are
//     the reference's sequential state -- the sync / AFC variables of ofdmProcessor::run (ofdm-processor.cpp:445-466), the
//     unconsumed samples, and the 15 CIFs of soft bits the time de-interleaver remembers (dab-concurrent.cpp:41-43,
//     162-175).  Two schemes, both exact:
//       - parallel: GPU 0 decodes a short lead-in (acquisition, AFC convergence); a locked receiver is predictable
//         (every frame T_F samples long, correctors at rest), so every shard starts from the closed-form predicted
//         state (dabgpu_host_state_predict), 16 CIFs early so that its de-interleaver is full of real data when its own
//         range begins (a fresh back end swallows exactly those 16 CIFs, dab-concurrent.cpp:172-175) -- no halo travels.
//         Afterwards every boundary is VERIFIED: the state a shard assumed at its first own frame must equal the true
//         final state of its left neighbour.  All boundaries agree -> the concatenation is the one-GPU output bit for bit;
//       - chain (the fallback when a boundary disagrees or the receiver is not locked, and on request): the shards run one
//         after the other and the whole stream state moves from GPU r to GPU r+1 DEVICE TO DEVICE (cudaMemcpyPeerAsync over
//         NVLink: sample tail + the 15-CIF byte-symbol halo, <= ~1.3 MB; the scalars go through the host).
// The Python mirror of the same logic over torch.distributed (one process per GPU) is sdr-j-dab_b200/parallel.py.
#include <thread>
#include <algorithm>
#include "dabgpu_engine.h"

struct dabgpu_group {
	std::vector<dabgpu *> h;
	std::string err;
	std::vector<dabgpu_subch> subch;
	dabgpu_config cfg {};
};

static int group_fail (dabgpu_group *g, int code, const std::string &msg) { if (g) g -> err = msg; return code; }

extern "C" int32_t dabgpu_device_count (void) {
	int n = 0;
	if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
	return n;
}

extern "C" const char *dabgpu_group_last_error (const dabgpu_group_t *g) { return g ? g -> err. c_str () : "null group"; }
extern "C" int32_t dabgpu_group_size (const dabgpu_group_t *g) { return g ? (int32_t) g -> h. size () : 0; }
extern "C" dabgpu_t *dabgpu_group_handle (dabgpu_group_t *g, int32_t i) { return g && i >= 0 && i < (int32_t) g -> h. size () ? g -> h [i] : nullptr; }

extern "C" void dabgpu_group_destroy (dabgpu_group_t *g) {
	if (!g) return;
	for (auto *h : g -> h) dabgpu_destroy (h);
	delete g;
}

extern "C" int dabgpu_group_create (const dabgpu_config *cfg, const int32_t *devices, int32_t ndev, dabgpu_group_t **out) {
	if (!cfg || !out || ndev <= 0 || ndev > 16) return DABGPU_ERR_ARG;
	*out = nullptr;
	dabgpu_group *g = new dabgpu_group ();
	g -> cfg = *cfg;
	for (int i = 0; i < ndev; i ++) {
		dabgpu_config c = *cfg;
		c. device = devices ? devices [i] : i;
		dabgpu *h = nullptr;
		const int rc = dabgpu_create (&c, &h);
		if (rc) { dabgpu_group_destroy (g); return rc; }         // (text: dabgpu_last_error (NULL))
		g -> h. push_back (h);
	}
	// peer access both ways where the hardware offers it (NVLink / NVSwitch): the chain hand-over then copies device to device
	for (int i = 0; i < ndev; i ++)
		for (int j = 0; j < ndev; j ++) {
			if (i == j || g -> h [i] -> device == g -> h [j] -> device) continue;
			int can = 0;
			cudaDeviceCanAccessPeer (&can, g -> h [i] -> device, g -> h [j] -> device);
			if (can) { cudaSetDevice (g -> h [i] -> device); cudaDeviceEnablePeerAccess (g -> h [j] -> device, 0); cudaGetLastError (); }
		}
	*out = g;
	return DABGPU_OK;
}

extern "C" int dabgpu_group_set_subchannels (dabgpu_group_t *g, const dabgpu_subch *sc, int32_t nsub) {
	if (!g || nsub < 0 || (nsub > 0 && !sc)) return DABGPU_ERR_ARG;
	for (auto *h : g -> h) {
		const int rc = dabgpu_set_subchannels (h, sc, nsub);
		if (rc) return group_fail (g, rc, h -> err);
	}
	g -> subch. assign (sc, sc + nsub);
	return DABGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// independent streams: stream i -> GPU i mod n, every GPU runs its share through one dabgpu_decode_multi call
// ---------------------------------------------------------------------------------------------------
extern "C" int dabgpu_group_decode_multi (dabgpu_group_t *g, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format) {
	if (!g || nstreams < 0 || (nstreams > 0 && !jobs)) return DABGPU_ERR_ARG;
	const int n = (int) g -> h. size ();
	std::vector<std::vector<dabgpu_stream_job>> share (n);
	for (int i = 0; i < nstreams; i ++) share [i % n]. push_back (jobs [i]);
	std::vector<int> rc (n, DABGPU_OK);
	std::vector<std::thread> th;
	for (int r = 0; r < n; r ++)
		th. emplace_back ([&, r] { if (!share [r]. empty ()) rc [r] = dabgpu_decode_multi (g -> h [r], share [r]. data (), (int32_t) share [r]. size (), sample_format); });
	for (auto &t : th) t. join ();
	for (int r = 0; r < n; r ++) if (rc [r]) return group_fail (g, rc [r], g -> h [r] -> err);
	return DABGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// the whole stream state of handle a -> handle b, device to device (the chain hand-over)
// ---------------------------------------------------------------------------------------------------
static int state_handover (dabgpu *a, dabgpu *b) {
	Engine *A = a -> engine, *B = b -> engine;
	if (A -> backends. size () != B -> backends. size ()) return dab_fail (b, DABGPU_ERR_STATE, "hand-over between handles with different sub-channel sets");
	CUDA_TRY (b, cudaSetDevice (b -> device));
	const size_t tail_bytes = (size_t) A -> tail_len * A -> sample_bytes ();
	if (tail_bytes) CUDA_TRY (b, B -> tail. ensure (tail_bytes));
	CUDA_TRY (b, B -> d_hist8. ensure ((size_t) 15 * CIF_BITS));
	// the source handle is idle (its decode call has returned); the copies run on the destination's stream
	if (tail_bytes) CUDA_TRY (b, cudaMemcpyPeerAsync (B -> tail. p, b -> device, A -> tail. p, a -> device, tail_bytes, b -> stream));
	if (A -> hist_init && A -> d_hist8. p)
		CUDA_TRY (b, cudaMemcpyPeerAsync (B -> d_hist8. p, b -> device, A -> d_hist8. p, a -> device, (size_t) 15 * CIF_BITS, b -> stream));
	else
		CUDA_TRY (b, cudaMemsetAsync (B -> d_hist8. p, 127, (size_t) 15 * CIF_BITS, b -> stream));
	CUDA_TRY (b, cudaStreamSynchronize (b -> stream));
	B -> hist_init = true;
	B -> ctl = A -> ctl; B -> ctl. fault = 0;
	B -> abs_base = A -> abs_base; B -> frames_total = A -> frames_total; B -> cifs_total = A -> cifs_total;
	B -> tail_len = A -> tail_len; B -> cf32 = A -> cf32; B -> chunk = A -> chunk;
	for (size_t i = 0; i < A -> backends. size (); i ++) dab_backend_set_cifs_seen (B -> backends [i], dab_backend_cifs_seen (A -> backends [i]));
	B -> needs_reset = false;
	return DABGPU_OK;
}

// a view of the caller's result buffers starting at frame f / MSC block blk [i] of sub-channel i
struct ResultView {
	dabgpu_result r {};
	std::vector<uint8_t *> msc;
	std::vector<int32_t> nblk;
	std::vector<long long> blk0;                            // MSC block index (per sub-channel) the view starts at
	std::vector<dabgpu_frame_info> info_tmp;                // the boundary check needs the per-frame records even if the caller does not
};
static void make_view (const dabgpu *h, const std::vector<dabgpu_subch> &subch, dabgpu_result *out, long long f, const std::vector<long long> &blk, long long max_frames, ResultView *v) {
	const DabParams &p = h -> p;
	v -> r = dabgpu_result {};
	v -> r. max_frames = (int32_t) std::max (0ll, std::min (max_frames, (long long) out -> max_frames - f));
	v -> blk0 = blk;
	if (out -> info) v -> r. info = out -> info + f;
	else { v -> info_tmp. resize ((size_t) std::max (1, v -> r. max_frames)); v -> r. info = v -> info_tmp. data (); }
	v -> r. soft = out -> soft ? out -> soft + (size_t) f * (p. L - 1) * 2 * p. K : nullptr;
	v -> r. fic_bits = out -> fic_bits ? out -> fic_bits + (size_t) f * p. ficGroups * 768 : nullptr;
	v -> r. fic_crc = out -> fic_crc ? out -> fic_crc + (size_t) f * p. ficGroups * 3 : nullptr;
	v -> msc. assign (subch. size (), nullptr);
	v -> nblk. assign (subch. size (), 0);
	for (size_t i = 0; i < subch. size (); i ++)
		if (out -> msc_bits && out -> msc_bits [i]) v -> msc [i] = out -> msc_bits [i] + (size_t) blk [i] * 24 * subch [i]. bitRate;
	v -> r. msc_bits = out -> msc_bits ? v -> msc. data () : nullptr;
	v -> r. msc_nblocks = v -> nblk. data ();
}

extern "C" int dabgpu_group_decode (dabgpu_group_t *g, const uint8_t *iq_u8, size_t nsamples_, dabgpu_result *out, int32_t lead_frames, int32_t scheme, int32_t *scheme_used) {
	if (!g || !out || (nsamples_ > 0 && !iq_u8)) return DABGPU_ERR_ARG;
	const int n = (int) g -> h. size ();
	dabgpu *h0 = g -> h [0];
	const DabParams &p = h0 -> p;
	if (p. dabMode == 3) return group_fail (g, DABGPU_ERR_ARG, "stream decode is not available for Mode III");
	const long long nsamp = (long long) nsamples_, T_F = p. T_F, cpf = p. cifsPerFrame;
	const long long need = 2ll * p. T_u + (long long) (p. L - 1) * p. T_s + p. T_null;
	const long long overlap = (16 + cpf - 1) / cpf;               // frames that hold the 16 warm-up CIFs
	const size_t nsub = g -> subch. size ();
	if (lead_frames <= 0) lead_frames = 24;
	out -> nframes = 0; out -> consumed = 0;
	if (out -> msc_nblocks) for (size_t i = 0; i < nsub; i ++) out -> msc_nblocks [i] = 0;
	// fresh streams on every member
	for (auto *h : g -> h) {
		dabgpu_stream_state z {};
		z. f2Correction = 1; z. previous_1 = 1000; z. previous_2 = 999;
		dabgpu_state_set (h, &z);
		h -> engine -> hist_init = false;
		for (auto *b : h -> engine -> backends) dab_backend_set_cifs_seen (b, 0);
	}
	std::vector<long long> blk (nsub, 0);
	long long frames_done = 0;
	auto account = [&] (const ResultView &v) {
		frames_done += v. r. nframes;
		for (size_t i = 0; i < nsub; i ++) blk [i] += v. nblk [i];
	};
	// ---- 1. lead-in on GPU 0: acquisition + AFC convergence -> a locked state
	const long long lead_samples = std::min (nsamp, (long long) lead_frames * T_F);
	ResultView v0;
	make_view (h0, g -> subch, out, 0, blk, lead_frames + 2, &v0);
	int rc = dabgpu_decode (h0, iq_u8, (size_t) lead_samples, &v0. r);
	if (rc) return group_fail (g, rc, h0 -> err);
	account (v0);
	dabgpu_stream_state s0;
	dabgpu_state_get (h0, &s0);
	const bool locked = s0. synced == 1 && s0. f2Correction == 0;
	const long long total = nsamp - s0. abs_pos >= need ? (nsamp - s0. abs_pos - need) / T_F + 1 : 0;    // whole frames after the lead-in
	bool parallel = scheme != 0 && n > 1 && locked && total >= (long long) n * overlap;
	long long fed = lead_samples;                                  // samples handed to the chain so far
	if (parallel) {
		// ---- 2. frame ranges [a, b) counted from the first frame after the lead-in; every member decodes its range from the predicted state
		std::vector<long long> a (n), b (n);
		{	const long long base = total / n, extra = total % n;
			long long x = 0;
			for (int r = 0; r < n; r ++) { a [r] = x; x += base + (r < extra ? 1 : 0); b [r] = x; }
			if (b [0] < overlap) { b [0] = std::min (overlap, total); for (int r = 1; r < n; r ++) { a [r] = std::max (a [r], b [r - 1]); b [r] = std::max (b [r], a [r]); } }
		}
		std::vector<ResultView> v (n);
		std::vector<int> rcs (n, DABGPU_OK);
		std::vector<dabgpu_stream_state> fin (n);
		std::vector<long long> end_abs (n);
		// MSC blocks a member's range starts at: the lead-in + the frames before it, minus the 16 warm-up CIFs of the stream
		for (int r = 0; r < n; r ++) {
			std::vector<long long> blk_r (nsub);
			const long long cifs_before = (frames_done + a [r]) * cpf;
			for (size_t i = 0; i < nsub; i ++) blk_r [i] = r == 0 ? blk [i] : std::max (0ll, cifs_before - 16);
			make_view (h0, g -> subch, out, frames_done + a [r], blk_r, b [r] - a [r], &v [r]);
			end_abs [r] = r == n - 1 ? nsamp : std::min (nsamp, s0. abs_pos + b [r] * T_F + need - T_F + 64);    // enough for frame b-1, not for frame b
		}
		std::vector<std::thread> th;
		for (int r = 0; r < n; r ++)
			th. emplace_back ([&, r] {
				dabgpu *h = g -> h [r];
				if (r == 0) {
					if (end_abs [0] > lead_samples) rcs [0] = dabgpu_decode (h, iq_u8 + 2 * lead_samples, (size_t) (end_abs [0] - lead_samples), &v [0]. r);
				} else {
					dabgpu_stream_state sp;
					rcs [r] = dabgpu_host_state_predict (p. dabMode, &s0, a [r] - overlap, &sp);
					if (!rcs [r]) {
						sp. frames = 0; sp. cifs = 0;
						dabgpu_state_set (h, &sp);
						h -> engine -> out_skip = (int) overlap;       // the overlap frames only fill the de-interleaver
						rcs [r] = dabgpu_decode (h, iq_u8 + 2 * sp. abs_pos, (size_t) (end_abs [r] - sp. abs_pos), &v [r]. r);
					}
				}
				if (!rcs [r]) dabgpu_state_get (h, &fin [r]);
			});
		for (auto &t : th) t. join ();
		for (int r = 0; r < n; r ++) if (rcs [r]) return group_fail (g, rcs [r], g -> h [r] -> err);
		// ---- 3. verify every boundary: the state member r assumed at its first own frame == the true final state of member r - 1
		bool ok = fin [0]. synced == 1 && fin [0]. f2Correction == 0 && v [0]. r. nframes == b [0] - a [0];
		for (int r = 1; r < n && ok; r ++) {
			if (a [r] == b [r]) { ok = v [r]. r. nframes == 0; continue; }
			const dabgpu_stream_state &L = fin [r - 1];
			ok = v [r]. r. nframes == b [r] - a [r] && fin [r]. synced == 1 && fin [r]. f2Correction == 0;
			if (ok) {
				const dabgpu_frame_info &fi = v [r]. r. info [0];
				ok = fi. pos == L. abs_pos && fi. coarse == L. coarse && fi. fine == L. fine && fi. phase0 == L. localPhase;
			}
		}
		if (getenv ("DABGPU_GROUP_DEBUG")) {
			fprintf (stderr, "group: lead frames %lld, s0 pos %lld coarse %d fine %d lp %d, total %lld, ok %d\n", frames_done, (long long) s0. abs_pos, s0. coarse, s0. fine, s0. localPhase, total, (int) ok);
			for (int r = 0; r < n; r ++)
				fprintf (stderr, "  member %d: [%lld, %lld) got %d frames, first pos %lld coarse %d fine %d lp %d | final pos %lld coarse %d fine %d lp %d synced %d f2 %d\n", r, a [r], b [r], v [r]. r. nframes,
				         v [r]. r. nframes ? (long long) v [r]. r. info [0]. pos : -1ll, v [r]. r. nframes ? v [r]. r. info [0]. coarse : 0, v [r]. r. nframes ? v [r]. r. info [0]. fine : 0, v [r]. r. nframes ? v [r]. r. info [0]. phase0 : 0,
				         (long long) fin [r]. abs_pos, fin [r]. coarse, fin [r]. fine, fin [r]. localPhase, fin [r]. synced, fin [r]. f2Correction);
		}
		if (ok) {
			for (int r = 0; r < n; r ++) {
				frames_done += v [r]. r. nframes;
				for (size_t i = 0; i < nsub; i ++) blk [i] = v [r]. blk0 [i] + v [r]. nblk [i];
			}
			out -> nframes = (int32_t) frames_done;
			out -> consumed = fin [n - 1]. abs_pos;
			if (out -> msc_nblocks) for (size_t i = 0; i < nsub; i ++) out -> msc_nblocks [i] = (int32_t) blk [i];
			if (scheme_used) *scheme_used = 1;
			return DABGPU_OK;
		}
		// a boundary disagrees (corrector still moving, sync lost inside a shard): member 0's own range is exact whatever the
		// others did -- keep it and run the rest as the exact chain
		account (v [0]);
		fed = std::max (end_abs [0], lead_samples);
	}
	// ---- chain: sample ranges one after the other, the stream state moves from GPU to GPU device to device
	{
		const long long rest = nsamp - fed;
		int prev = 0;
		for (int r = 0; r < n && fed < nsamp; r ++) {
			if (parallel && r == 0) continue;                      // member 0 has done its share already
			const int left = n - r;
			long long take = r == n - 1 ? nsamp - fed : std::min (nsamp - fed, std::max (T_F, ((nsamp - fed) / left) / T_F * T_F));
			(void) rest;
			dabgpu *h = g -> h [r];
			if (r != prev) { if ((rc = state_handover (g -> h [prev], h))) return group_fail (g, rc, h -> err); }
			ResultView v;
			make_view (h0, g -> subch, out, frames_done, blk, take / T_F + 3, &v);
			if ((rc = dabgpu_decode (h, iq_u8 + 2 * fed, (size_t) take, &v. r))) return group_fail (g, rc, h -> err);
			account (v);
			fed += take;
			prev = r;
		}
		dabgpu_stream_state fs;
		dabgpu_state_get (g -> h [prev], &fs);
		out -> consumed = fs. abs_pos;
	}
	out -> nframes = (int32_t) frames_done;
	if (out -> msc_nblocks) for (size_t i = 0; i < nsub; i ++) out -> msc_nblocks [i] = (int32_t) blk [i];
	if (scheme_used) *scheme_used = 0;
	return DABGPU_OK;
}
