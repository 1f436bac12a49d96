// dabgpu_api.cu -- C-ABI entry points of libdabgpu.so: handle management and the channel-decoding
// (Viterbi group) calls.  See include/dabgpu.h for the reference interfaces each one replaces.
#include <stdarg.h>
#include "dabgpu_internal.h"

thread_local std::string g_create_error;

int dab_fail (dabgpu *h, int code, const char *fmt, ...) {
	char buf [512];
	va_list ap;
	va_start (ap, fmt);
	vsnprintf (buf, sizeof (buf), fmt, ap);
	va_end (ap);
	if (h) h -> err = buf; else g_create_error = buf;
	return code;
}

extern "C" const char *dabgpu_last_error (const dabgpu_t *h) {
	return h ? h -> err. c_str () : g_create_error. c_str ();
}

extern "C" int dabgpu_create (const dabgpu_config *cfg, dabgpu_t **out) {
	if (!cfg || !out) return dab_fail (nullptr, DABGPU_ERR_ARG, "dabgpu_create: null argument");
	*out = nullptr;
	DabParams p;
	if (dab_mode_params (cfg -> dabMode, &p)) return dab_fail (nullptr, DABGPU_ERR_ARG, "dabMode %d not in 1..4", cfg -> dabMode);
	if (cfg -> freqSyncMethod < 0 || cfg -> freqSyncMethod > 2)
		return dab_fail (nullptr, DABGPU_ERR_ARG, "freqSyncMethod %d not in 0..2", cfg -> freqSyncMethod);
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount (&ndev);
	if (e != cudaSuccess || ndev <= 0)        // no CPU fallback: the engine refuses to exist without a GPU
		return dab_fail (nullptr, DABGPU_ERR_CUDA, "no usable CUDA device (%s)", cudaGetErrorString (e));
	if (cfg -> device < 0 || cfg -> device >= ndev)
		return dab_fail (nullptr, DABGPU_ERR_ARG, "device %d out of range (%d devices)", cfg -> device, ndev);
	if ((e = cudaSetDevice (cfg -> device)) != cudaSuccess)
		return dab_fail (nullptr, DABGPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString (e));
	dabgpu *h = new dabgpu ();
	h -> device = cfg -> device; h -> cfg = *cfg; h -> p = p;
	if (h -> cfg. threshold <= 0) h -> cfg. threshold = 3;
	if ((e = cudaStreamCreateWithFlags (&h -> stream, cudaStreamNonBlocking)) != cudaSuccess) {
		delete h;
		return dab_fail (nullptr, DABGPU_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString (e));
	}
	h -> vctx [0]. st = h -> stream;
	for (int i = 1; i < 4; i ++)
		if ((e = cudaStreamCreateWithFlags (&h -> vctx [i]. st, cudaStreamNonBlocking)) != cudaSuccess) {
			g_create_error = std::string ("cudaStreamCreate: ") + cudaGetErrorString (e);
			dabgpu_destroy (h);
			return DABGPU_ERR_CUDA;
		}
	int rc = dab_engine_init (h);
	if (rc) { g_create_error = h -> err; dabgpu_destroy (h); return rc; }
	*out = h;
	return DABGPU_OK;
}

extern "C" void dabgpu_destroy (dabgpu_t *h) {
	if (!h) return;
	cudaSetDevice (h -> device);
	if (h -> stream) cudaStreamSynchronize (h -> stream);
	for (int i = 1; i < 4; i ++) if (h -> vctx [i]. st) cudaStreamSynchronize (h -> vctx [i]. st);   // released buffers go to the process-wide cache: nothing may still use them
	dab_engine_free (h);
	for (auto &b : h -> d_table_bufs) b. release ();
	h -> d_in. release (); h -> d_out. release (); h -> d_aux. release ();
	for (int i = 0; i < 4; i ++) {
		dabgpu::VitCtx &c = h -> vctx [i];
		if (i > 0 && c. st) { cudaStreamSynchronize (c. st); cudaStreamDestroy (c. st); }
		c. d_dec. release (); c. d_jobs. release (); c. d_sym8. release (); c. h_jobs. release ();
	}
	h -> h_in. release (); h -> h_out. release ();
	if (h -> ev0) { cudaEventDestroy (h -> ev0); cudaEventDestroy (h -> ev1); }
	for (auto &pp : h -> prof_pending) { cudaEventDestroy (pp. a); cudaEventDestroy (pp. b); }
	for (auto e : h -> prof_pool) cudaEventDestroy (e);
	if (h -> stream) cudaStreamDestroy (h -> stream);
	delete h;
}

extern "C" int dabgpu_release_cached_memory (void) {
	BufCache &c = buf_cache ();
	std::vector<std::pair<int, void *>> dev; std::vector<void *> pin;
	{	std::lock_guard<std::mutex> g (c. m);
		for (int d = 0; d < 16; d ++) { for (auto &kv : c. dev [d]) dev. emplace_back (d, kv. second); c. dev [d]. clear (); c. dev_bytes [d] = 0; }
		for (auto &kv : c. pin) pin. push_back (kv. second);
		c. pin. clear (); c. pin_bytes = 0;
	}
	if (dev. empty () && pin. empty ()) return DABGPU_OK;
	int cur = 0;
	cudaGetDevice (&cur);
	for (auto &dp : dev) { cudaSetDevice (dp. first); cudaFree (dp. second); }
	for (void *p : pin) cudaFreeHost (p);
	cudaSetDevice (cur);
	return DABGPU_OK;
}

extern "C" int dabgpu_sync (dabgpu_t *h) {
	if (!h) return DABGPU_ERR_ARG;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	return DABGPU_OK;
}

extern "C" int dabgpu_timer_begin (dabgpu_t *h) {
	if (!h) return DABGPU_ERR_ARG;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	if (!h -> ev0) { CUDA_TRY (h, cudaEventCreate (&h -> ev0)); CUDA_TRY (h, cudaEventCreate (&h -> ev1)); }
	CUDA_TRY (h, cudaEventRecord (h -> ev0, h -> stream));
	return DABGPU_OK;
}

extern "C" int dabgpu_timer_end (dabgpu_t *h, float *ms) {
	if (!h || !ms || !h -> ev0) return dab_fail (h, DABGPU_ERR_STATE, "dabgpu_timer_end without dabgpu_timer_begin");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, cudaEventRecord (h -> ev1, h -> stream));
	CUDA_TRY (h, cudaEventSynchronize (h -> ev1));
	CUDA_TRY (h, cudaEventElapsedTime (ms, h -> ev0, h -> ev1));
	return DABGPU_OK;
}

ProfScope::ProfScope (dabgpu *h_, int cls_, cudaStream_t st_) : h (h_), cls (cls_), st (st_ ? st_ : h_ -> stream) {
	if (!h -> profiling) return;
	for (cudaEvent_t *e : { &a, &b }) {
		if (!h -> prof_pool. empty ()) { *e = h -> prof_pool. back (); h -> prof_pool. pop_back (); }
		else if (cudaEventCreate (e) != cudaSuccess) { *e = nullptr; }
	}
	if (a && b) cudaEventRecord (a, st);
}
ProfScope::~ProfScope () {
	if (!a || !b) return;
	cudaEventRecord (b, st);
	h -> prof_pending. push_back ({ cls, a, b });
}

extern "C" int dabgpu_profile_enable (dabgpu_t *h, int32_t on) {
	if (!h) return DABGPU_ERR_ARG;
	h -> profiling = on != 0;
	return DABGPU_OK;
}

extern "C" int dabgpu_profile_reset (dabgpu_t *h) {
	if (!h) return DABGPU_ERR_ARG;
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	for (auto &pp : h -> prof_pending) { h -> prof_pool. push_back (pp. a); h -> prof_pool. push_back (pp. b); }
	h -> prof_pending. clear ();
	for (int i = 0; i < KC_COUNT; i ++) { h -> prof_ms [i] = 0; h -> prof_n [i] = 0; }
	return DABGPU_OK;
}

extern "C" int dabgpu_profile_get (dabgpu_t *h, int32_t kernel_class, int64_t *launches, double *ms) {
	if (!h || kernel_class < 0 || kernel_class >= KC_COUNT || !launches || !ms) return DABGPU_ERR_ARG;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	for (auto &c : h -> vctx) if (c. st) CUDA_TRY (h, cudaStreamSynchronize (c. st));
	for (auto &pp : h -> prof_pending) {
		float t = 0;
		if (cudaEventElapsedTime (&t, pp. a, pp. b) == cudaSuccess) { h -> prof_ms [pp. cls] += t; h -> prof_n [pp. cls] ++; }
		h -> prof_pool. push_back (pp. a); h -> prof_pool. push_back (pp. b);
	}
	h -> prof_pending. clear ();
	*launches = h -> prof_n [kernel_class]; *ms = h -> prof_ms [kernel_class];
	return DABGPU_OK;
}

extern "C" int64_t dabgpu_launch_count (const dabgpu_t *h) { return h ? h -> launches : 0; }

int dab_device_table (dabgpu *h, long long key, const void *host, size_t bytes, void **dev) {
	auto it = h -> d_tables. find (key);
	if (it != h -> d_tables. end ()) { *dev = it -> second; return DABGPU_OK; }
	DevBuf b;
	CUDA_TRY (h, b. ensure (bytes));
	void *d = b. p;
	cudaError_t e = cudaMemcpyAsync (d, host, bytes, cudaMemcpyHostToDevice, h -> stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize (h -> stream);     // `host` may be a temporary
	if (e != cudaSuccess) { b. release (); return dab_fail (h, DABGPU_ERR_CUDA, "table upload: %s", cudaGetErrorString (e)); }
	h -> d_table_bufs. push_back (b);
	h -> d_tables [key] = d;
	*dev = d;
	return DABGPU_OK;
}

static long long profile_key (int kind, int bitRate, int uepFlag, int protLevel) {
	return ((long long) (kind == 3 ? 4 : kind ? 1 : 2) << 40) | ((long long) (uepFlag != 0) << 32) |
	       ((long long) (bitRate & 0xffff) << 16) | (protLevel & 0xffff);
}

int dab_get_profile (dabgpu *h, int kind, int bitRate, int uepFlag, int protLevel,
                     const ProtProfile **pp, const uint16_t **d_lut) {
	const long long key = profile_key (kind, bitRate, uepFlag, protLevel);
	auto it = h -> profiles. find (key);
	if (it == h -> profiles. end ()) {
		ProtProfile prof;
		const int rc = kind == 0 ? prot_build_fic (&prof) : kind == 3 ? prot_build_identity (bitRate, &prof) : prot_build_msc (bitRate, uepFlag, protLevel, &prof);
		if (rc) return dab_fail (h, DABGPU_ERR_PROFILE, "no %s profile for bitRate %d protLevel 0%o",
		                         uepFlag == 0 ? "UEP" : "EEP", bitRate, protLevel);
		it = h -> profiles. emplace (key, std::move (prof)). first;
	}
	*pp = &it -> second;
	void *d = nullptr;
	int rc = dab_device_table (h, key, it -> second. lut. data (), it -> second. lut. size () * sizeof (uint16_t), &d);
	if (rc) return rc;
	*d_lut = (const uint16_t *) d;
	return DABGPU_OK;
}

int dab_simd_job_profile (dabgpu *h, int kind, int bitRate, int uepFlag, int protLevel, int deint, VitSimdJob *job) {
	const ProtProfile *pp; const uint16_t *d_lut;
	int rc = dab_get_profile (h, kind, bitRate, uepFlag, protLevel, &pp, &d_lut);
	if (rc) return rc;
	const long long key = profile_key (kind, bitRate, uepFlag, protLevel);
	void *d = nullptr;
	const std::vector<uint32_t> &g = pp -> gather [deint ? 1 : 0];
	if ((rc = dab_device_table (h, key | ((deint ? 5ll : 7ll) << 44), g. data (), g. size () * sizeof (uint32_t), &d))) return rc;
	job -> gather = (const uint4 *) d;
	if ((rc = dab_device_table (h, key | (6ll << 44), pp -> chunk. data (), pp -> chunk. size () * sizeof (int32_t), &d))) return rc;
	job -> chunk = (const int2 *) d;
	job -> deint = deint ? 1 : 0; job -> ncols = pp -> nPunctured; job -> frameBits = pp -> frameBits; job -> nsteps = pp -> frameBits + 6;
	return DABGPU_OK;
}

// the one-code-word-per-thread kernels need ~20k code words in flight to fill 148 SMs, but already win from ~2k on; below that the
// warp-per-code-word kernel has the lower latency.  cfg.viterbi_path: 0 = auto, 1 = always warp, 2 = always SIMD.
bool dab_use_simd (const dabgpu *h, long long ncodewords) {
	if (h -> cfg. viterbi_path == 1) return false;
	if (h -> cfg. viterbi_path == 2) return true;
	return ncodewords >= 2048;
}

int dab_vit_simd_run (dabgpu *h, std::vector<VitSimdJob> &jobs) {
	if (jobs. empty ()) return DABGPU_OK;
	dabgpu::VitCtx &cx = h -> vctx [h -> cur];
	size_t dec_words = 0, sym_bytes = 0;
	int ctas = 0, ctas2 = 0;
	bool need_sym8 = false;
	const int cw2 = vit_simd_cw_per_cta (), cw1 = vit_simd_tb_cw_per_cta ();
	auto padded = [] (int nsteps) { return (size_t) vs_npad (nsteps); };
	for (auto &j : jobs) {
		j. cta_first = ctas; j. cta_first2 = ctas2; j. one = 1u;
		ctas += (j. ncw + cw1 - 1) / cw1; ctas2 += (j. ncw + cw2 - 1) / cw2;
		dec_words += padded (j. nsteps) * (size_t) ((j. ncw + 31) / 32 * 32);
		j. sym8_ready = j. sym8 != nullptr;                      // the caller already holds the byte symbols (stream engine)
		if (j. sym8_ready) continue;
		need_sym8 = true;
		j. stride8 = ((long long) j. ncols + 7 + 8) / 8 * 8;     // the staging may read up to 7 bytes past the last column
		sym_bytes += (size_t) (j. ncw + (j. deint ? 15 : 0)) * j. stride8;
	}
	// (re)allocation and the pinned job table are only touched once everything queued on this context is done
	const size_t jb = jobs. size () * sizeof (VitSimdJob);
	CUDA_TRY (h, cudaStreamSynchronize (cx. st));
	CUDA_TRY (h, cx. d_dec. ensure (dec_words * sizeof (uint2)));
	CUDA_TRY (h, cx. d_sym8. ensure (sym_bytes + 64));
	size_t off = 0, soff = 0;
	for (auto &j : jobs) {
		j. dec = (uint2 *) cx. d_dec. p + off; off += padded (j. nsteps) * (size_t) ((j. ncw + 31) / 32 * 32);
		if (!j. sym8_ready) { j. sym8 = (uint8_t *) cx. d_sym8. p + soff; soff += (size_t) (j. ncw + (j. deint ? 15 : 0)) * j. stride8; }
	}
	CUDA_TRY (h, cx. d_jobs. ensure (jb));
	CUDA_TRY (h, cx. h_jobs. ensure (jb));
	memcpy (cx. h_jobs. p, jobs. data (), jb);
	CUDA_TRY (h, cudaMemcpyAsync (cx. d_jobs. p, cx. h_jobs. p, jb, cudaMemcpyHostToDevice, cx. st));
	CUDA_TRY (h, vit_simd_launch (h, (const VitSimdJob *) cx. d_jobs. p, (int) jobs. size (), ctas, ctas2, need_sym8));
	return DABGPU_OK;
}

int dab_get_prbs (dabgpu *h, int nbits, const uint32_t **d_prbs) {
	const long long key = (3ll << 40) | nbits;
	auto it = h -> d_tables. find (key);
	if (it != h -> d_tables. end ()) { *d_prbs = (const uint32_t *) it -> second; return DABGPU_OK; }
	std::vector<uint32_t> w;
	prbs_packed (nbits, &w);
	void *d = nullptr;
	int rc = dab_device_table (h, key, w. data (), w. size () * 4, &d);
	if (rc) return rc;
	*d_prbs = (const uint32_t *) d;
	return DABGPU_OK;
}

// stage host input into the handle's device input buffer (through pinned memory)
static int stage_in (dabgpu *h, const void *host, size_t bytes, void **dev) {
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, h -> d_in. ensure (bytes));
	CUDA_TRY (h, h -> h_in. ensure (bytes));
	memcpy (h -> h_in. p, host, bytes);
	CUDA_TRY (h, cudaMemcpyAsync (h -> d_in. p, h -> h_in. p, bytes, cudaMemcpyHostToDevice, h -> stream));
	*dev = h -> d_in. p;
	return DABGPU_OK;
}
static int stage_out (dabgpu *h, void *host, const void *dev, size_t bytes) {
	CUDA_TRY (h, h -> h_out. ensure (bytes));
	CUDA_TRY (h, cudaMemcpyAsync (h -> h_out. p, dev, bytes, cudaMemcpyDeviceToHost, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	memcpy (host, h -> h_out. p, bytes);
	return DABGPU_OK;
}

// ---- viterbi::deconvolve x nblocks (viterbi.cpp:225-242) ----
extern "C" int dabgpu_viterbi_dev (dabgpu_t *h, const int16_t *soft, int32_t frameBits, int32_t nblocks, uint8_t *bits) {
	if (!h || !soft || !bits || frameBits <= 0 || frameBits > 32767 - 6 || nblocks < 0)
		return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_viterbi: bad argument");
	CUDA_TRY (h, cudaSetDevice (h -> device));
	VitJob j {};
	j. in = soft; j. in_stride = 4ll * (frameBits + 6); j. lut = nullptr;
	j. frameBits = frameBits; j. nsteps = frameBits + 6; j. nblocks = nblocks;
	j. deint = 0; j. prbs = nullptr; j. out = bits;
	if (dab_use_simd (h, nblocks) && 4 * (frameBits + 6) < 0xFFFF) {
		std::vector<VitSimdJob> jobs (1);
		VitSimdJob &s = jobs [0];
		memset (&s, 0, sizeof (s));
		int rc = dab_simd_job_profile (h, 3, frameBits, 1, 0, 0, &s);
		if (rc) return rc;
		s. in = soft; s. in_stride = j. in_stride; s. ncw = nblocks; s. out = bits;
		return dab_vit_simd_run (h, jobs);
	}
	CUDA_TRY (h, vit_launch (h, KC_VITERBI_API, j));
	return DABGPU_OK;
}

extern "C" int dabgpu_viterbi (dabgpu_t *h, const int16_t *soft, int32_t frameBits, int32_t nblocks, uint8_t *bits) {
	if (!h || !soft || !bits || frameBits <= 0 || frameBits > 32767 - 6 || nblocks < 0)
		return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_viterbi: bad argument");
	if (nblocks == 0) return DABGPU_OK;
	void *d_in;
	int rc = stage_in (h, soft, (size_t) nblocks * 4 * (frameBits + 6) * sizeof (int16_t), &d_in);
	if (rc) return rc;
	const size_t obytes = (size_t) nblocks * frameBits;
	CUDA_TRY (h, h -> d_out. ensure (obytes));
	rc = dabgpu_viterbi_dev (h, (const int16_t *) d_in, frameBits, nblocks, (uint8_t *) h -> d_out. p);
	if (rc) return rc;
	return stage_out (h, bits, h -> d_out. p, obytes);
}

// ---- eep_/uep_deconvolve::deconvolve x nblocks (deconvolve.cpp:186-237, 338-366) ----
extern "C" int dabgpu_protect_decode (dabgpu_t *h, int32_t bitRate, int32_t uepFlag, int32_t protLevel,
                                      const int16_t *v, int32_t size, int32_t nblocks, uint8_t *bits) {
	if (!h || !v || !bits || nblocks < 0 || size <= 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_protect_decode: bad argument");
	const ProtProfile *pp; const uint16_t *d_lut;
	int rc = dab_get_profile (h, 1, bitRate, uepFlag, protLevel, &pp, &d_lut);
	if (rc) return rc;
	if (size < pp -> nPunctured)
		return dab_fail (h, DABGPU_ERR_ARG, "size %d smaller than the %d punctured bits of the profile", size, pp -> nPunctured);
	if (nblocks == 0) return DABGPU_OK;
	void *d_in;
	rc = stage_in (h, v, (size_t) nblocks * size * sizeof (int16_t), &d_in);
	if (rc) return rc;
	const size_t obytes = (size_t) nblocks * pp -> frameBits;
	CUDA_TRY (h, h -> d_out. ensure (obytes));
	VitJob j {};
	j. in = (const int16_t *) d_in; j. in_stride = size; j. lut = d_lut;
	j. frameBits = pp -> frameBits; j. nsteps = pp -> frameBits + 6; j. nblocks = nblocks;
	j. out = (uint8_t *) h -> d_out. p;
	if (dab_use_simd (h, nblocks)) {
		std::vector<VitSimdJob> jobs (1);
		VitSimdJob &s = jobs [0];
		memset (&s, 0, sizeof (s));
		if ((rc = dab_simd_job_profile (h, 1, bitRate, uepFlag, protLevel, 0, &s))) return rc;
		s. in = j. in; s. in_stride = size; s. ncw = nblocks; s. out = j. out;
		if ((rc = dab_vit_simd_run (h, jobs))) return rc;
	} else
		CUDA_TRY (h, vit_launch (h, KC_VITERBI_API, j));
	return stage_out (h, bits, h -> d_out. p, obytes);
}

// ---- ficHandler::process_ficInput x ngroups (fic-handler.cpp:241-321) ----
int dab_fic_simd_job (dabgpu *h, const int16_t *d_soft, long long stride, int ngroups, uint8_t *d_bits, VitSimdJob *s) {
	memset (s, 0, sizeof (*s));
	int rc = dab_simd_job_profile (h, 0, 0, 1, 0, 0, s);
	if (rc) return rc;
	if ((rc = dab_get_prbs (h, 768, &s -> prbs))) return rc;
	s -> in = d_soft; s -> in_stride = stride; s -> ncw = ngroups; s -> out = d_bits;
	return DABGPU_OK;
}

int dab_fic_decode_dev (dabgpu *h, const int16_t *d_soft, const uint8_t *d_soft8, long long stride, int ngroups, uint8_t *d_bits, uint8_t *d_crc) {
	const ProtProfile *pp; const uint16_t *d_lut; const uint32_t *d_prbs;
	int rc = dab_get_profile (h, 0, 0, 1, 0, &pp, &d_lut);
	if (rc) return rc;
	if ((rc = dab_get_prbs (h, 768, &d_prbs))) return rc;
	VitJob j {};
	j. in = d_soft; j. in8 = d_soft8; j. in_stride = stride; j. lut = d_lut;
	j. frameBits = 768; j. nsteps = 774; j. nblocks = ngroups;
	j. prbs = d_prbs; j. out = d_bits;
	if (dab_use_simd (h, ngroups)) {
		std::vector<VitSimdJob> jobs (1);
		if ((rc = dab_fic_simd_job (h, d_soft, stride, ngroups, d_bits, &jobs [0]))) return rc;
		if (d_soft8) { jobs [0]. sym8 = const_cast<uint8_t *> (d_soft8); jobs [0]. stride8 = stride; }
		if ((rc = dab_vit_simd_run (h, jobs))) return rc;
	} else
		CUDA_TRY (h, vit_launch (h, KC_VITERBI_FIC, j));
	if (d_crc) CUDA_TRY (h, fib_crc_launch (h, d_bits, 3 * ngroups, d_crc));
	return DABGPU_OK;
}

extern "C" int dabgpu_fic_decode (dabgpu_t *h, const int16_t *soft, int32_t ngroups, uint8_t *bits, uint8_t *crc_ok) {
	if (!h || !soft || !bits || ngroups < 0) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_fic_decode: bad argument");
	if (ngroups == 0) return DABGPU_OK;
	void *d_in;
	int rc = stage_in (h, soft, (size_t) ngroups * 2304 * sizeof (int16_t), &d_in);
	if (rc) return rc;
	const size_t obytes = (size_t) ngroups * 768, cbytes = (size_t) ngroups * 3;
	CUDA_TRY (h, h -> d_out. ensure (obytes + cbytes));
	uint8_t *d_bits = (uint8_t *) h -> d_out. p, *d_crc = d_bits + obytes;
	if ((rc = dab_fic_decode_dev (h, (const int16_t *) d_in, nullptr, 2304, ngroups, d_bits, d_crc))) return rc;
	CUDA_TRY (h, h -> h_out. ensure (obytes + cbytes));
	CUDA_TRY (h, cudaMemcpyAsync (h -> h_out. p, d_bits, obytes + cbytes, cudaMemcpyDeviceToHost, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	memcpy (bits, h -> h_out. p, obytes);
	if (crc_ok) memcpy (crc_ok, (uint8_t *) h -> h_out. p + obytes, cbytes);
	return DABGPU_OK;
}

// ---- dabConcurrent (dab-concurrent.cpp:46-193) ----
struct dabgpu_backend {
	dabgpu *h;
	dabgpu_subch sc;
	int fragmentSize;
	const ProtProfile *pp; const uint16_t *d_lut; const uint32_t *d_prbs;
	DevBuf rows;                 // [15 history + ncif][fragmentSize] int16
	DevBuf hist;                 // the last 15 fragments (oldest first), device resident between calls
	int64_t cifs_seen;           // countforInterleaver, unbounded
};

extern "C" int dabgpu_backend_create (dabgpu_t *h, const dabgpu_subch *sc, dabgpu_backend_t **out) {
	if (!h || !sc || !out) return dab_fail (h, DABGPU_ERR_ARG, "dabgpu_backend_create: null argument");
	*out = nullptr;
	if (sc -> length <= 0 || sc -> length * 64 > 32767 || sc -> startAddr < 0 || sc -> startAddr + sc -> length > 864)
		return dab_fail (h, DABGPU_ERR_ARG, "sub-channel [%d, +%d) CUs out of range", sc -> startAddr, sc -> length);
	const ProtProfile *pp; const uint16_t *d_lut; const uint32_t *d_prbs;
	int rc = dab_get_profile (h, 1, sc -> bitRate, sc -> uepFlag, sc -> protLevel, &pp, &d_lut);
	if (rc) return rc;
	if (pp -> nPunctured > sc -> length * 64)
		return dab_fail (h, DABGPU_ERR_ARG, "profile needs %d soft bits, sub-channel holds %d", pp -> nPunctured, sc -> length * 64);
	if ((rc = dab_get_prbs (h, pp -> frameBits, &d_prbs))) return rc;
	dabgpu_backend *b = new dabgpu_backend ();
	b -> h = h; b -> sc = *sc; b -> fragmentSize = sc -> length * 64;
	b -> pp = pp; b -> d_lut = d_lut; b -> d_prbs = d_prbs; b -> cifs_seen = 0;
	const size_t hb = (size_t) 15 * b -> fragmentSize * sizeof (int16_t);
	cudaError_t e = b -> hist. ensure (hb);
	if (e == cudaSuccess) e = cudaMemsetAsync (b -> hist. p, 0, hb, h -> stream);      // dab-concurrent.cpp:70-74
	if (e != cudaSuccess) { b -> hist. release (); delete b; return dab_fail (h, DABGPU_ERR_CUDA, "backend alloc: %s", cudaGetErrorString (e)); }
	*out = b;
	return DABGPU_OK;
}

extern "C" void dabgpu_backend_destroy (dabgpu_backend_t *b) {
	if (!b) return;
	cudaSetDevice (b -> h -> device);
	cudaStreamSynchronize (b -> h -> stream);
	b -> rows. release (); b -> hist. release ();
	delete b;
}

// device-side core: rows = [15 history][ncif new] fragments; decodes the CIFs past the warm-up
int dab_backend_run_dev (dabgpu_backend *b, const int16_t *d_rows, const uint8_t *d_rows8, long long row_stride, int row0, int ncif,
                         uint8_t *d_out, int *nout, VitSimdJob *simd_job, int64_t cifs_seen) {
	dabgpu *h = b -> h;
	// dab-concurrent.cpp:172-175: the first 16 CIFs only fill the de-interleaver (cifs_seen >= 0: the caller keeps the
	// counter itself -- one backend object serving many independent streams)
	int64_t skip = 16 - (cifs_seen >= 0 ? cifs_seen : b -> cifs_seen);
	if (skip < 0) skip = 0;
	if (skip > ncif) skip = ncif;
	const int n = ncif - (int) skip;
	VitJob j {};
	j. in = d_rows; j. in8 = d_rows8; j. in_stride = row_stride; j. first_row = 15 + row0 + (int) skip; j. lut = b -> d_lut;
	j. frameBits = b -> pp -> frameBits; j. nsteps = j. frameBits + 6; j. nblocks = n;
	j. deint = 1; j. prbs = b -> d_prbs; j. out = d_out;
	*nout = n;
	if (simd_job) {                       // the caller batches several sub-channels into one SIMD launch
		memset (simd_job, 0, sizeof (*simd_job));
		int rc = dab_simd_job_profile (h, 1, b -> sc. bitRate, b -> sc. uepFlag, b -> sc. protLevel, 1, simd_job);
		if (rc) return rc;
		simd_job -> in = d_rows; simd_job -> in_stride = row_stride; simd_job -> first_row = j. first_row;
		simd_job -> ncw = n; simd_job -> prbs = b -> d_prbs; simd_job -> out = d_out;
		return DABGPU_OK;
	}
	CUDA_TRY (h, vit_launch (h, KC_VITERBI_MSC, j));
	return DABGPU_OK;
}

void dab_backend_note_cifs (dabgpu_backend *b, int ncif) { b -> cifs_seen += ncif; }
int64_t dab_backend_cifs_seen (const dabgpu_backend *b) { return b -> cifs_seen; }
void dab_backend_set_cifs_seen (dabgpu_backend *b, int64_t n) { b -> cifs_seen = n; }

extern "C" int dabgpu_backend_process (dabgpu_backend_t *b, const int16_t *frags, int32_t ncif, uint8_t *out, int32_t *nout) {
	if (!b || !frags || !out || !nout || ncif < 0) return dab_fail (b ? b -> h : nullptr, DABGPU_ERR_ARG, "dabgpu_backend_process: bad argument");
	dabgpu *h = b -> h;
	*nout = 0;
	if (ncif == 0) return DABGPU_OK;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	const size_t fs = (size_t) b -> fragmentSize, rowb = fs * sizeof (int16_t);
	CUDA_TRY (h, b -> rows. ensure ((15 + (size_t) ncif) * rowb));
	int16_t *d_rows = (int16_t *) b -> rows. p;
	CUDA_TRY (h, cudaMemcpyAsync (d_rows, b -> hist. p, 15 * rowb, cudaMemcpyDeviceToDevice, h -> stream));
	CUDA_TRY (h, h -> h_in. ensure ((size_t) ncif * rowb));
	memcpy (h -> h_in. p, frags, (size_t) ncif * rowb);
	CUDA_TRY (h, cudaMemcpyAsync (d_rows + 15 * fs, h -> h_in. p, (size_t) ncif * rowb, cudaMemcpyHostToDevice, h -> stream));
	const size_t obytes = (size_t) ncif * b -> pp -> frameBits;
	CUDA_TRY (h, h -> d_out. ensure (obytes));
	int n = 0;
	int rc;
	if (dab_use_simd (h, ncif)) {
		std::vector<VitSimdJob> jobs (1);
		if ((rc = dab_backend_run_dev (b, d_rows, nullptr, (long long) fs, 0, ncif, (uint8_t *) h -> d_out. p, &n, &jobs [0]))) return rc;
		if (n > 0 && (rc = dab_vit_simd_run (h, jobs))) return rc;
	} else if ((rc = dab_backend_run_dev (b, d_rows, nullptr, (long long) fs, 0, ncif, (uint8_t *) h -> d_out. p, &n, nullptr))) return rc;
	// new history = last 15 rows of [history | new]
	CUDA_TRY (h, cudaMemcpyAsync (b -> hist. p, d_rows + (size_t) ncif * fs, 15 * rowb, cudaMemcpyDeviceToDevice, h -> stream));
	b -> cifs_seen += ncif;
	*nout = n;
	if (n > 0) return stage_out (h, out, h -> d_out. p, (size_t) n * b -> pp -> frameBits);
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	return DABGPU_OK;
}

extern "C" int dabgpu_backend_get_state (dabgpu_backend_t *b, int16_t *hist, int32_t *cifs_seen) {
	if (!b || !hist || !cifs_seen) return DABGPU_ERR_ARG;
	dabgpu *h = b -> h;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, cudaMemcpyAsync (hist, b -> hist. p, (size_t) 15 * b -> fragmentSize * sizeof (int16_t), cudaMemcpyDeviceToHost, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	*cifs_seen = (int32_t) (b -> cifs_seen > 0x7fffffff ? 0x7fffffff : b -> cifs_seen);
	return DABGPU_OK;
}

extern "C" int dabgpu_backend_set_state (dabgpu_backend_t *b, const int16_t *hist, int32_t cifs_seen) {
	if (!b || !hist || cifs_seen < 0) return DABGPU_ERR_ARG;
	dabgpu *h = b -> h;
	CUDA_TRY (h, cudaSetDevice (h -> device));
	CUDA_TRY (h, cudaMemcpyAsync (b -> hist. p, hist, (size_t) 15 * b -> fragmentSize * sizeof (int16_t), cudaMemcpyHostToDevice, h -> stream));
	CUDA_TRY (h, cudaStreamSynchronize (h -> stream));
	b -> cifs_seen = cifs_seen;
	return DABGPU_OK;
}
