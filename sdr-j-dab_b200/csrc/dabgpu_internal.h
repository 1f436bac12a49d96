// dabgpu_internal.h -- shared declarations of libdabgpu.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/dabgpu.h"

#define DAB_INPUT_RATE 2048000          // includes/dab-constants.h:79

struct DabParams {                      // gui.cpp:1328-1372 + derived geometry
	int dabMode, L, K, T_null, T_F, T_s, T_u, T_g, carrierDiff;
	int ficGroups, cifsPerFrame, blocksPerCIF;
};
int dab_mode_params (int mode, DabParams *p);

// growable device / pinned-host buffers.  Buffers released by a handle go to a process-wide cache instead of back to the
// driver, and a new handle takes them from there: cudaMalloc / cudaFree / cudaMallocHost take milliseconds each and wait for
// every kernel running on the device -- with many short-lived handles (one per short stream, several host threads) they
// were what the threads queued on.  release () may only be called while none of the handle's streams still uses the
// buffer (the API is synchronous: every entry point returns with its streams idle); growing a live buffer frees the old
// block to the driver, which synchronises by itself.
#include <mutex>
struct BufCache {
	std::mutex m;
	std::multimap<size_t, void *> dev [16], pin;
	size_t dev_bytes [16] = {}, pin_bytes = 0;
	static constexpr size_t DEV_LIMIT = 8ull << 30, PIN_LIMIT = 2ull << 30;
	void *take (std::multimap<size_t, void *> &pool, size_t &total, size_t want, size_t *got) {
		std::lock_guard<std::mutex> g (m);
		auto it = pool. lower_bound (want);
		if (it == pool. end () || it -> first > 2 * want + (1u << 20)) return nullptr;
		void *p = it -> second; *got = it -> first; total -= it -> first;
		pool. erase (it);
		return p;
	}
	bool give (std::multimap<size_t, void *> &pool, size_t &total, size_t limit, size_t cap, void *p) {
		std::lock_guard<std::mutex> g (m);
		if (total + cap > limit) return false;
		pool. emplace (cap, p); total += cap;
		return true;
	}
};
inline BufCache &buf_cache () { static BufCache c; return c; }

struct DevBuf {
	void *p = nullptr; size_t cap = 0; int dev = 0;
	cudaError_t ensure (size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		// growing: the old buffer goes to the cache, not back to the driver (cudaFree costs about a millisecond each, and a handle
		// that starts small -- a lead-in call -- grows some twenty buffers at its first big call).  Work that may still read it is
		// waited for first, as cudaFree would have done implicitly.
		if (p) { cudaDeviceSynchronize (); release (); }
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 4 + 256;
		cudaGetDevice (&dev);
		if (dev < 0 || dev >= 16) dev = 0;
		BufCache &c = buf_cache ();
		if ((p = c. take (c. dev [dev], c. dev_bytes [dev], want, &cap))) return cudaSuccess;
		cudaError_t e = cudaMalloc (&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release () {
		BufCache &c = buf_cache ();
		if (p && !c. give (c. dev [dev], c. dev_bytes [dev], BufCache::DEV_LIMIT, cap, p)) cudaFree (p);
		p = nullptr; cap = 0;
	}
};
struct PinBuf {
	void *p = nullptr; size_t cap = 0;
	cudaError_t ensure (size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) { cudaDeviceSynchronize (); release (); }
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 4 + 256;
		BufCache &c = buf_cache ();
		if ((p = c. take (c. pin, c. pin_bytes, want, &cap))) return cudaSuccess;
		cudaError_t e = cudaMallocHost (&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release () {
		BufCache &c = buf_cache ();
		if (p && !c. give (c. pin, c. pin_bytes, BufCache::PIN_LIMIT, cap, p)) cudaFreeHost (p);
		p = nullptr; cap = 0;
	}
};

// ---- puncturing / protection profiles (host side, dabgpu_tables.cpp) ----
// mother-code index -> input index LUT (uint16, 0xFFFF = punctured/erasure), 4*(frameBits+6) entries
struct ProtProfile {
	int frameBits = 0;                  // 24 * bitRate (768 for the FIC)
	int nPunctured = 0;                 // input soft bits consumed per code word
	std::vector<uint16_t> lut;
	std::vector<int32_t>  chunk;        // throughput kernel: per VS_CHUNK-step chunk {first staged column a_k, 8-byte granules G_k}
	std::vector<uint32_t> gather [2];   // [deint][4 * step + j]: tile byte offset of symbol j relative to the code word's row
};
#define VS_CHUNK 40                     // trellis steps per staged chunk of the throughput Viterbi (one renormalisation each)
// steps per code word in the decision buffer: whole chunks, plus one pair so that the chain-back may fetch the (unused) odd
// partner of the last step
__host__ __device__ inline int vs_npad (int nsteps) { return (nsteps + VS_CHUNK - 1) / VS_CHUNK * VS_CHUNK + 2; }
#define VS_PITCH 184                    // tile row pitch in bytes: 8 x odd (conflict-free over 16 code words), >= 21 granules + pad column
void prot_build_gather (ProtProfile *pp);
int  prot_build_identity (int frameBits, ProtProfile *pp);                    // viterbi.cpp:225-242 (no puncturing)
int  prot_build_fic (ProtProfile *pp);                                        // fic-handler.cpp:254-288
int  prot_build_msc (int bitRate, int uepFlag, int protLevel, ProtProfile *pp); // deconvolve.cpp:142-182, 244-319
void prbs_packed (int nbits, std::vector<uint32_t> *words);                   // fic-handler.cpp:100-108

// ---- Viterbi group (dabgpu_viterbi.cu) ----
struct VitJob {
	const int16_t *in;        // soft-bit source (int16), or
	const uint8_t *in8;       // the same as 0..255 Viterbi symbols (exactly one of the two is set)
	long long in_stride;      // elements between consecutive code words (rows when deint)
	int first_row;            // deint: buffer row of the CIF decoded by block 0
	const uint16_t *lut;      // device LUT [4*nsteps] or nullptr (identity); 0xFFFF = erasure
	int frameBits, nsteps, nblocks;
	int deint;                // 1: value(idx) is read from row (blk + first_row - D[idx & 15])
	const uint32_t *prbs;     // packed energy-dispersal sequence or nullptr
	uint8_t *out;             // [nblocks][frameBits], one bit per byte
};
cudaError_t vit_launch (dabgpu *h, int cls, const VitJob &job);

// ---- throughput Viterbi (dabgpu_vit_simd.cu): two threads per code word, several jobs per launch ----
struct VitSimdJob {
	const int16_t *in;        // soft-bit source
	long long in_stride;      // elements between consecutive code words (rows when deint)
	int first_row;            // deint: buffer row of the CIF decoded by code word 0
	int deint;
	int ncols;                // input columns a code word consumes (punctured length)
	const uint4 *gather;      // device: per trellis step the four tile byte offsets (ProtProfile::gather)
	const int2 *chunk;        // device: per chunk {a_k, G_k}
	uint8_t *sym8;            // 0..255 symbols of the rows this job reads: row r = source row first_row - dmax + r
	long long stride8;        // bytes between sym8 rows (multiple of 8)
	int sym8_ready;           // 1: sym8 was filled by the producer of the soft bits, the conversion pre-pass skips this job
	int frameBits, nsteps, ncw;
	int cta_first;            // first CTA of this job in the chain-back launch (64 code words per CTA)
	int cta_first2;           // same for the forward kernel (VS_CW code words per CTA)
	unsigned one;             // = 1, set by dab_vit_simd_run; opaque to the compiler on purpose (see vs_acs)
	uint2 *dec;               // [ncw / 32][vs_npad (nsteps) / 2][32] uint4 decision words of a step pair (dabgpu_vit_simd.cu)
	const uint32_t *prbs;     // packed dispersal sequence or nullptr
	uint8_t *out;             // [ncw][frameBits], or [ncw][frameBits / 8] when packed
	int packed;               // 1: eight decoded bits per output byte, first bit on top (frameBits must be a multiple of 32)
};
cudaError_t vit_simd_launch (dabgpu *h, const VitSimdJob *d_jobs, int njobs, int total_ctas, int total_ctas2, bool convert);
int vit_simd_cw_per_cta ();
int vit_simd_tb_cw_per_cta ();
cudaError_t fib_crc_launch (dabgpu *h, const uint8_t *bits, int nfibs, uint8_t *ok);
// FIG 0/1 scan (dabgpu_fig.cu): keys = 128 persistent 64-bit words per stream, order0 = global number of the first FIB
cudaError_t fig01_launch (dabgpu *h, const uint8_t *d_bits, const uint8_t *d_crc, int nfibs, unsigned long long order0, unsigned long long *d_keys, cudaStream_t st);
int fig01_table (dabgpu *h, const unsigned long long *d_keys, dabgpu_subch_info *host_table, cudaStream_t st);

// kernel classes for the optional per-launch CUDA-event profile (dabgpu_profile_*)
enum { KC_ACQUIRE = 0, KC_FRONT, KC_SYMBOL, KC_SCAN, KC_VITERBI_MSC, KC_VITERBI_FIC, KC_VITERBI_API, KC_CRC, KC_VITERBI_TB, KC_VITERBI_SYM, KC_COUNT };
struct ProfPair { int cls; cudaEvent_t a, b; };

struct Engine;
struct dabgpu {
	int device = 0;
	Engine *engine = nullptr;
	dabgpu_config cfg {};
	DabParams p {};
	cudaStream_t stream = nullptr;
	std::string err;
	int64_t launches = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	bool profiling = false;
	std::vector<ProfPair> prof_pending;
	std::vector<cudaEvent_t> prof_pool;
	double prof_ms [KC_COUNT] = {0};
	int64_t prof_n [KC_COUNT] = {0};
	// staging
	DevBuf d_in, d_out, d_aux;
	// channel-decoding contexts: 0 = the handle's main stream, 1..3 = side streams the stream engine uses to
	// overlap the Viterbi work of finished chunks with the OFDM work of the next one.  Launch helpers use the
	// current context (the handle is not re-entrant, so an implicit current context is safe).
	struct VitCtx { cudaStream_t st = nullptr; DevBuf d_dec, d_jobs, d_sym8; PinBuf h_jobs; };
	VitCtx vctx [4];
	int cur = 0;
	cudaStream_t vst () const { return vctx [cur]. st; }
	PinBuf h_in, h_out;
	// cached device tables
	std::map<long long, void *> d_tables;          // key -> device pointer (LUTs, PRBS)
	std::vector<DevBuf> d_table_bufs;              // their storage
	std::map<long long, ProtProfile> profiles;
};

// brackets one kernel launch with events when profiling is on
struct ProfScope {
	dabgpu *h; int cls; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
	ProfScope (dabgpu *h_, int cls_, cudaStream_t st_ = nullptr);
	~ProfScope ();
};
extern thread_local std::string g_create_error;
int  dab_fail (dabgpu *h, int code, const char *fmt, ...);
#define CUDA_TRY(h, expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
	return dab_fail ((h), DABGPU_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString (e__), __FILE__, __LINE__); } while (0)

// device copy of a host table, cached in the handle under `key`
int  dab_device_table (dabgpu *h, long long key, const void *host, size_t bytes, void **dev);
int  dab_get_profile (dabgpu *h, int kind /*0 fic, 1 msc*/, int bitRate, int uepFlag, int protLevel,
                      const ProtProfile **pp, const uint16_t **d_lut);
int  dab_get_prbs (dabgpu *h, int nbits, const uint32_t **d_prbs);
// fills the profile-dependent fields of a throughput-Viterbi job (gather / chunk tables, ncols, frameBits, nsteps); kind 0 fic, 1 msc, 3 unpunctured (bitRate = frameBits)
int  dab_simd_job_profile (dabgpu *h, int kind, int bitRate, int uepFlag, int protLevel, int deint, VitSimdJob *job);
// runs a set of jobs on the one-code-word-per-thread kernels (fills cta_first / dec itself)
int  dab_vit_simd_run (dabgpu *h, std::vector<VitSimdJob> &jobs);
bool dab_use_simd (const dabgpu *h, long long ncodewords);
int  dab_fic_decode_dev (dabgpu *h, const int16_t *d_soft, const uint8_t *d_soft8, long long stride, int ngroups, uint8_t *d_bits, uint8_t *d_crc);   // soft bits as int16 or as byte symbols
struct dabgpu_backend;
int  dab_backend_run_dev (dabgpu_backend *b, const int16_t *d_rows, const uint8_t *d_rows8, long long row_stride, int row0, int ncif, uint8_t *d_out, int *nout, VitSimdJob *simd_job, int64_t cifs_seen = -1);
int  dab_fic_simd_job (dabgpu *h, const int16_t *d_soft, long long stride, int ngroups, uint8_t *d_bits, VitSimdJob *s);
void soft16_launch (dabgpu *h, const uint8_t *fic8, const uint8_t *msc8, int16_t *out, int nframes, cudaStream_t st);
void dab_backend_note_cifs (dabgpu_backend *b, int ncif);
int64_t dab_backend_cifs_seen (const dabgpu_backend *b);
void dab_backend_set_cifs_seen (dabgpu_backend *b, int64_t n);
// OFDM / stream engine state (dabgpu_ofdm.cu, dabgpu_engine.cu)
int  dab_engine_init (dabgpu *h);
void dab_engine_free (dabgpu *h);
