/*
 * dab_adapters.h -- drop-in C++ classes with the reference's names and member signatures that forward to the
 * C ABI of libdabgpu.so (include/dabgpu.h).  Header only; link with -ldabgpu.
 *
 * A maintainer of the reference swaps src/backend/{viterbi,deconvolve}.cpp, src/ofdm/{phasereference,
 * ofdm-decoder}.cpp for this header.  Signatures kept (reference file:line):
 *   viterbi (int16_t), void deconvolve (int16_t *, uint8_t *)                         viterbi.h:43-45
 *   eep_deconvolve / uep_deconvolve (int16_t bitRate, int16_t protLevel),
 *       bool deconvolve (int16_t *, int32_t, uint8_t *)                               deconvolve.h:34-36, 53-55
 *   phaseReference (DabParams *, int16_t), int32_t findIndex (DSPCOMPLEX *), getTable ()   phasereference.h:37-40
 *   ofdmDecoder (DabParams *, RingBuffer<DSPCOMPLEX> *, DSPCOMPLEX *, RadioInterface *, uint8_t),
 *       int16_t processBlock_0 (DSPCOMPLEX *, bool), void processToken (DSPCOMPLEX *, int16_t *, int32_t)  ofdm-decoder.h:40-47
 *   dabVirtual / dabConcurrent / dabSerial (ctor as the reference), int32_t process (int16_t *, int16_t)   dab-virtual.h:36-47, dab-concurrent.h:44-53, dab-serial.h:41-50
 *   ficHandler (RadioInterface *, int16_t), void process_ficBlock (int16_t *, int16_t), get_ficRatio ()   fic-handler.h:44-52
 *   mscHandler (RadioInterface *, DabParams *, audioSink *, uint8_t), process_mscBlock (int16_t *, int16_t),
 *       set_audioChannel (audiodata *), set_dataChannel (packetdata *), stop (), stopProcessing ()           msc-handler.h:43-57
 *   (decoded FIBs / frames leave through sinks where the reference calls fib_processor::process_FIB / dabProcessor::addtoFrame)
 * Errors: the reference's classes cannot fail after construction; these throw std::runtime_error from the
 * constructor when the engine cannot be created (no GPU) and otherwise keep the reference's sentinels
 * (negative findIndex, 100 from processBlock_0).  One engine handle per DAB mode is shared by all adapter
 * objects of a thread (the engine, like the reference objects, is not re-entrant).
 */
#ifndef DAB_ADAPTERS_H
#define DAB_ADAPTERS_H
#include <atomic>
#include <complex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <stdint.h>
#include "../../include/dabgpu.h"

#ifndef DAB_ADAPTERS_NO_TYPES          /* define when the reference's dab-constants.h is included first */
typedef float DSPFLOAT;
typedef std::complex<DSPFLOAT> DSPCOMPLEX;
struct DabParams {                     /* includes/dab-constants.h:137-149 */
	uint8_t dabMode; int16_t L, K, T_null; int32_t T_F; int16_t T_s, T_u, guardLength, carrierDiff;
};
class RadioInterface;
template <class T> class RingBuffer;
typedef struct {                       /* includes/dab-constants.h:165-176 */
	int16_t subchId, startAddr; uint8_t uepFlag; int16_t protLevel, length, bitRate, ASCTy, language, programType;
} audiodata;
typedef struct {                       /* includes/dab-constants.h:152-163 */
	int16_t subchId, startAddr; uint8_t uepFlag; int16_t protLevel, DSCTy, length, bitRate, FEC_scheme, DGflag, packetAddress;
} packetdata;
class virtualInput {                   /* src/input/virtual-input.h:51-69: the two members the OFDM thread uses */
public:
	virtual ~virtualInput (void) {}
	virtual int32_t getSamples (DSPCOMPLEX *, int32_t) = 0;      /* size in I/Q pairs; returns the number delivered */
	virtual int32_t Samples (void) = 0;
};
#endif

namespace dabgpu_host {

inline dabgpu_t *engine (int mode, int threshold = 3, int freqSyncMethod = 1) {
	struct slot { dabgpu_t *h = nullptr; int thr = 0, meth = 0; };
	static thread_local slot slots [5];
	if (mode < 1 || mode > 4) throw std::runtime_error ("dabgpu: mode out of range");
	slot &s = slots [mode];
	if (s. h && (s. thr != threshold || s. meth != freqSyncMethod)) { dabgpu_destroy (s. h); s. h = nullptr; }
	if (!s. h) {
		dabgpu_config cfg = {};
		cfg. device = 0; cfg. dabMode = mode; cfg. threshold = threshold; cfg. freqSyncMethod = freqSyncMethod;
		if (dabgpu_create (&cfg, &s. h) != DABGPU_OK)
			throw std::runtime_error (std::string ("dabgpu_create: ") + dabgpu_last_error (nullptr));
		s. thr = threshold; s. meth = freqSyncMethod;
	}
	return s. h;
}
inline void check (dabgpu_t *h, int rc) { if (rc != DABGPU_OK) throw std::runtime_error (std::string ("dabgpu: ") + dabgpu_last_error (h)); }

}	/* namespace dabgpu_host */

/* ---- viterbi.h:43-45 ---- */
class viterbi {
public:
	viterbi (int16_t wordlength) : frameBits (wordlength), h (dabgpu_host::engine (1)) {}
	~viterbi (void) {}
	void deconvolve (int16_t *input, uint8_t *output) { dabgpu_host::check (h, dabgpu_viterbi (h, input, frameBits, 1, output)); }
	/* batch extension: nblocks independent code words per call (what makes a GPU worthwhile) */
	void deconvolve (int16_t *input, uint8_t *output, int32_t nblocks) { dabgpu_host::check (h, dabgpu_viterbi (h, input, frameBits, nblocks, output)); }
protected:
	int16_t frameBits;
	dabgpu_t *h;
};

/* ---- deconvolve.h:32-64 ---- */
class uep_deconvolve : public viterbi {
public:
	uep_deconvolve (int16_t bitRate, int16_t protLevel) : viterbi (24 * bitRate), bitRate (bitRate), protLevel (protLevel) {}
	bool deconvolve (int16_t *v, int32_t size, uint8_t *outBuffer) {
		return dabgpu_protect_decode (h, bitRate, 0, protLevel, v, size, 1, outBuffer) == DABGPU_OK;
	}
private:
	int16_t bitRate, protLevel;
};
class eep_deconvolve : public viterbi {
public:
	eep_deconvolve (int16_t bitRate, int16_t protLevel) : viterbi (24 * bitRate), bitRate (bitRate), protLevel (protLevel) {}
	bool deconvolve (int16_t *v, int32_t size, uint8_t *outBuffer) {
		return dabgpu_protect_decode (h, bitRate, 1, protLevel, v, size, 1, outBuffer) == DABGPU_OK;
	}
private:
	int16_t bitRate, protLevel;
};

/* ---- phasereference.h:35-53 ---- */
class phaseReference {
public:
	phaseReference (DabParams *p, int16_t threshold) : Tu (p -> T_u), h (dabgpu_host::engine (p -> dabMode, threshold)), refTable (p -> T_u) {
		dabgpu_host::check (h, dabgpu_host_ref_table (p -> dabMode, reinterpret_cast<float *> (refTable. data ())));
	}
	int32_t findIndex (DSPCOMPLEX *v) {
		int32_t idx = -1;
		dabgpu_host::check (h, dabgpu_find_index (h, reinterpret_cast<const float *> (v), 1, &idx));
		return idx;
	}
	DSPCOMPLEX *getTable (void) { return refTable. data (); }
private:
	int32_t Tu;
	dabgpu_t *h;
	std::vector<DSPCOMPLEX> refTable;
};

/* ---- ofdm-decoder.h:38-81 ---- */
class ofdmDecoder {
public:
	ofdmDecoder (DabParams *p, RingBuffer<DSPCOMPLEX> *, DSPCOMPLEX *, RadioInterface *, uint8_t freqSyncMethod)
	   : h (dabgpu_host::engine (p -> dabMode, 3, freqSyncMethod)) {}
	int16_t processBlock_0 (DSPCOMPLEX *v, bool flag) {
		int16_t c = 100;
		dabgpu_host::check (h, dabgpu_block0 (h, reinterpret_cast<const float *> (v), flag, &c));
		return c;
	}
	void processToken (DSPCOMPLEX *inv, int16_t *ibits, int32_t /*blkno*/) {
		dabgpu_host::check (h, dabgpu_token (h, reinterpret_cast<const float *> (inv), 1, ibits));
	}
	/* batch extension: nsym consecutive symbols */
	void processTokens (DSPCOMPLEX *inv, int16_t *ibits, int32_t nsym) {
		dabgpu_host::check (h, dabgpu_token (h, reinterpret_cast<const float *> (inv), nsym, ibits));
	}
private:
	dabgpu_t *h;
};

/* ---- dabConcurrent (dab-concurrent.cpp:46-193): process() takes one CIF fragment, frames go to a sink in place
 * of dabProcessor::addtoFrame (dab-processor.h:39) ---- */
class dabBackend {
public:
	typedef std::function<void (uint8_t *, int16_t)> frameSink;
	dabBackend (uint8_t dabMode, int16_t fragmentSize, int16_t bitRate, int16_t uepFlag, int16_t protLevel, int16_t startAddr, frameSink sink)
	   : h (dabgpu_host::engine (dabMode)), bitRate (bitRate), sink (sink), outV (24 * bitRate) {
		dabgpu_subch sc = { startAddr, fragmentSize / 64, bitRate, uepFlag, protLevel };
		dabgpu_host::check (h, dabgpu_backend_create (h, &sc, &b));
	}
	~dabBackend (void) { dabgpu_backend_destroy (b); }
	int32_t process (int16_t *v, int16_t cnt) {
		int32_t n = 0;
		(void) cnt;
		dabgpu_host::check (h, dabgpu_backend_process (b, v, 1, outV. data (), &n));
		if (n == 1 && sink) sink (outV. data (), 24 * bitRate);
		return 32768;
	}
private:
	dabgpu_t *h;
	dabgpu_backend_t *b = nullptr;
	int16_t bitRate;
	frameSink sink;
	std::vector<uint8_t> outV;
};

/* ---- dabVirtual (dab-virtual.h:36-47) and its two audio back ends with the reference's constructors:
 * dabConcurrent (dab-concurrent.h:44-53; warm-up 16 CIFs, dab-concurrent.cpp:172-175) and dabSerial (dab-serial.h:41-50;
 * warm-up 15 CIFs, dab-serial.cpp:126-129).  process () takes one CIF fragment (the slice mscHandler cuts out of the
 * CIF); the decoded, dispersal-descrambled frame goes to a sink in place of dabProcessor::addtoFrame (dab-processor.h:39;
 * the MP2 / AAC decoders behind it stay on the host).  `dabModus` is the reference's DAB / DAB_PLUS audio type, which the
 * channel decoder does not depend on; the FILE and audioSink arguments belong to the audio decoders and are ignored. ---- */
#ifndef DAB_ADAPTERS_NO_TYPES
class audioSink;
#endif
#ifndef CUSize
#define CUSize (4 * 16)
#endif
class dabVirtual {
public:
	typedef std::function<void (uint8_t *, int16_t)> frameSink;
	dabVirtual (void) {}
	virtual ~dabVirtual (void) {}
	virtual int32_t process (int16_t *, int16_t) { return 32768; }
	virtual void stopRunning (void) {}
	virtual void stop (void) {}
	virtual void setFiles (FILE *, FILE *) {}
	virtual void set_sink (frameSink) {}
};
class dabGpuBackend_ : public dabVirtual {
public:
	dabGpuBackend_ (int16_t fragmentSize, int16_t bitRate, int16_t uepFlag, int16_t protLevel, int warmup)
	   : h (dabgpu_host::engine (1)), bitRate (bitRate), outV (24 * bitRate) {
		dabgpu_subch sc = { 0, fragmentSize / CUSize, bitRate, uepFlag, protLevel };
		dabgpu_host::check (h, dabgpu_backend_create (h, &sc, &b));
		if (warmup != 16) {                /* dabSerial starts decoding one CIF earlier: same de-interleaver, counter one ahead */
			std::vector<int16_t> zero ((size_t) 15 * fragmentSize, 0);
			dabgpu_host::check (h, dabgpu_backend_set_state (b, zero. data (), 16 - warmup));
		}
	}
	~dabGpuBackend_ (void) { dabgpu_backend_destroy (b); }
	int32_t process (int16_t *v, int16_t cnt) {
		int32_t n = 0;
		(void) cnt;                        /* "sorry" (dab-serial.cpp:114): the reference ignores it as well */
		dabgpu_host::check (h, dabgpu_backend_process (b, v, 1, outV. data (), &n));
		if (n == 1 && sink) sink (outV. data (), (int16_t) (24 * bitRate));
		return 32768;
	}
	void set_sink (frameSink s) { sink = s; }
private:
	dabgpu_t *h;
	dabgpu_backend_t *b = nullptr;
	int16_t bitRate;
	frameSink sink;
	std::vector<uint8_t> outV;
};
class dabConcurrent : public dabGpuBackend_ {
public:
	dabConcurrent (uint8_t dabModus, int16_t fragmentSize, int16_t bitRate, int16_t uepFlag, int16_t protLevel,
	               RadioInterface *mr, FILE *, FILE *, audioSink *as)
	   : dabGpuBackend_ (fragmentSize, bitRate, uepFlag, protLevel, 16) { (void) dabModus; (void) mr; (void) as; }
};
class dabSerial : public dabGpuBackend_ {
public:
	dabSerial (uint8_t dabModus, int16_t fragmentSize, int16_t bitRate, int16_t uepFlag, int16_t protLevel,
	           RadioInterface *mr, FILE *, FILE *, audioSink *as)
	   : dabGpuBackend_ (fragmentSize, bitRate, uepFlag, protLevel, 15) { (void) dabModus; (void) mr; (void) as; }
};

/* ---- ficHandler::process_ficBlock (fic-handler.cpp:143-230): collects 2304-bit code words, decodes them, hands
 * CRC-clean FIBs to a sink in place of fib_processor::process_FIB (fib-processor.h:96) ---- */
class ficDecoder {
public:
	typedef std::function<void (uint8_t *, uint16_t)> fibSink;
	ficDecoder (DabParams *p, fibSink sink) : h (dabgpu_host::engine (p -> dabMode)), BitsperBlock (2 * p -> K), sink (sink) {}
	void process_ficBlock (int16_t *data, int16_t blkno) {
		if (blkno == 1) { index = 0; ficno = 0; }
		for (int i = 0; i < BitsperBlock; i ++) {
			ofdm_input [index ++] = data [i];
			if (index >= 2304) {
				uint8_t bits [768], crc [3];
				dabgpu_host::check (h, dabgpu_fic_decode (h, ofdm_input, 1, bits, crc));
				for (int f = 0; f < 3; f ++) if (crc [f] && sink) sink (&bits [256 * f], ficno);
				index = 0; ficno ++;
			}
		}
	}
private:
	dabgpu_t *h;
	int BitsperBlock, index = 0, ficno = 0;
	fibSink sink;
	int16_t ofdm_input [2304];
};

/* ---- mscDatagroup::run (data/msc-datagroup.cpp:149-198): the packet-mode / data sub-channel front half is the very
 * chain dabConcurrent runs (time de-interleave, eep/uep depuncture + Viterbi, dispersal); only the sink differs ---- */
typedef dabBackend dataBackend;

/* ---- fib_processor's FIG 0/1 path (fib-processor.cpp:123-158, 278-347): feed CRC-clean FIC groups, read ficList.
 * The other FIG types (labels, service lists) stay with the host's own fib_processor ---- */
class figSubchannels {
public:
	explicit figSubchannels (uint8_t dabMode) : h (dabgpu_host::engine (dabMode)) { memset (ficList, 0, sizeof (ficList)); }
	/* groups of 768 bits with their three CRC flags, as dabgpu_fic_decode returns them; later calls override earlier ones */
	void process_groups (const uint8_t *bits768, const uint8_t *crc3, int32_t ngroups) {
		dabgpu_subch_info t [64];
		dabgpu_host::check (h, dabgpu_fig01_scan (h, bits768, crc3, ngroups, t));
		for (int i = 0; i < 64; i ++) {
			if (!t [i]. defined) continue;
			ficList [i]. defined = 1; ficList [i]. startAddr = t [i]. startAddr; ficList [i]. uepFlag = t [i]. uepFlag;
			if (t [i]. bitRate || t [i]. length) { ficList [i]. length = t [i]. length; ficList [i]. protLevel = t [i]. protLevel; ficList [i]. bitRate = t [i]. bitRate; }
		}
	}
	dabgpu_subch_info ficList [64];
private:
	dabgpu_t *h;
};

/* ---- mp4Processor's super-frame front (mp4processor.cpp:107-275): addtoFrame per CIF as the reference calls it; every
 * repaired super frame goes to a sink together with its access-unit table (the AAC decoder stays on the host) ---- */
class mp4SuperframeFront {
public:
	typedef std::function<void (const uint8_t *, int32_t, const dabgpu_superframe_info &)> superframeSink;
	mp4SuperframeFront (uint8_t dabMode, int16_t bitRate, superframeSink sink)
	   : h (dabgpu_host::engine (dabMode)), bitRate (bitRate), sink (sink), out (110 * (bitRate / 8)) {
		dabgpu_host::check (h, dabgpu_dabplus_create (h, bitRate, &d));
	}
	~mp4SuperframeFront (void) { dabgpu_dabplus_destroy (d); }
	void addtoFrame (uint8_t *V, int16_t nbits) {
		int32_t n = 0;
		dabgpu_superframe_info fi;
		(void) nbits;
		dabgpu_host::check (h, dabgpu_dabplus_process (d, V, 1, out. data (), &fi, 1, &n));
		if (n == 1 && sink) sink (out. data (), (int32_t) out. size (), fi);
	}
private:
	dabgpu_t *h;
	dabgpu_dabplus_t *d = nullptr;
	int16_t bitRate;
	superframeSink sink;
	std::vector<uint8_t> out;
};

/* ------------------------------------------------------------------------------------------------------------
 * The throughput substitution: ofdmProcessor with the reference's constructor and control members
 * (ofdm-processor.h:49-68) whose run () hands blocks of samples from virtualInput::getSamples to the stream engine
 * (dabgpu_decode_cf32) instead of walking them sample by sample (ofdm-processor.cpp:247-474).  Decoded FIBs and
 * sub-channel frames come back through ficHandler / mscHandler objects that carry the sinks the reference's handlers
 * end in: fib_processor::process_FIB (fib-processor.h:96, called at fic-handler.cpp:309-319) and
 * dabProcessor::addtoFrame (dab-processor.h:39, called at dab-concurrent.cpp:191).  The same two classes also carry the
 * reference's per-symbol members process_ficBlock / process_mscBlock (substitution 1: the reference's own sample-serial
 * ofdmProcessor keeps calling them); with the throughput ofdmProcessor below soft bits never leave the GPU and the
 * engine calls deliver () instead.
 * ---------------------------------------------------------------------------------------------------------- */
class ficHandler {                     /* fic-handler.h:44-46 */
public:
	typedef std::function<void (uint8_t *, uint16_t)> fibSink;
	/* the reference's constructor (fic-handler.cpp:82-83: the second argument is the soft bits per OFDM symbol, 2 K; the
	 * DAB mode is accepted as well and mapped to it) */
	ficHandler (RadioInterface *mr, int16_t bitsperBlock_or_mode, fibSink sink = nullptr) : sink (sink) {
		(void) mr;
		static const int16_t k2 [5] = { 0, 2 * 1536, 2 * 384, 2 * 192, 2 * 768 };
		BitsperBlock = bitsperBlock_or_mode >= 1 && bitsperBlock_or_mode <= 4 ? k2 [bitsperBlock_or_mode] : bitsperBlock_or_mode;
	}
	void set_sink (fibSink s) { sink = s; }
	/* fic-handler.cpp:143-153 + run () :192-230: the soft bits of FIC symbol blkno (1, 2 or 3); every completed 2304-bit
	 * code word is decoded (process_ficInput, :241-321) and its CRC-clean FIBs go to the sink with their ficno */
	void process_ficBlock (int16_t *data, int16_t blkno) {
		if (!h) h = dabgpu_host::engine (1);
		if (blkno == 1) { index = 0; ficno = 0; }
		for (int i = 0; i < BitsperBlock; i ++) {
			ofdm_input [index ++] = data [i];
			if (index >= 2304) {
				uint8_t bits [768], crc [3];
				dabgpu_host::check (h, dabgpu_fic_decode (h, ofdm_input, 1, bits, crc));
				deliver_group (bits, crc, ficno);
				index = 0; ficno ++;
			}
		}
	}
	int16_t get_ficRatio (void) { return total ? (int16_t) (100 * good / total) : 0; }      /* fic-handler.cpp:323-325 */
	void clearEnsemble (void) {}           /* the service directory (fib_processor) stays on the host */
	void stop (void) {}
	/* engine side: the FIC groups of one decode call (768 bits + 3 CRC flags each, ficGroups per frame) */
	void deliver (uint8_t *bits768, const uint8_t *crc3, int32_t ngroups, int32_t groupsPerFrame) {
		for (int32_t g = 0; g < ngroups; g ++) deliver_group (&bits768 [768 * g], &crc3 [3 * g], (uint16_t) (g % groupsPerFrame));
	}
private:
	void deliver_group (uint8_t *bits768, const uint8_t *crc3, uint16_t no) {
		for (int f = 0; f < 3; f ++) {
			total ++;
			if (!crc3 [f]) continue;
			good ++;
			if (sink) sink (&bits768 [256 * f], no);                                 /* fic-handler.cpp:309-319 */
		}
	}
	fibSink sink;
	dabgpu_t *h = nullptr;
	int16_t BitsperBlock, index = 0, ficno = 0;
	int16_t ofdm_input [2304];
	long long good = 0, total = 0;
};

class mscHandler {                     /* msc-handler.h:43-57 */
public:
	typedef std::function<void (uint8_t *, int16_t)> frameSink;
	mscHandler (RadioInterface *mr, DabParams *p, audioSink *as, uint8_t concurrencyOn, frameSink sink = nullptr)
	   : sink (sink), concurrencyOn (concurrencyOn), myRadioInterface (mr), our_audioSink (as), cifVector (55296) {
		BitsperBlock = 2 * p -> K;                                               /* msc-handler.cpp:61-71 */
		numberofblocksperCIF = p -> dabMode == 4 ? 36 : p -> dabMode == 1 ? 18 : p -> dabMode == 2 ? 72 : 18;
	}
	~mscHandler (void) { delete dabHandler; }
	void set_sink (frameSink s) { sink = s; if (dabHandler) dabHandler -> set_sink (s); }
	void set_audioChannel (audiodata *d) {                        /* msc-handler.cpp:91-105: takes effect at the next block */
		std::lock_guard<std::mutex> g (m);
		sc = { d -> startAddr, d -> length, d -> bitRate, d -> uepFlag, d -> protLevel }; have = true; changed = true; newChannel = true;
		new_dabModus = d -> ASCTy == 077 ? 1 : 0;
	}
	void set_dataChannel (packetdata *d) {                        /* msc-handler.cpp:107-123 */
		std::lock_guard<std::mutex> g (m);
		sc = { d -> startAddr, d -> length, d -> bitRate, d -> uepFlag, d -> protLevel }; have = true; changed = true; newChannel = true;
		new_dabModus = 0;
	}
	/* msc-handler.cpp:125-193: the soft bits of MSC symbol blkno (4 .. L); CIF assembly, sub-channel slice, hand-over to
	 * the back end (dabConcurrent, or dabSerial when the object was built with concurrencyOn == 0) */
	void process_mscBlock (int16_t *fbits, int16_t blkno) {
		if (!work_to_be_done && !newChannel) return;
		const int16_t currentblk = (blkno - 4) % numberofblocksperCIF;
		if (newChannel) {
			std::lock_guard<std::mutex> g (m);
			newChannel = false;
			if (dabHandler) dabHandler -> stopRunning ();
			delete dabHandler;
			if (concurrencyOn)
				dabHandler = new dabConcurrent (new_dabModus, sc. length * CUSize, sc. bitRate, sc. uepFlag, sc. protLevel, myRadioInterface, nullptr, nullptr, our_audioSink);
			else
				dabHandler = new dabSerial (new_dabModus, sc. length * CUSize, sc. bitRate, sc. uepFlag, sc. protLevel, myRadioInterface, nullptr, nullptr, our_audioSink);
			dabHandler -> set_sink (sink);
			startAddr = sc. startAddr; Length = sc. length;
			work_to_be_done = true;
		}
		memcpy (&cifVector [currentblk * BitsperBlock], fbits, BitsperBlock * sizeof (int16_t));
		if (currentblk < numberofblocksperCIF - 1) return;
		(void) dabHandler -> process (&cifVector [startAddr * CUSize], Length * CUSize);
	}
	void stopProcessing (void) { std::lock_guard<std::mutex> g (m); have = false; changed = true; work_to_be_done = false; }
	void stop (void) { work_to_be_done = false; if (dabHandler) dabHandler -> stop (); }
	/* engine side (the throughput ofdmProcessor below) */
	bool take_change (dabgpu_subch *out, bool *active) {
		std::lock_guard<std::mutex> g (m);
		if (!changed) return false;
		changed = false; *out = sc; *active = have;
		return true;
	}
	void deliver (uint8_t *frames, int32_t nframes, int32_t frameBits) {
		for (int32_t i = 0; i < nframes && sink; i ++) sink (frames + (size_t) i * frameBits, (int16_t) frameBits);
	}
private:
	frameSink sink;
	uint8_t concurrencyOn;
	RadioInterface *myRadioInterface;
	audioSink *our_audioSink;
	std::vector<int16_t> cifVector;
	std::mutex m;
	dabgpu_subch sc {};
	dabVirtual *dabHandler = nullptr;
	bool have = false, changed = false, newChannel = false, work_to_be_done = false;
	uint8_t new_dabModus = 0;
	int16_t BitsperBlock, numberofblocksperCIF, startAddr = 0, Length = 0;
};

class ofdmProcessor {                  /* ofdm-processor.h:49-68 (the build without HAVE_SPECTRUM) */
public:
	ofdmProcessor (virtualInput *theRig, DabParams *p, RadioInterface *mr, mscHandler *msc, ficHandler *fic,
	               int16_t threshold, RingBuffer<DSPCOMPLEX> *iqBuffer, uint8_t freqSyncMethod, int32_t framesPerCall = 8)
	   : theRig (theRig), params (*p), my_mscHandler (msc), my_ficHandler (fic), framesPerCall (framesPerCall) {
		(void) mr; (void) iqBuffer;
		dabgpu_config cfg = {};
		cfg. device = 0; cfg. dabMode = p -> dabMode; cfg. threshold = threshold; cfg. freqSyncMethod = freqSyncMethod;
		if (dabgpu_create (&cfg, &h) != DABGPU_OK) throw std::runtime_error (std::string ("dabgpu_create: ") + dabgpu_last_error (nullptr));
		int32_t mp [12];
		dabgpu_host_mode_params (p -> dabMode, mp);
		ficGroups = mp [9]; cifsPerFrame = mp [10];
	}
	~ofdmProcessor (void) { dabgpu_destroy (h); }
	void reset (void) { pending_reset = true; }                   /* ofdm-processor.cpp:476-479 */
	void stop (void) { running = false; }
	void setOffset (int32_t) {}                                   /* declared by the reference, never defined there either */
	void coarseCorrectorOn (void) { pending_coarse = 1; }         /* :499-502 */
	void coarseCorrectorOff (void) { pending_coarse = 0; }        /* :504-506 */
	void startDumping (void *) {}                                 /* the raw-sample dump stays with the input side */
	void stopDumping (void) {}
	void set_scanMode (bool) {}
	int64_t frames_decoded (void) const { return nframes_total; }
	/* the thread body (QThread::run in the reference): returns when stop () was called or the input ran dry */
	void run (void) {
		running = true;
		const size_t want = (size_t) framesPerCall * params. T_F;
		std::vector<DSPCOMPLEX> iq (want);
		std::vector<uint8_t> fic ((size_t) (framesPerCall + 2) * ficGroups * 768), crc ((size_t) (framesPerCall + 2) * ficGroups * 3), msc;
		uint8_t *mscp [1] = { nullptr }; int32_t nblk [1] = { 0 };
		dabgpu_subch sc {}; bool active = false;
		while (running) {
			if (pending_reset. exchange (false)) dabgpu_host::check (h, dabgpu_reset (h));
			const int pc = pending_coarse. exchange (-1);
			if (pc >= 0) dabgpu_host::check (h, dabgpu_coarse_corrector (h, pc));
			if (my_mscHandler && my_mscHandler -> take_change (&sc, &active))
				dabgpu_host::check (h, dabgpu_set_subchannels (h, &sc, active ? 1 : 0));
			const int32_t n = theRig -> getSamples (iq. data (), (int32_t) want);
			if (n <= 0) break;
			if (active) msc. resize ((size_t) (framesPerCall + 2) * cifsPerFrame * 24 * sc. bitRate);
			mscp [0] = active ? msc. data () : nullptr;
			dabgpu_result r = {};
			r. max_frames = framesPerCall + 2; r. fic_bits = fic. data (); r. fic_crc = crc. data (); r. msc_bits = mscp; r. msc_nblocks = nblk;
			dabgpu_host::check (h, dabgpu_decode_cf32 (h, reinterpret_cast<const float *> (iq. data ()), (size_t) n, &r));
			nframes_total += r. nframes;
			if (my_ficHandler) my_ficHandler -> deliver (fic. data (), crc. data (), r. nframes * ficGroups, ficGroups);
			if (my_mscHandler && active) my_mscHandler -> deliver (msc. data (), nblk [0], 24 * sc. bitRate);
		}
		running = false;
	}
private:
	virtualInput *theRig;
	DabParams params;
	mscHandler *my_mscHandler;
	ficHandler *my_ficHandler;
	int32_t framesPerCall, ficGroups = 4, cifsPerFrame = 4;
	dabgpu_t *h = nullptr;
	std::atomic<bool> running { false }, pending_reset { false };
	std::atomic<int> pending_coarse { -1 };
	int64_t nframes_total = 0;
};

#endif
