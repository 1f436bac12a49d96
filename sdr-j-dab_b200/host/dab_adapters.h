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
 *   dabConcurrent::process (int16_t *, int16_t) -> here dabBackend with a frame sink    dab-concurrent.h, dab-virtual.h:40
 *   ficHandler::process_ficBlock (int16_t *, int16_t) -> here ficDecoder with a FIB sink   fic-handler.h:44-46
 * Errors: the reference's classes cannot fail after construction; these throw std::runtime_error from the
 * constructor when the engine cannot be created (no GPU) and otherwise keep the reference's sentinels
 * (negative findIndex, 100 from processBlock_0).  One engine handle per DAB mode is shared by all adapter
 * objects of a thread (the engine, like the reference objects, is not re-entrant).
 */
#ifndef DAB_ADAPTERS_H
#define DAB_ADAPTERS_H
#include <complex>
#include <cstring>
#include <functional>
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

#endif
