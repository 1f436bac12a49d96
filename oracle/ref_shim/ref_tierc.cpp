/*
 * ref_tierc.cpp -- drives the reference's OWN handler classes, compiled UNMODIFIED from /root/reference by oracle/Makefile
 * (fic-handler.cpp, fib-processor.cpp, msc-handler.cpp, dab-concurrent.cpp, dab-virtual.cpp, dab-processor.cpp), so that
 * the oracle's restated Tier-C loops (dab_oracle.c: FIC regrouping / depuncturing / PRBS / CRC, FIG 0/1, CIF assembly, time
 * de-interleaver + warm-up + energy dispersal) are pinned against compiled reference code, not only by reading.
 * Qt is replaced by the stand-ins in this directory (QThread on std::thread, QSemaphore, QMutex, QWaitCondition, QString),
 * the audio decoders and the audio sink by recorders.  Test infrastructure only; exported from oracle/_ref/libdabref.so only.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <condition_variable>
#include <QObject>
#include <QThread>
#include <QMutex>
#include <QSemaphore>
#include <QWaitCondition>
#include <sstream>
#include <iostream>
#include <complex>
#include <limits>
#include <cmath>
#include <cstring>
#include "dab-constants.h"
#include "viterbi.h"
#include "deconvolve.h"
#include "audiosink.h"
#define private public
#include "fic-handler.h"
#include "fib-processor.h"
#include "msc-handler.h"
#include "mp4processor.h"
#undef private
#include "dab-concurrent.h"
#include "dab-serial.h"
#include "charsets.h"
#include "gui.h"
#include "../dab_oracle.h"

/* ---- what Qt's moc would generate: the signals.  show_ficCRC doubles as the recorder of the FIC path ---- */
struct FicRec { std::vector<uint8_t> bits, crc; int calls = 0; };
static std::mutex g_fic_m;
static std::map<const ficHandler *, FicRec> g_fic;

void ficHandler::show_ficCRC (bool ok) {
	/* called once per FIB, in order, from process_ficInput right after the decode + PRBS of the 768-bit group (fic-handler.cpp:306-315) */
	std::lock_guard<std::mutex> l (g_fic_m);
	FicRec &r = g_fic [this];
	/* FIB k of the group is recorded at its own call; check_CRC_bits (dab-constants.h) has just inverted the FIB's 16 CRC bits IN
	 * PLACE, which is undone in the copy: the record holds the decoded, de-scrambled bits as the Viterbi + PRBS left them */
	const int k = r. calls % 3;
	const size_t at = r. bits. size ();
	r. bits. insert (r. bits. end (), bitBuffer_out + 256 * k, bitBuffer_out + 256 * k + 256);
	for (int i = 240; i < 256; i ++) r. bits [at + i] ^= 1;
	r. crc. push_back (ok ? 1 : 0);
	r. calls ++;
}
void mp4Processor::show_successRate (int) {}
void fib_processor::addEnsembleChar (char, int) {}
void fib_processor::addtoEnsemble (const QString &) {}
void fib_processor::nameofEnsemble (int, const QString &) {}
void fib_processor::technicalData (int, int, int, int, int, int, int) {}
QString toQStringUsingCharset (const char *buffer, CharacterSet, int size) {
	return QString (std::string (buffer, size < 0 ? strlen (buffer) : (size_t) size). c_str ());
}

static bool wait_until (const std::function<bool ()> &done, double seconds) {
	const auto t0 = std::chrono::steady_clock::now ();
	while (!done ()) {
		if (std::chrono::duration<double> (std::chrono::steady_clock::now () - t0). count () > seconds) return false;
		std::this_thread::sleep_for (std::chrono::microseconds (200));
	}
	return true;
}

extern "C" {

/* ficHandler::process_ficBlock over nframes frames: sym = [nframes][3][bitsPerBlock] int16 soft bits (symbols 1..3 of every frame);
 * bits_out [groups][768], crc_out [groups][3]; returns the number of 2304-bit groups decoded or -1 */
int ref_fic_frames (int bitsPerBlock, const int16_t *sym, int nframes, uint8_t *bits_out, uint8_t *crc_out) {
	RadioInterface mr;
	ficHandler *f = new ficHandler (&mr, (int16_t) bitsPerBlock);
	const int groups = nframes * (3 * bitsPerBlock / 2304);
	for (int fr = 0; fr < nframes; fr ++)
		for (int b = 1; b <= 3; b ++)
			f -> process_ficBlock (const_cast<int16_t *> (sym) + ((size_t) fr * 3 + (b - 1)) * bitsPerBlock, (int16_t) b);
	const bool ok = wait_until ([&] { std::lock_guard<std::mutex> l (g_fic_m); return g_fic [f]. calls >= 3 * groups; }, 60.0);
	int rc = -1;
	{
		std::lock_guard<std::mutex> l (g_fic_m);
		FicRec &r = g_fic [f];
		if (ok && (int) r. crc. size () == 3 * groups && (int) r. bits. size () == 768 * groups) {
			memcpy (bits_out, r. bits. data (), r. bits. size ());
			memcpy (crc_out, r. crc. data (), r. crc. size ());
			rc = groups;
		}
	}
	delete f;
	{ std::lock_guard<std::mutex> l (g_fic_m); g_fic. erase (f); }
	return rc;
}

/* fib_processor::process_FIB for every CRC-clean FIB of ngroups groups, in order (fic-handler.cpp:306-315), then the FIG 0/1
 * fields of ficList [64] (fib-processor.h:78-87); `list` is read first (the table carries over) */
int ref_fig01_scan (const uint8_t *bits, const uint8_t *crc_ok, int ngroups, orc_subch_info *list) {
	RadioInterface mr;
	fib_processor fp (&mr);
	for (int i = 0; i < 64; i ++) {
		fp. ficList [i]. StartAddr = list [i]. startAddr; fp. ficList [i]. Length = list [i]. length; fp. ficList [i]. uepFlag = list [i]. uepFlag;
		fp. ficList [i]. protLevel = list [i]. protLevel; fp. ficList [i]. BitRate = list [i]. bitRate;
	}
	for (int g = 0; g < ngroups; g ++) {
		/* the reference hands process_FIB a pointer INTO the 768-bit group buffer: a FIG whose length field runs past its FIB
		 * reads on into the following FIBs of the same group */
		uint8_t group [768 + 512];
		memset (group, 0, sizeof (group));
		memcpy (group, bits + (size_t) g * 768, 768);
		for (int k = 0; k < 3; k ++)
			if (crc_ok [3 * g + k]) fp. process_FIB (group + 256 * k, (uint16_t) (g & 3));
	}
	for (int i = 0; i < 64; i ++) {
		list [i]. startAddr = fp. ficList [i]. StartAddr; list [i]. length = fp. ficList [i]. Length; list [i]. uepFlag = fp. ficList [i]. uepFlag;
		list [i]. protLevel = fp. ficList [i]. protLevel; list [i]. bitRate = fp. ficList [i]. BitRate;
	}
	return 0;
}

/* mscHandler::process_mscBlock, symbol by symbol, behind it the reference's dabConcurrent (thread, ring buffer, 16-CIF time
 * de-interleaver, warm-up, EEP / UEP, energy dispersal): sym = [nframes][L - 1][2K] int16 soft bits as ofdmProcessor hands them
 * over; out receives the blocks delivered to addtoFrame ([nblocks][24 * bitRate]); returns nblocks or -1 */
int ref_msc_run (int mode, int L, int K, const int16_t *sym, int nframes, int startAddr, int length, int bitRate, int uepFlag, int protLevel,
                 uint8_t *out, int out_cap_blocks) {
	DabParams p;
	memset (&p, 0, sizeof (p));
	p. dabMode = (int16_t) mode; p. L = (int16_t) L; p. K = (int16_t) K;
	const int bpc = mode == 1 ? 18 : mode == 2 ? 72 : mode == 4 ? 36 : 18, cifs = nframes * ((L - 4) / bpc);
	RadioInterface mr;
	audioSink sink;
	mscHandler *m = new mscHandler (&mr, &p, &sink, 1);
	audiodata ad;
	memset (&ad, 0, sizeof (ad));
	ad. startAddr = (int16_t) startAddr; ad. length = (int16_t) length; ad. bitRate = (int16_t) bitRate; ad. uepFlag = (uint8_t) uepFlag;
	ad. protLevel = (int16_t) protLevel; ad. ASCTy = 0;                  /* not 077: plain DAB, the mp2Processor stand-in records */
	m -> set_audioChannel (&ad);
	const size_t symw = (size_t) 2 * K;
	std::vector<int16_t> zero (symw, 0);
	for (int fr = 0; fr < nframes; fr ++)
		for (int l = 4; l < L; l ++)
			m -> process_mscBlock (const_cast<int16_t *> (sym) + ((size_t) fr * (L - 1) + (l - 1)) * symw, (int16_t) l);
	/* dabConcurrent::run only takes a fragment when MORE than one is waiting (dab-concurrent.cpp:150): one extra CIF of zeros
	 * pushes the last real one through; it is never decoded itself */
	for (int l = 4; l < 4 + bpc; l ++) m -> process_mscBlock (zero. data (), (int16_t) l);
	const int expect = cifs > 16 ? cifs - 16 : 0;
	const bool ok = wait_until ([&] { return sink. nblocks () >= expect; }, 120.0);
	std::this_thread::sleep_for (std::chrono::milliseconds (5));          /* (a stray extra block would show up in the count) */
	m -> stop ();
	int rc = -1;
	{
		std::lock_guard<std::mutex> l (sink. m);
		if (ok && sink. blocks == expect && sink. blocks <= out_cap_blocks && (int) sink. bits. size () == expect * 24 * bitRate) {
			memcpy (out, sink. bits. data (), sink. bits. size ());
			rc = expect;
		}
	}
	delete m;
	return rc;
}

}

extern "C" {
/* dabSerial::process (dab-serial.cpp:110-147, the backend without a thread: warm-up of 15 CIFs where dabConcurrent has 16) over
 * ncif CIF fragments frags [ncif][fragmentSize]; out receives the blocks handed to addtoFrame; returns their number */
int ref_serial_run (const int16_t *frags, int ncif, int fragmentSize, int bitRate, int uepFlag, int protLevel, uint8_t *out, int out_cap_blocks) {
	RadioInterface mr;
	audioSink sink;
	dabSerial *b = new dabSerial (DAB /* plain DAB: the mp2Processor stand-in records */, (int16_t) fragmentSize, (int16_t) bitRate, (int16_t) uepFlag, (int16_t) protLevel, &mr, NULL, NULL, &sink);
	std::vector<int16_t> frag (fragmentSize);
	for (int c = 0; c < ncif; c ++) {                        /* process () de-interleaves IN PLACE: work on a copy */
		memcpy (frag. data (), frags + (size_t) c * fragmentSize, (size_t) fragmentSize * sizeof (int16_t));
		b -> process (frag. data (), (int16_t) fragmentSize);
	}
	const int nb = sink. blocks < out_cap_blocks ? sink. blocks : out_cap_blocks;
	memcpy (out, sink. bits. data (), (size_t) nb * 24 * bitRate);
	delete b;
	return sink. blocks;
}
}

extern "C" {
/* mp4Processor::addtoFrame (mp4processor.cpp:107-150, compiled unmodified: five-block window, Fire code, RS (120, 110) per column,
 * access-unit table, AU CRCs) CIF by CIF over bits [ncif][24 bitRate]; for every super frame it accepts: the corrected bytes
 * sf [n][110 bitRate / 8] and info [n] = { first_cif, -1 (the count of corrected symbols is a local of the reference), num_aus,
 * au_start [7], au_crc mask }.  Returns n. */
int ref_dabplus_run (const uint8_t *bits, int ncif, int bitRate, uint8_t *sf, orc_superframe_info *info, int max_sf) {
	RadioInterface mr;
	audioSink sink;
	mp4Processor *p = new mp4Processor (&mr, &sink, (int16_t) bitRate);
	const int nb = 24 * bitRate, sfb = 110 * (bitRate / 8);
	std::vector<uint8_t> row (nb);
	int n = 0;
	for (int c = 0; c < ncif; c ++) {
		memcpy (row. data (), bits + (size_t) c * nb, nb);
		const int before = p -> blocksInBuffer;
		const size_t aus_before = sink. aus. size ();
		p -> addtoFrame (row. data (), (int16_t) nb);
		if (before >= 4 && p -> blocksInBuffer == 0) {           /* accepted: Fire code held and processSuperframe returned true */
			if (n < max_sf) {
				memcpy (sf + (size_t) n * sfb, p -> outVector, sfb);
				orc_superframe_info &I = info [n];
				memset (&I, 0, sizeof (I));
				I. first_cif = c - 4; I. corrected = -1;
				const uint8_t *o = p -> outVector;
				const int kind = 2 * ((o [2] >> 6) & 1) + ((o [2] >> 5) & 1);
				I. num_aus = kind == 0 ? 4 : kind == 1 ? 2 : kind == 2 ? 6 : 3;
				for (int i = 0; i <= I. num_aus; i ++) I. au_start [i] = p -> au_start [i];
				size_t k = aus_before;                               /* the access units that passed their CRC, in order */
				for (int i = 0; i < I. num_aus && k < sink. aus. size (); i ++) {
					const int len = p -> au_start [i + 1] - p -> au_start [i] - 2;
					if (len >= 0 && (int) sink. aus [k]. size () == len && memcmp (sink. aus [k]. data (), o + p -> au_start [i], len) == 0) { I. au_crc |= 1 << i; k ++; }
				}
			}
			n ++;
		}
	}
	delete p;
	return n;
}
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The whole receive chain as the reference wires it (gui.cpp): ofdmProcessor (ofdm-processor.cpp, compiled unmodified:
 * null-symbol search, findIndex, NCO, coarse / fine AFC, the per-symbol loop) pulling samples from a virtualInput and pushing
 * soft bits into the reference's ficHandler and mscHandler.  The input is a memory buffer of raw u8 IQ converted as
 * rawFiles::getSamples does (rawfiles.cpp:113-116).  End of input: the reference then polls Samples () forever
 * (ofdm-processor.cpp:139-142, 186-190); the stand-in input notices the polling and the driver stops the processor.
 * --------------------------------------------------------------------------------------------------------------------- */
#define private public
#include "ofdm-processor.h"
#undef private

void ofdmProcessor::show_fineCorrector (int) {}
void ofdmProcessor::show_coarseCorrector (int) {}
void ofdmProcessor::show_avgTokenLength (int) {}
void ofdmProcessor::setSynced (char) {}
void ofdmProcessor::No_Signal_Found (void) {}

class memInput : public virtualInput {
public:
	memInput (const uint8_t *iq, int64_t n) : iq_ (iq), n_ (n), pos_ (0), polls_ (0) {}
	int32_t getSamples (DSPCOMPLEX *v, int32_t size) {
		polls_ = 0;
		int64_t k = n_ - pos_ < size ? n_ - pos_ : size;
		for (int64_t i = 0; i < k; i ++)
			v [i] = DSPCOMPLEX (float (iq_ [2 * (pos_ + i)] - 128) / 128.0, float (iq_ [2 * (pos_ + i) + 1] - 128) / 128.0);
		pos_ += k;
		return (int32_t) k;
	}
	int32_t Samples (void) { polls_ ++; const int64_t r = n_ - pos_; return (int32_t) (r > 0x7fffffff ? 0x7fffffff : r); }
	bool starved () const { return polls_. load () > 50; }   /* the processor keeps asking without taking: it wants more than is left */
	int64_t consumed () const { return pos_; }
private:
	const uint8_t *iq_; int64_t n_; std::atomic<int64_t> pos_; std::atomic<int> polls_;
};

extern "C" {

/* iq: nsamples u8 I,Q pairs.  One audio sub-channel.  fic_bits [max_groups][768], fic_crc [max_groups][3], msc [max_blocks][24 bitRate].
 * state_out [6] = coarseCorrector, fineCorrector, f2Correction, localPhase, samples taken from the input, 0.  Returns 0 or -1. */
int ref_receive (int mode, const uint8_t *iq, int64_t nsamples, int threshold, int freqSyncMethod,
                 int startAddr, int length, int bitRate, int uepFlag, int protLevel,
                 uint8_t *fic_bits, uint8_t *fic_crc, int max_groups, int *ngroups, uint8_t *msc, int max_blocks, int *nblocks, int64_t *state_out) {
	orc_params op;
	if (orc_mode_params (mode, &op)) return -1;
	DabParams p;
	memset (&p, 0, sizeof (p));
	p. dabMode = op. dabMode; p. L = op. L; p. K = op. K; p. T_null = op. T_null; p. T_F = op. T_F; p. T_s = op. T_s; p. T_u = op. T_u;
	p. guardLength = op. T_g; p. carrierDiff = op. carrierDiff;
	RadioInterface mr;
	audioSink sink;
	memInput in (iq, nsamples);
	mscHandler *msch = new mscHandler (&mr, &p, &sink, 1);
	audiodata ad;
	memset (&ad, 0, sizeof (ad));
	ad. startAddr = (int16_t) startAddr; ad. length = (int16_t) length; ad. bitRate = (int16_t) bitRate; ad. uepFlag = (uint8_t) uepFlag;
	ad. protLevel = (int16_t) protLevel; ad. ASCTy = 0;
	msch -> set_audioChannel (&ad);
	ficHandler *fich = new ficHandler (&mr, (int16_t) (2 * p. K));
	RingBuffer<DSPCOMPLEX> iqTap (32768);                    /* the GUI's constellation tap (ofdm-decoder.cpp:197-204): written, never read */
	ofdmProcessor *proc = new ofdmProcessor (&in, &p, &mr, msch, fich, (int16_t) threshold, &iqTap, (uint8_t) freqSyncMethod);
	const bool dbg = getenv ("ORC_TRACE") != NULL;
	if (dbg) fprintf (stderr, "[ref_receive] started\n");
	const bool fed = wait_until ([&] { return in. starved (); }, 600.0);
	if (dbg) fprintf (stderr, "[ref_receive] input exhausted (%d), consumed %lld\n", (int) fed, (long long) in. consumed ());
	/* everything the processor handed over is decoded before the handlers go: the FIC queue is empty when three calls per
	 * finished group were seen and no more arrive; the MSC ring buffer keeps its last fragment (dab-concurrent.cpp:150) */
	int last_f = -1, last_m = -1;
	wait_until ([&] {
		int f, m;
		{ std::lock_guard<std::mutex> l (g_fic_m); f = g_fic [fich]. calls; }
		m = sink. nblocks ();
		const bool same = f == last_f && m == last_m;
		last_f = f; last_m = m;
		if (!same) return false;
		std::this_thread::sleep_for (std::chrono::milliseconds (20));
		{ std::lock_guard<std::mutex> l (g_fic_m); f = g_fic [fich]. calls; }
		return f == last_f && sink. nblocks () == last_m && f % 3 == 0;
	}, 20.0);
	if (dbg) fprintf (stderr, "[ref_receive] handlers quiet: fic calls %d, msc blocks %d; fic queue filler %d reader %d free %d used %d index %d\n", last_f, last_m,
	                  (int) fich -> bufferFiller, (int) fich -> bufferReader, fich -> freeSlots. available (), fich -> usedSlots. available (), (int) fich -> index);
	state_out [0] = proc -> coarseCorrector; state_out [1] = proc -> fineCorrector; state_out [2] = proc -> f2Correction ? 1 : 0;
	state_out [3] = proc -> localPhase; state_out [4] = in. consumed (); state_out [5] = 0;
	delete proc;
	if (dbg) fprintf (stderr, "[ref_receive] processor gone\n");
	msch -> stop ();
	if (dbg) fprintf (stderr, "[ref_receive] backend stopped\n");
	int rc = fed ? 0 : -1;
	{
		std::lock_guard<std::mutex> l (g_fic_m);
		FicRec &r = g_fic [fich];
		const int g = (int) (r. crc. size () / 3);
		*ngroups = g < max_groups ? g : max_groups;
		memcpy (fic_bits, r. bits. data (), (size_t) *ngroups * 768);
		memcpy (fic_crc, r. crc. data (), (size_t) *ngroups * 3);
	}
	{
		std::lock_guard<std::mutex> l (sink. m);
		*nblocks = sink. blocks < max_blocks ? sink. blocks : max_blocks;
		memcpy (msc, sink. bits. data (), (size_t) *nblocks * 24 * bitRate);
	}
	delete fich;
	{ std::lock_guard<std::mutex> l (g_fic_m); g_fic. erase (fich); }
	delete msch;
	return rc;
}

}
