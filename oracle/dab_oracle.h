/*
 * dab_oracle.h -- C API of the CPU oracle for the DAB baseband decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Two shared libraries export exactly this API:
 *   oracle/_build/liboracle.so   plain-C restatement (dab_oracle.c + fft_standin.c), every function
 *                                citing the reference file:line it follows;
 *   oracle/_ref/libdabref.so     the same Tier-C loops (dab_oracle.c built with -DORC_USE_REF) but with
 *                                the Tier-A/B pieces routed into the UNMODIFIED reference classes compiled
 *                                from /root/reference (viterbi, eep/uep_deconvolve, protTables, mapper,
 *                                phasetable, phaseReference, ofdmDecoder, common_fft) -- see ref_shim/.
 * tests/test_oracle_vs_ref.py pins the restatement against the compiled reference.
 *
 * Parity status: Viterbi / depuncture / tables: pinned against the compiled reference (bit-exact).
 * FFT: the reference uses FFTW3f (external, absent) -- both builds use fft_standin.c, a labelled stand-in;
 * FFT-level parity is "parity unpinned" and tolerance based.
 */
#ifndef DAB_ORACLE_H
#define DAB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int32_t dabMode, L, K, T_null, T_F, T_s, T_u, T_g, carrierDiff;
	int32_t ficSymbols;      /* OFDM symbols carrying the FIC (3; the reference hard-codes 3) */
	int32_t ficGroups;       /* 2304-soft-bit FIC code words per frame                         */
	int32_t cifsPerFrame;    /* CIFs per frame                                                 */
	int32_t blocksPerCIF;    /* OFDM symbols per CIF                                           */
} orc_params;

/* ---- tables (gui.cpp:1328-1372, mapper.cpp:33-117, phasetable.cpp, protTables.cpp) ---- */
int   orc_mode_params (int mode, orc_params *p);
int   orc_perm_table (int mode, int16_t *out /* K */);
float orc_phi (int mode, int k);
int   orc_ref_table (int mode, float *out /* 2*T_u, interleaved re,im */);
int   orc_pcode (int n /* 1..24 */, int8_t *out /* 32 */);
/* returns 0 and fills L[4], PI[4] (PI[3] == 0 when unused) or -1 for an unknown profile */
int   orc_uep_profile (int bitRate, int protLevel, int16_t L[4], int16_t PI[4]);
int   orc_eep_profile (int bitRate, int protLevel, int16_t L[2], int16_t PI[2]);

/* ---- channel decoding ---- */
/* viterbi::deconvolve (viterbi.cpp:225-242): soft[4*(frameBits+6)] -> out[frameBits], 1 bit / byte */
int orc_viterbi (int frameBits, const int16_t *soft, uint8_t *out);
/* eep_/uep_deconvolve::deconvolve (deconvolve.cpp:186-237, 338-366) */
int orc_eep_deconvolve (int bitRate, int protLevel, const int16_t *v, int32_t size, uint8_t *out);
int orc_uep_deconvolve (int bitRate, int protLevel, const int16_t *v, int32_t size, uint8_t *out);
/* ficHandler::process_ficInput (fic-handler.cpp:241-321) minus the fib_processor hand-over:
 * soft2304 -> bits768 (after PRBS), crc_ok[3].  bits768 are NOT modified by the CRC check. */
int orc_fic_decode (const int16_t *soft2304, uint8_t *bits768, uint8_t *crc_ok);
/* check_CRC_bits (dab-constants.h:310-340) on a copy of in[size] */
int orc_check_crc (const uint8_t *in, int size);
/* energy-dispersal PRBS (fic-handler.cpp:100-108) */
void orc_prbs (uint8_t *out, int n);
/* dabConcurrent::run (dab-concurrent.cpp:144-193) over ncif consecutive CIF fragments:
 * in[ncif][fragmentSize] -> out[max(ncif-16,0)][24*bitRate]; returns the number of output blocks.
 * uepFlag follows the reference's inverted naming: 0 = UEP, otherwise EEP. */
int orc_msc_backend (const int16_t *in, int ncif, int fragmentSize, int bitRate, int uepFlag,
                     int protLevel, uint8_t *out);
/* time de-interleaver only (dab-concurrent.cpp:41-43, 162-169): out[t][i] = in[t-D[i&15]][i] with zero history */
void orc_time_deinterleave (const int16_t *in, int ncif, int fragmentSize, int16_t *out);

/* ---- frame level regrouping ---- */
/* soft bits of nframes frames: sym[frame][L-1][2K] (symbols 1..L-1) */
/* ficHandler::run regroup (fic-handler.cpp:192-230) + process_ficInput per group */
/* DAB+ super-frame layer (mp4processor.cpp:107-275): Fire code sync over 5 CIFs, RS(120,110) per column, AU table + CRCs */
typedef struct orc_dabplus orc_dabplus;
typedef struct { int64_t first_cif; int32_t corrected, num_aus, au_start [7], au_crc; } orc_superframe_info;
orc_dabplus *orc_dabplus_new (int bitRate);
void orc_dabplus_free (orc_dabplus *);
/* bits[ncif][24*bitRate] (1 bit per byte) -> superframes[n][110*bitRate/8], info[n]; returns n (may exceed max_sf: then truncated) */
int  orc_dabplus_process (orc_dabplus *, const uint8_t *bits, int ncif, uint8_t *superframes, orc_superframe_info *info, int max_sf);
/* primitives, exposed for the tests: firecode_checker::check, reedSolomon::dec / enc (cutlen 135) */
int  orc_firecode_check (const uint8_t *x11);
int  orc_rs_dec (const uint8_t *r120, uint8_t *d110);
void orc_rs_enc (const uint8_t *d110, uint8_t *r120);

/* airspyHandler's rate conversion (airspy-handler.cpp:138-148, 342-370): int16 I,Q at selectedRate -> floats at 2.048 MS/s */
int64_t orc_resample_i16 (const int16_t *sbuf, int64_t nSamples, int selectedRate, float *out);

/* ficList entry (fib-processor.h:78-87), the fields FIG 0/1 fills */
typedef struct { int32_t defined, startAddr, length, uepFlag, protLevel, bitRate; } orc_subch_info;
/* FIG 0/1 of every CRC-clean FIB of ngroups FIC groups, in order, into list[64] (fib-processor.cpp:123-158, 278-347) */
int orc_fig01_scan (const uint8_t *bits, const uint8_t *crc_ok, int ngroups, orc_subch_info *list);
int orc_fic_frames (int mode, const int16_t *sym, int nframes, uint8_t *bits /* [nframes*groups][768] */,
                    uint8_t *crc_ok /* [nframes*groups][3] */);
/* mscHandler::process_mscBlock (msc-handler.cpp:125-193): CIF assembly and sub-channel slice
 * -> frag[nframes*cifsPerFrame][Length*64] */
int orc_msc_slice (int mode, const int16_t *sym, int nframes, int startAddr, int Length, int16_t *frag);

/* ---- OFDM front end ---- */
typedef struct orc_ofdm orc_ofdm;
orc_ofdm *orc_ofdm_new (int mode, int threshold, int freqSyncMethod);
void      orc_ofdm_free (orc_ofdm *);
/* common_fft::do_FFT / common_ifft::do_IFFT (fft.cpp:53-55, 109-112) on n complex floats, in place */
int  orc_fft (float *v, int n, int inverse);
/* phaseReference::findIndex (phasereference.cpp:60-88) */
int32_t orc_find_index (orc_ofdm *, const float *v /* 2*T_u */);
/* ofdmDecoder::processBlock_0 (ofdm-decoder.cpp:85-162) */
int  orc_block0 (orc_ofdm *, const float *v /* 2*T_u */, int flag);
/* ofdmDecoder::processToken (ofdm-decoder.cpp:167-207) */
void orc_token (orc_ofdm *, const float *inv /* 2*T_s */, int16_t *ibits /* 2K */);
/* copy of the decoder's current phaseReference[T_u] (tests) */
void orc_get_phase_reference (orc_ofdm *, float *out /* 2*T_u */);

typedef struct {
	int64_t pos;          /* absolute sample index where SyncOnPhase started reading        */
	int32_t startIndex;   /* findIndex result                                               */
	int32_t coarse;       /* coarseCorrector while the data symbols were read               */
	int32_t fine;         /* fineCorrector while the data symbols were read                 */
	int32_t phase0;       /* localPhase before the first sample of SyncOnPhase              */
	int32_t correction;   /* processBlock_0 return value                                    */
	float   freqCorrRe, freqCorrIm;  /* FreqCorr accumulated over the frame                 */
} orc_frame_info;

/* ofdmProcessor::run (ofdm-processor.cpp:247-474) restated over an in-memory u8 IQ "rawfile"
 * (rawfiles.cpp:113-116).  Decodes at most max_frames frames; sym[frame][L-1][2K] receives the soft
 * bits handed to process_ficBlock / process_mscBlock.  Returns frames decoded. */
int orc_ofdm_run (int mode, int threshold, int freqSyncMethod, const uint8_t *iq, int64_t nsamples,
                  int max_frames, int16_t *sym, orc_frame_info *info);
/* the same loop fed with complex float samples (interleaved re, im), the form every input device of the reference
 * delivers (virtual-input.h:62-63; wavfiles.cpp:168-180 for .sdr files) */
int orc_ofdm_run_cf32 (int mode, int threshold, int freqSyncMethod, const float *iq, int64_t nsamples,
                       int max_frames, int16_t *sym, orc_frame_info *info);

const char *orc_build_kind (void);   /* "port" or "reference" */

#ifdef __cplusplus
}
#endif
#endif
