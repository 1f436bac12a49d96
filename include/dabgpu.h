/*
 * dabgpu.h -- C ABI of the B200-native DAB baseband decode engine (libdabgpu.so).
 *
 * This is the drop-in boundary for the reference's decode hot path (AlbrechtL/sdr-j-dab 0.997).  The
 * reference has no plugin/FFI mechanism: its "API" is C++ class composition (gui.cpp:160-179), so every
 * entry point below names the reference member function it replaces (file:line relative to the reference
 * tree).  sdr-j-dab_b200/host/dab_adapters.h wraps these calls in classes with the reference's names and
 * signatures; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; buffers are caller owned HOST memory unless the name ends in _dev
 *     (then they are device pointers on the handle's device and the call is asynchronous on the handle's
 *     stream until dabgpu_sync);
 *   - every function returns 0 on success or a negative dabgpu_status; text via dabgpu_last_error();
 *     nothing throws across the ABI; there is NO CPU fallback: without a usable CUDA device
 *     dabgpu_create fails with DABGPU_ERR_CUDA;
 *   - a handle is not re-entrant (like the reference's viterbi objects, viterbi.h:48-66); several handles
 *     may be used concurrently;
 *   - soft bits: int16 in [-127,127], positive = bit 1, 0 = erasure (viterbi.cpp:229-235);
 *     decoded bits: one bit per uint8 (viterbi.cpp:240-241).
 */
#ifndef DABGPU_H
#define DABGPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dabgpu dabgpu_t;

typedef enum {
	DABGPU_OK          =  0,
	DABGPU_ERR_ARG     = -1,   /* bad argument / unsupported parameter combination      */
	DABGPU_ERR_CUDA    = -2,   /* CUDA runtime error (see dabgpu_last_error)             */
	DABGPU_ERR_PROFILE = -3,   /* unknown UEP/EEP protection profile                     */
	DABGPU_ERR_NOMEM   = -4,
	DABGPU_ERR_STATE   = -5    /* call sequence error (e.g. token before block0)         */
} dabgpu_status;

typedef struct {
	int32_t device;            /* CUDA device ordinal                                              */
	int32_t dabMode;           /* 1..4 (DabParams.dabMode, gui.cpp:1328-1372)                      */
	int32_t threshold;         /* phaseReference level, default 3 (gui.cpp:98-99)                  */
	int32_t freqSyncMethod;    /* 0,1,2 as ofdmDecoder (main.cpp:91 default 1)                     */
	int32_t viterbi_path;      /* 0 = auto (by batch size), 1 = warp-per-code-word kernel, 2 = throughput (two threads per code word) kernel */
	int32_t host_batch_frames; /* dabgpu_decode from HOST memory: frames per channel-decoding launch while the input is
	                              still arriving over PCIe (0 = default 128)                                           */
	int32_t reserved [2];      /* tuning / A-B switches, 0 = default: [0] = 1 forces the generic symbol kernel in Mode I,
	                              [1] = frames per channel-decoding launch on the device-resident path (default: one
	                              launch per call, measured fastest) */
} dabgpu_config;

/* one MSC sub-channel, the fields of audiodata/packetdata the decode path uses (dab-constants.h:151-175;
 * msc-handler.cpp:91-105).  uepFlag keeps the reference's inverted meaning: 0 = UEP table, else EEP;
 * protLevel for EEP is 0100+level (A) / 0200+level (B) (fib-processor.cpp:313,330). */
typedef struct {
	int32_t startAddr;         /* first capacity unit (CU = 64 soft bits)                          */
	int32_t length;            /* size in CUs; fragmentSize = length * 64                          */
	int32_t bitRate;           /* kbit/s; decoded block = 24 * bitRate bits per CIF                */
	int32_t uepFlag;
	int32_t protLevel;
} dabgpu_subch;

int  dabgpu_create (const dabgpu_config *cfg, dabgpu_t **out);
void dabgpu_destroy (dabgpu_t *h);
const char *dabgpu_last_error (const dabgpu_t *h);      /* h may be NULL: error of the last failed create */
int  dabgpu_sync (dabgpu_t *h);                          /* wait for the handle's stream                   */
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t dabgpu_launch_count (const dabgpu_t *h);
/* CUDA-event stopwatch on the handle's stream: begin records an event, end records a second one, waits for
 * it and returns the device time between them in milliseconds */
int  dabgpu_timer_begin (dabgpu_t *h);
int  dabgpu_timer_end (dabgpu_t *h, float *ms);

/* optional per-kernel-class CUDA-event profile (bench.py's roofline numbers): when enabled every kernel
 * launch is bracketed by two events on the handle's stream; get() waits for the stream and returns the
 * accumulated launch count and device milliseconds of one class since the last reset */
enum { DABGPU_KC_ACQUIRE = 0, DABGPU_KC_FRONT, DABGPU_KC_SYMBOL, DABGPU_KC_SCAN, DABGPU_KC_VITERBI_MSC,
       DABGPU_KC_VITERBI_FIC, DABGPU_KC_VITERBI_API, DABGPU_KC_CRC, DABGPU_KC_VITERBI_TB, DABGPU_KC_VITERBI_SYM, DABGPU_KC_COUNT };
int  dabgpu_profile_enable (dabgpu_t *h, int32_t on);
int  dabgpu_profile_reset (dabgpu_t *h);
int  dabgpu_profile_get (dabgpu_t *h, int32_t kernel_class, int64_t *launches, double *ms);

/* integer-pipe peak of this GPU, measured live (the roofline denominator of the Viterbi group, which
 * MEASURED_PEAKS.json does not hold): ops[0] = add only, ops[1] = min only, ops[2] = add + mad.lo mix,
 * ops[3] = packed 16x2 add, ops[4] = packed 16x2 min, ops[5] = packed min + predicates + two predicated ORs
 * (one count per group); ops must hold 6 doubles; unit = instructions x 32-bit lanes per second */
int  dabgpu_int_peak (dabgpu_t *h, double *ops);

/* ------------------------------------------------------------------------------------------------
 * Channel decoding (Viterbi group)
 * ---------------------------------------------------------------------------------------------- */
/* viterbi::deconvolve (viterbi.cpp:225-242) on nblocks independent terminated code words:
 * soft[nblocks][4*(frameBits+6)] -> bits[nblocks][frameBits] */
int dabgpu_viterbi (dabgpu_t *h, const int16_t *soft, int32_t frameBits, int32_t nblocks, uint8_t *bits);
int dabgpu_viterbi_dev (dabgpu_t *h, const int16_t *soft, int32_t frameBits, int32_t nblocks, uint8_t *bits);

/* eep_deconvolve::deconvolve / uep_deconvolve::deconvolve (deconvolve.cpp:338-366 / 186-237): depuncture +
 * Viterbi, NO energy dispersal.  v[nblocks][size] -> bits[nblocks][24*bitRate]; size >= the number of
 * punctured bits of the profile (the reference ignores it, deconvolve.cpp:184). */
int dabgpu_protect_decode (dabgpu_t *h, int32_t bitRate, int32_t uepFlag, int32_t protLevel,
                           const int16_t *v, int32_t size, int32_t nblocks, uint8_t *bits);

/* ficHandler::process_ficInput (fic-handler.cpp:241-321) on ngroups 2304-soft-bit FIC code words:
 * depuncture (21xPI_16, 3xPI_15, PI_X), Viterbi(768), PRBS, CRC of the three FIBs.
 * soft[ngroups][2304] -> bits[ngroups][768], crc_ok[ngroups][3] (crc_ok may be NULL) */
int dabgpu_fic_decode (dabgpu_t *h, const int16_t *soft, int32_t ngroups, uint8_t *bits, uint8_t *crc_ok);

/* fib_processor::process_FIB restricted to what the MSC decoder needs: FIG 0/1, the sub-channel organisation
 * (fib-processor.cpp:123-158 process_FIB, 163-177 process_FIG0, 278-347 FIG0Extension1 / HandleFIG0Extension1),
 * applied to every FIB whose CRC holds (fic-handler.cpp:309-319), in order; table[SubChId] = the ficList entry
 * (fib-processor.h:78-87).  uepFlag / protLevel use the reference's encoding, so an entry can be handed to
 * dabgpu_set_subchannels as it is.  fic_bits[ngroups][768], crc_ok[ngroups][3] as dabgpu_fic_decode returns them. */
typedef struct { int32_t defined, startAddr, length, uepFlag, protLevel, bitRate; } dabgpu_subch_info;
int dabgpu_fig01_scan (dabgpu_t *h, const uint8_t *fic_bits, const uint8_t *crc_ok, int32_t ngroups, dabgpu_subch_info table [64]);
/* the running table of the stream engine: every FIB a dabgpu_decode call decodes is scanned on the device */
int dabgpu_get_subch_table (dabgpu_t *h, dabgpu_subch_info table [64]);

/* dabConcurrent::run (dab-concurrent.cpp:144-193) for one sub-channel over ncif consecutive CIF fragments:
 * 16-CIF time de-interleaving, 16-CIF warm-up skip, eep/uep depuncture + Viterbi, energy dispersal.
 * A backend object carries the de-interleaver history across calls like the reference object does.
 * frags[ncif][fragmentSize] -> out[*nout][24*bitRate]; *nout = blocks produced (ncif minus what the warm-up
 * swallowed). */
typedef struct dabgpu_backend dabgpu_backend_t;
int  dabgpu_backend_create (dabgpu_t *h, const dabgpu_subch *sc, dabgpu_backend_t **out);
void dabgpu_backend_destroy (dabgpu_backend_t *b);
int  dabgpu_backend_process (dabgpu_backend_t *b, const int16_t *frags, int32_t ncif, uint8_t *out, int32_t *nout);
/* the de-interleaver state (the multi-GPU halo): the last 15 fragments seen plus the warm-up counter.
 * hist[15][fragmentSize], oldest first. */
int  dabgpu_backend_get_state (dabgpu_backend_t *b, int16_t *hist, int32_t *cifs_seen);
int  dabgpu_backend_set_state (dabgpu_backend_t *b, const int16_t *hist, int32_t cifs_seen);

/* DAB+ super-frame layer, the stage right behind the MSC decoder for every DAB+ service: mp4Processor::addtoFrame
 * and the Fire-code / Reed-Solomon / access-unit-table part of processSuperframe (mp4processor.cpp:107-150,
 * 152-275; firecode-checker.cpp:77-95; reed-solomon.cpp:129-399 with galois.cpp; dabPlus_crc mp4processor.cpp:40-61).
 * An object carries the 5-CIF window and the sync counter across calls like the reference object does.
 * bits[ncif][24*bitRate] = the blocks dabgpu_backend_process / dabgpu_decode return for the sub-channel;
 * superframes[*nsf][110*bitRate/8] repaired super frames in order (at most max_sf are copied out), info[i] =
 * first CIF (counted from the object's creation), RS symbol errors corrected, the access-unit table
 * au_start[0..num_aus] and the per-AU CRC verdicts (bit i = AU i passes).  The AAC decoder is out of scope. */
typedef struct dabgpu_dabplus dabgpu_dabplus_t;
typedef struct { int64_t first_cif; int32_t corrected, num_aus, au_start [7], au_crc; } dabgpu_superframe_info;
int  dabgpu_dabplus_create (dabgpu_t *h, int32_t bitRate, dabgpu_dabplus_t **out);
void dabgpu_dabplus_destroy (dabgpu_dabplus_t *d);
int  dabgpu_dabplus_process (dabgpu_dabplus_t *d, const uint8_t *bits, int32_t ncif, uint8_t *superframes,
                             dabgpu_superframe_info *info, int32_t max_sf, int32_t *nsf);
/* same with the bits already on the handle's device (e.g. straight from the MSC decoder's output buffer) */
int  dabgpu_dabplus_process_dev (dabgpu_dabplus_t *d, const uint8_t *d_bits, int32_t ncif, uint8_t *superframes,
                                 dabgpu_superframe_info *info, int32_t max_sf, int32_t *nsf);

/* ------------------------------------------------------------------------------------------------
 * OFDM front end (FFT + demod group), per-call parity entry points
 * ---------------------------------------------------------------------------------------------- */
/* phaseReference::findIndex (phasereference.cpp:60-88) on n windows v[n][T_u] (interleaved re,im):
 * idx[i] = index of the correlation peak, or the reference's negative code when below threshold */
int dabgpu_find_index (dabgpu_t *h, const float *v, int32_t n, int32_t *idx);
/* ofdmDecoder::processBlock_0 (ofdm-decoder.cpp:85-162): v[T_u] is the PRS block; sets the handle's phase
 * reference; *correction = coarse offset in carriers (100 = not found; 0 when flag == 0) */
int dabgpu_block0 (dabgpu_t *h, const float *v, int32_t flag, int16_t *correction);
/* ofdmDecoder::processToken (ofdm-decoder.cpp:167-207) on nsym CONSECUTIVE symbols inv[nsym][T_s]:
 * ibits[nsym][2K]; the phase reference advances symbol by symbol as in the reference */
int dabgpu_token (dabgpu_t *h, const float *inv, int32_t nsym, int16_t *ibits);
/* copy of the current phaseReference[T_u] (tests) */
int dabgpu_get_phase_reference (dabgpu_t *h, float *out);
/* common_fft::do_FFT / common_ifft::do_IFFT (fft.cpp:53-55, 109-112) on n vectors of T_u complex floats */
int dabgpu_fft (dabgpu_t *h, float *v, int32_t n, int32_t inverse);

/* ------------------------------------------------------------------------------------------------
 * Batch decode: ofdmProcessor::run + ficHandler + mscHandler/dabConcurrent over a block of u8 IQ
 * (ofdm-processor.cpp:247-474; rawfiles.cpp:113-116 sample format).  The handle is a stream: sync/AFC
 * state, the unconsumed sample tail and the de-interleaver histories persist between calls.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {               /* per decoded frame, what the reference shows on its GUI + replay state */
	int64_t pos;               /* absolute sample index where SyncOnPhase started reading               */
	int32_t startIndex;        /* findIndex result                                                      */
	int32_t coarse, fine;      /* correctors while the data symbols were read                           */
	int32_t phase0;            /* NCO phase index before the first SyncOnPhase sample                   */
	int32_t correction;        /* processBlock_0 result                                                 */
	float   freqCorrRe, freqCorrIm;
} dabgpu_frame_info;

typedef struct {
	/* capacities (in), set by the caller */
	int32_t max_frames;
	/* outputs; any pointer may be NULL to skip that output */
	int32_t nframes;           /* frames decoded by this call                                           */
	dabgpu_frame_info *info;   /* [max_frames]                                                          */
	int16_t *soft;             /* [max_frames][L-1][2K] soft bits as handed to process_fic/mscBlock     */
	uint8_t *fic_bits;         /* [max_frames*ficGroups][768]                                           */
	uint8_t *fic_crc;          /* [max_frames*ficGroups][3]                                             */
	uint8_t **msc_bits;        /* [nsub] -> [max_frames*cifsPerFrame][24*bitRate]                       */
	int32_t *msc_nblocks;      /* [nsub] blocks written per sub-channel                                 */
	int64_t consumed;          /* samples of this call's input consumed (rest is kept in the handle)    */
} dabgpu_result;

int dabgpu_set_subchannels (dabgpu_t *h, const dabgpu_subch *sc, int32_t nsub);   /* set_audioChannel x nsub */
/* packed != 0: the stream engine (dabgpu_decode*, dabgpu_decode_multi, dabgpu_group_*) delivers MSC blocks with EIGHT bits
 * per byte, first bit in the top bit of byte 0 -- msc_bits[i] then holds [blocks][3*bitRate] bytes.  The reference hands
 * dabProcessor::addtoFrame one bit per byte (dab-concurrent.cpp:191) and the DAB+ processor packs them exactly this
 * way as its first step (mp4Processor::addtoFrame, audio/mp4processor.cpp:107-117); packing on the device cuts the device-to-host
 * traffic of the decoded bits by 8.  FIC bits stay one per byte (fib_processor::process_FIB reads them that way). */
int dabgpu_set_msc_output (dabgpu_t *h, int32_t packed);
int dabgpu_decode (dabgpu_t *h, const uint8_t *iq_u8, size_t nsamples, dabgpu_result *out);
/* Streaming from host memory: announces the block the NEXT dabgpu_decode / _cf32 / _i16 call of this handle will be given (same
 * pointer, same size, same format: 0 = u8, 1 = complex float, 2 = int16) and starts its host-to-device copy at once, so that it
 * runs while the call in between is still busy with its last frames -- the reference's reader thread does the same for its
 * consumer (rawfiles.cpp:136-161 fills the ring buffer ahead of ofdmProcessor).  Up to two blocks may be announced; call
 * order for blocks B0, B1, ...:  prefetch (B0); prefetch (B1); decode (B0); prefetch (B2); decode (B1); ...  An announced
 * buffer must stay unchanged until its decode call returns.  A decode call whose block was not announced uploads it itself (behind
 * whatever is already on its way).  Pinned host memory makes the copy asynchronous; pageable memory is staged (the staging memcpy
 * happens inside dabgpu_prefetch).  Returns DABGPU_ERR_STATE when two blocks are announced already. */
int dabgpu_prefetch (dabgpu_t *h, const void *iq, size_t nsamples, int32_t sample_format);
/* same with the input already resident on the handle's device (result pointers stay host pointers) */
int dabgpu_decode_dev (dabgpu_t *h, const uint8_t *d_iq_u8, size_t nsamples, dabgpu_result *out);
/* The same stream engine fed with complex float samples (interleaved re, im at 2.048 MS/s): the form in which
 * every input device of the reference hands samples to ofdmProcessor (virtualInput::getSamples, virtual-input.h:62-63;
 * wavfiles.cpp:168-180 for .sdr recordings; the dabstick / sdrplay / airspy handlers).  A stream may switch format
 * only when no unconsumed samples are pending (DABGPU_ERR_STATE otherwise). */
int dabgpu_decode_cf32 (dabgpu_t *h, const float *iq, size_t nsamples, dabgpu_result *out);
int dabgpu_decode_cf32_dev (dabgpu_t *h, const float *d_iq, size_t nsamples, dabgpu_result *out);
/* 16-bit PCM samples (interleaved I,Q at 2.048 MS/s), the payload of a 16-bit ".sdr" / WAV recording as wavFiles
 * reads it (wavfiles.cpp:186-197): libsndfile's sf_readf_float delivers x / 32768 (third-party, not vendored by the
 * reference: README "libsndfile"; the normalisation is libsndfile's documented default for 16-bit PCM).  The division is
 * exact in float, so the result equals dabgpu_decode_cf32 on the converted samples bit for bit; the conversion happens
 * in the kernels' sample fetch, no float copy of the recording is made. */
int dabgpu_decode_i16 (dabgpu_t *h, const int16_t *iq, size_t nsamples, dabgpu_result *out);
int dabgpu_decode_i16_dev (dabgpu_t *h, const int16_t *d_iq, size_t nsamples, dabgpu_result *out);
/* Many independent streams in one call (BASELINE configs[3]: short recordings of different ensembles).  The reference
 * runs one ofdmProcessor / ficHandler / mscHandler chain per stream (gui.cpp:160-179); its sample-serial acquisition
 * (ofdm-processor.cpp:275-338) and the frame-by-frame AFC convergence (:390-405, 445-466) are independent across streams,
 * so the engine runs the streams side by side: the null-symbol searches asynchronously (every stream joins the decoding
 * rounds the moment ITS search ends), the streams in sync through the same kernel launches round by round.
 * Stream i is decoded exactly as a FRESH handle with this handle's configuration and sub-channels would decode it with
 * one dabgpu_decode call (acquisition from the first sample, coarse search on, empty de-interleaver): jobs[i].out is
 * filled like dabgpu_decode fills it (out->consumed = samples of the stream consumed).  The handle's own stream state
 * is not touched.  sample_format: 0 = u8 I,Q (rawfiles.cpp:113-116), 1 = complex float, 2 = int16 I,Q.
 * Host input is uploaded in interleaved pieces and decoded as it arrives.  Result buffers in PINNED host memory (all of a
 * stream's fic_bits / fic_crc / msc_bits pointers, soft == NULL) are filled by asynchronous copies while later rounds run;
 * pageable ones are filled from a staging buffer at the end of the call.
 * _dev: jobs[i].iq are device pointers on the handle's device (result pointers stay host pointers). */
typedef struct {
	const void *iq;            /* nsamples complex samples of the given format                                     */
	size_t nsamples;
	dabgpu_result *out;
} dabgpu_stream_job;
int dabgpu_decode_multi (dabgpu_t *h, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format);
int dabgpu_decode_multi_dev (dabgpu_t *h, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format);
/* ------------------------------------------------------------------------------------------------
 * Several GPUs of one box: one process, one engine handle + one host thread per device.  The reference
 * has no counterpart (one ofdmProcessor thread per receiver, gui.cpp:160-179); what crosses a shard boundary is the
 * reference's sequential state: the sync / AFC variables of ofdmProcessor::run (ofdm-processor.cpp:445-466) and the 15
 * CIFs the time de-interleaver remembers (dab-concurrent.cpp:41-43, 162-175).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dabgpu_group dabgpu_group_t;
int32_t dabgpu_device_count (void);                                        /* usable CUDA devices (0 without a driver / device) */
/* one handle per entry of devices[] (NULL: devices 0 .. ndev-1; the same device may appear more than once), all with cfg */
int  dabgpu_group_create (const dabgpu_config *cfg, const int32_t *devices, int32_t ndev, dabgpu_group_t **out);
void dabgpu_group_destroy (dabgpu_group_t *g);
const char *dabgpu_group_last_error (const dabgpu_group_t *g);
int32_t dabgpu_group_size (const dabgpu_group_t *g);
dabgpu_t *dabgpu_group_handle (dabgpu_group_t *g, int32_t i);             /* member i (owned by the group) */
int  dabgpu_group_set_subchannels (dabgpu_group_t *g, const dabgpu_subch *sc, int32_t nsub);
/* independent streams (BASELINE configs[3]): stream i runs on member i mod n through dabgpu_decode_multi, no communication */
int  dabgpu_group_decode_multi (dabgpu_group_t *g, const dabgpu_stream_job *jobs, int32_t nstreams, int32_t sample_format);
/* ONE recording (BASELINE configs[4]) decoded from its first sample by all members; `out` is filled exactly as ONE fresh
 * handle's dabgpu_decode over the whole recording would fill it (bit for bit), out->consumed = samples consumed.
 * scheme 1 (default choice): GPU 0 decodes lead_frames frames (<= 0: 24) to lock, every member then decodes its
 * contiguous frame range IN PARALLEL from the closed-form predicted tracking state (dabgpu_host_state_predict),
 * starting 16 CIFs early instead of receiving a de-interleaver halo; every boundary is verified against the left
 * neighbour's true final state, and on any disagreement (or an unlocked receiver) the call falls back to
 * scheme 0: the exact serial chain -- member r decodes its sample range, then the whole stream state (sync / AFC
 * variables, unconsumed samples, 15-CIF soft-bit halo) moves to member r + 1 device to device (cudaMemcpyPeerAsync).
 * *scheme_used (may be NULL) reports which one produced the output. */
int  dabgpu_group_decode (dabgpu_group_t *g, const uint8_t *iq_u8, size_t nsamples, dabgpu_result *out,
                          int32_t lead_frames, int32_t scheme, int32_t *scheme_used);

/* airspyHandler's sample-rate conversion (airspy-handler.cpp:138-148, 342-370): int16 I,Q pairs at in_rate samples/s
 * (a multiple of 1000) -> complex floats at 2 048 000 samples/s by linear interpolation in 1 ms blocks, ready for
 * dabgpu_decode_cf32.  Block b reads input samples [b R, b R + R], R = in_rate / 1000, and writes 2048 samples; a call
 * converts the (n_in - 1) / R whole blocks it holds: *n_out = samples written to out (capacity needed:
 * ((n_in - 1) / R) * 2048 complex floats), *consumed = input samples the caller may drop (it keeps the rest,
 * including the sample shared with the next block, for its next call). */
int dabgpu_resample_i16 (dabgpu_t *h, const int16_t *iq, size_t n_in, int32_t in_rate, float *out, size_t *n_out, size_t *consumed);
int dabgpu_resample_i16_dev (dabgpu_t *h, const int16_t *d_iq, size_t n_in, int32_t in_rate, float *d_out, size_t *n_out, size_t *consumed);
int dabgpu_reset (dabgpu_t *h);                                                   /* ofdmProcessor::reset  */
/* ofdmProcessor::coarseCorrectorOn (on != 0: coarse search on, coarseCorrector = 0) / coarseCorrectorOff
 * (ofdm-processor.cpp:499-506); takes effect with the next frame decoded */
int dabgpu_coarse_corrector (dabgpu_t *h, int32_t on);

/* stream state for splitting a recording across calls / GPUs */
typedef struct {
	int32_t synced;            /* 0 = acquisition (notSynced), 1 = tracking (SyncOnPhase)               */
	int32_t coarse, fine, f2Correction, previous_1, previous_2, localPhase;
	int64_t abs_pos;           /* absolute index of the next sample to read                             */
	int64_t frames, cifs;
} dabgpu_stream_state;
int dabgpu_state_get (dabgpu_t *h, dabgpu_stream_state *s);
int dabgpu_state_set (dabgpu_t *h, const dabgpu_stream_state *s);
/* the complete stream state as one opaque blob: sync/AFC state, the unconsumed sample tail, the 15-CIF soft-bit
 * halo of the time de-interleaver (dab-concurrent.cpp:41-43, 162-175) and the warm-up counters of the configured
 * sub-channels.  This is what one GPU hands to the next when a recording is split (<= ~2.1 MB).
 * export with buf == NULL returns the size needed in *used. */
int dabgpu_state_export (dabgpu_t *h, void *buf, size_t capacity, size_t *used);
int dabgpu_state_import (dabgpu_t *h, const void *buf, size_t n);

/* The tracking state `nframes` frames further on for a LOCKED receiver, in closed form (host only, no GPU): every frame
 * is T_F samples long (findIndex returns T_g), the coarse search is off and the correctors do not move, so
 * abs_pos advances by nframes * T_F and localPhase by -(nframes * T_F) * (coarse + fine) mod 2048000
 * (ofdm-processor.cpp:217-226, 344-474 in the steady state).  This is what lets the shards of ONE recording start
 * in parallel on several GPUs: each shard assumes the predicted state, the assumption is verified afterwards
 * against the previous shard's true final state (sdr-j-dab_b200/parallel.py: decode_sharded).
 * Returns DABGPU_ERR_STATE when `in` is not a locked state (synced == 0 or coarse search still on). */
int dabgpu_host_state_predict (int32_t mode, const dabgpu_stream_state *in, int64_t nframes, dabgpu_stream_state *out);

/* ------------------------------------------------------------------------------------------------
 * Host-only helpers (no GPU needed): the constant tables the engine derives on the host, exported so the
 * CPU test-suite can compare them with the oracle.
 * ---------------------------------------------------------------------------------------------- */
/* out = {dabMode, L, K, T_null, T_F, T_s, T_u, T_g, carrierDiff, ficGroups, cifsPerFrame, blocksPerCIF} (gui.cpp:1328-1372) */
int dabgpu_host_mode_params (int32_t mode, int32_t out [12]);
/* permVector (mapper.cpp:33-117) with negative carriers wrapped by +T_u as ofdm-decoder.cpp:179-181 does: out[K] */
int dabgpu_host_perm_table (int32_t mode, int16_t *out);
/* phaseReference::refTable (phasereference.cpp:25-48): out[2*T_u] interleaved re,im */
int dabgpu_host_ref_table (int32_t mode, float *out);
/* mother-code index -> input index (-1 = punctured) of the FIC (fic != 0; fic-handler.cpp:254-288) or of an
 * MSC profile (deconvolve.cpp:142-182, 244-319); lut may be NULL to query the sizes only */
int dabgpu_host_depuncture_lut (int32_t fic, int32_t bitRate, int32_t uepFlag, int32_t protLevel,
                                int32_t *lut, int32_t lut_capacity, int32_t *lut_len, int32_t *n_punctured);
/* energy-dispersal sequence (fic-handler.cpp:100-108), one bit per byte */
int dabgpu_host_prbs (int32_t nbits, uint8_t *out);
/* ".sdr" recordings are RIFF/WAVE files (wavfiles.cpp:44-75 opens them with libsndfile and rejects everything that is
 * not 2 channels at 2 048 000 samples/s).  This parses the header of a file image (or of its first bytes: 64 KiB is
 * plenty) and says where the samples are and which decode call takes them:
 *   sample_format  1 -> dabgpu_decode_cf32 (32-bit IEEE float),  2 -> dabgpu_decode_i16 (16-bit PCM).
 * Returns DABGPU_ERR_ARG for anything else (not RIFF/WAVE, other rate / channel count -- "This is not a recorded dab
 * file", wavfiles.cpp:66-71 -- or a sample type the engine has no fetch for). */
typedef struct dabgpu_wav_info {
	int32_t format_tag;      /* 1 = PCM, 3 = IEEE float (WAVE_FORMAT_EXTENSIBLE resolved through its sub-format) */
	int32_t channels, samplerate, bits;
	int32_t sample_format;   /* see above */
	int32_t pad;
	int64_t data_offset;     /* byte offset of the first sample in the file */
	int64_t nsamples;        /* complex samples present in the bytes given (the data chunk may run past them) */
	int64_t nsamples_total;  /* complex samples the data chunk announces */
} dabgpu_wav_info;
int dabgpu_host_wav_parse (const void *file_image, size_t nbytes, dabgpu_wav_info *info);

#ifdef __cplusplus
}
#endif
#endif
