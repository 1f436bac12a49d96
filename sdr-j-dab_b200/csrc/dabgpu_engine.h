// dabgpu_engine.h -- stream engine state (ofdmProcessor + ficHandler + mscHandler replacement)
#pragma once
#include "dabgpu_ofdm.cuh"

#define CIF_BITS 55296                  // 864 CUs x 64 soft bits (msc-handler.cpp:42,52)
#define MAX_GROUPS 8                    // symbol groups (CTAs) per frame in the symbol kernel

// device/host mirrored control block of the sequential part of ofdmProcessor::run
struct StreamCtl {
	int synced;                         // 1 = next is SyncOnPhase, 0 = notSynced
	int lp;                             // localPhase before sample `pos`
	long long pos;                      // next sample to read, relative to the engine's sample window
	int coarse, fine, f2, prev1, prev2;
	int n_redo, pad0;                   // derive scan: frames whose inputs changed
	int n_valid;                        // frames accepted by the last pass
	int lost;                           // last pass ended with a failed findIndex
	int acq_done;                       // acquisition kernel: 1 = found the end of a null symbol
	int pad;
};

struct SampleWin {                      // two-segment sample window: [tail of earlier calls | this call's input]
	const uchar2 *seg0; long long len0;     // lengths in samples; the pointers are really const float2 * (cf32 = 1) / const short2 * (cf32 = 2)
	const uchar2 *seg1; long long len1;
	int cf32;                               // sample format: 0 = u8 I,Q pairs (rawfiles.cpp:113-116); 1 = complex float (virtual-input.h:62-63); 2 = int16 I,Q (16-bit .sdr/WAV, wavfiles.cpp:186-197)
};

struct FrameIn {                        // per chunk slot: the inputs a frame is (re)computed from
	long long P;                        // SyncOnPhase window start (window relative)
	int lp;                             // localPhase before sample P
	int phiA, phiB;                     // NCO frequency (Hz) for the PRS part / for the data symbols
	int active;                         // compute this slot in the current pass
	int pad;
};

struct FrameOut {                       // per chunk slot, written by the front / symbol kernels
	int startIndex, correction;
};

static inline size_t dab_sample_bytes (int fmt) { return fmt == 1 ? sizeof (float2) : fmt == 2 ? sizeof (short2) : sizeof (uchar2); }

struct dabgpu_backend;

struct Engine {
	OfdmTables T {};
	float2 *d_phaseRef = nullptr;       // per-call API phase reference
	bool have_phase_ref = false;
	// stream state
	StreamCtl ctl {};
	long long abs_base = 0;             // absolute sample index of window position 0
	long long frames_total = 0, cifs_total = 0;
	DevBuf tail, tail_spare; long long tail_len = 0;     // unconsumed samples, ping-pong
	int cf32 = 0;                       // sample format of the stream (fixed while a tail is pending)
	size_t sample_bytes () const { return dab_sample_bytes (cf32); }
	DevBuf d_ctl;                       // StreamCtl on the device
	PinBuf h_ctl;
	int chunk = 1, max_chunk = 1024;
	DevBuf d_frameout, d_fcpart, d_spec0, d_info, d_framein;
	DevBuf d_fic, d_msc, d_histtmp;     // soft bits: FIC [frames][3*2K], MSC rows [15 + cifs][55296]
	DevBuf d_fic8, d_msc8;              // the same as 0..255 Viterbi symbols (viterbi.cpp:229-235), written by the symbol kernels alongside
	long long cap_frames = 0;
	DevBuf d_ficbits, d_ficcrc;
	DevBuf d_figkeys;                   // FIG 0/1 scan: 64 x 2 persistent keys (dabgpu_fig.cu)
	std::vector<DevBuf> d_mscbits;
	bool hist_init = false;
	std::vector<dabgpu_subch> subch;
	std::vector<dabgpu_backend *> backends;
	int groups = 5;                     // symbol groups (CTAs) per frame for small chunks; big chunks use 3
	bool groups_fixed = false;          // DABGPU_GROUPS given: use it for every chunk
	cudaStream_t copy_st = nullptr;     // piecewise host-to-device input copies
	std::vector<cudaEvent_t> copy_events;
	int vit_batch_frames = 128;         // host-input path: frames per channel-decoding launch (cfg.host_batch_frames overrides)
	unsigned vrr = 0;                   // round robin over the channel-decoding side streams
};

int ofdm_tables_init (dabgpu *h, OfdmTables *T);
