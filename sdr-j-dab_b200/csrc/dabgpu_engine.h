// dabgpu_engine.h -- stream engine state (ofdmProcessor + ficHandler + mscHandler replacement)
#pragma once
#include "dabgpu_ofdm.cuh"

#define CIF_BITS 55296                  // 864 CUs x 64 soft bits (msc-handler.cpp:42,52)
#define MAX_GROUPS 8                    // symbol groups (CTAs) per frame in the symbol kernel

// device/host mirrored control block of the sequential part of ofdmProcessor::run
struct StreamCtl {
	int synced;                         // 1 = next is SyncOnPhase, 0 = notSynced
	int lp;                             // localPhase before sample `pos`
	long long pos;                      // next sample to read, relative to the stream's sample window
	int coarse, fine, f2, prev1, prev2;
	int n_redo, pad0;                   // derive scan: frames whose inputs changed
	int n_valid;                        // frames accepted by the last pass
	int lost;                           // last pass ended with a failed findIndex
	int acq_done;                       // acquisition kernel: 1 = found the end of a null symbol
	int fault;                          // set by a kernel that had to give up (bit 0: shared-memory layout contract, bit 1: a bulk copy never arrived)
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
	int stream;                         // which stream of the round the frame belongs to (StreamDev index)
	int slot;                           // frame slot in that stream's output buffers
};

struct FrameOut {                       // per chunk slot, written by the front / symbol kernels
	int startIndex, correction;
};

// Per stream and round, device resident: everything a kernel needs to know about the stream a chunk slot belongs to.
// One stream (dabgpu_decode) is the special case n = 1; dabgpu_decode_multi runs many through the same kernels.
struct StreamDev {
	SampleWin w;
	StreamCtl ctl;                      // tracking state: read by predict / scan, committed by scan, acquire
	uint8_t *fic8, *msc8;               // soft bits as 0..255 Viterbi symbols (viterbi.cpp:229-235): FIC [frame][3 * 2K], MSC rows [15 + cif][55296]
	dabgpu_frame_info *info;            // per-frame records [frame slot]
	long long abs_base;                 // absolute sample index of window position 0
	long long limit;                    // samples of the window known to be resident (host input arrives piecewise)
	int first, budget;                  // chunk slots [first, first + budget) belong to the stream in this round
	int nframes;                        // frames actually attempted: the budget clamped (predict kernel) by what is resident once the state is known
	int slot0;                          // output frame slot of chunk frame 0
	int do_acquire;                     // run the null-symbol search for this stream in this round
	int pad;
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
	DevBuf d_sd;                        // StreamDev [nstreams] on the device
	PinBuf h_sd;
	int chunk = 1, max_chunk = 1024;
	DevBuf d_frameout, d_fcpart, d_spec0, d_info, d_framein;
	DevBuf d_soft16;                    // int16 view of the soft bits, produced on demand (dabgpu_result.soft, state export, the warp-cooperative Viterbi)
	DevBuf d_hist8;                     // [15][55296] the time de-interleaver's history rows (dab-concurrent.cpp:162-175) as byte symbols, between calls
	DevBuf d_fic8, d_msc8;              // soft bits: FIC [frames][3*2K], MSC rows [15 + cifs][55296], as 0..255 Viterbi symbols
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
	std::vector<cudaEvent_t> copy_events [2];
	DevBuf d_inbuf [2]; PinBuf h_stage [2];                  // host input: two device buffers (+ pinned staging for pageable callers), ping-pong
	struct Pending { const void *host = nullptr; size_t nsamples = 0; int fmt = 0; unsigned long long seq = 0; bool valid = false; std::vector<cudaEvent_t> ready; };
	Pending pending [2];                // dabgpu_prefetch: the block announced for (and on its way into) input buffer 0 / 1
	unsigned long long pending_seq = 0;
	int vit_batch_frames = 128;         // host-input path: frames per channel-decoding launch (cfg.host_batch_frames overrides)
	unsigned vrr = 0;                   // round robin over the channel-decoding side streams
	bool msc_packed = false;            // dabgpu_set_msc_output: MSC blocks leave with 8 bits per byte
	int out_skip = 0;                   // next decode call: FIC / info / soft results of its first out_skip frames are not delivered (multi-GPU shards, dabgpu_group.cu)
	bool needs_reset = false;           // a call failed half way: the stream state is not trustworthy until dabgpu_state_set / import
	// multi-stream batch (dabgpu_decode_multi): scratch that lives with the handle
	DevBuf m_in, m_fic8, m_msc8, m_info, m_ficbits, m_ficcrc, m_mscbits, m_offs;
	PinBuf mh_in, mh_out, mh_acq;       // mh_acq: StreamDev [n] + completion flags of the asynchronous acquisitions (read / written by the kernel in place)
	cudaStream_t acq_st [4] = { nullptr, nullptr, nullptr, nullptr };
};

int ofdm_tables_init (dabgpu *h, OfdmTables *T);

// ---- kernel launchers (dabgpu_sync.cu: the sequential part; dabgpu_symbol.cu: FFT + demod) ----
struct ChunkBufs {                      // per-round scratch, indexed by chunk slot
	FrameIn *fin; FrameOut *fo; float2 *spec0; float2 *fcpart;
};
int  sync_init (dabgpu *h);
void acquire_launch (dabgpu *h, StreamDev *sd, int nstreams, cudaStream_t st, int wave = 0, int *flags = nullptr);
void predict_launch (dabgpu *h, StreamDev *sd, int nstreams, const ChunkBufs &cb, cudaStream_t st);
void scan_launch (dabgpu *h, StreamDev *sd, int nstreams, int max_frames, int groups, const ChunkBufs &cb, int derive, cudaStream_t st);
int  symbol_init (dabgpu *h);
// which = 0: the packed register-FFT kernels where they exist for (mode, sample format); 1: the generic kernels
void front_launch (dabgpu *h, const StreamDev *sd, int nslots, const ChunkBufs &cb, int which, cudaStream_t st);
void symbol_launch (dabgpu *h, const StreamDev *sd, int nslots, int groups, const ChunkBufs &cb, int which, int fmt, cudaStream_t st);   // fmt = sample format of the streams of this round
bool symbol_has_packed (const dabgpu *h, int fmt);
