#!/usr/bin/env python
"""bench.py -- Mode I frames/s through sync + FFT + demod + Viterbi (BASELINE.json metric).

Workload (BASELINE.json configs[1]): a Mode I stream of 1024 frames (after a 16-frame lead-in) carrying the
FIC and nine 96-CU EEP-3A 128 kbit/s sub-channels = all 864 capacity units, from the test-side modulator with
AWGN (15 dB) and a carrier offset; one step = one dabgpu_decode of the 1024-frame batch (stream state restored
to the locked state before every step).  Synthetic data, generated here; no reference code is read at run time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

N > 1: launched under torchrun, one engine per GPU, every rank decodes its own ensemble (weak scaling, no
data-path collective); timed with barriers on both sides, max over ranks.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODE = 1
T_F = 196608
SUBS = [(96 * i, 128, 1, 0o103) for i in range(9)]          # (startAddr, bitRate, uepFlag, protLevel)
LEAD_FRAMES = 64                                             # lead-in decoded before the timed region: acquisition, AFC convergence, de-interleaver fill
ALG_BYTES_PER_FRAME = 2 * T_F + 2 * 75 * 3072                # u8 IQ read once + int16 soft bits written once
INT_OPS_PER_STEP = 272                                       # 64 ACS x 4 + 16 branch-metric ops
STEPS_PER_FRAME = 4 * 9 * (3072 + 6) + 4 * (768 + 6)         # trellis steps per Mode I frame of this workload


def make_workload(nframes, seed, orc_mod, dabmod):
    """-> u8 IQ of (LEAD_FRAMES + nframes) frames plus margins, and the modulator (for sub-channel geometry)"""
    port = orc_mod.Oracle("port")                            # table provider for the modulator only
    mod = dabmod.Modulator(port, MODE, SUBS, seed)
    total = LEAD_FRAMES + nframes + 4
    truth = mod.frame_bits(total)
    lead, tail = 30000, 8000
    parts = []
    pos = 0
    chunk = 64
    for f0 in range(0, total, chunk):
        x = mod.modulate(truth["bits"][f0:f0 + chunk])
        first, last = f0 == 0, f0 + chunk >= total
        parts.append(mod.channel(x, cfo_hz=137.0, snr_db=15.0, rms=30.0, lead=lead if first else 0,
                                 tail=tail if last else 0, start_index=pos))
        pos += parts[-1].size // 2
    return np.concatenate(parts), mod, truth


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons with NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_run(iq, mod, nthreads, frames_per_piece, steps, warmup, orc_mod, min_seconds=0.0):
    """The reference's CPU chain (oracle/_ref when the compiled reference travelled with the repo, else the
    oracle port) on `nthreads` host threads: every thread runs the full chain -- sync/AFC loop, FFT + demod,
    FIC decode, time de-interleave + EEP decode of all nine sub-channels -- on its own piece of the stream
    (LEAD_FRAMES + frames_per_piece frames).  One decoder object per thread (the reference's are not
    re-entrant).  Returns (kind, frames/s, ms_per_step, frames per step)."""
    try:
        O = orc_mod.Oracle("ref")
    except Exception:
        O = orc_mod.Oracle("port")
    piece = (LEAD_FRAMES + frames_per_piece) * T_F + 30000 + 8000
    piece_iq = iq[:2 * piece]
    counts = [0] * nthreads

    def work(t):
        sym, info = O.ofdm_run(MODE, piece_iq, LEAD_FRAMES + frames_per_piece + 2)
        O.fic_frames(MODE, sym)
        for s in mod.sub:
            O.msc_backend(O.msc_slice(MODE, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
        counts[t] = len(info)

    def step():
        th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        step()
    times = [step() for _ in range(steps)]
    while min_seconds and sum(times) < min_seconds and len(times) < 40:     # bounded sample: ~10 s of wall clock
        times.append(step())
    frames = sum(counts)
    dt = sum(times) / len(times)
    return O.kind, frames / dt, dt * 1e3, frames


def cpu_viterbi_only(nthreads, orc_mod, seconds=3.0):
    """Viterbi-only leg of the CPU baseline: the reference's time de-interleave + EEP-3A depuncture +
    viterbi::deconvolve + dispersal on random soft bits, one decoder object per thread, decoded Mbit/s over all threads"""
    try:
        O = orc_mod.Oracle("ref")
    except Exception:
        O = orc_mod.Oracle("port")
    rng = np.random.default_rng(7)
    ncif = 16 + 96
    frags = rng.integers(-127, 128, (ncif, 96 * 64), dtype=np.int16)
    done = [0] * nthreads
    t_end = time.perf_counter() + seconds

    def work(t):
        while time.perf_counter() < t_end:
            O.msc_backend(frags, 128, 1, 0o103)
            done[t] += (ncif - 16) * 3072
    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    return sum(done) / (time.perf_counter() - t0) / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1024, help="frames per step (the metric's config is 1024)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dev-batch", type=int, default=0, help="frames per channel-decoding launch on the device-resident path; 0 = one launch per step")
    ap.add_argument("--host-batch", type=int, default=0, help="frames per channel-decoding launch on the host-input (e2e) path; 0 = engine default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import dabmod
    import orc as orc_mod
    config = {"workload": "Mode I %d-frame batch, FIC + 9 x 96-CU EEP-3A 128 kbit/s sub-channels (864 CU), AWGN 15 dB, CFO +137 Hz"
                          % args.frames, "frames_per_step": args.frames, "lead_in_frames": LEAD_FRAMES,
              "l2": "inputs larger than L2 (%.0f MB of u8 IQ per step)" % (args.frames * T_F * 2 / 1e6)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        nthreads = args.cpu_threads or (os.cpu_count() or 1)
        fpp = 24
        iq, mod, _ = make_workload(fpp + 2, 1002, orc_mod, dabmod)
        kind, fps, ms, frames = cpu_reference_run(iq, mod, nthreads, fpp, max(args.steps, 1), args.warmup, orc_mod)
        line = {"impl": "reference", "metric": "Mode I frames/s (sync+FFT+demod+Viterbi)", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": kind,
                                 "sample": "%d threads x (%d lead-in + %d frames) of the same stream, full chain incl. all 9 sub-channels; "
                                           "%d frames decoded per step; FFT = labelled FFTW stand-in" % (nthreads, LEAD_FRAMES, fpp, frames)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    pkg = importlib.import_module("sdr-j-dab_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    iq, mod, truth = make_workload(args.frames, 1002 + rank, orc_mod, dabmod)
    subs = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
    eng = pkg.DabGpu(mode=MODE, device=local_rank, host_batch_frames=args.host_batch, dev_batch_frames=args.dev_batch)
    eng.set_subchannels(subs)

    # lead-in: acquire + lock + fill the de-interleaver, then remember the locked stream state
    lead_samples = 30000 + LEAD_FRAMES * T_F - 20000          # stop inside the null symbol before frame LEAD_FRAMES
    out_lead = eng.alloc_result(LEAD_FRAMES + 2, want_soft=False)
    r0 = eng.decode(iq[:2 * lead_samples], out_lead)
    st = eng.state_get()
    hist_frames = r0.nframes
    # the batch = everything after what the lead-in consumed
    batch_first = st.abs_pos
    assert st.synced == 1 and st.f2Correction == 0, "lead-in did not lock (synced %d, coarse search %d)" % (st.synced, st.f2Correction)
    nsamp = args.frames * T_F + 6000                          # exactly `frames` frames + margin for the last PRS window
    batch = iq[2 * batch_first:2 * (batch_first + nsamp)]
    assert batch.size == 2 * nsamp
    h_in = torch.empty(batch.size, dtype=torch.uint8).pin_memory()
    h_in.numpy()[:] = batch
    d_in = h_in.to(dev)
    # pinned result buffers for the e2e leg
    def pinned(shape, dtype):
        assert dtype == np.uint8
        return torch.empty(shape, dtype=torch.uint8).pin_memory().numpy()
    out_e2e = eng.alloc_result(args.frames, want_soft=False, alloc=pinned)
    # device-resident leg: no result download
    out_dev = eng.alloc_result(args.frames, want_soft=False)
    out_dev.res.fic_bits = None; out_dev.res.fic_crc = None; out_dev.res.info = None
    for i in range(len(subs)):
        out_dev.ptrs[i] = None
    hist0 = None

    def restore():
        # back to the locked state right after the lead-in (the history rows are restored by the engine's
        # own lead-in decode only once; the de-interleaver history content does not change the work done)
        s = pkg.binding.StreamState.from_buffer_copy(bytes(st))
        eng.state_set(s)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_dev():
        restore()
        return eng.decode_dev(d_in.data_ptr(), nsamp, out_dev)

    def run_e2e():
        restore()
        return eng.decode((h_in.data_ptr(), nsamp), out_e2e)

    # correctness gate before timing: FIC CRCs of the batch are ok and the payload comes out
    r = run_e2e()
    assert r.nframes == args.frames, r.nframes
    crc_ok = float(r.fic_crc.mean())
    blocks = r.msc[0].shape[0]
    pay = truth["payloads"][0]
    k0 = next((k for k in range(pay.shape[0] - 4) if np.array_equal(r.msc[0][blocks - 4:blocks], pay[k:k + 4])), None)
    assert crc_ok > 0.99 and k0 is not None, (crc_ok, k0)
    frames_per_step = r.nframes

    for _ in range(warmup):
        run_dev()
    eng.profile_enable(True)
    eng.profile_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    eng.timer_begin()
    for _ in range(args.steps):
        run_dev()
    dev_ms = eng.timer_end()
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - l0
    prof = eng.profile()
    eng.profile_enable(False)
    clocks = sampler.result()

    # e2e: host (pinned) input in, decoded bits out, per step, wall clock with sync on both sides
    for _ in range(2):
        run_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        run_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t1

    elapsed = max(dev_ms / 1e3, 0.0)
    if dist is not None:
        tt = torch.tensor([elapsed, wall, e2e_wall], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed, wall, e2e_wall = tt.tolist()
        ff = torch.tensor([frames_per_step], device=dev, dtype=torch.float64)
        dist.all_reduce(ff, op=dist.ReduceOp.SUM)
        total_frames = ff.item()
    else:
        total_frames = frames_per_step
    value = total_frames * args.steps / elapsed
    e2e_value = total_frames * args.steps / e2e_wall

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        traffic = {}
        try:                                                     # per-launch DRAM bytes of the committed ncu capture (same launch shape only)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("frames_per_launch") == frames_per_step and not args.dev_batch:
                traffic = {k: v["read"] + v["write"] for k, v in tj.items() if isinstance(v, dict)}
        except Exception:
            pass
        ip = eng.int_peak()
        int_peak = max(ip[k] for k in ("add", "min", "add_mad"))
        n_vit, ms_vit = prof["viterbi_msc"]                   # forward (add-compare-select) launches: FIC + all sub-channels per launch
        n_tb, ms_tb = prof["viterbi_tb"]
        n_sym, ms_sym = prof["symbol"]
        steps_per_launch = frames_per_step * STEPS_PER_FRAME * args.steps / max(n_vit, 1)   # trellis steps per forward launch
        vit_avg_ms = ms_vit / max(n_vit, 1)
        vit_ops = INT_OPS_PER_STEP * steps_per_launch / (vit_avg_ms * 1e-3) if n_vit else 0.0
        # every chunk is followed by a verification pass whose launches exit at once when nothing changed (a locked
        # stream): half of the symbol-kernel launches do the work
        n_sym_work = max(n_sym // 2, 1)
        sym_frames = frames_per_step * args.steps / n_sym_work
        sym_avg_ms = ms_sym / n_sym_work
        sym_gbs = ALG_BYTES_PER_FRAME * sym_frames / (sym_avg_ms * 1e-3) / 1e9 if n_sym else 0.0
        shares = {k: round(v[1] / (dev_ms if dev_ms > 0 else 1.0), 4) for k, v in prof.items() if v[0]}
        line = {
            "metric": "Mode I frames/s (sync+FFT+demod+Viterbi)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": elapsed * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (FFT/demod) + u16/u32 (Viterbi metrics)", "data": "synthetic",
            "config": dict(config, frames_decoded_per_step=frames_per_step, parallelism="1 stream per GPU, frame-parallel inside",
                           fic_crc_ok=crc_ok),
            "msamples_per_s": value * T_F / 1e6,
            "viterbi_gbit_per_s": (frames_per_step * args.steps * (4 * 9 * 3072 + 4 * 768)) / ((ms_vit + ms_tb) * 1e-3) / 1e9 if n_vit else None,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(nsamp * 2),
                    "d2h_bytes_per_step": int(frames_per_step * (4 * 768 + 12) + sum(m.shape[0] * m.shape[1] for m in r.msc)),
                    "ms_per_step": e2e_wall * 1e3 / args.steps, "timing": "wall clock, synchronised on both sides, pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant kernel by time: the Viterbi forward pass; integer-ALU bound
            "roofline": {"kernel": "vit_simd_forward (add-compare-select of FIC + 9 sub-channels, one launch per step)", "bound": "alu", "achieved": vit_ops / 1e12, "peak": int_peak / 1e12,
                         "unit": "Tint-op/s", "frac": vit_ops / int_peak if int_peak else None, "traffic": traffic.get("vit_simd_forward"),
                         "avg_launch_ms": vit_avg_ms, "launches": n_vit, "share_of_step": shares.get("viterbi_msc"),
                         "peak_source": "measured live by dabgpu_int_peak (add / min / add+mad.lo micro-benchmark): %s" %
                                        {k: round(v / 1e12, 2) for k, v in ip.items()},
                         "algorithmic": "272 int-ops per trellis step (SURVEY.md 8d) x %d steps per launch" % steps_per_launch},
            "roofline_hbm": {"kernel": "symbol_kernel (FFT+demod group)", "bound": "hbm", "achieved": sym_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": sym_gbs / hbm_peak, "traffic": traffic.get("symbol_kernel_r8"), "avg_launch_ms": sym_avg_ms, "launches": n_sym_work,
                             "share_of_step": shares.get("symbol"), "peak_source": peak_src,
                             "algorithmic": "854016 B per Mode I frame (SURVEY.md 8d) x %.1f frames per working launch (+ %d no-op verification launches)" % (sym_frames, n_sym - n_sym_work),
                             "fp32_tflops": 11.9e6 * sym_frames / (sym_avg_ms * 1e-3) / 1e12 if n_sym else None},
            "kernel_shares": shares,
            "wall_ms_per_step": wall * 1e3 / args.steps,
        }
    # CPU baseline on rank 0 at N = 1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = args.cpu_threads or (os.cpu_count() or 1)
        fpp = 24
        kind, fps, ms, frames = cpu_reference_run(iq, mod, nthreads, fpp, 1, 0, orc_mod, min_seconds=10.0)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": kind,
                                "sample": "%d threads x (%d lead-in + %d frames) of the same stream, full chain incl. all 9 sub-channels, "
                                          "%d frames per %.2f s pass, passes repeated for >= 10 s; FFT = labelled FFTW stand-in" % (nthreads, LEAD_FRAMES, fpp, frames, ms / 1e3)}
        # the same chain on ONE thread, and the Viterbi-only rate (bounded samples, a few seconds each)
        _, fps1, _, _ = cpu_reference_run(iq, mod, 1, fpp, 1, 0, orc_mod, min_seconds=2.0)
        line["cpu_baseline"]["single_thread_frames_per_s"] = fps1
        line["cpu_baseline"]["viterbi_only_mbit_per_s"] = {"threads_%d" % nthreads: cpu_viterbi_only(nthreads, orc_mod), "threads_1": cpu_viterbi_only(1, orc_mod, 2.0)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
