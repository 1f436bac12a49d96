"""Device-resident decode rate of Modes II / IV (generic OFDM kernels) next to Mode I: frames/s and Msamples/s.
usage: python tools/mode_perf.py [mode ...]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle("port")
SUBS = [(96 * i, 128, 1, 0o103) for i in range(9)]
for mode in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    p = port.mode_params(mode)
    nframes = {1: 256, 2: 1024, 4: 512}[mode]
    lead_frames = {1: 64, 2: 256, 4: 128}[mode]                  # 6 s of signal: the fine corrector has settled (bench.py uses the same)
    mod = dabmod.Modulator(port, mode, SUBS, 77 + mode)
    total = lead_frames + nframes + 4
    truth = mod.frame_bits(total)
    parts, pos = [], 0
    for f0 in range(0, total, 64):
        x = mod.modulate(truth["bits"][f0:f0 + 64])
        parts.append(mod.channel(x, cfo_hz=137.0, snr_db=15.0, rms=30.0, lead=30000 if f0 == 0 else 0, tail=8000 if f0 + 64 >= total else 0, start_index=pos))
        pos += parts[-1].size // 2
    iq = np.concatenate(parts)
    eng = pkg.DabGpu(mode=mode)
    eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
    r0 = eng.decode(iq[:2 * (30000 + lead_frames * p.T_F - p.T_null // 2)], eng.alloc_result(lead_frames + 2, want_soft=False))
    st = eng.state_get()
    nsamp = nframes * p.T_F + 6000
    d_in = torch.from_numpy(iq[2 * st.abs_pos:2 * (st.abs_pos + nsamp)].copy()).cuda()
    out = eng.alloc_result(nframes, want_soft=False)
    out.res.fic_bits = None; out.res.fic_crc = None; out.res.info = None
    for i in range(len(SUBS)):
        out.ptrs[i] = None
    def run():
        eng.state_set(pkg.binding.StreamState.from_buffer_copy(bytes(st)))
        return eng.decode_dev(d_in.data_ptr(), nsamp, out)
    for _ in range(3):
        r = run()
    eng.profile_enable(True); eng.profile_reset()
    torch.cuda.synchronize(); eng.timer_begin()
    for _ in range(5):
        r = run()
    ms = eng.timer_end() / 5
    prof = eng.profile()
    print("mode", mode, "frames", r.nframes, "lock", st.synced, st.f2Correction, "ms/step %.3f" % ms, "frames/s %.0f" % (r.nframes / ms * 1e3),
          "Msamples/s %.0f" % (r.nframes * p.T_F / ms / 1e3), {k: round(v[1] / 5, 3) for k, v in prof.items() if v[0]})
    eng.close()
