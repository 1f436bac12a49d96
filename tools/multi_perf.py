"""BASELINE.json configs[3] at one GPU's share through dabgpu_decode_multi: NSTREAM independent ensemble streams
(16 + 32 Mode I frames each, FIC + 3 sub-channels) in ONE call.  Prints frames/s from host memory and device-resident,
with the per-kernel-class device times.   usage: python tools/multi_perf.py [nstreams ...]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle("port")
SUBS = [(0, 128, 1, 0o103), (96, 128, 0, 3), (200, 64, 1, 0o202)]
NDISTINCT, NFR = 8, 48
streams = []
for i in range(NDISTINCT):
    mod = dabmod.Modulator(port, 1, SUBS, 2000 + i)
    tr = mod.generate(NFR, cfo_hz=-3000.0 + 850.0 * i, snr_db=12.0 + 1.5 * i, lead=3000 + 977 * i, tail=6000)
    streams.append(tr["iq"])
sub_t = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
eng = pkg.DabGpu(mode=1)
eng.set_subchannels(sub_t)
for ns in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 128]:
    iqs = [streams[i % NDISTINCT] for i in range(ns)]
    outs = [eng.alloc_result(NFR + 2, want_soft=False) for _ in range(ns)]
    d_in = [torch.from_numpy(x.copy()).cuda() for x in iqs]
    torch.cuda.synchronize()
    dev = [(t.data_ptr(), t.numel() // 2) for t in d_in]
    for label, call in (("host", lambda: eng.decode_multi(iqs, outs)), ("dev ", lambda: eng.decode_multi(None, outs, dev_ptrs=dev))):
        call()
        eng.profile_enable(True); eng.profile_reset()
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            res = call()
        dt = (time.perf_counter() - t0) / reps
        prof = eng.profile()
        eng.profile_enable(False)
        frames = sum(r.nframes for r in res)
        print("%s %3d streams: %5d frames in %7.2f ms = %7.0f frames/s, %d launches" % (label, ns, frames, dt * 1e3, frames / dt, (eng.launch_count() - l0) // reps),
              {k: round(v[1] / reps, 2) for k, v in prof.items() if v[0]})
eng.close()
