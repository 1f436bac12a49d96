"""Where the time of a SHORT stream goes (48 Mode I frames from cold: allocation, acquisition, AFC convergence)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle("port")
SUBS = [(0, 128, 1, 0o103), (96, 128, 0, 3), (200, 64, 1, 0o202)]
mod = dabmod.Modulator(port, 1, SUBS, 2001)
iq = mod.generate(48, cfo_hz=-2150.0, snr_db=15.0, lead=3977, tail=6000)["iq"]
sub_t = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
warm = pkg.DabGpu(mode=1); warm.set_subchannels(sub_t); warm.decode(iq, warm.alloc_result(50, want_soft=False)); warm.close()   # CUDA context, module load
t0 = time.perf_counter(); e = pkg.DabGpu(mode=1); e.set_subchannels(sub_t); t1 = time.perf_counter()
out = e.alloc_result(50, want_soft=False)
s0 = e.state_get()
t2 = time.perf_counter(); r = e.decode(iq, out); t3 = time.perf_counter()
print("create %.1f ms; cold decode %.1f ms (%d frames)" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3, r.nframes))
for k in range(2):
    e.state_set(s0); e.profile_enable(True); e.profile_reset()
    l0 = e.launch_count()
    t4 = time.perf_counter(); r = e.decode(iq, out); t5 = time.perf_counter()
    prof = e.profile()
    print("warm decode %.1f ms (%d frames, %d launches)" % ((t5 - t4) * 1e3, r.nframes, e.launch_count() - l0), {k: (v[0], round(v[1], 2)) for k, v in prof.items() if v[0]})
import torch
pin = torch.empty(iq.size, dtype=torch.uint8).pin_memory(); pin.numpy()[:] = iq
for name, src in (("pageable", iq), ("pinned", (pin.data_ptr(), iq.size // 2))):
    for k in range(3):
        e.state_set(s0)
        t4 = time.perf_counter(); r = e.decode(src, out); t5 = time.perf_counter()
        print("warm decode, %s input, no profiling: %.1f ms (%d frames)" % (name, (t5 - t4) * 1e3, r.nframes))
e.close()
