"""time of the acquisition kernel per sample: a noise-only stream (never syncs: every sample goes through the search)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("sdr-j-dab_b200")
rng = np.random.default_rng(5)
n = 4_000_000
iq = np.clip(np.rint(rng.standard_normal(2 * n) * 20 + 128), 0, 255).astype(np.uint8)
eng = pkg.DabGpu(mode=1)
for rep in range(2):
    eng2 = pkg.DabGpu(mode=1)
    eng2.profile_enable(True); eng2.profile_reset()
    t0 = time.perf_counter()
    r = eng2.decode(iq, eng2.alloc_result(8, want_soft=False))
    dt = time.perf_counter() - t0
    p = eng2.profile()
    print("noise %d samples: call %.2f ms, acquire kernel %.2f ms in %d launches = %.2f ns per sample, frames %d" % (n, dt * 1e3, p["acquire"][1], p["acquire"][0], p["acquire"][1] * 1e6 / n, r.nframes))
    eng2.close()
