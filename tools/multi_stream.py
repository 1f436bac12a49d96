"""BASELINE.json configs[3] at one GPU's share: 32 independent ensemble streams (16 + 32 Mode I frames each, one handle per
stream, FIC + 3 sub-channels) decoded concurrently from host threads.  Prints aggregate frames/s for 1, 8 and 32 threads.
usage: python tools/multi_stream.py"""
import importlib, os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle("port")
SUBS = [(0, 128, 1, 0o103), (96, 128, 0, 3), (200, 64, 1, 0o202)]
NSTREAM, NDISTINCT, NFR = 32, 8, 48
streams = []
for i in range(NDISTINCT):
    mod = dabmod.Modulator(port, 1, SUBS, 2000 + i)
    tr = mod.generate(NFR, cfo_hz=-3000.0 + 850.0 * i, snr_db=12.0 + 1.5 * i, lead=3000 + 977 * i, tail=6000)
    streams.append(tr["iq"])
sub_t = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
engines = [pkg.DabGpu(mode=1) for _ in range(NSTREAM)]
outs = []
for e in engines:
    e.set_subchannels(sub_t)
    outs.append(e.alloc_result(NFR + 2, want_soft=False))

def run_all(nthreads):
    done = [0] * NSTREAM
    def work(t):
        for i in range(t, NSTREAM, nthreads):
            engines[i].reset_stream() if hasattr(engines[i], "reset_stream") else None
            done[i] = engines[i].decode(streams[i % NDISTINCT], outs[i]).nframes
    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return sum(done), time.perf_counter() - t0

# every handle decodes a NEW stream each round: fresh handles would be the honest way; re-create them per round instead
for nthreads in (1, 8, 32):
    for e in engines: e.close()
    engines = [pkg.DabGpu(mode=1) for _ in range(NSTREAM)]
    for e in engines: e.set_subchannels(sub_t)
    frames, dt = run_all(nthreads)
    print("threads %2d: %d streams, %d frames in %.3f s = %.0f frames/s (%.1f ms per stream incl. acquisition)" % (nthreads, NSTREAM, frames, dt, frames / dt, dt / NSTREAM * 1e3 * nthreads))
for e in engines: e.close()
