"""debug: configs[1] full stream, engine (bench-style two calls) vs oracle trajectory; prints every mismatching frame"""
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, orc, dabmod, importlib, torch, bench
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle('port')
T_F = 196608
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iq, mod, truth = bench.make_workload(nb, 1002, orc, dabmod)
nlead = bench.LEAD_FRAMES
subs = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
sym, info = port.ofdm_run(1, iq, nlead + nb + 8)
print("oracle frames", len(info))
eng = pkg.DabGpu(mode=1)
eng.set_subchannels(subs)
lead_samples = 30000 + nlead * T_F - 20000
r0 = eng.decode(iq[:2 * lead_samples], eng.alloc_result(nlead + 2, want_soft=False))
st = eng.state_get()
print("r0 frames", r0.nframes, "abs_pos", st.abs_pos, st.synced, st.f2Correction)
nsamp = nb * T_F + 6000
batch = iq[2 * st.abs_pos:2 * (st.abs_pos + nsamp)]
d_in = torch.from_numpy(batch.copy()).cuda()
import os
if os.environ.get('DBG_SYNC','1')=='1': torch.cuda.synchronize()
if os.environ.get('DBG_EXPORT','0')=='1': blob = eng.export_state()
eng.state_set(st)
r1 = eng.decode_dev(d_in.data_ptr(), nsamp, eng.alloc_result(nb, want_soft=False))
got = list(r0.info) + list(r1.info)
f = lambda a: (a.pos, a.startIndex, a.coarse, a.fine, a.phase0, a.correction, round(a.freqCorrRe, 1), round(a.freqCorrIm, 1))
bad = 0
for k, (a, b) in enumerate(zip(got, info)):
    if f(a)[:6] != f(b)[:6]:
        print("frame", k, "(r1 idx %d)" % (k - r0.nframes), "engine", f(a), "oracle", f(b))
        bad += 1
        if bad > 12: break
print("mismatches", bad, "of", len(got))
for rep in range(3):
    eng.state_set(st)
    r2 = eng.decode_dev(d_in.data_ptr(), nsamp, eng.alloc_result(nb, want_soft=False))
    d = [k for k in range(min(r1.nframes, r2.nframes)) if f(r1.info[k])[:6] != f(r2.info[k])[:6]]
    print("repeat", rep, "nframes", r2.nframes, "differs from first run at", d[:10])
