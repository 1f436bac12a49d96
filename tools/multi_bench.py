"""bench.py's multi_stream leg alone (32 cuts of the headline recording, 48 frames each, one dabgpu_decode_multi call).
DABGPU_TRACE=1 prints the rounds.   usage: python tools/multi_bench.py [nstreams] [frames_of_recording]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench, dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iq, mod, truth = bench.make_workload(nfr, 1002, orc, dabmod)
subs = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
bench.SUB5 = subs
eng = pkg.DabGpu(mode=1)
eng.set_subchannels(subs)
def barrier():
    torch.cuda.synchronize()
eng.profile_enable(True); eng.profile_reset()
r = bench.multi_stream_leg(pkg, eng, iq, 0, ns, 48, 3, barrier)
print({k: v for k, v in r.items() if k != "note"})
print({k: (v[0], round(v[1], 2)) for k, v in eng.profile().items() if v[0]})
