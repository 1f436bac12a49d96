import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, orc, dabmod, importlib, torch
import test_multi_stream_gpu as T
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle('port')
iqs, mods = T._streams(port, 1, 3, T.SUBS[:1])
subl = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mods[0].sub]
eng = pkg.DabGpu(mode=1); eng.set_subchannels(subl)
a = eng.decode_multi(iqs, [eng.alloc_result(40) for _ in iqs])
f32 = [((x.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)).astype(np.float32) for x in iqs]
b = eng.decode_multi(f32, [eng.alloc_result(40) for _ in iqs])
d_in = [torch.from_numpy(x.copy()).cuda() for x in iqs]
torch.cuda.synchronize()
c = eng.decode_multi(None, [eng.alloc_result(40) for _ in iqs], dev_ptrs=[(t.data_ptr(), t.numel() // 2) for t in d_in])
f = lambda a: (a.pos, a.startIndex, a.coarse, a.fine, a.phase0, a.correction)
for i, (x, y, z) in enumerate(zip(a, b, c)):
    print(i, x.nframes, y.nframes, z.nframes, [f(q) for q in x.info[:2]], [f(q) for q in y.info[:2]], [f(q) for q in z.info[:2]])
    for name, w in (("f32", y), ("dev", z)):
        n = min(x.nframes, w.nframes)
        print("   ", name, "frames with different soft:", [k for k in range(n) if not np.array_equal(x.soft[k], w.soft[k])][:12],
              "fic", np.array_equal(x.fic_bits, w.fic_bits), "msc", np.array_equal(x.msc[0], w.msc[0]))
