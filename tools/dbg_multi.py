import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, orc, dabmod, importlib
import test_multi_stream_gpu as T
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle('port')
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
subs = T.SUBS if mode != 2 else T.SUBS[:2]
iqs, mods = T._streams(port, mode, 7, subs)
sub_objs = mods[0].sub
subl = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in sub_objs]
eng = pkg.DabGpu(mode=mode); eng.set_subchannels(subl)
res = eng.decode_multi(iqs, [eng.alloc_result(50) for _ in iqs])
for i, (iq, r) in enumerate(zip(iqs, res)):
    sym, info, bits, crc, msc = T._oracle_chain(port, mode, iq, 50, sub_objs)
    e1 = pkg.DabGpu(mode=mode); e1.set_subchannels(subl)
    one = e1.decode(iq, e1.alloc_result(50))
    nf = r.nframes
    d = np.abs(r.soft.astype(int) - sym[:nf].astype(int))
    print("stream", i, "frames", nf, one.nframes, "soft diff max", d.max(), "frac", (d > 0).mean(), "soft multi==single", np.array_equal(one.soft, r.soft),
          "fic crc ok", r.fic_crc.mean())
    for k, (got, w, o1) in enumerate(zip(r.msc, msc, one.msc)):
        nb = got.shape[0]
        bad = [b for b in range(nb) if not np.array_equal(got[b], w[b])]
        bad1 = [b for b in range(nb) if not np.array_equal(got[b], o1[b])]
        print("   sub", k, "blocks", nb, "differ from oracle:", bad[:10], "from single:", bad1[:10])
        if bad:
            # the oracle's backend on the ENGINE's soft bits: is it the soft bits or the decoder?
            s = sub_objs[k]
            frag = port.msc_slice(mode, r.soft, s.startAddr, s.length)
            w2 = port.msc_backend(frag, s.bitRate, s.uepFlag, s.protLevel)
            bad2 = [b for b in range(nb) if not np.array_equal(got[b], w2[b])]
            print("      vs oracle backend fed with the engine's soft bits:", bad2[:10], " bit errors vs oracle in block", bad[0], int((got[bad[0]] != w[bad[0]]).sum()))
    e1.close()
