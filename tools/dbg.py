import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, orc, dabmod, importlib
pkg=importlib.import_module("sdr-j-dab_b200")
port=orc.Oracle('port')
mode,cfo,snr=2,-9300.0,18.0
subs=[(0, 128, 1, 0o103), (96, 128, 0, 3)]
mod = dabmod.Modulator(port, mode, subs, 1001)
tr = mod.generate(40, cfo_hz=cfo, snr_db=snr, lead=12345, tail=5000)
sym,info=port.ofdm_run(mode,tr['iq'],44)
bits,crc=port.fic_frames(mode,sym)
print(len(info),crc.all(1).astype(int))
print([(i.pos,i.startIndex,i.coarse,i.fine,i.correction) for i in info[-6:]])
eng=pkg.DabGpu(mode=mode); eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
res=eng.decode(tr['iq'],eng.alloc_result(44))
print(res.nframes,res.fic_crc.all(1).astype(int))
print([(i.pos,i.startIndex,i.coarse,i.fine,i.correction) for i in res.info[-6:]])
