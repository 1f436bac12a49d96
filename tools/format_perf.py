"""Device-resident decode rate of one Mode I stream in the three sample formats (u8 rawfile, complex float, int16 .sdr / WAV payload):
the symbol / front kernels' sample fetch is the only difference.   usage: python tools/format_perf.py"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, dabmod, orc
pkg = importlib.import_module("sdr-j-dab_b200")
port = orc.Oracle("port")
SUBS = [(96 * i, 128, 1, 0o103) for i in range(9)]
mode, nframes, lead_frames = 1, 512, 64
p = port.mode_params(mode)
mod = dabmod.Modulator(port, mode, SUBS, 78)
total = lead_frames + nframes + 4
truth = mod.frame_bits(total)
parts, pos = [], 0
for f0 in range(0, total, 64):
    x = mod.modulate(truth["bits"][f0:f0 + 64])
    parts.append(mod.channel(x, cfo_hz=137.0, snr_db=15.0, rms=30.0, lead=30000 if f0 == 0 else 0, tail=8000 if f0 + 64 >= total else 0, start_index=pos))
    pos += parts[-1].size // 2
iq = np.concatenate(parts)
ref_fic = None
for name, conv, fn in (("u8", lambda a: a, "dabgpu_decode_dev"),
                       ("cf32", lambda a: ((a.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)), "dabgpu_decode_cf32_dev"),
                       ("int16", lambda a: ((a.astype(np.int16) - 128) * 256).astype(np.int16), "dabgpu_decode_i16_dev")):
    eng = pkg.DabGpu(mode=mode)
    eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
    lead_n = 30000 + lead_frames * p.T_F - p.T_null // 2
    eng.decode(conv(iq[:2 * lead_n]) if name != "int16" else iq[:2 * lead_n], eng.alloc_result(lead_frames + 2, want_soft=False))     # lock on the u8 / float form
    st = eng.state_get()
    nsamp = nframes * p.T_F + 6000
    d_in = torch.from_numpy(np.ascontiguousarray(conv(iq[2 * st.abs_pos:2 * (st.abs_pos + nsamp)]))).cuda()
    out = eng.alloc_result(nframes, want_soft=False)
    f = getattr(eng.lib, fn)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]

    def run():
        eng.state_set(pkg.binding.StreamState.from_buffer_copy(bytes(st)))
        rc = f(eng.h, d_in.data_ptr(), nsamp, C.addressof(out.res))
        assert rc == 0, rc
    for _ in range(3):
        run()
    eng.profile_enable(True); eng.profile_reset()
    torch.cuda.synchronize(); eng.timer_begin()
    for _ in range(5):
        run()
    ms = eng.timer_end() / 5
    prof = eng.profile()
    r = eng._trim(out)
    if ref_fic is None:
        ref_fic = r.fic_bits.copy()
    print("%-5s frames %d  ms/step %.3f  frames/s %.0f  Msamples/s %.0f  FIC equal to u8 run: %s " % (name, r.nframes, ms, r.nframes / ms * 1e3, r.nframes * p.T_F / ms / 1e3, np.array_equal(ref_fic, r.fic_bits)),
          {k: round(v[1] / 5, 3) for k, v in prof.items() if v[0]})
    eng.close()
