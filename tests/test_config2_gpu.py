"""BASELINE.json configs[1] at FULL size against the oracle: the very stream bench.py times (64 lead-in + 1024 frames,
FIC + nine 96-CU EEP-3A sub-channels = 864 CU, AWGN 15 dB, CFO +137 Hz), decoded by the engine in the bench's own
two calls (lead-in, then the 1024-frame batch device-resident and once more from host memory) and by the oracle in one
run: the frame trajectory (positions, start indices, correctors, NCO phase) equal, every FIB and its CRC flag
bit-exact, every decoded bit of all nine sub-channels bit-exact.  About 10 s of oracle CPU time per oracle kind."""
import functools
import threading

import numpy as np
import pytest

import dabmod
import orc as orc_mod
from util import engine_pkg

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
T_F = 196608


@functools.lru_cache(maxsize=1)
def _workload():
    import bench
    return bench.make_workload(1024, 1002, orc_mod, dabmod)


def test_config2_full_stream_matches_oracle(port):
    import bench
    import torch
    pkg = engine_pkg()
    iq, mod, truth = _workload()
    nlead, nbatch = bench.LEAD_FRAMES, 1024
    subs = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
    # ---- oracle: one run over the whole stream; the nine sub-channel backends on threads (ctypes drops the GIL) ----
    sym, info = port.ofdm_run(1, iq, nlead + nbatch + 8)
    fic, crc = port.fic_frames(1, sym)
    want_msc = [None] * len(mod.sub)

    def backend(i, s):
        want_msc[i] = port.msc_backend(port.msc_slice(1, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
    th = [threading.Thread(target=backend, args=(i, s)) for i, s in enumerate(mod.sub)]
    [t.start() for t in th]
    [t.join() for t in th]
    # ---- engine, as bench.py drives it ----
    eng = pkg.DabGpu(mode=1)
    eng.set_subchannels(subs)
    lead_samples = 30000 + nlead * T_F - 20000
    r0 = eng.decode(iq[:2 * lead_samples], eng.alloc_result(nlead + 2, want_soft=False))
    st = eng.state_get()
    assert st.synced == 1 and st.f2Correction == 0
    nsamp = nbatch * T_F + 6000
    batch = iq[2 * st.abs_pos:2 * (st.abs_pos + nsamp)]
    d_in = torch.from_numpy(batch.copy()).cuda()
    blob = eng.export_state()
    eng.state_set(st)                                        # as bench.py does: drop the pending samples, the batch starts at abs_pos
    r1 = eng.decode_dev(d_in.data_ptr(), nsamp, eng.alloc_result(nbatch, want_soft=False))
    assert r1.nframes == nbatch and r0.nframes >= nlead - 4
    n0 = r0.nframes
    got_info = list(r0.info) + list(r1.info)
    for k, (a, b) in enumerate(zip(got_info, info)):
        assert (a.pos, a.startIndex, a.coarse, a.fine, a.phase0, a.correction) == \
               (b.pos, b.startIndex, b.coarse, b.fine, b.phase0, b.correction), (k, n0, st.abs_pos)
    n = n0 + nbatch
    assert np.array_equal(np.concatenate([r0.fic_bits, r1.fic_bits]), fic[:4 * n])
    assert np.array_equal(np.concatenate([r0.fic_crc, r1.fic_crc]), crc[:4 * n])
    assert r1.fic_crc.mean() > 0.99
    for i in range(len(subs)):
        got = np.concatenate([r0.msc[i], r1.msc[i]])
        assert got.shape[0] == 4 * n - 16 and np.array_equal(got, want_msc[i][:got.shape[0]]), i
    # the same batch from host memory (the e2e leg) after restoring the stream state: identical output
    eng.import_state(blob)
    eng.state_set(st)
    r2 = eng.decode(batch, eng.alloc_result(nbatch, want_soft=False))
    assert r2.nframes == nbatch and np.array_equal(r2.fic_bits, r1.fic_bits)
    for i in range(len(subs)):
        assert np.array_equal(r2.msc[i], r1.msc[i])
    # and the decoded payload is what the modulator sent
    pay = truth["payloads"][0]
    k0 = next((k for k in range(pay.shape[0] - 8) if np.array_equal(r1.msc[0][-8:], pay[k:k + 8])), None)
    assert k0 is not None
    eng.close()
