"""GPU: airspy-style sample-rate conversion (dabgpu_resample_i16; airspy-handler.cpp:138-148, 342-370) against the
oracle's streaming restatement: bit-exact floats, whole-block consumption, piecewise == one shot, and the converted
stream decodes like the 2.048 MS/s original."""
import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rate,n", [(2500000, 2500 * 40 + 1), (2500000, 2500 * 7 + 1234), (3000000, 3000 * 5 + 1), (2048000, 2048 * 3 + 1), (2500000, 2500), (2500000, 0)])
def test_resample_matches_oracle(port, rate, n):
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    rng = np.random.default_rng(n % 1000 + 1)
    iq = rng.integers(-2048, 2048, 2 * n).astype(np.int16)
    got, consumed = eng.resample_i16(iq, rate)
    want = port.resample_i16(iq, rate)
    R = rate // 1000
    assert consumed == (max(n - 1, 0) // R) * R
    assert got.size == want.size and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    eng.close()


def test_resample_in_pieces_and_decode(port):
    """a Mode II stream is up-sampled to 2.5 MS/s int16 on the test side, converted back in ragged pieces (the caller
    carries the unconsumed samples), and decoded through the float-sample entry point: FIC CRCs hold"""
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 2, [(0, 64, 1, 0o102)], 31)
    iq = mod.generate(20, cfo_hz=500.0, snr_db=25.0, rms=30.0, lead=3000, tail=4000)["iq"]
    x = ((iq[0::2].astype(np.float64) - 128) + 1j * (iq[1::2].astype(np.float64) - 128)) * 40.0
    t_out = np.arange(int(x.size * 2500 / 2048) - 2) * (2048.0 / 2500.0)
    k = np.floor(t_out).astype(np.int64)
    f = t_out - k
    y = x[k] * (1 - f) + x[k + 1] * f                            # linear up-sampling, good enough for a functional check
    i16 = np.empty(2 * y.size, np.int16)
    i16[0::2] = np.clip(np.rint(y.real), -2048, 2047); i16[1::2] = np.clip(np.rint(y.imag), -2048, 2047)
    eng = pkg.DabGpu(mode=2)
    eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
    one, _ = eng.resample_i16(i16, 2500000)
    assert np.array_equal(one.view(np.uint32), port.resample_i16(i16, 2500000).view(np.uint32))
    pieces, pending = [], np.zeros(0, np.int16)
    for a, b in zip([0, 12346, 500000, 500002], [12346, 500000, 500002, i16.size // 2]):
        buf = np.concatenate([pending, i16[2 * a:2 * b]])
        out, consumed = eng.resample_i16(buf, 2500000)
        pieces.append(out)
        pending = buf[2 * consumed:]
    assert np.array_equal(np.concatenate(pieces).view(np.uint32), one.view(np.uint32))
    r = eng.decode(one, eng.alloc_result(24))
    assert r.nframes >= 16 and r.fic_crc[-8:].all()
    sym, info = port.ofdm_run(2, one, 24)
    assert [(i.pos, i.fine) for i in r.info] == [(i.pos, i.fine) for i in info[:r.nframes]]
    eng.close()
