"""Size-independent properties at sizes the oracle would take minutes for (BASELINE.json configs[1] and [4]):

  * encode -> channel -> decode round trip of a FULL ensemble: every decoded sub-channel block equals the payload the
    test-side modulator transmitted, every FIB comes back bit for bit with a clean CRC (the whole chain -- sync, FFT,
    demodulation, both de-interleavers, depuncturing, Viterbi, dispersal -- inverted the transmitter), for the nine
    EEP-3A sub-channels of the benchmark workload and for the UEP / EEP mix of configs[4];
  * the throughput Viterbi path and the warp-cooperative one decode the same stream to the same bits;
  * feeding the stream in two calls gives the same output as one call.
No oracle is involved: the truth is what was transmitted."""
import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu

BENCH_SUBS = [(96 * i, 128, 1, 0o103) for i in range(9)]                       # 9 x 96 CU = 864 CU
# configs[4]: 32k/L5 (16 CU), 64k/L3 (48), 128k/L3 (96), 192k/L3 (140), 256k/L3 (192), 384k/L5 (192) UEP + EEP filler
UEP_SUBS = [(0, 32, 0, 5), (16, 64, 0, 3), (64, 128, 0, 3), (160, 192, 0, 3), (300, 256, 0, 3), (492, 384, 0, 5),
            (684, 128, 1, 0o103), (780, 64, 1, 0o202)]


SETTLE = 12          # decoded frames allowed for the coarse / fine frequency correction to converge (the reference's own behaviour)


def _check_against_truth(res, truth, mod):
    g = mod.p.ficGroups
    n = res.nframes
    fibs = truth["fibs"]
    # frames come out in transmission order; the receiver skips the first one(s) while it synchronises: align on a late frame
    j = n - 4
    first = next(k for k in range(8) if np.array_equal(res.fic_bits[j * g:(j + 1) * g], fibs[(k + j) * g:(k + j + 1) * g]))
    assert np.array_equal(res.fic_bits[SETTLE * g:], fibs[(first + SETTLE) * g:(first + n) * g])
    assert res.fic_crc[SETTLE * g:].all()
    cpf = mod.p.cifsPerFrame
    for got, pay in zip(res.msc, truth["payloads"]):
        nb = got.shape[0]
        assert nb >= n * cpf - 16
        k1 = next(k for k in range(nb - 1, pay.shape[0]) if np.array_equal(got[nb - 1], pay[k]))     # CIF of the last block
        b0 = SETTLE * cpf
        assert np.array_equal(got[b0:], pay[k1 - (nb - 1 - b0):k1 + 1])


@pytest.mark.parametrize("subs,seed", [(BENCH_SUBS, 1002), (UEP_SUBS, 1005)])
def test_full_ensemble_round_trip(port, subs, seed):
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 1, subs, seed)
    nframes = 96
    tr = mod.generate(nframes, cfo_hz=137.0, snr_db=15.0, lead=20000, tail=8000)
    sub_t = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
    assert sum(s.length for s in mod.sub) <= 864
    eng = pkg.DabGpu(mode=1)                                      # auto: the throughput Viterbi for a batch this size
    eng.set_subchannels(sub_t)
    one = eng.decode(tr["iq"], eng.alloc_result(nframes + 2, want_soft=False))
    assert one.nframes >= nframes - 2
    _check_against_truth(one, tr, mod)
    # the other Viterbi path, and the stream cut in two calls
    e2 = pkg.DabGpu(mode=1, viterbi_path=1)
    e2.set_subchannels(sub_t)
    cut = 2 * (tr["iq"].size // 5)
    a = e2.decode(tr["iq"][:cut], e2.alloc_result(nframes + 2, want_soft=False))
    b = e2.decode(tr["iq"][cut:], e2.alloc_result(nframes + 2, want_soft=False))
    assert a.nframes + b.nframes == one.nframes
    assert np.array_equal(np.concatenate([a.fic_bits, b.fic_bits]), one.fic_bits)
    for k in range(len(sub_t)):
        assert np.array_equal(np.concatenate([a.msc[k], b.msc[k]]), one.msc[k])
    eng.close(); e2.close()
