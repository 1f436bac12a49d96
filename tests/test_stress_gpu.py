"""BASELINE.json configs[2] and configs[3] at test size.

configs[2]: Modes II / IV with a carrier offset of several carriers plus a fraction and AWGN down to 6 dB, one
EEP-2A 64 kbit/s and one UEP 128 kbit/s / level 3 sub-channel.  Low SNR is where the speculative frame-parallel
pass of the engine has the most to get wrong (failed findIndex -> resynchronisation, coarse search that does not
settle, fine corrector moving every frame): whatever the reference does with such an input, the engine must do
exactly the same -- frame positions and AFC trajectory equal, soft bits within +-1, decoded bits bit-exact.

configs[3]: several independent ensemble streams on ONE GPU, one handle per stream, decoded concurrently from
host threads (handles are independent, like one `viterbi` object per thread in the reference, viterbi.h:48-66):
every stream's output equals the same stream decoded alone."""
import threading

import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu

STRESS_SUBS = [(0, 64, 1, 0o102), (64, 128, 0, 3)]             # EEP-2A 64 kbit/s, UEP 128 kbit/s level 3
CARRIER_DIFF = {1: 1000, 2: 4000, 4: 2000}


def _oracle_chain(port, mode, iq, nmax, subs):
    sym, info = port.ofdm_run(mode, iq, nmax)
    bits, crc = port.fic_frames(mode, sym)
    msc = [port.msc_backend(port.msc_slice(mode, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel) for s in subs]
    return sym, info, bits, crc, msc


@pytest.mark.parametrize("mode,snr,carriers,frac", [(2, 6.0, 3, 0.31), (2, 10.0, -8, -0.45), (4, 6.0, -5, 0.12), (4, 10.0, 8, 0.49),
                                                    (1, 6.0, 7, -0.37), (1, 10.0, -8, 0.05), (2, 25.0, 8, 0.5), (4, 15.0, 0, -0.5)])
def test_cfo_awgn_stress_matches_oracle(port, mode, snr, carriers, frac):
    pkg = engine_pkg()
    cd = CARRIER_DIFF[mode]
    cfo = (carriers + frac) * cd
    mod = dabmod.Modulator(port, mode, STRESS_SUBS, 1003 + mode)
    nframes = 24 if mode == 1 else 48
    tr = mod.generate(nframes, cfo_hz=cfo, snr_db=snr, lead=7777 + 131 * mode, tail=5000)
    sym, info, bits, crc, msc = _oracle_chain(port, mode, tr["iq"], nframes + 4, mod.sub)
    eng = pkg.DabGpu(mode=mode, viterbi_path=2)                  # the throughput Viterbi in every mode (Modes II / IV: one CIF per frame)
    eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
    res = eng.decode(tr["iq"], eng.alloc_result(nframes + 4))
    n = res.nframes
    assert len(info) - n in (0, 1), (len(info), n)             # the oracle counts a last frame whose trailing null is cut off
    for a, b in zip(res.info, info[:n]):
        assert (a.pos, a.startIndex, a.coarse, a.fine, a.phase0, a.correction) == \
               (b.pos, b.startIndex, b.coarse, b.fine, b.phase0, b.correction)
    if n:
        assert np.abs(res.soft.astype(int) - sym[:n].astype(int)).max() <= 1
    g = mod.p.ficGroups
    assert np.array_equal(res.fic_bits, bits[:n * g]) and np.array_equal(res.fic_crc, crc[:n * g])
    for got, w in zip(res.msc, msc):
        assert np.array_equal(got, w[:got.shape[0]])
        assert n < 8 or got.shape[0] > 0
    eng.close()


def test_independent_streams_concurrent_handles(port):
    pkg = engine_pkg()
    nstreams, nframes = 6, 20
    subs = [(0, 128, 1, 0o103)]
    streams = []
    for i in range(nstreams):
        mod = dabmod.Modulator(port, 1, subs, 2000 + i)
        tr = mod.generate(nframes, cfo_hz=-3000.0 + 1170.0 * i, snr_db=12.0 + 2.5 * i, lead=3000 + 977 * i, tail=5000)
        streams.append((mod, tr["iq"]))
    sub_t = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in streams[0][0].sub]

    def decode(iq, pieces):
        eng = pkg.DabGpu(mode=1)
        eng.set_subchannels(sub_t)
        cuts = np.linspace(0, iq.size // 2, pieces + 1).astype(np.int64) * 2
        parts = [eng.decode(iq[a:b], eng.alloc_result(nframes + 4)) for a, b in zip(cuts[:-1], cuts[1:])]
        eng.close()
        return (sum(p.nframes for p in parts), np.concatenate([p.soft for p in parts]), np.concatenate([p.fic_bits for p in parts]),
                np.concatenate([p.msc[0] for p in parts]))

    alone = [decode(iq, 1) for _, iq in streams]
    together = [None] * nstreams
    errors = []

    def work(i):
        try:
            together[i] = decode(streams[i][1], 3)                # fed in three pieces while the other handles run
        except Exception as e:                                    # noqa: BLE001 -- reported below
            errors.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(nstreams)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for a, b in zip(alone, together):
        assert a[0] == b[0], (a[0], b[0])
        assert a[0] > 0
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)
    # and one of them against the oracle
    mod, iq = streams[3]
    sym, info = port.ofdm_run(1, iq, nframes + 4)
    bits, crc = port.fic_frames(1, sym)
    assert np.array_equal(alone[3][2], bits[:alone[3][0] * 4])
