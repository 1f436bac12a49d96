"""CPU: the oracle restatement against the COMPILED REFERENCE (oracle/_ref, built from /root/reference) on
larger seeded random inputs.  Skipped where neither /root/reference nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

import dabmod


def test_build_kinds(port, ref):
    assert port.kind == "port" and ref.kind == "reference"


@pytest.mark.parametrize("frameBits", [24, 768, 3072])
def test_viterbi_random(port, ref, frameBits):
    rng = np.random.default_rng(frameBits)
    for i in range(12):
        lim = (127, 127, 2, 400)[i % 4]
        s = rng.integers(-lim, lim + 1, 4 * (frameBits + 6)).astype(np.int16)
        assert np.array_equal(port.viterbi(frameBits, s), ref.viterbi(frameBits, s))


def test_protection_profiles_random(port, ref):
    rng = np.random.default_rng(3)
    for br, flag, lvl in [(128, 1, 0o103), (8, 1, 0o102), (64, 1, 0o204), (128, 0, 3), (80, 0, 1), (32, 0, 5), (384, 0, 1)]:
        mask = dabmod.puncture_mask(port, br, flag, lvl)
        v = rng.integers(-127, 128, -(-int(mask.sum()) // 64) * 64).astype(np.int16)
        a = port.uep_deconvolve(br, lvl, v) if flag == 0 else port.eep_deconvolve(br, lvl, v)
        b = ref.uep_deconvolve(br, lvl, v) if flag == 0 else ref.eep_deconvolve(br, lvl, v)
        assert np.array_equal(a, b), (br, flag, lvl)


@pytest.mark.parametrize("mode,method", [(1, 1), (1, 2), (1, 0), (2, 1), (4, 2), (3, 1)])
def test_ofdm_classes(port, ref, mode, method):
    rng = np.random.default_rng(mode * 10 + method)
    p = port.mode_params(mode)
    a, b = port.ofdm(mode, 3, method), ref.ofdm(mode, 3, method)
    mod = dabmod.Modulator(port, mode, [], mode)
    bits = rng.integers(0, 2, (1, p.L - 1, 2 * p.K), dtype=np.uint8)
    x = mod.modulate(bits)[p.T_null:]
    x = (x * np.exp(2j * np.pi * (-5) * np.arange(x.size) / p.T_u) + 0.05 * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
    for start in (0, 31, p.T_g, 3 * p.T_s):
        w = x[start:start + p.T_u]
        assert a.find_index(w) == b.find_index(w)
    prs = x[p.T_g:p.T_g + p.T_u]
    assert a.block0(prs, True) == b.block0(prs, True)
    for l in range(3):
        s = x[(l + 1) * p.T_s:(l + 2) * p.T_s]
        assert np.array_equal(a.token(s), b.token(s))
    assert np.array_equal(a.phase_reference().view(np.uint32), b.phase_reference().view(np.uint32))


def test_whole_chain_random_stream(port, ref):
    mod = dabmod.Modulator(port, 4, [(0, 64, 1, 0o103)], 99)
    tr = mod.generate(20, cfo_hz=-4400.0, snr_db=14.0, lead=999, tail=3000)
    sa, ia = port.ofdm_run(4, tr["iq"], 24)
    sb, ib = ref.ofdm_run(4, tr["iq"], 24)
    assert len(ia) == len(ib) > 10 and np.array_equal(sa, sb)
    assert [(i.pos, i.startIndex, i.coarse, i.fine) for i in ia] == [(i.pos, i.startIndex, i.coarse, i.fine) for i in ib]
    s = mod.sub[0]
    fa = port.msc_backend(port.msc_slice(4, sa, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
    fb = ref.msc_backend(ref.msc_slice(4, sb, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
    assert np.array_equal(fa, fb)
