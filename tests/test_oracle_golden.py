"""CPU: the oracle restatement (oracle/_build/liboracle.so) against the committed golden vectors, which are
outputs of the compiled reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import os

import numpy as np
import pytest

import dabmod
from golden.make_golden import CHAIN, chain_case, profiles, sha

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_tables(port, mode):
    p = port.mode_params(mode)
    assert np.array_equal(port.perm_table(mode), G["perm_m%d" % mode])
    ks = [k for k in range(-p.K // 2, p.K // 2 + 1) if k != 0]
    assert np.array_equal(np.array([port.phi(mode, k) for k in ks], np.float32), G["phi_m%d" % mode])
    assert np.array_equal(port.ref_table(mode).view(np.uint32), G["ref_m%d" % mode].view(np.uint32))


def test_table_kats_from_the_survey(port):
    # SURVEY.md Appendix D (probes on the compiled reference)
    assert list(port.perm_table(1)[:8]) == [-513, -14, 329, 692, -733, 13, 680, 273]
    assert list(port.perm_table(2)[:8]) == [-129, -14, -55, -76, 163, 141, -88, 7]
    assert list(port.perm_table(4)[:4]) == [-257, -14, 73, 180]
    assert [round(port.phi(1, k) / (np.pi / 2)) for k in (1, 2, 3, 4, -768, 768)] == [3, 5, 3, 3, 1, 1]


def test_puncture_vectors(port):
    assert np.array_equal(np.stack([port.pcode(n) for n in range(1, 25)]), G["pcodes"])
    for n in range(1, 25):
        assert port.pcode(n).sum() == 8 + n


def test_viterbi(port):
    for v, want in zip(G["vit768_in"], G["vit768_out"]):
        assert np.array_equal(np.packbits(port.viterbi(768, v)), want)


def test_all_protection_profiles(port):
    prof = profiles(port)
    assert np.array_equal(np.array(prof, np.int32), G["prof_list"])
    for (br, flag, lvl), want in zip(prof, G["prof_hash"]):
        r = np.random.default_rng(br * 1000 + flag * 500 + lvl)
        mask = dabmod.puncture_mask(port, br, flag, lvl)
        v = r.integers(-127, 128, -(-int(mask.sum()) // 64) * 64).astype(np.int16)
        out = port.uep_deconvolve(br, lvl, v) if flag == 0 else port.eep_deconvolve(br, lvl, v)
        assert np.array_equal(sha(out), want), (br, flag, lvl)


def test_fic_and_msc_backend(port):
    for v, b, c in zip(G["fic_in"], G["fic_bits"], G["fic_crc"]):
        bits, crc = port.fic_decode(v)
        assert np.array_equal(np.packbits(bits), b) and np.array_equal(crc, c)
    assert G["fic_crc"][0].all() and not G["fic_crc"][1].any()
    assert np.array_equal(np.packbits(port.msc_backend(G["msc_in"], 16, 1, 0o103), axis=1), G["msc_out"])
    assert np.array_equal(sha(port.time_deinterleave(G["msc_in"])), G["deint_out_hash"])


def test_ofdm_pieces(port):
    x = G["ofdm_x"]
    p = port.mode_params(2)
    o = port.ofdm(2)
    assert o.find_index(x[17:17 + p.T_u]) == G["ofdm_find_index"][0]
    assert o.find_index(x[p.T_s + 40:p.T_s + 40 + p.T_u]) == G["ofdm_find_index"][1]
    prs = x[p.T_g:p.T_g + p.T_u]
    assert o.block0(prs, True) == G["ofdm_block0"][0]
    assert port.ofdm(2, freqSyncMethod=2).block0(prs, True) == G["ofdm_block0"][1]
    # both builds share the stand-in FFT, so these are exact
    assert np.array_equal(o.phase_reference().view(np.uint32), G["ofdm_phase_ref"].view(np.uint32))
    tok = np.stack([o.token(x[(l + 1) * p.T_s:(l + 2) * p.T_s]) for l in range(2)])
    assert np.array_equal(tok, G["ofdm_token"])


def test_whole_chain(port):
    tr, sym, traj, fic, crc, msc = chain_case(port, **CHAIN)
    if not np.array_equal(sha(tr["iq"]), G["chain_iq_hash"]):
        pytest.skip("the modulator produced another stream than when the golden file was made (numpy/scipy version)")
    assert np.array_equal(traj, G["chain_traj"])
    assert np.array_equal(sha(sym), G["chain_sym_hash"]) and np.array_equal(sym[-1, :4, :64], G["chain_sym_first"])
    assert np.array_equal(np.packbits(fic, axis=1), G["chain_fic"]) and np.array_equal(crc, G["chain_crc"])
    for i, m in enumerate(msc):
        assert np.array_equal(np.packbits(m, axis=1), G["chain_msc%d" % i])
    assert crc[-8:].all()


def test_chain_equals_the_references_own_chain_golden(port):
    """tests/golden/golden_chain.npz = what the reference's own classes, compiled unmodified and wired as the reference wires them,
    delivered for three seeded recordings (make_golden_chain.py): the restated chain reproduces every FIC group, CRC flag and MSC
    block -- with or without /root/reference at hand"""
    import hashlib
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_chain as mg
    g = np.load(os.path.join(here, "golden", "golden_chain.npz"))
    for k, case in enumerate(mg.CASES):
        mod, tr = mg.recording(port, case)
        assert np.array_equal(np.frombuffer(hashlib.sha256(tr["iq"].tobytes()).digest(), np.uint8), g["iq_sha_%d" % k]), "the recording is not the one the golden was made from"
        s = mod.sub[0]
        sym, info = port.ofdm_run(case[0], tr["iq"], case[2] + 4)
        fic, crc = port.fic_frames(case[0], sym)
        msc = port.msc_backend(port.msc_slice(case[0], sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
        gf, gc, gm = np.unpackbits(g["fic_%d" % k], axis=1), g["crc_%d" % k], np.unpackbits(g["msc_%d" % k], axis=1)
        n, m = min(fic.shape[0], gf.shape[0]), min(msc.shape[0], gm.shape[0])
        assert n >= gf.shape[0] - 2 * mod.p.ficGroups and m >= gm.shape[0] - mod.p.cifsPerFrame - 1 and m > 0
        assert np.array_equal(fic[:n], gf[:n]) and np.array_equal(crc[:n], gc[:n]) and np.array_equal(msc[:m], gm[:m])
