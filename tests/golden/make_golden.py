"""Generates tests/golden/golden.npz from the COMPILED REFERENCE (oracle/_ref/libdabref.so, built from
/root/reference by oracle/Makefile).  Run in the dev container:  python tests/golden/make_golden.py
The reference repository ships no test vectors of its own, so these known-answer vectors are
outputs of the reference's own classes on seeded inputs; the inputs that are cheap to store are stored, the
others are regenerated from the seed by the test."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import dabmod   # noqa: E402
import orc      # noqa: E402


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def profiles(R):
    out = [(br, 0, lvl) for br in (32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384) for lvl in (1, 2, 3, 4, 5)
           if R.uep_profile(br, lvl) is not None]
    out += [(br, 1, lvl) for br in (8, 32, 64, 128, 192, 384) for lvl in (0o101, 0o102, 0o103, 0o104)]
    out += [(br, 1, lvl) for br in (32, 64, 128, 384) for lvl in (0o201, 0o202, 0o203, 0o204)]
    return out


def chain_case(O, mode, seed, nframes, cfo, snr, subs):
    mod = dabmod.Modulator(O, mode, subs, seed)
    tr = mod.generate(nframes, cfo_hz=cfo, snr_db=snr, lead=7000, tail=4000)
    sym, info = O.ofdm_run(mode, tr["iq"], nframes + 2)
    fic, crc = O.fic_frames(mode, sym)
    msc = [O.msc_backend(O.msc_slice(mode, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
    traj = np.array([(i.pos, i.startIndex, i.coarse, i.fine, i.phase0, i.correction) for i in info], np.int64)
    return tr, sym, traj, fic, crc, msc


CHAIN = dict(mode=2, seed=4242, nframes=30, cfo=1300.0, snr=18.0, subs=[(0, 64, 1, 0o103), (40, 32, 0, 5)])


def main():
    R = orc.Oracle("ref")
    g = {}
    for m in (1, 2, 3, 4):
        p = R.mode_params(m)
        g["perm_m%d" % m] = R.perm_table(m)
        ks = [k for k in range(-p.K // 2, p.K // 2 + 1) if k != 0]
        g["phi_m%d" % m] = np.array([R.phi(m, k) for k in ks], np.float32)
        g["ref_m%d" % m] = R.ref_table(m)
    g["pcodes"] = np.stack([R.pcode(n) for n in range(1, 25)])
    rng = np.random.default_rng(20261018)
    vin = rng.integers(-127, 128, (4, 4 * 774)).astype(np.int16)
    vin[3] = rng.integers(-2, 3, 4 * 774)                      # tie-heavy
    g["vit768_in"] = vin
    g["vit768_out"] = np.stack([np.packbits(R.viterbi(768, v)) for v in vin])
    prof = profiles(R)
    g["prof_list"] = np.array(prof, np.int32)
    hashes = []
    for br, flag, lvl in prof:
        r = np.random.default_rng(br * 1000 + flag * 500 + lvl)
        mask = dabmod.puncture_mask(R, br, flag, lvl)
        v = r.integers(-127, 128, -(-int(mask.sum()) // 64) * 64).astype(np.int16)
        out = R.uep_deconvolve(br, lvl, v) if flag == 0 else R.eep_deconvolve(br, lvl, v)
        hashes.append(sha(out))
    g["prof_hash"] = np.stack(hashes)
    fin = rng.integers(-127, 128, (3, 2304)).astype(np.int16)
    mod = dabmod.Modulator(R, 1, [], 5)
    fibs, punct = mod.make_fic(1)
    fin[0] = dabmod.soft_from_bits(punct[0], rng, amp=80, flip=0.02, jitter=40)
    g["fic_in"] = fin
    fo = [R.fic_decode(v) for v in fin]
    g["fic_bits"] = np.stack([np.packbits(b) for b, _ in fo]); g["fic_crc"] = np.stack([c for _, c in fo])
    frags = rng.integers(-127, 128, (20, 12 * 64)).astype(np.int16)
    g["msc_in"] = frags
    g["msc_out"] = np.packbits(R.msc_backend(frags, 16, 1, 0o103), axis=1)           # 16 kbit/s EEP 3-A = 12 CU
    g["deint_out_hash"] = sha(R.time_deinterleave(frags))
    # OFDM per-symbol pieces, Mode II (T_u = 512)
    m2 = dabmod.Modulator(R, 2, [], 77)
    p = m2.p
    bits = m2.rng.integers(0, 2, (1, p.L - 1, 2 * p.K), dtype=np.uint8)
    x = m2.modulate(bits)[p.T_null:]
    x = (x * np.exp(2j * np.pi * 3 * np.arange(x.size) / p.T_u)).astype(np.complex64)     # +3 carriers
    x += ((m2.rng.standard_normal(x.size) + 1j * m2.rng.standard_normal(x.size)) * 0.05).astype(np.complex64)
    g["ofdm_x"] = x[:p.T_g + p.T_u + 2 * p.T_s].copy()
    o = R.ofdm(2)
    g["ofdm_find_index"] = np.array([o.find_index(x[17:17 + p.T_u]), o.find_index(x[p.T_s + 40:p.T_s + 40 + p.T_u])], np.int32)
    prs = x[p.T_g:p.T_g + p.T_u]
    g["ofdm_block0"] = np.array([o.block0(prs, True), R.ofdm(2, freqSyncMethod=2).block0(prs, True)], np.int32)
    g["ofdm_phase_ref"] = o.phase_reference()
    g["ofdm_token"] = np.stack([o.token(x[(l + 1) * p.T_s:(l + 2) * p.T_s]) for l in range(2)])
    # the whole chain on a regenerated synthetic stream
    tr, sym, traj, fic, crc, msc = chain_case(R, **CHAIN)
    g["chain_iq_hash"] = sha(tr["iq"])
    g["chain_traj"] = traj
    g["chain_sym_hash"] = sha(sym)
    g["chain_sym_first"] = sym[-1, :4, :64].copy()
    g["chain_fic"] = np.packbits(fic, axis=1); g["chain_crc"] = crc
    for i, m in enumerate(msc):
        g["chain_msc%d" % i] = np.packbits(m, axis=1)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **g)
    print("wrote golden.npz:", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
