"""Generates tests/golden/golden_chain.npz: what the reference's OWN receive chain -- ofdmProcessor, ofdmDecoder, ficHandler,
mscHandler and dabConcurrent, compiled unmodified into oracle/_ref/libdabref.so and wired as gui.cpp wires them
(oracle/ref_shim/ref_tierc.cpp: ref_receive) -- delivers for seeded synthetic recordings: every FIC group, CRC flag and decoded
MSC block, bit-packed.  The recordings are regenerated from the seeds by the tests (tests/dabmod.py, well-formed FIBs).
Run in the dev container:  python tests/golden/make_golden_chain.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import dabmod   # noqa: E402
import orc      # noqa: E402

# (mode, seed, frames, cfo Hz, snr dB, lead samples, sub-channel (startAddr, bitRate, uepFlag, protLevel))
CASES = [(1, 7101, 30, 1937.0, 15.0, 23000, (0, 128, 1, 0o103)),
         (2, 7102, 70, -1130.0, 16.0, 20000, (10, 64, 1, 0o103)),
         (4, 7104, 40, -730.0, 18.0, 41000, (96, 128, 1, 0o202))]


def recording(O, case):
    mode, seed, nfr, cfo, snr, lead, sub = case
    mod = dabmod.Modulator(O, mode, [sub], seed)
    mod.wellformed_fibs = True                     # the reference's FIB parser runs on every CRC-clean FIB of its chain
    return mod, mod.generate(nfr, cfo_hz=cfo, snr_db=snr, lead=lead, tail=9000)


def main():
    R = orc.Oracle("ref")
    out = {}
    for k, case in enumerate(CASES):
        mod, tr = recording(R, case)
        s = mod.sub[0]
        fic, crc, msc, state = R.ref_receive(case[0], tr["iq"], (s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel), max_frames=case[2] + 4)
        assert crc[-16:-6].mean() > 0.8 and msc.shape[0] > 0, crc[-16:].T
        out["fic_%d" % k] = np.packbits(fic, axis=1)
        out["crc_%d" % k] = crc
        out["msc_%d" % k] = np.packbits(msc, axis=1)
        out["iq_sha_%d" % k] = np.frombuffer(__import__("hashlib").sha256(tr["iq"].tobytes()).digest(), np.uint8).copy()
        print("case", k, "mode", case[0], "groups", fic.shape[0], "blocks", msc.shape[0], "final state", state[:4])
    np.savez_compressed(os.path.join(HERE, "golden_chain.npz"), **out)


if __name__ == "__main__":
    main()
