"""The drop-in C++ adapter classes (sdr-j-dab_b200/host/dab_adapters.h): compile check on CPU; on the GPU box the
demo program drives them like the reference's own code does and its outputs are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, engine_pkg

HOST = os.path.join(ROOT, "sdr-j-dab_b200", "host")
EXE = os.path.join(HOST, "adapter_demo")


def _build():
    import importlib
    importlib.import_module("sdr-j-dab_b200.build").build()
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", EXE, os.path.join(HOST, "adapter_demo.cpp"),
                    "-L" + os.path.join(ROOT, "sdr-j-dab_b200"), "-ldabgpu", "-Wl,-rpath," + os.path.join(ROOT, "sdr-j-dab_b200")],
                   check=True)


def test_adapters_compile_and_link():
    _build()
    assert os.path.exists(EXE)


def _lcg_stream(seed):
    s = seed
    while True:
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        yield (s >> 8)


def _fnv(a):
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_adapter_demo_matches_oracle(port):
    _build()
    import dabplus
    rng = np.random.default_rng(4)
    rows = [rng.integers(0, 2, (3, 768), dtype=np.uint8)]
    for i in range(6):
        sf, coded, _ = dabplus.make_superframe(32, rng, dac_rate=i & 1, sbr=(i >> 1) & 1)
        if i == 2:
            coded[3 * 4 + 1::4][:3] ^= 0x21                      # three byte errors in column 1
        rows.append(dabplus.to_cif_bits(coded, 32))
    cifs = np.concatenate(rows)
    path = os.path.join(HOST, "adapter_demo_cifs.bin")
    cifs.tofile(path)
    # a Mode I recording as complex floats for the throughput ofdmProcessor adapter
    import dabmod
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103)], 5150)
    tr = mod.generate(20, cfo_hz=2345.0, snr_db=20.0, lead=4000, tail=6000)
    f32 = ((tr["iq"].astype(np.float32) - np.float32(128.0)) / np.float32(128.0)).astype(np.float32)
    iqpath = os.path.join(HOST, "adapter_demo_iq.bin")
    f32.tofile(iqpath)
    # the oracle's soft bits of the same recording for the per-symbol handler members (process_ficBlock / process_mscBlock)
    sym, finfo = port.ofdm_run(1, f32, 30)
    softpath = os.path.join(HOST, "adapter_demo_soft.bin")
    np.ascontiguousarray(sym, np.int16).tofile(softpath)
    # a longer u8 recording for the C++ multi-GPU group (dabgpu_group_*): two sub-channels, 64 frames
    mod2 = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103), (96, 128, 0, 3)], 5151)
    u8path = os.path.join(HOST, "adapter_demo_u8.bin")
    mod2.generate(64, cfo_hz=137.0, snr_db=18.0, lead=9000, tail=6000)["iq"].tofile(u8path)
    r = subprocess.run([EXE, path, iqpath, softpath, u8path], capture_output=True, text=True, timeout=300)
    os.remove(path); os.remove(iqpath); os.remove(softpath); os.remove(u8path)
    assert r.returncode == 0, r.stderr
    got = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    g = _lcg_stream(12345)
    draw = lambda n: np.array([next(g) % 255 - 127 for _ in range(n)], np.int16)
    assert int(got["viterbi768"], 16) == _fnv(port.viterbi(768, draw(4 * 774)))
    v = draw(96 * 64)
    assert int(got["eep128_3A"], 16) == _fnv(port.eep_deconvolve(128, 0o103, v))
    assert int(got["uep128_3"], 16) == _fnv(port.uep_deconvolve(128, 3, v))
    frags = np.stack([draw(12 * 64) for _ in range(20)])
    out = port.msc_backend(frags, 16, 1, 0o103)
    acc = 0
    for i, blk in enumerate(out):
        acc ^= (_fnv(blk) + i + 1) & 0xFFFFFFFFFFFFFFFF
    frames, h = got["backend"].split()
    assert int(frames) == 4 and int(h, 16) == acc
    bits = np.array([next(g) & 1 for _ in range(40 * 768)], np.uint8)
    table = port.fig01_scan(bits, np.ones((40, 3), np.uint8))
    assert int(got["fig01"], 16) == _fnv(table)
    sfs, info = port.dabplus(32).process(cifs)
    acc = 0
    for sf, k in zip(sfs, info):
        acc ^= (_fnv(sf) + k[0] * 1315423911 + k[1]) & 0xFFFFFFFFFFFFFFFF
    count, h = got["dabplus"].split()
    assert int(count) == len(info) == 6 and int(h, 16) == acc
    # the stream through ofdmProcessor / ficHandler / mscHandler: every CRC-clean FIB (with its ficno) and every frame
    nfr, nfib, fh, nmsc, mh, ratio = got["stream"].split()
    n = int(nfr)
    assert len(finfo) - n in (0, 1) and n >= 14
    bits, crc = port.fic_frames(1, sym[:n])
    M = 0xFFFFFFFFFFFFFFFF
    facc, cnt = 0, 0
    for g in range(4 * n):
        for f in range(3):
            if crc[g][f]:
                facc = (facc * 31 + _fnv(bits[g][256 * f:256 * f + 256]) + (g % 4)) & M
                cnt += 1
    assert int(nfib) == cnt and cnt > 50 and int(fh, 16) == facc
    s0 = mod.sub[0]
    frames = port.msc_backend(port.msc_slice(1, sym[:n], s0.startAddr, s0.length), s0.bitRate, s0.uepFlag, s0.protLevel)
    macc = 0
    for blk in frames:
        macc = (macc * 31 + _fnv(blk)) & M
    assert int(nmsc) == len(frames) > 30 and int(mh, 16) == macc
    assert int(ratio) == 100 * cnt // (12 * n)
    # dabgpu_group_decode from C++ (two GPUs when the box has them, else two handles on one): equal to one handle, both schemes
    for scheme in (1, 0):
        used, n1, n2, h1, h2, dev1 = got["group_scheme%d" % scheme].split()
        assert int(n1) == int(n2) >= 56 and h1 == h2, (scheme, got["group_scheme%d" % scheme])
        assert int(used) == scheme
    # ficHandler::process_ficBlock / mscHandler::process_mscBlock fed symbol by symbol with the oracle's soft bits of ALL
    # its frames: the same FIBs (with ficno) and frames as the oracle's FIC / MSC chain; behind mscHandler once
    # dabConcurrent (16-CIF warm-up) and once dabSerial (15: one frame more at the start, the rest identical)
    na = len(finfo)
    bits, crc = port.fic_frames(1, sym)
    facc, cnt = 0, 0
    for g in range(4 * na):
        for f in range(3):
            if crc[g][f]:
                facc = (facc * 31 + _fnv(bits[g][256 * f:256 * f + 256]) + (g % 4)) & M
                cnt += 1
    frames = port.msc_backend(port.msc_slice(1, sym, s0.startAddr, s0.length), s0.bitRate, s0.uepFlag, s0.protLevel)
    macc = 0
    for blk in frames:
        macc = (macc * 31 + _fnv(blk)) & M
    for key, extra in (("handlers_concurrent", 0), ("handlers_serial", 1)):
        nfib, fh, nmsc, mh, mh_skip1, ratio = got[key].split()
        assert int(nfib) == cnt and int(fh, 16) == facc, key
        assert int(nmsc) == len(frames) + extra, key
        assert int(mh_skip1 if extra else mh, 16) == macc, key
        assert int(ratio) == 100 * cnt // (12 * na), key
