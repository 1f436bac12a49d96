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


# ---- Tier C: the restated handler loops against the reference's own handler classes, compiled unmodified with Qt stand-ins
# (oracle/ref_shim/ref_tierc.cpp: fic-handler.cpp, fib-processor.cpp, msc-handler.cpp, dab-concurrent.cpp) ----
@pytest.mark.parametrize("mode", [1, 2, 4])
def test_fic_handler_class(port, ref, mode):
    """ficHandler::process_ficBlock (queue + thread + regrouping into 2304-bit words + depuncturing + Viterbi + PRBS + CRC)"""
    p = port.mode_params(mode)
    rng = np.random.default_rng(40 + mode)
    mod = dabmod.Modulator(port, mode, [], 7 + mode)
    mod.wellformed_fibs = True          # (ficHandler feeds every CRC-clean FIB to the reference's FIB parser, which cannot take random bits)
    nfr = 6
    truth = mod.frame_bits(nfr)
    soft = (2 * truth["bits"].astype(np.int16) - 1) * 90 + rng.integers(-120, 121, truth["bits"].shape).astype(np.int16)
    soft = np.clip(soft, -127, 127).astype(np.int16)
    a_bits, a_crc = port.fic_frames(mode, soft)
    b_bits, b_crc = ref.ref_fic_frames(mode, soft)
    assert np.array_equal(a_bits, b_bits) and np.array_equal(a_crc, b_crc)
    assert 0 < a_crc.mean() <= 1.0
    noise = rng.integers(-127, 128, (3, p.L - 1, 2 * p.K)).astype(np.int16)
    a_bits, a_crc = port.fic_frames(mode, noise)
    b_bits, b_crc = ref.ref_fic_frames(mode, noise)
    assert np.array_equal(a_bits, b_bits) and np.array_equal(a_crc, b_crc)


def test_fib_processor_class(port, ref):
    """fib_processor::process_FIB -> FIG 0/1 sub-channel table, on encoder-built FIBs and on random bits (every FIG type and
    extension the random bits happen to spell goes through the reference's real dispatch)"""
    import figutil
    f1 = figutil.fib([figutil.fig01([("short", 3, 0, 35), ("long", 7, 96, 0, 3, 96)])])
    f2 = figutil.fib([figutil.fig01([("long", 9, 200, 1, 2, 84), ("short", 63, 863, 0)]), figutil.fig01([("short", 1, 500, 63)])])
    f3 = figutil.fib([figutil.fig01([("long", 5, 20, 5, 1, 77)])])
    g = figutil.groups([f1, figutil.fib([]), f2, f3, f1, f2])
    crc = np.array([[port.check_crc(g[i, 256 * j:256 * j + 256]) for j in range(3)] for i in range(g.shape[0])], np.uint8)
    a = port.fig01_scan(g, crc)
    b = ref.ref_fig01_scan(g, crc)
    assert np.array_equal(a[:, 1:], b[:, 1:])
    # many random but WELL-FORMED FIBs (random bits with a good CRC make the reference's parser read wild or loop forever --
    # nothing a CRC-protected broadcast ever contains): random FIG 0/1 entry lists, several FIGs per FIB, random CRC verdicts
    rng = np.random.default_rng(9)
    table_a = table_b = None
    for rep in range(6):
        fibs = []
        for k in range(3 * 40):
            figs, room = [], 240
            while room > 40 and rng.integers(0, 3):
                ent = [("short", int(rng.integers(0, 64)), int(rng.integers(0, 864)), int(rng.integers(0, 64))) if rng.integers(0, 2) else
                       ("long", int(rng.integers(0, 64)), int(rng.integers(0, 864)), int(rng.integers(0, 2)), int(rng.integers(1, 5)), int(rng.integers(6, 400)))
                       for _ in range(int(rng.integers(1, 4)))]
                f = figutil.fig01(ent)
                if len(f) > room:
                    break
                figs.append(f); room -= len(f)
            fibs.append(figutil.fib(figs))
        g = figutil.groups(fibs)
        ok = rng.integers(0, 4, (g.shape[0], 3)).astype(bool).astype(np.uint8)
        table_a = port.fig01_scan(g, ok, table_a)
        table_b = ref.ref_fig01_scan(g, ok, table_b if table_b is not None else None)
        assert np.array_equal(table_a[:, 1:], table_b[:, 1:]), rep
    assert table_a[:, 0].sum() > 40


@pytest.mark.parametrize("mode,sub", [(1, (0, 96, 128, 1, 0o103)), (1, (100, 84, 128, 1, 0o202)), (1, (300, 96, 128, 0, 3)), (2, (10, 48, 64, 1, 0o103)), (4, (0, 84, 80, 0, 1))])
def test_msc_handler_and_dab_concurrent_classes(port, ref, mode, sub):
    """mscHandler::process_mscBlock (CIF assembly, sub-channel slice) + dabConcurrent (ring buffer, thread, 16-CIF time
    de-interleaver, warm-up of 16 CIFs, EEP / UEP deconvolution, energy dispersal)"""
    p = port.mode_params(mode)
    startAddr, Length, bitRate, uepFlag, protLevel = sub
    rng = np.random.default_rng(mode * 100 + startAddr)
    nfr = 8 if mode == 1 else (24 if mode == 2 else 12)
    sym = rng.integers(-127, 128, (nfr, p.L - 1, 2 * p.K)).astype(np.int16)
    a = port.msc_backend(port.msc_slice(mode, sym, startAddr, Length), bitRate, uepFlag, protLevel)
    b = ref.ref_msc_run(mode, sym, startAddr, Length, bitRate, uepFlag, protLevel)
    assert a.shape == b.shape and a.shape[0] == nfr * p.cifsPerFrame - 16
    assert np.array_equal(a, b)


@pytest.mark.parametrize("mode,cfo,snr,lead,sub", [
    (1, 137.0, 15.0, 30000, (0, 96, 128, 1, 0o103)),
    (1, -2310.0, 12.0, 77777, (200, 96, 128, 0, 3)),
    (2, -1130.0, 16.0, 20000, (10, 48, 64, 1, 0o103)),
    (4, -730.0, 20.0, 41000, (96, 84, 128, 1, 0o202)),
])
def test_whole_receive_chain_classes(port, ref, mode, cfo, snr, lead, sub):
    """The reference's OWN receive chain -- ofdmProcessor::run (null-symbol search, findIndex, NCO, coarse and fine AFC, the
    symbol loop), ofdmDecoder, ficHandler, mscHandler, dabConcurrent, every class compiled unmodified and wired as gui.cpp wires
    them -- against the oracle's restated chain on the same raw IQ: every FIC group, CRC flag and MSC block equal.  This is what
    pins dab_oracle.c's restatement of ofdm-processor.cpp:247-474 (the FFT under both is the labelled FFTW stand-in)."""
    p = port.mode_params(mode)
    nfr = {1: 30, 2: 70, 4: 40}[mode]
    mod = dabmod.Modulator(port, mode, [(sub[0], sub[2], sub[3], sub[4])], 500 + mode)
    assert mod.sub[0].length == sub[1]
    mod.wellformed_fibs = True                               # the reference's FIB parser runs on every CRC-clean FIB of this chain
    tr = mod.generate(nfr, cfo_hz=cfo, snr_db=snr, lead=lead, tail=9000)
    iq = tr["iq"]
    sym, info = port.ofdm_run(mode, iq, nfr + 4)
    a_fic, a_crc = port.fic_frames(mode, sym)
    a_msc = port.msc_backend(port.msc_slice(mode, sym, sub[0], sub[1]), sub[2], sub[3], sub[4])
    b_fic, b_crc, b_msc, state = ref.ref_receive(mode, iq, sub, max_frames=nfr + 4)
    g, cpf = p.ficGroups, p.cifsPerFrame
    # the oracle counts a frame whose trailing null symbol is cut off by the end of the input; the reference, waiting for those
    # samples, has handed over that frame's symbols all the same -- and its backend keeps the last CIF in the ring buffer
    assert len(info) >= nfr // 2 and (len(info) - 1) * g <= b_fic.shape[0] <= (len(info) + 1) * g, (len(info), b_fic.shape)   # (the reference hands over symbol by symbol)
    n = min(a_fic.shape[0], b_fic.shape[0])
    assert np.array_equal(a_fic[:n], b_fic[:n]) and np.array_equal(a_crc[:n], b_crc[:n])
    assert a_crc[n - 2 * g:n].mean() > 0.9                    # (the first frames are decoded while the correctors still move)
    m = min(a_msc.shape[0], b_msc.shape[0])
    assert m >= (len(info) - 1) * cpf - 17 and m > 0, (m, len(info))
    assert np.array_equal(a_msc[:m], b_msc[:m])
    last = info[min(len(info), b_fic.shape[0] // g) - 1]
    assert state[0] == last.coarse or abs(state[0] - last.coarse) % p.carrierDiff == 0


@pytest.mark.parametrize("sub", [(96, 128, 1, 0o103), (84, 128, 1, 0o202), (96, 128, 0, 3)])
def test_dab_serial_class(port, ref, sub):
    """dabSerial (the backend without a thread) starts decoding one CIF EARLIER than dabConcurrent (15 against 16 CIFs of warm-up,
    dab-serial.cpp:126-129 / dab-concurrent.cpp:172-175): from its second block on its output is the concurrent backend's, the
    first is the decode of the de-interleaver's 16th output row -- which is how host/dab_adapters.h models it"""
    Length, bitRate, uepFlag, protLevel = sub
    rng = np.random.default_rng(Length + protLevel)
    ncif = 40
    frags = rng.integers(-127, 128, (ncif, Length * 64)).astype(np.int16)
    serial = ref.ref_serial_run(frags, bitRate, uepFlag, protLevel)
    conc = port.msc_backend(frags, bitRate, uepFlag, protLevel)
    assert serial.shape[0] == ncif - 15 and conc.shape[0] == ncif - 16
    assert np.array_equal(serial[1:], conc)
    row = port.time_deinterleave(frags)[15]
    first = (port.uep_deconvolve(bitRate, protLevel, row) if uepFlag == 0 else port.eep_deconvolve(bitRate, protLevel, row)) ^ port.prbs(24 * bitRate)
    assert np.array_equal(serial[0], first)
