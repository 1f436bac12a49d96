"""CPU: the DAB+ super-frame layer of the oracle (Fire code, RS(120,110), super-frame sync state machine, AU table;
mp4processor.cpp:40-61, 107-275, firecode-checker.cpp, reed-solomon.cpp, galois.cpp) pinned three ways: against an
independent encoder written from ETSI TS 102 563 (tests/dabplus.py), against the reference's own compiled classes
(oracle/_ref, when present), and through golden vectors generated from them."""
import numpy as np
import pytest

import dabplus
import orc as orc_mod


def _oracles():
    out = [orc_mod.Oracle("port")]
    try:
        out.append(orc_mod.Oracle("ref"))
    except Exception:
        pass
    return out


def test_firecode_and_rs_against_independent_encoder():
    rng = np.random.default_rng(1)
    for o in _oracles():
        for _ in range(50):
            d = rng.integers(0, 256, 9)
            fc = dabplus.firecode(d)
            x = np.concatenate([[fc >> 8, fc & 255], d]).astype(np.uint8)
            assert o.firecode_check(x)
            x[int(rng.integers(0, 11))] ^= 1 << int(rng.integers(0, 8))
            assert not o.firecode_check(x)
        for nerr in (0, 1, 3, 5):
            for _ in range(20):
                data = rng.integers(0, 256, 110).astype(np.uint8)
                cw = np.array(dabplus.rs_encode(data), np.uint8)
                assert np.array_equal(o.rs_enc(data), cw)                 # the reference's encoder agrees with the standard's
                pos = rng.choice(120, nerr, replace=False)
                bad = cw.copy()
                bad[pos] ^= rng.integers(1, 256, nerr).astype(np.uint8)
                ret, out = o.rs_dec(bad)
                assert np.array_equal(out, data)
                assert ret == int((pos < 110).sum())                     # errors in the parity bytes are not counted (reed-solomon.cpp:218-219)


def test_rs_port_equals_reference_on_garbage():
    """beyond the correction radius the decoder's verdict and output are implementation detail: port == compiled reference"""
    os_ = _oracles()
    if len(os_) < 2:
        pytest.skip("no compiled reference in this checkout")
    rng = np.random.default_rng(2)
    for nerr in (6, 7, 9, 20, 120):
        for _ in range(60):
            data = rng.integers(0, 256, 110).astype(np.uint8)
            bad = np.array(dabplus.rs_encode(data), np.uint8)
            pos = rng.choice(120, nerr, replace=False)
            bad[pos] ^= rng.integers(1, 256, nerr).astype(np.uint8)
            a, b = os_[0].rs_dec(bad), os_[1].rs_dec(bad)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])


def _stream(bitRate, rng, nsf, junk_before=3, damage=True):
    blocks, sfs = [rng.integers(0, 2, (junk_before, 24 * bitRate), dtype=np.uint8)], []
    for i in range(nsf):
        sf, coded, starts = dabplus.make_superframe(bitRate, rng, dac_rate=i & 1, sbr=(i >> 1) & 1)
        if damage and i % 3 == 1:                                # a few byte errors per column, inside the RS radius -- but not in
            for j in range(bitRate // 8):                        # the first 11 bytes: the Fire code is checked BEFORE the repair
                k = rng.choice(np.arange(3, 120), 4, replace=False)       # (mp4processor.cpp:135), such a frame is simply lost
                coded[j + k * (bitRate // 8)] ^= rng.integers(1, 256, 4).astype(np.uint8)
        if damage and i == 7:
            coded[4] ^= 0x10                                     # one bit inside the Fire code's reach: lost although RS could repair it
        if damage and i == 5:                                    # one super frame beyond repair: sync slides CIF by CIF
            coded[::2] ^= 0x5A
        blocks.append(dabplus.to_cif_bits(coded, bitRate))
        sfs.append((sf, starts))
    return np.concatenate(blocks), sfs


@pytest.mark.parametrize("bitRate", [32, 72, 128])
def test_superframe_sync_and_repair(bitRate):
    rng = np.random.default_rng(bitRate)
    bits, sfs = _stream(bitRate, rng, 9)
    res = []
    for o in _oracles():
        dp = o.dabplus(bitRate)
        a = dp.process(bits[:7])                                 # ragged pieces: the object carries ring buffer and counters
        b = dp.process(bits[7:8])
        c = dp.process(bits[8:])
        sf = np.concatenate([a[0], b[0], c[0]])
        info = a[1] + b[1] + c[1]
        res.append((sf, info))
        good = [i for i in range(9) if i not in (5, 7)]
        assert len(info) == len(good)
        for (first, corrected, n, au, crc), i in zip(info, good):
            assert first == 3 + 5 * i and np.array_equal(sf[good.index(i)], sfs[i][0])
            assert n == len(sfs[i][1]) - 1 and list(au[:n + 1]) == sfs[i][1] and crc == (1 << n) - 1
            assert (corrected > 0) == (i % 3 == 1)
    if len(res) == 2:
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


@pytest.mark.parametrize("bitRate", [32, 72, 128])
def test_superframe_layer_equals_the_references_mp4processor(bitRate):
    """the oracle's restated super-frame front against the reference's OWN mp4Processor (mp4processor.cpp compiled unmodified,
    AAC decoder and PAD handler replaced by recorders): the same super frames are accepted at the same CIFs, with the same
    corrected bytes, access-unit tables and AU CRC verdicts -- on a damaged stream and on noise with planted frames"""
    os_ = _oracles()
    if len(os_) < 2:
        pytest.skip("no compiled reference in this checkout")
    port, ref = os_
    rng = np.random.default_rng(100 + bitRate)
    bits, sfs = _stream(bitRate, rng, 9)
    more, _ = _stream(bitRate, rng, 4, junk_before=7, damage=False)
    more[rng.integers(0, more.shape[0], 3)] ^= 1                          # three CIFs inverted: their super frames fail
    stream = np.concatenate([bits, rng.integers(0, 2, (11, 24 * bitRate), dtype=np.uint8), more])
    a_sf, a_info = port.dabplus(bitRate).process(stream)
    b_sf, b_info = ref.ref_dabplus_run(stream, bitRate)
    assert len(a_info) == len(b_info) >= 8
    assert np.array_equal(a_sf, b_sf)
    for (first, corrected, n, au, crc), (bfirst, bn, bau, bcrc) in zip(a_info, b_info):
        assert (first, n, tuple(au[:n + 1]), crc) == (bfirst, bn, tuple(bau[:bn + 1]), bcrc)
