"""GPU: the DAB+ super-frame layer (dabgpu_dabplus_*) against the oracle -- bit-exact super frames, identical sync
decisions, RS error counts, AU tables and CRC verdicts; streams with repairable and unrepairable damage, junk between
super frames, ragged call sizes, and pure noise (where the decoder's verdict beyond the correction radius must still
equal the reference's)."""
import numpy as np
import pytest

import dabplus
from util import engine_pkg

pytestmark = pytest.mark.gpu


def _stream(bitRate, rng, nsf, junk_every=4):
    blocks = [rng.integers(0, 2, (2, 24 * bitRate), dtype=np.uint8)]
    R = bitRate // 8
    for i in range(nsf):
        kind = i % 7
        sf, coded, starts = dabplus.make_superframe(bitRate, rng, dac_rate=i & 1, sbr=(i >> 1) & 1, mangle_table=kind == 6)
        if kind in (1, 4):                                       # inside the RS radius, outside the Fire code's reach
            for j in range(R):
                ne = int(rng.integers(1, 6))
                k = rng.choice(np.arange(3, 120), ne, replace=False)
                coded[j + k * R] ^= rng.integers(1, 256, ne).astype(np.uint8)
        elif kind == 5:                                          # one column beyond repair
            k = rng.choice(np.arange(3, 120), 9, replace=False)
            coded[1 + k * R] ^= rng.integers(1, 256, 9).astype(np.uint8)
        blocks.append(dabplus.to_cif_bits(coded, bitRate))
        if i % junk_every == junk_every - 1:
            blocks.append(rng.integers(0, 2, (int(rng.integers(1, 4)), 24 * bitRate), dtype=np.uint8))
    return np.concatenate(blocks)


@pytest.mark.parametrize("bitRate,nsf,seed", [(32, 12, 1), (48, 25, 2), (128, 40, 3), (192, 9, 4)])
def test_superframes_match_oracle(port, bitRate, nsf, seed):
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    rng = np.random.default_rng(seed)
    bits = _stream(bitRate, rng, nsf)
    want_sf, want_info = port.dabplus(bitRate).process(bits)
    assert len(want_info) >= nsf // 2
    dp = pkg.binding.DabPlus(eng, bitRate)
    got_sf, got_info = dp.process(bits)
    assert got_info == want_info and np.array_equal(got_sf, want_sf)
    # the same stream in ragged pieces through a second object
    dp2 = pkg.binding.DabPlus(eng, bitRate)
    cuts = sorted(set([0, 1, 3, 4, 9, 10, 37, bits.shape[0] // 2, bits.shape[0]]))
    cuts = [c for c in cuts if c <= bits.shape[0]]
    parts = [dp2.process(bits[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    assert sum((p[1] for p in parts), []) == want_info
    assert np.array_equal(np.concatenate([p[0] for p in parts]), want_sf)
    dp.close(); dp2.close(); eng.close()


def test_noise_and_false_fire_code_hits(port):
    """random blocks whose first 11 bytes are forced to pass the Fire code: the RS decoder runs on garbage, its failure
    modes (root count mismatch, zero denominator, 'repairs' inside the padding) must equal the reference's"""
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    rng = np.random.default_rng(9)
    bitRate = 64
    ncif = 400
    by = rng.integers(0, 256, (ncif, 3 * bitRate)).astype(np.uint8)
    for r in range(0, ncif, 3):
        fc = dabplus.firecode(by[r, 2:11])
        by[r, 0], by[r, 1] = fc >> 8, fc & 255
    # a few rows with only a handful of byte errors relative to a valid code word layout: near-miss decodes
    bits = np.unpackbits(by, axis=1)
    want = port.dabplus(bitRate).process(bits)
    dp = pkg.binding.DabPlus(eng, bitRate)
    got = dp.process(bits)
    assert got[1] == want[1] and np.array_equal(got[0], want[0])
    dp.close(); eng.close()


def test_rs_columns_vs_oracle_beyond_radius(port):
    """single super frames with 0..12 byte errors in every column: the GPU's per-column verdict decides acceptance
    exactly like the oracle's"""
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    rng = np.random.default_rng(10)
    bitRate, R = 40, 5
    rows = []
    for t in range(60):
        sf, coded, _ = dabplus.make_superframe(bitRate, rng)
        ne = t % 13
        for j in range(R):
            if ne:
                k = rng.choice(np.arange(3, 120), ne, replace=False)
                coded[j + k * R] ^= rng.integers(1, 256, ne).astype(np.uint8)
        rows.append(dabplus.to_cif_bits(coded, bitRate))
    bits = np.concatenate(rows)
    want = port.dabplus(bitRate).process(bits)
    dp = pkg.binding.DabPlus(eng, bitRate)
    got = dp.process(bits)
    assert len(want[1]) >= 25
    assert got[1] == want[1] and np.array_equal(got[0], want[0])
    dp.close(); eng.close()
