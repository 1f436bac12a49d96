"""GPU: FIG 0/1 sub-channel table extraction (dabgpu_fig01_scan, the stream engine's running table) against the
oracle's sequential restatement of fib-processor.cpp:123-158, 278-347 -- bit-exact, including the order-dependent
"last write wins" semantics that the device resolves with ordered atomics."""
import numpy as np
import pytest

import dabmod
import figutil
from util import engine_pkg

pytestmark = pytest.mark.gpu


def test_fig01_handmade(port):
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    pad = figutil.fib([])
    fibs = [figutil.fib([figutil.fig01([("short", 3, 0, 35), ("long", 7, 96, 0, 3, 96)])]), pad,
            figutil.fib([figutil.fig01([("long", 9, 200, 1, 2, 84)]), figutil.fig01([("short", 1, 500, 63)])]),
            figutil.fib([figutil.fig01([("long", 7, 97, 6, 1, 5)])]), figutil.fib([figutil.fig01([("short", 3, 1, 0)])], corrupt=True), pad]
    g = figutil.groups(fibs)
    crc = np.array([[port.check_crc(g[i, 256 * j:256 * j + 256]) for j in range(3)] for i in range(g.shape[0])], np.uint8)
    got = eng.fig01_scan(g, crc)
    assert np.array_equal(got, port.fig01_scan(g, crc))
    assert got[7].tolist() == [1, 97, 96, 1, 0o103, 128] and got[3].tolist() == [1, 0, 96, 0, 3, 128]
    assert np.array_equal(eng.fig01_scan(g[:0], crc[:0]), np.zeros((64, 6), np.int32))
    eng.close()


@pytest.mark.parametrize("ngroups,seed", [(1, 1), (7, 2), (4096, 3)])
def test_fig01_random_fibs_match_oracle(port, ngroups, seed):
    """random FIB bodies exercise every path of the parser (lengths running past the FIB, all option values, 64
    sub-channel ids rewritten thousands of times): the final table must be the sequential one"""
    pkg = engine_pkg()
    eng = pkg.DabGpu(mode=1)
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 2, (ngroups, 768), dtype=np.uint8)
    g[:, ::17] &= rng.integers(0, 2, g[:, ::17].shape, dtype=np.uint8)          # more zeros: more FIG type 0
    crc = (rng.random((ngroups, 3)) < 0.8).astype(np.uint8)
    assert np.array_equal(eng.fig01_scan(g, crc), port.fig01_scan(g, crc))
    eng.close()


def test_stream_engine_running_table(port):
    """the table after a stream decode = the oracle's scan of the same FIC bits in order, across two calls"""
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103)], 21)
    iq = mod.generate(14, cfo_hz=300.0, snr_db=22.0, lead=2000, tail=6000)["iq"]
    eng = pkg.DabGpu(mode=1); eng.set_subchannels([(0, 96, 128, 1, 0o103)])
    half = (iq.size // 4) * 2
    r1 = eng.decode(iq[:half], eng.alloc_result(20))
    t1 = eng.subch_table()
    r2 = eng.decode(iq[half:], eng.alloc_result(20))
    t2 = eng.subch_table()
    assert r1.nframes > 0 and r2.nframes > 0
    want1 = port.fig01_scan(r1.fic_bits, r1.fic_crc)
    assert np.array_equal(t1, want1)
    assert np.array_equal(t2, port.fig01_scan(r2.fic_bits, r2.fic_crc, want1.copy()))
    eng.close()
