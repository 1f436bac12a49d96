"""CPU: the oracle's restatement of the FIG 0/1 path (fib-processor.cpp:123-158, 278-347) against FIBs built by an
independent encoder (tests/figutil.py): the sub-channel organisation that went in comes out, later FIBs override
earlier ones field group by field group, FIBs with a bad CRC are ignored."""
import numpy as np

import figutil

PROT = {35: (96, 3, 128), 0: (16, 5, 32), 63: (416, 1, 384)}     # table index -> (size, level, bit rate), ETSI EN 300 401 table 8


def _crc(port, g):
    return np.array([[port.check_crc(g[i, 256 * j:256 * j + 256]) for j in range(3)] for i in range(g.shape[0])], np.uint8)


def test_fig01_roundtrip(port):
    pad = figutil.fib([])
    f1 = figutil.fib([figutil.fig01([("short", 3, 0, 35), ("long", 7, 96, 0, 3, 96)])])
    f2 = figutil.fib([figutil.fig01([("long", 9, 200, 1, 2, 84), ("short", 63, 863, 0)]), figutil.fig01([("short", 1, 500, 63)])])
    g = figutil.groups([f1, pad, f2])
    crc = _crc(port, g)
    assert crc.all()
    t = port.fig01_scan(g, crc)
    assert t[3].tolist() == [1, 0, 96, 0, 3, 128]
    assert t[7].tolist() == [1, 96, 96, 1, 0o103, 128]              # EEP 3-A: size / 6 * 8
    assert t[9].tolist() == [1, 200, 84, 1, 0o202, 128]             # EEP 2-B: size / 21 * 32
    assert t[63].tolist() == [1, 863, 16, 0, 5, 32]
    assert t[1].tolist() == [1, 500, 416, 0, 1, 384]
    assert t[:, 0].sum() == 5


def test_fig01_last_write_wins_per_field_group(port):
    a = figutil.fib([figutil.fig01([("short", 5, 10, 35)])])
    b = figutil.fib([figutil.fig01([("long", 5, 20, 5, 1, 77)])])      # option 5: only StartAddr / uepFlag are written
    c = figutil.fib([figutil.fig01([("long", 5, 30, 0, 4, 64)])], corrupt=True)
    g = figutil.groups([a, b, c])
    crc = _crc(port, g)
    assert crc.tolist() == [[1, 1, 0]]
    t = port.fig01_scan(g, crc)
    assert t[5].tolist() == [1, 20, 96, 1, 3, 128]                   # the reference's mixed state, reproduced
    t = port.fig01_scan(figutil.groups([c, a, a]), np.array([[1, 1, 1]], np.uint8), t)   # table carries over; c taken as clean
    assert t[5].tolist() == [1, 10, 96, 0, 3, 128]


def test_fig01_random_fibs_do_not_run_away(port):
    rng = np.random.default_rng(3)
    g = rng.integers(0, 2, (64, 768), dtype=np.uint8)
    t = port.fig01_scan(g, np.ones((64, 3), np.uint8))
    assert t[:, 0].sum() > 0 and (t[:, 1] < 1024).all()
