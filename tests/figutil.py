"""Test-side FIG 0/1 encoder (ETSI EN 300 401 6.2.1 layout, written from the standard, not from the reference's
parser): builds FIBs that carry a known sub-channel organisation."""
import numpy as np

import dabmod


def _bits(v, n):
    return [(v >> (n - 1 - i)) & 1 for i in range(n)]


def fig01(entries):
    """entries: ('short', SubChId, StartAddr, tableIndex) | ('long', SubChId, StartAddr, option, level(1..4), size) -> bit list"""
    body = []
    for e in entries:
        if e[0] == "short":
            body += _bits(e[1], 6) + _bits(e[2], 10) + [0, 0] + _bits(e[3], 6)
        else:
            body += _bits(e[1], 6) + _bits(e[2], 10) + [1] + _bits(e[3], 3) + _bits(e[4] - 1, 2) + _bits(e[5], 10)
    nbytes = 1 + len(body) // 8
    return _bits(0, 3) + _bits(nbytes, 5) + [0, 0, 0] + _bits(1, 5) + body


def fib(figs, corrupt=False):
    """256 bits: the FIGs, 0xFF padding (FIG type 7 = end marker), CRC"""
    b = sum(figs, [])
    assert len(b) <= 240
    b = np.array(b + [1] * (240 - len(b)), np.uint8)
    out = np.concatenate([b, dabmod.crc16(b)])
    if corrupt:
        out[5] ^= 1
    return out


def groups(fibs):
    """list of FIBs (multiple of 3) -> bits[ngroups, 768]"""
    a = np.array(fibs, np.uint8)
    assert a.shape[0] % 3 == 0
    return a.reshape(-1, 768)
