"""Test-side DAB+ super-frame encoder, written from ETSI TS 102 563 (not from the reference's decoder): Fire code
over bytes 2..10, access-unit table and per-AU CRC-16, RS(120,110) over GF(2^8) (poly 0x11D, roots alpha^0..alpha^9)
on the column-interleaved super frame, split into five CIF blocks of 24*bitRate bits."""
import numpy as np

# ---- GF(256), primitive polynomial x^8+x^4+x^3+x^2+1 ----
EXP = np.zeros(512, np.int64)
LOG = np.zeros(256, np.int64)
_x = 1
for _i in range(255):
    EXP[_i] = _x
    LOG[_x] = _i
    _x <<= 1
    if _x & 0x100:
        _x ^= 0x11D
EXP[255:510] = EXP[:255]


def gmul(a, b):
    return 0 if a == 0 or b == 0 else int(EXP[LOG[a] + LOG[b]])


def _genpoly():
    g = [1]
    for r in range(10):                                       # (x - alpha^r), r = 0..9
        ng = [0] * (len(g) + 1)
        for i, c in enumerate(g):
            ng[i] ^= gmul(c, int(EXP[r]))
            ng[i + 1] ^= c
        g = ng
    return g                                                  # g[i] = coefficient of x^i, monic degree 10


GEN = _genpoly()


def rs_encode(data110):
    """systematic RS(120,110): -> 120 bytes, parity = data(x) x^10 mod g(x), highest power first"""
    rem = [0] * 10
    for d in data110:
        fb = int(d) ^ rem[0]
        rem = rem[1:] + [0]
        if fb:
            for j in range(10):
                rem[j] ^= gmul(fb, GEN[9 - j])
    return list(map(int, data110)) + rem


def crc16_ccitt(data):
    acc = 0xFFFF
    for b in data:
        acc ^= int(b) << 8
        for _ in range(8):
            acc = ((acc << 1) ^ 0x1021) & 0xFFFF if acc & 0x8000 else (acc << 1) & 0xFFFF
    return (~acc) & 0xFFFF


def firecode(data9):
    """CRC with g(x) = (x^11+1)(x^5+x^3+x^2+x+1) = x^16+x^14+x^13+x^12+x^11+x^5+x^3+x^2+x+1, zero start"""
    acc = 0
    for b in data9:
        acc ^= int(b) << 8
        for _ in range(8):
            acc = ((acc << 1) ^ 0x782F) & 0xFFFF if acc & 0x8000 else (acc << 1) & 0xFFFF
    return acc


def make_superframe(bitRate, rng, dac_rate=1, sbr=0, mangle_table=False):
    """-> (data[110*R] before RS, coded[120*R] as transmitted = 5 CIF blocks of 3*bitRate bytes)"""
    R = bitRate // 8
    size = 110 * R
    table = {(0, 0): (4, 8), (0, 1): (2, 5), (1, 0): (6, 11), (1, 1): (3, 6)}
    if (size - table[(dac_rate, sbr)][1]) // table[(dac_rate, sbr)][0] > 940:      # access units would exceed 960 bytes at this
        dac_rate, sbr = 1, 0                                                        # bit rate: use the six-AU layout instead
    n, first = table[(dac_rate, sbr)]
    step = (size - first) // n                                # roughly equal access units with some jitter
    jit = max(0, min(step // 4, (955 - step) // 2))             # an access unit must stay below 960 bytes (mp4processor.cpp:246)
    cuts = [first + k * step + int(rng.integers(-jit, jit + 1)) for k in range(1, n)]
    starts = [first] + cuts + [size]
    assert all(3 < b - a < 960 for a, b in zip(starts[:-1], starts[1:])), starts
    sf = np.zeros(size, np.int64)
    sf[2] = (dac_rate << 6) | (sbr << 5) | (int(rng.integers(0, 2)) << 4) | int(rng.integers(0, 8))
    nib = []
    for s in starts[1:-1]:
        nib += [(s >> 8) & 15, (s >> 4) & 15, s & 15]
    if len(nib) % 2:
        nib.append(0)
    for i in range(0, len(nib), 2):
        sf[3 + i // 2] = (nib[i] << 4) | nib[i + 1]
    for a, b in zip(starts[:-1], starts[1:]):
        body = rng.integers(0, 256, b - a - 2)
        c = crc16_ccitt(body)
        sf[a:b - 2] = body
        sf[b - 2], sf[b - 1] = c >> 8, c & 255
    if mangle_table:                                          # second AU 'starts' before the first: valid Fire code and RS,
        sf[3], sf[4] = 0, (1 << 4) | (sf[4] & 15)             # but processSuperframe must refuse it (mp4processor.cpp:241-244)
    fc = firecode(sf[2:11])
    sf[0], sf[1] = fc >> 8, fc & 255
    coded = np.zeros(120 * R, np.int64)
    for j in range(R):
        coded[j::R] = rs_encode(sf[j::R])
    return sf.astype(np.uint8), coded.astype(np.uint8), starts


def to_cif_bits(coded, bitRate):
    """120*R bytes -> [5][24*bitRate] one bit per byte, MSB first (mp4processor.cpp:114-119)"""
    return np.unpackbits(coded.astype(np.uint8)).reshape(5, 24 * bitRate)
