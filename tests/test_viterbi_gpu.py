"""GPU parity of the Viterbi group through the C ABI: bit-exact against the
oracle on identical soft bits."""
import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 2], ids=["warp", "simd"])
def eng(request):
    """both Viterbi kernels are forced in turn: 1 = warp per code word, 2 = code word per thread (SIMD)"""
    e = engine_pkg().DabGpu(mode=1, viterbi_path=request.param)
    yield e
    e.close()


def _noisy(rng, bits, flip=0.08, erase=0.1):
    return dabmod.soft_from_bits(bits, rng, amp=70, flip=flip, erase=erase, jitter=57)


@pytest.mark.parametrize("frameBits", [8, 24, 192, 768, 3072, 9216])
def test_viterbi_random_soft_bits(eng, port, frameBits):
    """uniform random soft bits: every tie-break and the weak 63/0 start bias get exercised"""
    rng = np.random.default_rng(frameBits)
    nblocks = 37 if frameBits <= 3072 else 5
    soft = rng.integers(-127, 128, (nblocks, 4 * (frameBits + 6))).astype(np.int16)
    soft[0] = 0                       # all erasures
    soft[1] = 127
    soft[2] = -127
    soft[3, ::2] = -300               # out of range values clamp like the reference (viterbi.cpp:232-233)
    soft[3, 1::2] = 300
    got = eng.viterbi(soft, frameBits)
    for i in range(nblocks):
        assert np.array_equal(got[i], port.viterbi(frameBits, soft[i])), i


def test_viterbi_small_alphabet_ties(eng, port):
    """soft bits from {-1,0,1}: almost every ACS is a tie -> the strict '>' rule decides"""
    rng = np.random.default_rng(7)
    soft = rng.integers(-1, 2, (64, 4 * 774)).astype(np.int16)
    got = eng.viterbi(soft, 768)
    for i in range(64):
        assert np.array_equal(got[i], port.viterbi(768, soft[i]))


def test_viterbi_encoded_roundtrip(eng):
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2, (50, 3072), dtype=np.uint8)
    soft = _noisy(rng, dabmod.conv_encode(bits), flip=0.03, erase=0.05)
    assert np.array_equal(eng.viterbi(soft, 3072), bits)


EEP = [(br, 1, lvl) for br in (8, 32, 64, 128, 192) for lvl in (0o101, 0o102, 0o103, 0o104)] + \
      [(br, 1, lvl) for br in (32, 64, 128, 384) for lvl in (0o201, 0o202, 0o203, 0o204)]


@pytest.mark.parametrize("bitRate,uepFlag,protLevel", EEP)
def test_eep_profiles(eng, port, bitRate, uepFlag, protLevel):
    rng = np.random.default_rng(bitRate * 1000 + protLevel)
    mask = dabmod.puncture_mask(port, bitRate, uepFlag, protLevel)
    size = -(-int(mask.sum()) // 64) * 64
    v = rng.integers(-127, 128, (3, size)).astype(np.int16)
    bits = rng.integers(0, 2, 24 * bitRate, dtype=np.uint8)
    v[0, :mask.sum()] = _noisy(rng, dabmod.conv_encode(bits)[mask], flip=0.02, erase=0.0)
    got = eng.protect_decode(bitRate, uepFlag, protLevel, v)
    for i in range(3):
        assert np.array_equal(got[i], port.eep_deconvolve(bitRate, protLevel, v[i])), i


def test_uep_all_profiles(eng, port):
    """every row of the reference's UEP table (deconvolve.cpp:39-113)"""
    rng = np.random.default_rng(5)
    n = 0
    for bitRate in (32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384):
        for lvl in (1, 2, 3, 4, 5):
            if port.uep_profile(bitRate, lvl) is None:
                with pytest.raises(Exception):
                    eng.protect_decode(bitRate, 0, lvl, np.zeros((1, 64 * 416), np.int16))
                continue
            mask = dabmod.puncture_mask(port, bitRate, 0, lvl)
            size = -(-int(mask.sum()) // 64) * 64
            v = rng.integers(-127, 128, (2, size)).astype(np.int16)
            got = eng.protect_decode(bitRate, 0, lvl, v)
            for i in range(2):
                assert np.array_equal(got[i], port.uep_deconvolve(bitRate, lvl, v[i])), (bitRate, lvl)
            n += 1
    assert n == 60


def test_fic_decode(eng, port):
    rng = np.random.default_rng(3)
    mod = dabmod.Modulator(port, 1, [], 42)
    fibs, punct = mod.make_fic(8)                     # 32 code words with valid CRCs
    soft = _noisy(rng, punct, flip=0.03, erase=0.0)
    soft[5] = rng.integers(-127, 128, 2304)           # garbage -> CRC must fail like the reference's
    bits, crc = eng.fic_decode(soft)
    for i in range(soft.shape[0]):
        b, c = port.fic_decode(soft[i])
        assert np.array_equal(bits[i], b) and np.array_equal(crc[i], c), i
    assert crc[0].all() and np.array_equal(bits[0], fibs[0]) and not crc[5].all()


@pytest.mark.parametrize("sub", [(0, 128, 1, 0o103), (10, 128, 0, 3), (3, 32, 0, 5), (0, 64, 1, 0o202)])
def test_msc_backend_stateful(eng, port, sub):
    """dabConcurrent: de-interleave + warm-up + decode + dispersal, fed in ragged pieces"""
    startAddr, bitRate, uepFlag, protLevel = sub
    rng = np.random.default_rng(bitRate)
    sc = dabmod.SubChannel(port, startAddr, bitRate, uepFlag, protLevel)
    ncif = 45
    frags = rng.integers(-127, 128, (ncif, sc.fragmentSize)).astype(np.int16)
    want = port.msc_backend(frags, bitRate, uepFlag, protLevel)
    b = eng.backend(startAddr, sc.length, bitRate, uepFlag, protLevel)
    got = np.concatenate([b.process(frags[a:z]) for a, z in ((0, 1), (1, 7), (7, 16), (16, 17), (17, 40), (40, 45))])
    assert got.shape == want.shape == (ncif - 16, 24 * bitRate)
    assert np.array_equal(got, want)
    # state hand-over (the multi-GPU halo): a second backend continues bit-exactly
    b1 = eng.backend(startAddr, sc.length, bitRate, uepFlag, protLevel)
    b1.process(frags[:20])
    hist, seen = b1.get_state()
    assert seen == 20 and np.array_equal(hist, frags[5:20])
    b2 = eng.backend(startAddr, sc.length, bitRate, uepFlag, protLevel)
    b2.set_state(hist, seen)
    assert np.array_equal(b2.process(frags[20:]), want[4:])
    for x in (b, b1, b2):
        x.close()


def test_modulated_subchannel_decodes_to_payload(eng, port):
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103), (96, 64, 0, 3)], 9)
    payloads, cif = mod.make_msc(40)
    rng = np.random.default_rng(1)
    for s, pay in zip(mod.sub, payloads):
        frags = _noisy(rng, cif[:, s.startAddr * 64:s.startAddr * 64 + s.fragmentSize], flip=0.005, erase=0.0)
        b = eng.backend(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel)
        out = b.process(frags)
        assert np.array_equal(out, pay[1:1 + out.shape[0]])      # 15 CIFs of interleaving + 16 warm-up
        b.close()


def test_errors(eng):
    pkg = engine_pkg()
    with pytest.raises(pkg.DabGpuError):
        eng.protect_decode(128, 1, 0o105, np.zeros((1, 6144), np.int16))     # no such EEP level
    with pytest.raises(pkg.DabGpuError):
        eng.protect_decode(128, 1, 0o103, np.zeros((1, 100), np.int16))      # too few soft bits
    with pytest.raises(pkg.DabGpuError):
        eng.backend(860, 96, 128, 1, 0o103)                                    # beyond the 864 CUs
    assert eng.viterbi(np.zeros((0, 4 * 774), np.int16), 768).shape == (0, 768)
