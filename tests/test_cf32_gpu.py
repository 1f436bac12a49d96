"""GPU: the stream engine fed with complex float samples (dabgpu_decode_cf32; the form every input device of the
reference delivers, virtual-input.h:62-63 / wavfiles.cpp:168-180) against the oracle run on the same floats."""
import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu
SUBS = [(0, 128, 1, 0o103), (96, 128, 0, 3)]


def _engine(pkg, mod, mode):
    eng = pkg.DabGpu(mode=mode)
    eng.set_subchannels([(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub])
    return eng


@pytest.mark.parametrize("mode,cfo", [(1, 2210.0), (2, -1300.0), (4, 3050.0)])
def test_float_samples_equal_u8_samples(port, mode, cfo):
    """(x - 128) / 128 as floats is exactly what the u8 path computes internally: every output must be identical,
    through different kernels (Mode I: the generic symbol kernel instead of the register-FFT one: soft bits +-1)"""
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, mode, SUBS if mode != 2 else SUBS[:1], 15)
    n = 14 if mode == 1 else 30
    iq = mod.generate(n, cfo_hz=cfo, snr_db=20.0, lead=4321, tail=5000)["iq"]
    fq = ((iq.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    a = _engine(pkg, mod, mode); ra = a.decode(iq, a.alloc_result(n + 4))
    b = _engine(pkg, mod, mode); rb = b.decode(fq, b.alloc_result(n + 4))
    assert ra.nframes == rb.nframes > 0
    assert [(i.pos, i.coarse, i.fine, i.phase0, i.startIndex) for i in ra.info] == [(i.pos, i.coarse, i.fine, i.phase0, i.startIndex) for i in rb.info]
    d = np.abs(ra.soft.astype(int) - rb.soft.astype(int))
    assert d.max() <= (1 if mode == 1 else 0)
    assert np.array_equal(ra.fic_bits, rb.fic_bits)
    for x, y in zip(ra.msc, rb.msc):
        assert np.array_equal(x, y)
    a.close(); b.close()


def test_float_stream_matches_oracle_in_pieces(port):
    """floats that no u8 file could hold (gain 0.37, tiny dither), fed in ragged pieces, against the oracle's float run"""
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 4, SUBS, 16)
    iq = mod.generate(24, cfo_hz=-2400.0, snr_db=18.0, lead=999, tail=4000)["iq"]
    rng = np.random.default_rng(3)
    fq = (((iq.astype(np.float32) - 128.0) / 128.0) * np.float32(0.37) + rng.standard_normal(iq.size).astype(np.float32) * np.float32(1e-4)).astype(np.float32)
    sym, info = port.ofdm_run(4, fq, 30)
    bits, crc = port.fic_frames(4, sym)
    eng = _engine(pkg, mod, 4)
    cuts = [0, 2 * 30001, 2 * 250000, 2 * 250001, fq.size // 2 // 2 * 2, fq.size]
    parts = [eng.decode(fq[a:b], eng.alloc_result(30)) for a, b in zip(cuts[:-1], cuts[1:])]
    n = sum(p.nframes for p in parts)
    assert len(info) - n in (0, 1) and n > 15
    got_info = [i for p in parts for i in p.info]
    assert [(i.pos, i.coarse, i.fine, i.phase0) for i in got_info] == [(i.pos, i.coarse, i.fine, i.phase0) for i in info[:n]]
    soft = np.concatenate([p.soft for p in parts])
    assert np.abs(soft.astype(int) - sym[:n].astype(int)).max() <= 1
    g = mod.p.ficGroups
    assert np.array_equal(np.concatenate([p.fic_bits for p in parts]), bits[:n * g])
    for k, s in enumerate(mod.sub):
        want = port.msc_backend(port.msc_slice(4, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel)
        got = np.concatenate([p.msc[k] for p in parts])
        assert got.shape[0] > 0 and np.array_equal(got, want[:got.shape[0]])
    eng.close()


def test_format_switch_needs_an_empty_tail(port):
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 2, SUBS[:1], 17)
    iq = mod.generate(6, cfo_hz=0.0, snr_db=25.0, lead=500, tail=3000)["iq"]
    eng = _engine(pkg, mod, 2)
    eng.decode(iq[:200000], eng.alloc_result(8))                 # leaves unconsumed u8 samples behind
    with pytest.raises(pkg.DabGpuError, match="sample format"):
        eng.decode(np.zeros(1000, np.float32), eng.alloc_result(8))
    eng.close()


def test_int16_recording_equals_its_float_form(port):
    """a 16-bit .sdr / WAV payload (dabgpu_decode_i16): sf_readf_float hands the reference x / 32768 (wavfiles.cpp:186-197);
    the engine converts in the sample fetch -- same frames, same soft bits, same decoded bits as the float stream, which
    in turn is checked against the oracle's float run"""
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103), (96, 64, 0, 3)], 31)
    tr = mod.generate(22, cfo_hz=-1870.0, snr_db=18.0, lead=6001, tail=5000)
    i16 = ((tr["iq"].astype(np.int16) - 128) * 200).astype(np.int16)          # an int16 recording of the same signal
    f32 = (i16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    subs = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mod.sub]
    ea = pkg.DabGpu(mode=1); ea.set_subchannels(subs)
    eb = pkg.DabGpu(mode=1); eb.set_subchannels(subs)
    cuts = [0, 2 * 700001, 2 * 2100000, i16.size]
    a_parts = [ea.decode(i16[x:y], ea.alloc_result(30)) for x, y in zip(cuts[:-1], cuts[1:])]   # ragged pieces: the int16 tail carries over
    b = eb.decode(f32, eb.alloc_result(30))
    assert sum(p.nframes for p in a_parts) == b.nframes >= 16
    assert np.array_equal(np.concatenate([p.soft for p in a_parts]), b.soft)
    assert np.array_equal(np.concatenate([p.fic_bits for p in a_parts]), b.fic_bits)
    for k in range(len(subs)):
        assert np.array_equal(np.concatenate([p.msc[k] for p in a_parts]), b.msc[k])
    sym, info = port.ofdm_run(1, f32, 30)
    bits, crc = port.fic_frames(1, sym)
    assert np.array_equal(b.fic_bits, bits[:b.nframes * 4]) and np.array_equal(b.fic_crc, crc[:b.nframes * 4])
    assert b.fic_crc[-8:].all()
    # a state blob of an int16 stream carries its sample tail
    blob = ea.export_state()
    ec = pkg.DabGpu(mode=1); ec.set_subchannels(subs); ec.import_state(blob)
    ea.close(); eb.close(); ec.close()
