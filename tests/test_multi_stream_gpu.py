"""dabgpu_decode_multi (BASELINE configs[3]: many independent ensemble streams in one call): every stream's output
equals the oracle's run over that stream alone (trajectory exact, soft bits +-1, FIC / MSC bits bit-exact) and equals
what a fresh single-stream handle returns for it; streams of different length, CFO, SNR and lead-in, one of them pure
noise, one too short to hold a frame, an empty one."""
import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu

SUBS = [(0, 128, 1, 0o103), (96, 128, 0, 3), (200, 64, 1, 0o202)]


def _oracle_chain(port, mode, iq, nmax, subs_objs):
    sym, info = port.ofdm_run(mode, iq, nmax)
    bits, crc = port.fic_frames(mode, sym)
    msc = [port.msc_backend(port.msc_slice(mode, sym, s.startAddr, s.length), s.bitRate, s.uepFlag, s.protLevel) for s in subs_objs]
    return sym, info, bits, crc, msc


def _streams(port, mode, n, subs):
    rng = np.random.default_rng(100 + mode)
    out, mods = [], []
    for i in range(n):
        mod = dabmod.Modulator(port, mode, subs, 2000 + i)
        nfr = int(rng.integers(10, 22)) * (1 if mode == 1 else 2)
        cd = mod.p.carrierDiff
        cfo = float(rng.integers(-6, 7) * cd + rng.integers(-cd // 2 + 20, cd // 2 - 20))
        tr = mod.generate(nfr, cfo_hz=cfo, snr_db=float(rng.choice([10.0, 15.0, 25.0])), lead=int(rng.integers(500, 150000)), tail=6000)
        out.append(tr["iq"]); mods.append(mod)
    return out, mods


@pytest.mark.parametrize("mode", [1, 2, 4])
def test_multi_equals_oracle_and_single(port, mode):
    pkg = engine_pkg()
    subs = SUBS if mode != 2 else SUBS[:2]
    n = 7
    iqs, mods = _streams(port, mode, n, subs)
    rng = np.random.default_rng(5)
    iqs.append(np.clip(np.rint(rng.standard_normal(2 * 500000) * 20 + 128), 0, 255).astype(np.uint8))   # noise: never syncs
    iqs.append(iqs[0][:2 * 30000].copy())                                                                  # too short for a frame
    iqs.append(np.zeros(0, np.uint8))                                                                      # empty
    sub_objs = mods[0].sub
    subl = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in sub_objs]
    eng = pkg.DabGpu(mode=mode)
    eng.set_subchannels(subl)
    cap = 50
    res = eng.decode_multi(iqs, [eng.alloc_result(cap) for _ in iqs])
    g = mods[0].p.ficGroups
    for i, (iq, r) in enumerate(zip(iqs, res)):
        sym, info, bits, crc, msc = _oracle_chain(port, mode, iq, cap, sub_objs)
        # the oracle counts a frame whose trailing null symbol is cut off by the end of the input; the engine does not
        assert len(info) - r.nframes in (0, 1), (i, len(info), r.nframes)
        nf = r.nframes
        for k, (a, b) in enumerate(zip(r.info, info[:nf])):
            assert (a.pos, a.startIndex, a.coarse, a.fine, a.phase0, a.correction) == \
                   (b.pos, b.startIndex, b.coarse, b.fine, b.phase0, b.correction), (i, k)
        if nf:
            d = np.abs(r.soft.astype(int) - sym[:nf].astype(int))
            assert d.max() <= 1, (i, d.max())
        # Channel decoding is bit-exact on identical soft bits: the oracle's decoders fed with the ENGINE's soft bits give the
        # engine's output.  Against the oracle's own soft bits (+-1 on ~1e-5 of them: the FFT is not the reference's FFTW) a
        # block that sits on an error event of a noisy stream may come out differently: rare, and only there.
        if nf:
            b2, c2 = port.fic_frames(mode, r.soft)
            assert np.array_equal(r.fic_bits, b2) and np.array_equal(r.fic_crc, c2), i
        assert (r.fic_bits != bits[:nf * g]).any(axis=1).mean() <= 0.02 if nf else True, i
        for got, w, sc in zip(r.msc, msc, sub_objs):
            assert got.shape[0] == max(0, nf * mods[0].p.cifsPerFrame - 16), (i, got.shape)
            if got.shape[0]:
                w2 = port.msc_backend(port.msc_slice(mode, r.soft, sc.startAddr, sc.length), sc.bitRate, sc.uepFlag, sc.protLevel)
                assert np.array_equal(got, w2[:got.shape[0]]), i
                assert (got != w[:got.shape[0]]).any(axis=1).mean() <= 0.04, i
        # a fresh single-stream handle gives the same
        e1 = pkg.DabGpu(mode=mode)
        e1.set_subchannels(subl)
        one = e1.decode(iq, e1.alloc_result(cap))
        assert one.nframes == nf and np.array_equal(one.fic_bits, r.fic_bits) and one.consumed == r.consumed, (i, one.nframes, nf, one.consumed, r.consumed)
        if nf:
            assert np.array_equal(one.soft, r.soft), i
        for a, b in zip(one.msc, r.msc):
            assert np.array_equal(a, b), i
        e1.close()
    assert res[n].nframes == 0 and res[n + 1].nframes == 0 and res[n + 2].nframes == 0
    assert sum(r.nframes for r in res) > 60
    # the handle's own stream is untouched by the batch call
    st = eng.state_get()
    assert (st.synced, st.frames, st.abs_pos) == (0, 0, 0)
    eng.close()


def test_multi_float_and_device_input(port):
    """the same streams as complex floats, and u8 streams already resident on the device"""
    import torch
    pkg = engine_pkg()
    iqs, mods = _streams(port, 1, 3, SUBS[:1])
    subl = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mods[0].sub]
    eng = pkg.DabGpu(mode=1)
    eng.set_subchannels(subl)
    a = eng.decode_multi(iqs, [eng.alloc_result(40) for _ in iqs])
    f32 = [((x.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)).astype(np.float32) for x in iqs]
    b = eng.decode_multi(f32, [eng.alloc_result(40) for _ in iqs])
    d_in = [torch.from_numpy(x.copy()).cuda() for x in iqs]
    torch.cuda.synchronize()
    c = eng.decode_multi(None, [eng.alloc_result(40) for _ in iqs], dev_ptrs=[(t.data_ptr(), t.numel() // 2) for t in d_in])
    for x, y, z in zip(a, b, c):
        assert x.nframes == y.nframes == z.nframes > 5
        assert np.array_equal(x.soft, y.soft) and np.array_equal(x.soft, z.soft)
        assert np.array_equal(x.msc[0], y.msc[0]) and np.array_equal(x.msc[0], z.msc[0])
        assert np.array_equal(x.fic_bits, z.fic_bits)
    eng.close()


def test_multi_pinned_results_and_late_start(port):
    """pinned result buffers (results are copied straight into them while later rounds run) give what pageable ones give;
    a stream whose signal starts after the first upload piece (the null search runs out of resident samples and is
    repeated when more have arrived) equals its single-handle decode"""
    import torch
    pkg = engine_pkg()
    iqs, mods = _streams(port, 1, 4, SUBS)
    mod = dabmod.Modulator(port, 1, SUBS, 3001)
    iqs.append(mod.generate(12, cfo_hz=410.0, snr_db=15.0, lead=1500000, tail=6000)["iq"])      # 7.6 frames of noise first
    subl = [(s.startAddr, s.length, s.bitRate, s.uepFlag, s.protLevel) for s in mods[0].sub]
    eng = pkg.DabGpu(mode=1)
    eng.set_subchannels(subl)

    def pinned(shape, dtype):
        assert dtype == np.uint8
        return torch.empty(shape, dtype=torch.uint8).pin_memory().numpy()
    a = eng.decode_multi(iqs, [eng.alloc_result(40, want_soft=False) for _ in iqs])
    b = eng.decode_multi(iqs, [eng.alloc_result(40, want_soft=False, alloc=pinned) for _ in iqs])
    for i, (x, y) in enumerate(zip(a, b)):
        assert x.nframes == y.nframes > 5, (i, x.nframes, y.nframes)
        assert np.array_equal(x.fic_bits, y.fic_bits) and np.array_equal(x.fic_crc, y.fic_crc), i
        assert [(f.pos, f.fine) for f in x.info] == [(f.pos, f.fine) for f in y.info], i
        for s, t in zip(x.msc, y.msc):
            assert s.shape == t.shape and np.array_equal(s, t), i
    e1 = pkg.DabGpu(mode=1)
    e1.set_subchannels(subl)
    one = e1.decode(iqs[-1], e1.alloc_result(40, want_soft=False))
    assert one.nframes == b[-1].nframes and np.array_equal(one.fic_bits, b[-1].fic_bits) and one.consumed == b[-1].consumed
    for s, t in zip(one.msc, b[-1].msc):
        assert np.array_equal(s, t)
    e1.close()
    eng.close()
