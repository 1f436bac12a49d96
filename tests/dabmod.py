"""Test-side DAB modulator: the inverse of the reference receive chain, used only to
manufacture deterministic synthetic IQ and soft-bit workloads for the tests and bench.py.

payload -> energy dispersal -> K=7 rate-1/4 conv. code (+6 tail bits) -> puncturing -> time interleaving ->
CIF / FIC multiplex -> QPSK -> frequency interleaving -> differential modulation -> IFFT + cyclic prefix + null
symbol -> CFO / AWGN / gain -> u8 "rawfile" samples ((x-128)/128 on the receive side, rawfiles.cpp:113-116).

The code tables it needs (puncturing vectors, UEP/EEP profiles, PRS phases, frequency interleaver) are taken
from the oracle, which is pinned against the compiled reference.
"""
import numpy as np

POLYS = (0o155, 0o117, 0o123, 0o155)          # viterbi.cpp:63, in the sr = (sr << 1) | bit convention
TX_DELAY = 15 - np.array([15, 7, 11, 3, 13, 5, 9, 1, 14, 6, 10, 2, 12, 4, 8, 0])   # dab-concurrent.cpp:41-43
INPUT_RATE = 2048000


def prbs(n):
    """x^9 + x^5 + 1, all-ones start (fic-handler.cpp:100-108)."""
    sr = [1] * 9
    out = np.zeros(n, np.uint8)
    for i in range(n):
        b = sr[8] ^ sr[4]
        sr = [b] + sr[:8]
        out[i] = b
    return out


def conv_encode(bits):
    """bits[..., n] -> mother code word [..., 4 * (n + 6)] (6 zero tail bits appended)."""
    bits = np.asarray(bits, np.uint8)
    n = bits.shape[-1]
    pad = np.zeros(bits.shape[:-1] + (n + 12,), np.uint8)
    pad[..., 6:6 + n] = bits                       # 6 zeros of history, data, 6 tail zeros
    out = np.zeros(bits.shape[:-1] + (n + 6, 4), np.uint8)
    for k, poly in enumerate(POLYS):
        acc = np.zeros(bits.shape[:-1] + (n + 6,), np.uint8)
        for d in range(7):                         # sr bit d at step t is input bit t - d
            if (poly >> d) & 1:
                acc ^= pad[..., 6 - d:6 - d + n + 6]
        out[..., k] = acc
    return out.reshape(bits.shape[:-1] + (4 * (n + 6),))


def crc16(bits):
    """CRC-CCITT x^16+x^12+x^5+1, all-ones start, complemented (dab-constants.h:310-340) -> 16 bits."""
    reg = 0xFFFF
    for b in bits:
        top = (reg >> 15) & 1
        reg = (reg << 1) & 0xFFFF
        if top ^ int(b):
            reg ^= 0x1021
    reg ^= 0xFFFF
    return np.array([(reg >> (15 - i)) & 1 for i in range(16)], np.uint8)


class SubChannel:
    """One MSC sub-channel.  uepFlag follows the reference's inverted naming: 0 = UEP table, 1 = EEP."""

    def __init__(self, orc, startAddr, bitRate, uepFlag, protLevel):
        self.startAddr, self.bitRate, self.uepFlag, self.protLevel = startAddr, bitRate, uepFlag, protLevel
        self.nbits = 24 * bitRate
        self.mask = puncture_mask(orc, bitRate, uepFlag, protLevel)
        self.npunct = int(self.mask.sum())
        self.length = -(-self.npunct // 64)        # size in capacity units
        self.fragmentSize = self.length * 64


def puncture_mask(orc, bitRate, uepFlag, protLevel):
    prof = orc.uep_profile(bitRate, protLevel) if uepFlag == 0 else orc.eep_profile(bitRate, protLevel)
    if prof is None:
        raise ValueError("unknown protection profile")
    L, PI = prof
    n = 24 * bitRate
    mask = np.zeros(4 * n + 24, bool)
    pos = 0
    for l, pi in zip(L, PI):
        if l <= 0:
            continue
        v = orc.pcode(int(pi)).astype(bool)
        seg = np.tile(v, 4 * int(l))
        mask[pos:pos + seg.size] = seg
        pos += seg.size
    mask[pos:pos + 24] = np.tile([True, True, False, False], 6)      # PI_X, at the running position
    return mask


def fic_mask(orc):
    m = np.concatenate([np.tile(orc.pcode(16).astype(bool), 4 * 21), np.tile(orc.pcode(15).astype(bool), 4 * 3),
                        np.tile([True, True, False, False], 6)])
    assert m.sum() == 2304
    return m


class Modulator:
    def __init__(self, orc, mode, subchannels, seed):
        self.orc, self.mode = orc, mode
        self.p = orc.mode_params(mode)
        self.sub = [SubChannel(orc, *s) for s in subchannels]
        self.rng = np.random.default_rng(seed)
        self.perm = orc.perm_table(mode).astype(np.int64)
        self.prs = orc.ref_table(mode).astype(np.complex128)         # PRS spectrum, FFT bin order
        self.fmask = fic_mask(orc)

    # ---- bit level ----
    def make_fic(self, nframes):
        """-> fibs[nframes*groups, 768] (ground truth after dispersal removal), punctured[nframes*groups, 2304]"""
        g = nframes * self.p.ficGroups
        fibs = np.zeros((g, 768), np.uint8)
        for i in range(g):
            for j in range(3):
                if getattr(self, "wellformed_fibs", False):
                    # FIG 0/1 with a few random (valid) sub-channel entries, then the end marker: what a parser of real FIBs can
                    # digest (the reference's fib_processor loops forever / reads wild on random bits with a good CRC)
                    import figutil
                    ent = [("short", int(self.rng.integers(0, 64)), int(self.rng.integers(0, 864)), int(self.rng.integers(0, 64)))
                           if self.rng.integers(0, 2) else
                           ("long", int(self.rng.integers(0, 64)), int(self.rng.integers(0, 864)), int(self.rng.integers(0, 2)), int(self.rng.integers(1, 5)), int(self.rng.integers(6, 400)))
                           for _ in range(int(self.rng.integers(0, 5)))]
                    fibs[i, 256 * j:256 * j + 256] = figutil.fib([figutil.fig01(ent)] if ent else [])
                    continue
                body = self.rng.integers(0, 2, 240, dtype=np.uint8)
                fibs[i, 256 * j:256 * j + 240] = body
                fibs[i, 256 * j + 240:256 * j + 256] = crc16(body)
        scr = fibs ^ prbs(768)
        return fibs, conv_encode(scr)[:, self.fmask]

    def make_msc(self, ncif, history=None):
        """-> payload[sub][ncif, 24*bitRate], cif[ncif, 55296] code bits (time interleaved, zero history)."""
        cif = self.rng.integers(0, 2, (ncif, 55296), dtype=np.uint8)     # unused capacity: random QPSK
        payloads = []
        for s in self.sub:
            pay = self.rng.integers(0, 2, (ncif, s.nbits), dtype=np.uint8)
            payloads.append(pay)
            code = conv_encode(pay ^ prbs(s.nbits))[:, s.mask]
            a = np.zeros((ncif, s.fragmentSize), np.uint8)
            a[:, :s.npunct] = code
            tx = np.zeros_like(a)
            for d in range(16):
                cols = np.arange(s.fragmentSize) % 16 == d
                delay = int(TX_DELAY[d])
                if delay < ncif:
                    tx[delay:, cols] = a[:ncif - delay][:, cols]
            cif[:, s.startAddr * 64:s.startAddr * 64 + s.fragmentSize] = tx
        return payloads, cif

    def frame_bits(self, nframes):
        """-> dict with ground truth and symbol bits[nframes, L-1, 2K]."""
        p = self.p
        fibs, ficp = self.make_fic(nframes)
        ncif = nframes * p.cifsPerFrame
        payloads, cif = self.make_msc(ncif)
        bits = np.zeros((nframes, p.L - 1, 2 * p.K), np.uint8)
        ficflat = ficp.reshape(nframes, -1)
        nfic = 3 * 2 * p.K
        used = ficflat.shape[1]
        ficsym = self.rng.integers(0, 2, (nframes, nfic), dtype=np.uint8)
        ficsym[:, :used] = ficflat
        bits[:, :3, :] = ficsym.reshape(nframes, 3, 2 * p.K)
        msc = cif.reshape(nframes, p.cifsPerFrame * 55296)
        nmsc = (p.L - 4) * 2 * p.K
        bits[:, 3:, :] = msc[:, :nmsc].reshape(nframes, p.L - 4, 2 * p.K)
        return dict(fibs=fibs, payloads=payloads, bits=bits)

    # ---- waveform ----
    def modulate(self, bits):
        """bits[nframes, L-1, 2K] -> complex64 baseband [nframes * T_F], unit power in the symbols."""
        import scipy.fft
        p = self.p
        nframes = bits.shape[0]
        K, T_u, T_g = p.K, p.T_u, p.T_g
        # QPSK symbol ((1-2b0) + j(1-2b1))/sqrt2 = exp(j(pi/4 + k pi/2)); differential modulation = running
        # sum of the quadrant numbers k plus pi/4 per symbol, i.e. a running sum of odd multiples of pi/4
        b0, b1 = bits[..., :K].astype(np.int16), bits[..., K:].astype(np.int16)
        eighth = 1 + 2 * b0 + 6 * b1 - 4 * (b0 & b1)          # (b0,b1) -> 1,3,7,5 in units of pi/4
        idx = np.where(self.perm < 0, self.perm + T_u, self.perm)
        prs8 = np.rint(np.angle(self.prs[idx]) / (np.pi / 4)).astype(np.int16)
        acc = (np.cumsum(eighth, axis=1, dtype=np.int16) + prs8[None, None, :]) & 7
        table = np.exp(1j * np.pi / 4 * np.arange(8)).astype(np.complex64)
        spec = np.zeros((nframes, p.L, T_u), np.complex64)
        spec[:, 0, :] = self.prs.astype(np.complex64)
        spec[:, 1:, idx] = table[acc]
        t = scipy.fft.ifft(spec, axis=-1, workers=-1) * np.float32(T_u / np.sqrt(K))     # unit power per sample
        frames = np.zeros((nframes, p.T_F), np.complex64)
        sym = frames[:, p.T_null:].reshape(nframes, p.L, p.T_s)
        sym[..., T_g:] = t
        sym[..., :T_g] = t[..., T_u - T_g:]                                               # cyclic prefix
        return frames.reshape(-1)

    def channel(self, x, cfo_hz=0.0, snr_db=30.0, rms=30.0, lead=0, tail=0, phase0=0.0, start_index=0):
        """CFO, AWGN, gain, u8 quantisation.  `lead`/`tail` noise-only samples around the signal.
        `start_index` = absolute index of the first output sample (keeps the CFO phase continuous when a
        long stream is produced in pieces)."""
        n = lead + x.size + tail
        sig = np.zeros(n, np.complex64)
        sig[lead:lead + x.size] = x
        if cfo_hz != 0.0 or phase0 != 0.0:
            blk = 1 << 16                                    # e^{j w k} = e^{j w (k0 + r)}: two small tables
            w = 2.0 * np.pi * cfo_hz / INPUT_RATE
            fine = np.exp(1j * w * np.arange(blk)).astype(np.complex64)
            nb = -(-n // blk)
            coarse = np.exp(1j * (w * (start_index + blk * np.arange(nb, dtype=np.float64)) + phase0)).astype(np.complex64)
            pad = np.zeros(nb * blk, np.complex64)
            pad[:n] = sig
            pad = pad.reshape(nb, blk)
            pad *= fine[None, :]
            pad *= coarse[:, None]
            sig = pad.reshape(-1)[:n]
        sigma = np.float32(10.0 ** (-snr_db / 20.0) / np.sqrt(2.0))
        r = np.empty(2 * n, np.float32)
        r[0::2] = sig.real; r[1::2] = sig.imag
        r += self.rng.standard_normal(2 * n, dtype=np.float32) * sigma
        r *= np.float32(rms)                                 # unit-power complex signal -> |x| rms = `rms` LSB
        r += np.float32(128.0)
        np.rint(r, out=r)
        np.clip(r, 0, 255, out=r)
        return r.astype(np.uint8)

    def generate(self, nframes, cfo_hz=0.0, snr_db=30.0, rms=30.0, lead=0, tail=0):
        truth = self.frame_bits(nframes)
        iq = self.channel(self.modulate(truth["bits"]), cfo_hz, snr_db, rms, lead, tail)
        truth["iq"] = iq
        return truth


def soft_from_bits(bits, rng=None, amp=100, flip=0.0, erase=0.0, jitter=20):
    """code bits -> int16 soft bits in the reference convention (+ => 1), with optional noise."""
    bits = np.asarray(bits)
    s = np.where(bits > 0, amp, -amp).astype(np.int32)
    if rng is not None:
        s = s + rng.integers(-jitter, jitter + 1, s.shape)
        if flip > 0:
            s = np.where(rng.random(s.shape) < flip, -s, s)
        if erase > 0:
            s = np.where(rng.random(s.shape) < erase, 0, s)
    return np.clip(s, -127, 127).astype(np.int16)
