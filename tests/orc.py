"""ctypes binding of the CPU oracle (oracle/dab_oracle.h).  TEST INFRASTRUCTURE ONLY.

Two builds export the same API: ``oracle/_build/liboracle.so`` (plain-C restatement, kind "port") and
``oracle/_ref/libdabref.so`` (Tier-A/B = the reference's own classes compiled unmodified, kind "reference").
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dabMode", "L", "K", "T_null", "T_F", "T_s", "T_u", "T_g", "carrierDiff",
                                         "ficSymbols", "ficGroups", "cifsPerFrame", "blocksPerCIF")]


class FrameInfo(C.Structure):
    _fields_ = [("pos", C.c_int64), ("startIndex", C.c_int32), ("coarse", C.c_int32), ("fine", C.c_int32),
                ("phase0", C.c_int32), ("correction", C.c_int32), ("freqCorrRe", C.c_float), ("freqCorrIm", C.c_float)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def build(target="all"):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True)


class Oracle:
    def __init__(self, kind="port"):
        path = os.path.join(ORACLE_DIR, "_build", "liboracle.so") if kind == "port" else \
            os.path.join(ORACLE_DIR, "_ref", "libdabref.so")
        if not os.path.exists(path):
            build("port" if kind == "port" else "ref")
        self.lib = L = C.CDLL(path)
        L.orc_build_kind.restype = C.c_char_p
        L.orc_phi.restype = C.c_float
        L.orc_ofdm_new.restype = C.c_void_p
        L.orc_ofdm_free.argtypes = [C.c_void_p]
        L.orc_find_index.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_block0.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_token.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_phase_reference.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ofdm_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        self.kind = L.orc_build_kind().decode()
        assert self.kind == ("port" if kind == "port" else "reference")

    # ---- tables ----
    def mode_params(self, mode):
        p = Params()
        if self.lib.orc_mode_params(mode, C.byref(p)) != 0:
            raise ValueError("bad mode")
        return p

    def perm_table(self, mode):
        out = np.zeros(self.mode_params(mode).K, np.int16)
        assert self.lib.orc_perm_table(mode, _p(out, C.c_int16)) == 0
        return out

    def phi(self, mode, k):
        return float(self.lib.orc_phi(mode, k))

    def ref_table(self, mode):
        out = np.zeros(2 * self.mode_params(mode).T_u, np.float32)
        assert self.lib.orc_ref_table(mode, _p(out, C.c_float)) == 0
        return out.view(np.complex64)

    def pcode(self, n):
        out = np.zeros(32, np.int8)
        assert self.lib.orc_pcode(n, _p(out, C.c_int8)) == 0
        return out

    def uep_profile(self, bitRate, protLevel):
        L = np.zeros(4, np.int16); PI = np.zeros(4, np.int16)
        if self.lib.orc_uep_profile(bitRate, protLevel, _p(L, C.c_int16), _p(PI, C.c_int16)) != 0:
            return None
        return L, PI

    def eep_profile(self, bitRate, protLevel):
        L = np.zeros(2, np.int16); PI = np.zeros(2, np.int16)
        if self.lib.orc_eep_profile(bitRate, protLevel, _p(L, C.c_int16), _p(PI, C.c_int16)) != 0:
            return None
        return L, PI

    # ---- channel decoding ----
    def viterbi(self, frameBits, soft):
        soft = np.ascontiguousarray(soft, np.int16)
        assert soft.size == 4 * (frameBits + 6)
        out = np.zeros(frameBits, np.uint8)
        assert self.lib.orc_viterbi(frameBits, _p(soft, C.c_int16), _p(out, C.c_uint8)) == 0
        return out

    def eep_deconvolve(self, bitRate, protLevel, v):
        v = np.ascontiguousarray(v, np.int16)
        out = np.zeros(24 * bitRate, np.uint8)
        assert self.lib.orc_eep_deconvolve(bitRate, protLevel, _p(v, C.c_int16), v.size, _p(out, C.c_uint8)) == 0
        return out

    def uep_deconvolve(self, bitRate, protLevel, v):
        v = np.ascontiguousarray(v, np.int16)
        out = np.zeros(24 * bitRate, np.uint8)
        assert self.lib.orc_uep_deconvolve(bitRate, protLevel, _p(v, C.c_int16), v.size, _p(out, C.c_uint8)) == 0
        return out

    def fic_decode(self, soft2304):
        v = np.ascontiguousarray(soft2304, np.int16)
        assert v.size == 2304
        bits = np.zeros(768, np.uint8); crc = np.zeros(3, np.uint8)
        assert self.lib.orc_fic_decode(_p(v, C.c_int16), _p(bits, C.c_uint8), _p(crc, C.c_uint8)) == 0
        return bits, crc

    def check_crc(self, bits):
        b = np.ascontiguousarray(bits, np.uint8)
        return bool(self.lib.orc_check_crc(_p(b, C.c_uint8), b.size))

    def prbs(self, n):
        out = np.zeros(n, np.uint8)
        self.lib.orc_prbs(_p(out, C.c_uint8), n)
        return out

    def msc_backend(self, frags, bitRate, uepFlag, protLevel):
        frags = np.ascontiguousarray(frags, np.int16)
        ncif, fragmentSize = frags.shape
        out = np.zeros((max(ncif - 16, 0), 24 * bitRate), np.uint8)
        n = self.lib.orc_msc_backend(_p(frags, C.c_int16), ncif, fragmentSize, bitRate, uepFlag, protLevel,
                                     _p(out, C.c_uint8))
        assert n == out.shape[0], n
        return out

    def time_deinterleave(self, frags):
        frags = np.ascontiguousarray(frags, np.int16)
        out = np.zeros_like(frags)
        self.lib.orc_time_deinterleave(_p(frags, C.c_int16), frags.shape[0], frags.shape[1], _p(out, C.c_int16))
        return out

    def fic_frames(self, mode, sym):
        p = self.mode_params(mode)
        sym = np.ascontiguousarray(sym, np.int16)
        nframes = sym.shape[0]
        bits = np.zeros((nframes * p.ficGroups, 768), np.uint8)
        crc = np.zeros((nframes * p.ficGroups, 3), np.uint8)
        n = self.lib.orc_fic_frames(mode, _p(sym, C.c_int16), nframes, _p(bits, C.c_uint8), _p(crc, C.c_uint8))
        assert n == bits.shape[0]
        return bits, crc

    def fig01_scan(self, bits, crc, table=None):
        """FIG 0/1 table after the given FIC groups (updates and returns `table`, (64, 6) int32)"""
        bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, 768)
        crc = np.ascontiguousarray(crc, np.uint8).reshape(-1, 3)
        table = np.zeros((64, 6), np.int32) if table is None else table
        self.lib.orc_fig01_scan(_p(bits, C.c_uint8), _p(crc, C.c_uint8), bits.shape[0], _p(table, C.c_int32))
        return table

    # ---- the reference's OWN handler classes (oracle/ref_shim/ref_tierc.cpp; in libdabref.so only) ----
    def ref_fic_frames(self, mode, sym):
        """ficHandler::process_ficBlock (compiled unmodified, its own thread) over the FIC symbols of sym [nframes][L-1][2K]"""
        p = self.mode_params(mode)
        sym = np.ascontiguousarray(np.asarray(sym, np.int16)[:, :3, :])
        nframes = sym.shape[0]
        bits = np.zeros((nframes * p.ficGroups, 768), np.uint8)
        crc = np.zeros((nframes * p.ficGroups, 3), np.uint8)
        n = self.lib.ref_fic_frames(2 * p.K, _p(sym, C.c_int16), nframes, _p(bits, C.c_uint8), _p(crc, C.c_uint8))
        assert n == bits.shape[0], n
        return bits, crc

    def ref_fig01_scan(self, bits, crc, table=None):
        """fib_processor::process_FIB (compiled unmodified) for every CRC-clean FIB, then ficList's FIG 0/1 fields;
        column 0 (`valid`) is not a reference field and is left alone"""
        bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, 768)
        crc = np.ascontiguousarray(crc, np.uint8).reshape(-1, 3)
        table = np.zeros((64, 6), np.int32) if table is None else table
        self.lib.ref_fig01_scan(_p(bits, C.c_uint8), _p(crc, C.c_uint8), bits.shape[0], _p(table, C.c_int32))
        return table

    def ref_msc_run(self, mode, sym, startAddr, Length, bitRate, uepFlag, protLevel):
        """mscHandler::process_mscBlock symbol by symbol with the reference's dabConcurrent behind it (both compiled unmodified)"""
        p = self.mode_params(mode)
        sym = np.ascontiguousarray(sym, np.int16)
        nframes = sym.shape[0]
        cap = max(nframes * p.cifsPerFrame - 16, 0)
        out = np.zeros((cap, 24 * bitRate), np.uint8)
        n = self.lib.ref_msc_run(mode, p.L, p.K, _p(sym, C.c_int16), nframes, startAddr, Length, bitRate, uepFlag, protLevel, _p(out, C.c_uint8), cap)
        assert n == cap, n
        return out

    def ref_dabplus_run(self, bits, bitRate):
        """mp4Processor::addtoFrame (compiled unmodified) CIF by CIF -> (superframes, [(first_cif, num_aus, au_start, au_crc)])"""
        bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, 24 * bitRate)
        cap = bits.shape[0] // 5 + 2
        sf = np.zeros((cap, 110 * (bitRate // 8)), np.uint8)
        info = (SuperframeInfo * cap)()
        self.lib.ref_dabplus_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        n = self.lib.ref_dabplus_run(bits.ctypes.data, bits.shape[0], bitRate, sf.ctypes.data, C.addressof(info), cap)
        assert n <= cap
        return sf[:n], [(info[i].first_cif, info[i].num_aus, tuple(info[i].au_start), info[i].au_crc) for i in range(n)]

    def ref_serial_run(self, frags, bitRate, uepFlag, protLevel):
        """dabSerial::process (compiled unmodified) over CIF fragments [ncif][fragmentSize] -> blocks [ncif - 15][24 bitRate]"""
        frags = np.ascontiguousarray(frags, np.int16)
        ncif, fragmentSize = frags.shape
        out = np.zeros((max(ncif - 15, 0), 24 * bitRate), np.uint8)
        n = self.lib.ref_serial_run(_p(frags, C.c_int16), ncif, fragmentSize, bitRate, uepFlag, protLevel, _p(out, C.c_uint8), out.shape[0])
        assert n == out.shape[0], n
        return out

    def ref_receive(self, mode, iq, sub, threshold=3, method=1, max_frames=64):
        """the reference's whole receive chain (ofdmProcessor -> ofdmDecoder -> ficHandler / mscHandler -> dabConcurrent, all
        compiled unmodified) over raw u8 IQ with one audio sub-channel sub = (startAddr, Length, bitRate, uepFlag, protLevel)
        -> (fic_bits, fic_crc, msc_blocks, state = [coarse, fine, f2Correction, localPhase, samples taken])"""
        p = self.mode_params(mode)
        iq = np.ascontiguousarray(iq, np.uint8)
        startAddr, Length, bitRate, uepFlag, protLevel = sub
        mg, mb = max_frames * p.ficGroups, max_frames * p.cifsPerFrame
        fic = np.zeros((mg, 768), np.uint8); crc = np.zeros((mg, 3), np.uint8); msc = np.zeros((mb, 24 * bitRate), np.uint8)
        ng, nb = C.c_int(0), C.c_int(0)
        state = np.zeros(6, np.int64)
        self.lib.ref_receive.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        rc = self.lib.ref_receive(mode, iq.ctypes.data, iq.size // 2, threshold, method, startAddr, Length, bitRate, uepFlag, protLevel,
                                  fic.ctypes.data, crc.ctypes.data, mg, C.addressof(ng), msc.ctypes.data, mb, C.addressof(nb), state.ctypes.data)
        assert rc == 0, rc
        return fic[:ng.value], crc[:ng.value], msc[:nb.value], state

    def resample_i16(self, iq, rate):
        iq = np.ascontiguousarray(iq, np.int16)
        n = iq.size // 2
        out = np.zeros(2 * 2048 * (n // (rate // 1000) + 1), np.float32)
        self.lib.orc_resample_i16.restype = C.c_int64
        self.lib.orc_resample_i16.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        k = self.lib.orc_resample_i16(iq.ctypes.data, n, rate, out.ctypes.data)
        return out[:2 * k]

    # ---- DAB+ super-frame layer ----
    def firecode_check(self, x11):
        b = np.ascontiguousarray(x11, np.uint8)
        return bool(self.lib.orc_firecode_check(_p(b, C.c_uint8)))

    def rs_dec(self, r120):
        r = np.ascontiguousarray(r120, np.uint8)
        d = np.zeros(110, np.uint8)
        return int(self.lib.orc_rs_dec(_p(r, C.c_uint8), _p(d, C.c_uint8))), d

    def rs_enc(self, d110):
        d = np.ascontiguousarray(d110, np.uint8)
        r = np.zeros(120, np.uint8)
        self.lib.orc_rs_enc(_p(d, C.c_uint8), _p(r, C.c_uint8))
        return r

    def dabplus(self, bitRate):
        return _DabPlus(self, bitRate)

    def msc_slice(self, mode, sym, startAddr, Length):
        p = self.mode_params(mode)
        sym = np.ascontiguousarray(sym, np.int16)
        nframes = sym.shape[0]
        frag = np.zeros((nframes * p.cifsPerFrame, Length * 64), np.int16)
        n = self.lib.orc_msc_slice(mode, _p(sym, C.c_int16), nframes, startAddr, Length, _p(frag, C.c_int16))
        assert n == frag.shape[0]
        return frag

    # ---- OFDM ----
    def fft(self, v, inverse=False):
        a = np.ascontiguousarray(v, np.complex64).copy()
        assert self.lib.orc_fft(_p(a, C.c_float), a.size, int(inverse)) == 0
        return a

    def ofdm(self, mode, threshold=3, freqSyncMethod=1):
        return _Ofdm(self, mode, threshold, freqSyncMethod)

    def ofdm_run(self, mode, iq, max_frames, threshold=3, freqSyncMethod=1):
        """iq: uint8 (rawfile format) or float32 (interleaved re, im: what the reference's input devices deliver)"""
        p = self.mode_params(mode)
        sym = np.zeros((max_frames, p.L - 1, 2 * p.K), np.int16)
        info = (FrameInfo * max_frames)()
        if np.asarray(iq).dtype == np.float32:
            iq = np.ascontiguousarray(iq, np.float32)
            self.lib.orc_ofdm_run_cf32.argtypes = self.lib.orc_ofdm_run.argtypes
            n = self.lib.orc_ofdm_run_cf32(mode, threshold, freqSyncMethod, iq.ctypes.data, iq.size // 2, max_frames,
                                           sym.ctypes.data, C.addressof(info))
        else:
            iq = np.ascontiguousarray(iq, np.uint8)
            n = self.lib.orc_ofdm_run(mode, threshold, freqSyncMethod, iq.ctypes.data, iq.size // 2, max_frames,
                                      sym.ctypes.data, C.addressof(info))
        return sym[:n], [info[i] for i in range(n)]


class SuperframeInfo(C.Structure):
    _fields_ = [("first_cif", C.c_int64), ("corrected", C.c_int32), ("num_aus", C.c_int32), ("au_start", C.c_int32 * 7),
                ("au_crc", C.c_int32)]

    def key(self):
        return (self.first_cif, self.corrected, self.num_aus, tuple(self.au_start), self.au_crc)


class _DabPlus:
    """mp4Processor's super-frame front (stateful, like the reference object)"""

    def __init__(self, orc, bitRate):
        self.lib, self.bitRate = orc.lib, bitRate
        self.lib.orc_dabplus_new.restype = C.c_void_p
        self.lib.orc_dabplus_free.argtypes = [C.c_void_p]
        self.lib.orc_dabplus_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        self.h = self.lib.orc_dabplus_new(bitRate)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_dabplus_free(self.h)
            self.h = None

    def process(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, 24 * self.bitRate)
        cap = bits.shape[0] // 5 + 2
        sf = np.zeros((cap, 110 * (self.bitRate // 8)), np.uint8)
        info = (SuperframeInfo * cap)()
        n = self.lib.orc_dabplus_process(self.h, bits.ctypes.data, bits.shape[0], sf.ctypes.data, C.addressof(info), cap)
        assert n <= cap
        return sf[:n], [info[i].key() for i in range(n)]


class _Ofdm:
    def __init__(self, orc, mode, threshold, method):
        self.lib = orc.lib
        self.p = orc.mode_params(mode)
        self.h = self.lib.orc_ofdm_new(mode, threshold, method)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_ofdm_free(self.h)
            self.h = None

    def find_index(self, v):
        a = np.ascontiguousarray(v, np.complex64)
        assert a.size == self.p.T_u
        return int(self.lib.orc_find_index(self.h, a.ctypes.data))

    def block0(self, v, flag=True):
        a = np.ascontiguousarray(v, np.complex64)
        assert a.size == self.p.T_u
        return int(self.lib.orc_block0(self.h, a.ctypes.data, int(flag)))

    def token(self, inv):
        a = np.ascontiguousarray(inv, np.complex64)
        assert a.size == self.p.T_s
        out = np.zeros(2 * self.p.K, np.int16)
        self.lib.orc_token(self.h, a.ctypes.data, out.ctypes.data)
        return out

    def phase_reference(self):
        out = np.zeros(self.p.T_u, np.complex64)
        self.lib.orc_get_phase_reference(self.h, out.ctypes.data)
        return out
