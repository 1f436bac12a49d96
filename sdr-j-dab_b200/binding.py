"""ctypes binding of libdabgpu.so (include/dabgpu.h).  Mirrors the C entry points one to one."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdabgpu.so")
_lib = None


class DabGpuError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("dabMode", C.c_int32), ("threshold", C.c_int32),
                ("freqSyncMethod", C.c_int32), ("viterbi_path", C.c_int32), ("host_batch_frames", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class SubCh(C.Structure):
    _fields_ = [("startAddr", C.c_int32), ("length", C.c_int32), ("bitRate", C.c_int32),
                ("uepFlag", C.c_int32), ("protLevel", C.c_int32)]


class FrameInfo(C.Structure):
    _fields_ = [("pos", C.c_int64), ("startIndex", C.c_int32), ("coarse", C.c_int32), ("fine", C.c_int32),
                ("phase0", C.c_int32), ("correction", C.c_int32), ("freqCorrRe", C.c_float), ("freqCorrIm", C.c_float)]


class Result(C.Structure):
    _fields_ = [("max_frames", C.c_int32), ("nframes", C.c_int32), ("info", C.POINTER(FrameInfo)),
                ("soft", C.POINTER(C.c_int16)), ("fic_bits", C.POINTER(C.c_uint8)), ("fic_crc", C.POINTER(C.c_uint8)),
                ("msc_bits", C.POINTER(C.POINTER(C.c_uint8))), ("msc_nblocks", C.POINTER(C.c_int32)),
                ("consumed", C.c_int64)]


class StreamJob(C.Structure):
    _fields_ = [("iq", C.c_void_p), ("nsamples", C.c_size_t), ("out", C.POINTER(Result))]


class StreamState(C.Structure):
    _fields_ = [("synced", C.c_int32), ("coarse", C.c_int32), ("fine", C.c_int32), ("f2Correction", C.c_int32),
                ("previous_1", C.c_int32), ("previous_2", C.c_int32), ("localPhase", C.c_int32),
                ("abs_pos", C.c_int64), ("frames", C.c_int64), ("cifs", C.c_int64)]


def load_library():
    """Loads the in-tree libdabgpu.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DabGpuError("libdabgpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.dabgpu_last_error.restype = C.c_char_p
    L.dabgpu_last_error.argtypes = [C.c_void_p]
    L.dabgpu_launch_count.restype = C.c_int64
    L.dabgpu_launch_count.argtypes = [C.c_void_p]
    L.dabgpu_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.dabgpu_destroy.argtypes = [C.c_void_p]
    L.dabgpu_destroy.restype = None
    L.dabgpu_sync.argtypes = [C.c_void_p]
    L.dabgpu_timer_begin.argtypes = [C.c_void_p]
    L.dabgpu_timer_end.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    for name in ("dabgpu_viterbi", "dabgpu_viterbi_dev"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.dabgpu_protect_decode.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                        C.c_int32, C.c_void_p]
    L.dabgpu_fic_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.dabgpu_backend_create.argtypes = [C.c_void_p, C.POINTER(SubCh), C.POINTER(C.c_void_p)]
    L.dabgpu_backend_destroy.argtypes = [C.c_void_p]
    L.dabgpu_backend_destroy.restype = None
    L.dabgpu_backend_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.dabgpu_backend_get_state.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.dabgpu_backend_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.dabgpu_profile_enable.argtypes = [C.c_void_p, C.c_int32]
    L.dabgpu_profile_reset.argtypes = [C.c_void_p]
    L.dabgpu_profile_get.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.dabgpu_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.dabgpu_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    L.dabgpu_find_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.dabgpu_block0.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int16)]
    L.dabgpu_token.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.dabgpu_get_phase_reference.argtypes = [C.c_void_p, C.c_void_p]
    L.dabgpu_set_subchannels.argtypes = [C.c_void_p, C.POINTER(SubCh), C.c_int32]
    L.dabgpu_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Result)]
    L.dabgpu_decode_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Result)]
    L.dabgpu_reset.argtypes = [C.c_void_p]
    L.dabgpu_state_get.argtypes = [C.c_void_p, C.POINTER(StreamState)]
    L.dabgpu_state_set.argtypes = [C.c_void_p, C.POINTER(StreamState)]
    L.dabgpu_fig01_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.dabgpu_get_subch_table.argtypes = [C.c_void_p, C.c_void_p]
    L.dabgpu_host_state_predict.argtypes = [C.c_int32, C.POINTER(StreamState), C.c_int64, C.POINTER(StreamState)]
    L.dabgpu_state_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.dabgpu_state_import.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    _lib = L
    return L


MODE_PARAMS = {  # L, K, T_null, T_F, T_s, T_u, cifsPerFrame (gui.cpp:1328-1372, msc-handler.cpp:61-71)
    1: (76, 1536, 2656, 196608, 2552, 2048, 4), 2: (76, 384, 664, 49152, 638, 512, 1),
    3: (153, 192, 345, 49152, 319, 256, 0), 4: (76, 768, 1328, 98304, 1276, 1024, 2)}


def state_predict(mode, s, nframes):
    out = StreamState()
    rc = load_library().dabgpu_host_state_predict(mode, C.byref(s), int(nframes), C.byref(out))
    if rc != 0:
        raise DabGpuError("dabgpu_host_state_predict failed (%d): not a locked state?" % rc)
    return out


class SuperframeInfo(C.Structure):
    _fields_ = [("first_cif", C.c_int64), ("corrected", C.c_int32), ("num_aus", C.c_int32), ("au_start", C.c_int32 * 7),
                ("au_crc", C.c_int32)]

    def key(self):
        return (self.first_cif, self.corrected, self.num_aus, tuple(self.au_start), self.au_crc)


class DabPlus:
    """DAB+ super-frame layer object (dabgpu_dabplus_*): Fire-code sync, RS(120,110) repair, AU table"""

    def __init__(self, engine, bitRate):
        self.eng, self.bitRate = engine, bitRate
        self.lib = engine.lib
        self.lib.dabgpu_dabplus_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        self.lib.dabgpu_dabplus_destroy.argtypes = [C.c_void_p]
        self.lib.dabgpu_dabplus_destroy.restype = None
        for f in (self.lib.dabgpu_dabplus_process, self.lib.dabgpu_dabplus_process_dev):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        self.d = C.c_void_p()
        engine._check(self.lib.dabgpu_dabplus_create(engine.h, bitRate, C.byref(self.d)))

    def close(self):
        if getattr(self, "d", None):
            self.lib.dabgpu_dabplus_destroy(self.d)
            self.d = None

    __del__ = close

    def process(self, bits, dev_ptr=None, ncif=None):
        """bits[ncif][24*bitRate] (host numpy) or a device pointer -> (superframes[n][110*R], [info keys])"""
        if dev_ptr is None:
            bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, 24 * self.bitRate)
            ncif = bits.shape[0]
        cap = ncif // 5 + 2
        sf = np.zeros((cap, 110 * (self.bitRate // 8)), np.uint8)
        info = (SuperframeInfo * cap)()
        n = C.c_int32(0)
        if dev_ptr is None:
            rc = self.lib.dabgpu_dabplus_process(self.d, bits.ctypes.data, ncif, sf.ctypes.data, C.addressof(info), cap, C.byref(n))
        else:
            rc = self.lib.dabgpu_dabplus_process_dev(self.d, dev_ptr, ncif, sf.ctypes.data, C.addressof(info), cap, C.byref(n))
        self.eng._check(rc)
        assert n.value <= cap
        return sf[:n.value], [info[i].key() for i in range(n.value)]


class WavInfo(C.Structure):
    _fields_ = [("format_tag", C.c_int32), ("channels", C.c_int32), ("samplerate", C.c_int32), ("bits", C.c_int32),
                ("sample_format", C.c_int32), ("pad", C.c_int32), ("data_offset", C.c_int64), ("nsamples", C.c_int64),
                ("nsamples_total", C.c_int64)]


def wav_parse(lib, image):
    """dabgpu_host_wav_parse on a bytes-like file image -> WavInfo, or None when the engine does not take the file"""
    buf = np.frombuffer(image, np.uint8)
    info = WavInfo()
    lib.dabgpu_host_wav_parse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(WavInfo)]
    rc = lib.dabgpu_host_wav_parse(buf.ctypes.data if buf.size else None, buf.size, C.byref(info))
    return info if rc == 0 else None


class _InfoView:
    """the first n per-frame records of a result buffer, read on demand (building a list of a thousand ctypes
    structs per call costs more host time than the call itself)"""

    def __init__(self, arr, n, valid=True):
        self.arr, self.n = arr, (n if valid else 0)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.arr[k] for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return self.arr[i]

    def __iter__(self):
        return (self.arr[k] for k in range(self.n))


class DecodeOut:
    """Host-side result buffers of one dabgpu_decode call."""
    pass


class DabGpu:
    """One engine handle (dabgpu_t)."""

    def __init__(self, mode=1, device=0, threshold=3, freqSyncMethod=1, viterbi_path=0, host_batch_frames=0, generic_symbol_kernel=False, dev_batch_frames=0):
        self.lib = load_library()
        cfg = Config(device=device, dabMode=mode, threshold=threshold, freqSyncMethod=freqSyncMethod, viterbi_path=viterbi_path)
        cfg.host_batch_frames = host_batch_frames
        cfg.reserved[0] = 1 if generic_symbol_kernel else 0
        cfg.reserved[1] = dev_batch_frames
        self.h = C.c_void_p()
        rc = self.lib.dabgpu_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise DabGpuError("dabgpu_create failed (%d): %s" % (rc, self.lib.dabgpu_last_error(None).decode()))
        self.mode = mode

    def close(self):
        if getattr(self, "h", None):
            self.lib.dabgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise DabGpuError("dabgpu error %d: %s" % (rc, self.lib.dabgpu_last_error(self.h).decode()))

    def sync(self):
        self._check(self.lib.dabgpu_sync(self.h))

    def timer_begin(self):
        self._check(self.lib.dabgpu_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._check(self.lib.dabgpu_timer_end(self.h, C.byref(ms)))
        return ms.value

    KERNEL_CLASSES = ("acquire", "front", "symbol", "scan", "viterbi_msc", "viterbi_fic", "viterbi_api", "crc", "viterbi_tb", "viterbi_sym")

    def profile_enable(self, on=True):
        self._check(self.lib.dabgpu_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.lib.dabgpu_profile_reset(self.h))

    def profile(self):
        """{class: (launches, device ms)} since the last reset"""
        out = {}
        for i, name in enumerate(self.KERNEL_CLASSES):
            n, ms = C.c_int64(0), C.c_double(0)
            self._check(self.lib.dabgpu_profile_get(self.h, i, C.byref(n), C.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    def int_peak(self):
        ops = (C.c_double * 6)()
        self._check(self.lib.dabgpu_int_peak(self.h, ops))
        return {"add": ops[0], "min": ops[1], "add_mad": ops[2], "vadd2": ops[3], "vmin2": ops[4], "vibmin2_or2": ops[5]}

    def launch_count(self):
        return int(self.lib.dabgpu_launch_count(self.h))

    # ---- Viterbi group ----
    def viterbi(self, soft, frameBits):
        soft = np.ascontiguousarray(soft, np.int16).reshape(-1, 4 * (frameBits + 6))
        out = np.empty((soft.shape[0], frameBits), np.uint8)
        self._check(self.lib.dabgpu_viterbi(self.h, soft.ctypes.data, frameBits, soft.shape[0], out.ctypes.data))
        return out

    def viterbi_dev(self, d_soft_ptr, frameBits, nblocks, d_bits_ptr):
        self._check(self.lib.dabgpu_viterbi_dev(self.h, d_soft_ptr, frameBits, nblocks, d_bits_ptr))

    def protect_decode(self, bitRate, uepFlag, protLevel, v):
        v = np.ascontiguousarray(v, np.int16)
        if v.ndim == 1:
            v = v[None, :]
        out = np.empty((v.shape[0], 24 * bitRate), np.uint8)
        self._check(self.lib.dabgpu_protect_decode(self.h, bitRate, uepFlag, protLevel, v.ctypes.data, v.shape[1],
                                                   v.shape[0], out.ctypes.data))
        return out

    def fic_decode(self, soft):
        soft = np.ascontiguousarray(soft, np.int16).reshape(-1, 2304)
        bits = np.empty((soft.shape[0], 768), np.uint8)
        crc = np.empty((soft.shape[0], 3), np.uint8)
        self._check(self.lib.dabgpu_fic_decode(self.h, soft.ctypes.data, soft.shape[0], bits.ctypes.data,
                                               crc.ctypes.data))
        return bits, crc

    def resample_i16(self, iq, rate):
        """airspy-style rate conversion: int16 I,Q at `rate` -> (float32 interleaved re,im at 2.048 MS/s, consumed input samples)"""
        iq = np.ascontiguousarray(iq, np.int16)
        n = iq.size // 2
        out = np.zeros(2 * 2048 * (max(n - 1, 0) // (rate // 1000)) + 2, np.float32)
        n_out, consumed = C.c_size_t(0), C.c_size_t(0)
        self.lib.dabgpu_resample_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        self._check(self.lib.dabgpu_resample_i16(self.h, iq.ctypes.data, n, rate, out.ctypes.data, C.byref(n_out), C.byref(consumed)))
        return out[:2 * n_out.value], consumed.value

    def fig01_scan(self, fic_bits, crc_ok):
        """FIG 0/1 sub-channel table of the given FIC groups -> (64, 6) int32: defined, startAddr, length, uepFlag, protLevel, bitRate"""
        bits = np.ascontiguousarray(fic_bits, np.uint8).reshape(-1, 768)
        crc = np.ascontiguousarray(crc_ok, np.uint8).reshape(-1, 3)
        assert bits.shape[0] == crc.shape[0]
        table = np.zeros((64, 6), np.int32)
        self._check(self.lib.dabgpu_fig01_scan(self.h, bits.ctypes.data, crc.ctypes.data, bits.shape[0], table.ctypes.data))
        return table

    def subch_table(self):
        """the stream engine's running FIG 0/1 table (same layout as fig01_scan)"""
        table = np.zeros((64, 6), np.int32)
        self._check(self.lib.dabgpu_get_subch_table(self.h, table.ctypes.data))
        return table

    # ---- OFDM group, per-call ----
    def fft(self, v, inverse=False):
        T_u = MODE_PARAMS[self.mode][5]
        a = np.ascontiguousarray(v, np.complex64).reshape(-1, T_u).copy()
        self._check(self.lib.dabgpu_fft(self.h, a.ctypes.data, a.shape[0], int(inverse)))
        return a

    def find_index(self, v):
        T_u = MODE_PARAMS[self.mode][5]
        a = np.ascontiguousarray(v, np.complex64).reshape(-1, T_u)
        idx = np.empty(a.shape[0], np.int32)
        self._check(self.lib.dabgpu_find_index(self.h, a.ctypes.data, a.shape[0], idx.ctypes.data))
        return idx

    def block0(self, v, flag=True):
        a = np.ascontiguousarray(v, np.complex64)
        assert a.size == MODE_PARAMS[self.mode][5]
        c = C.c_int16(0)
        self._check(self.lib.dabgpu_block0(self.h, a.ctypes.data, int(flag), C.byref(c)))
        return c.value

    def token(self, inv):
        L, K, _, _, T_s, _, _ = MODE_PARAMS[self.mode]
        a = np.ascontiguousarray(inv, np.complex64).reshape(-1, T_s)
        out = np.empty((a.shape[0], 2 * K), np.int16)
        self._check(self.lib.dabgpu_token(self.h, a.ctypes.data, a.shape[0], out.ctypes.data))
        return out

    def phase_reference(self):
        out = np.empty(MODE_PARAMS[self.mode][5], np.complex64)
        self._check(self.lib.dabgpu_get_phase_reference(self.h, out.ctypes.data))
        return out

    # ---- stream decode ----
    def set_subchannels(self, subs):
        """subs: list of (startAddr, length, bitRate, uepFlag, protLevel)"""
        self._subs = [SubCh(*s) for s in subs]
        arr = (SubCh * max(len(subs), 1))(*self._subs)
        self._check(self.lib.dabgpu_set_subchannels(self.h, arr, len(subs)))

    def alloc_result(self, max_frames, want_soft=True, alloc=None):
        """alloc(shape, dtype) -> ndarray lets the caller place the result buffers (e.g. in pinned memory)"""
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        L, K, _, _, _, _, cpf = MODE_PARAMS[self.mode]
        subs = getattr(self, "_subs", [])
        o = DecodeOut()
        o.max_frames = max_frames
        o.info = (FrameInfo * max(max_frames, 1))()
        o.soft = alloc((max_frames, L - 1, 2 * K), np.int16) if want_soft else None
        groups = 3 * 2 * K // 2304
        o.fic_bits = alloc((max_frames * groups, 768), np.uint8)
        o.fic_crc = alloc((max_frames * groups, 3), np.uint8)
        o.msc = [alloc((max_frames * cpf, (3 if getattr(self, "_packed", False) else 24) * s.bitRate), np.uint8) for s in subs]
        o.ptrs = (C.POINTER(C.c_uint8) * max(len(subs), 1))(*[m.ctypes.data_as(C.POINTER(C.c_uint8)) for m in o.msc])
        o.nblocks = (C.c_int32 * max(len(subs), 1))()
        o.res = Result(max_frames=max_frames, nframes=0, info=o.info,
                       soft=o.soft.ctypes.data_as(C.POINTER(C.c_int16)) if want_soft else None,
                       fic_bits=o.fic_bits.ctypes.data_as(C.POINTER(C.c_uint8)),
                       fic_crc=o.fic_crc.ctypes.data_as(C.POINTER(C.c_uint8)),
                       msc_bits=o.ptrs, msc_nblocks=o.nblocks, consumed=0)
        return o

    def decode(self, iq_u8, out):
        """iq_u8: numpy uint8 (interleaved I,Q; rawfile format) or float32 (interleaved re, im: dabgpu_decode_cf32), or an
        int host address of u8 samples with `nsamples` given via a tuple"""
        if isinstance(iq_u8, tuple):
            ptr, nsamples = iq_u8
        elif np.asarray(iq_u8).dtype == np.float32:
            iq = np.ascontiguousarray(iq_u8, np.float32)
            self.lib.dabgpu_decode_cf32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Result)]
            self._check(self.lib.dabgpu_decode_cf32(self.h, iq.ctypes.data, iq.size // 2, C.byref(out.res)))
            return self._trim(out)
        elif np.asarray(iq_u8).dtype == np.int16:                 # 16-bit .sdr / WAV payload (dabgpu_decode_i16)
            iq = np.ascontiguousarray(iq_u8, np.int16)
            self.lib.dabgpu_decode_i16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Result)]
            self._check(self.lib.dabgpu_decode_i16(self.h, iq.ctypes.data, iq.size // 2, C.byref(out.res)))
            return self._trim(out)
        else:
            iq = np.ascontiguousarray(iq_u8, np.uint8)
            ptr, nsamples = iq.ctypes.data, iq.size // 2
        self._check(self.lib.dabgpu_decode(self.h, ptr, nsamples, C.byref(out.res)))
        return self._trim(out)

    def prefetch(self, iq):
        """dabgpu_prefetch: announce the block a later decode() call will be given (numpy array kept alive by the caller, or a
        (host address, nsamples) tuple of u8 samples); its upload starts at once"""
        self.lib.dabgpu_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32]
        if isinstance(iq, tuple):
            ptr, ns, fmt = iq[0], iq[1], 0
        else:
            assert iq.flags["C_CONTIGUOUS"]
            fmt = 1 if iq.dtype == np.float32 else 2 if iq.dtype == np.int16 else 0
            ptr, ns = iq.ctypes.data, iq.size // 2
        self._check(self.lib.dabgpu_prefetch(self.h, ptr, ns, fmt))

    def decode_dev(self, d_ptr, nsamples, out):
        self._check(self.lib.dabgpu_decode_dev(self.h, d_ptr, nsamples, C.byref(out.res)))
        return self._trim(out)

    def set_msc_output(self, packed):
        """dabgpu_set_msc_output: MSC blocks with 8 bits per byte (first bit on top); affects result buffers allocated afterwards"""
        self.lib.dabgpu_set_msc_output.argtypes = [C.c_void_p, C.c_int32]
        self._check(self.lib.dabgpu_set_msc_output(self.h, 1 if packed else 0))
        self._packed = bool(packed)

    def decode_multi(self, streams, outs, dev_ptrs=None, host_ptrs=False):
        """dabgpu_decode_multi: `streams` = list of numpy arrays (all uint8, float32 or int16, interleaved I,Q), one per
        independent stream; outs = one alloc_result buffer per stream.  dev_ptrs: [(device pointer, nsamples)] instead of
        host arrays (dabgpu_decode_multi_dev, sample format u8).  -> list of trimmed results"""
        n = len(outs)
        jobs = (StreamJob * max(n, 1))()
        keep = []
        fmt = 0
        if dev_ptrs is None:
            dt = np.asarray(streams[0]).dtype if n and not host_ptrs else np.dtype(np.uint8)
            fmt = 1 if dt == np.float32 else 2 if dt == np.int16 else 0
            for i, x in enumerate(streams):
                if host_ptrs:                                # (host address of u8 samples, nsamples): e.g. pinned memory
                    jobs[i].iq, jobs[i].nsamples = x
                    continue
                a = np.ascontiguousarray(x, dt)
                keep.append(a)
                jobs[i].iq, jobs[i].nsamples = a.ctypes.data, a.size // 2
        else:
            for i, (ptr, ns) in enumerate(dev_ptrs):
                jobs[i].iq, jobs[i].nsamples = ptr, ns
        for i, o in enumerate(outs):
            jobs[i].out = C.pointer(o.res)
        f = self.lib.dabgpu_decode_multi if dev_ptrs is None else self.lib.dabgpu_decode_multi_dev
        f.argtypes = [C.c_void_p, C.POINTER(StreamJob), C.c_int32, C.c_int32]
        self._check(f(self.h, jobs, n, fmt))
        return [self._trim(o) for o in outs]

    def _trim(self, o):
        L, K, _, _, _, _, cpf = MODE_PARAMS[self.mode]
        n = o.res.nframes
        groups = 3 * 2 * K // 2304
        r = DecodeOut()
        r.nframes, r.consumed = n, o.res.consumed
        r.info = _InfoView(o.info, n, bool(o.res.info))
        r.soft = o.soft[:n] if o.soft is not None else None
        r.fic_bits, r.fic_crc = o.fic_bits[:n * groups], o.fic_crc[:n * groups]
        r.msc = [m[:o.nblocks[i]] for i, m in enumerate(o.msc)]
        return r

    def reset(self):
        self._check(self.lib.dabgpu_reset(self.h))

    def coarse_corrector(self, on):
        """ofdmProcessor::coarseCorrectorOn / coarseCorrectorOff"""
        self.lib.dabgpu_coarse_corrector.argtypes = [C.c_void_p, C.c_int32]
        self._check(self.lib.dabgpu_coarse_corrector(self.h, 1 if on else 0))

    # ---- geometry and result helpers used by parallel.decode_sharded
    @property
    def frame_len(self):
        return MODE_PARAMS[self.mode][3]

    @property
    def cifs_per_frame(self):
        return MODE_PARAMS[self.mode][6]

    @property
    def frame_need(self):
        """samples a frame needs from its SyncOnPhase window start on, worst case (ofdm-processor.cpp:344-442)"""
        L, _, T_null, _, T_s, T_u, _ = MODE_PARAMS[self.mode]
        return 2 * T_u + (L - 1) * T_s + T_null

    @staticmethod
    def make_state(**kw):
        return StreamState(**kw)

    def drop_frames(self, r, n):
        """the result without its first n frames (MSC blocks are not touched: a fresh engine's warm-up already swallowed them)"""
        _, K, _, _, _, _, _ = MODE_PARAMS[self.mode]
        g = 3 * 2 * K // 2304
        o = DecodeOut()
        n = min(n, r.nframes)
        o.nframes, o.consumed, o.info = r.nframes - n, r.consumed, r.info[n:]
        o.soft = r.soft[n:] if r.soft is not None else None
        o.fic_bits, o.fic_crc, o.msc = r.fic_bits[n * g:], r.fic_crc[n * g:], r.msc
        return o

    @staticmethod
    def concat_results(a, b):
        o = DecodeOut()
        o.nframes, o.consumed, o.info = a.nframes + b.nframes, a.consumed + b.consumed, list(a.info) + list(b.info)
        o.soft = np.concatenate([a.soft, b.soft]) if a.soft is not None and b.soft is not None else None
        o.fic_bits, o.fic_crc = np.concatenate([a.fic_bits, b.fic_bits]), np.concatenate([a.fic_crc, b.fic_crc])
        o.msc = [np.concatenate([x, y]) for x, y in zip(a.msc, b.msc)]
        return o

    def state_get(self):
        s = StreamState()
        self._check(self.lib.dabgpu_state_get(self.h, C.byref(s)))
        return s

    def state_set(self, s):
        self._check(self.lib.dabgpu_state_set(self.h, C.byref(s)))

    def state_predict(self, s, nframes):
        """closed-form tracking state `nframes` frames after the locked state s (dabgpu_host_state_predict)"""
        return state_predict(self.mode, s, nframes)

    def export_state(self):
        """-> uint8 ndarray: the whole stream state (sync, sample tail, de-interleaver halo)"""
        n = C.c_size_t(0)
        self._check(self.lib.dabgpu_state_export(self.h, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._check(self.lib.dabgpu_state_export(self.h, buf.ctypes.data, buf.size, C.byref(n)))
        return buf

    def import_state(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        self._check(self.lib.dabgpu_state_import(self.h, blob.ctypes.data, blob.size))

    def backend(self, startAddr, length, bitRate, uepFlag, protLevel):
        return Backend(self, SubCh(startAddr, length, bitRate, uepFlag, protLevel))


class DabGroup:
    """dabgpu_group_t: several engine handles (one per GPU, or several on one GPU for tests) behind one object"""

    def __init__(self, devices, mode=1, threshold=3, freqSyncMethod=1, viterbi_path=0):
        self.lib = load_library()
        L = self.lib
        L.dabgpu_group_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
        L.dabgpu_group_destroy.argtypes = [C.c_void_p]
        L.dabgpu_group_destroy.restype = None
        L.dabgpu_group_last_error.argtypes = [C.c_void_p]
        L.dabgpu_group_last_error.restype = C.c_char_p
        L.dabgpu_group_set_subchannels.argtypes = [C.c_void_p, C.POINTER(SubCh), C.c_int32]
        L.dabgpu_group_decode_multi.argtypes = [C.c_void_p, C.POINTER(StreamJob), C.c_int32, C.c_int32]
        L.dabgpu_group_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Result), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        cfg = Config(device=0, dabMode=mode, threshold=threshold, freqSyncMethod=freqSyncMethod, viterbi_path=viterbi_path)
        devs = (C.c_int32 * len(devices))(*devices)
        self.g = C.c_void_p()
        rc = L.dabgpu_group_create(C.byref(cfg), devs, len(devices), C.byref(self.g))
        if rc != 0:
            self.g = None
            raise DabGpuError("dabgpu_group_create failed (%d): %s" % (rc, L.dabgpu_last_error(None).decode()))
        self.mode, self.n = mode, len(devices)
        self._shape = DabGpu.__new__(DabGpu)                 # result-buffer helper (alloc_result / _trim), no handle of its own
        self._shape.mode, self._shape.h, self._shape.lib = mode, None, L

    def close(self):
        if getattr(self, "g", None):
            self.lib.dabgpu_group_destroy(self.g)
            self.g = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise DabGpuError("dabgpu group error %d: %s" % (rc, self.lib.dabgpu_group_last_error(self.g).decode()))

    def set_subchannels(self, subs):
        arr = (SubCh * max(len(subs), 1))(*[SubCh(*s) for s in subs])
        self._check(self.lib.dabgpu_group_set_subchannels(self.g, arr, len(subs)))
        self._shape._subs = [SubCh(*s) for s in subs]

    def alloc_result(self, max_frames, **kw):
        return self._shape.alloc_result(max_frames, **kw)

    def decode(self, iq_u8, out, lead_frames=24, scheme=1):
        """ONE recording over all members -> (trimmed result, scheme used: 1 parallel / 0 chain)"""
        if isinstance(iq_u8, tuple):
            ptr, ns = iq_u8
        else:
            iq = np.ascontiguousarray(iq_u8, np.uint8)
            ptr, ns = iq.ctypes.data, iq.size // 2
        used = C.c_int32(-1)
        self._check(self.lib.dabgpu_group_decode(self.g, ptr, ns, C.byref(out.res), lead_frames, scheme, C.byref(used)))
        return self._shape._trim(out), used.value

    def decode_multi(self, streams, outs):
        n = len(outs)
        jobs = (StreamJob * max(n, 1))()
        dt = np.asarray(streams[0]).dtype if n else np.dtype(np.uint8)
        fmt = 1 if dt == np.float32 else 2 if dt == np.int16 else 0
        keep = []
        for i, x in enumerate(streams):
            if isinstance(x, tuple):
                jobs[i].iq, jobs[i].nsamples = x
            else:
                a = np.ascontiguousarray(x, dt)
                keep.append(a)
                jobs[i].iq, jobs[i].nsamples = a.ctypes.data, a.size // 2
            jobs[i].out = C.pointer(outs[i].res)
        self._check(self.lib.dabgpu_group_decode_multi(self.g, jobs, n, fmt))
        return [self._shape._trim(o) for o in outs]


class Backend:
    """dabConcurrent stand-in: stateful time de-interleave + protection decode + dispersal."""

    def __init__(self, eng, sc):
        self.eng, self.sc = eng, sc
        self.b = C.c_void_p()
        eng._check(eng.lib.dabgpu_backend_create(eng.h, C.byref(sc), C.byref(self.b)))

    def close(self):
        if getattr(self, "b", None) and self.eng.h:
            self.eng.lib.dabgpu_backend_destroy(self.b)
        self.b = None

    __del__ = close

    def process(self, frags):
        frags = np.ascontiguousarray(frags, np.int16).reshape(-1, self.sc.length * 64)
        out = np.empty((frags.shape[0], 24 * self.sc.bitRate), np.uint8)
        n = C.c_int32(0)
        self.eng._check(self.eng.lib.dabgpu_backend_process(self.b, frags.ctypes.data, frags.shape[0],
                                                            out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def get_state(self):
        hist = np.empty((15, self.sc.length * 64), np.int16)
        n = C.c_int32(0)
        self.eng._check(self.eng.lib.dabgpu_backend_get_state(self.b, hist.ctypes.data, C.byref(n)))
        return hist, n.value

    def set_state(self, hist, cifs_seen):
        hist = np.ascontiguousarray(hist, np.int16)
        self.eng._check(self.eng.lib.dabgpu_backend_set_state(self.b, hist.ctypes.data, cifs_seen))
