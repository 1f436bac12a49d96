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
                ("freqSyncMethod", C.c_int32), ("reserved", C.c_int32 * 4)]


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
    _lib = L
    return L


class DabGpu:
    """One engine handle (dabgpu_t)."""

    def __init__(self, mode=1, device=0, threshold=3, freqSyncMethod=1):
        self.lib = load_library()
        cfg = Config(device=device, dabMode=mode, threshold=threshold, freqSyncMethod=freqSyncMethod)
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

    def backend(self, startAddr, length, bitRate, uepFlag, protLevel):
        return Backend(self, SubCh(startAddr, length, bitRate, uepFlag, protLevel))


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
