"""sdr-j-dab_b200 -- B200-native DAB baseband decode engine.

The product is the C-ABI shared library ``libdabgpu.so`` (include/dabgpu.h) built from ``csrc/``; this
package only holds its build recipe and a thin ctypes binding used by the tests and bench.py.  There is no
CPU fallback: without the built library, or without a CUDA device, everything here raises.
"""
from .binding import DabGpu, DabGroup, DabGpuError, Backend, SubCh, load_library, LIB_PATH   # noqa: F401
