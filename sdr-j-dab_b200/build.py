"""Builds libdabgpu.so (the C-ABI engine) in-tree with nvcc for sm_100a.  No torch involved."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdabgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "--use_fast_math" if False else "-Xptxas", "-v"]


def sources():
    src = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cpp")))
    hdr = sorted(glob.glob(os.path.join(HERE, "csrc", "*.h")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) +
                 glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    return src, hdr


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    src, hdr = sources()
    return any(os.path.getmtime(f) > t for f in src + hdr)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    src, _ = sources()
    cmd = [NVCC] + FLAGS + ["-o", LIB] + src
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdabgpu.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
