import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("sdr-j-dab_b200")
eng = pkg.DabGpu(mode=1, viterbi_path=2)
frameBits, nblocks = 3072, 36864
soft = torch.randint(-127, 128, (nblocks, 4 * (frameBits + 6)), dtype=torch.int16, device="cuda")
out = torch.empty((nblocks, frameBits), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.viterbi_dev(soft.data_ptr(), frameBits, nblocks, out.data_ptr())
eng.sync()
print("ok")
