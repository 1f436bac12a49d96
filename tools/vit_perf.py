"""Quick device-side timing of the Viterbi kernels (development aid, not the bench)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("sdr-j-dab_b200")
for path in (1, 2, 3):
    eng = pkg.DabGpu(mode=1, viterbi_path=min(path, 2))
    for frameBits, nblocks in ((3072, 4096), (3072, 36864), (768, 4096)):
        soft = torch.randint(-127, 128, (nblocks, 4 * (frameBits + 6)), dtype=torch.int16, device="cuda")
        out = torch.empty((nblocks, frameBits), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for _ in range(2):
            eng.viterbi_dev(soft.data_ptr(), frameBits, nblocks, out.data_ptr())
        eng.sync()
        eng.profile_enable(True); eng.profile_reset()
        eng.timer_begin()
        reps = 3
        for _ in range(reps):
            eng.viterbi_dev(soft.data_ptr(), frameBits, nblocks, out.data_ptr())
        ms = eng.timer_end() / reps
        prof = {k: round(v[1] / reps, 3) for k, v in eng.profile().items() if v[0]}
        steps = nblocks * (frameBits + 6)
        print(f"path {path} frameBits {frameBits} nblocks {nblocks}: {ms:.3f} ms  {steps / ms / 1e6:.2f} Gstep/s  {prof}", flush=True)
    eng.close()
