import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("sdr-j-dab_b200")
e = pkg.DabGpu(mode=1)
print({k: round(v / 1e12, 2) for k, v in e.int_peak().items()})
