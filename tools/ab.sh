#!/bin/bash
# A/B of prebuilt library variants (sdr-j-dab_b200/variants/lib_<name>.so) on the GPU box: bench.py per variant
cd "$(dirname "$0")/.."
cp sdr-j-dab_b200/libdabgpu.so /tmp/libdabgpu.orig.so
for v in "$@"; do
  cp sdr-j-dab_b200/variants/lib_$v.so sdr-j-dab_b200/libdabgpu.so
  python bench.py --steps 5 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    print(v, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 4), {k: round(s * d["ms_per_step"], 4) for k, s in d["kernel_shares"].items()})
except Exception as e:
    print(v, "failed", e)
PY
done
cp /tmp/libdabgpu.orig.so sdr-j-dab_b200/libdabgpu.so
