#!/bin/bash
# A/B of an environment knob on the GPU box: tools/ab_env.sh NAME v1 v2 ...
cd "$(dirname "$0")/.."
name=$1; shift
for v in "$@"; do
  env $name=$v python bench.py --steps 5 --no-cpu-baseline --no-extras > gpurun_out/abe_$v.json 2> gpurun_out/abe_$v.err
  python - "$name=$v" gpurun_out/abe_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 4), {k: round(s * d["ms_per_step"], 4) for k, s in d["kernel_shares"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
