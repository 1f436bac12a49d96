#!/bin/bash
# e2e (host-input) throughput for several channel-decoding batch sizes
for hb in 64 128 256 512; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --host-batch $hb 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('host_batch', $hb, 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],2), 'value', round(d['value']))"
done
