#!/bin/bash
# e2e (host-input, streaming caller) throughput for several upload piece sizes (DABGPU_PIECE_MSAMPLES, 2 bytes per sample)
for pm in 4 8 16 32 64; do
  DABGPU_PIECE_MSAMPLES=$pm python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('piece_msamples', $pm, 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],3), 'plain', round(d['e2e']['plain_calls']['ms_per_step'],3), 'packed', round(d['e2e']['packed_output']['ms_per_step'],3), 'floor', round(d['e2e']['h2d_floor_ms'],3))"
done
