#!/bin/bash
# dabgpu_decode_multi from host memory: number of stream groups of the upload / size of the heads (frames)
for cfg in "1 6.6" "2 6.6" "4 6.6" "8 6.6" "16 6.6" "32 6.6" "8 3"; do
  set -- $cfg
  DABGPU_MULTI_PIECES=$1 DABGPU_MULTI_FIRST=$2 python tools/multi_bench.py 32 256 2>/dev/null | python -c "
import sys, ast
d = ast.literal_eval(sys.stdin.readline())
print('groups $1 heads $2 frames: host %.2f ms (%.0f frames/s), dev %.2f ms, equal %s %s' % (d['ms_per_call_host'], d['e2e_frames_per_s'], d['ms_per_call_dev'], d['equal_to_single_handle'], d['dev_equals_host_input']))"
done
