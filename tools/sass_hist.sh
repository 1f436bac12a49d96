#!/bin/bash
# usage: tools/sass_hist.sh <kernel-substring> [lib]  -- opcode histogram of one kernel's SASS + loop ranges
LIB=${2:-sdr-j-dab_b200/libdabgpu.so}
cuobjdump -sass $LIB | awk -v k="$1" '/Function :/ {on = index($0, k) > 0} on {print}' > /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass | wc -l
grep -nE "BRA 0x" /tmp/k.sass | awk '{print}' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//' | head -40
