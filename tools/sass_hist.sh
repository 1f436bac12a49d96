#!/bin/bash
# usage: tools/sass_hist.sh <kernel-substring> <out-name> [lib]  -- opcode histogram of one kernel's SASS -> profiles/r02_sass_<out-name>.txt
LIB=${3:-sdr-j-dab_b200/libdabgpu.so}
OUT=profiles/r02_sass_$2.txt
cuobjdump -sass $LIB | awk -v k="$1" '/Function :/ {on = index($0, k) > 0} on {print}' > /tmp/k.sass
{
  echo "cuobjdump -sass $LIB, kernel matching '$1' (sm_100a); static instruction counts"
  grep -m1 "Function :" /tmp/k.sass
  echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/k.sass)"
  grep -E "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn
} > $OUT
head -12 $OUT
