#!/bin/bash
# ncu --set full capture of the LARGEST-grid launch of a kernel inside bench.py (the 1024-frame step, not the lead-in chunks):
#   tools/ncu_big.sh <kernel regex> <output name>      (run under gpurun; writes gpurun_out/<name>.ncu-rep and the launch list)
# pass 1 lists every launch of the kernel with its grid size and duration, pass 2 captures the first launch of the largest grid
cd "$(dirname "$0")/.."
K=$1; NAME=$2
CMD=${NCU_CMD:-"python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"}      # NCU_CMD: another workload (e.g. "python tools/mode_perf.py 2")
$CMD > gpurun_out/plain_$NAME.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$NAME.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K --csv --log-file gpurun_out/${NAME}_list.csv $CMD > /dev/null 2>&1
SKIP=$(python - gpurun_out/${NAME}_list.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; gi = h.index("Grid Size"); ii = h.index("ID")
seen, order = {}, []
for r in rows[1:]:
    if r[ii] not in seen:
        seen[r[ii]] = r[gi]; order.append(r[ii])
def size(g):
    g = g.strip("() ").split(",")
    n = 1
    for x in g: n *= int(x)
    return n
sizes = [size(seen[i]) for i in order]
big = max(sizes)
idx = [k for k, s in enumerate(sizes) if s == big]
print(idx[min(2, len(idx) - 1)])          # the third launch of that size when there are that many (warm)
PY
)
echo "capturing launch index $SKIP of $K"
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/$NAME -f $CMD > gpurun_out/ncu_$NAME.log 2>&1
tail -2 gpurun_out/ncu_$NAME.log
