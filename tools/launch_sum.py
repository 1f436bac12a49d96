"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel totals and per grid-size medians.
usage: python tools/launch_sum.py launches.csv "command line that was profiled" """
import csv, sys, collections, statistics, re
rows = []
with open(sys.argv[1]) as f:
    for r in csv.reader(f):
        if len(r) >= 15 and r[12] == "gpu__time_duration.sum":
            rows.append((re.sub(r"\(.*", "", r[4]), r[8], float(r[14]) / 1e3))
print("ncu --metrics gpu__time_duration.sum --clock-control none; command:", sys.argv[2] if len(sys.argv) > 2 else "?")
tot = sum(r[2] for r in rows)
by = collections.defaultdict(list)
for n, g, t in rows: by[n].append(t)
for n, ts in sorted(by.items(), key=lambda kv: -sum(kv[1])):
    print("%-36s n=%4d total_us=%11.1f avg_us=%9.1f share=%.3f last=%s" % (n, len(ts), sum(ts), sum(ts) / len(ts), sum(ts) / tot, [round(x, 1) for x in ts[-3:]]))
print("\nper kernel and grid size (median device time, cold-cache serialised ncu replay):")
bg = collections.defaultdict(list)
for n, g, t in rows: bg[(n, g)].append(t)
for (n, g), ts in sorted(bg.items()):
    print("%-30s grid=%-15s n=%3d median_us=%9.1f" % (n, g, len(ts), statistics.median(ts)))
