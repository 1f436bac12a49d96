"""Key metrics + opcode/stall summary of one kernel in an .ncu-rep.  usage: python tools/ncu_sum.py file.ncu-rep"""
import csv, subprocess, io, collections, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[0]; v = rows[2]
keys = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread ', 'gpu__time_duration.sum', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', '_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'pipe_fma.avg.pct_of_peak_sustained_active',
        'pipe_alu.avg.pct_of_peak_sustained_active', 'pipe_lsu.avg.pct_of_peak_sustained_active', 'pipe_xu.avg.pct_of_peak_sustained_active',
        'pipe_fp64', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct', 'launch__occupancy_limit', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct', 'l1tex__t_sector_hit_rate']
for i, n in enumerate(h):
    if any(k in n for k in keys):
        try:
            if '_per_issue_active' in n and float(v[i]) < 0.05: continue
        except ValueError: pass
        print(n, '[%s]' % rows[1][i], '=', v[i])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
S, I, W, WI = ci["# Samples"], ci["Instructions Executed"], ci["L1 Wavefronts Shared"], ci["L1 Wavefronts Shared Ideal"]
print("smem wavefronts", sum(int(r[W] or 0) for r in data), "ideal", sum(int(r[WI] or 0) for r in data))
by = collections.defaultdict(lambda: [0, 0])
ti = sum(int(r[I]) for r in data); ts = sum(int(r[S]) for r in data)
for r in data:
    t = r[1].split(); op = t[1] if t[0].startswith('@') else t[0]
    by[op.rstrip(';')][0] += int(r[I]); by[op.rstrip(';')][1] += int(r[S])
print("%-26s %6s %6s" % ("opcode", "inst%", "samp%"))
for op, (i, s_) in sorted(by.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print("%-26s %6.2f %6.2f" % (op, 100 * i / ti, 100 * s_ / ts))
print("--- hottest instructions (index, sass, samples, executions)")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][S]))[:14]:
    print(k, r[1].strip()[:64], r[S], r[I])
