"""Summarise the SASS source page of an .ncu-rep: per-opcode instruction counts and stall samples, and the
hottest instructions.  usage: python tools/ncu_src.py file.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
S, I, T = ci["# Samples"], ci["Instructions Executed"], ci["Thread Instructions Executed"]
tot_s = sum(int(r[S]) for r in data); tot_i = sum(int(r[I]) for r in data)
print("instructions(warp)", tot_i, "samples", tot_s, "sass lines", len(data))
by = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    op = op.rstrip(";")
    by[op][0] += int(r[I]); by[op][1] += int(r[S])
print("%-28s %8s %8s" % ("opcode", "inst%", "samp%"))
for op, (i, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%-28s %8.2f %8.2f" % (op, 100 * i / tot_i, 100 * s / max(tot_s, 1)))
print("--- hottest instructions by samples")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][S]))[:25]:
    print(k, r[1].strip()[:70], r[S], r[I])
