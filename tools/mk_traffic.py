"""profiles/r02_traffic.json from ncu --set full captures of the 1024-frame launches:
   python tools/mk_traffic.py name=file.ncu-rep [name=file.ncu-rep ...]     (name = the key bench.py looks up)"""
import csv, io, json, subprocess, sys
out = {"source": "ncu --set full --clock-control none, one launch each inside `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras` "
                 "(tools/ncu_big.sh: the third launch of the largest grid); dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch of the 1024-frame batch",
       "frames_per_launch": 1024}
for arg in sys.argv[1:]:
    name, rep = arg.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    def get(metric):
        i = h.index(metric)
        x = float(v[i].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "": 1}.get(u[i], 1)
        return x * scale
    out[name] = {"read": int(get("dram__bytes_read.sum")), "write": int(get("dram__bytes_write.sum")), "warp_instructions": int(get("smsp__inst_executed.sum")),
                 "duration_us_under_ncu": get("gpu__time_duration.sum"), "grid": v[h.index("launch__grid_size")], "kernel": v[h.index("Kernel Name")]}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
