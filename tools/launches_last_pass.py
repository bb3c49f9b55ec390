"""Turns an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
`bench.py --profile --steps 1 --warmup 1` into (a) the launch list of the LAST pass over the pipeline (one row per launch:
id, kernel, grid, ms, DRAM bytes read / written) and (b) a per-kernel summary.
usage: python tools/launches_last_pass.py ncu_log.csv out_last_pass.csv out_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

src, out_csv, out_txt = sys.argv[1:4]
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
launch = {}
for r in csv.DictReader(lines):
    i = int(r["ID"])
    d = launch.setdefault(i, {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d["read" if "read" in r["Metric Name"] else "write"] = v * scale
ids = sorted(launch)
# the pipeline starts with the rasteriser's first kernel: the last pass begins at its last occurrence
starts = [i for i in ids if "raster_xform" in launch[i]["kernel"] or "raster_clear" in launch[i]["kernel"]]
first = max(i for i in starts if not any(j in starts for j in (i - 1,)))  # first kernel of the last raster group
last_pass = [i for i in ids if i >= first]
with open(out_csv, "w") as f:
    f.write("id,kernel,grid,ms,dram_read_bytes,dram_write_bytes\n")
    for i in last_pass:
        d = launch[i]
        f.write(f'{i},"{d["kernel"].split("(")[0]}","{d["grid"]}",{d["ms"]:.6f},{int(d.get("read", 0))},{int(d.get("write", 0))}\n')
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in last_pass:
    d = launch[i]
    name = re.sub(r"^.*::", "", d["kernel"].split("(")[0])
    a = agg[name]
    a[0] += 1; a[1] += d["ms"]; a[2] += d.get("read", 0); a[3] += d.get("write", 0)
tot = sum(a[1] for a in agg.values())
rd = sum(a[2] for a in agg.values()); wr = sum(a[3] for a in agg.values())
cnn = [a for n, a in agg.items() if n.startswith(("conv_", "pool", "bn_relu", "image_to", "peaks", "moment"))]
with open(out_txt, "w") as f:
    f.write("One scan of the headline workload under ncu (python bench.py --profile --steps 1 --warmup 1; --clock-control none; per-launch times are cold-cache and\n"
            f"serialised: compare shares).  Last pass = {len(last_pass)} launches, {tot:.3f} ms; DRAM {rd / 1e9:.2f} GB read + {wr / 1e9:.2f} GB written;\n"
            f"CNN-stage kernels {sum(a[1] for a in cnn):.3f} ms = {100 * sum(a[1] for a in cnn) / tot:.1f} % of the pass.\n\n")
    f.write("      ms   share  launches   DRAM read GB  written GB   kernel\n")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{a[1]:8.3f}  {100 * a[1] / tot:5.1f}%   x{a[0]:4d}     {a[2] / 1e9:8.3f}    {a[3] / 1e9:8.3f}    {n}\n")
print(open(out_txt).read())
