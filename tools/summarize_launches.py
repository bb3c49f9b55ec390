"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and,
for the conv kernel, per-grid/shape rows.  Usage: python tools/summarize_launches.py launches.csv [first_id last_id]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], ns))
    if len(sys.argv) > 3:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        rows = [r for r in rows if lo <= r[0] <= hi]
    tot = sum(r[4] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, grid, block, ns in rows:
        short = name.split("(")[0]
        agg[short][0] += 1
        agg[short][1] += ns
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total (ncu per-launch times are serialised / cold-cache: compare shares)")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:9.3f} ms {100 * ns / tot:5.1f}%  x{n:4d}  {name}")
    return rows


if __name__ == "__main__":
    main()
