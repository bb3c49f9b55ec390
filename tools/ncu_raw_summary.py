"""Key metrics per kernel launch from `ncu -i report.ncu-rep --page raw --csv` (text table for profiles/).
usage: python tools/ncu_raw_summary.py raw.csv [more_raw.csv ...]"""
import csv
import sys

KEYS = [
    "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
]
for path in sys.argv[1:]:
    with open(path) as f:
        rows = list(csv.reader(ln for ln in f if not ln.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        print(f"  {name}")
        for k in KEYS:
            if k in col:
                print(f"    {k:95s} {r[col[k]]:>16s} {units[col[k]]}")
        print()
