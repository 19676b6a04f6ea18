"""Summarise an `ncu --page raw --csv` dump into the per-kernel table kept under profiles/.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv out.csv ["label 0" "label 1" ...]"""
import csv
import sys

COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
cols = [c for c in COLS if c in ci]
labels = sys.argv[3:]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[ci[c]] for c in cols])
    for i, b in enumerate(body):
        r = [b[ci[c]] for c in cols]
        if i < len(labels):
            r[1] += "  // " + labels[i]
        w.writerow(r)
print(f"{len(body)} kernels -> {sys.argv[2]}")
