"""Reduce an `ncu -i X.ncu-rep --page raw --csv` dump to the columns DESIGN.md / bench.py cite (one row per kernel launch).
    python tools/ncu_reduce.py gpurun_out/<tag>/prof_trunk_raw.csv profiles/<name>.csv"""
import csv, sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum"]

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["%s [%s]" % (hdr[i], units[i]) for i in idx])
    for r in rows[2:]:
        if len(r) == len(hdr):
            name = r[idx[0]]
            r = list(r)
            r[idx[0]] = name[:90]
            w.writerow([r[i] for i in idx])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
