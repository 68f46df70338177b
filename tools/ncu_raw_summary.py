"""Print the roofline-relevant counters of every kernel in an `ncu --page raw --csv` dump.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

r = csv.reader(open(sys.argv[1]))
hdr = next(r)
units = next(r)
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
extra = [i for i, h in enumerate(hdr) if "pipe_tensor" in h or "warp_issue_stalled" in h and h.endswith("ratio")]
for row in r:
    print("----")
    for w, i in idx:
        print(f"  {w}: {row[i][:100]} {units[i]}")
    if len(sys.argv) > 2:
        for i in extra:
            print(f"  {hdr[i]}: {row[i]} {units[i]}")
