"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list of
bench.py into a per-kernel-family table (one forward, picked by index among the patch_gather launches).

    python tools/summarize_launches.py gpurun_out/r01_launches.csv [forward_index] > profiles/rNN_launches_summary.md

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes."""
import collections
import csv
import re
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, mi, vi, ui, idi = (hdr.index(h) for h in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
recs = collections.OrderedDict()
for x in r:
    d = recs.setdefault(x[idi], {"k": x[ki], "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(x[vi].replace(",", "")) * scale.get(x[ui], 1.0)
    if x[mi].startswith("gpu__time"):
        d["t"] = v
    elif "read" in x[mi]:
        d["rd"] = v
    elif "write" in x[mi]:
        d["wr"] = v
data = list(recs.values())
starts = [i for i, d in enumerate(data) if "patch_gather" in d["k"]]
s, e = starts[which], starts[which + 1] if which + 1 < len(starts) else len(data)
agg = collections.OrderedDict()
for d in data[s:e]:
    name = re.sub(r"<.*", "", d["k"].split("(")[0]).replace("void ", "").replace("lrce::", "")
    a = agg.setdefault(name or "(unnamed)", [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["t"]
    a[2] += d["rd"]
    a[3] += d["wr"]
tot = sum(v[1] for v in agg.values())
have_bytes = any(v[2] or v[3] for v in agg.values())
print(f"# ncu launch list summary: {path}, forward #{which} of {len(starts)} ({e - s} launches, {tot / 1e6:.3f} ms summed)\n")
print("| kernel | launches | ms (sum) | share |" + (" DRAM read MB | DRAM write MB |" if have_bytes else ""))
print("|---|---|---|---|" + ("---|---|" if have_bytes else ""))
for k, (n, v, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    row = f"| `{k[:80]}` | {n} | {v / 1e6:.3f} | {100 * v / tot:.1f}% |"
    if have_bytes:
        row += f" {rd / 1e6:.1f} | {wr / 1e6:.1f} |"
    print(row)
