"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel-family table.

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
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(x[ki], float(x[vi].replace(",", ""))) for x in r]
starts = [i for i, (k, _) in enumerate(data) if "patch_gather" in k]
s, e = starts[which], starts[which + 1] if which + 1 < len(starts) else len(data)
agg = collections.OrderedDict()
for k, v in data[s:e]:
    name = re.sub(r"<.*", "", k.split("(")[0]).replace("void ", "").replace("lrce::", "")
    d = agg.setdefault(name or "(unnamed)", [0, 0.0])
    d[0] += 1
    d[1] += v
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list summary: {path}, forward #{which} of {len(starts)} ({e - s} launches, {tot / 1e6:.3f} ms summed)\n")
print("| kernel | launches | ms (sum) | share |\n|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:80]}` | {n} | {v / 1e6:.3f} | {100 * v / tot:.1f}% |")
