"""Opcode histogram and hottest SASS lines of one kernel from `ncu --page source --csv` output.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-skip N --launch-count 1 > src.csv
    python tools/ncu_source_summary.py src.csv [n_top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if "# Samples" in r)
si, smp, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows:
    if len(r) > max(smp, ie) and r[smp].isdigit() and r[0].startswith("0x"):
        data.append((r[si].strip(), int(r[smp]), int(r[ie] or 0)))
tot = sum(d[1] for d in data)
toti = sum(d[2] for d in data)
print("total samples", tot, "warp-instructions", toti, "sass lines", len(data))
h, hs = collections.Counter(), collections.Counter()
for s, n, i in data:
    t = s.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    h[op] += i
    hs[op] += n
for op, c in h.most_common(28):
    print(f"{op:10s} instr {100 * c / toti:5.1f}%  samples {100 * hs[op] / tot:5.1f}%")
print("--- top sampled instructions")
for s, n, i in sorted(data, key=lambda d: -d[1])[:ntop]:
    print(f"{100 * n / tot:5.1f}% exec={i:8d} {s[:110]}")
