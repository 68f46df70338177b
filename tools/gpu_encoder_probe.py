"""GPU probe: per-launch timing of the recurrent encoder's kernels by shape tag."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrce_b200
from lrce_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [3], 32).cuda().eval()
vf = torch.randn(B, 3, 3, 49, 1024, device="cuda").bfloat16()
tf = torch.randn(B, 32, 768, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(vf, tf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(vf, tf)
    e1.record(); torch.cuda.synchronize()
    print(f"encoder forward B={B}: {e0.elapsed_time(e1)/5:.3f} ms")
    ops.trace = []
    m(vf, tf)
    torch.cuda.synchronize()
    tr, ops.trace = ops.trace, None
agg = collections.OrderedDict()
for i, (name, tag, fl, by, a, b) in enumerate(tr):
    key = (name, tag) if name != "lrce_skinny_linear" else (name, f"call{(i - 5) % 6 if i >= 5 else i}")
    d = agg.setdefault(key, [0.0, 0])
    d[0] += a.elapsed_time(b); d[1] += 1
for k, (ms, n) in agg.items():
    print(f"{k[0]:28s} {k[1]:24s} n={n:4d} avg {ms/n*1e3:8.1f} us total {ms:7.3f} ms")
