"""GPU probe: per-entry-point CUDA-event times of ONE encoder training step (forward + backward, direct launches), to see
where the step's device time goes.   python tools/gpu_train_profile.py [batch]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200
from lrce_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [3], 30).cuda().train()
vf = torch.randn(B, 3, 3, 49, 1024, device="cuda").bfloat16()
tf = torch.randn(B, 30, 768, device="cuda")
tgt = torch.randint(0, 1000, (B,), device="cuda")


def step():
    for p in m.parameters():
        p.grad = None
    torch.nn.functional.cross_entropy(m(vf, tf, None), tgt).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"encoder train step (graphs): {e0.elapsed_time(e1) / 5:.2f} ms")
ops.trace = []
step()
torch.cuda.synchronize()
tr, ops.trace = ops.trace, None
agg = collections.OrderedDict()
for name, tag, fl, by, a, b in tr:
    k = (name, tag if "gemm" in name else "")
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
print(f"traced (direct launches, events around every launch): {len(tr)} launches, {tot:.2f} ms summed")
for (name, tag), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {name:34s} {tag:24s} x{n:4d}  {ms:8.3f} ms  ({1e3 * ms / n:7.1f} us each)")
