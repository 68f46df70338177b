"""GPU probe: timing of the encoder (K/V GEMMs + persistent token walk) and, through the lrce_debug_walk_timing hook,
the per-phase breakdown of the walk kernel as seen by CTA 0 (work time vs grid-barrier wait per phase)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200
from lrce_b200 import _lib, ops

if os.environ.get('ATTN_LIB'):
    _lib.LIB_PATH = os.environ['ATTN_LIB']  # A/B variants of the library (tools only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "oe"
S = 3
if kind == "mc":
    m = lrce_b200.LRCEMultipleChoice(768, 1, 0.1, [7, 7], 1024, 5, [S], 40).cuda().eval()
    tf = torch.randn(B, 5, 40, 768, device="cuda")
else:
    m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [S], 32).cuda().eval()
    tf = torch.randn(B, 32, 768, device="cuda")
vf = torch.randn(B, S, 3, 49, 1024, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(3):
        m(vf, tf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(vf, tf)
    e1.record(); torch.cuda.synchronize()
    print(f"encoder forward B={B} {kind}: {e0.elapsed_time(e1)/5:.3f} ms")
    ops.trace = []
    m(vf, tf)
    torch.cuda.synchronize()
    tr, ops.trace = ops.trace, None
    for name, tag, fl, by, a, b in tr:
        print(f"  {name:28s} {tag:24s} {a.elapsed_time(b)*1e3:8.1f} us")
    n = 3 + 18 * S * 12
    buf = torch.zeros(n, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().lrce_debug_walk_timing(buf.data_ptr()), "timing hook")
    m(vf, tf)
    torch.cuda.synchronize()
    _lib.lib().lrce_debug_walk_timing(0)
t = buf.cpu().tolist()
names = ["P1 self", "P2 q", "P3 attn", "P4 out", "P5 fc1", "P6 fc2"]
pro, work, wait = [0.0] * 6, [0.0] * 6, [0.0] * 6
# stamps: [start] then per phase (after prologue / staging, before barrier, after barrier)
for i in range(S * 12 * 6):
    prev = t[3 * i]  # after the previous barrier (or kernel start)
    a, done, released = t[3 * i + 1], t[3 * i + 2], t[3 * i + 3]
    pro[i % 6] += a - prev
    work[i % 6] += done - a
    wait[i % 6] += released - done
k = S * 12
for j in range(6):
    print(f"  {names[j]:8s} prologue {pro[j]/k/1e3:6.2f} us  tiles {work[j]/k/1e3:6.2f} us  barrier wait {wait[j]/k/1e3:6.2f} us"
          f"   (CTA 1, mean over {k} layer-steps)")
print(f"  head     {(t[-1]-t[-3])/1e3:7.2f} us ; whole walk {(t[-1]-t[0])/1e3:8.1f} us")
