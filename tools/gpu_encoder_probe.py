"""GPU probe: timing of the encoder (projection + pos-embeds + K/V GEMMs + the cluster-sharded token walk), per launch."""
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
