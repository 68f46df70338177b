"""GPU stress: run the encoder (K/V GEMMs + cluster-sharded token walk) many times per shape and require bit-identical
results, to expose rare races / hangs in the walk's mbarrier, ring and exchange protocols. Run under `timeout`; every line
is flushed, so the last line printed tells which launch never returned. Also runs the fused MLP kernel the same way."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200
from lrce_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
torch.manual_seed(0)
bad = 0
for kind, B in [("oe", 1), ("oe", 5), ("oe", 32), ("oe", 60), ("mc", 32), ("count", 7)]:
    if kind == "mc":
        m = lrce_b200.LRCEMultipleChoice(768, 1, 0.1, [7, 7], 1024, 5, [3], 40).cuda().eval()
        tf = torch.randn(B, 5, 40, 768, device="cuda")
    elif kind == "count":
        m = lrce_b200.LRCECount(768, 1, 0.1, [7, 7], 1024, 5, [3], 30).cuda().eval()
        tf = torch.randn(B, 30, 768, device="cuda")
    else:
        m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [3], 32).cuda().eval()
        tf = torch.randn(B, 32, 768, device="cuda")
    vf = torch.randn(B, 3, 3, 49, 1024, device="cuda").bfloat16()
    with torch.no_grad():
        ref = m(vf, tf).clone()
        torch.cuda.synchronize()
        for i in range(reps):
            y = m(vf, tf)
            if i % 7 == 0:  # vary the timing between launches
                torch.cuda.synchronize()
            if not torch.equal(y, ref):
                bad += 1
                print(f"{kind} B={B} launch {i}: differs from launch 0 (max abs {(y - ref).abs().max().item():.3e})", flush=True)
        torch.cuda.synchronize()
    print(f"{kind} B={B}: {reps} launches ok, finite={torch.isfinite(ref).all().item()}", flush=True)

M, C = 148 * 128 * 3 + 50, 128
x = torch.randn(M, C, device="cuda").bfloat16()
w1, w2 = (torch.randn(4 * C, C, device="cuda") * 0.05).bfloat16(), (torch.randn(C, 4 * C, device="cuda") * 0.05).bfloat16()
b1, c1, b2 = torch.randn(4 * C, device="cuda"), w1.float().sum(1).contiguous(), torch.randn(C, device="cuda")
v = x.float().view(M, 4, 32)
mean = v.mean(-1)
st_in = torch.stack([mean, ((v - mean[..., None]) ** 2).sum(-1)], -1).contiguous().view(-1)
ref = ops.mlp_fused(x, w1, b1, c1, st_in, 1e-5, w2, b2).clone()
for i in range(reps):
    y = ops.mlp_fused(x, w1, b1, c1, st_in, 1e-5, w2, b2)
    if not torch.equal(y, ref):
        bad += 1
        print(f"mlp_fused launch {i}: differs from launch 0", flush=True)
torch.cuda.synchronize()
print(f"mlp_fused M={M}: {reps} launches ok", flush=True)
print("STRESS_DONE" if bad == 0 else f"STRESS_FAILED ({bad} mismatches)", flush=True)
