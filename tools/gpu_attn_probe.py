"""GPU probe: window-attention kernel timing per Swin stage (CUDA events), config-2 shapes (96 segments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrce_b200
from lrce_b200 import ops, _lib as _l
if os.environ.get('ATTN_LIB'):
    _l.LIB_PATH = os.environ['ATTN_LIB']  # A/B variants of the library (tools only)
    print('library', _l.LIB_PATH)

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 96
stages = [("s1", 56, 128, 4), ("s2", 28, 256, 8), ("s3", 14, 512, 16), ("s4", 7, 1024, 32)]
only = sys.argv[2] if len(sys.argv) > 2 else None
for name, hw, C, heads in stages:
    if only and name != only:
        continue
    T = 3 * hw * hw
    qkv = (torch.randn(n_seg * T, 3 * C, device="cuda") * 1.0).bfloat16()
    bias = ops.window_bias_pack(torch.randn(2535, heads, device="cuda") * 0.5)
    out = torch.empty(n_seg * T, C, device="cuda", dtype=torch.bfloat16)
    for shift in ((0, 0), (3, 3) if hw > 7 else (0, 0)):
        for _ in range(2):
            ops.window_attention(qkv, bias, n_seg, 3, hw, hw, C, heads, shift, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        iters = 5
        for _ in range(iters):
            ops.window_attention(qkv, bias, n_seg, 3, hw, hw, C, heads, shift, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        items = n_seg * (hw // 7) ** 2 * heads
        flops = items * 4.0 * 147 * 147 * 32
        byts = qkv.numel() * 2 + out.numel() * 2
        print(f"{name} shift={shift} {ms*1e3:8.1f} us  items={items} core {flops/ms/1e9:6.1f} TFLOP/s  {byts/ms/1e6:7.1f} GB/s  {ms*1e6/items*296:.0f} ns/item/CTA", flush=True)

# ---- where the warps of CTA 0 wait (instrumented instantiation), stage 3 shifted
from lrce_b200 import _lib
hw, C, heads = 14, 512, 16
T = 3 * hw * hw
qkv = torch.randn(n_seg * T, 3 * C, device="cuda").bfloat16()
bias = ops.window_bias_pack(torch.randn(2535, heads, device="cuda") * 0.5)
buf = torch.zeros(224, dtype=torch.int64, device="cuda")
ops.window_attention(qkv, bias, n_seg, 3, hw, hw, C, heads, (3, 3), prof=buf)
torch.cuda.synchronize()
t = buf.cpu().tolist()
n, total = t[192], t[193]
print(f"CTA 0: {n} items, {total} cycles = {total / max(n, 1):.0f} cycles/item")
roles = {0: "softmax t0 q0 c0", 5: "softmax t0 q1 c1", 15: "softmax t0 q3 c3", 16: "softmax t1 q0", 19: "softmax t1 q3", 20: "loader",
         21: "mma tile 0", 22: "mma tile 1"}
for w, name in roles.items():
    print(f"  warp {w:2d} {name:20s} " + "  ".join(f"[{i}] {t[w * 8 + i] / max(n, 1):6.0f}" for i in range(8)) + "   cycles/item")
