"""One launch each of the kernels changed late in round 2, for a single `ncu --set full` capture:
mlp_l2_kernel at the stage-3 and stage-2 shapes, window_attention_kernel at stage 3 (96 segments, unshifted and shifted)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200  # noqa: F401
from lrce_b200 import ops

dev = "cuda"
torch.manual_seed(0)
for M, C in ((56448, 512), (225792, 256)):
    x = (torch.randn(M, C, device=dev) * 0.8).bfloat16()
    w1 = (torch.randn(4 * C, C, device=dev) * 0.04).bfloat16()
    w2 = (torch.randn(C, 4 * C, device=dev) * 0.04).bfloat16()
    b1, c1, b2 = torch.randn(4 * C, device=dev) * 0.1, torch.randn(4 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
    nc = C // ops.stats_chunk(C)
    st_in, st_out = torch.rand(M * nc * 2, device=dev) + 0.5, torch.empty(M * nc * 2, device=dev)
    out = torch.empty_like(x)
    torch.cuda.synchronize()
    ops.mlp_l2(x, w1, b1, c1, st_in, 1e-5, w2, b2, out=out, stats_out=st_out)
    torch.cuda.synchronize()
    del x, out
n_seg, hw, C, heads = 96, 14, 512, 16
qkv = torch.randn(n_seg * 3 * hw * hw, 3 * C, device=dev).bfloat16()
bias = ops.window_bias_pack(torch.randn(2535, heads, device=dev) * 0.5)
for shift in ((0, 0), (3, 3)):
    torch.cuda.synchronize()
    ops.window_attention(qkv, bias, n_seg, 3, hw, hw, C, heads, shift)
    torch.cuda.synchronize()
print("PICK_DONE")
