"""GPU probe: the fused stage-1 MLP kernel (lrce_mlp_fused_bf16) against the two GEMMs it replaces, at the batch-32 shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lrce_b200 import _lib, ops

if os.environ.get("LRCE_LIB"):
    _lib.LIB_PATH = os.environ["LRCE_LIB"]  # timing variants of the library (tools/build_variants.sh)

M, C = int(sys.argv[1]) if len(sys.argv) > 1 else 903168, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(M, C, generator=g).bfloat16().cuda()
w1 = (torch.randn(4 * C, C, generator=g) * 0.05).bfloat16().cuda()
w2 = (torch.randn(C, 4 * C, generator=g) * 0.05).bfloat16().cuda()
b1, c1, b2 = torch.randn(4 * C).cuda(), w1.float().sum(1).contiguous(), torch.randn(C).cuda()
v = x.float().view(M, 4, 32)
mean = v.mean(-1)
st_in = torch.stack([mean, ((v - mean[..., None]) ** 2).sum(-1)], -1).contiguous().view(-1)
st_out = torch.zeros_like(st_in)
out = torch.empty_like(x)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def two():
    hid = ops.gemm(x, w1, b1, epilogue=ops.EPI_BIAS_GELU, ln_in=(st_in, c1, 1e-5))
    ops.gemm(hid, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=x, out=out, stats_out=st_out)


t2 = timed(two)
tf = timed(lambda: ops.mlp_fused(x, w1, b1, c1, st_in, 1e-5, w2, b2, out=out, stats_out=st_out))
flops = 2.0 * M * C * 4 * C * 2
print(f"M={M} C={C}: two GEMMs {t2:.1f} us, fused {tf:.1f} us ({flops / tf / 1e6:.0f} TFLOP/s; x in + out = {4 * M * C / tf / 1e3:.0f} GB/s; "
      f"weights re-streamed from L2: {(M + 127) // 128 * 8 * C * C * 2 / tf / 1e6:.2f} TB/s)")
