"""GPU probe: launch every distinct (kernel, shape) of the config-2 forward (B=32 clips -> 96 segments) exactly once, in a
fixed printed order, so that one `ncu --set full` capture of this script holds one record per hot-path kernel shape.

    python tools/gpu_kernel_zoo.py [gemm|attn|rows|enc|all]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200  # noqa: F401
from lrce_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
torch.manual_seed(0)
N_SEG = int(os.environ.get("ZOO_SEG", "96"))  # use ZOO_SEG=24 under `ncu --set full` (replays save/restore every buffer)
SC = N_SEG / 96.0
order = []


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


def gemm(tag, M, N, K, epi, lnin=False, stats=False):
    M = int(M * SC)
    a, w = rnd(M, K, scale=0.5), rnd(N, K, scale=0.05)
    bias = torch.randn(N, device=dev)
    res = rnd(M, N) if epi == ops.EPI_BIAS_RESIDUAL else None
    ln = (torch.ones(N, device=dev), torch.zeros(N, device=dev), 1e-5) if epi == ops.EPI_BIAS_LN else None
    kw = {}
    if lnin:
        kw["ln_in"] = (torch.rand(K // 32 * M * 2, device=dev) + 0.5, torch.randn(N, device=dev), 1e-5)
    if stats:
        kw["stats_out"] = torch.empty(N // 32 * M * 2, device=dev)
    torch.cuda.synchronize()
    ops.gemm(a, w, bias, epilogue=epi, residual=res, ln=ln, **kw)
    torch.cuda.synchronize()
    order.append(f"gemm {tag} M{M} N{N} K{K} epi{epi} lnin{int(lnin)} stats{int(stats)}")


if which in ("gemm", "all"):
    gemm("patch-embed", 903168, 128, 96, ops.EPI_BIAS_LN)
    for st, (M, C) in enumerate([(903168, 128), (225792, 256), (56448, 512), (14112, 1024)], 1):
        gemm(f"s{st}.qkv", M, 3 * C, C, ops.EPI_BIAS, lnin=True)
        gemm(f"s{st}.proj", M, C, C, ops.EPI_BIAS_RESIDUAL, stats=True)
        gemm(f"s{st}.fc1", M, 4 * C, C, ops.EPI_BIAS_GELU, lnin=True)
        gemm(f"s{st}.fc2", M, C, 4 * C, ops.EPI_BIAS_RESIDUAL, stats=True)
        if st < 4:
            gemm(f"merge{st}", M // 4, 2 * C, 4 * C, ops.EPI_BIAS)
    gemm("video-proj", 14112, 768, 1024, ops.EPI_BIAS)
    gemm("enc-kv", 14400, 18432, 768, ops.EPI_BIAS)

if which == "pick":  # the few launches worth an `ncu --set full --import-source on` capture
    gemm("s1.qkv", 903168, 384, 128, ops.EPI_BIAS, lnin=True)
    gemm("s1.fc1", 903168, 512, 128, ops.EPI_BIAS_GELU, lnin=True)
    gemm("s3.proj", 56448, 512, 512, ops.EPI_BIAS_RESIDUAL, stats=True)
    gemm("s3.fc1", 56448, 2048, 512, ops.EPI_BIAS_GELU, lnin=True)
    qkv = rnd(N_SEG * 588, 1536)
    bias = ops.window_bias_pack(torch.randn(2535, 16, device=dev) * 0.5)
    torch.cuda.synchronize()
    ops.window_attention(qkv, bias, N_SEG, 3, 14, 14, 512, 16, (3, 3))
    torch.cuda.synchronize()
    order.append("attn s3 shift(3,3)")

if which == "s2":  # the four GEMMs of a stage-2 block (K = 256 / 1024: short main loops)
    gemm("s2.qkv", 225792, 768, 256, ops.EPI_BIAS, lnin=True)
    gemm("s2.proj", 225792, 256, 256, ops.EPI_BIAS_RESIDUAL, stats=True)
    gemm("s2.fc1", 225792, 1024, 256, ops.EPI_BIAS_GELU, lnin=True)
    gemm("s2.fc2", 225792, 256, 1024, ops.EPI_BIAS_RESIDUAL, stats=True)

if which == "attn1":  # one stage-3 shifted attention launch (for ncu --set full --import-source on)
    qkv = rnd(N_SEG * 588, 1536)
    bias = ops.window_bias_pack(torch.randn(2535, 16, device=dev) * 0.5)
    torch.cuda.synchronize()
    ops.window_attention(qkv, bias, N_SEG, 3, 14, 14, 512, 16, (3, 3))
    torch.cuda.synchronize()
    order.append("attn s3 shift(3,3)")

if which in ("attn", "all"):
    for name, hw, C, heads in [("s1", 56, 128, 4), ("s2", 28, 256, 8), ("s3", 14, 512, 16), ("s4", 7, 1024, 32)]:
        T = 3 * hw * hw
        qkv = rnd(N_SEG * T, 3 * C)
        bias = ops.window_bias_pack(torch.randn(2535, heads, device=dev) * 0.5)
        for shift in ((0, 0), (3, 3)) if hw > 7 else ((0, 0),):
            torch.cuda.synchronize()
            ops.window_attention(qkv, bias, N_SEG, 3, hw, hw, C, heads, shift)
            torch.cuda.synchronize()
            order.append(f"attn {name} shift{shift}")

if which in ("rows", "all"):
    clips = torch.rand(N_SEG, 5, 3, 224, 224, device=dev)
    ops.patch_gather(clips)
    order.append("patch_gather")
    for M, C in [(903168, 128), (225792, 256), (56448, 512), (14112, 1024)]:
        M = int(M * SC)
        x = rnd(M, C)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ops.layernorm(x, g, b, 1e-5)
        order.append(f"layernorm M{M} C{C}")
    for hw, C in [(56, 128), (28, 256), (14, 512)]:
        x = rnd(N_SEG * 3 * hw * hw, C)
        g, b = torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev)
        ops.patch_merge_ln(x, g, b, 1e-5, N_SEG, 3, hw, hw, C)
        order.append(f"patch_merge_ln C{C}")

if which == "mem":  # the memory-bound pieces north_star names (remap, LN, bias gather, patch gather, merge-LN), full size
    clips = torch.rand(N_SEG, 5, 3, 224, 224, device=dev)
    ops.patch_gather(clips)
    order.append(f"patch_gather_f32 in {clips.numel() * 4 / 1e6:.0f} MB out {N_SEG * 9408 * 96 * 2 / 1e6:.0f} MB")
    ops.patch_gather((clips * 255).to(torch.uint8))
    order.append("patch_gather_u8")
    for hw, C in [(56, 128), (28, 256), (14, 512)]:
        x = rnd(N_SEG * 3 * hw * hw, C)
        ops.window_remap(x, N_SEG, (3, hw, hw), (3, 7, 7), (0, 3, 3))
        order.append(f"window_remap C{C}: algorithmic {2 * x.numel() * 2 / 1e6:.0f} MB")
        g, b = torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev)
        ops.patch_merge_ln(x, g, b, 1e-5, N_SEG, 3, hw, hw, C)
        order.append(f"patch_merge_ln C{C}: algorithmic {2 * x.numel() * 2 / 1e6:.0f} MB")
    x = rnd(N_SEG * 147, 1024)
    ops.layernorm(x, torch.ones(1024, device=dev), torch.zeros(1024, device=dev), 1e-5)
    order.append(f"layernorm C1024: algorithmic {2 * x.numel() * 2 / 1e6:.0f} MB")
    for heads in (4, 16, 32):
        ops.window_bias_pack(torch.randn(2535, heads, device=dev) * 0.5)
        order.append(f"window_bias_pack heads{heads}: table {2535 * heads * 4 / 1e3:.0f} KB -> dense {heads * 160 * 160 * 2 / 1e3:.0f} KB")
    vf = rnd(32 * 3 * 3 * 49, 768)
    e = lambda *s_: torch.randn(*s_, device=dev)
    ops.video_posembed_ln(vf, e(768), e(50, 768), e(3, 768), e(3, 768), e(768), e(768), 1e-12, 32, 3, 3, 49)
    order.append(f"video_posembed_ln: algorithmic {(vf.numel() + 32 * 3 * 150 * 768) * 2 / 1e6:.0f} MB")

if which in ("enc", "all"):
    m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [1], 32).cuda().eval()  # S=1: one recurrent step
    vf = rnd(32, 1, 3, 49, 1024)
    tf = torch.randn(32, 32, 768, device=dev)
    with torch.no_grad():
        m(vf, tf)
    torch.cuda.synchronize()
    order.append("encoder S=1 forward (K/V GEMMs + one persistent walk kernel)")

print("\n".join(order))
