"""compute-sanitizer driver: every hand-written kernel of liblrce_b200 once at a SMALL shape (the sanitizer slows kernels
10-100x), results checked against torch so a silent corruption is caught too.

    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
    compute-sanitizer --tool synccheck python tools/gpu_sanitize.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch

import lrce_b200
from lrce_b200 import ops

dev = "cuda"
torch.manual_seed(0)
which = sys.argv[1:] or ["gemm", "attn", "rows", "enc", "bert"]


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()


if "gemm" in which:
    for M, N, K, epi in [(300, 128, 96, ops.EPI_BIAS_LN), (300, 384, 128, ops.EPI_BIAS), (520, 512, 256, ops.EPI_BIAS_GELU),
                         (520, 256, 512, ops.EPI_BIAS_RESIDUAL), (300, 128, 128, ops.EPI_BIAS_RESIDUAL)]:
        a, w, bias = rnd(M, K, scale=0.5), rnd(N, K, scale=0.1), torch.randn(N, device=dev)
        res = rnd(M, N) if epi == ops.EPI_BIAS_RESIDUAL else None
        ln = (torch.ones(N, device=dev), torch.zeros(N, device=dev), 1e-5) if epi == ops.EPI_BIAS_LN else None
        st = torch.empty(N // ops.stats_chunk(N) * M * 2, device=dev) if epi == ops.EPI_BIAS_RESIDUAL else None
        y = ops.gemm(a, w, bias, epilogue=epi, residual=res, ln=ln, stats_out=st)
        ref = a.float() @ w.float().T + bias
        if epi == ops.EPI_BIAS_GELU:
            ref = torch.nn.functional.gelu(ref)
        if epi == ops.EPI_BIAS_RESIDUAL:
            ref = ref + res.float()
        if epi == ops.EPI_BIAS_LN:
            ref = torch.nn.functional.layer_norm(ref, (N,))
        print(f"gemm M{M} N{N} K{K} epi{epi}: rel {rel(y, ref):.2e}")
        assert rel(y, ref) < 1e-2
    if hasattr(ops, "mlp_fused"):
        pass

if "attn" in which:
    for hw, C, heads, shift in [(14, 128, 4, (3, 3)), (7, 64, 2, (0, 0))]:
        n_seg = 1
        qkv = rnd(n_seg * 3 * hw * hw, 3 * C)
        bias = ops.window_bias_pack(torch.randn(2535, heads, device=dev) * 0.5)
        out = ops.window_attention(qkv, bias, n_seg, 3, hw, hw, C, heads, shift)
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all()
        print(f"attn hw{hw} C{C} heads{heads} shift{shift}: ok, |out| {out.float().abs().mean().item():.3f}")

if "rows" in which:
    clips = torch.rand(1, 5, 3, 56, 56, device=dev)
    a = ops.patch_gather(clips)
    a8 = ops.patch_gather((clips * 255).round().to(torch.uint8))
    x = rnd(3 * 14 * 14, 128)
    ops.layernorm(x, torch.ones(128, device=dev), torch.zeros(128, device=dev), 1e-5)
    ops.patch_merge_ln(x, torch.ones(512, device=dev), torch.zeros(512, device=dev), 1e-5, 1, 3, 14, 14, 128)
    y = ops.window_remap(x, 1, (3, 14, 14), (3, 7, 7), (0, 3, 3))
    z = ops.window_remap(y, 1, (3, 14, 14), (3, 7, 7), (0, 3, 3), inverse=True)
    assert torch.equal(x, z)
    ops.remap_index((3, 14, 14), (3, 7, 7), (0, 3, 3))
    torch.cuda.synchronize()
    print("row kernels: ok")

if "enc" in which:
    import weights as W

    for kind, cls, ncls, L in (("oe", lrce_b200.LRCEOpenEnded, 1000, 32), ("mc", lrce_b200.LRCEMultipleChoice, 1, 40)):
        m = cls(768, ncls, 0.1, [7, 7], 1024, 5, [1], L)
        m.load_state_dict(W.make_fusion_state_dict(ncls, L, 1, seed=0), strict=True)
        m = m.cuda().eval()
        vf = rnd(2, 1, 3, 49, 1024)
        tf = torch.randn((2, 5, L, 768) if kind == "mc" else (2, L, 768), device=dev)
        with torch.no_grad():
            y = m(vf, tf)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
        print(f"encoder {kind} S=1: ok {tuple(y.shape)}")

if "bert" in which and hasattr(ops, "bert_attention"):
    import weights as W

    te = lrce_b200.TextExtractor(pretrained=False)
    te.bert.load_state_dict(W.make_bert_state_dict(seed=0), strict=True)  # any weights do here
    te = te.cuda().eval()
    _, ids, mask, types = W.make_inputs(2, 1, 32, seed=1)
    with torch.no_grad():
        t = te(ids.cuda(), mask.cuda(), types.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(t.float()).all()
    print("bert: ok")
print("SANITIZE_RUN_OK")
