"""Torch-tensor front end of the C ABI: argument checking, output allocation and the current-stream handle.
PyTorch is plumbing here (device memory + streams); all arithmetic happens inside liblrce_b200.so."""
import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_BIAS_LN = 0, 1, 2, 3


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.LrceError(f"{name} must be a CUDA tensor: the LRCE hot path has no CPU fallback")
    if t.dtype != dtype:
        raise _lib.LrceError(f"{name} must be {dtype}, got {t.dtype}")


def gemm(a, w, bias=None, *, epilogue=EPI_BIAS, residual=None, out=None, out_fp32=False, ln=None):
    """out = epilogue(a @ w.T); a (M,K) bf16 with unit inner stride, w (N,K) bf16; see lrce_gemm_bf16."""
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w"); _req(bias, torch.float32, "bias")
    _req(residual, torch.bfloat16, "residual")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    assert out.stride(1) == 1 and out.shape == (M, N)
    g = b = None
    eps = 0.0
    if ln is not None:
        g, b, eps = ln
        _req(g, torch.float32, "ln gamma"); _req(b, torch.float32, "ln beta")
    rc = _lib.lib().lrce_gemm_bf16(
        _ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(bias), _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), epilogue, int(out_fp32),
        _ptr(g), _ptr(b), float(eps), _stream())
    _lib.check(rc, "lrce_gemm_bf16")
    return out
