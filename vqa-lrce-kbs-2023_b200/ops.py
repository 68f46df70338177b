"""Torch-tensor front end of the C ABI: argument checking, output allocation and the current-stream handle.
PyTorch is plumbing here (device memory + streams); all arithmetic happens inside liblrce_b200.so."""
import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_BIAS_LN = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
WINDOW = (3, 7, 7)  # effective Swin window on LRCE's 5-frame segments
BIAS_PITCH = 160

# count of kernels launched through this module (bench.py reports it as gpu_launches)
launches = 0
# label appended to the entry-point names in `trace` ("bert", "train"): keeps BERT's / the training step's GEMMs out of the
# Swin + encoder GEMM family that bench.py's roofline is computed on
scope = ""


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.LrceError(f"{name} must be a CUDA tensor: the LRCE hot path has no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        # launches go to the current device's current stream; a tensor of another GPU would be dereferenced there
        raise _lib.LrceError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                             "wrap the call in torch.cuda.device(...)")
    if t.dtype != dtype:
        raise _lib.LrceError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous() and t.dim() != 2:
        raise _lib.LrceError(f"{name} must be contiguous")


# when `trace` is a list, every launch is bracketed by CUDA events on the launching stream and appended as
# (entry point, tag, flops, bytes, start_event, end_event); bench.py uses it for the per-kernel roofline figures
trace = None


def _call(name, *args, work=None):
    global launches
    launches += 1
    if trace is None:
        _lib.check(getattr(_lib.lib(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    e1.record()
    tag, flops, nbytes = work if work is not None else ("", 0, 0)
    trace.append((name + ("@" + scope if scope else ""), tag, flops, nbytes, e0, e1))


def stats_chunk(width):
    """column width of the LayerNorm partials a GEMM with `width` output features emits (lrce_gemm_bf16 out_stats)"""
    return 64 if width % 256 == 0 else 32


def gemm(a, w, bias=None, *, epilogue=EPI_BIAS, residual=None, out=None, out_fp32=False, ln=None, ln_in=None,
         stats_out=None):
    """out = epilogue(a @ w.T); a (M,K) bf16 with unit inner stride, w (N,K) bf16; see lrce_gemm_bf16.
    ln_in = (stats fp32 [M, K/cw, 2], colsum fp32 [N], eps): LayerNorm folded into the A operand (w, bias pre-folded);
    stats_out = fp32 [M, N/cw, 2] buffer that receives the (mean, M2) partials of the rows written.
    cw = stats_chunk(width) of the PRODUCING GEMM: 64 columns when its N % 256 == 0, else 32."""
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
    st_in = cs = None
    eps_in = 0.0
    if ln_in is not None:
        st_in, cs, eps_in = ln_in
        _req(st_in, torch.float32, "ln_in stats"); _req(cs, torch.float32, "ln_in colsum")
        assert st_in.is_contiguous() and st_in.numel() >= (K // stats_chunk(K)) * M * 2 and cs.numel() == N
    if stats_out is not None:
        _req(stats_out, torch.float32, "stats_out")
        assert stats_out.is_contiguous() and stats_out.numel() >= (N // stats_chunk(N)) * M * 2
    _call("lrce_gemm_bf16", _ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(bias), _ptr(residual),
          residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), epilogue, int(out_fp32),
          _ptr(g), _ptr(b), float(eps), _ptr(st_in), stats_chunk(K), _ptr(cs), float(eps_in), _ptr(stats_out), _stream(),
          work=(f"M{M}N{N}K{K}e{epilogue}", 2.0 * M * N * K,
                2.0 * (M * K + N * K + M * N * (2 if residual is not None else 1)) * (2 if out_fp32 else 1)))
    return out


def mlp_fused(x, w1, b1, colsum1, stats_in, eps_in, w2, b2, *, out=None, stats_out=None):
    """out = x + fc2(gelu(fc1(LN(x)))) for C = 128 rows in one kernel (lrce_mlp_fused_bf16); LayerNorm folded as in gemm()"""
    _req(x, torch.bfloat16, "x"); _req(w1, torch.bfloat16, "w1"); _req(w2, torch.bfloat16, "w2")
    for t, nme in ((b1, "b1"), (colsum1, "colsum1"), (stats_in, "stats_in"), (b2, "b2"), (stats_out, "stats_out")):
        _req(t, torch.float32, nme)
    M, C = x.shape
    assert x.stride(1) == 1 and w1.is_contiguous() and w2.is_contiguous() and w1.shape == (4 * C, C) and w2.shape == (C, 4 * C)
    assert stats_in.is_contiguous() and stats_in.numel() >= M * (C // stats_chunk(C)) * 2
    if out is None:
        out = torch.empty_like(x)
    assert out.shape == x.shape and out.stride(1) == 1
    if stats_out is not None:
        assert stats_out.is_contiguous() and stats_out.numel() >= M * (C // stats_chunk(C)) * 2
    _call("lrce_mlp_fused_bf16", _ptr(x), x.stride(0), _ptr(w1), _ptr(b1), _ptr(colsum1), _ptr(stats_in), float(eps_in), _ptr(w2),
          _ptr(b2), _ptr(out), out.stride(0), _ptr(stats_out), M, C, _stream(),
          work=(f"M{M}C{C}", 2.0 * M * C * 4 * C * 2, 2.0 * (2 * M * C + 8 * C * C)))
    return out


_mlp_l2_scratch = {}  # (device, stream, C) -> uint8 scratch of lrce_mlp_l2_scratch_bytes(C), reused by every block launched there


def mlp_l2(x, w1, b1, colsum1, stats_in, eps_in, w2, b2, *, out=None, stats_out=None):
    """out = x + fc2(gelu(fc1(LN(x)))) for C = 256 / 512 rows in one kernel whose hidden rows stay in L2 (lrce_mlp_l2_bf16)"""
    _req(x, torch.bfloat16, "x"); _req(w1, torch.bfloat16, "w1"); _req(w2, torch.bfloat16, "w2")
    for t, nme in ((b1, "b1"), (colsum1, "colsum1"), (stats_in, "stats_in"), (b2, "b2"), (stats_out, "stats_out")):
        _req(t, torch.float32, nme)
    M, C = x.shape
    assert x.stride(1) == 1 and w1.is_contiguous() and w2.is_contiguous() and w1.shape == (4 * C, C) and w2.shape == (C, 4 * C)
    assert stats_in.is_contiguous() and stats_in.numel() >= M * (C // stats_chunk(C)) * 2
    if out is None:
        out = torch.empty_like(x)
    assert out.shape == x.shape and out.stride(1) == 1
    if stats_out is not None:
        assert stats_out.is_contiguous() and stats_out.numel() >= M * (C // stats_chunk(C)) * 2
    key = (x.device.index, _stream(), C)
    scratch = _mlp_l2_scratch.get(key)
    if scratch is None:
        if len(_mlp_l2_scratch) >= 8:  # (device, stream, width) combinations are few; drop the oldest rather than grow
            _mlp_l2_scratch.pop(next(iter(_mlp_l2_scratch)))
        scratch = _mlp_l2_scratch[key] = torch.empty(_lib.lib().lrce_mlp_l2_scratch_bytes(C), device=x.device, dtype=torch.uint8)
    _call("lrce_mlp_l2_bf16", _ptr(x), x.stride(0), _ptr(w1), _ptr(b1), _ptr(colsum1), _ptr(stats_in), stats_chunk(C), float(eps_in),
          _ptr(w2), _ptr(b2), _ptr(out), out.stride(0), _ptr(stats_out), _ptr(scratch), scratch.numel(), M, C, _stream(),
          work=(f"M{M}C{C}", 2.0 * M * C * 4 * C * 2, 2.0 * (2 * M * C + 8 * C * C)))
    return out


def layernorm(x, gamma, beta, eps, *, out=None, out_fp32=False):
    """x bf16 (rows, C) contiguous -> LayerNorm over C."""
    _req(x, torch.bfloat16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    assert x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    _call("lrce_layernorm_bf16", _ptr(x), _ptr(out), _ptr(gamma), _ptr(beta), float(eps), rows, C, int(out_fp32), _stream(),
          work=(f"C{C}", 8.0 * rows * C, rows * C * (6.0 if out_fp32 else 4.0)))
    return out


def patch_merge_ln(x, gamma, beta, eps, n_seg, D, H, W, C):
    """x bf16 [n_seg*D*H*W, C] -> bf16 [n_seg*D*(H/2)*(W/2), 4C] (2x2 gather + LayerNorm)."""
    _req(x, torch.bfloat16, "x")
    assert x.is_contiguous() and x.numel() == n_seg * D * H * W * C
    out = torch.empty((n_seg * D * (H // 2) * (W // 2), 4 * C), device=x.device, dtype=torch.bfloat16)
    _call("lrce_patch_merge_ln_bf16", _ptr(x), _ptr(out), _ptr(gamma), _ptr(beta), float(eps), n_seg, D, H, W, C, _stream(),
          work=(f"C{C}", 8.0 * x.numel(), 4.0 * x.numel()))
    return out


def patch_gather(clips):
    """clips (n_seg, T, 3, Hin, Win), fp32 in [0,1] or the uint8 frames themselves (x / 255 is then applied in the kernel)
    -> bf16 [n_seg*ceil(T/2)*(Hin/4)*(Win/4), 96]."""
    if clips.dtype not in (torch.float32, torch.uint8):
        raise _lib.LrceError(f"clips must be fp32 or uint8, got {clips.dtype}")
    _req(clips, clips.dtype, "clips")
    assert clips.is_contiguous() and clips.dim() == 5 and clips.shape[2] == 3
    n, T, _, Hin, Win = clips.shape
    out = torch.empty((n * ((T + 1) // 2) * (Hin // 4) * (Win // 4), 96), device=clips.device, dtype=torch.bfloat16)
    sym = "lrce_patch_gather_f32" if clips.dtype == torch.float32 else "lrce_patch_gather_u8"
    _call(sym, _ptr(clips), _ptr(out), n, T, Hin, Win, _stream(),
          work=("", 2.0 * clips.numel(), clips.element_size() * clips.numel() + 2.0 * out.numel()))
    return out


def window_remap(x, n_seg, dims, window, shift, inverse=False):
    """x bf16 [n_seg*D*H*W, C] -> same shape, rows permuted by shift+partition (or its inverse)."""
    _req(x, torch.bfloat16, "x")
    assert x.is_contiguous()
    out = torch.empty_like(x)
    _call("lrce_window_remap_bf16", _ptr(x), _ptr(out), n_seg, *dims, x.shape[-1], *window, *shift, int(inverse), _stream(),
          work=(f"C{x.shape[-1]}", 0.0, 4.0 * x.numel()))
    return out


def remap_index(dims, window, shift, device="cuda"):
    """(gather [nWin, N], region [nWin, N], relpos [N]) int32 tables from the kernels' own index functions."""
    n = dims[0] * dims[1] * dims[2]
    N = window[0] * window[1] * window[2]
    gather = torch.empty((n // N, N), device=device, dtype=torch.int32)
    region = torch.empty((n // N, N), device=device, dtype=torch.int32)
    relpos = torch.empty((N,), device=device, dtype=torch.int32)
    _call("lrce_remap_index", _ptr(gather), _ptr(region), _ptr(relpos), *dims, *window, *shift, _stream())
    return gather, region, relpos


def window_bias_pack(table):
    """relative_position_bias_table fp32 [2535, nH] -> dense bf16 [nH, 160, 160] (pre-multiplied by log2 e; rows and columns
    in the kernel's class-grouped slot order, pad columns -inf; rows 155-156 hold the fp32 row maxima)."""
    _req(table, torch.float32, "table")
    assert table.is_contiguous() and table.shape[0] == 2535
    nh = table.shape[1]
    out = torch.empty((nh, BIAS_PITCH, BIAS_PITCH), device=table.device, dtype=torch.bfloat16)
    _call("lrce_window_bias_pack", _ptr(table), _ptr(out), nh, _stream())
    return out


def window_attention(qkv, bias_dense, n_seg, D, H, W, C, n_heads, shift_hw, out=None, prof=None):
    """qkv bf16 [n_seg*D*H*W, 3C] (natural order) -> bf16 [n_seg*D*H*W, C] (natural order). `prof` (tools only): int64 [224]
    device buffer -> the instrumented instantiation (lrce_window_attention_profile)."""
    _req(qkv, torch.bfloat16, "qkv"); _req(bias_dense, torch.bfloat16, "bias_dense")
    assert qkv.is_contiguous() and qkv.shape == (n_seg * D * H * W, 3 * C)
    if out is None:
        out = torch.empty((qkv.shape[0], C), device=qkv.device, dtype=torch.bfloat16)
    if prof is not None:
        assert prof.dtype == torch.int64 and prof.is_cuda and prof.numel() >= 224
        _call("lrce_window_attention_profile", _ptr(qkv), _ptr(out), _ptr(bias_dense), n_seg, D, H, W, C, n_heads,
              shift_hw[0], shift_hw[1], _stream(), _ptr(prof))
        return out
    _call("lrce_window_attention_bf16", _ptr(qkv), _ptr(out), _ptr(bias_dense), n_seg, D, H, W, C, n_heads,
          shift_hw[0], shift_hw[1], _stream(),
          # core FLOPs as SURVEY.md 8(d) counts them: QK^T + PV = 4 * 147^2 * 32 per (window, head)
          work=(f"C{C}", 4.0 * 147 * 147 * 32 * (qkv.shape[0] // 147) * n_heads, 2.0 * qkv.numel() + 2.0 * out.numel()))
    return out


def video_posembed_ln(proj, emb_cls, emb_pos, emb_len, emb_clip, gamma, beta, eps, B, S, T, P, out=None):
    _req(proj, torch.bfloat16, "proj")
    assert proj.is_contiguous() and proj.shape == (B * S * T * P, 768)
    if out is None:
        out = torch.empty((B, S, T * (P + 1), 768), device=proj.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out")
    assert out.is_contiguous() and out.numel() == B * S * T * (P + 1) * 768
    _call("lrce_video_posembed_ln", _ptr(proj), _ptr(emb_cls), _ptr(emb_pos), _ptr(emb_len), _ptr(emb_clip), _ptr(gamma),
          _ptr(beta), float(eps), _ptr(out), B, S, T, P, _stream())
    return out


def text_posembed_ln(text, emb_cls, emb_pos, gamma, beta, eps, out=None):
    if text.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.LrceError(f"text features must be bf16 or fp32, got {text.dtype}")
    _req(text, text.dtype, "text")
    assert text.is_contiguous() and text.dim() == 3 and text.shape[2] == 768
    Bt, L, _ = text.shape
    if out is None:
        out = torch.empty((Bt, L + 1, 768), device=text.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out")
    assert out.is_contiguous() and out.numel() == Bt * (L + 1) * 768
    _call("lrce_text_posembed_ln", _ptr(text), int(text.dtype == torch.float32), _ptr(emb_cls), _ptr(emb_pos), _ptr(gamma), _ptr(beta), float(eps), _ptr(out),
          Bt, L, _stream())
    return out


def encoder_walk_pack(layer_table, n_layers, fc_w, fc_b, n_out):
    """decoder weights (fp32 masters, 16 device pointers per layer) -> the walk's per-CTA streaming order (lrce_encoder_walk_pack)"""
    _req(fc_w, torch.float32, "fc_w"); _req(fc_b, torch.float32, "fc_b")
    assert layer_table.dtype == torch.int64 and layer_table.is_cuda and layer_table.shape == (n_layers, 16)
    assert fc_w.is_contiguous() and fc_w.shape == (n_out, 768) and fc_b.shape == (n_out,)
    n = _lib.lib().lrce_encoder_walk_pack_bytes(int(n_layers), int(n_out))
    packed = torch.empty(n, device=fc_w.device, dtype=torch.uint8)
    assert packed.data_ptr() % 128 == 0
    _call("lrce_encoder_walk_pack", _ptr(layer_table), n_layers, _ptr(fc_w), _ptr(fc_b), n_out, _ptr(packed), _stream())
    return packed


def encoder_walk(packed, n_layers, kv_video, kv_text, tok0, f_g, f_b, eps, n_out, act, out, rows, S, Tv, Lt, n_cand,
                 tokens_tap=None):
    """the whole summarisation-token walk + answer head as one kernel of row-sharded 16-CTA clusters (lrce_encoder_walk)"""
    _req(kv_video, torch.bfloat16, "kv_video"); _req(kv_text, torch.bfloat16, "kv_text"); _req(packed, torch.uint8, "packed")
    _req(tok0, torch.float32, "tok0"); _req(out, torch.float32, "out"); _req(tokens_tap, torch.float32, "tokens_tap")
    assert kv_video.stride(0) == kv_text.stride(0) and out.is_contiguous() and out.shape == (rows, n_out)
    assert packed.numel() == _lib.lib().lrce_encoder_walk_pack_bytes(int(n_layers), int(n_out))
    if tokens_tap is not None:
        assert tokens_tap.is_contiguous() and tokens_tap.shape == (S, rows, 768)
    _call("lrce_encoder_walk", _ptr(packed), n_layers, _ptr(kv_video), _ptr(kv_text), kv_video.stride(0), _ptr(tok0),
          _ptr(f_g), _ptr(f_b), float(eps), n_out, act, _ptr(out), _ptr(tokens_tap), rows, S, Tv, Lt, n_cand, _stream(),
          work=(f"R{rows}S{S}", 2.0 * rows * S * n_layers * (4 * 768 * 768 + 2 * 768 * 3072 + 2 * (Tv + Lt) * 768),
                2.0 * S * n_layers * (3 * 768 * 768 + 2 * 768 * 3072) + 2.0 * rows * S * n_layers * (Tv + Lt) * 1536 * 2))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# row / sequence kernels (csrc/seqops.cu): BERT-base and the encoder's training step
# ----------------------------------------------------------------------------------------------------------------------
ROWS_CAST, ROWS_GELU_FWD, ROWS_GELU_BWD = 0, 1, 2
XATTN_MAXK = 256


def _seed_args(seed):
    """(host seed, device seed pointer): `seed` is an int, or a (int, int64 device tensor) pair whose tensor is XOR-ed in at run
    time (fresh dropout masks for launches replayed from a CUDA graph)"""
    if isinstance(seed, tuple):
        return int(seed[0]), seed[1].data_ptr()
    return int(seed), 0


def _t(tT):
    """(pointer, ld, row offset) of an optional transposed-copy target given as (tensor [cols, ld], row offset)"""
    if tT is None:
        return 0, 0, 0
    t, r0 = tT
    _req(t, torch.bfloat16, "transposed copy")
    assert t.dim() == 2 and t.stride(1) == 1
    return t.data_ptr(), t.stride(0), int(r0)


def add_ln(a, res, gamma, beta, eps, *, res_bcast=False, u_out=None, y_f32=None, y_bf16=None, yT=None, p_a=0.0, site_a=0,
           p_out=0.0, site_out=0, seed=0):
    """y = drop_out(LN(res + drop_a(a))), fp32 rows of 768 (lrce_add_ln_768)"""
    _req(a, torch.float32, "a"); _req(res, torch.float32, "res"); _req(u_out, torch.float32, "u_out")
    _req(y_f32, torch.float32, "y_f32"); _req(y_bf16, torch.bfloat16, "y_bf16")
    n = (a if a is not None else (y_f32 if y_f32 is not None else y_bf16)).numel() // 768
    tp, ld, r0 = _t(yT)
    _call("lrce_add_ln_768", _ptr(a), _ptr(res), int(res_bcast), _ptr(gamma), _ptr(beta), float(eps), _ptr(u_out), _ptr(y_f32),
          _ptr(y_bf16), tp, ld, r0, n, float(p_a), int(site_a), float(p_out), int(site_out), *_seed_args(seed), _stream())


def ln_bwd(dy_a, dy_b, u, gamma, eps, dgamma, dbeta, *, du=None, dub=None, dubT=None, p_a=0.0, site_a=0, p_out=0.0, site_out=0,
           seed=0):
    _req(dy_a, torch.float32, "dy_a"); _req(dy_b, torch.float32, "dy_b"); _req(u, torch.float32, "u")
    _req(du, torch.float32, "du"); _req(dub, torch.bfloat16, "dub"); _req(dgamma, torch.float32, "dgamma")
    n = u.numel() // 768
    tp, ld, r0 = _t(dubT)
    _call("lrce_ln_bwd_768", _ptr(dy_a), _ptr(dy_b), _ptr(u), _ptr(gamma), float(eps), _ptr(du), _ptr(dub), tp, ld, r0,
          _ptr(dgamma), _ptr(dbeta), n, float(p_a), int(site_a), float(p_out), int(site_out), *_seed_args(seed), _stream())


def rows_to_bf16(x, *, aux=None, y=None, yT=None, mode=ROWS_CAST, group=1, p=0.0, site=0, seed=0):
    """fp32 [n, C] -> bf16 (+ transposed copy) with optional GELU forward / backward and dropout (lrce_rows_f32_to_bf16)"""
    _req(x, torch.float32, "x"); _req(aux, torch.float32, "aux"); _req(y, torch.bfloat16, "y")
    assert x.is_contiguous() and x.dim() == 2
    n, C = x.shape
    tp, ld, r0 = _t(yT)
    _call("lrce_rows_f32_to_bf16", _ptr(x), _ptr(aux), _ptr(y), tp, ld, r0, n, C, mode, group, float(p), int(site), *_seed_args(seed),
          _stream())
    return y


def dropout_bf16_(x, p, site, seed):
    _req(x, torch.bfloat16, "x")
    assert x.is_contiguous()
    if p > 0:
        _call("lrce_dropout_bf16", _ptr(x), x.numel(), float(p), int(site), *_seed_args(seed), _stream())
    return x


def xattn_fwd(q, kv_video, kv_text, kcol, R, S, seg, Tv, Lt, n_cand, P, ctx, ctxT=None, p=0.0, site=0, seed=0):
    _req(q, torch.float32, "q"); _req(kv_video, torch.bfloat16, "kv_video"); _req(kv_text, torch.bfloat16, "kv_text")
    _req(P, torch.float32, "P"); _req(ctx, torch.bfloat16, "ctx")
    assert kv_video.stride(0) == kv_text.stride(0) and P.numel() >= R * 12 * XATTN_MAXK
    tp, ld, r0 = _t(ctxT)
    _call("lrce_xattn_fwd", _ptr(q), _ptr(kv_video), _ptr(kv_text), kv_video.stride(0), kcol, R, S, seg, Tv, Lt, n_cand, _ptr(P),
          _ptr(ctx), tp, ld, r0, float(p), int(site), *_seed_args(seed), _stream())


def xattn_bwd(q, kv_video, kv_text, kcol, R, S, seg, Tv, Lt, n_cand, P, dctx, dq, dkv_video, dkv_text, dqT=None, p=0.0, site=0,
              seed=0):
    _req(q, torch.float32, "q"); _req(dctx, torch.float32, "dctx"); _req(dq, torch.bfloat16, "dq")
    _req(dkv_video, torch.bfloat16, "dkv_video"); _req(dkv_text, torch.bfloat16, "dkv_text")
    assert dkv_video.stride(0) == kv_video.stride(0) == dkv_text.stride(0) == kv_text.stride(0)
    tp, ld, r0 = _t(dqT)
    _call("lrce_xattn_bwd", _ptr(q), _ptr(kv_video), _ptr(kv_text), kv_video.stride(0), kcol, R, S, seg, Tv, Lt, n_cand, _ptr(P),
          _ptr(dctx), _ptr(dq), tp, ld, r0, _ptr(dkv_video), _ptr(dkv_text), float(p), int(site), *_seed_args(seed), _stream())


def rowsum_bf16(src, cols, out, accumulate=False):
    _req(src, torch.bfloat16, "src"); _req(out, torch.float32, "out")
    assert src.dim() == 2 and src.stride(1) == 1 and out.numel() == src.shape[0]
    _call("lrce_rowsum_bf16", _ptr(src), src.stride(0), cols, _ptr(out), src.shape[0], int(accumulate), _stream())


def colsum(src, out):
    """out[c] += sum_r src[r, c] (out fp32, initialised by the caller)"""
    assert src.dtype in (torch.float32, torch.bfloat16) and src.dim() == 2 and src.stride(1) == 1
    _req(src, src.dtype, "src"); _req(out, torch.float32, "out")
    _call("lrce_colsum", _ptr(src), int(src.dtype == torch.bfloat16), src.shape[0], src.shape[1], src.stride(0), _ptr(out), _stream())


def transpose_bf16(src, dst=None):
    _req(src, torch.bfloat16, "src")
    assert src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    if dst is None:
        dst = torch.empty((cols, (rows + 7) // 8 * 8), device=src.device, dtype=torch.bfloat16)
        if dst.shape[1] != rows:
            dst[:, rows:].zero_()
    _call("lrce_transpose_bf16", _ptr(src), rows, cols, src.stride(0), _ptr(dst), dst.stride(0), _stream())
    return dst


def add_bf16(dst, a, b, c=None):
    _call("lrce_add_bf16", _ptr(dst), _ptr(a), _ptr(b), _ptr(c), dst.numel(), _stream())
    return dst


def posembed_bwd(dy, *, proj=None, text=None, emb_cls, emb_pos, emb_len=None, emb_clip=None, gamma, eps, dproj=None, d_cls, d_pos,
                 d_len=None, d_clip=None, dgamma, dbeta, B, S, T, P, is_text, p=0.0, site=0, seed=0):
    _req(dy, torch.float32, "dy")
    _call("lrce_posembed_bwd", _ptr(dy), _ptr(proj), _ptr(text), int(text is not None and text.dtype == torch.float32),
          _ptr(emb_cls), _ptr(emb_pos), _ptr(emb_len), _ptr(emb_clip), _ptr(gamma), float(eps), _ptr(dproj), _ptr(d_cls),
          _ptr(d_pos), _ptr(d_len), _ptr(d_clip), _ptr(dgamma), _ptr(dbeta), B, S, T, P, int(is_text), float(p), int(site),
          *_seed_args(seed), _stream())


def bert_embed_ln(ids, type_ids, word, pos, type_emb, gamma, beta, eps):
    assert ids.dtype == torch.int64 and ids.is_cuda and ids.is_contiguous() and ids.dim() == 2
    assert type_ids is None or (type_ids.dtype == torch.int64 and type_ids.is_contiguous() and type_ids.shape == ids.shape)
    _req(word, torch.float32, "word embeddings"); _req(pos, torch.float32, "position embeddings")
    n, L = ids.numel(), ids.shape[1]
    out_f32 = torch.empty((n, 768), device=ids.device, dtype=torch.float32)
    out_bf16 = torch.empty((n, 768), device=ids.device, dtype=torch.bfloat16)
    _call("lrce_bert_embed_ln", _ptr(ids), _ptr(type_ids), _ptr(word), _ptr(pos), _ptr(type_emb), _ptr(gamma), _ptr(beta),
          float(eps), _ptr(out_f32), _ptr(out_bf16), n, L, word.shape[0], type_emb.shape[0], pos.shape[0], _stream())
    return out_f32, out_bf16


def bert_attention(qkv, mask, n_seq, L, n_heads=12):
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.is_contiguous() and qkv.shape == (n_seq * L, 3 * 64 * n_heads)
    assert mask is None or (mask.dtype == torch.int64 and mask.is_contiguous() and mask.shape == (n_seq, L))
    out = torch.empty((n_seq * L, 64 * n_heads), device=qkv.device, dtype=torch.bfloat16)
    _call("lrce_bert_attention", _ptr(qkv), _ptr(mask), _ptr(out), n_seq, L, n_heads, _stream(),
          work=(f"L{L}", 4.0 * n_seq * n_heads * L * L * 64, 2.0 * qkv.numel() + 2.0 * out.numel()))
    return out


def gemm_skinny(a, w, bias, out):
    """out fp32 [M, N] = a bf16 [M, K] @ w bf16 [N, K].T (+ bias), K in {768, 3072}, few rows (lrce_gemm_skinny_bf16)"""
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w"); _req(bias, torch.float32, "bias"); _req(out, torch.float32, "out")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] >= K and out.shape == (M, N)
    _call("lrce_gemm_skinny_bf16", _ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(bias), _ptr(out), out.stride(0),
          _stream(), work=(f"skinny M{M}N{N}K{K}", 2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N))
    return out
