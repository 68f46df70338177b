"""Training step of the recurrent cross-modal encoder on liblrce_b200 (BASELINE.json configs[4]: cross-modal encoder forward +
backward; the reference runs it as `scaler.scale(loss).backward()` over fusionv3.py:41-51 / :168-198 under DDP,
agent_oe.py:28-42, agent_base.py:75-76).

`encoder_train(module, video_features, text_features, n_cand)` is a `torch.autograd.Function` over the module's own
parameters: autograd sees ONE node whose backward hands every encoder parameter its gradient, so the reference agent's
GradScaler / AdamW / plain `DDP(model)` work unchanged. Both passes are sequences of hand-written kernels:

  forward   projection GEMM -> pos-embed kernels (+ dropout) -> K/V GEMM of all 12 layers -> for every recurrent step and
            layer: v-proj / out-proj GEMMs of the length-1 self-attention, LayerNorm steps (lrce_add_ln_768), q GEMM,
            single-query cross attention (lrce_xattn_fwd), FFN GEMMs around a GELU row pass -> answer-head GEMM.
            Activations the backward needs are kept: pre-LayerNorm sums, q, attention probabilities, FFN pre-activations,
            and — TRANSPOSED and stacked over the recurrent steps — every Linear's input.
  backward  the mirror image: LayerNorm / GELU / cross-attention backward kernels produce the bf16 operands of
            dX = dY W (one lrce_gemm_bf16 per Linear and step) and, transposed and stacked over the steps, of
            dW = dY^T X (one lrce_gemm_bf16 per weight: the sum over recurrent steps and rows is the GEMM's K loop).
            Layer n's weight gradients are complete when the s = 0 pass leaves layer n; with `grad_sync="overlap"` their
            NCCL all-reduce is issued right there and runs under the backward of layers n-1 .. 0 and of the memory path.

Dropout (train mode) is a counter-based hash keyed by (seed, site, index), regenerated in the backward pass
(csrc/seqops.cu); it cannot reproduce torch's RNG stream, so gradient parity is checked with drop_out_rate = 0.
The extractors feeding the encoder are forward-only (frozen by E2EBase): no gradient is returned for the features.
"""
import torch
import torch.distributed as dist

from . import ops

EPS = 1e-12
D = 768
FF = 3072
_counter = [0]


def _round8(n):
    """leading dimension of a stacked / transposed operand: a multiple of 8 elements (16-byte TMA pitch), at least one
    64-element k-block wide; the pad columns are zero"""
    return max(64, (n + 7) // 8 * 8)


def _transposed(w):
    """[N, K] bf16 -> [K, _round8(N)] (zero pad columns)"""
    dst = torch.zeros((w.shape[1], _round8(w.shape[0])), device=w.device, dtype=torch.bfloat16)
    return ops.transpose_bf16(w, dst)


def _seed():
    _counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _counter[0] * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


class _Sites:
    VIDEO, TEXT = 1, 2

    @staticmethod
    def outer(s):
        return 8 + s

    @staticmethod
    def layer(s, n, n_layers, k):  # k: 0 self-attn head, 1 drop1, 2 attention probs, 3 drop2, 4 FFN inner, 5 drop3
        return 64 + 8 * (s * n_layers + n) + k


def pack_train(module):
    """bf16 operands of the training step: every Linear weight as stored ([N, K]: forward and dW GEMMs) and transposed
    ([K, N]: the dX GEMMs). Rebuilt whenever a parameter changed (every optimizer step)."""
    ft = module.fusion_transformer
    dev = module.final_fc.weight.device
    bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
    tr = _transposed
    pk = {"layers": []}
    kv_w, kv_b = [], []
    for lyr in ft.transformer.layers:
        sa, ca = lyr.self_attn, lyr.multihead_attn
        wv, wso = bf(sa.in_proj_weight[2 * D:]), bf(sa.out_proj.weight)
        wq, wo = bf(ca.in_proj_weight[:D]), bf(ca.out_proj.weight)
        w1, w2 = bf(lyr.linear1.weight), bf(lyr.linear2.weight)
        pk["layers"].append(dict(
            wv=wv, bv=sa.in_proj_bias.detach()[2 * D:].contiguous(), wso=wso, bso=sa.out_proj.bias.detach(),
            wq=wq, bq=ca.in_proj_bias.detach()[:D].contiguous(), wo=wo, bo=ca.out_proj.bias.detach(),
            w1=w1, b1=lyr.linear1.bias.detach(), w2=w2, b2=lyr.linear2.bias.detach(),
            wvT=tr(wv), wsoT=tr(wso), wqT=tr(wq), woT=tr(wo), w1T=tr(w1), w2T=tr(w2),
            n1=(lyr.norm1.weight.detach(), lyr.norm1.bias.detach()), n2=(lyr.norm2.weight.detach(), lyr.norm2.bias.detach()),
            n3=(lyr.norm3.weight.detach(), lyr.norm3.bias.detach())))
        kv_w.append(ca.in_proj_weight.detach()[D:])
        kv_b.append(ca.in_proj_bias.detach()[D:])
    pk["kv_w"] = bf(torch.cat(kv_w))
    pk["kv_b"] = torch.cat(kv_b).float().contiguous()
    pk["kv_wT"] = tr(pk["kv_w"])  # [768, 18432]
    if hasattr(module, "projection_layer"):
        pk["proj_w"], pk["proj_b"] = bf(module.projection_layer.weight), module.projection_layer.bias.detach()
    n_out = module.final_fc.out_features
    n_pad = (n_out + 127) // 128 * 128
    fc_w = torch.zeros((n_pad, D), device=dev, dtype=torch.bfloat16)
    fc_w[:n_out] = module.final_fc.weight.detach()
    fc_b = torch.zeros(n_pad, device=dev, dtype=torch.float32)
    fc_b[:n_out] = module.final_fc.bias.detach()
    pk["fc_w"], pk["fc_b"], pk["fc_wT"] = fc_w, fc_b, tr(fc_w[:n_out].contiguous())  # [768, round8(n_out)]
    return pk


class _EncoderTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, video_features, text_features, n_cand, names, *params):
        m = module
        pk = m._train_pack()
        ve, te, ft = m.video_pos_embed, m.question_pos_embed, m.fusion_transformer
        f32 = lambda t: t.detach().float().contiguous()
        B, S, T, P, Dv = video_features.shape
        R, L, _ = text_features.shape
        dev = video_features.device
        NL = len(pk["layers"])
        Tv, Lt = T * (P + 1), L + 1
        p = float(m.video_dropout.p) if m.training else 0.0
        seed = _seed()
        ld = _round8(S * R)
        st = dict(B=B, S=S, T=T, P=P, R=R, L=L, Tv=Tv, Lt=Lt, NL=NL, p=p, seed=seed, ld=ld, n_cand=n_cand, pk=pk, names=names)

        vf = video_features.reshape(B * S * T * P, Dv).contiguous()
        text = text_features.contiguous()
        proj = ops.gemm(vf, pk["proj_w"], pk["proj_b"]) if "proj_w" in pk else vf
        emb = dict(v_cls=f32(ve.emb_cls.reshape(D)), v_pos=f32(ve.emb_pos.reshape(-1, D)), v_len=f32(ve.emb_len.reshape(-1, D)),
                   v_clip=f32(ve.emb_clip.reshape(-1, D)), v_g=f32(ve.layer_norm.weight), v_b=f32(ve.layer_norm.bias),
                   t_cls=f32(te.emb_cls.reshape(D)), t_pos=f32(te.emb_pos.reshape(-1, D)), t_g=f32(te.layer_norm.weight),
                   t_b=f32(te.layer_norm.bias), f_g=f32(ft.fusion_layer_norm.weight), f_b=f32(ft.fusion_layer_norm.bias))
        vemb = ops.video_posembed_ln(proj, emb["v_cls"], emb["v_pos"], emb["v_len"], emb["v_clip"], emb["v_g"], emb["v_b"], EPS,
                                     B, S, T, P).view(B * S * Tv, D)
        temb = ops.text_posembed_ln(text, emb["t_cls"], emb["t_pos"], emb["t_g"], emb["t_b"], EPS).view(R * Lt, D)
        ops.dropout_bf16_(vemb, p, _Sites.VIDEO, seed)
        ops.dropout_bf16_(temb, p, _Sites.TEXT, seed)
        kv_video = ops.gemm(vemb, pk["kv_w"], pk["kv_b"])
        kv_text = ops.gemm(temb, pk["kv_w"], pk["kv_b"])

        z = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)
        zeros_bf = lambda *s: torch.zeros(s, device=dev, dtype=torch.bfloat16)
        U = z(S, NL, 3, R, D)
        Q = z(S, NL, R, D)
        Pst = z(S, NL, R * 12, ops.XATTN_MAXK)
        F = z(S, NL, R, FF)
        UF = z(S, R, D)
        # transposed, step-stacked inputs of every Linear of a layer: [features, ld], column s * R + row
        XT = [dict(x=zeros_bf(D, ld), v=zeros_bf(D, ld), h1=zeros_bf(D, ld), ctx=zeros_bf(D, ld), h2=zeros_bf(D, ld),
                   g=zeros_bf(FF, ld)) for _ in range(NL)]
        ldR = _round8(R)
        tokT = zeros_bf(D, ldR)  # final token, transposed: the answer head's dW operand

        tok0 = f32(ft.summarization_token.reshape(1, D))
        x32, xb = z(R, D), z(R, D, dt=torch.bfloat16)
        h1_32, h1b = z(R, D), z(R, D, dt=torch.bfloat16)
        h2_32, h2b = z(R, D), z(R, D, dt=torch.bfloat16)
        vb, ctxb, gb = z(R, D, dt=torch.bfloat16), z(R, D, dt=torch.bfloat16), z(R, FF, dt=torch.bfloat16)
        tok32 = tok0.expand(R, D).contiguous()
        for s in range(S):
            if s == 0:
                x32.copy_(tok32)
                ops.rows_to_bf16(x32, y=xb, yT=(XT[0]["x"], 0))
            for n, lw in enumerate(pk["layers"]):
                site = lambda k: _Sites.layer(s, n, NL, k)
                v32 = ops.gemm(xb, lw["wv"], lw["bv"], out_fp32=True)
                ops.rows_to_bf16(v32, y=vb, yT=(XT[n]["v"], s * R), group=64, p=p, site=site(0), seed=seed)
                sa32 = ops.gemm(vb, lw["wso"], lw["bso"], out_fp32=True)
                ops.add_ln(sa32, x32, lw["n1"][0], lw["n1"][1], EPS, u_out=U[s, n, 0], y_f32=h1_32, y_bf16=h1b,
                           yT=(XT[n]["h1"], s * R), p_a=p, site_a=site(1), seed=seed)
                ops.gemm(h1b, lw["wq"], lw["bq"], out_fp32=True, out=Q[s, n])
                ops.xattn_fwd(Q[s, n], kv_video, kv_text, n * 2 * D, R, S, s, Tv, Lt, n_cand, Pst[s, n], ctxb,
                              ctxT=(XT[n]["ctx"], s * R), p=p, site=site(2), seed=seed)
                o32 = ops.gemm(ctxb, lw["wo"], lw["bo"], out_fp32=True)
                ops.add_ln(o32, h1_32, lw["n2"][0], lw["n2"][1], EPS, u_out=U[s, n, 1], y_f32=h2_32, y_bf16=h2b,
                           yT=(XT[n]["h2"], s * R), p_a=p, site_a=site(3), seed=seed)
                ops.gemm(h2b, lw["w1"], lw["b1"], out_fp32=True, out=F[s, n])
                ops.rows_to_bf16(F[s, n], y=gb, yT=(XT[n]["g"], s * R), mode=ops.ROWS_GELU_FWD, p=p, site=site(4), seed=seed)
                y32 = ops.gemm(gb, lw["w2"], lw["b2"], out_fp32=True)
                nxt = (XT[n + 1]["x"], s * R) if n + 1 < NL else None
                ops.add_ln(y32, h2_32, lw["n3"][0], lw["n3"][1], EPS, u_out=U[s, n, 2], y_f32=x32, y_bf16=xb, yT=nxt,
                           p_a=p, site_a=site(5), seed=seed)
            # tok_{s+1} = dropout(LN_f(tok_s + x_12))   (fusionv3.py:47-49)
            new_tok = z(R, D)
            yT = (XT[0]["x"], (s + 1) * R) if s + 1 < S else (tokT, 0)
            ops.add_ln(x32, tok32, emb["f_g"], emb["f_b"], EPS, u_out=UF[s], y_f32=new_tok, y_bf16=xb, yT=yT,
                       p_out=p, site_out=_Sites.outer(s), seed=seed)
            tok32 = new_tok
            x32.copy_(tok32)
        n_out = m.final_fc.out_features
        logits = ops.gemm(xb, pk["fc_w"], pk["fc_b"], out_fp32=True)[:, :n_out].contiguous()
        st.update(vf=vf, text=text, proj=proj, vemb=vemb, temb=temb, kv_video=kv_video, kv_text=kv_text, U=U, Q=Q, Pst=Pst, F=F,
                  UF=UF, XT=XT, tokT=tokT, emb=emb, n_out=n_out, ldR=ldR)
        ctx.st = st
        ctx.module = m
        ctx.params = params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        st, m = ctx.st, ctx.module
        pk, emb, XT = st["pk"], st["emb"], st["XT"]
        B, S, T, P, R, L, Tv, Lt, NL, p, seed, ld = (st[k] for k in ("B", "S", "T", "P", "R", "L", "Tv", "Lt", "NL", "p", "seed", "ld"))
        n_cand, n_out, ldR = st["n_cand"], st["n_out"], st["ldR"]
        U, Q, Pst, F, UF = st["U"], st["Q"], st["Pst"], st["F"], st["UF"]
        kv_video, kv_text = st["kv_video"], st["kv_text"]
        dev = dlogits.device
        z = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)
        zeros_bf = lambda *s: torch.zeros(s, device=dev, dtype=torch.bfloat16)

        # ---- one flat fp32 gradient buffer in parameters() order: a layer's 18 tensors are contiguous (all-reduce slices)
        names, params = st["names"], ctx.params
        sizes = [q.numel() for q in params]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        G, off, offsets = {}, 0, {}
        for name, q, n_el in zip(names, params, sizes):
            G[name] = flat[off:off + n_el].view(q.shape)
            offsets[name] = (off, off + n_el)
            off += n_el
        sync = getattr(m, "grad_sync", "none") == "overlap" and dist.is_available() and dist.is_initialized() \
            and dist.get_world_size() > 1
        works = []

        def reduce_span(lo, hi):
            if sync and hi > lo:
                works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

        LP = "fusion_transformer.transformer.layers.%d."

        # ---- answer head: dtok = dlogits W_fc ; dW_fc = dlogits^T tok ; db_fc
        Kp = pk["fc_wT"].shape[1]
        dl = torch.zeros((R, Kp), device=dev, dtype=torch.float32)
        dl[:, :n_out] = dlogits
        dlb, dlT = z(R, Kp, dt=torch.bfloat16), zeros_bf(Kp, ldR)
        ops.rows_to_bf16(dl, y=dlb, yT=(dlT, 0))
        dtok = ops.gemm(dlb, pk["fc_wT"], None, out_fp32=True)
        ops.gemm(dlT[:n_out], st["tokT"], None, out_fp32=True, out=G["final_fc.weight"])
        ops.rowsum_bf16(dlT[:n_out], R, G["final_fc.bias"])

        # ---- memory path operands for the per-layer K/V weight gradient: X_mem^T = [vemb^T | temb^T x S]
        Nv, Nt = st["vemb"].shape[0], st["temb"].shape[0]
        Ktot = _round8(Nv + S * Nt)
        XmT = zeros_bf(D, Ktot)
        ops.transpose_bf16(st["vemb"], XmT[:, :Nv])
        ops.transpose_bf16(st["temb"], XmT[:, Nv:Nv + Nt])
        for s in range(1, S):
            XmT[:, Nv + s * Nt:Nv + (s + 1) * Nt].copy_(XmT[:, Nv:Nv + Nt])
        dkv_video = (torch.zeros_like(kv_video) if n_cand > 1 else torch.empty_like(kv_video))
        dkv_text = [torch.empty_like(kv_text) for _ in range(S)]
        dKVT = zeros_bf(2 * D, Ktot)  # one layer's [dK | dV]^T, reused layer after layer (stream-ordered)

        DYT = [dict(y=zeros_bf(D, ld), f=zeros_bf(FF, ld), o=zeros_bf(D, ld), q=zeros_bf(D, ld), sa=zeros_bf(D, ld), v=zeros_bf(D, ld))
               for _ in range(NL)]
        dyb, dfb, dob, dqb, dsab, dvb = (z(R, D, dt=torch.bfloat16), z(R, FF, dt=torch.bfloat16), z(R, D, dt=torch.bfloat16),
                                         z(R, D, dt=torch.bfloat16), z(R, D, dt=torch.bfloat16), z(R, D, dt=torch.bfloat16))
        du3, du2, du1, du_f = z(R, D), z(R, D), z(R, D), z(R, D)
        d_tok0 = torch.zeros(D, device=dev, dtype=torch.float32)
        dtok32 = dtok
        for s in reversed(range(S)):
            ops.ln_bwd(dtok32, None, UF[s], emb["f_g"], EPS, G["fusion_transformer.fusion_layer_norm.weight"],
                       G["fusion_transformer.fusion_layer_norm.bias"], du=du_f, p_out=p, site_out=_Sites.outer(s), seed=seed)
            dx_a, dx_b = du_f, None
            for n in reversed(range(NL)):
                lw, lp = pk["layers"][n], LP % n
                site = lambda k: _Sites.layer(s, n, NL, k)
                # LN3 / FFN
                ops.ln_bwd(dx_a, dx_b, U[s, n, 2], lw["n3"][0], EPS, G[lp + "norm3.weight"], G[lp + "norm3.bias"], du=du3, dub=dyb,
                           dubT=(DYT[n]["y"], s * R), p_a=p, site_a=site(5), seed=seed)
                dg32 = ops.gemm(dyb, lw["w2T"], None, out_fp32=True)
                ops.rows_to_bf16(dg32, aux=F[s, n], y=dfb, yT=(DYT[n]["f"], s * R), mode=ops.ROWS_GELU_BWD, p=p, site=site(4), seed=seed)
                dh2 = ops.gemm(dfb, lw["w1T"], None, out_fp32=True)
                # LN2 / cross attention
                ops.ln_bwd(du3, dh2, U[s, n, 1], lw["n2"][0], EPS, G[lp + "norm2.weight"], G[lp + "norm2.bias"], du=du2, dub=dob,
                           dubT=(DYT[n]["o"], s * R), p_a=p, site_a=site(3), seed=seed)
                dctx = ops.gemm(dob, lw["woT"], None, out_fp32=True)
                ops.xattn_bwd(Q[s, n], kv_video, kv_text, n * 2 * D, R, S, s, Tv, Lt, n_cand, Pst[s, n], dctx, dqb, dkv_video,
                              dkv_text[s], dqT=(DYT[n]["q"], s * R), p=p, site=site(2), seed=seed)
                dh1 = ops.gemm(dqb, lw["wqT"], None, out_fp32=True)
                # LN1 / length-1 self-attention
                ops.ln_bwd(du2, dh1, U[s, n, 0], lw["n1"][0], EPS, G[lp + "norm1.weight"], G[lp + "norm1.bias"], du=du1, dub=dsab,
                           dubT=(DYT[n]["sa"], s * R), p_a=p, site_a=site(1), seed=seed)
                dv32 = ops.gemm(dsab, lw["wsoT"], None, out_fp32=True)
                ops.rows_to_bf16(dv32, y=dvb, yT=(DYT[n]["v"], s * R), group=64, p=p, site=site(0), seed=seed)
                dx_sa = ops.gemm(dvb, lw["wvT"], None, out_fp32=True)
                dx_a, dx_b = du1, dx_sa
                if s == 0:
                    # every recurrent step has now contributed to layer n: its weight gradients are K-loops over the steps
                    sa_w, sa_b = G[lp + "self_attn.in_proj_weight"], G[lp + "self_attn.in_proj_bias"]
                    ca_w, ca_b = G[lp + "multihead_attn.in_proj_weight"], G[lp + "multihead_attn.in_proj_bias"]
                    ops.gemm(DYT[n]["v"], XT[n]["x"], None, out_fp32=True, out=sa_w[2 * D:])
                    ops.rowsum_bf16(DYT[n]["v"], S * R, sa_b[2 * D:])
                    ops.gemm(DYT[n]["sa"], XT[n]["v"], None, out_fp32=True, out=G[lp + "self_attn.out_proj.weight"])
                    ops.rowsum_bf16(DYT[n]["sa"], S * R, G[lp + "self_attn.out_proj.bias"])
                    ops.gemm(DYT[n]["q"], XT[n]["h1"], None, out_fp32=True, out=ca_w[:D])
                    ops.rowsum_bf16(DYT[n]["q"], S * R, ca_b[:D])
                    ops.gemm(DYT[n]["o"], XT[n]["ctx"], None, out_fp32=True, out=G[lp + "multihead_attn.out_proj.weight"])
                    ops.rowsum_bf16(DYT[n]["o"], S * R, G[lp + "multihead_attn.out_proj.bias"])
                    ops.gemm(DYT[n]["f"], XT[n]["h2"], None, out_fp32=True, out=G[lp + "linear1.weight"])
                    ops.rowsum_bf16(DYT[n]["f"], S * R, G[lp + "linear1.bias"])
                    ops.gemm(DYT[n]["y"], XT[n]["g"], None, out_fp32=True, out=G[lp + "linear2.weight"])
                    ops.rowsum_bf16(DYT[n]["y"], S * R, G[lp + "linear2.bias"])
                    # K / V in-projection of this layer: dW = [dK | dV]^T X_mem over every memory token (video once, text per step)
                    c0 = n * 2 * D
                    ops.transpose_bf16(dkv_video[:, c0:c0 + 2 * D], dKVT[:, :Nv])
                    for s2 in range(S):
                        ops.transpose_bf16(dkv_text[s2][:, c0:c0 + 2 * D], dKVT[:, Nv + s2 * Nt:Nv + (s2 + 1) * Nt])
                    ops.gemm(dKVT, XmT, None, out_fp32=True, out=ca_w[D:])
                    ops.rowsum_bf16(dKVT, Nv + S * Nt, ca_b[D:])
                    span = [offsets[k] for k in names if k.startswith(lp)]
                    reduce_span(min(a for a, _ in span), max(b for _, b in span))
            # the layer-0 input of step s is tok_s: gradient = residual of LN_f + both layer-0 contributions
            if s > 0:
                nxt = z(R, D)
                torch.add(du_f, dx_a, out=nxt)
                nxt.add_(dx_b)
                dtok32 = nxt
            else:
                for t in (du_f, dx_a, dx_b):
                    ops.colsum(t, d_tok0)
        G["fusion_transformer.summarization_token"].copy_(d_tok0.view(1, 1, D))

        # ---- memory path: d(embedded memory) = dKV W_kv ; pos-embed / projection gradients
        dkv_text_sum = dkv_text[0]
        if S == 2:
            dkv_text_sum = ops.add_bf16(torch.empty_like(kv_text), dkv_text[0], dkv_text[1])
        elif S >= 3:
            dkv_text_sum = ops.add_bf16(torch.empty_like(kv_text), dkv_text[0], dkv_text[1], dkv_text[2])
            for s2 in range(3, S):
                ops.add_bf16(dkv_text_sum, dkv_text_sum, dkv_text[s2])
        dvemb = ops.gemm(dkv_video, pk["kv_wT"], None, out_fp32=True)
        dtemb = ops.gemm(dkv_text_sum, pk["kv_wT"], None, out_fp32=True)
        has_proj = "proj_w" in pk
        dproj = z(B * S * T * P, D, dt=torch.bfloat16)
        ops.posembed_bwd(dvemb, proj=st["proj"], emb_cls=emb["v_cls"], emb_pos=emb["v_pos"], emb_len=emb["v_len"],
                         emb_clip=emb["v_clip"], gamma=emb["v_g"], eps=EPS, dproj=dproj, d_cls=G["video_pos_embed.emb_cls"],
                         d_pos=G["video_pos_embed.emb_pos"], d_len=G["video_pos_embed.emb_len"], d_clip=G["video_pos_embed.emb_clip"],
                         dgamma=G["video_pos_embed.layer_norm.weight"], dbeta=G["video_pos_embed.layer_norm.bias"], B=B, S=S, T=T,
                         P=P, is_text=False, p=p, site=_Sites.VIDEO, seed=seed)
        ops.posembed_bwd(dtemb, text=st["text"], emb_cls=emb["t_cls"], emb_pos=emb["t_pos"], gamma=emb["t_g"], eps=EPS,
                         d_cls=G["question_pos_embed.emb_cls"], d_pos=G["question_pos_embed.emb_pos"],
                         dgamma=G["question_pos_embed.layer_norm.weight"], dbeta=G["question_pos_embed.layer_norm.bias"], B=R, S=1,
                         T=1, P=L, is_text=True, p=p, site=_Sites.TEXT, seed=seed)
        if has_proj:
            n_tok = dproj.shape[0]
            dprojT = ops.transpose_bf16(dproj, zeros_bf(D, _round8(n_tok)))
            featT = ops.transpose_bf16(st["vf"], zeros_bf(st["vf"].shape[1], _round8(n_tok)))
            ops.gemm(dprojT, featT, None, out_fp32=True, out=G["projection_layer.weight"])
            ops.rowsum_bf16(dprojT, dproj.shape[0], G["projection_layer.bias"])
        if sync:
            layer_spans = [offsets[k] for k in names if k.startswith("fusion_transformer.transformer.layers.")]
            first_layer, last_layer = min(a for a, _ in layer_spans), max(b for _, b in layer_spans)
            reduce_span(0, first_layer)
            reduce_span(last_layer, flat.numel())
            for w in works:
                w.wait()
            flat.div_(dist.get_world_size())
        ctx.st = None
        grads = tuple(G[name] if q.requires_grad else None for name, q in zip(names, params))
        return (None, None, None, None, None) + grads


def encoder_train(module, video_features, text_features, n_cand):
    """logits fp32 [rows, n_out] (before the counting head's ReLU), differentiable w.r.t. every parameter of `module`"""
    names, params = zip(*module.named_parameters())
    return _EncoderTrain.apply(module, video_features, text_features, n_cand, names, *params)
