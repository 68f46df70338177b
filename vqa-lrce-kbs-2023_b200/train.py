"""Training step of the recurrent cross-modal encoder on liblrce_b200 (BASELINE.json configs[4]: cross-modal encoder forward +
backward; the reference runs it as `scaler.scale(loss).backward()` over fusionv3.py:41-51 / :168-198 under DDP,
agent_oe.py:28-42, agent_base.py:75-76).

`encoder_train(module, video_features, text_features, n_cand)` is a `torch.autograd.Function` over the module's own
parameters: autograd sees ONE node whose backward hands every encoder parameter its gradient, so the reference agent's
GradScaler / AdamW / plain `DDP(model)` work unchanged. Both passes are sequences of hand-written kernels:

  forward   bf16 repack of the weights -> projection GEMM -> pos-embed kernels (+ dropout) -> K/V GEMM of all 12 layers ->
            for every recurrent step and layer: v-proj / out-proj GEMMs of the length-1 self-attention, LayerNorm steps
            (lrce_add_ln_768), q GEMM, single-query cross attention (lrce_xattn_fwd), FFN GEMMs around a GELU row pass ->
            answer-head GEMM. Activations the backward needs are kept: pre-LayerNorm sums, q, attention probabilities, FFN
            pre-activations, and — TRANSPOSED and stacked over the recurrent steps — every Linear's input.
  backward  the mirror image: LayerNorm / GELU / cross-attention backward kernels produce the bf16 operands of
            dX = dY W (one lrce_gemm_bf16 per Linear and step) and, transposed and stacked over the steps, of
            dW = dY^T X (one lrce_gemm_bf16 per weight: the sum over recurrent steps and rows is the GEMM's K loop).
            Layer n's weight gradients are complete when the s = 0 pass leaves layer n; with `grad_sync="overlap"` their
            NCCL all-reduce is issued right there and runs under the backward of layers n-1 .. 0 and of the memory path.

The step is ~1500 short launches: launch-bound from Python (31 ms of host time for ~5 ms of device work). All buffers of one
problem shape therefore live in a `_Plan` with fixed addresses, and the launch sequence is captured once into CUDA graphs —
one for the forward, and for the backward one per all-reduce boundary (head + steps S-1..1, then one per layer of the s = 0
pass, then the memory path) — and replayed afterwards (LRCE_B200_TRAIN_GRAPH=0 launches directly). Consequence of the
fixed addresses: a plan holds the activations of ONE forward; calling forward twice before backward raises in backward.

Dropout (train mode) is a counter-based hash keyed by (seed, site, index), regenerated in the backward pass
(csrc/seqops.cu); the per-step seed lives in device memory so that graph replays draw fresh masks. It cannot reproduce
torch's RNG stream, so gradient parity is checked with drop_out_rate = 0.
The extractors feeding the encoder are forward-only (frozen by E2EBase): no gradient is returned for the features.
"""
import os

import torch
import torch.distributed as dist

from . import ops

EPS = 1e-12
D = 768
FF = 3072
_counter = [0]
LP = "fusion_transformer.transformer.layers.%d."


def _round8(n):
    """leading dimension of a stacked / transposed operand: a multiple of 8 elements (16-byte TMA pitch), at least one
    64-element k-block wide; the pad columns are zero"""
    return max(64, (n + 7) // 8 * 8)


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


class _Plan:
    """fixed-address buffers + captured launch sequences of the training step for one problem shape"""

    def __init__(self, module, B, S, T, P, Dv, R, L, n_cand, text_dtype, p, dev):
        self.m = module
        self.B, self.S, self.T, self.P, self.Dv, self.R, self.L, self.n_cand, self.p = B, S, T, P, Dv, R, L, n_cand, p
        self.Tv, self.Lt = T * (P + 1), L + 1
        self.NL = NL = len(module.fusion_transformer.transformer.layers)
        self.n_out = n_out = module.final_fc.out_features
        self.has_proj = hasattr(module, "projection_layer")
        self.ld, self.ldR = _round8(S * R), _round8(R)
        self.gen = 0           # forward generation whose activations the buffers hold
        self.graphs = None     # {"fwd": graph, "bwd": [graphs]} once captured
        self.n_launch = {}
        self.seed_buf = torch.zeros(1, device=dev, dtype=torch.int64)
        self.names, self.params = zip(*module.named_parameters())
        self.param_key = tuple(q.data_ptr() for q in self.params)
        z = lambda *s, dt=torch.float32: torch.empty(s, device=dev, dtype=dt)
        zb = lambda *s: torch.zeros(s, device=dev, dtype=torch.bfloat16)
        bf = torch.bfloat16
        Nf, Nv, Nt = B * S * T * P, B * S * self.Tv, R * self.Lt
        self.Nv, self.Nt = Nv, Nt
        ld, ldR = self.ld, self.ldR
        # ---- packed weights (refilled in place by pack())
        W = {"layers": []}
        for _ in range(NL):
            W["layers"].append(dict(wv=z(D, D, dt=bf), wso=z(D, D, dt=bf), wq=z(D, D, dt=bf), wo=z(D, D, dt=bf), w1=z(FF, D, dt=bf),
                                    w2=z(D, FF, dt=bf), wvT=zb(D, _round8(D)), wsoT=zb(D, _round8(D)), wqT=zb(D, _round8(D)),
                                    woT=zb(D, _round8(D)), w1T=zb(D, _round8(FF)), w2T=zb(FF, _round8(D))))
        W["kv_w"], W["kv_b"], W["kv_wT"] = z(NL * 2 * D, D, dt=bf), z(NL * 2 * D), zb(D, _round8(NL * 2 * D))
        if self.has_proj:
            W["proj_w"] = z(D, Dv, dt=bf)
        n_pad = (n_out + 127) // 128 * 128
        W["fc_w"], W["fc_b"], W["fc_wT"] = zb(n_pad, D), torch.zeros(n_pad, device=dev), zb(D, _round8(n_out))
        self.W = W
        # ---- forward
        self.vf, self.text = z(Nf, Dv, dt=bf), z(R, L, D, dt=text_dtype)
        self.proj = z(Nf, D, dt=bf) if self.has_proj else self.vf
        self.vemb, self.temb = z(Nv, D, dt=bf), z(Nt, D, dt=bf)
        self.kv_video, self.kv_text = z(Nv, NL * 2 * D, dt=bf), z(Nt, NL * 2 * D, dt=bf)
        self.U, self.Q, self.F = z(S, NL, 3, R, D), z(S, NL, R, D), z(S, NL, R, FF)
        self.Pst, self.UF = z(S, NL, R * 12, ops.XATTN_MAXK), z(S, R, D)
        self.XT = [dict(x=zb(D, ld), v=zb(D, ld), h1=zb(D, ld), ctx=zb(D, ld), h2=zb(D, ld), g=zb(FF, ld)) for _ in range(NL)]
        self.tokT = zb(D, ldR)
        self.tok = z(S + 1, R, D)
        self.x32, self.h1_32, self.h2_32, self.t32 = z(R, D), z(R, D), z(R, D), z(R, D)
        self.xb, self.h1b, self.h2b, self.vb, self.ctxb = (z(R, D, dt=bf) for _ in range(5))
        self.gb = z(R, FF, dt=bf)
        self.logits_pad = z(R, n_pad)
        # ---- backward
        sizes = [q.numel() for q in self.params]
        self.flat = torch.zeros(sum(sizes), device=dev)
        self.G, self.span, off = {}, {}, 0
        for name, q, n_el in zip(self.names, self.params, sizes):
            self.G[name] = self.flat[off:off + n_el].view(q.shape)
            self.span[name] = (off, off + n_el)
            off += n_el
        self.dlogits = z(R, n_out)
        Kp = W["fc_wT"].shape[1]
        self.dl, self.dlb, self.dlT = torch.zeros((R, Kp), device=dev), z(R, Kp, dt=bf), zb(Kp, ldR)
        self.Ktot = Ktot = _round8(Nv + S * Nt)
        self.XmT, self.dKVT = zb(D, Ktot), zb(2 * D, Ktot)
        self.dkv_video = torch.zeros_like(self.kv_video)
        self.dkv_text = [z(Nt, NL * 2 * D, dt=bf) for _ in range(S)]
        self.dkv_text_sum = z(Nt, NL * 2 * D, dt=bf) if S > 1 else self.dkv_text[0]
        self.DYT = [dict(y=zb(D, ld), f=zb(FF, ld), o=zb(D, ld), q=zb(D, ld), sa=zb(D, ld), v=zb(D, ld)) for _ in range(NL)]
        self.dyb, self.dob, self.dqb, self.dsab, self.dvb = (z(R, D, dt=bf) for _ in range(5))
        self.dfb = z(R, FF, dt=bf)
        self.du3, self.du2, self.du1, self.du_f, self.g32, self.dtok = (z(R, D) for _ in range(6))
        self.g32b = z(R, D)
        self.dg32 = z(R, FF)
        self.d_tok0 = torch.zeros(D, device=dev)
        self.dvemb, self.dtemb = z(Nv, D), z(Nt, D)
        self.dproj = z(Nf, D, dt=bf)
        self.dprojT, self.featT = zb(D, _round8(Nf)), zb(Dv, _round8(Nf))

    # ------------------------------------------------------------------------------------------------------------
    def pack(self):
        """bf16 operands of the step from the live parameters: every Linear weight as stored ([N, K]: forward and dW GEMMs)
        and transposed ([K, N]: the dX GEMMs). Part of the captured forward: replays re-read the parameters."""
        m, W = self.m, self.W
        for n, lyr in enumerate(m.fusion_transformer.transformer.layers):
            sa, ca, lw = lyr.self_attn, lyr.multihead_attn, W["layers"][n]
            for dst, src in (("wv", sa.in_proj_weight.detach()[2 * D:]), ("wso", sa.out_proj.weight.detach()),
                             ("wq", ca.in_proj_weight.detach()[:D]), ("wo", ca.out_proj.weight.detach()),
                             ("w1", lyr.linear1.weight.detach()), ("w2", lyr.linear2.weight.detach())):
                lw[dst].copy_(src)
                ops.transpose_bf16(lw[dst], lw[dst + "T"])
            W["kv_w"][n * 2 * D:(n + 1) * 2 * D].copy_(ca.in_proj_weight.detach()[D:])
            W["kv_b"][n * 2 * D:(n + 1) * 2 * D].copy_(ca.in_proj_bias.detach()[D:])
        ops.transpose_bf16(W["kv_w"], W["kv_wT"])
        if self.has_proj:
            W["proj_w"].copy_(m.projection_layer.weight.detach())
        W["fc_w"][:self.n_out].copy_(m.final_fc.weight.detach())
        W["fc_b"][:self.n_out].copy_(m.final_fc.bias.detach())
        ops.transpose_bf16(W["fc_w"][:self.n_out], W["fc_wT"])

    def _lnp(self, n, k):
        lyr = self.m.fusion_transformer.transformer.layers[n]
        norm = getattr(lyr, "norm%d" % k)
        return norm.weight.detach(), norm.bias.detach()

    def _bias(self, n):
        lyr = self.m.fusion_transformer.transformer.layers[n]
        sa, ca = lyr.self_attn, lyr.multihead_attn
        return dict(bv=sa.in_proj_bias.detach()[2 * D:], bso=sa.out_proj.bias.detach(), bq=ca.in_proj_bias.detach()[:D],
                    bo=ca.out_proj.bias.detach(), b1=lyr.linear1.bias.detach(), b2=lyr.linear2.bias.detach())

    def _emb(self):
        m = self.m
        ve, te, ft = m.video_pos_embed, m.question_pos_embed, m.fusion_transformer
        r = lambda t, *s: t.detach().reshape(*s)
        return dict(v_cls=r(ve.emb_cls, D), v_pos=r(ve.emb_pos, -1, D), v_len=r(ve.emb_len, -1, D), v_clip=r(ve.emb_clip, -1, D),
                    v_g=ve.layer_norm.weight.detach(), v_b=ve.layer_norm.bias.detach(), t_cls=r(te.emb_cls, D),
                    t_pos=r(te.emb_pos, -1, D), t_g=te.layer_norm.weight.detach(), t_b=te.layer_norm.bias.detach(),
                    f_g=ft.fusion_layer_norm.weight.detach(), f_b=ft.fusion_layer_norm.bias.detach(),
                    tok0=r(ft.summarization_token, 1, D))

    # ------------------------------------------------------------------------------------------------------------
    def forward_launches(self):
        B, S, T, P, R, NL, Tv, Lt, p, n_cand = self.B, self.S, self.T, self.P, self.R, self.NL, self.Tv, self.Lt, self.p, self.n_cand
        W, XT, emb = self.W, self.XT, self._emb()
        seed = (0, self.seed_buf)
        self.pack()
        if self.has_proj:
            ops.gemm(self.vf, W["proj_w"], self.m.projection_layer.bias.detach(), out=self.proj)
        ops.video_posembed_ln(self.proj, emb["v_cls"], emb["v_pos"], emb["v_len"], emb["v_clip"], emb["v_g"], emb["v_b"], EPS,
                              B, S, T, P, out=self.vemb)
        ops.text_posembed_ln(self.text, emb["t_cls"], emb["t_pos"], emb["t_g"], emb["t_b"], EPS, out=self.temb)
        ops.dropout_bf16_(self.vemb, p, _Sites.VIDEO, seed)
        ops.dropout_bf16_(self.temb, p, _Sites.TEXT, seed)
        ops.gemm(self.vemb, W["kv_w"], W["kv_b"], out=self.kv_video)
        ops.gemm(self.temb, W["kv_w"], W["kv_b"], out=self.kv_text)
        x32, xb, h1_32, h1b, h2_32, h2b, t32 = self.x32, self.xb, self.h1_32, self.h1b, self.h2_32, self.h2b, self.t32
        self.tok[0].copy_(emb["tok0"].expand(R, D))
        x32.copy_(self.tok[0])
        ops.rows_to_bf16(x32, y=xb, yT=(XT[0]["x"], 0))
        for s in range(S):
            for n, lw in enumerate(W["layers"]):
                site = lambda k: _Sites.layer(s, n, NL, k)
                b = self._bias(n)
                n1, n2, n3 = self._lnp(n, 1), self._lnp(n, 2), self._lnp(n, 3)
                ops.gemm_skinny(xb, lw["wv"], b["bv"], t32)
                ops.rows_to_bf16(t32, y=self.vb, yT=(XT[n]["v"], s * R), group=64, p=p, site=site(0), seed=seed)
                ops.gemm_skinny(self.vb, lw["wso"], b["bso"], t32)
                ops.add_ln(t32, x32, n1[0], n1[1], EPS, u_out=self.U[s, n, 0], y_f32=h1_32, y_bf16=h1b,
                           yT=(XT[n]["h1"], s * R), p_a=p, site_a=site(1), seed=seed)
                ops.gemm_skinny(h1b, lw["wq"], b["bq"], self.Q[s, n])
                ops.xattn_fwd(self.Q[s, n], self.kv_video, self.kv_text, n * 2 * D, R, S, s, Tv, Lt, n_cand, self.Pst[s, n], self.ctxb,
                              ctxT=(XT[n]["ctx"], s * R), p=p, site=site(2), seed=seed)
                ops.gemm_skinny(self.ctxb, lw["wo"], b["bo"], t32)
                ops.add_ln(t32, h1_32, n2[0], n2[1], EPS, u_out=self.U[s, n, 1], y_f32=h2_32, y_bf16=h2b,
                           yT=(XT[n]["h2"], s * R), p_a=p, site_a=site(3), seed=seed)
                ops.gemm_skinny(h2b, lw["w1"], b["b1"], self.F[s, n])
                ops.rows_to_bf16(self.F[s, n], y=self.gb, yT=(XT[n]["g"], s * R), mode=ops.ROWS_GELU_FWD, p=p, site=site(4), seed=seed)
                ops.gemm_skinny(self.gb, lw["w2"], b["b2"], t32)
                nxt = (XT[n + 1]["x"], s * R) if n + 1 < NL else None
                ops.add_ln(t32, h2_32, n3[0], n3[1], EPS, u_out=self.U[s, n, 2], y_f32=x32, y_bf16=xb, yT=nxt,
                           p_a=p, site_a=site(5), seed=seed)
            # tok_{s+1} = dropout(LN_f(tok_s + x_12))   (fusionv3.py:47-49)
            yT = (XT[0]["x"], (s + 1) * R) if s + 1 < S else (self.tokT, 0)
            ops.add_ln(x32, self.tok[s], emb["f_g"], emb["f_b"], EPS, u_out=self.UF[s], y_f32=self.tok[s + 1], y_bf16=xb, yT=yT,
                       p_out=p, site_out=_Sites.outer(s), seed=seed)
            x32.copy_(self.tok[s + 1])
        ops.gemm(xb, W["fc_w"], W["fc_b"], out_fp32=True, out=self.logits_pad)

    # ------------------------------------------------------------------------------------------------------------
    def _layer_backward(self, s, n, dx_a, dx_b):
        """one decoder layer of one recurrent step; returns the two contributions to the gradient of the layer's input"""
        R, S, NL, Tv, Lt, p, n_cand = self.R, self.S, self.NL, self.Tv, self.Lt, self.p, self.n_cand
        lw, lp, G, DYT = self.W["layers"][n], LP % n, self.G, self.DYT[n]
        seed = (0, self.seed_buf)
        site = lambda k: _Sites.layer(s, n, NL, k)
        n1, n2, n3 = self._lnp(n, 1), self._lnp(n, 2), self._lnp(n, 3)
        # LN3 / FFN
        ops.ln_bwd(dx_a, dx_b, self.U[s, n, 2], n3[0], EPS, G[lp + "norm3.weight"], G[lp + "norm3.bias"], du=self.du3, dub=self.dyb,
                   dubT=(DYT["y"], s * R), p_a=p, site_a=site(5), seed=seed)
        ops.gemm_skinny(self.dyb, lw["w2T"], None, self.dg32)
        ops.rows_to_bf16(self.dg32, aux=self.F[s, n], y=self.dfb, yT=(DYT["f"], s * R), mode=ops.ROWS_GELU_BWD, p=p, site=site(4),
                         seed=seed)
        ops.gemm_skinny(self.dfb, lw["w1T"], None, self.g32)
        # LN2 / cross attention
        ops.ln_bwd(self.du3, self.g32, self.U[s, n, 1], n2[0], EPS, G[lp + "norm2.weight"], G[lp + "norm2.bias"], du=self.du2,
                   dub=self.dob, dubT=(DYT["o"], s * R), p_a=p, site_a=site(3), seed=seed)
        ops.gemm_skinny(self.dob, lw["woT"], None, self.g32)
        ops.xattn_bwd(self.Q[s, n], self.kv_video, self.kv_text, n * 2 * D, R, S, s, Tv, Lt, n_cand, self.Pst[s, n], self.g32, self.dqb,
                      self.dkv_video, self.dkv_text[s], dqT=(DYT["q"], s * R), p=p, site=site(2), seed=seed)
        ops.gemm_skinny(self.dqb, lw["wqT"], None, self.g32)
        # LN1 / length-1 self-attention
        ops.ln_bwd(self.du2, self.g32, self.U[s, n, 0], n1[0], EPS, G[lp + "norm1.weight"], G[lp + "norm1.bias"], du=self.du1,
                   dub=self.dsab, dubT=(DYT["sa"], s * R), p_a=p, site_a=site(1), seed=seed)
        ops.gemm_skinny(self.dsab, lw["wsoT"], None, self.g32)
        ops.rows_to_bf16(self.g32, y=self.dvb, yT=(DYT["v"], s * R), group=64, p=p, site=site(0), seed=seed)
        ops.gemm_skinny(self.dvb, lw["wvT"], None, self.g32b)
        return self.du1, self.g32b

    def _layer_weight_grads(self, n):
        """every recurrent step has contributed to layer n: its weight gradients are K-loops over the stacked steps"""
        S, R, Nv, Nt, G, XT, DYT, lp = self.S, self.R, self.Nv, self.Nt, self.G, self.XT[n], self.DYT[n], LP % n
        sa_w, sa_b = G[lp + "self_attn.in_proj_weight"], G[lp + "self_attn.in_proj_bias"]
        ca_w, ca_b = G[lp + "multihead_attn.in_proj_weight"], G[lp + "multihead_attn.in_proj_bias"]
        for dy, x, w, b in ((DYT["v"], XT["x"], sa_w[2 * D:], sa_b[2 * D:]),
                            (DYT["sa"], XT["v"], G[lp + "self_attn.out_proj.weight"], G[lp + "self_attn.out_proj.bias"]),
                            (DYT["q"], XT["h1"], ca_w[:D], ca_b[:D]),
                            (DYT["o"], XT["ctx"], G[lp + "multihead_attn.out_proj.weight"], G[lp + "multihead_attn.out_proj.bias"]),
                            (DYT["f"], XT["h2"], G[lp + "linear1.weight"], G[lp + "linear1.bias"]),
                            (DYT["y"], XT["g"], G[lp + "linear2.weight"], G[lp + "linear2.bias"])):
            ops.gemm(dy, x, None, out_fp32=True, out=w)
            ops.rowsum_bf16(dy, S * R, b)
        # K / V in-projection of this layer: dW = [dK | dV]^T X_mem over every memory token (video once, text once per step)
        c0 = n * 2 * D
        ops.transpose_bf16(self.dkv_video[:, c0:c0 + 2 * D], self.dKVT[:, :Nv])
        for s2 in range(S):
            ops.transpose_bf16(self.dkv_text[s2][:, c0:c0 + 2 * D], self.dKVT[:, Nv + s2 * Nt:Nv + (s2 + 1) * Nt])
        ops.gemm(self.dKVT, self.XmT, None, out_fp32=True, out=ca_w[D:])
        ops.rowsum_bf16(self.dKVT, Nv + S * Nt, ca_b[D:])

    def backward_segments(self):
        """the backward pass as a list of (launch closure, flat-gradient span completed by it or None)"""
        S, R, NL, Nv, Nt, p, G, emb = self.S, self.R, self.NL, self.Nv, self.Nt, self.p, self.G, self._emb()
        seed = (0, self.seed_buf)

        def ln_f(s, dy):
            ops.ln_bwd(dy, None, self.UF[s], emb["f_g"], EPS, G["fusion_transformer.fusion_layer_norm.weight"],
                       G["fusion_transformer.fusion_layer_norm.bias"], du=self.du_f, p_out=p, site_out=_Sites.outer(s), seed=seed)

        def head_and_late_steps():
            self.flat.zero_()
            if self.n_cand > 1:
                self.dkv_video.zero_()  # candidates of a clip accumulate into shared video rows
            self.d_tok0.zero_()
            # answer head: dtok = dlogits W_fc ; dW_fc = dlogits^T tok ; db_fc
            self.dl[:, :self.n_out].copy_(self.dlogits)
            ops.rows_to_bf16(self.dl, y=self.dlb, yT=(self.dlT, 0))
            ops.gemm(self.dlb, self.W["fc_wT"], None, out_fp32=True, out=self.dtok)
            ops.gemm(self.dlT[:self.n_out], self.tokT, None, out_fp32=True, out=G["final_fc.weight"])
            ops.rowsum_bf16(self.dlT[:self.n_out], R, G["final_fc.bias"])
            # memory operand of the per-layer K/V weight gradient: X_mem^T = [vemb^T | temb^T x S]
            ops.transpose_bf16(self.vemb, self.XmT[:, :Nv])
            for s2 in range(S):
                ops.transpose_bf16(self.temb, self.XmT[:, Nv + s2 * Nt:Nv + (s2 + 1) * Nt])
            dtok = self.dtok
            for s in range(S - 1, 0, -1):
                ln_f(s, dtok)
                dx_a, dx_b = self.du_f, None
                for n in reversed(range(NL)):
                    dx_a, dx_b = self._layer_backward(s, n, dx_a, dx_b)
                # the layer-0 input of step s is tok_s: residual of LN_f + both layer-0 contributions
                torch.add(self.du_f, dx_a, out=self.dtok)
                self.dtok.add_(dx_b)
            ln_f(0, dtok)

        segs = [(head_and_late_steps, None)]
        for n in reversed(range(NL)):
            def layer_seg(n=n):
                # gradient of the layer's output: LN_f's input gradient for the top layer, else the two contributions the
                # layer above left in fixed buffers
                dx = (self.du_f, None) if n == NL - 1 else (self.du1, self.g32b)
                self._layer_backward(0, n, *dx)
                self._layer_weight_grads(n)
            spans = [self.span[k] for k in self.names if k.startswith(LP % n)]
            segs.append((layer_seg, (min(a for a, _ in spans), max(b for _, b in spans))))

        def memory_path():
            B, S_, T, P, L = self.B, self.S, self.T, self.P, self.L
            for t in (self.du_f, self.du1, self.g32b):
                ops.colsum(t, self.d_tok0)
            G["fusion_transformer.summarization_token"].copy_(self.d_tok0.view(1, 1, D))
            if S_ == 2:
                ops.add_bf16(self.dkv_text_sum, self.dkv_text[0], self.dkv_text[1])
            elif S_ >= 3:
                ops.add_bf16(self.dkv_text_sum, self.dkv_text[0], self.dkv_text[1], self.dkv_text[2])
                for s2 in range(3, S_):
                    ops.add_bf16(self.dkv_text_sum, self.dkv_text_sum, self.dkv_text[s2])
            ops.gemm(self.dkv_video, self.W["kv_wT"], None, out_fp32=True, out=self.dvemb)
            ops.gemm(self.dkv_text_sum, self.W["kv_wT"], None, out_fp32=True, out=self.dtemb)
            ops.posembed_bwd(self.dvemb, proj=self.proj, emb_cls=emb["v_cls"], emb_pos=emb["v_pos"], emb_len=emb["v_len"],
                             emb_clip=emb["v_clip"], gamma=emb["v_g"], eps=EPS, dproj=self.dproj, d_cls=G["video_pos_embed.emb_cls"],
                             d_pos=G["video_pos_embed.emb_pos"], d_len=G["video_pos_embed.emb_len"],
                             d_clip=G["video_pos_embed.emb_clip"], dgamma=G["video_pos_embed.layer_norm.weight"],
                             dbeta=G["video_pos_embed.layer_norm.bias"], B=B, S=S_, T=T, P=P, is_text=False, p=p, site=_Sites.VIDEO,
                             seed=seed)
            ops.posembed_bwd(self.dtemb, text=self.text, emb_cls=emb["t_cls"], emb_pos=emb["t_pos"], gamma=emb["t_g"], eps=EPS,
                             d_cls=G["question_pos_embed.emb_cls"], d_pos=G["question_pos_embed.emb_pos"],
                             dgamma=G["question_pos_embed.layer_norm.weight"], dbeta=G["question_pos_embed.layer_norm.bias"],
                             B=self.R, S=1, T=1, P=L, is_text=True, p=p, site=_Sites.TEXT, seed=seed)
            if self.has_proj:
                ops.transpose_bf16(self.dproj, self.dprojT)
                ops.transpose_bf16(self.vf, self.featT)
                ops.gemm(self.dprojT, self.featT, None, out_fp32=True, out=G["projection_layer.weight"])
                ops.rowsum_bf16(self.dprojT, self.dproj.shape[0], G["projection_layer.bias"])

        segs.append((memory_path, "rest"))
        return segs

    # ------------------------------------------------------------------------------------------------------------
    def _run(self, key, fn):
        """launch `fn` directly the first time (one-time kernel attribute setup happens outside any capture), capture it into a
        CUDA graph the second time, replay afterwards"""
        if os.environ.get("LRCE_B200_TRAIN_GRAPH", "1") == "0" or ops.trace is not None or torch.cuda.is_current_stream_capturing():
            fn()
            return
        g = self.graph_cache.get(key)
        if g is None:
            if key not in self.warm:
                self.warm.add(key)
                fn()
                return
            g = torch.cuda.CUDAGraph()
            n0 = ops.launches
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                fn()
            self.graph_cache[key] = g
            self.n_launch[key] = ops.launches - n0
        g.replay()
        ops.launches += self.n_launch[key]

    graph_cache, warm, pool = None, None, None

    def run_forward(self):
        if self.graph_cache is None:
            self.graph_cache, self.warm, self.pool = {}, set(), torch.cuda.graph_pool_handle()
        prev, ops.scope = ops.scope, "train"
        try:
            self._run("fwd", self.forward_launches)
        finally:
            ops.scope = prev

    def run_backward(self, sync):
        works = []
        prev, ops.scope = ops.scope, "train"
        try:
            segs = self.backward_segments()
            layer_lo = min(a for _, sp in segs if isinstance(sp, tuple) for a in sp[:1])
            layer_hi = max(sp[1] for _, sp in segs if isinstance(sp, tuple))
            for i, (fn, span) in enumerate(segs):
                self._run(("bwd", i), fn)
                if not sync:
                    continue
                if isinstance(span, tuple):
                    works.append(dist.all_reduce(self.flat[span[0]:span[1]], op=dist.ReduceOp.SUM, async_op=True))
                elif span == "rest":
                    if layer_lo > 0:
                        works.append(dist.all_reduce(self.flat[:layer_lo], op=dist.ReduceOp.SUM, async_op=True))
                    if layer_hi < self.flat.numel():
                        works.append(dist.all_reduce(self.flat[layer_hi:], op=dist.ReduceOp.SUM, async_op=True))
        finally:
            ops.scope = prev
        for w in works:
            w.wait()
        if sync:
            self.flat.div_(dist.get_world_size())


class _EncoderTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, video_features, text_features, n_cand, *params):
        B, S, T, P, Dv = video_features.shape
        R, L, _ = text_features.shape
        p = float(module.video_dropout.p) if module.training else 0.0
        plan = module._train_plan(B, S, T, P, Dv, R, L, n_cand, text_features.dtype, p, video_features.device)
        plan.vf.copy_(video_features.reshape(B * S * T * P, Dv))
        plan.text.copy_(text_features)
        plan.seed_buf.fill_(_seed())
        plan.gen += 1
        plan.run_forward()
        ctx.plan, ctx.gen = plan, plan.gen
        return plan.logits_pad[:, :plan.n_out].clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        if plan.gen != ctx.gen:
            raise ops._lib.LrceError(
                "lrce_b200 training step: another forward pass through this module overwrote the activations of the one being "
                "differentiated (the training plan keeps the buffers of ONE forward at fixed addresses for CUDA-graph replay); "
                "run backward before the next forward")
        plan.dlogits.copy_(dlogits)
        sync = getattr(plan.m, "grad_sync", "none") == "overlap" and dist.is_available() and dist.is_initialized() \
            and dist.get_world_size() > 1
        plan.run_backward(sync)
        out = plan.flat.clone()  # the plan's buffer is rewritten by the next step; .grad must not alias it
        grads, off = [], 0
        for q in plan.params:
            n_el = q.numel()
            grads.append(out[off:off + n_el].view(q.shape) if q.requires_grad else None)
            off += n_el
        return (None, None, None, None) + tuple(grads)


def encoder_train(module, video_features, text_features, n_cand):
    """logits fp32 [rows, n_out] (before the counting head's ReLU), differentiable w.r.t. every parameter of `module`"""
    return _EncoderTrain.apply(module, video_features, text_features, n_cand, *module.parameters())
