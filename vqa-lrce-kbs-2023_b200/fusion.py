"""Host-side mirror of lrce/models/{fusionv3.py, embedding.py}: same class names, constructor arguments, parameter
names and shapes as the reference; ``forward`` runs on liblrce_b200 kernels.

Execution plan of one forward (eval semantics — dropouts are identities, fusionv3.py:190-191, :49):
  1. projection_layer 1024->768 on every video token: one tcgen05 GEMM (fusionv3.py:184-185)
  2. VideoPosEmbed / TextPosEmbed: one fused row kernel each (embedding.py:47-63, :17-23)
  3. K/V in-projections of ALL 12 decoder layers for every memory token in ONE GEMM per modality
     ([Wk;Wv] of the 12 layers concatenated -> N = 18432). Text rows are projected once, not once per segment, and for
     multiple choice the video rows are projected once per clip, not once per candidate (fusionv3.py:259 expands them).
  4. the summarisation token: S x 12 layer-steps of 6 dependent sub-steps plus final_fc (ReLU fused for the counting
     head, fusionv3.py:368) run as ONE kernel of row-sharded 16-CTA clusters (csrc/encoder_walk.cu): the decoder weights
     are re-tiled once per weight version into the order each CTA streams them (lrce_encoder_walk_pack). Self-attention
     over a length-1 target is softmax over a single key = 1, so it reduces to out_proj(v_proj(x)); the two matrices are
     folded into one 768x768 product at pack time.
"""
from typing import Iterable, List

import torch
from torch import nn

from . import ops
from .feature_extractor import _PackedModule, _PackedWeights

EPS = 1e-12  # fusionv3.py:14,18 ; embedding.py:15,45


def init_weight(size):
    """embedding.py:4-7"""
    w = torch.empty(size)
    torch.nn.init.xavier_normal_(w)
    return nn.Parameter(w, requires_grad=True)


class TextPosEmbed(nn.Module):
    def __init__(self, seq_len: int, feature_dim: int) -> None:
        super().__init__()
        self.emb_cls = init_weight((1, 1, feature_dim))
        self.emb_pos = init_weight((1, 1 + seq_len, feature_dim))
        self.layer_norm = nn.LayerNorm(feature_dim, eps=EPS)


class VideoPosEmbed(nn.Module):
    def __init__(self, feature_dim: int, video_feature_res: Iterable[int] = (7, 7), frame_sample_size: int = 5,
                 clip_size: int = 6) -> None:
        super().__init__()
        self.emb_cls = init_weight((1, 1, 1, 1, feature_dim))
        self.emb_pos = init_weight((1, 1, 1, 1 + video_feature_res[0] * video_feature_res[1], feature_dim))
        self.emb_len = init_weight((1, 1, (frame_sample_size + 1) // 2, 1, feature_dim))
        self.emb_clip = init_weight((1, clip_size, 1, 1, feature_dim))
        self.layer_norm = nn.LayerNorm(feature_dim, eps=EPS)


class FusionTransformer(nn.Module):
    """fusionv3.py:5-51. The nn.TransformerDecoder is kept as the parameter container (identical state_dict keys)."""

    def __init__(self, feature_dim: int = 768, drop_out_rate: float = 0.1) -> None:
        super().__init__()
        decoder_layer = nn.TransformerDecoderLayer(d_model=feature_dim, nhead=12, dropout=drop_out_rate,
                                                   dim_feedforward=3072, batch_first=True, layer_norm_eps=EPS,
                                                   activation=torch.nn.functional.gelu)
        self.transformer = nn.TransformerDecoder(decoder_layer=decoder_layer, num_layers=12)
        self.fusion_layer_norm = nn.LayerNorm(feature_dim, eps=EPS)
        self.dropout = nn.Dropout(drop_out_rate)
        self.summarization_token = init_weight((1, 1, feature_dim))


class LRCEOpenEnded(_PackedModule):
    KIND = "oe"

    def __init__(self, feature_dim: int, num_classes: int, drop_out_rate: float = 0.1,
                 video_feature_res: Iterable[int] = (7, 7), video_feature_dim: int = 768, frame_sample_size: int = 5,
                 temporal_scale: List[int] = [1, 2, 3], question_seq_len: int = 30) -> None:
        super().__init__()
        if feature_dim != 768:
            raise ops._lib.LrceError("liblrce_b200's encoder kernels are specialised for feature_dim = 768")
        self.feature_dim, self.video_feature_dim, self.num_classes = feature_dim, video_feature_dim, num_classes
        self.video_pos_embed = VideoPosEmbed(feature_dim, video_feature_res, frame_sample_size, clip_size=sum(temporal_scale))
        self.question_pos_embed = TextPosEmbed(question_seq_len, feature_dim)
        if video_feature_dim != feature_dim:
            self.projection_layer = nn.Linear(video_feature_dim, feature_dim)
        self.video_dropout = nn.Dropout(drop_out_rate)
        self.question_dropout = nn.Dropout(drop_out_rate)
        self.fusion_transformer = FusionTransformer(feature_dim, drop_out_rate=drop_out_rate)
        self.final_fc = nn.Linear(feature_dim, num_classes)
        self._packed = _PackedWeights()
        self._states = {}

    # -------------------------------------------------------------------------------------------------------------
    def packed(self):
        sig = self._packed.signature(self)
        if self._packed.sig != sig:
            self._packed.data = self._pack()
            self._packed.sig = sig
            self._states = {}
        return self._packed.data

    @torch.no_grad()
    def _pack(self):
        dev = self.final_fc.weight.device
        if dev.type != "cuda":
            raise ops._lib.LrceError("LRCE fusion parameters must live on a CUDA device (no CPU fallback)")
        d = self.feature_dim
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()

        def pad8(w):
            n = w.shape[0]
            if n % 8:
                w = torch.cat([w, w.new_zeros((8 - n % 8, w.shape[1]))])
            return w

        ve, te, ft = self.video_pos_embed, self.question_pos_embed, self.fusion_transformer
        pk = dict(
            v_cls=f32(ve.emb_cls.reshape(d)), v_pos=f32(ve.emb_pos.reshape(-1, d)), v_len=f32(ve.emb_len.reshape(-1, d)),
            v_clip=f32(ve.emb_clip.reshape(-1, d)), v_g=f32(ve.layer_norm.weight), v_b=f32(ve.layer_norm.bias),
            t_cls=f32(te.emb_cls.reshape(d)), t_pos=f32(te.emb_pos.reshape(-1, d)), t_g=f32(te.layer_norm.weight),
            t_b=f32(te.layer_norm.bias), tok=f32(ft.summarization_token.reshape(1, d)),
            f_g=f32(ft.fusion_layer_norm.weight), f_b=f32(ft.fusion_layer_norm.bias),
            fc_w=bf(pad8(self.final_fc.weight.detach().float())), fc_b=f32(self.final_fc.bias), layers=[])
        walk_rows = []  # fp32 masters for lrce_encoder_walk_pack (the bf16 copies in pk["layers"] feed the training kernels)
        if hasattr(self, "projection_layer"):
            pk["proj_w"], pk["proj_b"] = bf(self.projection_layer.weight), f32(self.projection_layer.bias)
        kv_w, kv_b = [], []
        for lyr in ft.transformer.layers:
            sa_w, sa_b = lyr.self_attn.in_proj_weight.detach().double(), lyr.self_attn.in_proj_bias.detach().double()
            so_w, so_b = lyr.self_attn.out_proj.weight.detach().double(), lyr.self_attn.out_proj.bias.detach().double()
            ca_w, ca_b = lyr.multihead_attn.in_proj_weight.detach().float(), lyr.multihead_attn.in_proj_bias.detach().float()
            scale = (d // 12) ** -0.5
            walk_rows.append([f32((so_w @ sa_w[2 * d:]).float()), f32(ca_w[:d] * scale), f32(lyr.multihead_attn.out_proj.weight),
                              f32(lyr.linear1.weight), f32(lyr.linear2.weight), f32((so_w @ sa_b[2 * d:] + so_b).float()),
                              f32(ca_b[:d] * scale), f32(lyr.multihead_attn.out_proj.bias), f32(lyr.linear1.bias),
                              f32(lyr.linear2.bias), f32(lyr.norm1.weight), f32(lyr.norm1.bias), f32(lyr.norm2.weight),
                              f32(lyr.norm2.bias), f32(lyr.norm3.weight), f32(lyr.norm3.bias)])
            pk["layers"].append(dict(
                # length-1 self-attention == out_proj(v_proj(x)): fold the two linears (exact algebra, fp64 product)
                sa_w=bf((so_w @ sa_w[2 * d:]).float()), sa_b=f32((so_w @ sa_b[2 * d:] + so_b).float()),
                q_w=bf(ca_w[:d] * scale), q_b=f32(ca_b[:d] * scale),
                o_w=bf(lyr.multihead_attn.out_proj.weight), o_b=f32(lyr.multihead_attn.out_proj.bias),
                w1=bf(lyr.linear1.weight), b1=f32(lyr.linear1.bias), w2=bf(lyr.linear2.weight), b2=f32(lyr.linear2.bias),
                n1=(f32(lyr.norm1.weight), f32(lyr.norm1.bias)), n2=(f32(lyr.norm2.weight), f32(lyr.norm2.bias)),
                n3=(f32(lyr.norm3.weight), f32(lyr.norm3.bias))))
            kv_w.append(ca_w[d:])
            kv_b.append(ca_b[d:])
        pk["kv_w"], pk["kv_b"] = bf(torch.cat(kv_w)), f32(torch.cat(kv_b))  # [12*1536, 768]: layer-major, [k | v]
        # device table of per-layer pointers in the order lrce_encoder_walk_pack documents (include/lrce_b200.h). The fp32
        # masters are only needed by the pack kernels, which are enqueued on the current stream: the caching allocator hands
        # their blocks out again only to work ordered after those kernels on the same stream, so no synchronisation is needed.
        table = torch.tensor([[t.data_ptr() for t in row] for row in walk_rows], dtype=torch.int64, device=dev)
        pk["walk_packed"] = ops.encoder_walk_pack(table, len(walk_rows), f32(self.final_fc.weight), pk["fc_b"],
                                                  self.final_fc.out_features)
        return pk

    def _validate(self, video_features, text_features, n_cand):
        if video_features.dim() != 5 or text_features.dim() != 3:
            raise ops._lib.LrceError(f"expected video features (B, S, T, P, Dv) and text features (B, L, 768), got "
                                     f"{tuple(video_features.shape)} / {tuple(text_features.shape)}")
        B, S, T, P, Dv = video_features.shape
        Bq, L, d = text_features.shape
        # the pos-embed kernels index emb_clip[s], emb_len[t], emb_pos[p] / emb_pos[l] directly: a dataset / model mismatch in
        # temporal_scale, frame_sample_size, video_feature_res or text_seq_len must fail here, as the reference's broadcast
        # adds do (embedding.py:21, :55-57), not read out of bounds
        ve, te = self.video_pos_embed, self.question_pos_embed
        if (S != ve.emb_clip.shape[1] or T != ve.emb_len.shape[2] or P + 1 != ve.emb_pos.shape[3]
                or L + 1 != te.emb_pos.shape[1] or d != self.feature_dim or Dv != self.video_feature_dim
                or Bq != B * n_cand):
            raise ops._lib.LrceError(
                f"shape mismatch with the embedding tables: video (B={B}, S={S}, T={T}, P={P}, Dv={Dv}) vs emb_clip rows "
                f"{ve.emb_clip.shape[1]}, emb_len rows {ve.emb_len.shape[2]}, emb_pos rows {ve.emb_pos.shape[3]}, feature dim "
                f"{self.video_feature_dim}; text (rows={Bq}, L={L}, d={d}) vs emb_pos rows {te.emb_pos.shape[1]}, "
                f"{n_cand} candidate(s) per clip")

    # -------------------------------------------------------------------------------------------------------------
    def _encode(self, video_features, text_features, n_cand, act=ops.ACT_NONE, taps=None):
        """video_features (B, S, T, P, Dv) bf16; text_features (B*n_cand, L, 768) bf16/fp32 -> (B*n_cand, classes)."""
        pk = self.packed()
        self._validate(video_features, text_features, n_cand)
        B, S, T, P, Dv = video_features.shape
        Bq, L, d = text_features.shape
        dev = video_features.device
        vf = video_features.reshape(B * S * T * P, Dv)
        if vf.dtype != torch.bfloat16:
            raise ops._lib.LrceError(f"video features must be bf16 (the Swin kernels emit bf16), got {vf.dtype}")
        proj = ops.gemm(vf.contiguous(), pk["proj_w"], pk["proj_b"]) if "proj_w" in pk else vf.contiguous()
        vemb = ops.video_posembed_ln(proj, pk["v_cls"], pk["v_pos"], pk["v_len"], pk["v_clip"], pk["v_g"], pk["v_b"], EPS,
                                     B, S, T, P)
        temb = ops.text_posembed_ln(text_features.contiguous(), pk["t_cls"], pk["t_pos"], pk["t_g"], pk["t_b"], EPS)
        Tv, Lt = T * (P + 1), L + 1
        if taps is not None:
            taps["video_embedded"], taps["text_embedded"] = vemb, temb
        n_out = self.final_fc.out_features
        st = self._token_state(pk, dev, B, Bq, S, Tv, Lt)
        ops.gemm(vemb.view(B * S * Tv, d), pk["kv_w"], pk["kv_b"], out=st["kv_video"])
        ops.gemm(temb.view(Bq * Lt, d), pk["kv_w"], pk["kv_b"], out=st["kv_text"])
        out = torch.empty((Bq, n_out), device=dev, dtype=torch.float32)
        tap = torch.empty((S, Bq, d), device=dev, dtype=torch.float32) if taps is not None else None
        ops.encoder_walk(pk["walk_packed"], len(pk["layers"]), st["kv_video"], st["kv_text"], pk["tok"], pk["f_g"], pk["f_b"],
                         EPS, n_out, act, out, Bq, S, Tv, Lt, n_cand, tokens_tap=tap)
        if taps is not None:
            for s in range(S):
                taps[f"token.s{s}"] = tap[s]
        return out

    def _token_state(self, pk, dev, B, Bq, S, Tv, Lt):
        """K/V buffers for one problem shape (reused across forwards)"""
        key = (dev, B, Bq, S, Tv, Lt)
        st = self._states.get(key)
        if st is None:
            if len(self._states) > 8:
                self._states.clear()
            h = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)
            st = dict(kv_video=h(B * S * Tv, pk["kv_w"].shape[0]), kv_text=h(Bq * Lt, pk["kv_w"].shape[0]))
            self._states[key] = st
        return st

    # -------------------------------------------------------------------------------------------------------------
    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _train_plan(self, B, S, T, P, Dv, R, L, n_cand, text_dtype, p, dev):
        """fixed-address buffers + CUDA graphs of the training step for one problem shape (train._Plan)"""
        from . import train

        key = (B, S, T, P, Dv, R, L, n_cand, text_dtype, p, dev, tuple(q.data_ptr() for q in self.parameters()))
        plans = self.__dict__.setdefault("_plans", {})
        plan = plans.get(key)
        if plan is None:
            if len(plans) >= 2:  # a plan holds ~2 GB at batch 32: keep the two most recent shapes (train + validation batch)
                plans.pop(next(iter(plans)))
            plan = plans[key] = train._Plan(self, B, S, T, P, Dv, R, L, n_cand, text_dtype, p, dev)
        return plan

    def _encode_train(self, video_features, text_features, n_cand, act=ops.ACT_NONE):
        """differentiable forward (grad enabled): the hand-written training kernels behind one autograd node (train.py);
        `grad_sync = "overlap"` on this module additionally all-reduces the gradients inside the backward pass."""
        from . import train

        self._validate(video_features, text_features, n_cand)
        out = train.encoder_train(self, video_features, text_features, n_cand)
        return torch.relu(out) if act == ops.ACT_RELU else out  # LRCECount's ReLU (fusionv3.py:368) on a (B,) vector

    def forward(self, video_features, text_features, texts_attention_mask=None, taps=None):
        """(B, S, T, 49, Dv), (B, L, 768) -> (B, num_classes) fp32 (fusionv3.py:168-198). `texts_attention_mask` is
        accepted and ignored, exactly like the reference (fusionv3.py:31)."""
        B = video_features.shape[0]
        if self._wants_grad() and taps is None:
            return self._encode_train(video_features, text_features, 1).view(B, -1)
        return self._encode(video_features, text_features, 1, taps=taps).view(B, -1)


class LRCEMultipleChoice(LRCEOpenEnded):
    KIND = "mc"

    def __init__(self, feature_dim, num_classes, drop_out_rate=0.1, video_feature_res=(7, 7), video_feature_dim=768,
                 frame_sample_size=5, temporal_scale=[1, 2, 3], qa_seq_len=40) -> None:
        super().__init__(feature_dim, num_classes, drop_out_rate, video_feature_res, video_feature_dim, frame_sample_size,
                         temporal_scale, qa_seq_len)

    def forward(self, video_features, text_features, texts_attention_mask=None, taps=None):
        """text_features (B, n_cand, L, 768) -> (B, n_cand) (fusionv3.py:230-265)"""
        B, n_cand = text_features.shape[:2]
        if self._wants_grad() and taps is None:
            return self._encode_train(video_features, text_features.flatten(0, 1), n_cand).view(B, n_cand)
        return self._encode(video_features, text_features.flatten(0, 1), n_cand, taps=taps).view(B, n_cand)


class LRCECount(LRCEOpenEnded):
    KIND = "count"

    def __init__(self, feature_dim, num_classes=1, drop_out_rate=0.1, video_feature_res=(7, 7), video_feature_dim=768,
                 frame_sample_size=5, temporal_scale=[1, 2, 3], question_seq_len=30) -> None:
        super().__init__(feature_dim, 1, drop_out_rate, video_feature_res, video_feature_dim, frame_sample_size,
                         temporal_scale, question_seq_len)

    def forward(self, video_features, text_features, texts_attention_mask=None, taps=None):
        """(B,) = relu(final_fc(token)) (fusionv3.py:360-369)"""
        B = video_features.shape[0]
        if self._wants_grad() and taps is None:
            return self._encode_train(video_features, text_features, 1, act=ops.ACT_RELU).view(B)
        return self._encode(video_features, text_features, 1, act=ops.ACT_RELU, taps=taps).view(B)
