"""TEST INFRASTRUCTURE — deterministic synthetic weights for the LRCE E2E model, keyed exactly like the reference's
``state_dict()`` (`video_extractor.swin.*`, `text_extractor.bert.*`, `fusion_model.*`).

There is no network for checkpoints, so parity runs use random weights. Every tensor is drawn from its own
``torch.Generator`` seeded with crc32(key) ^ seed, which makes the values independent of construction order and lets
the golden-vector script (reference model) and the tests (oracle + CUDA path) rebuild identical weights on any machine
without shipping them. Scales are chosen so activations stay O(1) and every term of the path matters numerically
(non-zero biases, non-unit LayerNorm gains, relative-position biases of order 0.5) — a stricter test than the
reference's own init (zero biases, 0.02 trunc-normal tables, video_swin_ori.py:155, :647-654).
"""
import zlib

import torch

SWIN_DEPTHS = (2, 2, 18, 2)
SWIN_HEADS = (4, 8, 16, 32)
SWIN_EMBED = 128
SWIN_WINDOW = (8, 7, 7)  # configured window (video.py:15); the effective window is clamped to (3, 7, 7)
BERT_LAYERS = 12


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _randn(key, seed, shape, std):
    return torch.randn(shape, generator=_gen(key, seed), dtype=torch.float32) * std


def relative_position_index(window=SWIN_WINDOW):
    """Index of every (query token, key token) pair of a window into the relative-position-bias table
    (restates video_swin_ori.py:134-147): idx = (dd + Wd-1) * (2Wh-1)(2Ww-1) + (dh + Wh-1) * (2Ww-1) + (dw + Ww-1)."""
    wd, wh, ww = window
    d, h, w = torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")
    d, h, w = d.reshape(-1), h.reshape(-1), w.reshape(-1)
    idx = ((d[:, None] - d[None, :] + wd - 1) * ((2 * wh - 1) * (2 * ww - 1))
           + (h[:, None] - h[None, :] + wh - 1) * (2 * ww - 1) + (w[:, None] - w[None, :] + ww - 1))
    return idx.to(torch.int64)


def swin_schema():
    """[(key, shape, kind)] for SwinTransformer3D as VideoExtractor builds it (video.py:10-18)."""
    out = [("patch_embed.proj.weight", (SWIN_EMBED, 3, 2, 4, 4), "w"), ("patch_embed.proj.bias", (SWIN_EMBED,), "b"),
           ("patch_embed.norm.weight", (SWIN_EMBED,), "g"), ("patch_embed.norm.bias", (SWIN_EMBED,), "beta")]
    table = (2 * SWIN_WINDOW[0] - 1) * (2 * SWIN_WINDOW[1] - 1) * (2 * SWIN_WINDOW[2] - 1)
    n_tok = SWIN_WINDOW[0] * SWIN_WINDOW[1] * SWIN_WINDOW[2]
    for i, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        c = SWIN_EMBED << i
        for j in range(depth):
            p = f"layers.{i}.blocks.{j}."
            out += [(p + "norm1.weight", (c,), "g"), (p + "norm1.bias", (c,), "beta"),
                    (p + "attn.relative_position_bias_table", (table, heads), "rpb"),
                    (p + "attn.relative_position_index", (n_tok, n_tok), "rpi"),
                    (p + "attn.qkv.weight", (3 * c, c), "w"), (p + "attn.qkv.bias", (3 * c,), "b"),
                    (p + "attn.proj.weight", (c, c), "w"), (p + "attn.proj.bias", (c,), "b"),
                    (p + "norm2.weight", (c,), "g"), (p + "norm2.bias", (c,), "beta"),
                    (p + "mlp.fc1.weight", (4 * c, c), "w"), (p + "mlp.fc1.bias", (4 * c,), "b"),
                    (p + "mlp.fc2.weight", (c, 4 * c), "w"), (p + "mlp.fc2.bias", (c,), "b")]
        if i < 3:
            p = f"layers.{i}.downsample."
            out += [(p + "reduction.weight", (2 * c, 4 * c), "w"), (p + "norm.weight", (4 * c,), "g"),
                    (p + "norm.bias", (4 * c,), "beta")]
    out += [("norm.weight", (8 * SWIN_EMBED,), "g"), ("norm.bias", (8 * SWIN_EMBED,), "beta")]
    return out


def bert_schema():
    """bert-base-uncased as HF `BertModel(BertConfig())` names it (text.py:9)."""
    out = [("embeddings.word_embeddings.weight", (30522, 768), "emb"),
           ("embeddings.position_embeddings.weight", (512, 768), "emb"),
           ("embeddings.token_type_embeddings.weight", (2, 768), "emb"),
           ("embeddings.LayerNorm.weight", (768,), "g"), ("embeddings.LayerNorm.bias", (768,), "beta")]
    for n in range(BERT_LAYERS):
        p = f"encoder.layer.{n}."
        for name in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            out += [(p + name + ".weight", (768, 768), "w"), (p + name + ".bias", (768,), "b")]
        out += [(p + "attention.output.LayerNorm.weight", (768,), "g"), (p + "attention.output.LayerNorm.bias", (768,), "beta"),
                (p + "intermediate.dense.weight", (3072, 768), "w"), (p + "intermediate.dense.bias", (3072,), "b"),
                (p + "output.dense.weight", (768, 3072), "w"), (p + "output.dense.bias", (768,), "b"),
                (p + "output.LayerNorm.weight", (768,), "g"), (p + "output.LayerNorm.bias", (768,), "beta")]
    out += [("pooler.dense.weight", (768, 768), "w"), ("pooler.dense.bias", (768,), "b")]
    return out


def fusion_schema(num_classes, text_seq_len, n_segments, feature_dim=768, video_feature_dim=1024, res=(7, 7),
                  frame_sample_size=5, n_layers=12, ff=3072):
    """LRCEOpenEnded / LRCEMultipleChoice / LRCECount parameters (fusionv3.py:141-160, embedding.py:13-15, :37-45)."""
    d = feature_dim
    out = [("video_pos_embed.emb_cls", (1, 1, 1, 1, d), "emb"),
           ("video_pos_embed.emb_pos", (1, 1, 1, 1 + res[0] * res[1], d), "emb"),
           ("video_pos_embed.emb_len", (1, 1, (frame_sample_size + 1) // 2, 1, d), "emb"),
           ("video_pos_embed.emb_clip", (1, n_segments, 1, 1, d), "emb"),
           ("video_pos_embed.layer_norm.weight", (d,), "g"), ("video_pos_embed.layer_norm.bias", (d,), "beta"),
           ("question_pos_embed.emb_cls", (1, 1, d), "emb"),
           ("question_pos_embed.emb_pos", (1, 1 + text_seq_len, d), "emb"),
           ("question_pos_embed.layer_norm.weight", (d,), "g"), ("question_pos_embed.layer_norm.bias", (d,), "beta")]
    if video_feature_dim != feature_dim:
        out += [("projection_layer.weight", (d, video_feature_dim), "w"), ("projection_layer.bias", (d,), "b")]
    out += [("fusion_transformer.summarization_token", (1, 1, d), "emb")]
    for n in range(n_layers):
        p = f"fusion_transformer.transformer.layers.{n}."
        for att in ("self_attn", "multihead_attn"):
            out += [(p + att + ".in_proj_weight", (3 * d, d), "w"), (p + att + ".in_proj_bias", (3 * d,), "b"),
                    (p + att + ".out_proj.weight", (d, d), "w"), (p + att + ".out_proj.bias", (d,), "b")]
        out += [(p + "linear1.weight", (ff, d), "w"), (p + "linear1.bias", (ff,), "b"),
                (p + "linear2.weight", (d, ff), "w"), (p + "linear2.bias", (d,), "b")]
        for k in (1, 2, 3):
            out += [(p + f"norm{k}.weight", (d,), "g"), (p + f"norm{k}.bias", (d,), "beta")]
    out += [("fusion_transformer.fusion_layer_norm.weight", (d,), "g"),
            ("fusion_transformer.fusion_layer_norm.bias", (d,), "beta"),
            ("final_fc.weight", (num_classes, d), "head"), ("final_fc.bias", (num_classes,), "b")]
    return out


def _materialise(schema, prefix, seed, rpi_cache):
    sd = {}
    for key, shape, kind in schema:
        full = prefix + key
        if kind == "rpi":
            if "rpi" not in rpi_cache:
                rpi_cache["rpi"] = relative_position_index()
            t = rpi_cache["rpi"].clone()
        elif kind == "w":  # linear / conv weight: unit-gain fan-in scaling keeps activations O(1)
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = _randn(full, seed, shape, fan_in ** -0.5)
        elif kind == "head":  # answer head: larger gain so that per-sample argmax varies (SURVEY.md §8d caveat)
            t = _randn(full, seed, shape, 4.0 * shape[1] ** -0.5)
        elif kind == "b":
            t = _randn(full, seed, shape, 0.05)
        elif kind == "g":
            t = 1.0 + _randn(full, seed, shape, 0.1)
        elif kind == "beta":
            t = _randn(full, seed, shape, 0.05)
        elif kind == "rpb":
            t = _randn(full, seed, shape, 0.5)
        elif kind == "emb":
            t = _randn(full, seed, shape, 0.5 if "word_embeddings" not in key else 1.0)
        else:
            raise ValueError(kind)
        sd[full] = t
    return sd


def make_swin_state_dict(seed=0, prefix=""):
    return _materialise(swin_schema(), prefix, seed, {})


def make_fusion_state_dict(num_classes, text_seq_len, n_segments, seed=0, prefix="", **kw):
    return _materialise(fusion_schema(num_classes, text_seq_len, n_segments, **kw), prefix, seed, {})


def make_bert_state_dict(seed=0, prefix=""):
    return _materialise(bert_schema(), prefix, seed, {})


def make_e2e_state_dict(num_classes, text_seq_len, n_segments, seed=0):
    """Full E2E* state dict in the reference's key order families (text, video, fusion)."""
    sd = {}
    sd.update(make_bert_state_dict(seed, "text_extractor.bert."))
    sd.update(make_swin_state_dict(seed, "video_extractor.swin."))
    sd.update(make_fusion_state_dict(num_classes, text_seq_len, n_segments, seed, "fusion_model."))
    return sd


def make_inputs(batch, n_segments, text_seq_len, seed=1, n_candidates=0, real_len=20):
    """Synthetic batch shaped like e2e_dataset.py:118-124 output: clips uniform [0,1) fp32 (ToTensor range),
    BERT ids with [CLS]=101 ... [SEP]=102 then zero padding (so padded positions exist; the fusion model attends them,
    fusionv3.py:27-51)."""
    g = torch.Generator()
    g.manual_seed(seed)
    clips = torch.rand((batch, n_segments, 5, 3, 224, 224), generator=g, dtype=torch.float32)
    lead = (batch, n_candidates) if n_candidates else (batch,)
    ids = torch.randint(1000, 30522, lead + (text_seq_len,), generator=g, dtype=torch.int64)
    ids[..., 0] = 101
    ids[..., real_len - 1] = 102
    ids[..., real_len:] = 0
    mask = torch.zeros(lead + (text_seq_len,), dtype=torch.int64)
    mask[..., :real_len] = 1
    types = torch.zeros(lead + (text_seq_len,), dtype=torch.int64)
    if n_candidates:
        types[..., real_len // 2:real_len] = 1
    return clips, ids, mask, types
