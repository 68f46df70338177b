"""Drop-in replacements for lrce.models.e2e.{E2EOpenEnded, E2EMultipleChoice, E2ECount} (e2e.py:7-106): same class
names, constructor arguments (positional order of train_ddp.py:89-98 and keywords of eval.py:66-74), sub-module
attributes (`text_extractor`, `video_extractor`, `fusion_model` — agent_base.py:30-39 builds optimizer groups from
them), state_dict keys and forward signature. The forward pass runs on liblrce_b200 (sm_100a) and
returns fp32 logits; ambient fp16 autocast (agent_oe.py:28) is ignored by construction because no autocast-able torch
op is on the path except BERT, which is pinned to bf16.

One extension: the keyword-only `pretrained` flag. True (default) reproduces the reference's behaviour — the Swin
checkpoint must exist at ./pretrained_models/... (e2e.py:11) and BERT comes from `from_pretrained`; False builds both
with random weights (no network / checkpoint needed: benchmarks and parity tests).
"""
import os
from typing import Iterable, List

import torch
from torch import nn

from . import ops
from .feature_extractor import SWIN_CKPT, TextExtractor, VideoExtractor
from .fusion import LRCECount, LRCEMultipleChoice, LRCEOpenEnded


class E2EBase(nn.Module):
    def __init__(self, pretrained: bool = True) -> None:
        super().__init__()
        if pretrained:
            assert os.path.exists(SWIN_CKPT)
        self.text_extractor = TextExtractor(pretrained=pretrained)
        self.video_extractor = VideoExtractor(SWIN_CKPT if pretrained else None)
        # The extractors of this path are forward-only kernels (Swin / BERT backward is "next"): freeze
        # them explicitly, so that the reference agent's AdamW groups (agent_base.py:27-41), its L2 term
        # (agent_base.py:103-108 skips parameters that do not require grad) and plain DDP(model) (agent_base.py:76) all
        # skip them cleanly instead of waiting for gradients that never come. The reference fine-tunes all three groups.
        for p in list(self.text_extractor.parameters()) + list(self.video_extractor.parameters()):
            p.requires_grad_(False)

    def extract_text_features(self, texts, attention_mask, texts_type_ids):
        return self.text_extractor(texts, attention_mask, texts_type_ids)

    def extract_video_features(self, video_clips):
        return self.video_extractor(video_clips)

    def forward(self, video_clips, texts, texts_attention_mask, texts_type_ids):
        if not video_clips.is_cuda:
            raise ops._lib.LrceError("E2E forward needs CUDA inputs on a B200: the hot path has no CPU fallback")
        # The two extractors are independent (e2e.py:23-24). Swin's ~170 long kernels are enqueued first on the current
        # stream; BERT's ~200 short library kernels then go to a side stream, so they run inside Swin's shadow instead
        # of in front of it, and their launch overhead is hidden too.
        cur = torch.cuda.current_stream()
        # (per-launch tracing, bench.py's kernel table: everything on one stream so that event intervals do not overlap)
        side = cur if ops.trace is not None else self._side_stream(video_clips.device)
        side.wait_stream(cur)
        # The extractors are forward-only kernels: in a training step (grad enabled) they run without autograd and only
        # the cross-modal encoder is differentiated — the trainable part BASELINE.json's config 5 names.
        with torch.no_grad():
            video_features = self.extract_video_features(video_clips)
            with torch.cuda.stream(side):
                texts_features = self.extract_text_features(texts, texts_attention_mask, texts_type_ids)
        cur.wait_stream(side)
        texts_features.record_stream(cur)
        return self.fusion_model(video_features, texts_features, texts_attention_mask)

    def _side_stream(self, device):
        s = getattr(self, "_side", None)
        if s is None or s.device != device:
            s = self._side = torch.cuda.Stream(device=device)
        return s


class E2EOpenEnded(E2EBase):
    def __init__(self, feature_dim: int, num_classes: int, drop_out_rate: float = 0.1,
                 video_feature_res: Iterable[int] = (7, 7), video_feature_dim: int = 768, frame_sample_size: int = 5,
                 temporal_scale: List[int] = [1, 2, 3], text_seq_len: int = 30, *, pretrained: bool = True) -> None:
        super().__init__(pretrained)
        self.fusion_model = LRCEOpenEnded(feature_dim, num_classes, drop_out_rate, video_feature_res, video_feature_dim,
                                          frame_sample_size, temporal_scale, text_seq_len)


class E2EMultipleChoice(E2EBase):
    def __init__(self, feature_dim: int, num_classes: int, drop_out_rate: float = 0.1,
                 video_feature_res: Iterable[int] = (7, 7), video_feature_dim: int = 768, frame_sample_size: int = 5,
                 temporal_scale: List[int] = [1, 2, 3], text_seq_len: int = 40, *, pretrained: bool = True) -> None:
        super().__init__(pretrained)
        self.fusion_model = LRCEMultipleChoice(feature_dim, num_classes, drop_out_rate, video_feature_res,
                                               video_feature_dim, frame_sample_size, temporal_scale, text_seq_len)

    def extract_text_features(self, texts, attention_mask, texts_type_ids):
        batch_size, total_choice, seq_len = texts.shape  # e2e.py:77-81
        out = self.text_extractor(texts.flatten(0, 1), attention_mask.flatten(0, 1), texts_type_ids.flatten(0, 1))
        return out.view(batch_size, total_choice, seq_len, -1)


class E2ECount(E2EBase):
    def __init__(self, feature_dim: int, num_classes: int = 1, drop_out_rate: float = 0.1,
                 video_feature_res: Iterable[int] = (7, 7), video_feature_dim: int = 768, frame_sample_size: int = 5,
                 temporal_scale: List[int] = [1, 2, 3], text_seq_len: int = 30, *, pretrained: bool = True) -> None:
        super().__init__(pretrained)
        self.fusion_model = LRCECount(feature_dim, num_classes, drop_out_rate, video_feature_res, video_feature_dim,
                                      frame_sample_size, temporal_scale, text_seq_len)


def install():
    """Make `from lrce.models.e2e import E2EOpenEnded, E2EMultipleChoice, E2ECount` (eval.py:5, train_ddp.py) resolve to
    the B200-native classes, so the reference's launchers and agents run unmodified. Call before importing them."""
    import sys

    sys.modules["lrce.models.e2e"] = sys.modules[__name__]
