"""GPU parity, round 2: the configs the first round left untested (msrvtt-qa-oe = BASELINE configs[2], tgif-frameqa =
configs[4]'s model, tgif-transition), 32 DISTINCT clips at configs[1]'s batch size with top-1 agreement reported over all
32, the VideoPosEmbed / TextPosEmbed kernels asserted directly, and the stated tolerances.

Tolerances (bf16 activations + fp32 accumulation against the reference's fp32 CPU run, fixtures from
oracle/make_golden.py): Swin features rel-L2 <= 1.4e-2 (SURVEY.md 8d anchor: the reference's own bf16-autocast run differs
from its fp32 run by 1.4e-2), answer logits max-abs <= 0.2 and rel-L2 <= 1.2e-2 on logits of std ~4 (measured over 32 clips x
1000 classes: 0.134 / 7.9e-3; the bounds are 1.5 x measured), encoder-only heads max-abs <= 0.15 (measured 0.114). Top-1 must agree wherever the reference's own top-1/top-2 margin exceeds twice the
logit tolerance; agreement over all clips is printed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import weights as W  # noqa: E402

CFG = dict(feature_dim=768, video_feature_res=[7, 7], video_feature_dim=1024, frame_sample_size=5, temporal_scale=[3])
FEAT_TOL, LOGIT_TOL, LOGIT_REL_TOL, HEAD_TOL = 1.4e-2, 0.2, 1.2e-2, 0.15


def seeded(shape, seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(shape, generator=g)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def build_e2e(kind, ncls, L):
    import lrce_b200

    cls = {"oe": lrce_b200.E2EOpenEnded, "mc": lrce_b200.E2EMultipleChoice, "count": lrce_b200.E2ECount}[kind]
    m = cls(num_classes=ncls, text_seq_len=L, pretrained=False, **CFG)
    m.load_state_dict(W.make_e2e_state_dict(ncls, L, 3, seed=0), strict=True)
    return m.cuda().eval()


def top1_report(y, ref, tol):
    """(agreements, decided, n): `decided` = clips whose reference top-1/top-2 margin exceeds 2*tol — those must agree"""
    top2 = ref.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    agree = y.argmax(-1) == ref.argmax(-1)
    decided = margin > 2 * tol
    assert bool(agree[decided].all()), (y.argmax(-1).tolist(), ref.argmax(-1).tolist(), margin.tolist())
    return int(agree.sum()), int(decided.sum()), y.shape[0]


def test_msvd_b32_distinct_clips_vs_reference(golden):
    """BASELINE configs[1] at full size: 32 distinct clips + questions, against the reference's fp32 logits."""
    g = golden["e2e_b32"]
    m = build_e2e("oe", 1000, 32)
    clips, ids, mask, types = W.make_inputs(32, 3, 32, seed=2)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):  # the agent's ambient autocast (agent_oe.py:28)
        feats = m.video_extractor(clips.cuda())
        y = m(clips.cuda(), ids.cuda(), mask.cuda(), types.cuda()).cpu()
    ref = torch.from_numpy(g["msvd-qa-oe.logits"])
    ferr = rel_l2(feats.reshape(-1)[::997], torch.from_numpy(g["msvd-qa-oe.video_features.sample"]))
    d, r = (y - ref).abs().max().item(), rel_l2(y, ref)
    agree, decided, n = top1_report(y, ref, LOGIT_TOL)
    print(f"msvd b32 distinct: features rel-L2 {ferr:.3e}, logits max-abs {d:.3e} rel-L2 {r:.3e}, "
          f"top-1 agreement {agree}/{n} ({decided} clips with a reference margin > {2 * LOGIT_TOL})")
    assert ferr < FEAT_TOL and d < LOGIT_TOL and r < LOGIT_REL_TOL, (ferr, d, r)
    assert agree >= n - 3  # near-ties excepted (reference margins down to 0.003)


@pytest.mark.parametrize("name,kind,ncls,L,in_seed", [("msrvtt-qa-oe", "oe", 1500, 37, 1), ("tgif-frameqa", "oe", 1000, 30, 1),
                                                      ("tgif-transition", "mc", 1, 40, 3)])
def test_e2e_r2_b2_vs_reference(golden, name, kind, ncls, L, in_seed):
    g = golden["e2e_r2"]
    m = build_e2e(kind, ncls, L)
    clips, ids, mask, types = W.make_inputs(2, 3, L, seed=in_seed, n_candidates=5 if kind == "mc" else 0)
    with torch.no_grad():
        y = m(clips.cuda(), ids.cuda(), mask.cuda(), types.cuda()).cpu()
    ref = torch.from_numpy(g[f"{name}.logits"])
    assert y.shape == ref.shape
    d, r = (y - ref).abs().max().item(), rel_l2(y, ref)
    print(name, "logits max-abs", d, "rel-L2", r)
    assert d < LOGIT_TOL and r < LOGIT_REL_TOL, (d, r)
    top1_report(y, ref, LOGIT_TOL)


@pytest.mark.parametrize("name,kind,ncls,L,seed", [("msvd-qa-oe", "oe", 1000, 32, 0), ("msrvtt-qa-oe", "oe", 1500, 37, 0),
                                                   ("tgif-frameqa", "oe", 1000, 30, 0), ("tgif-transition", "mc", 1, 40, 1)])
def test_fusion_heads_r2_vs_reference(golden, name, kind, ncls, L, seed):
    import lrce_b200

    g = golden["fusion_r2"]
    cls = {"oe": lrce_b200.LRCEOpenEnded, "mc": lrce_b200.LRCEMultipleChoice}[kind]
    m = cls(768, ncls, 0.1, [7, 7], 1024, 5, [3], L)
    m.load_state_dict(W.make_fusion_state_dict(ncls, L, 3, seed=seed), strict=True)
    m = m.cuda().eval()
    vf = seeded((2, 3, 3, 49, 1024), 310)
    tf = seeded((2, 5, L, 768) if kind == "mc" else (2, L, 768), 311)
    taps = {}
    with torch.no_grad():
        y = m(vf.bfloat16().cuda(), tf.cuda(), None, taps=taps).cpu()
    ref = torch.from_numpy(g[f"{name}.logits"])
    toks = torch.stack([taps[f"token.s{s}"] for s in range(3)]).cpu()
    terr = rel_l2(toks, torch.from_numpy(g[f"{name}.tokens"]).view(toks.shape))
    d = (y - ref).abs().max().item()
    print(name, "tokens rel-L2", terr, "logits max-abs", d, "rel-L2", rel_l2(y, ref))
    assert terr < 1.2e-2 and d < HEAD_TOL, (terr, d)
    if name == "msvd-qa-oe":
        # lrce_video_posembed_ln / lrce_text_posembed_ln outputs against embedding.py:47-63 / :17-23 of the reference. The
        # video rows see a bf16 projection GEMM first (inputs rounded to bf16): 2^-8 relative per element -> rel-L2 <= 6e-3
        ve = rel_l2(taps["video_embedded"].view(2, 3, 150, 768), torch.from_numpy(g[f"{name}.video_embedded"].astype(np.float32)))
        te = rel_l2(taps["text_embedded"], torch.from_numpy(g[f"{name}.text_embedded"]))
        print("video_embedded rel-L2", ve, "text_embedded rel-L2", te)
        assert ve < 6e-3 and te < 4e-3, (ve, te)  # outputs are bf16: 2^-9 relative rounding alone is 2.3e-3 rms


def test_posembed_rejects_mismatched_shapes():
    """a dataset/model mismatch in text_seq_len, temporal_scale or resolution must raise, not read out of bounds."""
    import lrce_b200

    m = lrce_b200.LRCEOpenEnded(768, 10, 0.1, [7, 7], 1024, 5, [3], 32).cuda().eval()
    vf = torch.zeros((1, 3, 3, 49, 1024), dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        m(vf, torch.zeros((1, 32, 768), device="cuda"), None)
        for bad_v, bad_t in ((vf[:, :, :, :48], (1, 32, 768)), (vf.expand(1, 3, 3, 49, 1024).repeat(1, 2, 1, 1, 1), (1, 32, 768)),
                             (vf, (1, 33, 768)), (vf, (1, 31, 768))):
            with pytest.raises(lrce_b200.LrceError):
                m(bad_v.contiguous(), torch.zeros(bad_t, device="cuda"), None)


def test_training_step_gradients_vs_reference(golden):
    """BASELINE configs[4]: gradients of the cross-entropy loss w.r.t. every encoder parameter from the hand-written
    forward + backward kernels (train.py) against the reference LRCEOpenEnded with drop_out_rate=0 (fp32 CPU autograd,
    oracle/make_golden.py golden_grad). bf16 operands / fp32 accumulation: per-parameter norms within 4 %, relative L2 error of the
    sampled entries of every tensor within 5 %, loss within 5e-2."""
    import lrce_b200

    g = golden["grad"]
    m = lrce_b200.LRCEOpenEnded(768, 1000, 0.0, [7, 7], 1024, 5, [3], 30)
    m.load_state_dict(W.make_fusion_state_dict(1000, 30, 3, seed=0), strict=True)
    m = m.cuda().train()
    vf, tf = seeded((4, 3, 3, 49, 1024), 320), seeded((4, 30, 768), 321)
    y = m(vf.bfloat16().cuda(), tf.cuda(), None)
    loss = torch.nn.functional.cross_entropy(y, torch.from_numpy(g["target"]).cuda())
    loss.backward()
    torch.cuda.synchronize()
    print("train fwd: logits max-abs", (y.detach().cpu() - torch.from_numpy(g["logits"])).abs().max().item(),
          "loss", loss.item(), "ref", float(g["loss"][0]))
    assert abs(loss.item() - float(g["loss"][0])) < 5e-2
    params = dict(m.named_parameters())
    worst_norm, worst_samp, worst = 0.0, 0.0, None
    tot_ref = float(np.sqrt((g["norms"] ** 2).sum()))
    sq = 0.0
    for name, norm in zip(g["names"].tolist(), g["norms"].tolist()):
        gr = params[name].grad
        assert gr is not None and torch.isfinite(gr).all(), name
        gr = gr.float().cpu()
        ref = torch.from_numpy(g["g." + name])
        samp = gr.reshape(-1)[::1999]
        # sampled entries: relative L2 error over the stride-1999 sample (tensors with fewer than 8 samples are covered by the
        # norm check alone; gradients are heavy-tailed, so an entry-wise bound relative to the RMS would be meaningless)
        e_s = ((samp - ref).norm() / ref.norm().clamp_min(1e-6 * tot_ref)).item() if ref.numel() >= 8 else 0.0
        e_n = abs(gr.double().norm().item() - norm) / max(norm, 1e-3 * tot_ref / 233 ** 0.5)
        sq += ((samp - ref) ** 2).sum().item()
        if max(e_s, e_n) > max(worst_norm, worst_samp):
            worst = name
        worst_norm, worst_samp = max(worst_norm, e_n), max(worst_samp, e_s)
        assert e_n < 0.04 and e_s < 0.05, (name, e_n, e_s)
    print(f"gradient parity over 233 tensors: worst norm error {worst_norm:.3e}, worst sampled-entry error {worst_samp:.3e} ({worst})")


def test_bert_mc_shape_vs_reference(golden):
    """the multiple-choice text shape (5 candidates x L = 40 -> 10 sequences): native BERT features vs the reference sample"""
    import lrce_b200

    te = lrce_b200.TextExtractor(pretrained=False)
    pre = "text_extractor.bert."  # oracle/weights.py seeds every tensor by its FULL key: keep the E2E prefix while drawing
    te.bert.load_state_dict({k[len(pre):]: v for k, v in W.make_bert_state_dict(0, pre).items()}, strict=True)
    te = te.cuda().eval()
    _, ids, mask, types = W.make_inputs(2, 3, 40, seed=3, n_candidates=5)
    with torch.no_grad():
        t = te(ids.flatten(0, 1).cuda(), mask.flatten(0, 1).cuda(), types.flatten(0, 1).cuda())
    err = rel_l2(t.reshape(-1)[::997], torch.from_numpy(golden["e2e_r2"]["tgif-transition.text_features.sample"]))
    print("BERT MC-shape features rel-L2", err)
    assert err < 1.5e-2, err
