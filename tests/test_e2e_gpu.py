"""GPU end-to-end parity: the drop-in E2E* modules (CUDA path through the C ABI) against the golden vectors produced by
the unmodified reference on identical seeded weights and inputs (BASELINE.json configs[0], [3]) and, at the full batch
of configs[1], through batch-invariance (every clip is independent).

Stated tolerances (bf16 activations / fp32 accumulation vs the reference's fp32 CPU run): Swin features rel-L2 <= 3e-2,
summary logits max-abs <= 0.25 on logits of std ~4 (answer-head gain 4, oracle/weights.py), top-1 must agree."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import weights as W  # noqa: E402

CFG = dict(feature_dim=768, video_feature_res=[7, 7], video_feature_dim=1024, frame_sample_size=5, temporal_scale=[3])


# tolerances = 1.5 x measured (features 9.0e-3, logit rel-L2 8e-3); see test_msvd_b2_vs_reference for the max-abs bound
FEAT_TOL, LOGIT_TOL, LOGIT_REL_TOL = 1.4e-2, 0.2, 1.2e-2

def build(kind, ncls, L):
    import lrce_b200

    cls = {"oe": lrce_b200.E2EOpenEnded, "mc": lrce_b200.E2EMultipleChoice, "count": lrce_b200.E2ECount}[kind]
    m = cls(num_classes=ncls, text_seq_len=L, pretrained=False, **CFG)
    m.load_state_dict(W.make_e2e_state_dict(ncls, L, 3, seed=0), strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def msvd():
    return build("oe", 1000, 32)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_msvd_b2_vs_reference(golden, msvd):
    g = golden["e2e"]
    clips, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    taps = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):  # the agent's ambient autocast (agent_oe.py:28)
        feats = msvd.video_extractor(clips.cuda(), taps=taps)
        y = msvd(clips.cuda(), ids.cuda(), mask.cuda(), types.cuda())
    assert y.dtype == torch.float32 and y.shape == (2, 1000)
    for k in ("patch_embed", "stage0.out", "stage1.out", "stage2.out"):
        err = rel_l2(taps[k].reshape(-1)[::997], torch.from_numpy(g[f"msvd-qa-oe.{k}.sample"]))
        assert err < FEAT_TOL, (k, err)
    err = rel_l2(feats.reshape(-1)[::997], torch.from_numpy(g["msvd-qa-oe.video_features.sample"]))
    assert err < FEAT_TOL, err  # measured 9.0e-3
    ref = torch.from_numpy(g["msvd-qa-oe.logits"])
    d = (y.cpu() - ref).abs().max().item()
    r = rel_l2(y, ref)
    print("msvd b2: feature rel_l2", err, "logit max abs", d, "rel_l2", r)
    # measured: rel-L2 7.9e-3 .. 8.5e-3; the maximum over 2000 logits of std ~4 is an extreme-value statistic (0.10 .. 0.15
    # between kernel builds of identical rel-L2), so the relative norm carries the tight bound
    assert r < LOGIT_REL_TOL and d < LOGIT_TOL, (r, d)
    assert torch.equal(y.cpu().argmax(-1), ref.argmax(-1))


@pytest.mark.parametrize("name,kind,ncls,L", [("tgif-action", "mc", 1, 40), ("tgif-count", "count", 1, 30)])
def test_mc_count_b2_vs_reference(golden, name, kind, ncls, L):
    m = build(kind, ncls, L)
    clips, ids, mask, types = W.make_inputs(2, 3, L, seed=1, n_candidates=5 if kind == "mc" else 0)
    with torch.no_grad():
        y = m(clips.cuda(), ids.cuda(), mask.cuda(), types.cuda())
    ref = torch.from_numpy(golden["e2e"][f"{name}.logits"])
    assert y.shape == ref.shape
    d = (y.cpu() - ref).abs().max().item()
    print(name, "max abs", d, y.cpu().tolist(), ref.tolist())
    assert d < 0.06, d  # measured 0.025 .. 0.034 on logits of magnitude ~5


def test_msvd_b32_batch_invariance(golden, msvd):
    """configs[1] shape (batch 32, 96 segments): 16 copies of the 2 golden clips must give the golden logits in every
    copy, and identical results across copies (the path has no cross-clip term)."""
    clips, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    rep = lambda t: t.repeat((16,) + (1,) * (t.dim() - 1))
    with torch.no_grad():
        y = msvd(rep(clips).cuda(), rep(ids).cuda(), rep(mask).cuda(), rep(types).cuda())
    ref = torch.from_numpy(golden["e2e"]["msvd-qa-oe.logits"])
    assert y.shape == (32, 1000)
    y = y.cpu().view(16, 2, 1000)
    assert (y - ref[None]).abs().max().item() < LOGIT_TOL and rel_l2(y, ref[None].expand(16, 2, 1000)) < LOGIT_REL_TOL
    assert torch.equal(y.argmax(-1), ref.argmax(-1)[None].expand(16, 2))
    assert (y - y[:1]).abs().max().item() < 1e-3  # same kernels, same data -> same answer wherever the clip sits


def test_training_step_matches_inference_path_and_has_grads():
    """BASELINE config 5 (tgif-frameqa training step): with dropout off, the differentiable encoder path (hand-written
    training kernels behind one autograd node, train.py) must reproduce the inference kernels' logits, and backward must
    reach every encoder parameter and nothing else."""
    import lrce_b200

    m = lrce_b200.E2EOpenEnded(num_classes=1000, text_seq_len=30, drop_out_rate=0.0, pretrained=False, **CFG)
    m.load_state_dict(W.make_e2e_state_dict(1000, 30, 3, seed=0), strict=True)
    m = m.cuda().train()
    clips, ids, mask, types = W.make_inputs(2, 3, 30, seed=1)
    args = (clips.cuda(), ids.cuda(), mask.cuda(), types.cuda())
    with torch.no_grad():
        y_kernel = m(*args)
    y = m(*args)
    assert y.requires_grad and y.shape == y_kernel.shape
    assert (y - y_kernel).abs().max().item() < 0.1, (y - y_kernel).abs().max().item()
    loss = torch.nn.functional.cross_entropy(y, torch.tensor([3, 7], device="cuda"))
    loss.backward()
    enc = list(m.fusion_model.parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc)
    assert sum(p.grad.abs().sum().item() for p in enc) > 0
    assert all(p.grad is None for p in m.video_extractor.parameters())
    assert all(p.grad is None for p in m.text_extractor.parameters())
    # dropout on: a different mask every call, same expectation, finite gradients
    m2 = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [3], 30)
    m2.load_state_dict(W.make_fusion_state_dict(1000, 30, 3, seed=0), strict=True)
    m2 = m2.cuda().train()
    g = torch.Generator().manual_seed(5)
    vf = torch.randn((2, 3, 3, 49, 1024), generator=g).bfloat16().cuda()
    tf = torch.randn((2, 30, 768), generator=g).cuda()
    ya = m2(vf, tf).detach()
    for _ in range(3):  # direct launches, graph capture, graph replay: a fresh mask every time
        yb = m2(vf, tf)
        assert not torch.equal(ya, yb) and torch.isfinite(yb).all()
        for q in m2.parameters():
            q.grad = None
        yb.logsumexp(-1).sum().backward()
        assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in m2.parameters())
        ya = yb.detach()
    # two forwards before a backward: the plan holds ONE forward's activations -> loud error, not silent garbage
    y1 = m2(vf, tf)
    m2(vf, tf)
    with pytest.raises(RuntimeError):
        y1.sum().backward()


def test_text_extractor_native_bert(golden, msvd, monkeypatch):
    """BERT-base on liblrce_b200: against the reference's text features (golden, fp32 CPU), against the
    HuggingFace module on the same GPU, and CUDA-graph replay == direct launches (fresh inputs, second shape, weight update)."""
    te = msvd.text_extractor
    _, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    ids, mask, types = ids.cuda(), mask.cuda(), types.cuda()
    ref = torch.from_numpy(golden["e2e_r2"]["msvd-qa-oe.text_features"])
    with torch.no_grad():
        a = te(ids, mask, types)                      # capture + replay
        assert a.dtype == torch.float32 and a.shape == (2, 32, 768)
        err = rel_l2(a, ref)
        err_hf = rel_l2(te.hf_forward(ids, mask, types), ref)
        print(f"BERT text features vs reference: native rel-L2 {err:.3e} (HF bf16 autocast on the same GPU: {err_hf:.3e})")
        assert err < 1.5e-2, err  # measured 1.03e-2 (the HF module under bf16 autocast on the same GPU: 1.09e-2)
        ids2 = torch.roll(ids, 1, dims=0)
        b = te(ids2, mask, types)                     # replay with new inputs
        monkeypatch.setenv("LRCE_B200_BERT_GRAPH", "0")
        a_ref, b_ref = te(ids, mask, types), te(ids2, mask, types)
        monkeypatch.setenv("LRCE_B200_BERT_GRAPH", "1")
        assert len(te._graphs) >= 1
        assert torch.equal(a, a_ref) and torch.equal(b, b_ref)
        c = te(ids[:1], mask[:1], types[:1])          # another shape: its own graph
        assert (c - a_ref[:1]).abs().max().item() < 2e-2
        # a parameter update re-packs the bf16 weights and re-captures
        w = te.bert.embeddings.word_embeddings.weight
        old = w.detach().clone()
        w.mul_(1.5)
        d = te(ids, mask, types)
        monkeypatch.setenv("LRCE_B200_BERT_GRAPH", "0")
        d_ref = te(ids, mask, types)
        w.copy_(old)
        assert torch.equal(d, d_ref) and not torch.equal(d, a)
        monkeypatch.setenv("LRCE_B200_BERT_GRAPH", "1")
        # padding is really excluded: changing the ids behind the mask must not change the attended positions' features
        ids3 = ids.clone()
        ids3[:, 25:] = 1234
        e = te(ids3, mask, types)
        assert (e[:, :20] - a[:, :20]).abs().max().item() < 1e-5


def test_prefetch_feed_ring_delivers_every_batch():
    """PrefetchFeed (pinned host batches -> two-slot device ring, copy of batch i+1 under the work on batch i)"""
    from lrce_b200.feed import PrefetchFeed

    dev = torch.device("cuda", 0)
    batches = [[torch.full((4, 1024, 1024), float(i)).pin_memory(), torch.arange(8).add(i).pin_memory()] for i in range(7)]
    seen = []
    for x, y in PrefetchFeed(batches, dev):
        assert x.device.type == "cuda" and y.device.type == "cuda"
        z = x.sum() / x.numel()  # consume on the compute stream while the next copy is in flight
        for _ in range(20):
            z = z + (x * 0).sum()
        seen.append((z.item(), y.cpu().tolist()))
    assert [round(v) for v, _ in seen] == list(range(7))
    assert [t for _, t in seen] == [list(range(i, i + 8)) for i in range(7)]


def test_uint8_frames_give_the_fp32_logits(msvd):
    """E2E extension: the module accepts the uint8 frames themselves; logits equal those of the ToTensor()-ed clips."""
    _, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    frames = torch.randint(0, 256, (2, 3, 5, 3, 224, 224), generator=torch.Generator().manual_seed(9), dtype=torch.uint8)
    with torch.no_grad():
        y8 = msvd(frames.cuda(), ids.cuda(), mask.cuda(), types.cuda())
        yf = msvd(frames.float().div(255).cuda(), ids.cuda(), mask.cuda(), types.cuda())
    assert torch.equal(y8, yf)
