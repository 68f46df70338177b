"""CPU: pins oracle/lrce_oracle.py to the golden vectors produced by the unmodified reference (oracle/make_golden.py).
Integer artefacts must be bit-exact; fp32 activations agree to fp32 round-off."""
import numpy as np
import pytest
import torch

import lrce_oracle as O
import weights as W

STAGES = {"s1": (3, 56, 56), "s2": (3, 28, 28), "s3": (3, 14, 14), "s4": (3, 7, 7)}


def seeded(shape, seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(shape, generator=g)


def close(a, b, rtol=2e-4, atol=2e-4):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {err}"


@pytest.mark.parametrize("name", list(STAGES))
def test_window_index_bit_exact(golden, name):
    g = golden["index"]
    dims = STAGES[name]
    win, shift = O.clamp_window(dims, O.CONFIGURED_WINDOW, (4, 3, 3))
    assert tuple(g[f"{name}.window"]) == win and tuple(g[f"{name}.shift"]) == shift
    assert np.array_equal(O.window_gather_index(dims, win, (0, 0, 0)).numpy(), g[f"{name}.gather_plain"])
    assert np.array_equal(O.window_gather_index(dims, win, shift).numpy(), g[f"{name}.gather_shifted"])
    if name != "s4":
        assert np.array_equal(O.merge_gather_index(dims).numpy(), g[f"{name}.merge_gather"])
        m = O.shift_mask(dims, win, shift)
        bits = np.unpackbits(g[f"{name}.mask_bits"])[: m.numel()].reshape(tuple(g[f"{name}.mask_shape"]))
        assert np.array_equal((m != 0).numpy().astype(np.uint8), bits)
        assert set(m.unique().tolist()) <= {0.0, -100.0}
        # only 4 distinct window masks per stage: interior / last column / last row / corner
        assert len({m[i].numpy().tobytes() for i in range(m.shape[0])}) == 4


def test_rel_pos_index_bit_exact(golden):
    idx = O.relative_position_index((3, 7, 7))
    assert np.array_equal(idx.numpy().astype(np.int16), golden["index"]["rel_pos_index_147"])
    assert np.array_equal(W.relative_position_index()[:147, :147].numpy(), idx.numpy())
    # closed form used by the CUDA kernel: idx = f(i) - f(j) + 1267 with f(t) = 169 d + 13 h + w
    t = torch.arange(147)
    f = 169 * (t // 49) + 13 * ((t // 7) % 7) + t % 7
    assert torch.equal(idx, f[:, None] - f[None, :] + 7 * 169 + 6 * 13 + 6)


def test_patch_embed(golden):
    sd = W.make_swin_state_dict(seed=0)
    clips = torch.rand((2, 5, 3, 32, 32), generator=torch.Generator().manual_seed(11))
    close(O.patch_embed(sd, "", clips), golden["swin_modules"]["patch_embed.out"])


@pytest.mark.parametrize("tag,layer,blk,dim,heads,hw", [("s1b0", 0, 0, 128, 4, 14), ("s1b1", 0, 1, 128, 4, 14),
                                                         ("s3b1", 2, 1, 512, 16, 14), ("s4b1", 3, 1, 1024, 32, 7)])
def test_swin_block(golden, tag, layer, blk, dim, heads, hw):
    sd = W.make_swin_state_dict(seed=0)
    x = seeded((1, 3, hw, hw, dim), 100 + layer * 10 + blk)
    y = O.swin_block(sd, f"layers.{layer}.blocks.{blk}.", x, heads, shifted=bool(blk % 2))
    ref = golden["swin_modules"][f"{tag}.out"]
    close(y if dim == 128 else y.reshape(-1)[::7], ref, rtol=5e-4, atol=5e-4)


def test_window_attention_module(golden):
    g = golden["swin_modules"]
    sd = W.make_swin_state_dict(seed=0)
    xw = torch.from_numpy(g["s1b1.attn_in"])
    mask = O.shift_mask((3, 14, 14), (3, 7, 7), (0, 3, 3))
    close(O.window_attention(sd, "layers.0.blocks.1.attn.", xw, 4, (3, 7, 7), mask), g["s1b1.attn_out"])


def test_patch_merging(golden):
    sd = W.make_swin_state_dict(seed=0)
    close(O.patch_merging(sd, "layers.0.downsample.", seeded((1, 3, 14, 14, 128), 200)), golden["swin_modules"]["merge.out"])


@pytest.mark.parametrize("name,kind,ncls,L", [("msvd-qa-oe", "oe", 1000, 32), ("tgif-action", "mc", 1, 40),
                                              ("tgif-count", "count", 1, 30)])
def test_fusion_heads(golden, name, kind, ncls, L):
    g = golden["fusion"]
    sd = W.make_fusion_state_dict(ncls, L, 3, seed=0)
    vf = seeded((2, 3, 3, 49, 1024), 300)
    tf = seeded((2, 5, L, 768) if kind == "mc" else (2, L, 768), 301)
    taps = {}
    y = O.lrce_head(sd, vf, tf, kind, pre="", taps=taps)
    close(y, g[f"{name}.logits"], rtol=1e-3, atol=1e-3)
    toks = torch.stack([taps[f"token.s{s}"] for s in range(3)])
    close(toks, g[f"{name}.tokens"], rtol=1e-3, atol=1e-3)


def test_weight_fingerprint(golden):
    sd = W.make_e2e_state_dict(1000, 32, 3, seed=0)
    fp = np.array([sd[k].double().sum().item() for k in
                   ("video_extractor.swin.layers.2.blocks.7.mlp.fc1.weight",
                    "text_extractor.bert.encoder.layer.3.output.dense.weight", "fusion_model.final_fc.weight")])
    assert np.allclose(fp, golden["e2e"]["fingerprint"], rtol=0, atol=1e-6)
    assert len(sd) == 783


def test_e2e_msvd_b2(golden):
    """BASELINE.json configs[0]: msvd-qa-oe, batch 2, fp32 on CPU — the reference parity run."""
    g = golden["e2e"]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = W.make_e2e_state_dict(1000, 32, 3, seed=0)
    clips, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    taps = {}
    with torch.no_grad():
        y = O.e2e_forward(sd, clips, ids, mask, types, "oe", taps=taps)
    ref = torch.from_numpy(g["msvd-qa-oe.logits"])
    assert torch.equal(y.argmax(-1), ref.argmax(-1))
    close(y, ref, rtol=2e-3, atol=2e-3)
    for k in ("patch_embed", "stage0.out", "stage1.out", "stage2.out", "text_features", "video_features"):
        t = taps[k]
        if k == "stage3.out" or k.startswith("stage"):
            pass
        close(t.reshape(-1)[::997], g[f"msvd-qa-oe.{k}.sample"], rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("name,kind,ncls,L", [("tgif-action", "mc", 1, 40), ("tgif-count", "count", 1, 30)])
def test_e2e_mc_count_b2(golden, name, kind, ncls, L):
    """BASELINE.json configs[3] (multiple choice, 5 candidates) and the counting head, batch 2, fp32 CPU."""
    sd = W.make_e2e_state_dict(ncls, L, 3, seed=0)
    clips, ids, mask, types = W.make_inputs(2, 3, L, seed=1, n_candidates=5 if kind == "mc" else 0)
    with torch.no_grad():
        y = O.e2e_forward(sd, clips, ids, mask, types, kind)
    close(y, golden["e2e"][f"{name}.logits"], rtol=2e-3, atol=2e-3)


# ---------------------------------------------------------------------------------------------------------------------
# round-2 fixtures (oracle/make_golden.py golden_*_r2 / golden_e2e_b32 / golden_grad): the remaining configs/*.json
# shapes, direct pos-embed taps, BERT features, a 32-distinct-clip batch (first clips here, all 32 on the GPU) and the
# gradients of the configs[4] training step
R2 = [("msvd-qa-oe", "oe", 1000, 32, 0), ("msrvtt-qa-oe", "oe", 1500, 37, 0), ("tgif-frameqa", "oe", 1000, 30, 0),
      ("tgif-transition", "mc", 1, 40, 1)]


@pytest.mark.parametrize("name,kind,ncls,L,seed", R2)
def test_fusion_heads_r2(golden, name, kind, ncls, L, seed):
    g = golden["fusion_r2"]
    sd = W.make_fusion_state_dict(ncls, L, 3, seed=seed)
    vf = seeded((2, 3, 3, 49, 1024), 310)
    tf = seeded((2, 5, L, 768) if kind == "mc" else (2, L, 768), 311)
    taps = {}
    y = O.lrce_head(sd, vf, tf, kind, pre="", taps=taps)
    close(y, g[f"{name}.logits"], rtol=1e-3, atol=1e-3)
    close(torch.stack([taps[f"token.s{s}"] for s in range(3)]), g[f"{name}.tokens"], rtol=1e-3, atol=1e-3)
    if name == "msvd-qa-oe":  # VideoPosEmbed / TextPosEmbed outputs (embedding.py:47-63, :17-23) asserted directly
        close(taps["text_embedded"], g[f"{name}.text_embedded"], rtol=1e-4, atol=1e-4)
        close(taps["video_embedded"], torch.from_numpy(g[f"{name}.video_embedded"].astype(np.float32)), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("name,kind,ncls,L,in_seed", [("msrvtt-qa-oe", "oe", 1500, 37, 1), ("tgif-frameqa", "oe", 1000, 30, 1),
                                                      ("tgif-transition", "mc", 1, 40, 3)])
def test_e2e_r2_b2(golden, name, kind, ncls, L, in_seed):
    """BASELINE.json configs[2] (msrvtt-qa-oe), configs[4]'s model (tgif-frameqa) and tgif-transition, batch 2, fp32 CPU."""
    g = golden["e2e_r2"]
    sd = W.make_e2e_state_dict(ncls, L, 3, seed=0)
    clips, ids, mask, types = W.make_inputs(2, 3, L, seed=in_seed, n_candidates=5 if kind == "mc" else 0)
    taps = {}
    with torch.no_grad():
        y = O.e2e_forward(sd, clips, ids, mask, types, kind, taps=taps)
    close(y, g[f"{name}.logits"], rtol=2e-3, atol=2e-3)
    close(taps["text_features"].reshape(-1)[::997], g[f"{name}.text_features.sample"], rtol=2e-3, atol=2e-3)


def test_bert_features_full(golden):
    sd = W.make_e2e_state_dict(1000, 32, 3, seed=0)
    _, ids, mask, types = W.make_inputs(2, 3, 32, seed=1)
    with torch.no_grad():
        t = O.bert_forward(sd, ids, mask, types)
    close(t, golden["e2e_r2"]["msvd-qa-oe.text_features"], rtol=2e-3, atol=2e-3)


def test_e2e_b32_first_clips(golden):
    """the first 4 of the 32 distinct clips of tests/golden/e2e_b32.npz (the GPU suite checks all 32)"""
    sd = W.make_e2e_state_dict(1000, 32, 3, seed=0)
    clips, ids, mask, types = W.make_inputs(32, 3, 32, seed=2)
    with torch.no_grad():
        y = O.e2e_forward(sd, clips[:4], ids[:4], mask[:4], types[:4], "oe")
    close(y, golden["e2e_b32"]["msvd-qa-oe.logits"][:4], rtol=2e-3, atol=2e-3)


def test_training_step_gradients(golden):
    """configs[4]: gradients of the cross-entropy loss w.r.t. every encoder parameter, oracle autograd vs the reference
    LRCEOpenEnded with drop_out_rate=0 (oracle/make_golden.py golden_grad)."""
    g = golden["grad"]
    sd = {k: v.clone().requires_grad_(True) for k, v in W.make_fusion_state_dict(1000, 30, 3, seed=0).items()}
    y = O.lrce_head(sd, seeded((4, 3, 3, 49, 1024), 320), seeded((4, 30, 768), 321), "oe", pre="")
    loss = torch.nn.functional.cross_entropy(y, torch.from_numpy(g["target"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"][0])) < 1e-3
    close(y.detach(), g["logits"], rtol=1e-3, atol=1e-3)
    for name, norm in zip(g["names"].tolist(), g["norms"].tolist()):
        gr = sd[name].grad
        if name.endswith(("self_attn.in_proj_weight", "self_attn.in_proj_bias")):
            # length-1 self-attention: q and k rows receive exactly zero gradient (softmax over one key)
            assert gr is None or gr[: 2 * 768].abs().max().item() == 0.0
        got = torch.zeros_like(sd[name]) if gr is None else gr
        ref = torch.from_numpy(g["g." + name])
        samp = got.reshape(-1)[::1999]
        assert (samp - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()) + 1e-5, name
        assert abs(got.double().norm().item() - norm) <= 2e-3 * max(norm, 1e-3) + 1e-5, name
