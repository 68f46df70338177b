"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_harness.py) on seeded synthetic weights (oracle/weights.py) and inputs.

Run in the build container only:   python oracle/make_golden.py
The fixtures are small (inputs are re-derivable from seeds and are not stored; large activations are stored as strided
samples plus moments) and are what pins oracle/lrce_oracle.py — and through it the CUDA path — to the reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
import weights  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
SAMPLE_STRIDE = 997  # prime stride for activation samples

CONFIGS = {
    "msvd-qa-oe": dict(kind="oe", num_classes=1000, text_seq_len=32),
    "msrvtt-qa-oe": dict(kind="oe", num_classes=1500, text_seq_len=37),
    "tgif-frameqa": dict(kind="oe", num_classes=1000, text_seq_len=30),
    "tgif-action": dict(kind="mc", num_classes=1, text_seq_len=40),
    "tgif-transition": dict(kind="mc", num_classes=1, text_seq_len=40),
    "tgif-count": dict(kind="count", num_classes=1, text_seq_len=30),
}


def model_cfg(c):
    return dict(feature_dim=768, num_classes=c["num_classes"], video_feature_res=[7, 7], video_feature_dim=1024,
                frame_sample_size=5, temporal_scale=[3], text_seq_len=c["text_seq_len"])


def sample(t):
    f = t.detach().reshape(-1).to(torch.float32)
    return f[::SAMPLE_STRIDE].numpy().copy()


def moments(t):
    f = t.detach().double()
    return np.array([f.mean().item(), f.abs().mean().item(), f.pow(2).mean().sqrt().item(), f.abs().max().item()])


def seeded(shape, seed, scale=1.0):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def golden_index(ref):
    """integer artefacts: window gather tables, shift masks, relative-position index, patch-merging gather."""
    out = {}
    sw = ref.swin
    for name, dims in (("s1", (3, 56, 56)), ("s2", (3, 28, 28)), ("s3", (3, 14, 14)), ("s4", (3, 7, 7))):
        win, shift = sw.get_window_size(dims, (8, 7, 7), (4, 3, 3))
        n = dims[0] * dims[1] * dims[2]
        ids = torch.arange(n, dtype=torch.float32).view(1, *dims, 1)
        plain = sw.window_partition(ids, win).squeeze(-1)
        rolled = sw.window_partition(torch.roll(ids, shifts=tuple(-s for s in shift), dims=(1, 2, 3)), win).squeeze(-1)
        out[f"{name}.window"] = np.array(win, dtype=np.int32)
        out[f"{name}.shift"] = np.array(shift, dtype=np.int32)
        out[f"{name}.gather_plain"] = plain.to(torch.int32).numpy()
        out[f"{name}.gather_shifted"] = rolled.to(torch.int32).numpy()
        # inverse: window_reverse + roll(+shift) of the rolled table must give back arange
        back = sw.window_reverse(rolled.view(-1, *win, 1), win, 1, *dims)
        back = torch.roll(back, shifts=tuple(shift), dims=(1, 2, 3))
        assert torch.equal(back.reshape(-1), torch.arange(n, dtype=torch.float32))
        if any(s > 0 for s in shift):
            m = sw.compute_mask(dims[0], dims[1], dims[2], win, shift, torch.device("cpu"))
            assert set(m.unique().tolist()) <= {0.0, -100.0}
            out[f"{name}.mask_bits"] = np.packbits((m != 0).numpy().reshape(-1))
            out[f"{name}.mask_shape"] = np.array(m.shape, dtype=np.int32)
        pm = sw.PatchMerging(1)
        pm.norm, pm.reduction = torch.nn.Identity(), torch.nn.Identity()
        if name != "s4":
            out[f"{name}.merge_gather"] = pm(ids).reshape(-1, 4).to(torch.int32).numpy()
    attn = sw.WindowAttention3D(32, (8, 7, 7), 1)
    out["rel_pos_index_147"] = attn.relative_position_index[:147, :147].to(torch.int16).numpy()
    np.savez_compressed(os.path.join(OUT, "index.npz"), **out)
    print("index.npz", sum(v.nbytes for v in out.values()))


def golden_swin_modules(ref):
    """module-level activations on small shapes with the seeded Swin weights (stage-1 and stage-3 blocks)."""
    sw = ref.swin
    sd = weights.make_swin_state_dict(seed=0)
    out = {}

    def load(mod, prefix):
        mod.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)
        return mod.eval()

    with torch.no_grad():
        # PatchEmbed3D through the same pre-processing video.py:35-37 applies, small spatial size
        import torchvision as TV
        clips = torch.rand((2, 5, 3, 32, 32), generator=torch.Generator().manual_seed(11))
        pe = load(sw.PatchEmbed3D(patch_size=(2, 4, 4), in_chans=3, embed_dim=128, norm_layer=torch.nn.LayerNorm),
                  "patch_embed.")
        y = pe(TV.transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(clips).transpose(1, 2))
        out["patch_embed.out"] = y.permute(0, 2, 3, 4, 1).contiguous().numpy()  # channels-last

        for tag, layer, blk, dim, heads, hw in (("s1b0", 0, 0, 128, 4, 14), ("s1b1", 0, 1, 128, 4, 14),
                                                ("s3b1", 2, 1, 512, 16, 14), ("s4b1", 3, 1, 1024, 32, 7)):
            shift = (0, 0, 0) if blk % 2 == 0 else (4, 3, 3)
            b = load(sw.SwinTransformerBlock3D(dim=dim, num_heads=heads, window_size=(8, 7, 7), shift_size=shift,
                                               qkv_bias=True), f"layers.{layer}.blocks.{blk}.")
            x = seeded((1, 3, hw, hw, dim), 100 + layer * 10 + blk)
            win, sh = sw.get_window_size((3, hw, hw), (8, 7, 7), (4, 3, 3))
            mask = sw.compute_mask(3, hw, hw, win, sh, torch.device("cpu"))
            y = b(x, mask)
            # wide stages are stored as a stride-7 sample to keep the fixture small
            out[f"{tag}.out"] = y.numpy() if dim == 128 else y.reshape(-1)[::7].numpy().copy()
            # attention module alone (after norm1 + roll + partition) for the kernel-level test
            if tag == "s1b1":
                xw = sw.window_partition(torch.roll(b.norm1(x), shifts=(-sh[0], -sh[1], -sh[2]), dims=(1, 2, 3)), win)
                out[f"{tag}.attn_in"] = xw.numpy()
                out[f"{tag}.attn_out"] = b.attn(xw, mask=mask).numpy()
        pm = load(sw.PatchMerging(128), "layers.0.downsample.")
        x = seeded((1, 3, 14, 14, 128), 200)
        out["merge.out"] = pm(x).numpy()
    np.savez_compressed(os.path.join(OUT, "swin_modules.npz"), **out)
    print("swin_modules.npz", sum(v.nbytes for v in out.values()))


def golden_fusion(ref):
    """LRCE head variants on seeded features (no Swin/BERT): logits + per-segment summarisation tokens."""
    out = {}
    with torch.no_grad():
        for name in ("msvd-qa-oe", "tgif-action", "tgif-count"):
            c = CONFIGS[name]
            cls = {"oe": ref.fusionv3.LRCEOpenEnded, "mc": ref.fusionv3.LRCEMultipleChoice,
                   "count": ref.fusionv3.LRCECount}[c["kind"]]
            m = cls(768, c["num_classes"], 0.1, [7, 7], 1024, 5, [3], c["text_seq_len"])
            sd = weights.make_fusion_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=0)
            m.load_state_dict(sd, strict=True)
            m.eval()
            B = 2
            vf = seeded((B, 3, 3, 49, 1024), 300)
            tshape = (B, 5, c["text_seq_len"], 768) if c["kind"] == "mc" else (B, c["text_seq_len"], 768)
            tf = seeded(tshape, 301)
            mask = torch.ones(tshape[:-1], dtype=torch.int64)
            toks = []
            h = m.fusion_transformer.fusion_layer_norm.register_forward_hook(lambda mod, i, o: toks.append(o.clone()))
            y = m(vf, tf, mask)
            h.remove()
            out[f"{name}.logits"] = y.numpy()
            out[f"{name}.tokens"] = torch.stack(toks).numpy()
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), **out)
    print("fusion.npz", sum(v.nbytes for v in out.values()))


def golden_e2e(ref):
    """whole E2E forward, B=2 (BASELINE.json configs[0]) for msvd-qa-oe, plus MC and Count variants."""
    out = {}
    with torch.no_grad():
        for name in ("msvd-qa-oe", "tgif-action", "tgif-count"):
            c = CONFIGS[name]
            sd = weights.make_e2e_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=0)
            m = ref_harness.build_reference_e2e(c["kind"], model_cfg(c), sd)
            clips, ids, mask, types = weights.make_inputs(2, 3, c["text_seq_len"], seed=1,
                                                          n_candidates=5 if c["kind"] == "mc" else 0)
            taps = {}
            hooks = []
            if name == "msvd-qa-oe":
                swin = m.video_extractor.swin
                hooks.append(swin.patch_embed.register_forward_hook(
                    lambda mod, i, o: taps.setdefault("patch_embed", []).append(o.permute(0, 2, 3, 4, 1))))
                for li, layer in enumerate(swin.layers):
                    hooks.append(layer.register_forward_hook(
                        lambda mod, i, o, li=li: taps.setdefault(f"stage{li}.out", []).append(o.permute(0, 2, 3, 4, 1))))
                hooks.append(m.text_extractor.register_forward_hook(
                    lambda mod, i, o: taps.setdefault("text_features", []).append(o)))
                hooks.append(m.video_extractor.register_forward_hook(
                    lambda mod, i, o: taps.setdefault("video_features", []).append(o)))
            y = m(clips, ids, mask, types)
            for h in hooks:
                h.remove()
            out[f"{name}.logits"] = y.numpy()
            for k, lst in taps.items():
                # the reference runs the S segments as separate Swin calls of batch B (video.py:33): re-order to (B,S)
                t = torch.stack(lst, dim=1).flatten(0, 1) if len(lst) > 1 else lst[0]
                out[f"{name}.{k}.sample"] = sample(t)
                out[f"{name}.{k}.moments"] = moments(t)
            print(name, "logits", tuple(y.shape), "argmax", y.argmax(-1).tolist() if y.dim() > 1 else y.tolist())
            del m
    # weight fingerprint: detects RNG drift between the machine that wrote the fixtures and the one that tests
    sd = weights.make_e2e_state_dict(1000, 32, 3, seed=0)
    out["fingerprint"] = np.array([sd[k].double().sum().item() for k in
                                   ("video_extractor.swin.layers.2.blocks.7.mlp.fc1.weight",
                                    "text_extractor.bert.encoder.layer.3.output.dense.weight",
                                    "fusion_model.final_fc.weight")])
    np.savez_compressed(os.path.join(OUT, "e2e.npz"), **out)
    print("e2e.npz", sum(v.nbytes for v in out.values()))


# ---------------------------------------------------------------------------------------------------------------------
# round 2: the remaining configs/*.json shapes, a 32-distinct-clip batch, direct pos-embed taps, BERT features, gradients
def golden_fusion_r2(ref):
    """LRCE heads for msrvtt-qa-oe / tgif-frameqa / tgif-transition, plus the VideoPosEmbed / TextPosEmbed outputs
    (embedding.py:47-63, :17-23) of the msvd-qa-oe head as direct taps."""
    out = {}
    with torch.no_grad():
        for name in ("msvd-qa-oe", "msrvtt-qa-oe", "tgif-frameqa", "tgif-transition"):
            c = CONFIGS[name]
            cls = {"oe": ref.fusionv3.LRCEOpenEnded, "mc": ref.fusionv3.LRCEMultipleChoice}[c["kind"]]
            m = cls(768, c["num_classes"], 0.1, [7, 7], 1024, 5, [3], c["text_seq_len"])
            # transition shares tgif-action's shapes: a different weight seed makes it a distinct case
            seed = 1 if name == "tgif-transition" else 0
            m.load_state_dict(weights.make_fusion_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=seed), strict=True)
            m.eval()
            B = 2
            vf = seeded((B, 3, 3, 49, 1024), 310)
            tshape = (B, 5, c["text_seq_len"], 768) if c["kind"] == "mc" else (B, c["text_seq_len"], 768)
            tf = seeded(tshape, 311)
            taps, toks, hooks = {}, [], []
            hooks.append(m.fusion_transformer.fusion_layer_norm.register_forward_hook(lambda mod, i, o: toks.append(o.clone())))
            if name == "msvd-qa-oe":
                hooks.append(m.video_pos_embed.register_forward_hook(lambda mod, i, o: taps.__setitem__("video_embedded", o.clone())))
                hooks.append(m.question_pos_embed.register_forward_hook(lambda mod, i, o: taps.__setitem__("text_embedded", o.clone())))
            y = m(vf, tf, torch.ones(tshape[:-1], dtype=torch.int64))
            for h in hooks:
                h.remove()
            out[f"{name}.logits"] = y.numpy()
            out[f"{name}.tokens"] = torch.stack(toks).numpy()
            for k, v in taps.items():
                out[f"{name}.{k}"] = v.numpy().astype(np.float16) if k == "video_embedded" else v.numpy()
    np.savez_compressed(os.path.join(OUT, "fusion_r2.npz"), **out)
    print("fusion_r2.npz", sum(v.nbytes for v in out.values()))


def golden_e2e_r2(ref):
    """whole E2E forward, B=2, for msrvtt-qa-oe (configs[2]), tgif-frameqa (configs[4]) and tgif-transition; BERT's
    last_hidden_state (text.py:11-17) in full for msvd-qa-oe and as a sample for the multiple-choice shape."""
    out = {}
    with torch.no_grad():
        for name in ("msvd-qa-oe", "msrvtt-qa-oe", "tgif-frameqa", "tgif-transition"):
            c = CONFIGS[name]
            sd = weights.make_e2e_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=0)
            m = ref_harness.build_reference_e2e(c["kind"], model_cfg(c), sd)
            n_cand = 5 if c["kind"] == "mc" else 0
            clips, ids, mask, types = weights.make_inputs(2, 3, c["text_seq_len"], seed=3 if name == "tgif-transition" else 1,
                                                          n_candidates=n_cand)
            if name == "msvd-qa-oe":
                out[f"{name}.text_features"] = m.text_extractor(ids, mask, types).numpy()
                del m
                continue
            text = []
            h = m.text_extractor.register_forward_hook(lambda mod, i, o: text.append(o))
            y = m(clips, ids, mask, types)
            h.remove()
            out[f"{name}.logits"] = y.numpy()
            out[f"{name}.text_features.sample"] = sample(text[0])
            print(name, "logits", tuple(y.shape), "argmax", y.argmax(-1).tolist())
            del m
    np.savez_compressed(os.path.join(OUT, "e2e_r2.npz"), **out)
    print("e2e_r2.npz", sum(v.nbytes for v in out.values()))


def golden_e2e_b32(ref):
    """configs[1]: 32 DISTINCT clips + questions through the reference E2EOpenEnded (fp32, CPU): logits for the top-1
    agreement figure over 32 clips."""
    c = CONFIGS["msvd-qa-oe"]
    sd = weights.make_e2e_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=0)
    m = ref_harness.build_reference_e2e(c["kind"], model_cfg(c), sd)
    clips, ids, mask, types = weights.make_inputs(32, 3, c["text_seq_len"], seed=2)
    ys, fs = [], []
    h = m.video_extractor.register_forward_hook(lambda mod, i, o: fs.append(o))
    with torch.no_grad():
        for b0 in range(0, 32, 4):
            ys.append(m(clips[b0:b0 + 4], ids[b0:b0 + 4], mask[b0:b0 + 4], types[b0:b0 + 4]))
    h.remove()
    y, f = torch.cat(ys), torch.cat(fs)
    out = {"msvd-qa-oe.logits": y.numpy(), "msvd-qa-oe.video_features.sample": sample(f),
           "msvd-qa-oe.video_features.moments": moments(f)}
    print("b32 argmax", y.argmax(-1).tolist(), "distinct", len(set(y.argmax(-1).tolist())))
    top2 = y.topk(2, dim=-1).values
    print("top1-top2 margins: min %.3f median %.3f" % ((top2[:, 0] - top2[:, 1]).min().item(), (top2[:, 0] - top2[:, 1]).median().item()))
    np.savez_compressed(os.path.join(OUT, "e2e_b32.npz"), **out)
    print("e2e_b32.npz", sum(v.nbytes for v in out.values()))


GRAD_STRIDE = 1999


def golden_grad(ref):
    """configs[4] (tgif-frameqa) training step of the reference LRCEOpenEnded with drop_out_rate=0 (dropout RNG cannot be
    matched): cross-entropy loss, backward; per-parameter gradient norms + strided samples."""
    c = CONFIGS["tgif-frameqa"]
    m = ref.fusionv3.LRCEOpenEnded(768, c["num_classes"], 0.0, [7, 7], 1024, 5, [3], c["text_seq_len"])
    m.load_state_dict(weights.make_fusion_state_dict(c["num_classes"], c["text_seq_len"], 3, seed=0), strict=True)
    m.train()
    B = 4
    vf = seeded((B, 3, 3, 49, 1024), 320)
    tf = seeded((B, c["text_seq_len"], 768), 321)
    target = torch.tensor([3, 7, 11, 500])
    y = m(vf, tf, torch.ones((B, c["text_seq_len"]), dtype=torch.int64))
    loss = torch.nn.functional.cross_entropy(y, target)
    loss.backward()
    out = {"logits": y.detach().numpy(), "loss": np.array([loss.item()]), "target": target.numpy()}
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(p.grad.double().norm().item())
        out["g." + k] = p.grad.reshape(-1)[::GRAD_STRIDE].numpy().copy()
    out["names"] = np.array(names)
    out["norms"] = np.array(norms)
    print("grad golden: loss %.4f, %d params, total grad norm %.4f" % (loss.item(), len(names), float(np.sqrt((np.array(norms) ** 2).sum()))))
    np.savez_compressed(os.path.join(OUT, "grad.npz"), **out)
    print("grad.npz", sum(v.nbytes for v in out.values()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_harness.import_reference()
    which = sys.argv[1:] or ["index", "swin", "fusion", "e2e"]
    if "index" in which:
        golden_index(ref)
    if "swin" in which:
        golden_swin_modules(ref)
    if "fusion" in which:
        golden_fusion(ref)
    if "e2e" in which:
        golden_e2e(ref)
    for key, fn in (("fusion_r2", golden_fusion_r2), ("e2e_r2", golden_e2e_r2), ("e2e_b32", golden_e2e_b32), ("grad", golden_grad)):
        if key in which or not sys.argv[1:]:
            fn(ref)
