"""GPU probe: where the cluster-sharded token walk (csrc/encoder_walk.cu) spends its dependency chain. Runs the encoder once
through lrce_encoder_walk_profile (per-call instrumentation: cycles thread 0 of every CTA spent per sub-step) and prints the
mean per layer-step of CTA 0 next to the slowest CTA, for several cluster counts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200
from lrce_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "oe"
S = 3
if kind == "mc":
    m = lrce_b200.LRCEMultipleChoice(768, 1, 0.1, [7, 7], 1024, 5, [S], 40).cuda().eval()
    tf = torch.randn(B, 5, 40, 768, device="cuda")
else:
    m = lrce_b200.LRCEOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [S], 32).cuda().eval()
    tf = torch.randn(B, 32, 768, device="cuda")
vf = torch.randn(B, S, 3, 49, 1024, device="cuda").bfloat16()
NAMES = ["params wait", "sa: MMA wait", "sa: exchange", "LN1 + signal", "q: MMA wait", "q: exchange", "attn: K wait", "attn: scores",
         "attn: V wait", "attn: PV + combine", "ctx send", "o: MMA wait", "o: exchange", "LN2 + signal", "fc1: MMA wait",
         "fc1: gelu epilogue", "fc2: MMA wait", "fc2: partial staging", "fc2: reduce-scatter", "fc2: reduce + bias",
         "fc2: all-gather", "LN3 + signal", "head", ""]
with torch.no_grad():
    for _ in range(3):
        y_ref = m(vf, tf)
    torch.cuda.synchronize()
    pk = m.packed()
    real = ops.encoder_walk
    # uninstrumented timing of the code variants (same box, interleaved)
    for rep in range(3):
        for variant in (1, 0):
            tms = []

            def timed(packed, n_layers, kv_video, kv_text, tok0, f_g, f_b, eps, n_out, act, out, rows, S_, Tv, Lt, n_cand, tokens_tap=None):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(_lib.lib().lrce_encoder_walk_profile(
                    packed.data_ptr(), n_layers, kv_video.data_ptr(), kv_text.data_ptr(), kv_video.stride(0), tok0.data_ptr(),
                    f_g.data_ptr(), f_b.data_ptr(), float(eps), n_out, act, out.data_ptr(), 0, rows, S_, Tv, Lt, n_cand,
                    torch.cuda.current_stream().cuda_stream, 0, 0, variant), "walk variant")
                e1.record()
                tms.append((e0, e1))
                return out

            ops.encoder_walk = timed
            try:
                for _ in range(5):
                    y = m(vf, tf)
            finally:
                ops.encoder_walk = real
            torch.cuda.synchronize()
            print(f"variant {variant}: walk kernel {min(a.elapsed_time(b) for a, b in tms)*1e3:.1f} us (best of 5), "
                  f"max |y - y_ref| = {(y - y_ref).abs().max().item():.2e}")

    for max_clusters, variant in ((0, 1),):
        prof = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

        def profiled(packed, n_layers, kv_video, kv_text, tok0, f_g, f_b, eps, n_out, act, out, rows, S_, Tv, Lt, n_cand, tokens_tap=None):
            ev[0].record()
            _lib.check(_lib.lib().lrce_encoder_walk_profile(
                packed.data_ptr(), n_layers, kv_video.data_ptr(), kv_text.data_ptr(), kv_video.stride(0), tok0.data_ptr(),
                f_g.data_ptr(), f_b.data_ptr(), float(eps), n_out, act, out.data_ptr(), 0, rows, S_, Tv, Lt, n_cand,
                torch.cuda.current_stream().cuda_stream, prof.data_ptr(), max_clusters, variant), "walk profile")
            ev[1].record()
            return out

        ops.encoder_walk = profiled
        try:
            y = m(vf, tf)
        finally:
            ops.encoder_walk = real
        torch.cuda.synchronize()
        t = prof.view(148, 32).cpu().double()
        used = (t.sum(1) > 0).sum().item()
        steps = S * 12
        tot = t[:used, :24].sum(1)
        slow = int(tot.argmax())
        print(f"B={B} {kind} max_clusters={max_clusters or 'all'} variant={variant}: {used} CTAs, kernel {ev[0].elapsed_time(ev[1])*1e3:.1f} us, "
              f"max |y - y_ref| = {(y - y_ref).abs().max().item():.2e}")
        print(f"  {'cycles per layer-step':28s} {'CTA 0':>10s} {'slowest CTA ' + str(slow):>16s}")
        for j, nme in enumerate(NAMES[:23]):
            print(f"  {nme:28s} {t[0, j].item()/steps:10.0f} {t[slow, j].item()/steps:16.0f}")
        print(f"  {'sum':28s} {t[0, :24].sum().item()/steps:10.0f} {t[slow, :24].sum().item()/steps:16.0f}")
        for j, nme in enumerate(["MMA warp: stream stall sa/q/o", "MMA warp: stream stall fc1", "MMA warp: stream stall fc2", "MMA warp: waiting for B"]):
            print(f"  {nme:28s} {t[0, 24 + j].item()/steps:10.0f} {t[slow, 24 + j].item()/steps:16.0f}")
