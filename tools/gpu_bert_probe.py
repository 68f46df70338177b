"""GPU probe: what BERT (HF module on a side stream) costs the E2E forward: step time with the real text extractor vs with
its output cached, and BERT alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import lrce_b200

dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["msvd-qa-oe"]
torch.manual_seed(0)
model = lrce_b200.E2EOpenEnded(pretrained=False, **bench.model_kwargs(cfg)).to(dev).eval()
inputs = [t.to(dev) for t in bench.synth_inputs(32, cfg, seed=1)]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    full = timeit(lambda: model(*inputs))
    bert = timeit(lambda: model.extract_text_features(*inputs[1:]))
    cached = model.extract_text_features(*inputs[1:])
    orig = model.text_extractor.forward
    model.text_extractor.forward = lambda *a, **k: cached
    nobert = timeit(lambda: model(*inputs))
    model.text_extractor.forward = orig
    swin = timeit(lambda: model.extract_video_features(inputs[0]))
print(f"E2E forward {full:.2f} ms | with cached text features {nobert:.2f} ms | BERT alone {bert:.2f} ms | Swin alone {swin:.2f} ms")

# ---- the same question for the Swin path: eager launches (ctypes -> liblrce_b200) vs one CUDA graph replay
with torch.no_grad():
    static = inputs[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            model.extract_video_features(static)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_static = model.extract_video_features(static)
    ref = model.extract_video_features(static)
    g.replay()
    torch.cuda.synchronize()
    print("graph == eager:", torch.equal(ref, out_static))
    t_graph = timeit(lambda: g.replay())
    t_eager = timeit(lambda: model.extract_video_features(static))
print(f"Swin eager {t_eager:.2f} ms | Swin as one CUDA graph {t_graph:.2f} ms")
