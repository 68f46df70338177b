"""GPU probe: what BERT costs the forward. Times the msvd-qa-oe E2E forward (batch 32) as shipped (BERT graph on a side stream, in
Swin's shadow), with the text extractor stubbed out (cached features), and with BERT enqueued on the main stream in front of Swin."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrce_b200

B, S, L = 32, 3, 32
m = lrce_b200.E2EOpenEnded(768, 1000, 0.1, [7, 7], 1024, 5, [S], L, pretrained=False).cuda().eval()
clips = torch.rand(B, S, 5, 3, 224, 224, device="cuda")
ids = torch.randint(1000, 20000, (B, L), device="cuda")
mask = torch.ones(B, L, dtype=torch.long, device="cuda")
types = torch.zeros(B, L, dtype=torch.long, device="cuda")


def timeit(fn, n=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    def full():
        return m(clips, ids, mask, types)

    for rep in range(2):
        t_full = timeit(full)
        real = m.extract_text_features
        cached = real(ids, mask, types).clone()
        m.extract_text_features = lambda *a: cached
        t_stub = timeit(full)
        m.extract_text_features = real
        side = m._side_stream
        m._side_stream = lambda dev: torch.cuda.current_stream()
        t_serial = timeit(full)
        m._side_stream = side
        print(f"shipped (side stream) {t_full:.3f} ms   text stubbed {t_stub:.3f} ms   BERT on the main stream {t_serial:.3f} ms", flush=True)
