"""GPU probe: the Swin MLP of stages 2 / 3 as two GEMM launches (hidden rows through DRAM) against lrce_mlp_l2_bf16 (hidden rows
in an L2-resident scratch), back to back for ~1 s each with the SM clock and board power sampled through NVML."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

import lrce_b200  # noqa: F401
from lrce_b200 import _lib
if os.environ.get("LRCE_LIB"):
    _lib.LIB_PATH = os.environ["LRCE_LIB"]
from lrce_b200 import ops

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def run(name, fn, flop):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    iters = max(10, int(seconds * 1e3 / e0.elapsed_time(e1)))
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.02)

    th = threading.Thread(target=sampler)
    th.start()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / iters
    tail = samples[len(samples) // 2:] or samples
    mhz = sorted(s[0] for s in tail)[len(tail) // 2]
    watts = sorted(s[1] for s in tail)[len(tail) // 2]
    print(f"{name:22s} {ms * 1e3:8.1f} us  {flop / ms / 1e9:7.1f} TFLOP/s  SM {mhz} MHz {watts:.0f} W", flush=True)
    return ms


for M, C in ((56448, 512), (225792, 256), (903168, 128)):
    x = (torch.randn(M, C, device=dev) * 0.8).bfloat16()
    w1 = (torch.randn(4 * C, C, device=dev) * 0.04).bfloat16()
    w2 = (torch.randn(C, 4 * C, device=dev) * 0.04).bfloat16()
    b1, c1, b2 = torch.randn(4 * C, device=dev) * 0.1, torch.randn(4 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
    nc = C // ops.stats_chunk(C)
    st_in = torch.rand(M * nc * 2, device=dev) + 0.5
    st_out = torch.empty(M * nc * 2, device=dev)
    out = torch.empty_like(x)
    flop = 2.0 * M * C * 4 * C * 2

    def two():
        hid = ops.gemm(x, w1, b1, epilogue=ops.EPI_BIAS_GELU, ln_in=(st_in, c1, 1e-5))
        ops.gemm(hid, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=x, out=out, stats_out=st_out)

    def fused():
        ops.mlp_l2(x, w1, b1, c1, st_in, 1e-5, w2, b2, out=out, stats_out=st_out)

    print(f"M={M} C={C}", flush=True)
    def sm_fused():
        ops.mlp_fused(x, w1, b1, c1, st_in, 1e-5, w2, b2, out=out, stats_out=st_out)

    for rep in range(2):
        run("two GEMM launches", two, flop)
        if C == 128:
            run("mlp_fused (in-SM)", sm_fused, flop)
        run("mlp_l2 (one launch)", fused, flop)
