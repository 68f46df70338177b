"""GPU probe: correctness of the tcgen05 GEMM on the shapes of Appendix B + a CUDA-event throughput sweep."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrce_b200
from lrce_b200 import _lib as _l
if os.environ.get('LRCE_LIB') or os.environ.get('ATTN_LIB'):
    _l.LIB_PATH = os.environ.get('LRCE_LIB') or os.environ['ATTN_LIB']  # A/B variants of the library (tools only)
from lrce_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def ref(a, w, bias, epi, res, ln):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if epi == ops.EPI_BIAS_GELU:
        y = torch.nn.functional.gelu(y)
    if epi == ops.EPI_BIAS_RESIDUAL:
        y = y + res.float()
    if epi == ops.EPI_BIAS_LN:
        y = torch.nn.functional.layer_norm(y, (y.shape[1],), ln[0], ln[1], ln[2])
    return y


def check(M, N, K, epi, fp32=False):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16() if epi == ops.EPI_BIAS_RESIDUAL else None
    ln = (torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev), 1e-5) if epi == ops.EPI_BIAS_LN else None
    y = ops.gemm(a, w, bias, epilogue=epi, residual=res, out_fp32=fp32, ln=ln)
    torch.cuda.synchronize()
    r = ref(a, w, bias, epi, res, ln)
    err = (y.float() - r).abs().max().item()
    rel = ((y.float() - r).norm() / r.norm()).item()
    ok = rel < (1e-5 if fp32 else 6e-3)
    print(f"M={M:7d} N={N:5d} K={K:5d} epi={epi} fp32={int(fp32)} max_abs={err:.4e} rel_l2={rel:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def bench(M, N, K, epi, iters=20, lnin=False, stats=False):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16() if epi == ops.EPI_BIAS_RESIDUAL else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw = {}
    if lnin:
        kw["ln_in"] = (torch.rand(K // 32 * M * 2, device=dev) + 0.5, torch.randn(N, device=dev), 1e-5)
    if stats:
        kw["stats_out"] = torch.empty(N // 32 * M * 2, device=dev)
    for _ in range(3):
        ops.gemm(a, w, bias, epilogue=epi, residual=res, out=out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, bias, epilogue=epi, residual=res, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparison (library baseline)
    for _ in range(3):
        torch.matmul(a, w.t())
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"BENCH M={M} N={N} K={K} epi={epi} lnin={int(lnin)} stats={int(stats)}: {ms*1e3:.1f} us  {tf:.1f} TFLOP/s  (cuBLAS matmul {ms2*1e3:.1f} us {2.0*M*N*K/ms2/1e9:.1f} TF)", flush=True)


if __name__ == "__main__" and "--sustain" not in sys.argv:
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= check(128, 128, 64, ops.EPI_BIAS, fp32=True)
    ok &= check(128, 128, 128, ops.EPI_BIAS, fp32=True)
    ok &= check(256, 256, 256, ops.EPI_BIAS, fp32=True)
    ok &= check(300, 384, 128, ops.EPI_BIAS)
    ok &= check(882, 3072, 1024, ops.EPI_BIAS)
    ok &= check(3528, 2048, 512, ops.EPI_BIAS_GELU)
    ok &= check(3528, 512, 2048, ops.EPI_BIAS_RESIDUAL)
    ok &= check(14112, 768, 256, ops.EPI_BIAS)
    ok &= check(56448, 128, 96, ops.EPI_BIAS_LN)
    ok &= check(1000, 128, 96, ops.EPI_BIAS_LN)
    ok &= check(14112, 256, 512, ops.EPI_BIAS)
    ok &= check(32, 768, 768, ops.EPI_BIAS, fp32=True)
    ok &= check(160, 1024, 768, ops.EPI_BIAS, fp32=True)
    print("ALL_OK" if ok else "SOME_FAIL", flush=True)
    if ok or "--force-bench" in sys.argv:
        bench(8192, 8192, 8192, ops.EPI_BIAS)
        for M, C in ((903168, 128), (225792, 256), (56448, 512), (14112, 1024)):  # one Swin block per stage, as executed
            bench(M, 3 * C, C, ops.EPI_BIAS, lnin=True)
            bench(M, C, C, ops.EPI_BIAS_RESIDUAL, stats=True)
            bench(M, 4 * C, C, ops.EPI_BIAS_GELU, lnin=True)
            bench(M, C, 4 * C, ops.EPI_BIAS_RESIDUAL, stats=True)
        bench(56448, 1536, 512, ops.EPI_BIAS)
        bench(56448, 2048, 512, ops.EPI_BIAS_GELU)
        bench(56448, 512, 2048, ops.EPI_BIAS_RESIDUAL)
        bench(56448, 512, 512, ops.EPI_BIAS_RESIDUAL)
        bench(903168, 384, 128, ops.EPI_BIAS)
        bench(903168, 512, 128, ops.EPI_BIAS_GELU)
        bench(903168, 128, 512, ops.EPI_BIAS_RESIDUAL)
        bench(14112, 4096, 1024, ops.EPI_BIAS_GELU)
        bench(15456, 18432, 768, ops.EPI_BIAS)


def sustain(M, N, K, epi, seconds=1.5):
    """Long back-to-back runs of one shape (this kernel, then cuBLAS) with the SM clock and board power sampled through NVML:
    separates 'slower per clock' from 'lower clock under the power cap'."""
    import threading
    import time

    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16() if epi == ops.EPI_BIAS_RESIDUAL else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for name, fn in (("own", lambda: ops.gemm(a, w, bias, epilogue=epi, residual=res, out=out)), ("cuBLAS", lambda: torch.matmul(a, w.t()))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        iters = max(20, int(seconds * 1e3 / e0.elapsed_time(e1)))
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
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / iters
        tail = samples[len(samples) // 2:]
        mhz = sorted(s[0] for s in tail)[len(tail) // 2]
        watts = sorted(s[1] for s in tail)[len(tail) // 2]
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"SUSTAIN {name:7s} M={M} N={N} K={K} epi={epi}: {ms*1e3:.1f} us {tf:.1f} TFLOP/s  SM {mhz} MHz {watts:.0f} W  "
              f"-> {tf * 1e12 / (148 * mhz * 1e6 * 8192) * 100:.1f} % of the tensor pipe at that clock", flush=True)


if __name__ == "__main__" and "--sustain" in sys.argv:
    sustain(8192, 8192, 8192, ops.EPI_BIAS)
    sustain(56448, 512, 2048, ops.EPI_BIAS_RESIDUAL)
    sustain(56448, 2048, 512, ops.EPI_BIAS_GELU)
    sustain(56448, 512, 512, ops.EPI_BIAS_RESIDUAL)
