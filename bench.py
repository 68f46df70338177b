#!/usr/bin/env python
"""bench.py — LRCE forward throughput on B200 (BASELINE.json metric: clips/sec of the E2E forward).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path (default workload: configs[1])
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                      # clip-sharded, weak scaling, NCCL logit gather
    python bench.py --impl reference --steps 2 --warmup 1           # CPU baseline arm (oracle port of the reference)

A "step" is one E2E forward over one batch of `--batch` synthetic clips per GPU (default 32 = configs[1]: MSVD-QA
open-ended eval forward, batch 32 bf16, temporal-scale 3). One JSON line is printed by rank 0.
  value  : clips/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    : clips/s through the public module call with HOST (pinned) inputs: H2D of the clips/ids and D2H of the logits
           are inside the timed region
  roofline : the dominant kernel family (the tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of its launches inside
           the timed region, against the measured sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (a torch restatement of the reference pinned to its golden vectors) timed on this
           box's host cores on a bounded sample (rank 0, N=1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line

CONFIGS = {  # reference configs/*.json
    "msvd-qa-oe": dict(kind="oe", num_classes=1000, text_seq_len=32),
    "msrvtt-qa-oe": dict(kind="oe", num_classes=1500, text_seq_len=37),
    "tgif-frameqa": dict(kind="oe", num_classes=1000, text_seq_len=30),
    "tgif-action": dict(kind="mc", num_classes=1, text_seq_len=40),
    "tgif-transition": dict(kind="mc", num_classes=1, text_seq_len=40),
    "tgif-count": dict(kind="count", num_classes=1, text_seq_len=30),
}
GFLOP_PER_CLIP = 303.96  # Swin-B 3 x 96.354 + canonical encoder 14.90 (msvd-qa-oe)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


# gemm_tc_kernel + mlp_l2_kernel launches of Swin + encoder in one batch-32 forward (ncu launch list; BERT's 48 gemm_tc_kernel
# launches excluded: 304 MB). mlp_l2_kernel is the same pipeline walking fc1 -> fc2 per row tile (20 launches replace 40).
GEMM_DRAM_BYTES_PER_STEP = (7037.9e6 - 303.7e6) + 5748.7e6 + 3247.5e6 + 3131.5e6
GEMM_LAUNCHES_NCU = 79.0
GEMM_TRAFFIC_SOURCE = "profiles/r02_launches_v4_summary.md"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["source"] = "fallback"
    return d


class ClockSampler:
    """samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line). In-process NVML
    (nvidia_ml_py) every 5 ms on a thread, so even a 100 ms timed region gets samples; `nvidia-smi -lms` as a fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.power, self.reasons = [], [], [], set()

    def start(self):
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            try:  # honour CUDA_VISIBLE_DEVICES remapping
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.reasons |= {name for bit, name in self.BITS.items() if r & bit}
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(sm),
                    "power_w": round(sorted(self.power)[len(self.power) // 2], 1) if self.power else None, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def model_kwargs(cfg):
    return dict(feature_dim=768, num_classes=cfg["num_classes"], video_feature_res=[7, 7], video_feature_dim=1024,
                frame_sample_size=5, temporal_scale=[3], text_seq_len=cfg["text_seq_len"])


def synth_inputs(batch, cfg, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    L = cfg["text_seq_len"]
    lead = (batch, 5) if cfg["kind"] == "mc" else (batch,)
    clips = torch.rand((batch, 3, 5, 3, 224, 224), generator=g)
    ids = torch.randint(1000, 30522, lead + (L,), generator=g)
    ids[..., 0] = 101
    ids[..., 19] = 102
    ids[..., 20:] = 0
    mask = torch.zeros(lead + (L,), dtype=torch.int64)
    mask[..., :20] = 1
    types = torch.zeros(lead + (L,), dtype=torch.int64)
    return clips, ids, mask, types


# ----------------------------------------------------------------------------------------------------------------------
def _cpu_forward_fn(cfg_name, batch):
    """(callable running ONE batch-`batch` E2E forward on the host cores, kind, description). The unmodified reference
    (its own E2E* modules imported from baseline/_ref or /root/reference through oracle/ref_harness.py) when a copy is
    present — kind "reference"; otherwise the oracle port (oracle/lrce_oracle.py) — kind "port"."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lrce_oracle as O
    import ref_harness
    import weights as W

    cfg = CONFIGS[cfg_name]
    sd = W.make_e2e_state_dict(cfg["num_classes"], cfg["text_seq_len"], 3, seed=0)
    inputs = W.make_inputs(batch, 3, cfg["text_seq_len"], seed=1, n_candidates=5 if cfg["kind"] == "mc" else 0)
    if ref_harness.find_reference() is not None:
        try:
            model = ref_harness.build_reference_e2e(cfg["kind"], model_kwargs(cfg), sd)
            return (lambda: model(*inputs)), "reference", "the unmodified reference E2E module (fp32, torch CPU)"
        except Exception as e:  # a broken copy must not take the bench down: fall back to the port and say so
            sys.stderr.write(f"[bench] reference import failed ({type(e).__name__}: {e}); timing the oracle port\n")
    return (lambda: O.e2e_forward(sd, *inputs, cfg["kind"])), "port", "oracle/lrce_oracle.py (fp32, torch CPU)"


def cpu_baseline(cfg_name, steps=3, warmup=1, batch=2, max_seconds=None):
    """times the reference's CPU path on all host cores: `warmup` untimed + exactly `steps` timed batch-`batch` forwards
    (mean). `max_seconds` bounds the sample for the in-line cpu_baseline of the B200 arm (fewer steps, never zero)."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, what = _cpu_forward_fn(cfg_name, batch)
    times = []
    t_start = time.time()
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        for _ in range(steps):
            t0 = time.time()
            fwd()
            times.append(time.time() - t0)
            if max_seconds is not None and time.time() - t_start > max_seconds:
                break
    mean = sum(times) / len(times)
    return {"value": batch / mean, "unit": "clips/s", "cores": cores, "kind": kind,
            "sample": f"{cfg_name} E2E forward, batch {batch} fp32, mean of {len(times)} forwards after {warmup} warm-up: "
                      f"{what}, {cores} threads"}, mean, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, mean, n = cpu_baseline(args.config, steps=args.steps, warmup=args.warmup, batch=2)
    line = {"impl": "reference", "metric": "clips/sec LRCE fwd", "value": base["value"], "unit": "clips/s",
            "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": mean * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} E2E eval forward (CPU; each step = a batch-2 sample of the batch-{args.batch} "
                                   "workload)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def torch_eager_on_b200(cfg_name, batch, dev, steps=5, warmup=2):
    """the reference's own modules on the same B200 — torch eager library kernels (cuBLAS / cuDNN / ATen)
    under the agent's fp16 autocast (agent_oe.py:28), device-resident inputs. This is the "PyTorch on the same box" bar
    the hand-written kernels are measured against; it needs a copy of the reference (baseline/_ref)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness
    import weights as W

    if ref_harness.find_reference() is None:
        return {"unavailable": "no copy of the reference under baseline/_ref (made by __graft_entry__.build() in the build container)"}
    cfg = CONFIGS[cfg_name]
    try:
        sd = W.make_e2e_state_dict(cfg["num_classes"], cfg["text_seq_len"], 3, seed=0)
        model = ref_harness.build_reference_e2e(cfg["kind"], model_kwargs(cfg), sd).to(dev).eval()
        inputs = [t.to(dev) for t in synth_inputs(batch, cfg, seed=1)]
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for _ in range(warmup):
                model(*inputs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(*inputs)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del model
        torch.cuda.empty_cache()
        return {"value": batch / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
                "what": f"unmodified reference {cfg_name} E2E module, torch {torch.__version__} eager, fp16 autocast, batch {batch}, "
                        "inputs resident in HBM"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores local to GPU `index` (NVML's ideal affinity) BEFORE the pinned host buffers are
    allocated, so that they are first-touched on the NUMA node the GPU's PCIe root hangs off: on two-socket hosts a remote
    pinned buffer halves the H2D rate of the 289 MB fp32 batch (the e2e figure's box-to-box spread in round 1)."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cores_before": before, "cores_after": len(os.sched_getaffinity(0)), "bound": True}
    except Exception as e:
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import lrce_b200
    from lrce_b200 import dist as ldist
    from lrce_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): the LRCE hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = CONFIGS[args.config]
    cls = {"oe": lrce_b200.E2EOpenEnded, "mc": lrce_b200.E2EMultipleChoice, "count": lrce_b200.E2ECount}[cfg["kind"]]
    numa = bind_to_gpu_numa_node(local)
    torch.manual_seed(0)
    model = cls(pretrained=False, **model_kwargs(cfg)).to(dev).eval()
    B = args.batch
    host = [t.pin_memory() for t in synth_inputs(B, cfg, seed=1 + rank)]
    devin = [t.to(dev) for t in host]
    n_out = (B, 5) if cfg["kind"] == "mc" else ((B,) if cfg["kind"] == "count" else (B, cfg["num_classes"]))
    host_out = torch.empty(n_out, dtype=torch.float32).pin_memory()

    def step(inputs):
        with torch.no_grad():
            y = model(*inputs)
            if world > 1:  # clip-sharded eval: the only collective is the final logit gather
                ldist.gather_logits(y)
            return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(devin)
    barrier()

    # ---- timed region 1: device-resident inputs (CUDA events on the launching stream, clocks sampled meanwhile)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(devin)
    e1.record()
    barrier()
    gpu_launches = ops.launches - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- same steps again with every launch of liblrce_b200 bracketed by CUDA events: the per-kernel table / roofline
    ops.trace = []
    for _ in range(args.steps):
        step(devin)
    barrier()
    trace, ops.trace = ops.trace, None

    # ---- timed region 2: end to end through the public API with HOST (pinned) buffers: every step's inputs are copied
    # host -> device inside the timed region (lrce_b200.feed.PrefetchFeed: the copy of batch i+1 runs on a side stream under
    # the forward of batch i) and every step's logits are read back to the host
    from lrce_b200.feed import PrefetchFeed

    for cur in PrefetchFeed([host] * 2, dev):
        step(cur)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for cur in PrefetchFeed([host] * args.steps, dev):
        y = step(cur)
        host_out.copy_(y, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks_e2e = sampler.stop()
    # copy-only timing of the same batches (no compute): attributes an e2e shortfall to PCIe / host memory vs everything else
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_copy = min(args.steps, 10)
    c0.record()
    for _ in range(n_copy):
        for d, h in zip(devin, host):
            d.copy_(h, non_blocking=True)
    c1.record()
    barrier()
    ms_h2d = c0.elapsed_time(c1) / n_copy

    # ---- timed region 3 (extra, not the contract's e2e): the same end-to-end loop fed with the uint8 frames themselves
    # (the module's extension over the reference's ToTensor()-ed fp32 clips: x / 255 happens in the first kernel), i.e. a
    # quarter of the host -> device bytes. With several GPUs on one host the fp32 feed is bound by the host side of PCIe.
    host8 = [(host[0] * 255).round().to(torch.uint8).pin_memory()] + host[1:]
    for cur in PrefetchFeed([host8] * 2, dev):
        step(cur)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for cur in PrefetchFeed([host8] * args.steps, dev):
        y = step(cur)
        host_out.copy_(y, non_blocking=True)
    e5.record()
    barrier()
    ms_e2e8 = e4.elapsed_time(e5)
    clocks_e2e8 = sampler.stop()
    # device-resident steps once more, now that the board has been under load for seconds: separates the clock droop of the
    # later regions from anything the host feed costs
    e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e6.record()
    for _ in range(args.steps):
        step(devin)
    e7.record()
    barrier()
    ms_late = e6.elapsed_time(e7)

    t = torch.tensor([ms, ms_e2e, ms_e2e8], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e8 = t.tolist()

    if rank == 0:
        peaks = load_peaks()
        fam = {}
        for name, tag, flops, nbytes, a, b in trace:
            d = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b)
            d["flops"] += flops
            d["bytes"] += nbytes
            d["launches"] += 1
        kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                       "tflops": v["flops"] / v["ms"] / 1e9 if v["ms"] else 0.0,
                       "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] else 0.0} for k, v in fam.items()}
        # the GEMM family = gemm_tc_kernel + mlp_l2_kernel (the same tcgen05 pipeline walking fc1 -> fc2 per row tile)
        g = {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0}
        for nm in ("lrce_gemm_bf16", "lrce_mlp_l2_bf16"):
            for key in g:
                g[key] += fam.get(nm, {}).get(key, 0)
        if not g["ms"]:
            g = {"ms": 1.0, "flops": 0.0, "bytes": 0.0, "launches": 1}
        achieved = g["flops"] / g["ms"] / 1e9
        # window attention, both figures of SURVEY.md 8(d): (i) the core QK^T + PV alone, (ii) the W-MSA block = qkv GEMM +
        # core + proj GEMM (the qkv / proj launches are recognised by their shapes: N == 3K with the plain epilogue, N == K
        # with the residual epilogue)
        att = fam.get("lrce_window_attention_bf16", {"ms": 0.0, "flops": 0.0})
        blk_ms, blk_flops = att["ms"], att["flops"]
        for name, tag, flops, nbytes, a, b in trace:
            if name == "lrce_gemm_bf16":
                Mg, rest = tag[1:].split("N")
                Ng, rest = rest.split("K")
                Kg, eg = rest.split("e")
                if int(Mg) % 147 == 0 and ((int(Ng) == 3 * int(Kg) and eg == "0") or (Ng == Kg and eg == "2")):
                    blk_ms += a.elapsed_time(b)
                    blk_flops += flops
        burst = peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])
        peak = peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])
        clips_per_s = world * B * args.steps / (ms / 1e3)
        h2d = sum(t.numel() * t.element_size() for t in host)
        line = {
            "metric": "clips/sec LRCE fwd", "value": clips_per_s, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.config} E2E eval forward (Video Swin-B + BERT-base + recurrent encoder + head), "
                                   f"batch {B} clips/GPU, temporal-scale 3, random-init weights",
                       "global_batch": world * B, "parallelism": f"clip-sharded dp{world}" + (" + NCCL logit all_gather" if world > 1 else ""),
                       "l2_policy": "inputs (289 MB of fp32 clips per step) and every stage tensor exceed the 126 MB L2"},
            "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": ms_e2e / args.steps, "clocks": clocks_e2e,
                    "h2d_only_ms_per_step": ms_h2d, "h2d_only_gbs": h2d / ms_h2d / 1e6,
                    "device_resident_ms_per_step_after_load": ms_late / args.steps,
                    "host_numa_binding": numa},
            "e2e_uint8_frames": {"value": world * B * args.steps / (ms_e2e8 / 1e3), "unit": "clips/s",
                                 "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host8),
                                 "ms_per_step": ms_e2e8 / args.steps, "clocks": clocks_e2e8,
                                 "note": "extension: uint8 frames as the module input (x/255 in the first kernel); not the "
                                         "reference-API e2e above"},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel + mlp_l2_kernel (tcgen05, all Linear layers but the stage-1 MLP)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": f"{peaks['source']} bf16_tflops_sustained",
                         # DRAM bytes (read + write) per launch, mean over the GEMM launches of one forward: a STATIC figure
                         # from the committed ncu launch list of this kernel set, not measured in this run
                         "traffic": GEMM_DRAM_BYTES_PER_STEP / GEMM_LAUNCHES_NCU if args.config == "msvd-qa-oe" and B == 32 else None,
                         "traffic_note": "static: bytes per launch, ncu dram__bytes_read.sum + dram__bytes_write.sum over the GEMM "
                                         "launches of " + GEMM_TRAFFIC_SOURCE,
                         "algorithmic_bytes_per_launch": g.get("bytes", 0.0) / max(g["launches"], 1),
                         "launches_per_step": g["launches"] / args.steps, "share_of_step": g["ms"] / ms,
                         "whole_forward_tflops": clips_per_s / world * GFLOP_PER_CLIP / 1e3,
                         "whole_forward_frac": clips_per_s / world * GFLOP_PER_CLIP / 1e3 / peak},
            "window_attention": {
                "core_tflops": att["flops"] / att["ms"] / 1e9 if att["ms"] else 0.0,
                "core_frac_of_sustained_peak": att["flops"] / att["ms"] / 1e9 / peak if att["ms"] else 0.0,
                "core_ms_per_step": att["ms"] / args.steps,
                "wmsa_block_tflops": blk_flops / blk_ms / 1e9 if blk_ms else 0.0,
                "wmsa_block_frac_of_sustained_peak": blk_flops / blk_ms / 1e9 / peak if blk_ms else 0.0,
                "wmsa_block_frac_of_burst_peak": blk_flops / blk_ms / 1e9 / burst if blk_ms else 0.0,
                "wmsa_block_ms_per_step": blk_ms / args.steps,
                "note": "core = QK^T + PV (4*147^2*32 FLOP per window-head), bound by the softmax warps' per-unit latency chain; "
                        "block = qkv GEMM + core + proj GEMM"},
            "kernels": kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_baseline(args.config, steps=10, warmup=1, max_seconds=25.0)
        if world == 1 and not args.no_eager:
            del model
            torch.cuda.empty_cache()
            line["torch_eager_b200"] = torch_eager_on_b200(args.config, B, dev)
            if "value" in line["torch_eager_b200"]:
                line["torch_eager_b200"]["speedup_of_this_repo"] = clips_per_s / line["torch_eager_b200"]["value"]
        emit(line)
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------------------------------------------------
def run_train_arm(args):
    """BASELINE.json configs[4]: TGIF-FrameQA training step, clip-sharded data parallel: extractors forward-only on
    liblrce_b200, cross-modal encoder forward + backward on the hand-written training kernels (lrce_b200/train.py), the
    115 M encoder gradients all-reduced over NCCL layer by layer INSIDE the backward pass (`grad_sync = "overlap"`), AdamW
    step on the encoder (torch's fused optimizer, as the reference agent uses torch's AdamW). Not the headline metric;
    printed as its own JSON line."""
    import torch
    import torch.distributed as dist

    import lrce_b200
    from lrce_b200 import dist as ldist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): the LRCE hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = CONFIGS[args.config]
    cls = {"oe": lrce_b200.E2EOpenEnded, "mc": lrce_b200.E2EMultipleChoice, "count": lrce_b200.E2ECount}[cfg["kind"]]
    torch.manual_seed(0)
    model = cls(pretrained=False, **model_kwargs(cfg)).to(dev).train()
    model.fusion_model.grad_sync = "overlap"  # all-reduce each layer's gradients as soon as the backward pass completes them
    enc = [p for p in model.fusion_model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(enc, lr=1e-5, fused=True)
    B = args.batch
    host = [t.pin_memory() for t in synth_inputs(B, cfg, seed=1 + rank)]
    g = torch.Generator().manual_seed(7 + rank)
    if cfg["kind"] == "oe":
        target = torch.randint(0, cfg["num_classes"], (B,), generator=g).to(dev)
        loss_fn = torch.nn.functional.cross_entropy
    elif cfg["kind"] == "mc":
        target = torch.randint(0, 5, (B,), generator=g).to(dev)
        loss_fn = torch.nn.functional.cross_entropy
    else:
        target = torch.randint(1, 10, (B,), generator=g).float().to(dev)
        loss_fn = torch.nn.functional.mse_loss
    grad_bytes = [sum(p.numel() * 4 for p in enc)]

    devin = [t.to(dev) for t in host]
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def step(inputs):
        loss = loss_fn(model(*inputs), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()  # includes the overlapped NCCL all-reduce of the encoder gradients
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from lrce_b200.feed import PrefetchFeed

    for _ in range(max(args.warmup, 3)):
        step(devin)
    barrier()
    # ---- timed region 1: the batch already resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(devin)
    e1.record()
    barrier()
    clocks = sampler.stop()
    # ---- timed region 2: end to end, pinned host batches through the prefetching feed, loss read back every step
    for cur in PrefetchFeed([host] * 2, dev):
        step(cur)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for cur in PrefetchFeed([host] * args.steps, dev):
        loss = step(cur)
        host_loss.copy_(loss.detach(), non_blocking=True)
    e3.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1), e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        v = world * B * args.steps / (ms / 1e3)
        line = {"metric": "clips/sec LRCE train step (encoder fwd+bwd + grad all-reduce)", "value": v, "unit": "clips/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.config} training step: Swin-B + BERT forward (liblrce_b200, frozen), cross-modal "
                                       f"encoder fwd+bwd (hand-written kernels, CUDA graphs), AdamW on the encoder, batch {B} "
                                       "clips/GPU, temporal-scale 3, random-init weights", "global_batch": world * B,
                           "parallelism": f"clip-sharded dp{world}" + (" + per-layer NCCL all-reduce overlapped with backward" if world > 1 else ""),
                           "grad_bytes_per_step": grad_bytes[0]},
                "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "clips/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host), "d2h_bytes_per_step": 4},
                "clocks": clocks, "final_loss": float(loss.item())}
        emit(line)
    if world > 1:
        dist.destroy_process_group()

_RESULT_FD = None


def emit(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    # Native libraries print to stdout too (NCCL's "NCCL version ..." banner at communicator creation): keep the original
    # stdout for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="msvd-qa-oe", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-B200 figure (reference modules, fp16 autocast)")
    ap.add_argument("--train", action="store_true", help="time the configs[4] training step instead of the eval forward")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.train:
        if a.config == "msvd-qa-oe":
            a.config = "tgif-frameqa"
        run_train_arm(a)
    else:
        run_b200_arm(a)
