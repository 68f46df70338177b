"""GPU probe: pinned host -> device copy bandwidth for one config-2 batch of clips (289 MB fp32), alone and next to a
running forward, plus the host-side cost of narrowing the batch to fp16 first."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

dev = torch.device("cuda", 0)
x = torch.rand(32, 3, 5, 3, 224, 224).pin_memory()
nbytes = x.numel() * 4
d = torch.empty_like(x, device=dev)
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D pinned fp32 {nbytes/1e6:.0f} MB: {ms:.2f} ms = {nbytes/ms/1e6:.1f} GB/s", flush=True)
# pageable for comparison
xp = torch.rand(32, 3, 5, 3, 224, 224)
t0 = time.perf_counter(); d.copy_(xp); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"H2D pageable: {(t1-t0)*1e3:.2f} ms", flush=True)
# host narrowing cost
print("torch threads", torch.get_num_threads(), "cpus", os.cpu_count(), flush=True)
h16 = torch.empty(x.shape, dtype=torch.float16).pin_memory()
for _ in range(2):
    h16.copy_(x)
t0 = time.perf_counter()
for _ in range(5):
    h16.copy_(x)
t1 = time.perf_counter()
print(f"host fp32->fp16 narrowing: {(t1-t0)/5*1e3:.2f} ms", flush=True)
d16 = torch.empty_like(h16, device=dev)
d16.copy_(h16, non_blocking=True); torch.cuda.synchronize()
e0.record()
for _ in range(5):
    d16.copy_(h16, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D pinned fp16 {nbytes/2e6:.0f} MB: {ms:.2f} ms = {nbytes/2/ms/1e6:.1f} GB/s", flush=True)
os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv")
os.system("nvidia-smi topo -m | head -5; lscpu | grep -E 'Model name|^CPU\\(s\\)|NUMA node\\(s\\)'")
