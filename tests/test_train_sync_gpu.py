"""GPU (needs >= 2 devices; skipped otherwise): the per-layer all-reduce issued INSIDE the encoder's backward pass
(`grad_sync = "overlap"`, lrce_b200/train.py) must leave every rank with the mean of the per-rank gradients — the same
result as computing the gradients unsynchronised and all-reducing them afterwards (what DDP does, agent_base.py:76)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist

    import lrce_b200
    import weights as W

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = lrce_b200.LRCEOpenEnded(768, 1000, 0.0, [7, 7], 1024, 5, [3], 30)
    m.load_state_dict(W.make_fusion_state_dict(1000, 30, 3, seed=0), strict=True)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(100 + rank)  # every rank its own clips
    vf = torch.randn((2, 3, 3, 49, 1024), generator=g).bfloat16().cuda()
    tf = torch.randn((2, 30, 768), generator=g).cuda()
    tgt = torch.tensor([3 + rank, 7], device="cuda")

    def grads(mode):
        m.grad_sync = mode
        for p in m.parameters():
            p.grad = None
        torch.nn.functional.cross_entropy(m(vf, tf, None), tgt).backward()
        return [p.grad.detach().clone() for p in m.parameters()]

    local = grads("none")
    for t in local:
        dist.all_reduce(t)
        t.div_(world)
    synced = grads("overlap")
    worst = max((a - b).abs().max().item() / (b.abs().max().item() + 1e-12) for a, b in zip(synced, local))
    assert worst < 1e-5, worst  # same kernels, same data: only the reduction order of the all-reduce may differ
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(str(worst))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_overlapped_allreduce_equals_allreduce_after_backward(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
