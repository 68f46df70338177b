"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of the path (clip sharding, the final logit gather, the
bucketed gradient all-reduce of the training step). The kernels themselves never communicate."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lrce_b200.dist as D  # the package imports without a GPU; only kernel calls need one

    # ---- clip sharding + logit gather: rank-major order reassembles the global batch
    g = torch.Generator().manual_seed(0)
    logits_all = torch.randn(n_clips, 7, generator=g)  # what a single process would compute
    lo, hi = D.shard_range(n_clips, rank, world)
    assert hi - lo == n_clips // world  # equal shards in this test (all_gather_into_tensor needs equal shapes)
    gathered = D.gather_logits(logits_all[lo:hi].clone())
    assert torch.equal(gathered, logits_all)
    # ragged last batch: shard_range hands out shards that differ by one; the gather pads and trims
    odd = torch.randn(n_clips - 1, 7, generator=g)
    lo2, hi2 = D.shard_range(n_clips - 1, rank, world)
    sizes = [b - a for a, b in (D.shard_range(n_clips - 1, r, world) for r in range(world))]
    assert len(set(sizes)) == 2
    assert torch.equal(D.gather_logits(odd[lo2:hi2].clone()), odd)                # sizes exchanged by the call
    assert torch.equal(D.gather_logits(odd[lo2:hi2].clone(), sizes=sizes), odd)   # sizes supplied by the caller

    # ---- gradient all-reduce: per-rank grads of a per-rank loss average to the full-batch gradient
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 7))
    x = torch.randn(n_clips, 16, generator=g)
    y = torch.randint(0, 7, (n_clips,), generator=g)
    loss = torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi])
    loss.backward()
    nbytes = D.allreduce_gradients(model.parameters(), bucket_bytes=1024)  # tiny buckets: exercise the bucketing
    assert nbytes == sum(p.numel() * 4 for p in model.parameters())
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 7))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.cross_entropy(ref(x), y).backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad - q.grad).abs().max()
    with open(os.path.join(result_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_shard_gather_allreduce_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 8, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import lrce_b200.dist as D

    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)
