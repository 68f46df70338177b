"""Multi-GPU plumbing of the LRCE hot path: one process per GPU, clips sharded across ranks, weights
replicated. The forward has NO data-path collective — every clip is independent — so evaluation needs exactly one
``all_gather`` of the fp32 logits at the end, and the training step one bucketed ``all_reduce`` of the cross-modal
encoder's gradients (the only trainable part in BASELINE.json's config 5). Both run on whatever backend the process
group was created with: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of `n_items` clips for `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_logits(local: torch.Tensor, group=None, sizes: Sequence[int] = None) -> torch.Tensor:
    """all_gather of per-rank logits `(B_rank, ...)` -> `(sum B_rank, ...)`, rank-major (the order `shard_range` hands clips
    out in). `sizes` = the per-rank batch sizes when they are known to differ (a ragged last evaluation batch:
    `shard_range` shards differ by one): every rank then pads to the largest shard, one `all_gather_into_tensor` moves the
    padded blocks and the padding is trimmed. `sizes=None` means equal shards (checked cheaply: a mismatch would corrupt
    the collective, so the sizes are exchanged once when the caller cannot vouch for them — pass `sizes` to skip that).
    A no-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if sizes is None:
        mine = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        allsz = torch.empty(world, device=local.device, dtype=torch.int64)
        dist.all_gather_into_tensor(allsz, mine, group=group)
        sizes = allsz.tolist()
    if len(sizes) != world or sizes[dist.get_rank(group)] != local.shape[0]:
        raise ValueError(f"gather_logits: sizes {list(sizes)} do not describe this rank's batch of {local.shape[0]}")
    big = max(sizes)
    block = local.contiguous()
    if block.shape[0] != big:
        block = torch.cat([block, block.new_zeros((big - block.shape[0],) + tuple(block.shape[1:]))])
    out = local.new_empty((world * big,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, block, group=group)
    if all(s == big for s in sizes):
        return out
    return torch.cat([out[r * big:r * big + s] for r, s in enumerate(sizes)])


def _buckets(params: Sequence[torch.nn.Parameter], bucket_bytes: int) -> List[List[torch.nn.Parameter]]:
    out, cur, size = [], [], 0
    for p in params:
        nbytes = p.grad.numel() * p.grad.element_size()
        if cur and (size + nbytes > bucket_bytes or p.grad.dtype != cur[0].grad.dtype):
            out.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nbytes
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 64 << 20) -> int:
    """Average `.grad` of `params` over the process group with a few large flattened all_reduce calls (launch-latency-
    bound collectives on NVSwitch: size buckets for overlap, not for link count). Asynchronous handles are issued for
    every bucket before the first is waited on. Returns the number of gradient bytes reduced."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sum(p.grad.numel() * p.grad.element_size() for p in params)
    world = dist.get_world_size(group)
    work = []
    for bucket in _buckets(params, bucket_bytes):
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        work.append((bucket, flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)))
    total = 0
    for bucket, flat, handle in work:
        handle.wait()
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        total += flat.numel() * flat.element_size()
    return total
