"""Multi-GPU plumbing.  The hot path has no exchange step: sparse MHA shards by batch x head, the
routed FFN by token — every rank simply runs the same kernels on its own slice.
The only collective of the reference's fine-tuning step is DDP's gradient all-reduce of the trainable
parameters (script/4-sparse-tuning-0.py:183-187 via Lightning); `allreduce_grads` is that step on
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of n_items units (heads, sequences or tokens) owned by `rank`;
    sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def trainable_grads(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    return [p.grad for p in params if p.requires_grad and p.grad is not None]


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Flat-bucket all-reduce of the gradients of the trainable parameters (LoRA factors, routers, PQ
    codebooks: ~20 MB fp32 for 4 LLaMA-7B-shape layers, so one or two buckets).  Buckets are sized for
    launch latency, not link count (NVSwitch).  Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    grads = trainable_grads(params)
    world = dist.get_world_size()
    n_calls, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, n_calls
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1).float() for g in bucket])
        dist.all_reduce(flat)
        if average:
            flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
        n_calls += 1

    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return n_calls
