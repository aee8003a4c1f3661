"""Multi-GPU plumbing.  The hot path has no exchange step: sparse MHA shards by batch x head, the
routed FFN by token — every rank simply runs the same kernels on its own slice.
The only collective of the reference's fine-tuning step is DDP's gradient all-reduce of the trainable
parameters (script/4-sparse-tuning-0.py:183-187 via Lightning).  Two forms of it live here, on
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests):

  * `GradReducer`  — the step's form: the gradients of the trainable parameters (LoRA factors, routers,
    PQ codebooks: ~20 MB for 4 LLaMA-7B-shape layers) live in ONE persistent flat buffer per dtype
    (`p.grad` are views into it: no torch.cat, no copy-back), cut into a few buckets in backward order;
    a post-accumulate-grad hook launches a bucket's all-reduce (async, on NCCL's stream) as soon as its
    last gradient has been accumulated, so the collective runs under the rest of the backward pass.
    Buckets are sized for launch latency, not link count (NVSwitch).
  * `allreduce_grads` — the plain post-backward form (kept for callers without hooks).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

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


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _avg_supported() -> bool:
    return dist.get_backend() == "nccl"


def _all_reduce_mean(t: torch.Tensor, async_op: bool = False):
    """Mean over ranks in place.  NCCL averages inside the collective; gloo sums and the division follows."""
    if _avg_supported():
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op), False
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op), True


class GradReducer:
    """Flat, bucketed, backward-overlapped gradient all-reduce (mean) of `params`.

        reducer = GradReducer(trainable, n_buckets=4)
        ...
        reducer.zero_grad()          # instead of optimizer.zero_grad(set_to_none=True): keeps the views
        loss.backward()              # hooks launch each bucket's all-reduce as it completes
        reducer.finish()             # waits for the collectives (and reduces buckets no hook completed)
        optimizer.step()

    Parameters are laid out in REVERSE registration order (the order in which backward produces their
    gradients), so bucket 0 is complete first.  With world size 1 nothing is communicated; the flat
    buffer and views are still set up so that the step code is the same at every N."""

    def __init__(self, params: Sequence[torch.nn.Parameter], n_buckets: int = 4, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.overlap = overlap
        self.world = _world()
        order = list(reversed(self.params))
        self.flats: List[torch.Tensor] = []
        self.buckets: List[Tuple[torch.Tensor, List[torch.nn.Parameter]]] = []
        by_dtype = {}
        for p in order:
            by_dtype.setdefault((p.dtype, p.device), []).append(p)
        for (dtype, device), group in by_dtype.items():
            total = sum(p.numel() for p in group)
            flat = torch.zeros(total, dtype=dtype, device=device)
            self.flats.append(flat)
            target = max(1, -(-total // max(1, n_buckets)))
            off = start = 0
            members: List[torch.nn.Parameter] = []
            for p in group:
                p.grad = flat[off: off + p.numel()].view_as(p)
                off += p.numel()
                members.append(p)
                if off - start >= target:
                    self.buckets.append((flat[start:off], members))
                    start, members = off, []
            if members:
                self.buckets.append((flat[start:off], members))
        self._bucket_of = {}
        for bi, (_, members) in enumerate(self.buckets):
            for p in members:
                self._bucket_of[id(p)] = bi
        self._pending = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works: List[Tuple[object, torch.Tensor, bool]] = []
        self._handles = []
        if self.overlap:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self._reset()

    # -- bookkeeping ---------------------------------------------------------------------------------
    @property
    def n_buckets(self) -> int:
        return len(self.buckets)

    @property
    def nbytes(self) -> int:
        return sum(f.numel() * f.element_size() for f in self.flats)

    def _reset(self) -> None:
        for bi, (_, members) in enumerate(self.buckets):
            self._pending[bi] = len(members)
            self._launched[bi] = False
        self._works = []

    def zero_grad(self) -> None:
        """Zero the flat buffers (the `.grad` views stay attached) and re-arm the hooks."""
        for f in self.flats:
            f.zero_()
        for p in self.params:      # an optimizer / user may have detached a view (set_to_none): re-attach
            if p.grad is None:
                raise RuntimeError("GradReducer: a .grad view was dropped; call reducer.zero_grad(), "
                                   "not optimizer.zero_grad(set_to_none=True)")
        self._reset()

    def _launch(self, bi: int) -> None:
        if self._launched[bi]:
            return
        self._launched[bi] = True
        if self.world == 1:
            return
        buf = self.buckets[bi][0]
        work, divide = _all_reduce_mean(buf, async_op=True)
        self._works.append((work, buf, divide))

    def _hook(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> int:
        """Launch what no hook completed (parameters without a gradient this step, or overlap=False), then make
        the current stream wait for every bucket.  Returns the number of collectives of this step."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        n = len(self._works)
        for work, buf, divide in self._works:
            work.wait()
            if divide:
                buf /= self.world
        self._works = []
        return n

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Post-backward flat-bucket all-reduce of the gradients of the trainable parameters.  Returns the number of
    collectives issued.  (The fine-tuning step uses GradReducer, which overlaps the collective with backward.)"""
    if _world() == 1:
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
