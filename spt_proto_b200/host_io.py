"""Host-buffer front end of the sparse-MHA path: runs a layer's forward+backward on inputs that live
in pinned HOST memory and returns outputs/gradients into pinned host memory, chunk by chunk, with the
H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlapping the kernels of chunk i (three CUDA
streams, persistent double-buffered device staging — no allocator traffic on the copy streams).  The
path shards by sequence (batch x head), so chunking along the batch dimension
changes no result.

This is what `bench.py` times as `e2e`: the reference-facing layer call with host buffers."""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch


class HostPipeline:
    """fwd+bwd of `layer(q, k, v)` with gradient `dy`; all four operands and all four results
    ([N, S, H, E] each) in pinned host memory.  `chunk` sequences per pipeline stage, `depth` device
    staging slots."""

    def __init__(self, layer: Callable, device: torch.device, chunk: int = 1, depth: int = 4):
        self.layer, self.device, self.chunk, self.depth = layer, device, chunk, depth
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self._slots: List[List[torch.Tensor]] = []
        self._slot_free: List[torch.cuda.Event] = []
        self._key = None
        self._pending = None          # results of the previous run(): referenced until their D2H copies are done
        self._pending_ev = None

    def _ensure_slots(self, like: torch.Tensor) -> None:
        key = (tuple(like.shape[1:]), like.dtype)
        if key == self._key:
            return
        shape = (self.chunk,) + tuple(like.shape[1:])
        self._slots = [[torch.empty(shape, dtype=like.dtype, device=self.device) for _ in range(4)]
                       for _ in range(self.depth)]
        self._slot_free = [torch.cuda.Event() for _ in range(self.depth)]
        for ev in self._slot_free:
            ev.record(torch.cuda.current_stream(self.device))
        self._key = key

    def run(self, host_in: Sequence[torch.Tensor], host_out: Sequence[torch.Tensor]) -> None:
        hq = host_in[0]
        n = hq.size(0)
        self._ensure_slots(hq)
        main = torch.cuda.current_stream(self.device)
        chunks = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)]
        ready: List[torch.cuda.Event] = [None] * len(chunks)

        def stage_in(i: int) -> None:
            lo, hi = chunks[i]
            slot = i % self.depth
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self._slot_free[slot])     # the slot's previous consumer is done
                for dst, src in zip(self._slots[slot], host_in):
                    dst[: hi - lo].copy_(src[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            ready[i] = ev

        for i in range(min(self.depth, len(chunks))):
            stage_in(i)
        keep = []   # results stay referenced until the final wait_stream orders main after the D2H copies
        for i, (lo, hi) in enumerate(chunks):
            slot = i % self.depth
            main.wait_event(ready[i])
            q, k, v, dy = (t[: hi - lo] for t in self._slots[slot])
            q = q.detach().requires_grad_()
            k = k.detach().requires_grad_()
            v = v.detach().requires_grad_()
            y = self.layer(q, k, v)
            y.backward(dy)
            self._slot_free[slot].record(main)
            if i + self.depth < len(chunks):
                stage_in(i + self.depth)
            done = torch.cuda.Event()
            done.record(main)
            results = (y.detach(), q.grad, k.grad, v.grad)
            keep.append(results)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for dst, src in zip(host_out, results):
                    dst[lo:hi].copy_(src, non_blocking=True)
        # Do not stall the compute stream on this step's D2H copies (the next run() may start its kernels while the
        # last results are still on their way to the host): keep the result tensors referenced and order the compute
        # stream only after the copies of the PREVIOUS run, which are long finished by now.  finish() closes a sequence.
        ev = torch.cuda.Event()
        ev.record(self.s_out)
        if self._pending_ev is not None:
            main.wait_event(self._pending_ev)
        self._pending, self._pending_ev = keep, ev

    def finish(self) -> None:
        """Order the current stream after every outstanding device-to-host copy (call before reading the host
        buffers or before recording the end of a timed region) and drop the references to the last results."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.s_out)
        self._pending, self._pending_ev = None, None
