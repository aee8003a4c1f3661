"""Host-buffer front end of the sparse-MHA path: runs a layer's forward+backward on inputs that live
in pinned HOST memory and returns outputs/gradients into pinned host memory, sequence by sequence,
with the H2D copy of sequence i+1 and the D2H copy of sequence i-1 overlapping the kernels of
sequence i (three CUDA streams, double-buffered device staging).  The path shards by sequence
(batch x head), so chunking along the batch dimension changes no result.

This is what `bench.py` times as `e2e`: the reference-facing layer call with host buffers."""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class HostPipeline:
    """fwd+bwd of `layer(q, k, v)` with gradient `dy`, all four operands and all four results
    ([N, S, H, E] each) in pinned host memory.  `chunk` sequences are processed per stage."""

    def __init__(self, layer: Callable, device: torch.device, chunk: int = 1):
        self.layer, self.device, self.chunk = layer, device, chunk
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)

    def run(self, host_in: Sequence[torch.Tensor], host_out: Sequence[torch.Tensor]) -> None:
        hq, hk, hv, hdy = host_in
        n = hq.size(0)
        main = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(main)
        self.s_out.wait_stream(main)
        staged = []
        # H2D of every chunk is queued up front on its own stream: copies run back to back while the
        # compute stream consumes chunks as their events fire
        for lo in range(0, n, self.chunk):
            hi = min(n, lo + self.chunk)
            with torch.cuda.stream(self.s_in):
                dev = [t[lo:hi].to(self.device, non_blocking=True) for t in (hq, hk, hv, hdy)]
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            for t in dev:
                t.record_stream(main)
            staged.append((lo, hi, dev, ev))
        for lo, hi, (q, k, v, dy), ev in staged:
            main.wait_event(ev)
            q.requires_grad_()
            k.requires_grad_()
            v.requires_grad_()
            y = self.layer(q, k, v)
            y.backward(dy)
            done = torch.cuda.Event()
            done.record(main)
            results = (y.detach(), q.grad, k.grad, v.grad)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for dst, src in zip(host_out, results):
                    src.record_stream(self.s_out)
                    dst[lo:hi].copy_(src, non_blocking=True)
        main.wait_stream(self.s_out)
