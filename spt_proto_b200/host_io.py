"""Host-buffer front end of the sparse-MHA path: runs a layer's forward+backward on inputs that live
in pinned HOST memory and returns outputs/gradients into pinned host memory, chunk by chunk, with the
H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlapping the kernels of chunk i (three CUDA
streams, persistent device staging — no allocator traffic on the copy streams).  The path shards by
sequence (batch x head), so chunking along the batch dimension changes no result.

Copies: one `cudaMemcpyAsync` per chunk and direction when the host buffers come from `alloc_host()`
(the four operands of a chunk are then contiguous in host memory: q,k,v,dO of chunk i travel as one
32 MB copy instead of four 8 MB ones), else one per operand.  `bind_to_gpu_numa()` pins the calling
process (and therefore its first-touch pinned allocations) to the CPUs of the GPU's NUMA node.

This is what `bench.py` times as `e2e`: the reference-facing layer call with host buffers."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple  # noqa: F401

import torch


def bind_to_gpu_numa(device_index: int) -> Optional[dict]:
    """Best effort: restrict this process to the CPUs local to the GPU's PCIe root (sysfs `local_cpulist`), so
    that the copy-issuing thread and the pages of later pinned allocations sit on the GPU's NUMA node.  Returns
    what was found ({"numa_node", "cpus"}) or None when the platform does not say (single node, VM, no sysfs)."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(os.path.join(base, "numa_node")) as f:
            node = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"numa_node": node, "cpus": len(allowed), "bound": False}
        os.sched_setaffinity(0, use)
        return {"numa_node": node, "cpus": len(use), "bound": True}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def alloc_host(n: int, tail_shape: Sequence[int], dtype: torch.dtype, chunk: int = 1, n_operands: int = 4) -> torch.Tensor:
    """Pinned host storage for `n_operands` tensors of shape [n, *tail_shape] laid out chunk-major:
    stacked[c, j] is operand j of chunk c ([chunk, *tail]), so one chunk's operands are contiguous in host memory
    and travel as one copy.  Fill / read whole operands with fill_operand() / read_operand()."""
    if n % chunk:
        raise ValueError("alloc_host: n must be a multiple of chunk")
    return torch.empty((n // chunk, n_operands, chunk) + tuple(tail_shape), dtype=dtype).pin_memory()


def fill_operand(stacked: torch.Tensor, j: int, src: torch.Tensor) -> None:
    """stacked[:, j] <- src [n, *tail] (n = n_chunks * chunk)."""
    dst = stacked[:, j]
    dst.copy_(src.reshape(dst.shape))


def read_operand(stacked: torch.Tensor, j: int) -> torch.Tensor:
    """Operand j of a stacked buffer as a new [n, *tail] tensor."""
    v = stacked[:, j]
    return v.reshape((v.size(0) * v.size(1),) + tuple(v.shape[2:]))


class HostPipeline:
    """fwd+bwd of `layer(q, k, v)` with gradient `dy`; all four operands and all four results
    ([N, S, H, E] each) in pinned host memory.  `chunk` sequences per pipeline stage, `depth` device
    staging slots."""

    def __init__(self, layer: Callable, device: torch.device, chunk: int = 1, depth: int = 4):
        self.layer, self.device, self.chunk, self.depth = layer, device, chunk, depth
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self._slots: List[torch.Tensor] = []          # [4, chunk, S, H, E] each
        self._out_slots: List[torch.Tensor] = []
        self._slot_free: List[torch.cuda.Event] = []
        self._out_free: List[torch.cuda.Event] = []
        self._key = None
        self._pending_ev = None

    def _ensure_slots(self, tail: Tuple[int, ...], dtype: torch.dtype) -> None:
        key = (tail, dtype)
        if key == self._key:
            return
        shape = (4, self.chunk) + tail
        self._slots = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self._out_slots = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self._slot_free = [torch.cuda.Event() for _ in range(self.depth)]
        self._out_free = [torch.cuda.Event() for _ in range(self.depth)]
        cur = torch.cuda.current_stream(self.device)
        for ev in self._slot_free + self._out_free:
            ev.record(cur)
        self._key = key

    def run_stacked(self, stacked_in: torch.Tensor, stacked_out: torch.Tensor) -> None:
        """Host buffers from alloc_host(): [n_chunks, 4, chunk, S, H, E].  One copy per chunk each way."""
        n_chunks = stacked_in.size(0)
        if stacked_in.size(2) != self.chunk or stacked_in.shape != stacked_out.shape:
            raise ValueError("run_stacked: buffers must come from alloc_host(chunk=self.chunk)")
        self._run(n_chunks, lambda c, dst: dst.copy_(stacked_in[c], non_blocking=True),
                  lambda c, src: stacked_out[c].copy_(src, non_blocking=True),
                  tuple(stacked_in.shape[3:]), stacked_in.dtype, [self.chunk] * n_chunks)

    def run(self, host_in: Sequence[torch.Tensor], host_out: Sequence[torch.Tensor]) -> None:
        """Four separate host tensors [N, S, H, E] in, four out (one copy per operand, chunk and direction)."""
        n = host_in[0].size(0)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)]

        def h2d(c, dst):
            lo, hi = bounds[c]
            for j, src in enumerate(host_in):
                dst[j, : hi - lo].copy_(src[lo:hi], non_blocking=True)

        def d2h(c, src):
            lo, hi = bounds[c]
            for j, dst in enumerate(host_out):
                dst[lo:hi].copy_(src[j, : hi - lo], non_blocking=True)

        self._run(len(bounds), h2d, d2h, tuple(host_in[0].shape[1:]), host_in[0].dtype, [hi - lo for lo, hi in bounds])

    def _run(self, n_chunks, h2d, d2h, tail, dtype, sizes) -> None:
        self._ensure_slots(tail, dtype)
        main = torch.cuda.current_stream(self.device)
        ready: List[torch.cuda.Event] = [None] * n_chunks

        def stage_in(i: int) -> None:
            slot = i % self.depth
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self._slot_free[slot])     # the slot's previous consumer is done
                h2d(i, self._slots[slot])
                ev = torch.cuda.Event()
                ev.record(self.s_in)
            ready[i] = ev

        for i in range(min(self.depth, n_chunks)):
            stage_in(i)
        for i in range(n_chunks):
            slot, rows = i % self.depth, sizes[i]
            main.wait_event(ready[i])
            q, k, v, dy = (self._slots[slot][j, :rows] for j in range(4))
            q = q.detach().requires_grad_()
            k = k.detach().requires_grad_()
            v = v.detach().requires_grad_()
            y = self.layer(q, k, v)
            y.backward(dy)
            # results gathered into one persistent device buffer (4 small D2D copies) so that they leave as ONE copy
            out = self._out_slots[slot]
            main.wait_event(self._out_free[slot])               # its previous D2H copy has finished
            for j, src in enumerate((y.detach(), q.grad, k.grad, v.grad)):
                out[j, :rows].copy_(src, non_blocking=True)
            self._slot_free[slot].record(main)
            if i + self.depth < n_chunks:
                stage_in(i + self.depth)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                d2h(i, out if rows == self.chunk else out[:, :rows])
                self._out_free[slot].record(self.s_out)
        # Do not stall the compute stream on this step's D2H copies (the next run() may start its kernels while the
        # last results are still on their way to the host): the persistent output slots are protected by their own
        # events.  finish() closes a sequence of runs.
        ev = torch.cuda.Event()
        ev.record(self.s_out)
        self._pending_ev = ev

    def finish(self) -> None:
        """Order the current stream after every outstanding device-to-host copy (call before reading the host
        buffers or before recording the end of a timed region)."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.s_out)
        self._pending_ev = None


def measure_host_link(device: torch.device, nbytes: int = 256 << 20, reps: int = 4, barrier: Optional[Callable] = None
                      ) -> dict:
    """Concurrent pinned H2D + D2H bandwidth of this process's GPU (both directions at once, two streams), in
    GB/s per direction.  Call on every rank at the same time (pass `barrier`) to see what the host fabric gives
    each GPU when all of them copy — the ceiling of any host-buffer (`e2e`) number at that GPU count."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def once(both: bool, h2d: bool):
        if barrier is not None:
            barrier()
        torch.cuda.synchronize(device)
        a = torch.cuda.Event(enable_timing=True)
        b1, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(device)
        a.record(cur)
        s1.wait_event(a)
        s2.wait_event(a)
        with torch.cuda.stream(s1):
            if both or h2d:
                for _ in range(reps):
                    d_in.copy_(h_in, non_blocking=True)
            b1.record(s1)
        with torch.cuda.stream(s2):
            if both or not h2d:
                for _ in range(reps):
                    h_out.copy_(d_out, non_blocking=True)
            b2.record(s2)
        b1.synchronize()
        b2.synchronize()
        return a.elapsed_time(b1) * 1e-3, a.elapsed_time(b2) * 1e-3

    once(True, True)                                  # warm-up (page mapping, first-touch)
    t_in, t_out = once(True, True)
    t_in_only, _ = once(False, True)
    _, t_out_only = once(False, False)
    gb = nbytes * reps / 1e9
    return {"h2d_GBps_duplex": gb / t_in, "d2h_GBps_duplex": gb / t_out, "h2d_GBps_alone": gb / t_in_only,
            "d2h_GBps_alone": gb / t_out_only, "bytes_per_copy": nbytes, "reps": reps}
