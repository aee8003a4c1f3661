#!/usr/bin/env python
"""BASELINE configs[4]: sparse-MHA sweep, seq 512-8192 x top-k 16-256, fwd+bwd tokens/s per GPU.

    python scripts/sweep.py [--cpu]        # --cpu adds the oracle port (CPU) at the smallest points
One JSON line per (S, k): 32 heads x d_head 64, bf16, PQ 8x16, sparse_coeff = S / k, sequences per step
chosen so that a step holds 8192 tokens.  Under torchrun every rank runs the same sweep on its own
sequences (weak scaling, no collective) and rank 0 reports the aggregate with the max-over-ranks time."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HEADS, D_HEAD = 32, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    from spt_proto_b200 import layers

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    attn = layers.SparseVanillaAttentionV2(d_head=D_HEAD, d_codeword=8, n_codewords=16, p_dropout=0.0).to(dev)
    attn.host_trigger = False
    torch.manual_seed(1234 + rank)
    for S in (512, 1024, 2048, 4096, 8192):
        n_seq = max(1, 8192 // S)
        q, k, v = (torch.randn(n_seq, S, HEADS, D_HEAD, device=dev).bfloat16().requires_grad_() for _ in range(3))
        dy = torch.randn(n_seq, S, HEADS, D_HEAD, device=dev).bfloat16()
        for topk in (16, 32, 64, 128, 256):
            attn.sparse_coeff = S // topk

            def step():
                q.grad = k.grad = v.grad = None
                attn(q, k, v).backward(dy)

            for _ in range(3):
                step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.iters):
                step()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b) * 1e-3 / args.iters
            if world > 1:
                tt = torch.tensor([t], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = tt.item()
            line = {"metric": "sparse_mha_fwd_bwd_tokens_per_s", "S": S, "top_k": topk, "sparse_coeff": S // topk,
                    "n_gpus": world, "seqs_per_gpu": n_seq, "ms_per_step": t * 1e3, "value": n_seq * S * world / t,
                    "unit": "tokens/s", "path": "fused" if attn._fused_ok(q) else "stage"}
            if args.cpu and rank == 0 and S <= 1024 and topk == 32:
                from oracle import spt_oracle as O
                torch.set_num_threads(os.cpu_count() or 1)
                g = torch.Generator().manual_seed(1)
                mk = lambda: torch.randn(1, S, 8, D_HEAD, generator=g).bfloat16().float().requires_grad_()
                qc, kc, vc = mk(), mk(), mk()
                w = torch.randn(8, 16, 8, generator=g)
                t0 = time.perf_counter()
                y = O.sparse_mha_layer(qc, kc, vc, w, S // topk)
                y.backward(torch.ones_like(y))
                tc = time.perf_counter() - t0
                line["cpu_port_tokens_per_s"] = S * 8 / HEADS / tc
                line["cpu_cores"] = os.cpu_count()
            if rank == 0:
                print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
