#!/usr/bin/env python
"""BASELINE configs[3]: LLaMA-7B-shape 4-layer SPT fine-tuning step, batch-sharded over N GPUs.

    python scripts/finetune_step.py --steps 5 --warmup 2 --seq 2048
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/finetune_step.py ...

Each rank owns `--batch` sequences (weak scaling).  A step = forward + backward through 4 pre-norm
`TransformerBlock`s (d_model 4096, 32 heads x d_head 128, d_ff 11008) built dense and then upgraded by
`ModuleUpgrader` + `SparseLoRAHandler` exactly as the reference does (script/0-profile.py:87-143,182-189 +
utils/adapter.py:94-97,155-184): frozen base weights, rank-16 LoRA
on q/k/v/o and gate/side/down, SparseRotaryAttentionV2 (PQ 16 subspaces x 16 codewords, top-k S/8,
PQ training loss armed every step like script/4-sparse-tuning-0.py:71-91), LoRARoutedLLaMaFFN
(block = d_ff/4, half the blocks active); then the bucketed NCCL all-reduce of the trainable gradients, launched from
backward hooks on a persistent flat buffer (spt_proto_b200.distributed.GradReducer), grad-clip 1.0 and AdamW.  Prints one JSON line (rank 0).

The attention runs the fused tcgen05 kernels (head dim 128 instantiation), the PQ loss the fused
pq_train kernels, the FFN the grouped GEMM."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch import nn  # noqa: E402


def build_model(dev, n_layers: int, d_lora: int):
    """4 dense LLaMA-7B-shape blocks, then the reference's four-pass upgrade (script/0-profile.py:87-143,182-189)."""
    from spt_proto_b200 import layers, utils

    d_model, n_heads, d_ff = 4096, 32, 11008
    torch.manual_seed(1234)                       # same weights on every rank (DDP starts from a broadcast)
    torch.set_default_device(dev)
    try:
        d_head = d_model // n_heads
        model = nn.Sequential(*[layers.TransformerBlock(
            d_model=d_model, n_heads=n_heads, layernorm_fn=layers.LlamaRMSNorm(d_model),
            attention_fn=layers.RotaryAttention(d_head=d_head, p_dropout=0.0),
            feedforward_fn=layers.LLaMaFeedforward(d_model=d_model, d_feedforward=d_ff, activation=nn.SiLU()),
            attention_bias=False, pre_norm=True) for _ in range(n_layers)])
        for stage in ("lora", "ffn", "mha_v1", "mha_v2"):
            model = utils.ModuleUpgrader(utils.SparseLoRAHandler(d_lora=d_lora, stage=stage, verbose=False)).visit(model)
    finally:
        torch.set_default_device("cpu")
    model = model.to(dev).bfloat16()
    for blk in model:
        nn.init.normal_(blk.mha.linear_q.lora.right.weight, std=0.02)   # non-zero LoRA so that every gradient is exercised
        nn.init.normal_(blk.ffd.down.lora.right.weight, std=0.02)
    return model, (d_model, n_heads, d_ff)


def run(dev, rank: int, world: int, steps: int = 5, warmup: int = 2, seq: int = 2048, batch: int = 1, n_layers: int = 4,
        d_lora: int = 16, graph: bool = True, overlap: bool = True, n_buckets: int = 4, data_seed=None) -> dict:
    """Times the step on this rank's GPU (collective when world > 1: torch.distributed must be initialised) and
    returns the result line as a dict (identical on every rank; times are the max over ranks)."""
    import torch.distributed as dist

    from spt_proto_b200 import ext
    from spt_proto_b200.distributed import GradReducer

    model, (d_model, n_heads, d_ff) = build_model(dev, n_layers, d_lora)
    trainable = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in trainable)
    reducer = GradReducer(trainable, n_buckets=n_buckets, overlap=overlap)
    opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=1e-2, capturable=graph)
    torch.manual_seed(1234 + rank if data_seed is None else data_seed)   # every rank its own batch
    x = torch.randn(batch, seq, d_model, device=dev).bfloat16()
    target = torch.randn(batch, seq, d_model, device=dev).bfloat16()

    def clip_(max_norm: float = 1.0):
        # nn.utils.clip_grad_norm_ on the flat buffers: two kernels per dtype group instead of ~4 per parameter
        total = torch.sqrt(sum(f.float().pow(2).sum() for f in reducer.flats))
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        for f in reducer.flats:
            f.mul_(coef.to(f.dtype))

    def step():
        for blk in model:
            blk.mha.attn_fn.host_trigger = True           # arm the PQ loss (no device->host sync)
        y = model(x)
        loss = nn.functional.mse_loss(y.float(), target.float())
        loss = loss + 1e-2 * sum(blk.mha.attn_fn.loss for blk in model)
        reducer.zero_grad()                                # flat buffer zeroed; .grad stay views into it
        loss.backward()                                    # hooks launch each bucket's all-reduce as it completes
        n_coll = reducer.finish()
        clip_()
        opt.step()
        return loss, n_coll

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if graph:
        # whole-network capture recipe: warm up on a side stream so that no autograd node (AccumulateGrad of the
        # parameters in particular) stays tied to the legacy default stream, which cannot be joined during capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 3)):
                loss, n_coll = step()
        torch.cuda.current_stream(dev).wait_stream(side)
    else:
        for _ in range(warmup):
            loss, n_coll = step()
    barrier()
    captured = 0
    if graph:
        # ~1300 small launches per step: replaying one captured graph removes the launch gaps.  Nothing on the path
        # synchronises with the host (host_trigger, device-side bucketing), so the whole step is capturable —
        # the bucketed NCCL all-reduces included (they fork onto NCCL's stream inside the capture and join in finish()).
        eager_step = step
        g = torch.cuda.CUDAGraph()
        captured0 = ext.launch_count()
        with torch.cuda.graph(g):
            loss, n_coll = eager_step()
        captured = ext.launch_count() - captured0     # launches of libspt_b200 kernels recorded in the graph

        def step():
            g.replay()
            return loss, n_coll

        for _ in range(2):
            step()
        barrier()
    launches0 = ext.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss, n_coll = step()
    b.record()
    barrier()
    elapsed = a.elapsed_time(b) * 1e-3
    # the collective alone (not overlapped with anything): the same buckets, back to back
    ar_ms = 0.0
    if world > 1:
        barrier()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a2.record()
        for _ in range(reps):
            for buf, _ in reducer.buckets:
                dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        b2.record()
        barrier()
        ar_ms = a2.elapsed_time(b2) / reps
        t = torch.tensor([elapsed, ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, ar_ms = t.tolist()
    tokens = batch * seq * world
    loss_val = float(loss.detach())
    reducer.remove()
    return {"metric": "spt_finetune_step_tokens_per_s", "value": tokens * steps / elapsed, "unit": "tokens/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"LLaMA-7B-shape {n_layers}-layer SPT fine-tuning step (sparse rotary MHA d_head 128 "
                                   f"+ LoRA routed FFN), seq {seq}, {batch} seq/GPU",
                       "d_model": d_model, "n_heads": n_heads, "d_ff": d_ff, "d_lora": d_lora,
                       "trainable_params": n_train, "allreduce_calls_per_step": n_coll,
                       "allreduce_bytes_per_step": reducer.nbytes,
                       "allreduce": ("bucketed, launched from backward hooks on a persistent flat gradient buffer "
                                     "(overlapped with backward)" if overlap else "after backward"),
                       "parallelism": f"dp{world} + NCCL all-reduce of trainable grads", "cuda_graph": bool(graph)},
            "allreduce_ms_standalone": ar_ms,
            "loss": loss_val,
            # a replayed graph launches its captured kernels without passing through the library's counter
            "gpu_launches": (captured * steps) if graph else ext.launch_count() - launches0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=1, help="sequences per GPU per step")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--d-lora", type=int, default=16)
    ap.add_argument("--graph", action="store_true",
                    help="capture forward + backward + all-reduce + optimizer of one step in a CUDA graph and replay it")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after backward instead of from backward hooks")
    ap.add_argument("--buckets", type=int, default=4)
    ap.add_argument("--data-seed", type=int, default=None,
                    help="seed of this rank's batch (default 1234 + rank): run one GPU on another rank's data")
    args = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    line = run(dev, rank, world, steps=args.steps, warmup=args.warmup, seq=args.seq, batch=args.batch,
               n_layers=args.layers, d_lora=args.d_lora, graph=args.graph, overlap=not args.no_overlap,
               n_buckets=args.buckets, data_seed=args.data_seed)
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
