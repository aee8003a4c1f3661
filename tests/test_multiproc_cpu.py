"""world_size-2 gloo tests of the N > 1 host logic (run on CPU): unit sharding is exact and disjoint,
and the fine-tuning step's gradient all-reduce of the trainable parameters averages across ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spt_proto_b200.distributed import allreduce_grads, shard_range

    # (1) batch x head sharding: every head owned exactly once
    begin, end = shard_range(32 * 3, rank, world)
    owned = torch.zeros(96)
    owned[begin:end] = 1
    dist.all_reduce(owned)
    ok_shard = bool((owned == 1).all())

    # (2) gradient all-reduce of trainable parameters only (frozen base weights are skipped)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    for p in model[0].parameters():
        p.requires_grad = False
    x = torch.full((2, 8), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model[1].parameters()]
    n_calls = allreduce_grads(model.parameters(), bucket_bytes=64)
    gathered = [torch.zeros_like(local[0]) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    want = sum(gathered) / world
    ok_grad = torch.allclose(model[1].weight.grad, want) and model[0].weight.grad is None

    # (3) GradReducer: flat persistent buffer, bucketed, launched from backward hooks; two steps (the second checks
    # that zero_grad() re-arms the hooks and that the views survive), mixed dtypes, one parameter without a gradient
    from spt_proto_b200.distributed import GradReducer
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 4))
    extra = torch.nn.Parameter(torch.ones(16, dtype=torch.float64))   # a second dtype group
    for p in net[0].parameters():
        p.requires_grad = False
    unused = torch.nn.Parameter(torch.ones(3))                  # never reached by backward: reduced by finish()
    trainable = [p for p in net.parameters() if p.requires_grad] + [extra, unused]
    reducer = GradReducer(trainable, n_buckets=3)
    ok_reducer = reducer.n_buckets >= 2
    for step in range(2):
        xs = torch.full((2, 8), float(rank + 1 + step))
        reducer.zero_grad()
        h = net[1](net[0](xs)) * extra.float()
        loss = net[2](h.float()).sum() * (rank + 1)
        # this rank's own gradients, taken WITHOUT accumulation (no hooks fire): once backward() runs, the hooks reduce
        # the flat buffer in place and asynchronously, so p.grad must not be read between backward() and finish()
        own = torch.autograd.grad(loss, trainable, allow_unused=True, retain_graph=True)
        local = {id(p): (torch.zeros_like(p) if g is None else g.clone()) for p, g in zip(trainable, own)}
        loss.backward()
        n_coll = reducer.finish()
        ok_reducer &= n_coll == reducer.n_buckets
        for p in trainable:
            parts = [torch.zeros_like(local[id(p)]) for _ in range(world)]
            dist.all_gather(parts, local[id(p)])
            ok_reducer &= torch.allclose(p.grad, sum(parts) / world)
            ok_reducer &= any(p.grad.data_ptr() >= f.data_ptr() and
                              p.grad.data_ptr() < f.data_ptr() + f.numel() * f.element_size() for f in reducer.flats)
    reducer.remove()

    # (4) the non-overlapped form the 2-GPU measurements compare against: no hooks, ONE bucket,
    # everything reduced by finish() after backward
    torch.manual_seed(2)
    lin = torch.nn.Linear(6, 3)
    plain = GradReducer(list(lin.parameters()), n_buckets=1, overlap=False)
    plain.zero_grad()
    loss = lin(torch.full((4, 6), float(rank + 1))).sum() * (rank + 2)
    own = torch.autograd.grad(loss, list(lin.parameters()), retain_graph=True)
    loss.backward()
    ok_plain = plain.n_buckets == 1 and plain.finish() == 1
    for p, g in zip(lin.parameters(), own):
        parts = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(parts, g.clone())
        ok_plain &= torch.allclose(p.grad, sum(parts) / world)
    ret[rank] = (ok_shard, ok_grad, n_calls, bool(ok_reducer) and bool(ok_plain))
    dist.destroy_process_group()


def test_two_rank_sharding_and_grad_allreduce():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    for attempt in range(3):          # the probed port can be taken between the probe and the rendezvous: retry on a new one
        ret.clear()
        try:
            mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
            break
        except Exception:
            if attempt == 2:
                raise
    assert len(ret) == world
    for rank in range(world):
        ok_shard, ok_grad, n_calls, ok_reducer = ret[rank]
        assert ok_shard and ok_grad and n_calls >= 1 and ok_reducer
