"""Experiment: the MHA step (configs[1], 4 sequences) as ONE call vs two half-batches on two streams / four quarter-batches
on two or four streams — do kernels with different bottlenecks (integer lookup, XU-bound attention) fill each other's gaps?"""
import sys, json
import torch
sys.path.insert(0, ".")
from spt_proto_b200 import layers

dev = torch.device("cuda:0")
torch.manual_seed(1234)
attn = layers.SparseVanillaAttentionV2(d_head=64, d_codeword=8, n_codewords=16, p_dropout=0.0).to(dev)
attn.sparse_coeff = 8
attn.host_trigger = False
n_seq, S, H, E = 4, 2048, 32, 64


def make(n):
    t = [torch.randn(n, S, H, E, device=dev).bfloat16().requires_grad_() for _ in range(3)]
    return t + [torch.randn(n, S, H, E, device=dev).bfloat16()]


def step_of(ts):
    q, k, v, dy = ts

    def f():
        q.grad = k.grad = v.grad = None
        attn(q, k, v).backward(dy)
    return f


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {}
whole = step_of(make(n_seq))
out["one_call_ms"] = timeit(whole)
for parts, n_streams in ((2, 2), (4, 2), (4, 4), (2, 1), (4, 1)):
    steps = [step_of(make(n_seq // parts)) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    main = torch.cuda.current_stream()

    def multi():
        if n_streams == 1:
            for s in steps:
                s()
            return
        for st in streams:
            st.wait_stream(main)
        for i, s in enumerate(steps):
            with torch.cuda.stream(streams[i % n_streams]):
                s()
        for st in streams:
            main.wait_stream(st)
    out[f"{parts}_parts_{n_streams}_streams_ms"] = timeit(multi)
print(json.dumps(out))
