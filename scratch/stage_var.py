"""Per-step times of the stage path (reference-style kernels) at the bench shape: distribution over 30 steps."""
import sys, time; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import layers
dev = 'cuda'
attn = layers.SparseVanillaAttentionV2(d_head=64, d_codeword=8, n_codewords=16, p_dropout=0.0).to(dev)
attn.host_trigger = False
attn.use_fused = False
q, k, v = (torch.randn(4, 2048, 32, 64, device=dev).bfloat16().requires_grad_() for _ in range(3))
dy = torch.randn(4, 2048, 32, 64, device=dev).bfloat16()
def step():
    q.grad = k.grad = v.grad = None
    attn(q, k, v).backward(dy)
for _ in range(3): step()
torch.cuda.synchronize()
ts, cs = [], []
for i in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); step(); b.record()
    t1 = time.perf_counter()
    b.synchronize()
    ts.append(a.elapsed_time(b)); cs.append((t1 - t0) * 1e3)
print("gpu ms per step:", " ".join(f"{t:.1f}" for t in ts))
print("cpu issue ms   :", " ".join(f"{t:.1f}" for t in cs))
print("mem GB", torch.cuda.max_memory_allocated() / 1e9, "reserved", torch.cuda.memory_reserved() / 1e9)
