"""configs[4] corner points: fused (dense tiles) vs stage (gathered) path, fwd+bwd ms per 8192 tokens."""
import sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import layers
dev = 'cuda'
attn = layers.SparseVanillaAttentionV2(d_head=64, d_codeword=8, n_codewords=16, p_dropout=0.0).to(dev)
attn.host_trigger = False
for S in (2048, 4096, 8192):
    n_seq = max(1, 8192 // S)
    q, k, v = (torch.randn(n_seq, S, 32, 64, device=dev).bfloat16().requires_grad_() for _ in range(3))
    dy = torch.randn(n_seq, S, 32, 64, device=dev).bfloat16()
    for topk in (16, 32, 64, 128, 256):
        attn.sparse_coeff = S // topk
        row = []
        for fused in (True, False):
            attn.use_fused = fused
            def step():
                q.grad = k.grad = v.grad = None
                attn(q, k, v).backward(dy)
            for _ in range(2): step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3): step()
            b.record(); torch.cuda.synchronize()
            row.append(a.elapsed_time(b) / 3)
        print(f"S {S} k {topk}: fused {row[0]:.3f} ms  stage {row[1]:.3f} ms", flush=True)
    del q, k, v, dy
