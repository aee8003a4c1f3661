"""configs[4] corner points: per-kernel-group times of the fused path (ms per 8192 tokens)."""
import sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import ext
dev = 'cuda'
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
w = torch.randn(8, 16, 8, device=dev)
for S in (2048, 4096, 8192):
    n_seq = max(1, 8192 // S)
    q, k, v, dy = (torch.randn(n_seq, S, 32, 64, device=dev).bfloat16() for _ in range(4))
    for topk in (16, 256):
        coeff = S // topk
        qc, kc = ext.pq_encode_pair(q, k, w)
        mask, extra0, _ = ext.lookup_mask(qc, kc, coeff)
        y, z = ext.sparse_attn_fwd(q, k, v, mask, extra0, 0.125)
        print(f"S {S} k {topk}: encode {ev(lambda: ext.pq_encode_pair(q, k, w)):.3f}  lookup {ev(lambda: ext.lookup_mask(qc, kc, coeff)):.3f}"
              f"  fwd {ev(lambda: ext.sparse_attn_fwd(q, k, v, mask, extra0, 0.125)):.3f}  bwd {ev(lambda: ext.sparse_attn_bwd(q, k, v, y, dy, mask, extra0, z, 0.125)):.3f}"
              f"  mask density {float(sum(bin(x).count('1') for x in mask.flatten()[:4096].tolist())) / (4096 * 32):.4f}", flush=True)
