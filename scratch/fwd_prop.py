import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spt_proto_b200 import ext
DEV = "cuda"
for B in (1, 4, 32, 128):
    S, d = 2048, 64
    g = torch.Generator().manual_seed(5)
    q = torch.randn(B, S, d, generator=g).bfloat16().to(DEV)
    k = torch.randn(B, S, d, generator=g).bfloat16().to(DEV)
    w = torch.randn(8, 16, 8, generator=g).to(DEV)
    mask, extra0, _ = ext.lookup_mask(ext.pq_encode(q, w), ext.pq_encode(k, w), 8)
    v = torch.ones(B, S, d, device=DEV, dtype=torch.bfloat16)
    bad = []
    for rep in range(5):
        y, z = ext.sparse_attn_fwd(q, k, v, mask, extra0, d ** -0.5)
        err = (y.float() - 1).abs()
        bad.append((int((err > 1e-2).sum()), float(err.max())))
    rows = (err > 1e-2).any(-1).nonzero()
    print(B, bad, rows[:8].tolist())
