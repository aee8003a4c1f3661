"""Fused attention fwd / bwd on the layer's interleaved [N, S, H, E] layout vs head-major [B, S, E] (same data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spt_proto_b200 import ext
N, S, H, E = 4, 2048, 32, 64
g = torch.Generator().manual_seed(1)
q4, k4, v4, dy4 = (torch.randn(N, S, H, E, generator=g).to("cuda", torch.bfloat16) for _ in range(4))
w = torch.randn(8, 16, 8, generator=g).cuda()
def ev(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
for name, conv in (("interleaved [N,S,H,E]", lambda t: t), ("head-major [B,S,E]", lambda t: t.transpose(1, 2).contiguous().view(N * H, S, E))):
    q, k, v, dy = (conv(t) for t in (q4, k4, v4, dy4))
    mask, extra0, _ = ext.lookup_mask(ext.pq_encode(q, w), ext.pq_encode(k, w), 8)
    y, z = ext.sparse_attn_fwd(q, k, v, mask, extra0, E ** -0.5)
    yr, zr = ext.sparse_attn_fwd(q, k, v, mask, extra0, E ** -0.5, reference_layout=True)
    print(name, "REFERENCE output layout: fwd ms %.4f" % ev(lambda: ext.sparse_attn_fwd(q, k, v, mask, extra0, E ** -0.5, reference_layout=True)),
          "bwd ms %.4f" % ev(lambda: ext.sparse_attn_bwd(q, k, v, yr, dy, mask, extra0, zr, E ** -0.5, reference_layout=True)))
    print(name, "fwd ms %.4f" % ev(lambda: ext.sparse_attn_fwd(q, k, v, mask, extra0, E ** -0.5)),
          "bwd ms %.4f" % ev(lambda: ext.sparse_attn_bwd(q, k, v, y, dy, mask, extra0, z, E ** -0.5)),
          "encode ms %.4f" % ev(lambda: ext.pq_encode_pair(q, k, w)),
          "lookup_mask ms %.4f" % ev(lambda: ext.lookup_mask(ext.pq_encode(q, w), ext.pq_encode(k, w), 8)))
