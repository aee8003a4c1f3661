"""Tile index build + transposed product on it at the bench shape, against csr2csc + the CSC kernel."""
import os, sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import ext
dev = 'cuda'
B, S, d, k = int(os.environ.get("B", 128)), 2048, 64, 256
g = torch.Generator().manual_seed(7)
q = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16); kk = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16)
w = torch.randn(8, 16, 8, generator=g).to(dev)
qc, kc = ext.pq_encode_pair(q, kk, w)
idx = ext.lookup_forward_cuda(torch.empty([8], device='meta'), qc, kc).flatten(1)
indptr = torch.arange(0, k * S + 1, k, dtype=torch.int32, device=dev)
vals = ext.sddmm_forward_cuda(False, True, indptr, idx, q, kk)
p = ext.softmax_forward_cuda(indptr, idx, torch.clamp(vals * d ** -0.5, -10, 10))
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
csc = ext.csr2csc(indptr, idx)
tiles = ext.csr_tiles(indptr, idx)
print("B", B, "csr2csc ms %.3f" % ev(lambda: ext.csr2csc(indptr, idx)), "csr_tiles ms %.3f" % ev(lambda: ext.csr_tiles(indptr, idx)),
      "spmm_csc ms %.3f" % ev(lambda: ext.spmm_csc(csc, p, q)), "spmm_tiles ms %.3f" % ev(lambda: ext.spmm_tiles(tiles, p, q)),
      "| direct: gathered ms %.3f" % ev(lambda: ext.spmm_forward_cuda(False, False, indptr, idx, p, q)),
      "tiles ms %.3f" % ev(lambda: ext.spmm_tiles(tiles, p, q, trans=False)),
      "| sddmm: gathered ms %.3f" % ev(lambda: ext.sddmm_scaled(indptr, idx, q, kk, 0.125, 10.0)),
      "tiles ms %.3f" % ev(lambda: ext.sddmm_tiles(tiles, q, kk, 0.125, 10.0)))
a, b = ext.spmm_csc(csc, p, q, out_dtype=torch.float32), ext.spmm_tiles(tiles, p, q, out_dtype=torch.float32)
print("max |csc - tiles| %.3e  rel %.3e" % ((a - b).abs().max().item(), ((a - b).norm() / a.norm()).item()))
