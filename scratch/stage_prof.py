"""The stage kernels alone at the bench shape (input for ncu): sddmm, spmm, csr2csc, transposed spmm."""
import sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import ext
dev = 'cuda'
B, S, d, k = 128, 2048, 64, 256
g = torch.Generator().manual_seed(7)
q = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16); kk = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16)
w = torch.randn(8, 16, 8, generator=g).to(dev)
qc, kc = ext.pq_encode_pair(q, kk, w)
idx = ext.lookup_forward_cuda(torch.empty([8], device='meta'), qc, kc).flatten(1)
indptr = torch.arange(0, k * S + 1, k, dtype=torch.int32, device=dev)
f, t = False, True
for _ in range(3):
    vals = ext.sddmm_forward_cuda(f, t, indptr, idx, q, kk)
    p = ext.softmax_forward_cuda(indptr, idx, torch.clamp(vals * d ** -0.5, -10, 10))
    y = ext.spmm_forward_cuda(f, f, indptr, idx, p, q)
    csc = ext.csr2csc(indptr, idx)
    yt = ext.spmm_csc(csc, p, q)
torch.cuda.synchronize()
