"""Transposed product at the bench shape: dense-tile kernel (default) vs the gathered kernel (SPT_SPMM_T_DENSE=0)."""
import os, sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import ext
dev = 'cuda'
import os
B, S, d, k = int(os.environ.get("B", 128)), 2048, 64, 256
g = torch.Generator().manual_seed(7)
q = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16); kk = torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16)
w = torch.randn(8, 16, 8, generator=g).to(dev)
qc, kc = ext.pq_encode_pair(q, kk, w)
idx = ext.lookup_forward_cuda(torch.empty([8], device='meta'), qc, kc).flatten(1)
indptr = torch.arange(0, k * S + 1, k, dtype=torch.int32, device=dev)
vals = ext.sddmm_forward_cuda(False, True, indptr, idx, q, kk)
p = ext.softmax_forward_cuda(indptr, idx, torch.clamp(vals * d ** -0.5, -10, 10))
csc = ext.csr2csc(indptr, idx)
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
print("SPT_SPMM_T_DENSE", os.environ.get("SPT_SPMM_T_DENSE", "1"), "spmm_t bf16 ms %.3f" % ev(lambda: ext.spmm_csc(csc, p, q)),
      "fp32-out ms %.3f" % ev(lambda: ext.spmm_csc(csc, p, q, out_dtype=torch.float32)))
y = ext.spmm_csc(csc, p, q, out_dtype=torch.float32)
print("checksum %.6f" % y.double().abs().sum().item())
