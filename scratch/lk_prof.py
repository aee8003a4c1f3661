"""The lookup kernels alone (input for ncu) + CUDA-event time of the mask-only lookup (SPT_LOOKUP_MASK_V1=1: first version)."""
import sys; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import ext
B, S, m = 128, 2048, 8
g = torch.Generator().manual_seed(0)
q = torch.randn(B, S, 64, generator=g).bfloat16().cuda(); k = torch.randn(B, S, 64, generator=g).bfloat16().cuda()
w = torch.randn(8, 16, 8, generator=g).cuda()
qc, kc = ext.pq_encode(q, w), ext.pq_encode(k, w)
for _ in range(3):
    ext.lookup_mask(qc, kc, 8)
    ext.lookup_forward_cuda(torch.empty([8]), qc, kc)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ext.lookup_mask(qc, kc, 8)
b.record(); b.synchronize()
print("lookup_mask ms %.4f" % (a.elapsed_time(b) / 10))
