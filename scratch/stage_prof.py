import sys; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import ext
dev='cuda'
B,S,d,k=128,2048,64,256
g=torch.Generator().manual_seed(7)
q=torch.randn(B,S,d,generator=g).to(dev,torch.bfloat16); kk=torch.randn(B,S,d,generator=g).to(dev,torch.bfloat16)
w=torch.randn(8,16,8,generator=g).to(dev)
qc,kc=ext.pq_encode_pair(q,kk,w)
idx=ext.lookup_forward_cuda(torch.empty([8],device='meta'),qc,kc).flatten(1)
indptr=torch.arange(0,k*S+1,k,dtype=torch.int32,device=dev)
vals=ext.sddmm_forward_cuda(False,True,indptr,idx,q,kk)
p=ext.softmax_forward_cuda(indptr,idx,torch.clamp(vals*d**-0.5,-10,10))
for _ in range(2):
    csc=ext.csr2csc(indptr,idx)
    y=ext.spmm_csc(csc,p,q)
torch.cuda.synchronize()
