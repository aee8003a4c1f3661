import sys; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import ext
dev='cuda'
B,S,d=128,2048,64
g=torch.Generator().manual_seed(7)
q,k,v,dy=(torch.randn(B,S,d,generator=g).to(dev,torch.bfloat16) for _ in range(4))
w=torch.randn(8,16,8,generator=g).to(dev)
qc,kc=ext.pq_encode_pair(q,k,w)
mask,extra0,_=ext.lookup_mask(qc,kc,8)
y,z=ext.sparse_attn_fwd(q,k,v,mask,extra0,d**-0.5)
for _ in range(3): ext.sparse_attn_bwd(q,k,v,y,dy,mask,extra0,z,d**-0.5)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ext.sparse_attn_bwd(q,k,v,y,dy,mask,extra0,z,d**-0.5)
b.record(); torch.cuda.synchronize()
print("bwd ms", a.elapsed_time(b)/10)
