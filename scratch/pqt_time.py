import sys; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import ext
dev='cuda'
g=torch.Generator().manual_seed(7)
z=torch.randn(2048*32,128,generator=g).to(dev,torch.bfloat16)
w=torch.randn(16,16,8,generator=g).to(dev)
gzq=torch.randn(2048*32,128,generator=g).to(dev)
gl=torch.ones(1,device=dev)
def t(f,n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); b=torch.cuda.Event(True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
print("fwd us", t(lambda: ext.pq_train_fwd(z,w)))
print("bwd us", t(lambda: ext.pq_train_bwd(z,w,gzq,gl)))
