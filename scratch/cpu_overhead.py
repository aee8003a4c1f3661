import sys, time; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import layers, ext
dev=torch.device('cuda:0')
attn=layers.SparseVanillaAttentionV2(d_head=64,d_codeword=8,n_codewords=16,p_dropout=0.0).to(dev)
for n in (1,4):
    q,k,v=(torch.randn(n,2048,32,64,device=dev).bfloat16().requires_grad_() for _ in range(3))
    dy=torch.randn(n,2048,32,64,device=dev).bfloat16()
    def step():
        q.grad=k.grad=v.grad=None
        attn(q,k,v).backward(dy)
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(20): step()
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print(f"n={n}: cpu enqueue {1e3*(t1-t0)/20:.3f} ms/step, total {1e3*(t2-t0)/20:.3f} ms/step")
import cProfile,pstats
q,k,v=(torch.randn(1,2048,32,64,device=dev).bfloat16().requires_grad_() for _ in range(3))
dy=torch.randn(1,2048,32,64,device=dev).bfloat16()
def step():
    q.grad=k.grad=v.grad=None
    attn(q,k,v).backward(dy)
pr=cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
