import sys, time; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import layers
from spt_proto_b200.host_io import HostPipeline
dev=torch.device('cuda:0')
attn=layers.SparseVanillaAttentionV2(d_head=64,d_codeword=8,n_codewords=16,p_dropout=0.0).to(dev); attn.host_trigger=False
shape=(4,2048,32,64)
host_in=[torch.randn(shape).bfloat16().pin_memory() for _ in range(4)]
outs=[torch.empty(shape,dtype=torch.bfloat16).pin_memory() for _ in range(4)]
for chunk,depth in ((1,3),(1,4),(1,8),(2,2),(2,4)):
    pipe=HostPipeline(attn,dev,chunk=chunk,depth=depth)
    for _ in range(3): pipe.run(host_in,outs)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): pipe.run(host_in,outs)
    b.record(); torch.cuda.synchronize()
    t=a.elapsed_time(b)/10
    print(f"chunk={chunk} depth={depth}: {t:.3f} ms/step  {8192/t*1e3/1e6:.2f} M tok/s")
