import sys, time; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import layers
dev=torch.device('cuda:0')
torch.manual_seed(0)
d,F,T=2048,8192,8192
for bs in (1024,2048):
    ffn=layers.RoutedFFN(d_model=d,d_feedforward=F,block_size=bs,activation=torch.nn.ReLU()).to(dev).bfloat16()
    x=torch.randn(16,T//16,d,device=dev).bfloat16().requires_grad_()
    dy=torch.randn(16,T//16,d,device=dev).bfloat16()
    def step():
        x.grad=None
        for p in ffn.parameters(): p.grad=None
        ffn(x).backward(dy)
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(20): step()
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print(f"bs={bs}: cpu enqueue {1e3*(t1-t0)/20:.3f} ms/step, total {1e3*(t2-t0)/20:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
