import sys, os; sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from spt_proto_b200 import layers
dev='cuda'
torch.manual_seed(1)
d, F, T, bs = 2048, 8192, 8192, 1024
ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=torch.nn.ReLU()).to(dev).bfloat16()
x = torch.randn(16, T // 16, d, device=dev).bfloat16().requires_grad_(); dy = torch.randn_like(x)
def step():
    x.grad = None
    for p in ffn.parameters(): p.grad = None
    ffn(x).backward(dy)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
ev=[e for e in prof.events() if e.device_type==torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
n=len(ev)//4
t0=ev[3*n].time_range.start
for e in ev[3*n:]:
    print(f"{e.time_range.start-t0:8.1f} {e.device_time:7.1f} {e.name[:110]}")
print('span', ev[-1].time_range.end-t0)
