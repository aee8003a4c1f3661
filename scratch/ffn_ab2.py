import sys, os; sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from spt_proto_b200 import layers
dev='cuda'
torch.manual_seed(1)
d, F, T, bs = 2048, 8192, 8192, 1024
ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=torch.nn.ReLU()).to(dev).bfloat16()
x = torch.randn(16, T // 16, d, device=dev).bfloat16().requires_grad_(); dy = torch.randn_like(x)
for _ in range(3): ffn(x).backward(dy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4): ffn(x).backward(dy)
    torch.cuda.synchronize()
ev=[e for e in prof.events() if 'grouped_gemm' in e.name]
ev.sort(key=lambda e: e.time_range.start)
d_=[e.device_time for e in ev]
per=[sum(d_[i::6])/4 for i in range(6)]
print(os.environ.get('SPT_GEMM_DEBUG','0'), 'fc1 fc2 dH dW2 dX dW1:', [round(v,1) for v in per], 'sum', round(sum(per),1))
