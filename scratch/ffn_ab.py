"""A/B of the grouped GEMM kernels on the FFN step (run with SPT_GEMM_CTA_PAIR=0 / 1)."""
import sys, os; sys.path.insert(0, '.')
import torch
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
r = bench.ffn_bench(dev, 1656.0)
print(os.environ.get('SPT_GEMM_CTA_PAIR', '1'), {k: (round(v['ms'], 3), round(v['algorithmic_TFLOPs'], 1)) for k, v in r.items()})
from spt_proto_b200 import layers
torch.manual_seed(1)
d, F, T, bs = 2048, 8192, 8192, 1024
ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=torch.nn.ReLU()).to(dev).bfloat16()
x = torch.randn(16, T // 16, d, device=dev).bfloat16().requires_grad_(); dy = torch.randn_like(x)
for _ in range(3): ffn(x).backward(dy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): ffn(x).backward(dy)
    torch.cuda.synchronize()
for k in sorted(prof.key_averages(), key=lambda k: -k.self_device_time_total)[:6]:
    print(f"{k.self_device_time_total/5:9.1f} us/step x{k.count/5:4.1f} {k.key[:80]}")
