import sys, os; sys.path.insert(0, '.')
import torch
from spt_proto_b200 import layers
dev='cuda'
torch.manual_seed(1)
d, F, T, bs = 2048, 8192, 8192, 1024
ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=torch.nn.ReLU()).to(dev).bfloat16()
x = torch.randn(16, T // 16, d, device=dev).bfloat16().requires_grad_(); dy = torch.randn_like(x)
ffn(x).backward(dy)
torch.cuda.synchronize()
