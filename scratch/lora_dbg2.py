import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from oracle import spt_oracle as O
from spt_proto_b200 import layers
import test_ffn_gpu as T
d, Fdim, bs, r, Tn = 256, 1024, 256, 16, 768
ffn, x = T._lora_setup(layers.LoRARoutedLLaMaFFN, torch.nn.SiLU(), d, Fdim, bs, r, Tn, 12)
y = ffn(x); dy = T._bf(torch.randn_like(y)); y.backward(dy)
sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
names = [n for n, p in ffn.named_parameters() if p.requires_grad]
p = {n: sd[n].clone().requires_grad_() for n in names}
xc = x.detach().cpu().requires_grad_()
y_ref = O.lora_routed_llama_ffn(xc, p["router.0.weight"], p["router.0.bias"], sd["gate.weight"], sd["side.weight"], sd["down.weight"],
        p["gate.lora.left.weight"], p["gate.lora.right.weight"], p["side.lora.left.weight"], p["side.lora.right.weight"], p["down.lora.left.weight"], p["down.lora.right.weight"], bs, (Fdim // bs) // 2)
y_ref.backward(dy.cpu())
rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
print('y', rel(y, y_ref.detach()), 'dx', rel(x.grad, xc.grad), y_ref.abs().max().item(), xc.grad.abs().max().item())
got = dict(ffn.named_parameters())
for n in names: print(n, rel(got[n].grad, p[n].grad), p[n].grad.norm().item())
