import sys, os; sys.path.insert(0,'.')
import torch
from spt_proto_b200 import layers
gd = torch.load("tests/golden/routed_ffn.pt", weights_only=False)
for key, cls, act in (("routed_ffn", layers.RoutedFFN, torch.nn.ReLU()), ("routed_llama_ffn", layers.RoutedLLaMaFFN, torch.nn.SiLU())):
    case = gd[key]; cfg = case["cfg"]
    ffn = cls(d_model=cfg["d_model"], d_feedforward=cfg["d_feedforward"], block_size=cfg["block_size"], activation=act).cuda()
    ffn.load_state_dict(case["state"])
    x = case["x"].cuda().requires_grad_()
    y = ffn(x); y.sum().backward()
    rel = lambda a,b: ((a.float().cpu()-b).norm()/b.norm()).item()
    print(key, 'y', rel(y.detach(), case['y']), 'dx', rel(x.grad, case['grads']['x']))
    for n,p in ffn.named_parameters():
        if n in case['grads'] and case['grads'][n] is not None: print('   ', n, rel(p.grad, case['grads'][n]))
    # same with bf16-rounded reference computed by oracle
    from oracle import spt_oracle as O
    sd = {k: v.bfloat16().float() for k, v in case["state"].items()}
    xc = case["x"].bfloat16().float().requires_grad_()
    if key == "routed_ffn":
        yr = O.routed_ffn(xc, case["state"]["router.0.weight"], case["state"]["router.0.bias"], sd["fc1.weight"], case["state"]["fc1.bias"], sd["fc2.weight"], case["state"]["fc2.bias"], cfg["block_size"], cfg["k_active"])
        yr.sum().backward()
        print('  vs bf16-rounded oracle: y', rel(y.detach(), yr.detach()), 'dx', rel(x.grad, xc.grad))
