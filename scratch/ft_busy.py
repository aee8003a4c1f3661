import sys, os, time
sys.path.insert(0, '.'); sys.argv=['x','--steps','3','--warmup','3']
import torch
from torch.profiler import profile, ProfilerActivity
import scripts.finetune_step as F
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    F.main()
ev=[e for e in prof.events() if e.device_type==torch.autograd.DeviceType.CUDA]
tot=sum(e.cuda_time if hasattr(e,'cuda_time') else e.device_time for e in ev)
print("kernel events", len(ev), "total cuda time ms", tot/1e3, "per step (6 steps)", tot/1e3/6, file=sys.stderr)
ka=prof.key_averages()
rows=sorted(ka, key=lambda k: -(k.self_device_time_total))[:25]
for k in rows: print(f"{k.self_device_time_total/1e3/6:8.3f} ms/step  x{k.count/6:6.1f}  {k.key[:90]}", file=sys.stderr)
