import sys, os
sys.path.insert(0, '.'); sys.argv=['x','--steps','1','--warmup','2']
import torch
from torch.profiler import profile, ProfilerActivity
import scripts.finetune_step as F
# monkeypatch: wrap main's timed loop with the profiler by running main under profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    F.main()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70), file=sys.stderr)
