import torch, time
dev=torch.device('cuda:0')
n=32<<20  # 32 MiB
h_in=[torch.empty(n,dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out=[torch.empty(n,dtype=torch.uint8).pin_memory() for _ in range(4)]
d=[torch.empty(n,dtype=torch.uint8,device=dev) for _ in range(4)]
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,it=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/it
def h2d():
    for a,b in zip(d,h_in): a.copy_(b,non_blocking=True)
def d2h():
    for a,b in zip(h_out,d): a.copy_(b,non_blocking=True)
def both():
    with torch.cuda.stream(s1): h2d()
    with torch.cuda.stream(s2): d2h()
gb=4*n/1e9
print('h2d GB/s',gb/t(h2d)); print('d2h GB/s',gb/t(d2h)); tb=t(both); print('both: each dir GB/s',gb/tb)
import os; print('cpus',os.cpu_count())
