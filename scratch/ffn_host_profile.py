"""Host-side profile of the eager routed-FFN step (configs[2] shape): where the Python time per step goes."""
import cProfile, pstats, time, sys, io
import torch
sys.path.insert(0, ".")
from spt_proto_b200 import layers, ext

dev = torch.device("cuda:0")
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(4321)
d, F, T = 2048, 8192, 8192
ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=torch.nn.ReLU()).to(dev).bfloat16()
x = torch.randn(16, T // 16, d, device=dev).bfloat16().requires_grad_()
dy = torch.randn(16, T // 16, d, device=dev).bfloat16()


def step():
    x.grad = None
    for p in ffn.parameters():
        p.grad = None
    ffn(x).backward(dy)


for _ in range(5):
    step()
torch.cuda.synchronize()
# host time per step with the GPU never the limiter: time only the launches of 20 steps, then sync
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host launch time per step {1e3 * (t1 - t0) / 20:.3f} ms, incl. drain {1e3 * (t2 - t0) / 20:.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
print(f"device time per eager step {e0.elapsed_time(e1) / 20:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
