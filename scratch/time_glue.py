"""Timings of the glue kernels beside torch's own copies: layout moves of the stage layer, FFN column sums."""
import sys, json
import torch
sys.path.insert(0, ".")
from spt_proto_b200 import ext


def t(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(1_500_000)          # the host queues the calls while the device spins: device time, not launch time
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


dev = "cuda"
out = {}
x = torch.randn(4, 2048, 32, 64, device=dev).bfloat16()
out["swap12_bf16_us"] = t(lambda: ext.swap12(x))
out["swap12_bf16_torch_us"] = t(lambda: x.transpose(1, 2).contiguous())
xf = x.float()
out["swap12_f32_us"] = t(lambda: ext.swap12(xf))
out["swap12_f32_torch_us"] = t(lambda: xf.transpose(1, 2).contiguous())
y = torch.randn(128, 2048, 64, device=dev).bfloat16()
out["transpose_last2_bf16_us"] = t(lambda: ext.transpose_last2(y))
out["transpose_last2_bf16_torch_us"] = t(lambda: y.transpose(1, 2).contiguous())
yf = y.float()
out["transpose_last2_f32_us"] = t(lambda: ext.transpose_last2(yf))
out["transpose_last2_f32_torch_us"] = t(lambda: yf.transpose(1, 2).contiguous())
g = torch.randn(8192, 2048, device=dev).bfloat16()
ptr = torch.tensor([0, 8192], dtype=torch.int32, device=dev)
out["colsum_T8192_C2048_us"] = t(lambda: ext.group_colsum(g, ptr))
gh = torch.randn(32768 + 1024, 1024, device=dev).bfloat16()
ptr8 = torch.arange(0, 9, dtype=torch.int32, device=dev) * 4224
out["colsum_8groups_R33792_C1024_us"] = t(lambda: ext.group_colsum(gh, ptr8))
print(json.dumps(out))
