"""In-kernel phase timers of the attention kernels at the bench shape (library built with -DSPT_ATTN_PROF)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spt_proto_b200 import ext
from spt_proto_b200._lib import lib
dev = "cuda"
B, S, d = int(os.environ.get("B", 128)), int(os.environ.get("S", 2048)), int(os.environ.get("D", 64))
g = torch.Generator().manual_seed(1)
q, k, v, dy = (torch.randn(B, S, d, generator=g).to(dev, torch.bfloat16) for _ in range(4))
w = torch.randn(d // 8, 16, 8, generator=g).to(dev)
mask, extra0, _ = ext.lookup_mask(ext.pq_encode(q, w), ext.pq_encode(k, w), 8)
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
y, z = ext.sparse_attn_fwd(q, k, v, mask, extra0, d ** -0.5)
print("fwd ms", ev(lambda: ext.sparse_attn_fwd(q, k, v, mask, extra0, d ** -0.5)))
print("bwd ms", ev(lambda: ext.sparse_attn_bwd(q, k, v, y, dy, mask, extra0, z, d ** -0.5)))
buf = (ctypes.c_ulonglong * 48)()
lib.spt_debug_attn_prof(buf, 1)
torch.cuda.synchronize()
ext.sparse_attn_fwd(q, k, v, mask, extra0, d ** -0.5)
ext.sparse_attn_bwd(q, k, v, y, dy, mask, extra0, z, d ** -0.5)
torch.cuda.synchronize()
rc = lib.spt_debug_attn_prof(buf, 0)
print("prof build:", rc)
names = ["fwd", "bwd_q", "bwd_kv"]
mlab = ["wait_scores", "tmem_ld", "math", "st+arrive", "iters", "loop_total"]
ilab = ["wait_operands", "wait_math", "issue", "loop_total", "iters", "issue_acc(dQ)"]
for kidx, name in enumerate(names):
    r = [buf[kidx * 16 + i] for i in range(16)]
    it_m, it_i = max(r[4], 1), max(r[12], 1)
    print(name, "math thread, clk per iteration:", {mlab[i]: round(r[i] / it_m, 1) for i in (0, 1, 2, 3, 5)}, "iters", r[4])
    print(name, "issuer, clk per iteration:", {ilab[i]: round(r[8 + i] / it_i, 1) for i in (0, 1, 2, 3, 5)}, "iters", r[12])

r = [buf[i] for i in range(16)]
n_cta = max(r[14], 1)
print("fwd per CTA (clk): entry->loop", r[6] / n_cta, " loop", r[5] / n_cta, " loop_end->exit(epilogue)", r[7] / n_cta, " whole CTA", r[15] / n_cta, " CTAs", r[14])
print("fwd: clk per ns over CTA lifetimes =", r[15] / max(r[13], 1), "(SM clock in GHz while the kernel runs)")
