"""Error statistics of the fused attention against the oracle (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import spt_oracle as O
from spt_proto_b200 import ext, kernels
DEV = "cuda"
for (B, S, scale_mul, d) in [(3, 256, 1.0, 64), (1, 1024, 1.0, 64), (1, 2048, 1.0, 64), (2, 256, 6.0, 64), (1, 1024, 1.0, 128)]:
    g = torch.Generator().manual_seed(S + B)
    q = (torch.randn(B, S, d, generator=g) * scale_mul ** 0.5).bfloat16()
    k = (torch.randn(B, S, d, generator=g) * scale_mul ** 0.5).bfloat16()
    v = torch.randn(B, S, d, generator=g).bfloat16()
    dy = torch.randn(B, S, d, generator=g).bfloat16()
    w = torch.randn(d // 8, 16, 8, generator=g)
    indptr, indices = O.sparse_attention_indices(q.float(), k.float(), w, 8)
    qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
    y_ref, _ = O.sparse_attention_values(indptr, indices, qf, kf, vf, d ** -0.5)
    y_ref.backward(dy.float())
    qd, kd, vd = (t.to(DEV).requires_grad_() for t in (q, k, v))
    q_c, k_c = ext.pq_encode(qd.detach(), w.to(DEV)), ext.pq_encode(kd.detach(), w.to(DEV))
    mask, extra0, idx = ext.lookup_mask(q_c, k_c, 8, want_indices=True)
    y = kernels.sparse_attention(qd, kd, vd, mask, extra0, d ** -0.5)
    y.backward(dy.to(DEV))
    for name, got, want in (("y", y, y_ref.detach()), ("dq", qd.grad, qf.grad), ("dk", kd.grad, kf.grad), ("dv", vd.grad, vf.grad)):
        diff = (got.float().cpu() - want)
        viol = (diff.abs() > 4e-2 + 3e-2 * want.abs())
        rows = viol.any(-1).nonzero()
        print(f"S={S} d={d} mul={scale_mul} {name}: rel_fro={diff.norm()/want.norm():.2e} max_abs={diff.abs().max():.3e} "
              f"max|want|={want.abs().max():.2f} viol={int(viol.sum())} rows={rows[:6].flatten().tolist()}")
