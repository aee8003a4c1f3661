import sys; sys.path.insert(0,'.')
import torch
from oracle import spt_oracle as O
from spt_proto_b200 import ext, kernels
DEV='cuda'
for (B,S,scale_mul,d) in [(2,256,6.0,64),(2,384,6.0,64),(2,384,6.0,128),(2,384,1.0,128)]:
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
    q_c, k_c = ext.pq_encode_pair(qd.detach(), kd.detach(), w.to(DEV))
    mask, extra0, idx = ext.lookup_mask(q_c, k_c, 8, want_indices=True)
    y = kernels.sparse_attention(qd, kd, vd, mask, extra0, d ** -0.5)
    y.backward(dy.to(DEV))
    for name, got, want in (("y", y, y_ref.detach()), ("dq", qd.grad, qf.grad), ("dk", kd.grad, kf.grad), ("dv", vd.grad, vf.grad)):
        gotf = got.float().cpu()
        err = (gotf - want)
        viol = (err.abs() > 4e-2 + 3e-2 * want.abs()).sum().item()
        print((B,S,scale_mul,d), name, "relF %.2e" % (err.norm()/want.norm()).item(), "maxabs %.3f" % err.abs().max().item(), "max|want| %.1f" % want.abs().max().item(), "viol", viol)
