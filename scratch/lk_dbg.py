import sys; sys.path.insert(0, '.')
import torch
from oracle import spt_oracle as O
from spt_proto_b200 import ext
sys.path.insert(0, 'tests')
from test_fused_gpu import _mask_from_indices
B, S, m, c, coeff = 1, 8192, 8, 16, 512
g = torch.Generator().manual_seed(S + c)
q = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
k = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
want_idx = O.lookup_forward(q, k, coeff)
want_words, want_extra = _mask_from_indices(want_idx, S)
mask_i, extra_i, idx = ext.lookup_mask(q.cuda(), k.cuda(), coeff, want_indices=True)
mask_m, extra_m, _ = ext.lookup_mask(q.cuda(), k.cuda(), coeff)
print("idx equal", torch.equal(idx.cpu(), want_idx))
for name, mk, ex in (("index-kernel mask", mask_i, extra_i), ("mask-only kernel", mask_m, extra_m)):
    d = (mk.cpu() != want_words)
    rows = d.any(-1).nonzero()
    print(name, "mismatching words", int(d.sum()), "rows", rows[:10].flatten().tolist(), "extra0 equal", torch.equal(ex.cpu(), want_extra))
    if len(rows):
        b, r = rows[0].tolist()
        w = d[b, r].nonzero().flatten().tolist()
        print("  row", r, "words", w[:8], "got", [hex(int(mk[b, r, x]) & 0xffffffff) for x in w[:4]], "want", [hex(int(want_words[b, r, x]) & 0xffffffff) for x in w[:4]])
        print("  want idx row", want_idx[b, r].tolist())
