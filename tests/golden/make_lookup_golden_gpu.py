"""Run ON THE GPU BOX (gpurun): executes the UNMODIFIED reference lookup kernel (oracle/_ref/ext_ref.so,
built by oracle/build_ref.sh from /root/reference/extension) on seeded codes, prints every slot where
the CPU emulator disagrees, and saves (q, k, reference output) to gpurun_out/lk_*.pt.  Four of those
dumps were packed (uint8 codes, int16 indices) into tests/golden/lookup_ref_kernel_b200.pt — the
golden vectors that pin the lookup oracle to the real kernel on B200.

    gpurun -- python tests/golden/make_lookup_golden_gpu.py
"""
import sys, os, importlib.util, json
sys.path.insert(0, '.')
import torch
from oracle import spt_oracle as O
spec = importlib.util.spec_from_file_location('ext_ref', 'oracle/_ref/ext_ref.so')
ref_ext = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_ext)
g = torch.Generator().manual_seed(11)
res = []
for (B, S, m, c) in [(4, 256, 8, 16), (2, 512, 8, 3), (2, 1024, 16, 16), (2, 512, 10, 4), (2, 256, 8, 1), (8,256,8,16), (8,512,8,16)]:
    q = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
    k = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
    outs = []
    for rep in range(3):
        ref = ref_ext.lookup_forward_cuda(torch.empty([8]), q.cuda(), k.cuda()); torch.cuda.synchronize()
        outs.append(ref.cpu())
    stable = all(torch.equal(outs[0], o) for o in outs)
    emu = O.lookup_forward(q, k, 8)
    diff = (outs[0] != emu).nonzero()
    print((B,S,m,c), 'mismatches', diff.shape[0], 'ref stable across runs', stable)
    for (b, r, p) in diff[:40].tolist():
        cnt = (q[b, r][None, :] == k[b, :r+1]).sum(-1)
        bucket = torch.clamp(cnt // (m // 4), max=3)
        rv, ev = outs[0][b, r, p].item(), emu[b, r, p].item()
        print('  b', b, 'r', r, 'p', p, 'ref', rv, 'emu', ev, 'bucket(ref)', bucket[rv].item() if rv <= r else None, 'bucket(emu)', bucket[ev].item(),
              'lane lens', [[int(((bucket == s) & (torch.arange(r+1) % 4 == t)).sum()) for s in range(4)] for t in range(4)])
    torch.save(dict(q=q, k=k, ref=outs[0]), f'gpurun_out/lk_{B}_{S}_{m}_{c}.pt')
