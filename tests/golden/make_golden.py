"""Generates tests/golden/*.pt by running the UNMODIFIED reference Python package from
/root/reference (ytgui/SPT-proto) on CPU.  Run in the authoring container only; the fixtures are
committed because /root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py

Import shim: `naive_gpt.ext` is CUDA-only, `naive_gpt.loaders` needs
lightning/torchtext.  For the pure-torch reference paths (PQV1, RoutedFFN, LoRARoutedFFN, ...) no
stub is ever called.  For the SparseVanillaAttentionV2 *glue* fixture the 7 ext entry points are
stubbed with oracle/spt_oracle.py so that the reference's own layer code (transposes, scaling, clamp,
CSR construction) produces the output; that fixture therefore pins the glue, not the stage math —
the stage math is pinned by the dense formulas of the reference's tests (fixture `stage_formulas`).
"""
import contextlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import spt_oracle as O  # noqa: E402

ext = types.ModuleType("naive_gpt.ext")
ext.cdist_forward_cuda = lambda q, t: list(O.cdist_forward(q, t))
ext.cdist_backward_cuda = lambda q, t, g: list(O.cdist_backward(q, t, g))
ext.lookup_forward_cuda = lambda cfg, q, k: O.lookup_forward(q, k, cfg.size(0))
ext.sddmm_forward_cuda = lambda tl, tr, ip, ix, q, k: O.sddmm_forward(ip, ix, q, k)
ext.spmm_forward_cuda = lambda tl, tr, ip, ix, v, x: O.spmm_forward(bool(tl.item()), ip, ix, v, x)
ext.softmax_forward_cuda = lambda ip, ix, v: O.softmax_forward(ip, ix, v)
ext.softmax_backward_cuda = lambda ip, ix, y, g: O.softmax_backward(ip, ix, y, g)
sys.modules["naive_gpt.ext"] = ext
sys.modules["naive_gpt.loaders"] = types.ModuleType("naive_gpt.loaders")


class _S:
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass


torch.cuda.current_stream = lambda *a, **k: _S()
torch.cuda.Stream = _S
torch.cuda.stream = lambda s: contextlib.nullcontext()

import naive_gpt  # noqa: E402
from naive_gpt import layers  # noqa: E402
from torch import nn  # noqa: E402


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def grads_of(module, x, y, names):
    y.sum().backward()
    out = {"x": x.grad.clone()}
    for n in names:
        p = dict(module.named_parameters())[n]
        out[n] = None if p.grad is None else p.grad.clone()
    return out


def golden_pq():
    torch.manual_seed(1234)
    cases = []
    for (B, S, m, c, dc, bf16) in [(3, 64, 8, 16, 8, False), (2, 128, 8, 16, 8, True), (2, 48, 16, 16, 8, False),
                                   (1, 96, 4, 32, 4, False)]:
        pq = layers.PQV1(d_codeword=dc, n_codewords=c, n_subspaces=m)
        z = torch.randn(B, S, m * dc)
        if bf16:  # bf16-rounded inputs create exact distance ties
            z = z.bfloat16().float()
            pq.weight.data = pq.weight.data.bfloat16().float()
        with torch.no_grad():
            codes = pq("encode", z=z)                       # torch.cdist(p=1) + argmin, quantizer.py:53-62
            z_q, loss = pq("train", z=z)
        cases.append(dict(z=z, weight=pq.weight.detach().clone(), codes=codes.squeeze(-1) if codes.dim() > 3 else codes,
                          z_q=z_q, loss=loss))
    save("pq_v1", cases)


def golden_routed_ffn():
    torch.manual_seed(4321)
    out = {}
    d, F, bs = 32, 128, 16
    x = torch.randn(3, 20, d, requires_grad=True)
    ffn = layers.RoutedFFN(d_model=d, d_feedforward=F, block_size=bs, activation=nn.ReLU())
    y = ffn(x)
    g = grads_of(ffn, x, y, ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"])
    out["routed_ffn"] = dict(x=x.detach().clone(), state=ffn.state_dict(), y=y.detach().clone(), grads=g,
                             cfg=dict(d_model=d, d_feedforward=F, block_size=bs, k_active=(F // bs) // 2))

    x = torch.randn(2, 24, d, requires_grad=True)
    ffn = layers.RoutedLLaMaFFN(d_model=d, d_feedforward=F, block_size=bs, activation=nn.SiLU())
    y = ffn(x)
    g = grads_of(ffn, x, y, ["gate.weight", "side.weight", "down.weight"])
    out["routed_llama_ffn"] = dict(x=x.detach().clone(), state=ffn.state_dict(), y=y.detach().clone(), grads=g,
                                   cfg=dict(d_model=d, d_feedforward=F, block_size=bs, k_active=(F // bs) // 4))

    x = torch.randn(2, 24, d, requires_grad=True)
    ffn = layers.LoRARoutedFFN(d_lora=4, block_size=bs * 2, d_model=d, d_feedforward=F, activation=nn.ReLU())
    for name, p in ffn.named_parameters():
        if "lora.right" in name:
            nn.init.normal_(p, std=0.1)
    y = ffn(x)
    names = [n for n, p in ffn.named_parameters() if p.requires_grad]
    g = grads_of(ffn, x, y, names)
    out["lora_routed_ffn"] = dict(x=x.detach().clone(), state=ffn.state_dict(), y=y.detach().clone(), grads=g,
                                  cfg=dict(d_lora=4, d_model=d, d_feedforward=F, block_size=bs * 2,
                                           k_active=(F // (bs * 2)) // 2))

    x = torch.randn(2, 24, d, requires_grad=True)
    ffn = layers.LoRARoutedLLaMaFFN(d_lora=4, block_size=bs * 2, d_model=d, d_feedforward=F, activation=nn.SiLU())
    for name, p in ffn.named_parameters():
        if "lora.right" in name:
            nn.init.normal_(p, std=0.1)
    y = ffn(x)
    names = [n for n, p in ffn.named_parameters() if p.requires_grad]
    g = grads_of(ffn, x, y, names)
    out["lora_routed_llama_ffn"] = dict(x=x.detach().clone(), state=ffn.state_dict(), y=y.detach().clone(), grads=g,
                                        cfg=dict(d_lora=4, d_model=d, d_feedforward=F, block_size=bs * 2,
                                                 k_active=(F // (bs * 2)) // 2))
    save("routed_ffn", out)


def golden_sparse_mha_glue():
    torch.manual_seed(2468)
    N, S, H, E = 2, 64, 3, 32
    attn = layers.SparseVanillaAttentionV2(d_head=E, d_codeword=8, n_codewords=16, p_dropout=0.0)
    q = torch.randn(N, S, H, E, requires_grad=True)
    k = torch.randn(N, S, H, E, requires_grad=True)
    v = torch.randn(N, S, H, E, requires_grad=True)
    y = attn(q, k, v)
    y.sum().backward()
    save("sparse_mha_glue", dict(q=q.detach().clone(), k=k.detach().clone(), v=v.detach().clone(),
                                 weight=attn.quantizer.weight.detach().clone(), y=y.detach().clone(),
                                 dq=q.grad.clone(), dk=k.grad.clone(), dv=v.grad.clone()))
    # dense attention on all-ones inputs: the reference's own layer test (test_sparse_mha.py:7-43)
    dense = layers.VanillaAttention(d_head=E, p_dropout=0.0)
    ones = torch.ones(N, S, H, E)
    mask = torch.full([S, S], float("-inf")).triu(1)
    save("dense_ones", dict(y=dense(ones, ones, ones, attn_mask=mask)))


def golden_stage_formulas():
    """The dense torch formulas the reference's kernel tests use as oracles (test_sddmm.py:58-62,
    test_softmax.py:70, test_spmm.py:56), evaluated on a seeded uniform-k pattern."""
    torch.manual_seed(1357)
    B, S, d = 2, 64, 32
    k_per = S // 8
    prob = torch.rand(B, S, S)
    tril = torch.tril(torch.ones(S, S, dtype=torch.bool))
    prob = torch.where(tril, prob, torch.zeros(()))
    topk = torch.topk(prob, k=k_per, dim=-1, sorted=False)
    mask = torch.scatter(torch.zeros_like(prob), -1, topk.indices, torch.ones_like(topk.values))
    q = torch.randn(B, S, d, requires_grad=True)
    k = torch.randn(B, S, d, requires_grad=True)
    v = torch.randn(B, S, d, requires_grad=True)
    scores = mask * torch.matmul(q, k.transpose(-1, -2))                       # sddmm oracle
    neg = torch.where((mask > 0) & tril, d ** -0.5 * scores, torch.full((), float("-inf")))
    probs = torch.softmax(neg, dim=-1)                                         # softmax oracle
    probs = torch.where(torch.isnan(probs), torch.zeros(()), probs)
    y = torch.matmul(probs, v)                                                 # spmm oracle
    w = torch.randn_like(y)
    (y * w).sum().backward()
    save("stage_formulas", dict(indices=topk.indices.flatten(1).to(torch.int32), k_per=k_per,
                                q=q.detach().clone(), k=k.detach().clone(), v=v.detach().clone(), w=w,
                                scores=scores.detach().clone(), probs=probs.detach().clone(), y=y.detach().clone(),
                                dq=q.grad.clone(), dk=k.grad.clone(), dv=v.grad.clone()))


if __name__ == "__main__":
    golden_pq()
    golden_routed_ffn()
    golden_sparse_mha_glue()
    golden_stage_formulas()
