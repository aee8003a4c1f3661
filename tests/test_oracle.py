"""CPU tests: pin the oracle (oracle/spt_oracle.py, oracle/spt_oracle_c.c) against the golden
vectors generated from the reference's own Python (tests/golden/make_golden.py) and against its
second, independent restatements.  No GPU needed."""
import os

import pytest
import torch

from oracle import spt_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


# ---- cdist / PQ codes: reference torch path PQV1 (quantizer.py:53-62) --------------------------------
def test_pq_codes_match_reference_pqv1():
    for case in gold("pq_v1"):
        codes = O.pq_encode(case["z"], case["weight"])
        ref = case["codes"].reshape(codes.shape)
        assert torch.equal(codes.long(), ref.long())


def test_cdist_c_loop_equals_vectorised_oracle():
    g = torch.Generator().manual_seed(0)
    for (m, n, c, dc) in [(8, 512, 16, 8), (3, 100, 40, 4), (2, 64, 16, 24)]:
        q = torch.randn(m, n, dc, generator=g).bfloat16().float()     # bf16 rounding => exact ties
        t = torch.randn(m, c, dc, generator=g).bfloat16().float()
        d1, i1 = O.cdist_forward(q, t)
        d2, i2 = O.cdist_forward_c(q.numpy(), t.numpy())
        assert torch.equal(d1, torch.from_numpy(d2)) and torch.equal(i1, torch.from_numpy(i2))


def test_pq_train_loss_matches_reference():
    for case in gold("pq_v1"):
        loss = O.pq_train_loss(case["z"], case["weight"])
        assert torch.allclose(loss, case["loss"], rtol=1e-5, atol=1e-6)


def test_cdist_backward_matches_autograd_of_l1():
    g = torch.Generator().manual_seed(1)
    q = torch.randn(4, 96, 8, generator=g, requires_grad=True)
    t = torch.randn(4, 16, 8, generator=g, requires_grad=True)
    go = torch.randn(4, 96, 16, generator=g)
    (torch.cdist(q, t, p=1.0) * go).sum().backward()
    gq, gt = O.cdist_backward(q, t, go)
    assert torch.allclose(gq, q.grad, atol=1e-5) and torch.allclose(gt, t.grad, atol=1e-4)


# ---- lookup: literal C emulation == abstract spec -------------------------------------------------------
@pytest.mark.parametrize("B,S,m,c,coeff", [(2, 64, 8, 16, 8), (2, 128, 8, 2, 8), (1, 128, 16, 4, 8),
                                            (1, 256, 8, 1, 8), (2, 64, 10, 3, 4), (1, 128, 4, 2, 4),
                                            (1, 96, 12, 5, 2), (1, 256, 8, 16, 8)])
def test_lookup_literal_equals_spec(B, S, m, c, coeff):
    g = torch.Generator().manual_seed(S * m + c)
    q = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
    k = torch.randint(0, c, (B, S, m), generator=g, dtype=torch.int32)
    assert torch.equal(O.lookup_forward(q, k, coeff), O.lookup_spec(q, k, coeff))


def test_lookup_oracle_matches_reference_kernel_on_b200():
    """Golden vectors = outputs of the UNMODIFIED reference CUDA kernel run on a B200
    (tests/golden/make_lookup_golden_gpu.py).  Pins both restatements bit-exactly, including the
    same-instruction store conflict (lowest lane wins on this hardware)."""
    for case in gold("lookup_ref_kernel_b200"):
        q, k, ref = case["q"].int(), case["k"].int(), case["ref"].int()
        assert torch.equal(O.lookup_forward(q, k, case["sparse_coeff"]), ref)
        assert torch.equal(O.lookup_spec(q[:1], k[:1], case["sparse_coeff"]), ref[:1])


def test_lookup_structure_properties():
    g = torch.Generator().manual_seed(2)
    B, S, m = 2, 256, 8
    q = torch.randint(0, 16, (B, S, m), generator=g, dtype=torch.int32)
    k = torch.randint(0, 16, (B, S, m), generator=g, dtype=torch.int32)
    out = O.lookup_forward(q, k, 8)
    nnz = S // 8
    rows = torch.arange(S).view(1, S, 1)
    assert (out >= 0).all() and (out <= rows).all()                    # causal
    # short rows whose lanes stay under the per-(lane,bucket) cap nnz/4-1: all keys, zero padded
    # (from r = nnz-2 on, lane 2/3 can overflow into lane 1/0's last slot — that IS the reference)
    for r in range(nnz - 4):
        assert sorted(out[0, r, : r + 1].tolist()) == list(range(r + 1))
        assert (out[0, r, r + 1:] == 0).all()
    pos = torch.arange(nnz).view(1, 1, nnz)
    full = out[:, nnz:, : nnz - 4]                                      # last 4 slots can hold a partner's key
    assert ((full % 4) == (pos[..., : nnz - 4] % 4)).all()              # position p holds a key of lane p % 4


def test_lookup_recall_like_reference_test():
    g = torch.Generator().manual_seed(3)
    B, S, m = 1, 512, 8
    q = torch.randint(0, 8, (B, S, m), generator=g, dtype=torch.int32)
    k = torch.randint(0, 8, (B, S, m), generator=g, dtype=torch.int32)
    out = O.lookup_forward(q, k, 8)
    score = O.exact_topk_match_count(q, k, S // 8)
    rec = []
    for r in range(0, S, 5):
        kk = min(r + 1, S // 8)
        sc = score[0, r, : r + 1]                                       # tie-aware recall, see GPU test
        thr = torch.topk(sc, k=kk).values.min()
        rec.append((sc[out[0, r, :kk].long()] >= thr).float().mean().item())
    assert sum(rec) / len(rec) > 0.8                                    # test_lookup.py:75


# ---- stage formulas of the reference's kernel tests --------------------------------------------------
def test_stage_oracles_match_reference_dense_formulas():
    gd = gold("stage_formulas")
    S = gd["q"].shape[1]
    indptr = O.fixed_indptr(S, gd["k_per"])
    idx = gd["indices"]
    rows = torch.arange(S).repeat_interleave(gd["k_per"])
    B = idx.shape[0]
    dense_scores = torch.zeros(B, S, S)
    vals = O.sddmm_forward(indptr, idx, gd["q"], gd["k"])
    dense_scores[torch.arange(B).view(-1, 1), rows.view(1, -1), idx.long()] = vals
    assert torch.allclose(dense_scores, gd["scores"], atol=1e-4)
    d = gd["q"].shape[-1]
    p = O.softmax_forward(indptr, idx, torch.clamp(d ** -0.5 * vals, -10, 10))
    dense_p = torch.zeros(B, S, S)
    dense_p[torch.arange(B).view(-1, 1), rows.view(1, -1), idx.long()] = p
    assert torch.allclose(dense_p, gd["probs"], atol=1e-5)
    y = O.spmm_forward(False, indptr, idx, p, gd["v"])
    assert torch.allclose(y, gd["y"], atol=1e-4)
    # backward through the differentiable gathered form
    q, k, v = (gd[n].clone().requires_grad_() for n in ("q", "k", "v"))
    y2, _ = O.sparse_attention_values(indptr, idx, q, k, v, d ** -0.5)
    (y2 * gd["w"]).sum().backward()
    assert torch.allclose(y2, gd["y"], atol=1e-4)
    for n, t in (("dq", q), ("dk", k), ("dv", v)):
        assert torch.allclose(t.grad, gd[n], atol=1e-4), n
    # stage-wise backward formulas (kernels/spmm.py:23-49, softmax.py:21-30, sddmm.py:25-51)
    dP = O.sddmm_forward(indptr, idx, gd["w"], gd["v"])
    dV = O.spmm_forward(True, indptr, idx, p, gd["w"])
    dS = O.softmax_backward(indptr, idx, p, dP)
    s_raw = d ** -0.5 * vals
    dS = dS * ((s_raw >= -10) & (s_raw <= 10)).float() * d ** -0.5
    dQ = O.spmm_forward(False, indptr, idx, dS, gd["k"])
    dK = O.spmm_forward(True, indptr, idx, dS, gd["q"])
    assert torch.allclose(dV, gd["dv"], atol=1e-4)
    assert torch.allclose(dQ, gd["dq"], atol=1e-4)
    assert torch.allclose(dK, gd["dk"], atol=1e-4)


def test_softmax_backward_reference_clamp_flag():
    indptr = torch.tensor([0, 4], dtype=torch.int32)
    idx = torch.zeros(1, 4, dtype=torch.int32)
    y = torch.full((1, 4), 0.25)
    dy = torch.tensor([[-1.0, -2.0, 0.5, 0.1]])
    true = O.softmax_backward(indptr, idx, y, dy)
    buggy = O.softmax_backward(indptr, idx, y, dy, reference_clamp=True)
    assert torch.allclose(true.sum(), torch.zeros(()), atol=1e-7)       # true softmax grads sum to zero
    assert not torch.allclose(true, buggy)


def test_sparse_mha_glue_matches_reference_layer():
    gd = gold("sparse_mha_glue")
    q, k, v = (gd[n].clone().requires_grad_() for n in ("q", "k", "v"))
    y = O.sparse_mha_layer(q, k, v, gd["weight"], sparse_coeff=8, reference_output_layout=True)
    y.sum().backward()
    assert torch.allclose(y, gd["y"], atol=1e-5)
    assert torch.allclose(q.grad, gd["dq"], atol=1e-5)
    assert torch.allclose(k.grad, gd["dk"], atol=1e-5)
    assert torch.allclose(v.grad, gd["dv"], atol=1e-5)


def test_sparse_mha_default_layout_is_masked_dense_attention():
    """Default layout == the reference's dense VanillaAttention restricted to the selected pattern."""
    gd = gold("sparse_mha_glue")
    q, k, v, w = gd["q"], gd["k"], gd["v"], gd["weight"]
    N, S, H, E = q.shape
    y = O.sparse_mha_layer(q, k, v, w, sparse_coeff=8)
    qh, kh, vh = (t.transpose(1, 2).reshape(N * H, S, E) for t in (q, k, v))
    indptr, idx = O.sparse_attention_indices(qh, kh, w, 8)
    rows = torch.arange(S).repeat_interleave(S // 8)
    mult = torch.zeros(N * H, S, S)                                     # multiplicity (zero padding repeats key 0)
    mult.index_put_((torch.arange(N * H).view(-1, 1).expand_as(idx), rows.view(1, -1).expand_as(idx), idx.long()),
                    torch.ones(idx.shape), accumulate=True)
    scores = torch.clamp(E ** -0.5 * qh @ kh.transpose(1, 2), -10, 10)
    e = torch.exp(scores) * mult * torch.tril(torch.ones(S, S))
    dense = (e / e.sum(-1, keepdim=True)) @ vh
    assert torch.allclose(y, dense.view(N, H, S, E).transpose(1, 2), atol=1e-5)
    assert torch.allclose(O.sparse_mha_layer(q, k, v, w, 8, reference_output_layout=True),
                          dense.transpose(1, 2).contiguous().view(N, S, H, E), atol=1e-5)


def test_csr2csc_oracle_is_stable_transpose():
    g = torch.Generator().manual_seed(4)
    B, S, k = 2, 64, 8
    idx = torch.randint(0, S, (B, S * k), generator=g, dtype=torch.int32)
    idx[:, :40] = 0
    indptr = O.fixed_indptr(S, k)
    cp, ri, pm = O.csr2csc(indptr, idx, S)
    for b in range(B):
        order = sorted(range(S * k), key=lambda e: (int(idx[b, e]), e))
        assert pm[b].tolist() == order
        assert ri[b].tolist() == [e // k for e in order]
        counts = torch.bincount(idx[b].long(), minlength=S)
        assert torch.equal(cp[b, 1:] - cp[b, :-1], counts.int())


# ---- routed FFN ------------------------------------------------------------------------------------------
def _ffn_args(state, prefix_map):
    return {k: state[v] for k, v in prefix_map.items()}


def test_routed_ffn_matches_reference_module():
    gd = gold("routed_ffn")["routed_ffn"]
    st, cfg = gd["state"], gd["cfg"]
    x = gd["x"].clone().requires_grad_()
    params = {n: st[n].clone().requires_grad_() for n in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")}
    y = O.routed_ffn(x, st["router.0.weight"], st["router.0.bias"], params["fc1.weight"], params["fc1.bias"],
                     params["fc2.weight"], params["fc2.bias"], cfg["block_size"], cfg["k_active"])
    y.sum().backward()
    assert torch.allclose(y, gd["y"], atol=1e-5)
    assert torch.allclose(x.grad, gd["grads"]["x"], atol=1e-5)
    for n, p in params.items():
        assert torch.allclose(p.grad, gd["grads"][n], atol=1e-4), n


def test_routed_llama_ffn_matches_reference_module():
    gd = gold("routed_ffn")["routed_llama_ffn"]
    st, cfg = gd["state"], gd["cfg"]
    x = gd["x"].clone().requires_grad_()
    params = {n: st[n].clone().requires_grad_() for n in ("gate.weight", "side.weight", "down.weight")}
    y = O.routed_llama_ffn(x, st["router.0.weight"], st["router.0.bias"], params["gate.weight"],
                           params["side.weight"], params["down.weight"], cfg["block_size"], cfg["k_active"])
    y.sum().backward()
    assert torch.allclose(y, gd["y"], atol=1e-5)
    assert torch.allclose(x.grad, gd["grads"]["x"], atol=1e-5)
    for n, p in params.items():
        assert torch.allclose(p.grad, gd["grads"][n], atol=1e-4), n


def test_lora_routed_ffn_matches_reference_module():
    gd = gold("routed_ffn")["lora_routed_ffn"]
    st, cfg = gd["state"], gd["cfg"]
    x = gd["x"].clone().requires_grad_()
    names = ["router.0.weight", "router.0.bias", "fc1.lora.left.weight", "fc1.lora.right.weight",
             "fc2.lora.left.weight", "fc2.lora.right.weight"]
    p = {n: st[n].clone().requires_grad_() for n in names}
    y = O.lora_routed_ffn(x, p["router.0.weight"], p["router.0.bias"], st["fc1.weight"], st["fc1.bias"],
                          st["fc2.weight"], st["fc2.bias"], p["fc1.lora.left.weight"], p["fc1.lora.right.weight"],
                          p["fc2.lora.left.weight"], p["fc2.lora.right.weight"], cfg["block_size"], cfg["k_active"])
    y.sum().backward()
    assert torch.allclose(y, gd["y"], atol=1e-5)
    assert torch.allclose(x.grad, gd["grads"]["x"], atol=1e-5)
    for n in names:
        assert torch.allclose(p[n].grad, gd["grads"][n], atol=1e-4), n


def test_lora_routed_llama_ffn_matches_reference_module():
    gd = gold("routed_ffn")["lora_routed_llama_ffn"]
    st, cfg = gd["state"], gd["cfg"]
    x = gd["x"].clone().requires_grad_()
    names = ["router.0.weight", "router.0.bias"] + [f"{l}.lora.{s}.weight" for l in ("gate", "side", "down")
                                                    for s in ("left", "right")]
    p = {n: st[n].clone().requires_grad_() for n in names}
    y = O.lora_routed_llama_ffn(
        x, p["router.0.weight"], p["router.0.bias"], st["gate.weight"], st["side.weight"], st["down.weight"],
        p["gate.lora.left.weight"], p["gate.lora.right.weight"], p["side.lora.left.weight"],
        p["side.lora.right.weight"], p["down.lora.left.weight"], p["down.lora.right.weight"],
        cfg["block_size"], cfg["k_active"])
    y.sum().backward()
    assert torch.allclose(y, gd["y"], atol=1e-5)
    assert torch.allclose(x.grad, gd["grads"]["x"], atol=1e-5)
    for n in names:
        assert torch.allclose(p[n].grad, gd["grads"][n], atol=1e-4), n
