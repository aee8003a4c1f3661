"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement ("oracle") of the SPT hot path: PQ sparse multi-head attention
(cdist -> lookup -> sddmm -> scale/clamp -> softmax -> spmm, forward and backward)
and the routed FFN.  Reference: ytgui/SPT-proto; every function cites the reference
file:line (relative to /root/reference) whose behaviour it restates.

Who may import this module: tests/, __graft_entry__.smoke(), and bench.py's
`cpu_baseline` / `--impl reference` legs — as the checker or the reported CPU
baseline only.  The product package (spt_proto_b200/) never imports it and fails
loudly if its CUDA library is missing.

Parity status ("pinning"):
  * cdist / PQ codes  : pinned against the reference's own torch path PQV1
                        (quantizer.py:53-62, torch.cdist(p=1)+argmin) via tests/golden/.
  * lookup            : the reference's tests only pin recall > 0.8 (test_lookup.py:75).
                        Two independent restatements live here (literal C emulation of
                        lookup.cu and the abstract per-lane spec) and are checked against each
                        other; on the GPU box they are checked against the compiled reference
                        kernel (oracle/_ref/ext_ref.so) for S in {256,512,1024}.
  * sddmm/softmax/spmm: pinned by the dense torch formulas the reference's tests use
                        (test_sddmm.py:58-62, test_softmax.py:70, test_spmm.py:56).
  * routed FFN        : pinned against the reference's RoutedFFN / LoRARoutedFFN modules
                        imported from /root/reference (fixtures in tests/golden/).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "liboracle_c.so")
_lib = None


def build_c(force: bool = False) -> str:
    """Compile oracle/spt_oracle_c.c with gcc (no fast-math: keeps fp32 summation order)."""
    src = os.path.join(_HERE, "spt_oracle_c.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-fno-fast-math", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def _clib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_lookup_literal.argtypes = [i32p, i32p, i32p] + [ctypes.c_int] * 4
        _lib.oracle_lookup_literal.restype = ctypes.c_int
        _lib.oracle_cdist_forward.argtypes = [f32p, f32p, f32p, i32p] + [ctypes.c_int] * 4
        _lib.oracle_cdist_forward.restype = ctypes.c_int
        _lib.oracle_csr2csc.argtypes = [i32p] * 5 + [ctypes.c_int] * 4
        _lib.oracle_csr2csc.restype = ctypes.c_int
    return _lib


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _f32p(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.POINTER(ctypes.c_float)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------------------
# cdist  (extension/cdist.cu, kernels/cdist.py, quantizer.py:43-77)
# --------------------------------------------------------------------------------------
def cdist_forward(query: torch.Tensor, table: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """query [m,n,dc], table [m,c,dc] -> (distance [m,n,c] f32, indices [m,n] i32).

    L1 distance accumulated over i ascending in fp32 (cdist.cu:46-51); running strict-'<'
    minimum => lowest index wins ties (cdist.cu:52-54).  Non-fp32 inputs are upcast first
    ("bf16 = upcast to fp32, then reference math", SURVEY.md §7 hard part 3).
    """
    q = query.detach().to(torch.float32).contiguous()
    t = table.detach().to(torch.float32).contiguous()
    dist = torch.zeros(q.shape[0], q.shape[1], t.shape[1], dtype=torch.float32)
    for i in range(q.shape[-1]):  # fixed order i = 0..dc-1; every step is one fp32 rounding
        dist = dist + (q[:, :, None, i] - t[:, None, :, i]).abs()
    idx = torch.argmin(dist, dim=-1).to(torch.int32)  # torch.argmin: first minimal index
    # argmin over ties: enforce "lowest index" explicitly rather than rely on torch
    mn = dist.min(dim=-1, keepdim=True).values
    first = (dist == mn).to(torch.int32).argmax(dim=-1).to(torch.int32)
    assert torch.equal(first, idx)
    return dist, idx


def cdist_forward_c(query: np.ndarray, table: np.ndarray, want_distance: bool = True):
    """Same as cdist_forward but through the scalar C loop (literal cdist.cu:42-55 order)."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    t = np.ascontiguousarray(table, dtype=np.float32)
    m, n, dc = q.shape
    c = t.shape[1]
    dist = np.empty((m, n, c), dtype=np.float32) if want_distance else None
    idx = np.empty((m, n), dtype=np.int32)
    rc = _clib().oracle_cdist_forward(_f32p(q), _f32p(t), _f32p(dist), _i32p(idx), m, n, c, dc)
    assert rc == 0
    return dist, idx


def cdist_backward(query: torch.Tensor, table: torch.Tensor, grad_distance: torch.Tensor):
    """cdist.cu:72-182.  sgn(x) = +1 if x > 0 else -1 (note sgn(0) = -1, cdist.cu:117,168-171).
    grad_query[s,n,i] =  sum_c sgn(q-t) g[s,n,c];  grad_table[s,c,i] = -sum_n sgn(q-t) g[s,n,c]."""
    q = query.detach().to(torch.float32)
    t = table.detach().to(torch.float32)
    g = grad_distance.detach().to(torch.float32)
    diff = q[:, :, None, :] - t[:, None, :, :]                      # [m,n,c,dc]
    sgn = torch.where(diff > 0, torch.ones_like(diff), -torch.ones_like(diff))
    gq = (sgn * g[..., None]).sum(dim=2)
    gt = -(sgn * g[..., None]).sum(dim=1)
    return gq, gt


def pq_encode(z: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """PQBase.forward(mode='encode'), quantizer.py:26-77: z [..., m*dc] -> codes [..., m] int32."""
    m, c, dc = weight.shape
    shape = list(z.shape[:-1]) + [m]
    zf = z.reshape(-1, m, dc).transpose(0, 1).contiguous()         # [m, n, dc]  (quantizer.py:43-48)
    _, idx = cdist_forward(zf, weight)
    return idx.transpose(0, 1).reshape(shape).contiguous()          # quantizer.py:74-77


def pq_train_loss(z: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """PQBase.forward(mode='train') loss, quantizer.py:81-111 (differentiable in z and weight)."""
    m, c, dc = weight.shape
    zf = z.reshape(-1, m, dc).transpose(0, 1)                        # [m, n, dc]
    dist = (zf[:, :, None, :].float() - weight[:, None, :, :].float()).abs().sum(-1)
    with torch.no_grad():
        _, idx = cdist_forward(zf, weight)
    z_q = torch.gather(weight, 1, idx.long().unsqueeze(-1).expand(-1, -1, dc))
    attn = torch.softmax(-torch.log(torch.clamp(dist, min=1e-5)), dim=-1)
    z_w = torch.matmul(attn, weight)
    mse = torch.nn.functional.mse_loss
    return mse(z_w, z_q) + mse(zf, z_q)


# --------------------------------------------------------------------------------------
# lookup  (extension/lookup.cu, kernels/lookup.py)
# --------------------------------------------------------------------------------------
def lookup_forward(query: torch.Tensor, key: torch.Tensor, sparse_coeff: int) -> torch.Tensor:
    """Literal emulation (C) of lookup_forward_kernel.  query/key [B,S,m] int32 -> [B,S,S//sparse_coeff]."""
    q = np.ascontiguousarray(query.detach().cpu().numpy(), dtype=np.int32)
    k = np.ascontiguousarray(key.detach().cpu().numpy(), dtype=np.int32)
    B, S, m = q.shape
    assert S % sparse_coeff == 0
    nnz = S // sparse_coeff
    out = np.zeros((B, S, nnz), dtype=np.int32)
    rc = _clib().oracle_lookup_literal(_i32p(q), _i32p(k), _i32p(out), B, S, m, nnz)
    if rc != 0:
        raise RuntimeError(f"oracle_lookup_literal: unsupported shape (rc={rc})")
    return torch.from_numpy(out)


def lookup_spec(query: torch.Tensor, key: torch.Tensor, sparse_coeff: int) -> torch.Tensor:
    """Second, independent restatement of lookup.cu as an abstract per-lane specification
    (SURVEY.md §8 a-2).  Pure numpy/Python, for small cases; must equal lookup_forward().

    Row r, lane t in 0..3 owns keys j = t (mod 4), j <= r, scanned ascending.
    bucket(j) = min(3, matches(r,j) // (m // 4)).  Lane t fills output positions t, t+4, ...
    (< min(r+1, k)) with its keys ordered (bucket descending, j ascending).  A (lane,bucket)
    list keeps at most cap_t = k/4 (t<2) or k/4-1 (t>=2) entries; later entries of lanes 2/3
    land on lane 1/0's LAST slot of that bucket (position k-3 / k-4); the latest writer in scan
    order wins, where scan order is the group-of-4 index j // 4 (one warp instruction handles keys
    4g..4g+3) and, inside one instruction, the LOWEST lane wins (measured on B200 against the
    compiled reference kernel); later entries of lanes 0/1 are dropped.
    """
    q = query.detach().cpu().numpy().astype(np.int64) & 0xFFFF
    kk = key.detach().cpu().numpy().astype(np.int64) & 0xFFFF
    B, S, m = q.shape
    nnz = S // sparse_coeff
    assert nnz % 4 == 0 and nnz >= 8 and m >= 4
    div = m // 4
    quarter = nnz // 4
    cap = [quarter, quarter, quarter - 1, quarter - 1]
    out = np.zeros((B, S, nnz), dtype=np.int32)
    for b in range(B):
        for r in range(S):
            cnt = (q[b, r][None, :] == kk[b, : r + 1]).sum(-1)
            bucket = np.minimum(3, cnt // div)
            lists = [[[] for _ in range(4)] for _ in range(4)]
            for j in range(r + 1):
                lists[j % 4][bucket[j]].append(j)
            lim = min(r + 1, nnz)
            for t in range(4):
                chain = []
                for s in (3, 2, 1, 0):
                    own = lists[t][s]
                    stored = list(own[: cap[t]])
                    if t < 2 and len(own) >= quarter:
                        partner = lists[3 - t][s]
                        if len(partner) >= quarter:           # partner overflowed into our last slot
                            if partner[-1] // 4 > stored[quarter - 1] // 4:   # same instruction: owner wins
                                stored[quarter - 1] = partner[-1]
                    chain.extend(stored)
                n_t = len(range(t, lim, 4))
                chain = chain[:n_t]
                out[b, r, t: t + 4 * len(chain): 4] = chain
    return torch.from_numpy(out)


def exact_topk_match_count(query: torch.Tensor, key: torch.Tensor, nnz: int):
    """The reference test's own oracle for lookup (test_lookup.py:46-50,66-69): exact top-k by
    match count under the causal mask.  Used only for the recall check."""
    cmp = (query.unsqueeze(-2) == key.unsqueeze(-3)).sum(-1)
    S = query.shape[1]
    mask = torch.tril(torch.ones(S, S, dtype=torch.bool))
    return torch.where(mask, cmp, torch.full_like(cmp, -1))


# --------------------------------------------------------------------------------------
# CSR helpers: sddmm / softmax / spmm / csr2csc   (kernels/sddmm.py, softmax.py, spmm.py)
# --------------------------------------------------------------------------------------
def _row_of_entry(indptr: torch.Tensor, nnz: int) -> torch.Tensor:
    counts = (indptr[1:] - indptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(counts.numel()), counts)
    assert rows.numel() == nnz
    return rows


def sddmm_forward(indptr, indices, query, key) -> torch.Tensor:
    """values[b, e] = <query[b, row(e), :], key[b, indices[b, e], :]>  (sddmm.cpp:52-69 with
    op(A)=N, op(B)=T as called from kernels/sddmm.py:19-22).  fp32 accumulation."""
    rows = _row_of_entry(indptr, indices.shape[1])
    qg = query.float()[:, rows, :]                                           # [B, nnz, d]
    kg = torch.gather(key.float(), 1, indices.long().unsqueeze(-1).expand(-1, -1, key.shape[-1]))
    return (qg * kg).sum(-1)


def spmm_forward(trans_lhs: bool, indptr, indices, values, x) -> torch.Tensor:
    """y = op(A) @ x with A the batched CSR (spmm.cpp:52-69).  trans_lhs=True is the transposed
    product used for dK / dV (kernels/sddmm.py:46-49, kernels/spmm.py:44-47)."""
    B, nnz = indices.shape
    S, d = x.shape[1], x.shape[2]
    rows = _row_of_entry(indptr, nnz).unsqueeze(0).expand(B, -1)
    cols = indices.long()
    src, dst = (rows, cols) if trans_lhs else (cols, rows)
    contrib = values.float().unsqueeze(-1) * torch.gather(x.float(), 1, src.unsqueeze(-1).expand(-1, -1, d))
    y = torch.zeros(B, S, d, dtype=torch.float32)
    y.scatter_add_(1, dst.unsqueeze(-1).expand(-1, -1, d), contrib)
    return y


def softmax_forward(indptr, indices, values) -> torch.Tensor:
    """softmax.cu:16-46: no max-subtraction, causal predicate (index <= row) as a 0/1 factor,
    denominator clamped to >= 1e-9."""
    rows = _row_of_entry(indptr, indices.shape[1]).unsqueeze(0)
    keep = (indices.long() <= rows).float()
    e = torch.exp(values.float()) * keep
    denom = torch.zeros(values.shape[0], indptr.numel() - 1, dtype=torch.float32)
    denom.scatter_add_(1, rows.expand_as(e), e)
    denom = torch.clamp(denom, min=1e-9)
    return e / torch.gather(denom, 1, rows.expand_as(e))


def softmax_backward(indptr, indices, output, grad_output, reference_clamp: bool = False) -> torch.Tensor:
    """True softmax gradient dv = y * (dy - sum(y*dy)) restricted to kept entries.  The shipped
    CUDA kernel additionally clamps sum(y*dy) to >= 1e-9 (softmax.cu:69), which is wrong when the
    sum is negative; north_star pins gradients to the reference's *torch* formulation, so the true
    gradient is the contract.  reference_clamp=True reproduces the kernel's behaviour."""
    rows = _row_of_entry(indptr, indices.shape[1]).unsqueeze(0)
    keep = (indices.long() <= rows).float()
    prod = output.float() * grad_output.float() * keep
    c = torch.zeros(output.shape[0], indptr.numel() - 1, dtype=torch.float32)
    c.scatter_add_(1, rows.expand_as(prod), prod)
    if reference_clamp:
        c = torch.clamp(c, min=1e-9)
    return output.float() * (grad_output.float() - torch.gather(c, 1, rows.expand_as(prod))) * keep


def csr2csc(indptr: torch.Tensor, indices: torch.Tensor, n_cols: int):
    """Stable transpose (legacy/csr2csc.cpp:3-54 semantics): returns (col_ptr [B,n_cols+1],
    row_idx [B,nnz], perm [B,nnz]) with rows ascending inside each column, ties in CSR order."""
    ip = np.ascontiguousarray(indptr.cpu().numpy(), dtype=np.int32)
    ix = np.ascontiguousarray(indices.cpu().numpy(), dtype=np.int32)
    B, nnz = ix.shape
    n_rows = ip.shape[0] - 1
    cp = np.empty((B, n_cols + 1), dtype=np.int32)
    ri = np.empty((B, nnz), dtype=np.int32)
    pm = np.empty((B, nnz), dtype=np.int32)
    rc = _clib().oracle_csr2csc(_i32p(ip), _i32p(ix), _i32p(cp), _i32p(ri), _i32p(pm), B, n_rows, n_cols, nnz)
    if rc != 0:
        raise RuntimeError("oracle_csr2csc failed")
    return torch.from_numpy(cp), torch.from_numpy(ri), torch.from_numpy(pm)


# --------------------------------------------------------------------------------------
# Sparse attention layer glue (layers/sparse/attention.py:84-142)
# --------------------------------------------------------------------------------------
def fixed_indptr(seq_length: int, top_k: int) -> torch.Tensor:
    """attention.py:115-119."""
    return torch.arange(0, top_k * seq_length + 1, step=top_k, dtype=torch.int32)


def sparse_attention_indices(q: torch.Tensor, k: torch.Tensor, weight: torch.Tensor, sparse_coeff: int = 8):
    """q,k [B,S,d] -> (indptr [S+1], indices [B, S*topk]) exactly as _get_attn builds them
    (attention.py:105-119)."""
    q_c = pq_encode(q, weight)
    k_c = pq_encode(k, weight)
    topk = lookup_forward(q_c, k_c, sparse_coeff)
    S = q.shape[1]
    return fixed_indptr(S, S // sparse_coeff), topk.flatten(1)


def sparse_attention_values(indptr, indices, q, k, v, scaling: float):
    """Differentiable (torch autograd) gathered-form restatement of
    sddmm -> clamp_(scaling * ., -10, 10) -> softmax -> spmm (attention.py:122-141).
    q,k,v [B,S,d] fp32 (requires_grad allowed) -> (y [B,S,d], probs [B, nnz])."""
    B, S, d = q.shape
    nnz = indices.shape[1]
    rows = _row_of_entry(indptr, nnz)
    cols = indices.long()
    kg = torch.gather(k, 1, cols.unsqueeze(-1).expand(-1, -1, d))
    s = (q[:, rows, :] * kg).sum(-1)
    s = torch.clamp(scaling * s, min=-10.0, max=10.0)
    keep = (cols <= rows.unsqueeze(0)).to(s.dtype)
    e = torch.exp(s) * keep
    denom = torch.zeros(B, S, dtype=s.dtype).scatter_add(1, rows.unsqueeze(0).expand(B, -1), e)
    denom = torch.clamp(denom, min=1e-9)
    p = e / denom[:, rows]
    vg = torch.gather(v, 1, cols.unsqueeze(-1).expand(-1, -1, d))
    y = torch.zeros(B, S, d, dtype=s.dtype).scatter_add(
        1, rows.view(1, -1, 1).expand(B, -1, d), p.unsqueeze(-1) * vg)
    return y, p


def sparse_mha_layer(q4, k4, v4, weight, sparse_coeff: int = 8, reference_output_layout: bool = False):
    """SparseVanillaAttentionV2.forward on [N,S,H,E] tensors (attention.py:84-142 via
    basic/attention.py:41-57): transposes to [N*H,S,E], builds the CSR, applies it, transposes back.

    reference_output_layout=True reproduces a quirk of the shipped layer: _apply_attn un-transposes
    the 3-D result [N*H, S, E] with `y.transpose(1, 2).contiguous().view(v_size)`
    (attention.py:139-142), which swaps S and E instead of S and H, so the returned [N,S,H,E] tensor
    is a re-interpretation of [N*H, E, S] memory.  The reference's only layer test feeds all-ones
    (test_sparse_mha.py:7-43) and cannot see it.  The default (False) returns the mathematically
    intended layout, i.e. what the reference's dense VanillaAttention returns on the same pattern."""
    N, S, H, E = q4.shape
    q = q4.transpose(1, 2).reshape(N * H, S, E)
    k = k4.transpose(1, 2).reshape(N * H, S, E)
    v = v4.transpose(1, 2).reshape(N * H, S, E)
    with torch.no_grad():
        indptr, indices = sparse_attention_indices(q.detach(), k.detach(), weight.detach(), sparse_coeff)
    y, _ = sparse_attention_values(indptr, indices, q.float(), k.float(), v.float(), float(E) ** -0.5)
    if reference_output_layout:
        return y.transpose(1, 2).contiguous().view(N, S, H, E)
    return y.reshape(N, H, S, E).transpose(1, 2).contiguous()


# --------------------------------------------------------------------------------------
# Routed FFN (layers/sparse/feedforward.py:47-103, tuning/lora_ffn.py:52-115)
# --------------------------------------------------------------------------------------
def route_topk_mask(prob: torch.Tensor, k: int) -> torch.Tensor:
    """Block membership mask [T, nb] of torch.topk(prob, k) (feedforward.py:57-70).  Only set
    membership matters; ties are broken towards the LOWEST block index (stable sort), which is
    the tie-break the CUDA router kernel defines."""
    order = torch.sort(prob, dim=-1, descending=True, stable=True).indices[:, :k]
    mask = torch.zeros_like(prob, dtype=torch.bool)
    mask.scatter_(1, order, True)
    return mask


def routed_ffn(x, router_w, router_b, w1, b1, w2, b2, block_size: int, k_active: int, activation=torch.relu):
    """Masked-dense restatement of RoutedFFN._apply_ffn (feedforward.py:47-85): equivalent to
    fc2(act(mask * fc1(x))) when act(0) = 0 — the form the reference's own test uses as oracle
    (test_sparse_ffn.py:8-38).  Here the mask is applied AFTER the activation so that it also
    holds for activations with act(0) != 0, matching the per-block loop exactly."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    prob = torch.sigmoid(x2 @ router_w.t() + router_b)
    mask = route_topk_mask(prob.detach(), k_active)
    mask_f = mask.repeat_interleave(block_size, dim=-1).to(x2.dtype)
    h = activation(x2 @ w1.t() + b1) * mask_f
    y = h @ w2.t() + b2
    return y.reshape(shape)


def routed_llama_ffn(x, router_w, router_b, w_gate, w_side, w_down, block_size: int, k_active: int,
                     activation=torch.nn.functional.silu):
    """RoutedLLaMaFFN._apply_ffn (feedforward.py:144-180)."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    prob = torch.sigmoid(x2 @ router_w.t() + router_b)
    mask = route_topk_mask(prob.detach(), k_active)
    mask_f = mask.repeat_interleave(block_size, dim=-1).to(x2.dtype)
    h = activation(x2 @ w_gate.t()) * (x2 @ w_side.t()) * mask_f
    return (h @ w_down.t()).reshape(shape)


def lora_routed_ffn(x, router_w, router_b, w1, b1, w2, b2, l1_left, l1_right, l2_left, l2_right,
                    block_size: int, k_active: int, activation=torch.relu):
    """LoRARoutedFFN.forward (tuning/lora_ffn.py:52-115) in masked-dense form:
       coeff[t,i] = 2 * prob[t,i] for active (t,i);
       h = act(coeff * (x W1_i^T + b1_i) + (x L1) R1_i^T);
       y += coeff * (h W2_i) + (h L2_i) R2^T;  y += b2.
    l1_left [d, r], l1_right [F, r], l2_left [F, r], l2_right [d, r] (nn.Embedding weights)."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    nb = w1.shape[0] // block_size
    prob = torch.sigmoid(x2 @ router_w.t() + router_b)
    mask = route_topk_mask(prob.detach(), k_active)
    y = torch.zeros_like(x2)
    for i in range(nb):
        sl = slice(i * block_size, (i + 1) * block_size)
        m_i = mask[:, i].to(x2.dtype).unsqueeze(-1)
        coeff = 2.0 * prob[:, i: i + 1]
        h = coeff * (x2 @ w1[sl].t() + b1[sl]) + (x2 @ l1_left) @ l1_right[sl].t()
        h = activation(h)
        y = y + m_i * (coeff * (h @ w2[:, sl].t()) + (h @ l2_left[sl]) @ l2_right.t())
    y = y + b2
    return y.reshape(shape)


def lora_routed_llama_ffn(x, router_w, router_b, w_gate, w_side, w_down, lg_left, lg_right, ls_left, ls_right,
                          ld_left, ld_right, block_size: int, k_active: int, activation=torch.nn.functional.silu):
    """LoRARoutedLLaMaFFN.forward (tuning/lora_ffn.py:164-225) in masked-dense form:
       g = coeff (x Wg_i^T) + (x Lg) Rg_i^T;  s = coeff (x Ws_i^T) + (x Ls) Rs_i^T;  h = act(g) * s;
       y += coeff (h Wd_i) + (h Ld_i) Rd^T.   *_left [in, r], *_right [out, r]."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    nb = w_gate.shape[0] // block_size
    prob = torch.sigmoid(x2 @ router_w.t() + router_b)
    mask = route_topk_mask(prob.detach(), k_active)
    y = torch.zeros_like(x2)
    for i in range(nb):
        sl = slice(i * block_size, (i + 1) * block_size)
        m_i = mask[:, i].to(x2.dtype).unsqueeze(-1)
        coeff = 2.0 * prob[:, i: i + 1]
        g = coeff * (x2 @ w_gate[sl].t()) + (x2 @ lg_left) @ lg_right[sl].t()
        s = coeff * (x2 @ w_side[sl].t()) + (x2 @ ls_left) @ ls_right[sl].t()
        h = activation(g) * s
        y = y + m_i * (coeff * (h @ w_down[:, sl].t()) + (h @ ld_left[sl]) @ ld_right.t())
    return y.reshape(shape)
