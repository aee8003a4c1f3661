"""GPU parity of the routed-FFN path: the tcgen05 grouped GEMM against torch.matmul on the same
bf16-rounded operands (fp32 accumulation both sides: atol 2e-2 + rtol 1e-2 on fp32 outputs of O(10)
magnitude), then the bucketing and the full layers against the oracle / golden fixtures."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _groups(sizes):
    """padded (x128) bucket layout for the given real group sizes"""
    ptr, tiles = [0], []
    for g, n in enumerate(sizes):
        pad = (n + 127) // 128 * 128
        ptr.append(ptr[-1] + pad)
        tiles += [g] * (pad // 128)
    return ptr, tiles


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("sizes,N,K", [([128, 256], 128, 64), ([300, 0, 77, 513], 256, 192), ([1000, 900], 384, 1024)])
def test_grouped_gemm_mode0(b_mn, sizes, N, K):
    from spt_proto_b200 import ext
    g = torch.Generator().manual_seed(N + K + sum(sizes))
    G = len(sizes)
    ptr, tiles = _groups(sizes)
    R = ptr[-1]
    A = torch.zeros(R, K)
    for gi, n in enumerate(sizes):
        A[ptr[gi]: ptr[gi] + n] = torch.randn(n, K, generator=g)
    A = A.bfloat16()
    W = torch.randn(G * N, K, generator=g).bfloat16()             # block g = rows [g*N, (g+1)*N)
    bias = torch.randn(G * N, generator=g)
    scale = torch.rand(R, generator=g) + 0.5
    tile_group = torch.tensor(tiles + [-1, -1], dtype=torch.int32)
    want = torch.zeros(R, N)
    for gi in range(G):
        rows = slice(ptr[gi], ptr[gi + 1])
        want[rows] = torch.relu(A[rows].float() @ W[gi * N:(gi + 1) * N].float().t() + bias[gi * N:(gi + 1) * N]) \
            * scale[rows, None]
    out = torch.full((R + 256, N), float("nan"), device=DEV)
    if not b_mn:   # B stored [G*N, K] (K-major), group offset along N (rows)
        ext.grouped_gemm(0, A.to(DEV), False, W.to(DEV), False, tile_group=tile_group.to(DEV), N=N, K=K,
                         b_mn_off=N, out=out, bias=bias.to(DEV), bias_stride=N, row_scale=torch.cat(
                             [scale, torch.ones(256)]).to(DEV), act=1)
    else:          # B stored [K, G*N] (MN-major), group offset along N (columns)
        Wt = W.t().contiguous()
        ext.grouped_gemm(0, A.to(DEV), False, Wt.to(DEV), True, tile_group=tile_group.to(DEV), N=N, K=K,
                         b_mn_off=N, out=out, bias=bias.to(DEV), bias_stride=N, row_scale=torch.cat(
                             [scale, torch.ones(256)]).to(DEV), act=1)
    got = out.cpu()
    assert (got[R:] == 0).all()                                    # tail tiles (-1) are defined: zeros
    assert torch.allclose(got[:R], want, atol=2e-2, rtol=1e-2)


def test_grouped_gemm_mode0_fc2_layout():
    """fc2: B_g = W2[:, g*bs:(g+1)*bs] — K-major with a K offset per group and leading dim F; bf16 output."""
    from spt_proto_b200 import ext
    g = torch.Generator().manual_seed(5)
    sizes, bs, d = [200, 129, 384], 128, 256
    G = len(sizes)
    ptr, tiles = _groups(sizes)
    R = ptr[-1]
    H = torch.randn(R, bs, generator=g).bfloat16()
    W2 = torch.randn(d, G * bs, generator=g).bfloat16()
    want = torch.zeros(R, d)
    for gi in range(G):
        rows = slice(ptr[gi], ptr[gi + 1])
        want[rows] = H[rows].float() @ W2[:, gi * bs:(gi + 1) * bs].float().t()
    out = torch.empty(R, d, device=DEV, dtype=torch.bfloat16)
    ext.grouped_gemm(0, H.to(DEV), False, W2.to(DEV), False, tile_group=torch.tensor(tiles, dtype=torch.int32).to(DEV),
                     N=d, K=bs, b_k_off=bs, out=out)
    assert torch.allclose(out.float().cpu(), want, atol=1e-1, rtol=2e-2)


@pytest.mark.parametrize("sizes,M,N", [([128, 256], 128, 128), ([384, 0, 128, 640], 256, 192), ([1024, 896], 512, 320)])
def test_grouped_gemm_mode1_weight_grad(sizes, M, N):
    """dW_g [M, N] = dU_g^T X_g over the group's (padded) rows: both operands MN-major, in place."""
    from spt_proto_b200 import ext
    g = torch.Generator().manual_seed(M + N)
    G = len(sizes)
    ptr = [0]
    for n in sizes:
        ptr.append(ptr[-1] + n)
    R = ptr[-1]
    dU = torch.randn(R, M, generator=g).bfloat16()
    X = torch.randn(R, N, generator=g).bfloat16()
    want = torch.stack([dU[ptr[i]:ptr[i + 1]].float().t() @ X[ptr[i]:ptr[i + 1]].float() for i in range(G)])
    out = torch.full((G * M, N), float("nan"), device=DEV)
    ext.grouped_gemm(1, dU.to(DEV), True, X.to(DEV), True, group_ptr=torch.tensor(ptr, dtype=torch.int32).to(DEV),
                     M=M, N=N, c_row_off=M, out=out)
    assert torch.allclose(out.cpu().view(G, M, N), want, atol=5e-2, rtol=1e-2)
    # dW2-style: output blocks side by side along the columns (c_col_off), leading dim G*N
    out2 = torch.full((M, G * N), float("nan"), device=DEV)
    ext.grouped_gemm(1, dU.to(DEV), True, X.to(DEV), True, group_ptr=torch.tensor(ptr, dtype=torch.int32).to(DEV),
                     M=M, N=N, c_col_off=N, out=out2)
    assert torch.allclose(out2.cpu().view(M, G, N).transpose(0, 1), want, atol=5e-2, rtol=1e-2)


# ---------------------------------------------------------------------------------- routing / bucketing
@pytest.mark.parametrize("T,nb,k", [(1000, 8, 4), (8192, 4, 2), (77, 16, 4), (513, 8, 2)])
def test_route_bucket_matches_oracle(T, nb, k):
    from oracle import spt_oracle as O
    from spt_proto_b200 import ext
    g = torch.Generator().manual_seed(T + nb)
    prob = torch.sigmoid(torch.randn(T, nb, generator=g))
    prob[::7, 1] = prob[::7, 0]                       # exact ties: lowest block index must win
    prob[::11] = 0.5
    mask = O.route_topk_mask(prob, k)
    b = ext.route_bucket(prob.to(DEV), k)
    ptr, rows = b.bucket_ptr.cpu(), b.bucket_rows.cpu()
    assert torch.equal(rows, mask.sum(0).int())
    row_token, row_prob, token_rows, tile_group = b.row_token.cpu(), b.row_prob.cpu(), b.token_rows.cpu(), b.tile_group.cpu()
    for gi in range(nb):
        assert ptr[gi] % 128 == 0
        want = torch.nonzero(mask[:, gi]).flatten().int()
        assert torch.equal(row_token[ptr[gi]: ptr[gi] + rows[gi]], want)          # ascending token order
        assert (row_token[ptr[gi] + rows[gi]: ptr[gi + 1]] == -1).all()           # padding
        assert torch.equal(row_prob[ptr[gi]: ptr[gi] + rows[gi]], prob[want.long(), gi])
        assert (tile_group[ptr[gi] // 128: ptr[gi + 1] // 128] == gi).all()
    assert (tile_group[ptr[nb] // 128:] == -1).all()
    # token_rows: the token's active blocks in ascending order
    for t in range(0, T, 13):
        blocks = torch.nonzero(mask[t]).flatten().tolist()
        for j, gi in enumerate(blocks):
            r = token_rows[t, j].item()
            assert ptr[gi] <= r < ptr[gi + 1] and row_token[r] == t


# ---------------------------------------------------------------------------------- layers vs oracle / golden
def _bf(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("T,d,Fdim,bs", [(512, 128, 1024, 128), (300, 256, 1024, 256), (2048, 256, 2048, 256)])
def test_routed_ffn_layer_matches_oracle(T, d, Fdim, bs):
    """RoutedFFN fwd+bwd vs the masked-dense oracle (the reference test's own formulation,
    test_sparse_ffn.py:8-38) on bf16-rounded weights/inputs.  Tolerances: bf16 GEMM operands and bf16
    intermediates (h, per-block partial outputs) => relative Frobenius error < 1e-2."""
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    torch.manual_seed(T + d)
    ffn = layers.RoutedFFN(d_model=d, d_feedforward=Fdim, block_size=bs, activation=torch.nn.ReLU()).to(DEV)
    with torch.no_grad():
        for p in ffn.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(4, T // 4, d)).to(DEV).requires_grad_()
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    xc = x.detach().cpu().requires_grad_()
    ps = {n: sd[n].clone().requires_grad_() for n in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")}
    y_ref = O.routed_ffn(xc, sd["router.0.weight"], sd["router.0.bias"], ps["fc1.weight"], ps["fc1.bias"],
                         ps["fc2.weight"], ps["fc2.bias"], bs, (Fdim // bs) // 2)
    y_ref.backward(dy.cpu())

    def rel(a, b):
        return ((a.float().cpu() - b).norm() / b.norm()).item()

    assert rel(y, y_ref.detach()) < 1e-2
    assert rel(x.grad, xc.grad) < 1e-2
    got = dict(ffn.named_parameters())
    for n, p in ps.items():
        assert rel(got[n].grad, p.grad) < 1e-2, n
    assert got["router.0.weight"].grad is None          # plain RoutedFFN: the router gets no gradient


def test_routed_ffn_matches_reference_golden():
    """Golden vectors from the reference's own RoutedFFN / RoutedLLaMaFFN modules (tests/golden/make_golden.py).
    d_model 32, block 16 are far below tensor-core tile sizes: exercises the padding / masking paths."""
    import os
    from spt_proto_b200 import layers
    gd = torch.load(os.path.join(os.path.dirname(__file__), "golden", "routed_ffn.pt"), weights_only=False)
    for key, cls, act in (("routed_ffn", layers.RoutedFFN, torch.nn.ReLU()),
                          ("routed_llama_ffn", layers.RoutedLLaMaFFN, torch.nn.SiLU())):
        case = gd[key]
        cfg = case["cfg"]
        ffn = cls(d_model=cfg["d_model"], d_feedforward=cfg["d_feedforward"], block_size=cfg["block_size"],
                  activation=act).to(DEV)
        ffn.load_state_dict(case["state"])
        x = case["x"].to(DEV).requires_grad_()
        y = ffn(x)
        y.sum().backward()
        # fp32 reference vs bf16 tensor-core path: absolute tolerance scaled to the output magnitude
        scale = case["y"].abs().max().item()
        assert (y.detach().cpu() - case["y"]).abs().max().item() < 3e-2 * max(scale, 1.0), key
        # gradients: a bf16 rounding can flip a ReLU gate that sits at ~0 in fp32, which moves single
        # elements by a whole term — compare in the Frobenius norm, like the weight gradients below
        # (measured: 4e-2 vs this fp32 golden, 3e-3 vs the oracle on bf16-rounded weights, see
        # test_routed_ffn_layer_matches_oracle which is the tight check)
        assert (x.grad.cpu() - case["grads"]["x"]).norm() / case["grads"]["x"].norm() < 6e-2, key
        for n, p in ffn.named_parameters():
            if n in case["grads"] and case["grads"][n] is not None:
                ref = case["grads"][n]
                assert (p.grad.cpu() - ref).norm() / ref.norm() < 6e-2, (key, n)


def test_routed_llama_ffn_layer_matches_oracle():
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    torch.manual_seed(3)
    d, Fdim, bs, T = 256, 2048, 256, 1024
    ffn = layers.RoutedLLaMaFFN(d_model=d, d_feedforward=Fdim, block_size=bs, activation=torch.nn.SiLU()).to(DEV)
    with torch.no_grad():
        for p in ffn.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(T, d)).to(DEV).requires_grad_()
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    xc = x.detach().cpu().requires_grad_()
    ps = {n: sd[n].clone().requires_grad_() for n in ("gate.weight", "side.weight", "down.weight")}
    y_ref = O.routed_llama_ffn(xc, sd["router.0.weight"], sd["router.0.bias"], ps["gate.weight"], ps["side.weight"],
                               ps["down.weight"], bs, (Fdim // bs) // 4)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1.5e-2
    assert rel(x.grad, xc.grad) < 1.5e-2
    got = dict(ffn.named_parameters())
    for n, p in ps.items():
        assert rel(got[n].grad, p.grad) < 1.5e-2, n


# ---------------------------------------------------------------------------------- LoRA routed FFN (a-10)
def _lora_setup(cls, act, d, Fdim, bs, r, T, seed):
    torch.manual_seed(seed)
    ffn = cls(d_lora=r, block_size=bs, d_model=d, d_feedforward=Fdim, activation=act).to(DEV)
    with torch.no_grad():
        for n, p in ffn.named_parameters():
            if "lora.right" in n:
                torch.nn.init.normal_(p, std=0.05)
            p.copy_(_bf(p))
    x = _bf(torch.randn(T, d)).to(DEV).requires_grad_()
    return ffn, x


def test_lora_routed_ffn_layer_matches_oracle():
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    d, Fdim, bs, r, T = 256, 1024, 256, 16, 1024
    ffn, x = _lora_setup(layers.LoRARoutedFFN, torch.nn.ReLU(), d, Fdim, bs, r, T, 11)
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    names = [n for n, p in ffn.named_parameters() if p.requires_grad]
    assert set(names) == {"router.0.weight", "router.0.bias", "fc1.lora.left.weight", "fc1.lora.right.weight",
                          "fc2.lora.left.weight", "fc2.lora.right.weight"}          # frozen base (lora.py:43-44)
    p = {n: sd[n].clone().requires_grad_() for n in names}
    xc = x.detach().cpu().requires_grad_()
    y_ref = O.lora_routed_ffn(xc, p["router.0.weight"], p["router.0.bias"], sd["fc1.weight"], sd["fc1.bias"],
                              sd["fc2.weight"], sd["fc2.bias"], p["fc1.lora.left.weight"], p["fc1.lora.right.weight"],
                              p["fc2.lora.left.weight"], p["fc2.lora.right.weight"], bs, (Fdim // bs) // 2)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1.5e-2
    assert rel(x.grad, xc.grad) < 2e-2
    got = dict(ffn.named_parameters())
    for n in names:
        assert rel(got[n].grad, p[n].grad) < 3e-2, (n, rel(got[n].grad, p[n].grad))
    assert got["fc1.weight"].grad is None and got["fc2.weight"].grad is None


def test_lora_routed_llama_ffn_layer_matches_oracle():
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    d, Fdim, bs, r, T = 256, 1024, 256, 16, 768
    ffn, x = _lora_setup(layers.LoRARoutedLLaMaFFN, torch.nn.SiLU(), d, Fdim, bs, r, T, 12)
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    names = [n for n, p in ffn.named_parameters() if p.requires_grad]
    p = {n: sd[n].clone().requires_grad_() for n in names}
    xc = x.detach().cpu().requires_grad_()
    y_ref = O.lora_routed_llama_ffn(
        xc, p["router.0.weight"], p["router.0.bias"], sd["gate.weight"], sd["side.weight"], sd["down.weight"],
        p["gate.lora.left.weight"], p["gate.lora.right.weight"], p["side.lora.left.weight"],
        p["side.lora.right.weight"], p["down.lora.left.weight"], p["down.lora.right.weight"], bs, (Fdim // bs) // 2)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1.5e-2
    assert rel(x.grad, xc.grad) < 2e-2
    got = dict(ffn.named_parameters())
    for n in names:
        assert rel(got[n].grad, p[n].grad) < 3e-2, (n, rel(got[n].grad, p[n].grad))


def test_lora_routed_ffn_matches_reference_golden():
    import os
    from spt_proto_b200 import layers
    gd = torch.load(os.path.join(os.path.dirname(__file__), "golden", "routed_ffn.pt"), weights_only=False)
    for key, cls, act in (("lora_routed_ffn", layers.LoRARoutedFFN, torch.nn.ReLU()),
                          ("lora_routed_llama_ffn", layers.LoRARoutedLLaMaFFN, torch.nn.SiLU())):
        case, cfg = gd[key], gd[key]["cfg"]
        ffn = cls(d_lora=cfg["d_lora"], block_size=cfg["block_size"], d_model=cfg["d_model"],
                  d_feedforward=cfg["d_feedforward"], activation=act).to(DEV)
        ffn.load_state_dict(case["state"])
        x = case["x"].to(DEV).requires_grad_()
        y = ffn(x)
        y.sum().backward()
        rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
        assert rel(y.detach(), case["y"]) < 2e-2, key                     # fp32 golden vs bf16 tensor-core path
        assert rel(x.grad, case["grads"]["x"]) < 8e-2, key
        for n, p in ffn.named_parameters():
            if p.requires_grad:
                assert rel(p.grad, case["grads"][n]) < 8e-2, (key, n, rel(p.grad, case["grads"][n]))


# ---------------------------------------------------------------------------- fused elementwise glue (lora_fuse.cu)
@pytest.mark.parametrize("a_dt,b_dt,o_dt", [(torch.float32, torch.float32, torch.float32),
                                            (torch.bfloat16, torch.bfloat16, torch.bfloat16),
                                            (torch.float32, torch.bfloat16, torch.bfloat16)])
def test_scale_add_matches_torch(a_dt, b_dt, o_dt):
    from spt_proto_b200.kernels import ffn as F
    g = torch.Generator(device="cpu").manual_seed(5)
    R, C = 384, 1032
    coeff = torch.rand(R, generator=g).to(DEV).requires_grad_()
    a = torch.randn(R, C, generator=g).to(DEV).to(a_dt).requires_grad_()
    b = torch.randn(R, C, generator=g).to(DEV).to(b_dt).requires_grad_()
    out = F.scale_add(coeff, a, b, o_dt)
    ref = (coeff.detach()[:, None] * a.detach().float() + b.detach().float())
    assert out.dtype == o_dt
    tol = 1e-6 if o_dt == torch.float32 else 1.6e-2
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    go = torch.randn(R, C, generator=g).to(DEV).to(o_dt)
    out.backward(go)
    gf = go.float()
    torch.testing.assert_close(a.grad.float(), coeff.detach()[:, None] * gf, atol=tol, rtol=tol)
    torch.testing.assert_close(b.grad.float(), gf, atol=0, rtol=0)
    torch.testing.assert_close(coeff.grad, (gf * a.detach().float()).sum(1), atol=1e-3, rtol=1e-4)


def test_lora_glu_matches_torch_autograd():
    from spt_proto_b200.kernels import ffn as F
    g = torch.Generator(device="cpu").manual_seed(6)
    R, C = 256, 520
    mk = lambda: torch.randn(R, C, generator=g).to(DEV)
    coeff = torch.rand(R, generator=g).to(DEV)
    ts = [mk() for _ in range(4)]
    leaf = [t.clone().requires_grad_() for t in ts]
    cf = coeff.clone().requires_grad_()
    h = F.lora_glu(cf, *leaf)
    ref_in = [t.clone().double().requires_grad_() for t in ts]
    cr = coeff.clone().double().requires_grad_()
    href = torch.nn.functional.silu(cr[:, None] * ref_in[0] + ref_in[1]) * (cr[:, None] * ref_in[2] + ref_in[3])
    assert h.dtype == torch.bfloat16
    torch.testing.assert_close(h.float(), href.float(), atol=1e-2, rtol=1e-2)
    go = torch.randn(R, C, generator=g).to(DEV).to(torch.bfloat16)
    h.backward(go)
    href.backward(go.double())
    for t, r in zip(leaf, ref_in):
        torch.testing.assert_close(t.grad, r.grad.float(), atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(cf.grad, cr.grad.float(), atol=1e-3, rtol=1e-4)


# ---------------------------------------------------------------------------- headline shapes (BENCH / configs[3])
def _exact_router(ffn, gain=4.0):
    """Router whose logits are single products (weight row i = gain * e_i): CPU and GPU then compute
    bit-identical logits whatever their summation order, so the top-k selection cannot flip on a near-tie
    between the two sides — the parity check below compares GEMM arithmetic, not a routing coin toss."""
    with torch.no_grad():
        w = ffn.router[0].weight
        w.zero_()
        for i in range(w.size(0)):
            w[i, i] = gain
        ffn.router[0].bias.zero_()


@pytest.mark.parametrize("bs", [1024, 2048])
def test_routed_ffn_bench_shape_matches_oracle(bs):
    """The shape bench.py times (d 2048, F 8192, block 1024 | 2048), reduced only in T (2048 tokens): K = 2048
    accumulation, 256x256 pair units.  Reference contract: test/layer/test_sparse_ffn.py:41-113."""
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    torch.manual_seed(bs)
    d, Fdim, T = 2048, 8192, 2048
    ffn = layers.RoutedFFN(d_model=d, d_feedforward=Fdim, block_size=bs, activation=torch.nn.ReLU()).to(DEV)
    _exact_router(ffn)
    with torch.no_grad():
        for p in ffn.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(4, T // 4, d)).to(DEV).requires_grad_()
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    xc = x.detach().cpu().requires_grad_()
    ps = {n: sd[n].clone().requires_grad_() for n in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")}
    y_ref = O.routed_ffn(xc, sd["router.0.weight"], sd["router.0.bias"], ps["fc1.weight"], ps["fc1.bias"],
                         ps["fc2.weight"], ps["fc2.bias"], bs, (Fdim // bs) // 2)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1e-2
    assert rel(x.grad, xc.grad) < 1e-2
    got = dict(ffn.named_parameters())
    for n, p in ps.items():
        assert rel(got[n].grad, p.grad) < 1e-2, (n, rel(got[n].grad, p.grad))


def test_routed_llama_ffn_llama7b_shape_matches_oracle():
    """LLaMA-7B shape (d 4096, F 11008, block 2752 = 21.5 x 128: ragged N / K tiles of the pair kernel), plain form
    (n_blocks // 4 = 1 active block, feedforward.py:156)."""
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    torch.manual_seed(7)
    d, Fdim, bs, T = 4096, 11008, 2752, 1024
    ffn = layers.RoutedLLaMaFFN(d_model=d, d_feedforward=Fdim, block_size=bs, activation=torch.nn.SiLU()).to(DEV)
    _exact_router(ffn)
    with torch.no_grad():
        for p in ffn.parameters():
            p.copy_(_bf(p))
    x = _bf(torch.randn(T, d)).to(DEV).requires_grad_()
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    xc = x.detach().cpu().requires_grad_()
    ps = {n: sd[n].clone().requires_grad_() for n in ("gate.weight", "side.weight", "down.weight")}
    y_ref = O.routed_llama_ffn(xc, sd["router.0.weight"], sd["router.0.bias"], ps["gate.weight"], ps["side.weight"],
                               ps["down.weight"], bs, (Fdim // bs) // 4)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1.5e-2
    assert rel(x.grad, xc.grad) < 1.5e-2
    got = dict(ffn.named_parameters())
    for n, p in ps.items():
        assert rel(got[n].grad, p.grad) < 1.5e-2, (n, rel(got[n].grad, p.grad))


def test_lora_routed_llama_ffn_llama7b_shape_matches_oracle():
    """configs[3]'s FFN: LoRARoutedLLaMaFFN at d 4096, F 11008, block 2752, rank 16, half the blocks active
    (tuning/lora_ffn.py:164-225); router, LoRA factors and x gradients vs the oracle."""
    from oracle import spt_oracle as O
    from spt_proto_b200 import layers
    d, Fdim, bs, r, T = 4096, 11008, 2752, 16, 1024
    ffn, x = _lora_setup(layers.LoRARoutedLLaMaFFN, torch.nn.SiLU(), d, Fdim, bs, r, T, 13)
    _exact_router(ffn, gain=1.0)
    y = ffn(x)
    dy = _bf(torch.randn_like(y))
    y.backward(dy)
    sd = {k: v.detach().cpu() for k, v in ffn.state_dict().items()}
    names = [n for n, p in ffn.named_parameters() if p.requires_grad]
    p = {n: sd[n].clone().requires_grad_() for n in names}
    xc = x.detach().cpu().requires_grad_()
    y_ref = O.lora_routed_llama_ffn(
        xc, p["router.0.weight"], p["router.0.bias"], sd["gate.weight"], sd["side.weight"], sd["down.weight"],
        p["gate.lora.left.weight"], p["gate.lora.right.weight"], p["side.lora.left.weight"],
        p["side.lora.right.weight"], p["down.lora.left.weight"], p["down.lora.right.weight"], bs, (Fdim // bs) // 2)
    y_ref.backward(dy.cpu())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(y, y_ref.detach()) < 1.5e-2
    assert rel(x.grad, xc.grad) < 2e-2
    got = dict(ffn.named_parameters())
    for n in names:
        assert rel(got[n].grad, p[n].grad) < 3e-2, (n, rel(got[n].grad, p[n].grad))


def test_route_bucket_nan_and_inf_rows():
    """torch.topk treats NaN as the greatest value and always returns exactly k indices; the bucketing must mark
    exactly k blocks for ANY row (a diverged bf16 step produces NaN / inf probabilities) — never more, which
    would write past token_rows / the static row capacity."""
    from spt_proto_b200 import ext
    T, nb, k = 300, 8, 4
    g = torch.Generator().manual_seed(1)
    prob = torch.sigmoid(torch.randn(T, nb, generator=g))
    prob[3] = float("nan")
    prob[5, 2] = float("nan")
    prob[7, :] = float("inf")
    prob[9, 1] = float("inf"); prob[9, 6] = float("nan"); prob[9, 0] = float("-inf")
    prob[T - 1] = float("nan")                                  # the last token: an overrun would leave the buffers
    b = ext.route_bucket(prob.to(DEV), k)
    torch.cuda.synchronize()
    rows = b.bucket_rows.cpu()
    assert int(rows.sum()) == T * k
    want = torch.topk(prob, k, dim=-1).indices
    ptr, row_token = b.bucket_ptr.cpu(), b.row_token.cpu()
    member = torch.zeros(T, nb, dtype=torch.bool)
    for gi in range(nb):
        toks = row_token[ptr[gi]: ptr[gi] + rows[gi]].long()
        assert (toks >= 0).all()
        member[toks, gi] = True
    assert (member.sum(1) == k).all()
    for t in (5, 9):                                            # unambiguous rows: same SET as torch.topk
        assert set(torch.nonzero(member[t]).flatten().tolist()) == set(want[t].tolist())
    assert member[3, :k].all() and member[7, :k].all() and member[T - 1, :k].all()   # all-equal rows: lowest indices
    tr = b.token_rows.cpu()
    assert ((tr >= 0) & (tr < b.R)).all()


def test_silu_mul_matches_torch_autograd():
    """The fused gated unit of the plain RoutedLLaMaFFN (act(gate) * side, feedforward.py:172-176) against torch in fp32."""
    from spt_proto_b200.kernels import ffn as F
    torch.manual_seed(5)
    g = (torch.randn(300, 512, device=DEV) * 2).bfloat16().requires_grad_()
    s = torch.randn(300, 512, device=DEV).bfloat16().requires_grad_()
    dh = torch.randn(300, 512, device=DEV).bfloat16()
    h = F.silu_mul(g, s)
    h.backward(dh)
    gf, sf = g.detach().float().requires_grad_(), s.detach().float().requires_grad_()
    hf = torch.nn.functional.silu(gf) * sf
    hf.backward(dh.float())
    assert torch.allclose(h.float(), hf, atol=2e-2, rtol=1e-2)
    assert torch.allclose(g.grad.float(), gf.grad, atol=2e-2, rtol=2e-2)
    assert torch.allclose(s.grad.float(), sf.grad, atol=2e-2, rtol=2e-2)


def test_row_coeff_matches_torch_index_chain():
    """layers.lora._row_coeff (2 * row_prob forward, spt_row_coeff_bwd backward) == the torch formulation it replaced:
    coeff[r] = 2 * prob[token(r), block(r)] * valid(r), differentiable in prob (lora_ffn.py:92,206), bit for bit."""
    from spt_proto_b200 import ext
    from spt_proto_b200.layers.lora import _row_coeff
    torch.manual_seed(11)
    for T, nb, k in ((1000, 8, 4), (77, 16, 4), (2048, 4, 2)):
        prob = torch.rand(T, nb, device=DEV).bfloat16().float()
        bucket = ext.route_bucket(prob.contiguous(), k)
        g = torch.randn(bucket.R, device=DEV)
        p1 = prob.clone().requires_grad_()
        c1 = _row_coeff(p1, bucket)
        c1.backward(g)
        p2 = prob.clone().requires_grad_()
        group = bucket.tile_group.clamp(min=0).long().repeat_interleave(128)
        valid = bucket.row_token >= 0
        flat = bucket.row_token.clamp(min=0).long() * nb + group
        c2 = 2.0 * p2.reshape(-1)[flat] * valid
        c2.backward(g)
        assert torch.equal(c1.detach(), c2.detach())
        assert torch.equal(p1.grad, p2.grad)
