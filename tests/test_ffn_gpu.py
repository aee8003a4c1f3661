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
    assert torch.isnan(got[R:]).all()                              # tail tiles (-1) are not touched
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
