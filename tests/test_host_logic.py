"""Host-side logic that needs no GPU: layer constructors / state-dict layout / from_pretrained match
the reference's names, the drop-in registration, the loud failure without CUDA, sharding helpers."""
import sys

import pytest
import torch
from torch import nn


def test_layer_constructors_and_state_dict_keys():
    from spt_proto_b200 import layers
    v1 = layers.SparseVanillaAttentionV1(d_head=64, p_dropout=0.0, d_codeword=8, n_codewords=16, n_subspaces=8)
    assert set(v1.state_dict()) == {"trigger", "quantizer.weight"}
    assert v1.quantizer.weight.shape == (8, 16, 8)
    v2 = layers.SparseVanillaAttentionV2.from_pretrained(v1)
    assert torch.equal(v2.quantizer.weight, v1.quantizer.weight) and v2.sparse_coeff == 8
    r1 = layers.SparseRotaryAttentionV1(d_head=64, p_dropout=0.0, d_codeword=8, n_codewords=16, n_subspaces=8)
    r2 = layers.SparseRotaryAttentionV2.from_pretrained(r1)
    assert {"trigger", "quantizer.weight", "cached_ids", "embedding.cos_cached", "embedding.sin_cached"} <= set(r2.state_dict())
    with pytest.raises(AssertionError):
        layers.SparseVanillaAttentionV2.from_pretrained(r1)

    ff = layers.Feedforward(32, 128, 0.0, nn.ReLU())
    routed = layers.RoutedFFN.from_pretrained(16, ff)
    assert set(routed.state_dict()) == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "router.0.weight", "router.0.bias"}
    assert routed.n_blocks == 8 and routed.k_active == 4
    lf = layers.LLaMaFeedforward(32, 128, nn.SiLU())
    lrouted = layers.RoutedLLaMaFFN.from_pretrained(16, lf)
    assert lrouted.k_active == 2 and "router.0.weight" in lrouted.state_dict()


def test_v1_layers_are_dense_attention_plus_pq_loss():
    from spt_proto_b200 import layers
    torch.manual_seed(0)
    v1 = layers.SparseVanillaAttentionV1(d_head=32, p_dropout=0.0, d_codeword=8, n_codewords=16, n_subspaces=4)
    dense = layers.VanillaAttention(d_head=32, p_dropout=0.0)
    q, k, v = (torch.randn(2, 16, 3, 32) for _ in range(3))
    mask = torch.full([16, 16], float("-inf")).triu(1)
    assert torch.allclose(v1(q, k, v, attn_mask=mask), dense(q, k, v, attn_mask=mask))
    assert v1.loss.dim() == 0 and v1.loss.item() > 0


def test_pq_modes_v1():
    from spt_proto_b200 import layers
    torch.manual_seed(1)
    pq = layers.PQV1(d_codeword=4, n_codewords=8, n_subspaces=6)
    z = torch.randn(5, 7, 24)
    codes = pq("encode", z=z)
    assert codes.shape == (5, 7, 6)
    zq = pq("decode", z=codes)
    assert zq.shape == z.shape and torch.allclose(zq, pq("quantize", z=z))
    z_q, loss = pq("train", z=z)
    assert torch.allclose(z_q, zq) and loss.requires_grad


def test_no_cpu_fallback():
    from spt_proto_b200 import ext, layers
    q = torch.zeros(8, 64, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.cdist_forward_cuda(q, torch.zeros(8, 16, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.lookup_forward_cuda(torch.empty([8]), torch.zeros(1, 64, 8, dtype=torch.int32), torch.zeros(1, 64, 8, dtype=torch.int32))
    ffn = layers.RoutedFFN(32, 128, 16, nn.ReLU())
    with pytest.raises(RuntimeError, match="no CPU path"):
        ffn(torch.zeros(2, 4, 32))
    attn = layers.SparseVanillaAttentionV2(d_head=64, d_codeword=8, n_codewords=16, p_dropout=0.0)
    with pytest.raises(RuntimeError):
        attn(torch.zeros(1, 64, 2, 64), torch.zeros(1, 64, 2, 64), torch.zeros(1, 64, 2, 64))


def test_product_never_imports_the_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spt_proto_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "spt_oracle" not in text.replace(
                    "oracle/spt_oracle", ""), f


def test_dropin_install():
    import spt_proto_b200.dropin as dropin
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "naive_gpt" or k.startswith("naive_gpt.")}
    try:
        dropin.install()
        from naive_gpt import ext, kernels, layers  # noqa: F401
        assert {"cdist", "lookup", "softmax", "sddmm", "spmm"} <= set(dir(kernels))
        for name in ("cdist_forward_cuda", "cdist_backward_cuda", "lookup_forward_cuda", "spmm_forward_cuda",
                     "sddmm_forward_cuda", "softmax_forward_cuda", "softmax_backward_cuda"):
            assert callable(getattr(ext, name))           # extension/entry.cpp:43-56
        assert layers.SparseVanillaAttentionV2 and layers.RoutedFFN
    finally:
        for k in list(sys.modules):
            if k == "naive_gpt" or k.startswith("naive_gpt."):
                del sys.modules[k]
        sys.modules.update(saved)


def test_shard_range_partitions_exactly():
    from spt_proto_b200.distributed import shard_range
    for n in (0, 1, 7, 32, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _small_block(kind):
    from torch import nn
    from spt_proto_b200 import layers
    torch.manual_seed(7)
    if kind == "llama":
        return layers.TransformerBlock(
            d_model=128, n_heads=2, layernorm_fn=layers.LlamaRMSNorm(128),
            attention_fn=layers.RotaryAttention(d_head=64, p_dropout=0.0),
            feedforward_fn=layers.LLaMaFeedforward(d_model=128, d_feedforward=512, activation=nn.SiLU()),
            attention_bias=False, pre_norm=True)
    return layers.TransformerBlock(
        d_model=128, n_heads=2, layernorm_fn=nn.LayerNorm(128),
        attention_fn=layers.VanillaAttention(d_head=64, p_dropout=0.0),
        feedforward_fn=layers.Feedforward(d_model=128, d_feedforward=512, p_dropout=0.0, activation=nn.ReLU()),
        attention_bias=True, pre_norm=True)


@pytest.mark.parametrize("kind", ["llama", "opt"])
def test_module_upgrader_matches_reference_upgrade(kind):
    """The four-pass sparse upgrade (reference utils/adapter.py) yields the same module classes, the same
    trainable set and the state_dict layout (keys, shapes, dtypes) of the reference's upgraded model
    (tests/golden/upgrader_state.pt is produced by the unmodified reference, make_upgrader_golden.py)."""
    import os
    from spt_proto_b200 import utils
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "upgrader_state.pt"))[kind]
    model = _small_block(kind)
    for stage in ("lora", "ffn", "mha_v1", "mha_v2"):
        model = utils.ModuleUpgrader(utils.SparseLoRAHandler(d_lora=4, stage=stage, verbose=False)).visit(model)
    assert {n: type(m).__name__ for n, m in model.named_modules()} == gold["classes"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == gold["trainable"]
    ours = model.state_dict()
    assert set(ours) == set(gold["state_dict"])
    assert all(tuple(ours[k].shape) == shape and str(ours[k].dtype) == dtype for k, (shape, dtype) in gold["state_dict"].items())
    ckpt = {k: torch.zeros(shape, dtype=getattr(torch, dtype.split(".")[1])) for k, (shape, dtype) in gold["state_dict"].items()}
    model.load_state_dict(ckpt, strict=True)      # a checkpoint with the reference's layout loads strictly


def test_module_upgrader_requires_default_and_skips_unknown():
    from spt_proto_b200 import utils
    with pytest.raises(RuntimeError):
        utils.ModuleUpgrader(object())
    lin = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.ReLU())
    out = utils.ModuleUpgrader(utils.LoRAHandler(d_lora=2, verbose=False)).visit(lin)
    assert type(out[0]).__name__ == "LoRALinear" and type(out[1]).__name__ == "ReLU"


# ------------------------------------------------------------------ grouped GEMM on CTA pairs: the unit schedule
def _pair_plan(tile_group, tiles_n):
    """Replay of the device schedule through the C ABI (host code shared with the kernel: ffn_gemm.cu get_unit)."""
    import ctypes

    import numpy as np
    from spt_proto_b200._lib import lib
    tg = np.asarray(tile_group, dtype=np.int32)
    n = lib.spt_grouped_gemm_plan(tg.ctypes.data_as(ctypes.c_void_p), len(tg), tiles_n, None, 0)
    assert n >= 0
    out = np.zeros((n, 6), dtype=np.int32)
    assert lib.spt_grouped_gemm_plan(tg.ctypes.data_as(ctypes.c_void_p), len(tg), tiles_n,
                                     out.ctypes.data_as(ctypes.c_void_p), n) == n
    return out


@pytest.mark.parametrize("sizes,tails,tiles_n", [
    ([1], 0, 1), ([2], 0, 3), ([3, 0, 1, 5], 2, 2), ([1, 1, 1, 1, 1], 1, 1), ([2, 2, 3], 0, 4), ([0, 0, 4], 3, 1),
    ([5, 1, 0, 0, 7, 2], 4, 2), ([16, 15, 17], 0, 8)])
def test_pair_gemm_schedule_covers_every_tile_once(sizes, tails, tiles_n):
    """Every (m-tile, n-tile) of a group is computed by exactly one active CTA on its own group's weights, every
    tail tile is zero-filled exactly once, and a pair that straddles two groups costs one extra unit."""
    tile_group = [g for g, n in enumerate(sizes) for _ in range(n)] + [-1] * tails
    if not tile_group:
        pytest.skip("no tiles")
    plan = _pair_plan(tile_group, tiles_n)
    tiles_m = len(tile_group)
    active, zero = {}, {}
    units = {}
    for unit, rank, g, mt, nt, flags in plan.tolist():
        role, mma = flags & 15, flags >> 4
        units.setdefault(unit, []).append((rank, g, mt, nt, role, mma))
        assert mt // 2 == units[unit][0][2] // 2 and nt == units[unit][0][3]      # the two CTAs of a unit: one m-pair, one n-tile
        if role == 1:
            assert mma == 1 and 0 <= mt < tiles_m and tile_group[mt] == g >= 0      # computed on its own group's weights
            active[(mt, nt)] = active.get((mt, nt), 0) + 1
        elif role == 2:
            assert 0 <= mt < tiles_m and tile_group[mt] == -1
            zero[(mt, nt)] = zero.get((mt, nt), 0) + 1
    for u, recs in units.items():
        assert sorted(r[0] for r in recs) == [0, 1]
        assert recs[0][5] == recs[1][5] and recs[0][1] == recs[1][1]               # both CTAs agree on mma and group
    for mt, g in enumerate(tile_group):
        for nt in range(tiles_n):
            if g >= 0:
                assert active.get((mt, nt)) == 1 and (mt, nt) not in zero
            else:
                assert zero.get((mt, nt)) == 1 and (mt, nt) not in active
    assert sum(active.values()) == sum(1 for g in tile_group if g >= 0) * tiles_n
    # cost: one unit per pair, plus one where a pair's two tiles belong to different (valid) groups
    straddle = sum(1 for t in range(0, tiles_m - 1, 2) if tile_group[t + 1] >= 0 and tile_group[t + 1] != tile_group[t])
    assert len(units) == ((tiles_m + 1) // 2 + straddle) * tiles_n


@pytest.mark.parametrize("tile_group", [[-1, 0, 0, 1], [0, -1, 1, 1, -1, 2], [-1], [-1, -1, 3], [2, 2, -1, -1, 0]])
def test_pair_gemm_schedule_with_unused_tiles_anywhere(tile_group):
    """The ABI allows -1 (unused) for any 128-row tile, not only in the tail: valid neighbours are still computed once."""
    tiles_n = 2
    plan = _pair_plan(tile_group, tiles_n)
    active, zero = {}, {}
    for unit, rank, g, mt, nt, flags in plan.tolist():
        role = flags & 15
        if role == 1:
            assert tile_group[mt] == g >= 0
            active[(mt, nt)] = active.get((mt, nt), 0) + 1
        elif role == 2:
            assert tile_group[mt] == -1
            zero[(mt, nt)] = zero.get((mt, nt), 0) + 1
    for mt, g in enumerate(tile_group):
        for nt in range(tiles_n):
            assert (active if g >= 0 else zero).get((mt, nt)) == 1
            assert (mt, nt) not in (zero if g >= 0 else active)


def test_pair_gemm_plan_rejects_bad_arguments():
    import ctypes

    import numpy as np
    from spt_proto_b200._lib import lib
    tg = np.zeros(4, dtype=np.int32)
    assert lib.spt_grouped_gemm_plan(None, 4, 1, None, 0) < 0
    assert lib.spt_grouped_gemm_plan(tg.ctypes.data_as(ctypes.c_void_p), 0, 1, None, 0) < 0
    assert lib.spt_grouped_gemm_plan(tg.ctypes.data_as(ctypes.c_void_p), 4, 0, None, 0) < 0
    assert lib.spt_grouped_gemm_plan(tg.ctypes.data_as(ctypes.c_void_p), 2000, 1, None, 0) < 0


def test_layout_helpers_fall_back_to_torch_on_cpu():
    """kernels.layout.swap12 / transpose_last2 use the CUDA copy kernels only where they apply; anything else (CPU tensors
    in the host-logic tests, odd row sizes) takes torch's own transpose(1, 2).contiguous() with the same result."""
    from spt_proto_b200 import ext
    from spt_proto_b200.kernels import layout
    x = torch.randn(2, 5, 3, 8, requires_grad=True)
    assert not ext.swap12_supported(x)
    y = layout.swap12(x)
    assert y.is_contiguous() and torch.equal(y, x.transpose(1, 2).contiguous())
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    z = torch.randn(3, 7, 5)
    assert not ext.transpose_last2_supported(z)
    assert torch.equal(layout.transpose_last2(z), z.transpose(1, 2).contiguous())
    with pytest.raises(RuntimeError):
        ext.swap12(x.detach())          # the ext entry points themselves never fall back


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the oracle port on the host cores; no GPU, no /root/reference at run time) prints one
    JSON line with the driver's keys, honours --steps / --warmup, and reports the same metric / unit as the GPU arm."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-heads", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sparse_mha_fwd_bwd_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
