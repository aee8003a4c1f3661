"""CSR->CSC cache shared by the backward passes.  The transposed products dK = dS^T Q and
dV = P^T dO (reference kernels/sddmm.py:44-49, kernels/spmm.py:42-47) use the same sparsity
pattern, so the CSC is built once per (indptr, indices) pair and reused."""
from collections import OrderedDict

import torch

from .. import ext

import os

_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_MAX = 4
_TILES_DIRECT = os.environ.get("SPT_SPMM_TILES_DIRECT", "1") != "0"    # A/B switch: the CSR-direction product on the tile index
_TILES_SDDMM = os.environ.get("SPT_SDDMM_TILES", "1") != "0"           # A/B switch: sddmm on the tile index


def get_csc(indptr: torch.Tensor, indices: torch.Tensor):
    key = (indptr.data_ptr(), indices.data_ptr(), indices._version, tuple(indices.shape),
           indices.device.index, torch.cuda.current_stream(indices.device).cuda_stream)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is indices:
        _CACHE.move_to_end(key)
        return hit[1]
    csc = ext.csr2csc(indptr, indices)
    _CACHE[key] = (indices, csc)  # holding `indices` keeps its storage (and data_ptr) alive
    while len(_CACHE) > _MAX:
        _CACHE.popitem(last=False)
    return csc


def get_tiles(indptr: torch.Tensor, indices: torch.Tensor):
    """The cached tile index of the pattern (entries bucketed by 64-column tile and 64-row chunk: 0.2 ms to build at the
    bench shape against 1.13 ms for the CSC); one index serves both directions of the product."""
    key = ("tiles", indptr.data_ptr(), indices.data_ptr(), indices._version, tuple(indices.shape),
           indices.device.index, torch.cuda.current_stream(indices.device).cuda_stream)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is indices:
        _CACHE.move_to_end(key)
        return hit[1]
    tiles = ext.csr_tiles(indptr, indices)
    _CACHE[key] = (indices, tiles)
    while len(_CACHE) > _MAX:
        _CACHE.popitem(last=False)
    return tiles


def transposed_product(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y = A^T x for the backward passes.  bf16 x with head dim 64 / 128 runs on the tile index; everything else goes
    through the cached CSC."""
    x = x.contiguous()
    if ext.csr_tiles_supported(indices, x):
        return ext.spmm_tiles(get_tiles(indptr, indices), values, x, trans=True)
    return ext.spmm_csc(get_csc(indptr, indices), values, x)


def direct_product(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y = A x (forward of spmm, dQ of sddmm): on the tile index when it applies, else the gathered CSR kernel."""
    x = x.contiguous()
    if _TILES_DIRECT and ext.csr_tiles_supported(indices, x):
        return ext.spmm_tiles(get_tiles(indptr, indices), values, x, trans=False)
    return ext.spmm_forward_cuda(False, False, indptr, indices, values, x)


def sddmm_product(indptr: torch.Tensor, indices: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                  scale: float = 1.0, clamp: float = 0.0) -> torch.Tensor:
    """values = clamp(scale * sddmm(q, k)): dense score tiles on the tile index when it applies, else the gathered kernel."""
    query, key = query.contiguous(), key.contiguous()
    if _TILES_SDDMM and key.dtype == query.dtype and ext.csr_tiles_supported(indices, query):
        return ext.sddmm_tiles(get_tiles(indptr, indices), query, key, scale, clamp)
    return ext.sddmm_scaled(indptr, indices, query, key, scale, clamp)


def clear():
    _CACHE.clear()
