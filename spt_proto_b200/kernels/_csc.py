"""CSR->CSC cache shared by the backward passes.  The transposed products dK = dS^T Q and
dV = P^T dO (reference kernels/sddmm.py:44-49, kernels/spmm.py:42-47) use the same sparsity
pattern, so the CSC is built once per (indptr, indices) pair and reused."""
from collections import OrderedDict

import torch

from .. import ext

_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_MAX = 4


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


def clear():
    _CACHE.clear()
