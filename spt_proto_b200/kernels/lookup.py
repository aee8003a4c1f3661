"""lookup Function (reference naive_gpt/kernels/lookup.py:6-24): integer output, no gradient."""
import torch
from torch import autograd

from .. import ext


class Lookup(autograd.Function):
    @staticmethod
    def forward(ctx, config: torch.Tensor, query: torch.Tensor, key: torch.Tensor):
        return ext.lookup_forward_cuda(config, query, key)

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        raise NotImplementedError


def lookup(query: torch.Tensor, key: torch.Tensor, sparse_coeff: int):
    """query, key [B, S, m] int32 PQ codes -> top-k candidate keys per query [B, S, S // sparse_coeff]."""
    # the reference smuggles sparse_coeff through the SHAPE of a CPU tensor (lookup.py:23); keep that
    config = torch.empty([sparse_coeff], device="meta")
    return Lookup.apply(config, query, key)
