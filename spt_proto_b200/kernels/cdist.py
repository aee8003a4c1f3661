"""cdist autograd Function (reference naive_gpt/kernels/cdist.py:6-30)."""
import torch
from torch import autograd

from .. import ext


class CDist(autograd.Function):
    @staticmethod
    def forward(ctx, query: torch.Tensor, table: torch.Tensor):
        ctx.save_for_backward(query, table)
        distance, indices = ext.cdist_forward_cuda(query, table)
        ctx.mark_non_differentiable(indices)
        return distance, indices

    @staticmethod
    def backward(ctx, grad_distance: torch.Tensor, grad_indices: torch.Tensor):
        query, table = ctx.saved_tensors
        grad_query, grad_table = ext.cdist_backward_cuda(
            query.float().contiguous(), table.float().contiguous(), grad_distance.float().contiguous())
        return grad_query.to(query.dtype), grad_table.to(table.dtype)


def cdist(query: torch.Tensor, table: torch.Tensor):
    """query [m, n, dc], table [m, c, dc] -> (L1 distance [m, n, c], argmin codes [m, n] int32)."""
    return CDist.apply(query, table)
