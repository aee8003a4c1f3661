"""spmm autograd Function (reference naive_gpt/kernels/spmm.py:6-58).
forward : y = A x
backward: dA = sddmm(dy, x) on the pattern;  dx = A^T dy through the cached CSC."""
from torch import autograd

from .. import ext
from ._csc import direct_product, sddmm_product, transposed_product


class SPMM(autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, values, x):
        ctx.save_for_backward(indptr, indices, values, x)
        return direct_product(indptr, indices, values, x)

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, values, x = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_a = grad_x = None
        if ctx.needs_input_grad[2]:
            grad_a = sddmm_product(indptr, indices, grad_output, x)
        if ctx.needs_input_grad[3]:
            grad_x = transposed_product(indptr, indices, values, grad_output)
        return None, None, grad_a, grad_x


def spmm(indptr, indices, values, x):
    return SPMM.apply(indptr, indices, values, x)
