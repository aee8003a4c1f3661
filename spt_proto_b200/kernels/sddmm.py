"""sddmm autograd Function (reference naive_gpt/kernels/sddmm.py:6-60).
forward : values = sample(Q K^T) on the CSR pattern
backward: dQ = spmm(dvalues, K);  dK = spmm^T(dvalues, Q) through the cached CSC (deterministic;
          the reference forks a side stream around cuSPARSE's transposed SpMM instead)."""
from torch import autograd

from .. import ext
from ._csc import get_csc


class SDDMM(autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, query, key):
        ctx.save_for_backward(indptr, indices, query, key)
        return ext.sddmm_forward_cuda(False, True, indptr, indices, query, key)

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, query, key = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_query = grad_key = None
        if ctx.needs_input_grad[2]:
            grad_query = ext.spmm_forward_cuda(False, False, indptr, indices, grad_output, key)
        if ctx.needs_input_grad[3]:
            grad_key = ext.spmm_csc(get_csc(indptr, indices), grad_output, query)
        return None, None, grad_query, grad_key


def sddmm(indptr, indices, query, key):
    return SDDMM.apply(indptr, indices, query, key)
