"""sddmm autograd Function (reference naive_gpt/kernels/sddmm.py:6-60).
forward : values = sample(Q K^T) on the CSR pattern
backward: dQ = spmm(dvalues, K);  dK = spmm^T(dvalues, Q) through the cached CSC (deterministic;
          the reference forks a side stream around cuSPARSE's transposed SpMM instead)."""
from torch import autograd

from .. import ext
from ._csc import direct_product, sddmm_product, transposed_product


class SDDMM(autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, query, key):
        ctx.save_for_backward(indptr, indices, query, key)
        return sddmm_product(indptr, indices, query, key)

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, query, key = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_query = grad_key = None
        if ctx.needs_input_grad[2]:
            grad_query = direct_product(indptr, indices, grad_output, key)
        if ctx.needs_input_grad[3]:
            grad_key = transposed_product(indptr, indices, grad_output, query)
        return None, None, grad_query, grad_key


def sddmm(indptr, indices, query, key):
    return SDDMM.apply(indptr, indices, query, key)


class SDDMMScaled(autograd.Function):
    """clamp(scale * sddmm(q, k), -clamp, clamp) in one kernel (the reference's eager `clamp_(scaling * values)`,
    layers/sparse/attention.py:125-127); the backward applies the clamp's zero-gradient mask and the scale in one pass
    before the two products of SDDMM.backward."""

    @staticmethod
    def forward(ctx, indptr, indices, query, key, scale, clamp):
        values = sddmm_product(indptr, indices, query, key, scale, clamp)
        ctx.save_for_backward(indptr, indices, query, key, values)
        ctx.scale, ctx.clamp = float(scale), float(clamp)
        return values

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, query, key, values = ctx.saved_tensors
        grad_raw = ext.clamp_scale_bwd(grad_output.contiguous(), values, ctx.scale, ctx.clamp)
        grad_query = grad_key = None
        if ctx.needs_input_grad[2]:
            grad_query = direct_product(indptr, indices, grad_raw, key)
        if ctx.needs_input_grad[3]:
            grad_key = transposed_product(indptr, indices, grad_raw, query)
        return None, None, grad_query, grad_key, None, None


def sddmm_scaled(indptr, indices, query, key, scale: float, clamp: float):
    return SDDMMScaled.apply(indptr, indices, query, key, scale, clamp)


class SDDMMSoftmax(autograd.Function):
    """softmax(clamp(scale * sddmm(q, k), -clamp, clamp)) on the CSR pattern: the chain of `_get_attn`
    (layers/sparse/attention.py:122-130) as one Function.  Same two forward kernels as sddmm_scaled + softmax; the backward
    goes from the gradient of the probabilities to the gradient of the raw scores in ONE pass over the nnz arrays
    (softmax backward, clamp mask and scale: bit-identical to the two-kernel chain, 0.30 -> 0.21 ms per 128 heads)."""

    @staticmethod
    def forward(ctx, indptr, indices, query, key, scale, clamp):
        values = sddmm_product(indptr, indices, query, key, scale, clamp)
        probs = ext.softmax_forward_cuda(indptr, indices, values)
        ctx.save_for_backward(indptr, indices, query, key, values, probs)
        ctx.scale, ctx.clamp = float(scale), float(clamp)
        return probs

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, query, key, values, probs = ctx.saved_tensors
        grad_raw = ext.softmax_clamp_bwd(indptr, indices, probs, grad_output.contiguous(), values, ctx.scale, ctx.clamp)
        grad_query = grad_key = None
        if ctx.needs_input_grad[2]:
            grad_query = direct_product(indptr, indices, grad_raw, key)
        if ctx.needs_input_grad[3]:
            grad_key = transposed_product(indptr, indices, grad_raw, query)
        return None, None, grad_query, grad_key, None, None


def sddmm_softmax(indptr, indices, query, key, scale: float, clamp: float):
    return SDDMMSoftmax.apply(indptr, indices, query, key, scale, clamp)
