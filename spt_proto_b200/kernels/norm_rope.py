"""Autograd wrappers of the fused RMSNorm / RoPE kernels (csrc/norm_rope.cu): the two elementwise layers of a
LLaMA-style block around the SPT operators (reference naive_gpt/layers/basic/utils.py:22-38, position.py:5-48),
one kernel per direction instead of a chain of torch elementwise ops."""
from torch import autograd

from .. import ext


class RMSNormFn(autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        x = x.contiguous()
        out, inv = ext.rmsnorm_fwd(x, weight, eps)
        ctx.save_for_backward(x, weight, inv)
        return out

    @staticmethod
    def backward(ctx, grad):
        x, weight, inv = ctx.saved_tensors
        dx, dw = ext.rmsnorm_bwd(grad.contiguous(), x, weight, inv)
        return dx, (dw.to(weight.dtype) if ctx.needs_input_grad[1] else None), None


class RopeFn(autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        ctx.save_for_backward(cos, sin)
        return ext.rope(x.contiguous(), cos, sin, False)

    @staticmethod
    def backward(ctx, grad):
        cos, sin = ctx.saved_tensors
        return ext.rope(grad.contiguous(), cos, sin, True), None, None


def rmsnorm(x, weight, eps):
    return RMSNormFn.apply(x, weight, eps)


def rope(x, cos, sin):
    return RopeFn.apply(x, cos, sin)
