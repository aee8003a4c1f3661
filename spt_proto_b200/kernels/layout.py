"""Layout copies of the stage path as autograd Functions: the reference layer's `transpose(1, 2).contiguous()`
(naive_gpt/layers/sparse/attention.py:92-95, 138-142) through the 16-byte-word copy kernels of csrc/layout.cu instead
of torch's element-wise strided copy.  Shapes the kernels do not cover raise in ext (callers check *_supported)."""
import torch
from torch import autograd

from .. import ext


class Swap12(autograd.Function):
    """y = x.transpose(1, 2).contiguous() for a contiguous 4-D x; the backward is the same move the other way."""

    @staticmethod
    def forward(ctx, x):
        return ext.swap12(x)

    @staticmethod
    def backward(ctx, grad):
        g = grad.contiguous()
        return ext.swap12(g) if ext.swap12_supported(g) else g.transpose(1, 2).contiguous()


class TransposeLast2(autograd.Function):
    """y = x.transpose(1, 2).contiguous() for a contiguous 3-D x (a real 2-D transpose per batch entry)."""

    @staticmethod
    def forward(ctx, x):
        return ext.transpose_last2(x)

    @staticmethod
    def backward(ctx, grad):
        g = grad.contiguous()
        return ext.transpose_last2(g) if ext.transpose_last2_supported(g) else g.transpose(1, 2).contiguous()


def swap12(x: torch.Tensor) -> torch.Tensor:
    """transpose(1, 2).contiguous() of a 4-D tensor; torch's own copy where the kernel does not apply."""
    if ext.swap12_supported(x):
        return Swap12.apply(x)
    return x.transpose(1, 2).contiguous()


def transpose_last2(x: torch.Tensor) -> torch.Tensor:
    if ext.transpose_last2_supported(x):
        return TransposeLast2.apply(x)
    return x.transpose(1, 2).contiguous()
