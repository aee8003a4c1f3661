"""CSR softmax autograd Function (reference naive_gpt/kernels/softmax.py:6-38)."""
from torch import autograd

from .. import ext


class Softmax(autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, values):
        output = ext.softmax_forward_cuda(indptr, indices, values)
        ctx.save_for_backward(indptr, indices, output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        indptr, indices, output = ctx.saved_tensors
        return None, None, ext.softmax_backward_cuda(indptr, indices, output, grad_output.contiguous())


def softmax(indptr, indices, values):
    return Softmax.apply(indptr, indices, values)
