"""spt_proto_b200 — B200 (sm_100a) implementation of the SPT hot path (ytgui/SPT-proto):
PQ sparse multi-head attention and the routed-FFN grouped GEMM, behind the reference's own
`naive_gpt.ext` / `naive_gpt.kernels` / `naive_gpt.layers` Python API.

Sub-modules (imported lazily so that `python -m spt_proto_b200.build` works before the CUDA
library exists):
    ext      the 7 reference extension entry points (+ new ones) over the C ABI of include/spt_b200.h
    kernels  torch.autograd Functions: cdist, lookup, sddmm, softmax, spmm (naive_gpt/kernels/*.py)
    layers   PQV2, Sparse{Vanilla,Rotary}AttentionV2, RoutedFFN, ... (naive_gpt/layers/*)
    utils    LoRAHandler / SparseLoRAHandler / ModuleUpgrader (naive_gpt/utils/adapter.py)
    dropin   install() registers this package as `naive_gpt` in sys.modules
"""
import importlib

__all__ = ["ext", "kernels", "layers", "utils", "dropin", "build"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
