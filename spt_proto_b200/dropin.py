"""Register this package under the reference's module names, so that code written against
ytgui/SPT-proto (`from naive_gpt import ext, kernels, layers`) runs on the B200 kernels unchanged.

    import spt_proto_b200.dropin as dropin
    dropin.install()                 # sys.modules['naive_gpt'(.ext|.kernels|.layers|.utils)] -> spt_proto_b200.*
    from naive_gpt import kernels, layers

Only the hot-path surface exists: ext, kernels, layers and utils (the module upgrader).
naive_gpt.models / loaders are out of scope and raise AttributeError."""
import sys
import types


def install(name: str = "naive_gpt", force: bool = False) -> types.ModuleType:
    if name in sys.modules and not force and not getattr(sys.modules[name], "__spt_b200__", False):
        raise RuntimeError(f"{name} is already imported from somewhere else; pass force=True to shadow it")
    from . import ext, kernels, layers, utils

    pkg = types.ModuleType(name)
    pkg.__spt_b200__ = True
    pkg.__path__ = []  # mark as package
    pkg.ext, pkg.kernels, pkg.layers, pkg.utils = ext, kernels, layers, utils
    sys.modules[name] = pkg
    sys.modules[name + ".ext"] = ext
    sys.modules[name + ".kernels"] = kernels
    sys.modules[name + ".layers"] = layers
    sys.modules[name + ".utils"] = utils
    return pkg
