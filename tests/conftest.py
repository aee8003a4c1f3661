import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_ext():
    """The UNMODIFIED reference CUDA extension compiled by oracle/build_ref.sh (GPU-side oracle for
    fp32, S in {256,512,1024}); None when it was not built."""
    import importlib.util

    path = os.path.join(ROOT, "oracle", "_ref", "ext_ref.so")
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location("ext_ref", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception as exc:  # pragma: no cover - depends on the box
        print("ext_ref.so present but not loadable:", exc)
        return None
