"""The C-ABI library loads without a GPU and exports every symbol include/spt_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_seven_reference_entry_points():
    syms = _declared_symbols()
    # replacements of extension/entry.cpp:43-56
    for s in ("spt_cdist_fwd", "spt_cdist_bwd", "spt_lookup_fwd", "spt_spmm_fwd", "spt_sddmm_fwd", "spt_softmax_fwd",
              "spt_softmax_bwd"):
        assert s in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "spt_proto_b200", "lib", "libspt_b200.so"))
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.spt_abi_version.restype = ctypes.c_int
    assert lib.spt_abi_version() == 1
    lib.spt_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.spt_last_error(), bytes)


def test_argument_validation_needs_no_gpu():
    """Null pointers / bad sizes are rejected before any CUDA call: status 1 + message."""
    lib = ctypes.CDLL(os.path.join(ROOT, "spt_proto_b200", "lib", "libspt_b200.so"))
    lib.spt_last_error.restype = ctypes.c_char_p
    rc = lib.spt_lookup_fwd(None, None, None, None, 1, 64, 8, 8, None)
    assert rc == 1 and b"null" in lib.spt_last_error()
    buf = (ctypes.c_int32 * 16)()
    rc = lib.spt_lookup_fwd(buf, buf, buf, None, 1, 64, 2, 8, None)        # m < 4
    assert rc == 1 and b"n_subspaces" in lib.spt_last_error()
    rc = lib.spt_lookup_fwd(buf, buf, buf, None, 1, 64, 8, 6, None)        # nnz % 4
    assert rc == 1 and b"multiple of 4" in lib.spt_last_error()


def test_sass_contains_blackwell_tensor_and_tma_ops():
    """cuobjdump evidence that the grouped GEMM is tcgen05/TMEM/TMA code (B200_PROFILING.md table)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", os.path.join(ROOT, "spt_proto_b200", "lib", "libspt_b200.so")],
                          capture_output=True, text=True).stdout
    for op in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA"):
        assert op in sass, op
