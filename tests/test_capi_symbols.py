"""The C-ABI library loads on a GPU-less box and exports every symbol include/syntex_b200.h declares; argument
validation and the error convention work without a device (no compute is attempted here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "syntex_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("stx_vgg_create", "stx_plan_create", "stx_plan_step", "stx_plan_step_host", "stx_lbfgs_run",
                 "stx_conv_bf16", "stx_conv_bf16x3", "stx_gram_bf16", "stx_last_error"):
        assert must in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header(built_lib):
    from syntex_b200 import _lib
    _lib.lib()
    assert set(declared_symbols()) == set(_lib.EXPORTED.keys())


def test_error_convention_without_device(built_lib):
    from syntex_b200 import _lib
    L = _lib.lib()
    assert L.stx_abi_version() == 1
    h = ctypes.c_void_p()
    rc = L.stx_plan_create(ctypes.byref(h), None, 1, 64, 64, 15, 0, None, None, 0)
    assert rc == -1 and b"null" in L.stx_last_error()
    assert L.stx_debug_set(999, 0) == -1
    assert L.stx_gram_workspace_bytes(0, 0, 0) == 0
    assert L.stx_gram_workspace_bytes(1, 65536, 64) > 0
    with pytest.raises(_lib.SyntexLibraryError):
        _lib.check(rc)


def test_sass_contains_blackwell_tensor_and_tma_instructions(built_lib):
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # cp.async.bulk.tensor (TMA)
    assert "LDTM" in sass         # tcgen05.ld
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path
