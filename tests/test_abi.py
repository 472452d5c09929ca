"""The C-ABI library loads and exports every symbol include/dc_cuda.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import data_compressor_b200 as d
    d.build()
    return d.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dc_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from data_compressor_b200 import _lib
    decl = declared_functions()
    assert decl == sorted(_lib.SYMBOLS), "include/dc_cuda.h and data_compressor_b200/_lib.py disagree"
    for name in decl:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    assert lib.dc_abi_version() == 1


def test_layout_stream_count(lib):
    from data_compressor_b200 import Layout
    assert lib.dc_layout_streams(C.byref(Layout(4096, 86400, 1024, 0, None))) == 4096 * 85
    assert lib.dc_layout_streams(C.byref(Layout(4096, 86400, 0, 0, None))) == 4096
    assert lib.dc_layout_streams(C.byref(Layout(3, 0, 16, 0, None))) == 3
    assert lib.dc_layout_streams(C.byref(Layout(0, 0, 0, 7, C.c_void_p(8)))) == 7


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.dc_create(C.byref(h), 0) == -10       # ERROR_LIBRARY_INIT
    from data_compressor_b200 import Context, DcError
    with pytest.raises(DcError):
        Context(0)


def test_product_does_not_import_the_oracle():
    """The product package must never route through oracle/ (checked textually)."""
    pkg = os.path.join(ROOT, "data_compressor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dc_oracle" not in text and "oracle.binding" not in text and "libdcref" not in text, f
