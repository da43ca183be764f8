"""CPU tests of the drop-in boundary: libeaglegpu.so loads, exports every symbol that
include/eagle_gpu.h declares, and refuses to compute without a GPU (no fallback)."""
import os
import re

import pytest

from eagleeverything_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "eagle_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[A-Za-z0-9_]+)\s*\(", src)) - {"eg_message_fn"})


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libeaglegpu.so does not export {s}"
    # and the ctypes table covers the header exactly
    assert sorted(_lib.SIGNATURES) == syms


def test_reference_facing_names_match_the_rcpp_exports():
    # RcppExports.cpp:154-170 registers these on the hot path; the C ABI carries one entry point each
    for name in ["ReadBlock", "calculateMMt_rcpp", "calculate_a_and_vara_rcpp", "calculate_reduced_a_rcpp",
                 "extract_geno_rcpp"]:
        assert "eg_" + name in _lib.SIGNATURES
        assert callable(getattr(api, name))


def test_abi_version_and_error_text():
    lib = _lib.load()
    assert lib.eg_abi_version() == 1
    assert isinstance(lib.eg_last_error(), bytes)


def test_no_gpu_means_loud_failure(tmp_path):
    lib = _lib.load()
    if lib.eg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.EagleGpuError, match="no CPU fallback"):
        api.calculateMMt_rcpp(str(tmp_path / "M.ascii"), 8, 1, [api.NA_REAL], (3, 3))
    assert lib.eg_init(0) == _lib.EG_ERR_CUDA
    assert b"no CPU fallback" in lib.eg_last_error()


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "eagleeverything_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
