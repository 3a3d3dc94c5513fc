"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/vinsat_b200.h declares,
and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from vinsat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vinsat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vinsat_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    lib = _lib.load()
    assert lib.vinsat_abi_version() == 1


def test_every_header_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 36
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "missing export: " + s
        assert s in _lib.SIGNATURES, "no ctypes signature for " + s
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    if lib.vinsat_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.VinsatError, match="no CUDA device|no CPU fallback"):
        _lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vinsat_b200")
    pat = re.compile(r"^\s*(from|import)\s+(oracle|ba_oracle|satcam_oracle|ref_loader)\b", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not pat.search(txt), f
                assert "oracle" not in txt.replace("no oracle", ""), f
