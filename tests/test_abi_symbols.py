"""The C-ABI library loads on a CPU-only box, exports every symbol include/apo_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "apo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apo_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(apo):
    assert header_symbols() == sorted(apo.ABI_SYMBOLS)


def test_library_exports_every_symbol(apo):
    apo.build_library()
    lib = ctypes.CDLL(apo.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert lib.apo_abi_version() == 1


def test_struct_sizes_match_header(apo):
    from importlib import import_module
    eng = import_module("senweaver-ide_b200.engine")
    assert ctypes.sizeof(eng.Pattern) == 40
    assert ctypes.sizeof(eng.DimStat) == 32
    assert ctypes.sizeof(eng.ScoreOpts) == 32
    assert ctypes.sizeof(eng.Timing) == 28
    assert eng.RECORD_DTYPE.itemsize == 32


def test_no_cpu_fallback(apo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    apo.build_library()
    with pytest.raises(apo.ApoError) as ei:
        apo.Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "senweaver-ide_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "apo_oracle" not in txt, f
