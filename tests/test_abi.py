"""The C-ABI library loads without a GPU and exports every entry point include/shrimp_b200.h declares
(no compute calls here).  Without a usable device the entry points fail loudly -- there is no CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "shrimp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shrimp_gpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from shrimp_b200._lib import lib
    L = lib()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_product_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "shrimp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shrimp_b200._lib import ShrimpGpuError
    import shrimp_b200
    with pytest.raises(ShrimpGpuError):
        shrimp_b200.GpuContext(0)
