"""CPU: the C-ABI library loads and exports every symbol include/rbvfit_b200.h declares; without a GPU it
fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    text = open(os.path.join(ROOT, "include", "rbvfit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from rbvfit_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbvfit_b200.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in rbvfit_b200/_lib.py"
    assert lib.rbv_version().startswith(b"rbvfit_b200")


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rbvfit_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    st = lib.rbv_create(0, C.byref(h))
    assert st == 2   # RBV_ECUDA
    assert b"no CPU fallback" in lib.rbv_last_error()
    from rbvfit_b200.engine import Engine
    with pytest.raises(_lib.RbvError):
        Engine()


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "rbvfit_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                if f.endswith(".py"):      # no CPU special-function path inside the product
                    assert not re.search(r"^\s*(from|import)\s+scipy\.special", text, flags=re.M), f
