"""The C-ABI library loads on a GPU-less box and exports every symbol include/opsagent_b200.h declares;
no compute is attempted without a GPU, and the product fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from opsagent_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "opsagent_b200.h")).read()
    return sorted(set(re.findall(r"OA_API [^;(]*?\b(oa_[a-z_0-9]+)\(", h)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/opsagent_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.load()
    h = C.c_void_p()
    rc = L.oa_engine_create(b'{"model": "llama-3-8b"}', C.byref(h))
    assert rc == 500 and not h
    assert "no CPU fallback" in _lib.last_error()


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu baseline may touch oracle/"""
    pkg = os.path.join(ROOT, "opsagent_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh")):
                src = open(os.path.join(dp, f), errors="replace").read()
                for line in src.splitlines():
                    code = line.split("//")[0].split("#")[0]
                    assert "import oracle" not in code and "from oracle" not in code and "liboracle" not in code, (f, line)
