"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/tagan_b200.h
declares (no compute calls -- there is no GPU here), and the host layer refuses CPU tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "tagan_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tagan_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from tagan_b200 import _lib, build
    build.build(verbose=False)
    lib = _lib.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tagan_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES out of sync with the header"
    assert lib.tagan_abi_version() >= 1


def test_workspace_queries_run_without_gpu():
    from tagan_b200 import _lib
    lib = _lib.load()
    assert lib.tagan_csr_workspace_bytes(1000, 100) > 0
    assert lib.tagan_layernorm_bwd_workspace_bytes(1000, 128) > 0
    assert lib.tagan_gemm_workspace_bytes(2, 128, 384, 100000) > 0


def test_no_cpu_fallback():
    import tagan_b200
    layer = tagan_b200.TAGANGraphAttention(32, 2, dropout=0.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.randn(4, 32), torch.randint(0, 4, (2, 6)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tagan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "ref_loader" not in src and "/root/reference" not in src, f
