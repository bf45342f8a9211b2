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


def test_stack_rows_is_a_view_of_sliced_buffers():
    """ops.stack_rows (host logic): consecutive slices of one allocation stack for free; anything else falls back to
    torch.stack with identical values."""
    import torch
    from tagan_b200 import ops
    buf = torch.arange(4 * 3 * 5, dtype=torch.float32).reshape(4, 3, 5)
    xs = list(buf.unbind(0))
    out = ops.stack_rows(xs)
    assert out.data_ptr() == buf.data_ptr() and torch.equal(out, buf)
    shuffled = [xs[1], xs[0], xs[2], xs[3]]                      # same storage, wrong order -> copy
    out2 = ops.stack_rows(shuffled)
    assert out2.data_ptr() != buf.data_ptr() and torch.equal(out2, torch.stack(shuffled, 0))
    separate = [x.clone() for x in xs]
    assert torch.equal(ops.stack_rows(separate), buf)
    grads = [x.clone().requires_grad_(True) for x in xs]         # gradient-carrying inputs keep autograd's stack
    assert ops.stack_rows(grads).grad_fn is not None


def test_workspace_query_of_fused_linear_backward():
    from tagan_b200 import _lib
    lib = _lib.load()
    assert lib.tagan_gemm_tn_colsum_workspace_bytes(384, 128, 1_600_000) >= lib.tagan_gemm_workspace_bytes(2, 384, 128, 1_600_000)
