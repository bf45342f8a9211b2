"""GPU: tcgen05 3xTF32 GEMM (precision 1) through the C ABI against an fp64 reference, all three
operand layouts (NT forward Linear, NN dX, TN dW with split-K), ragged sizes and strided operands.
Tolerance: fp32 parity bar rtol 1e-4 / atol 1e-5 relative to the output scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(op, a, b, bias, m, n, k, precision, accumulate=False, c=None):
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    dev = a.device
    if c is None:
        c = torch.empty(m, n, dtype=torch.float32, device=dev)
    ws_bytes = lib.tagan_gemm_workspace_bytes(op, m, n, k)
    ws = ops.workspace(ws_bytes, dev) if ws_bytes else None
    rc = lib.tagan_gemm(op, m, n, k, ops._ptr(a), a.stride(0), ops._ptr(b), b.stride(0), ops._ptr(bias), ops._ptr(c),
                        c.stride(0), int(accumulate), precision, ops._ptr(ws), ws.numel() if ws is not None else 0,
                        ops._stream())
    _lib.check(rc, "tagan_gemm")
    return c


def _ref(op, a, b, bias):
    a, b = a.double(), b.double()
    if op == 0:
        r = a @ b.t()
    elif op == 1:
        r = a @ b
    else:
        r = a.t() @ b
    if bias is not None:
        r = r + bias.double()
    return r


@pytest.mark.parametrize("precision", [1, 3, 5])  # 1/3 = TMA-fed kernel (3 / 4 MMAs per k-step), 5 = LDG-fed kernel
@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (128, 128, 128), (256, 384, 128), (1000, 384, 128), (4096, 768, 256),
                                   (333, 200, 100), (130, 70, 36), (5000, 128, 512)])
def test_gemm_tc_matches_fp64(op, m, n, k, precision):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(m + n + k + op)
    if op == 0:
        a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g)
    elif op == 1:
        a, b = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g)
    else:
        a, b = torch.randn(k, m, generator=g), torch.randn(k, n, generator=g)
    bias = torch.randn(n, generator=g) if op == 0 else None
    a, b = a.to(dev), b.to(dev)
    bias_d = bias.to(dev) if bias is not None else None
    ref = _ref(op, a.cpu(), b.cpu(), bias)
    out = _gemm(op, a, b, bias_d, m, n, k, precision=precision)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-4, atol=1e-5 * max(1.0, scale))
    # fp32-level accuracy: error well below what a single TF32 pass gives (~1e-3 relative)
    assert float((out.cpu().double() - ref).abs().max()) < 2e-5 * max(1.0, scale)


def test_gemm_tc_dw_long_reduction_and_accumulate():
    """dW = dY^T X with a 200k-row reduction (split-K, deterministic) and accumulate into C."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    rows, n_out, k_in = 200_000, 384, 128
    dy = (torch.randn(rows, n_out, generator=g) / 30).to(dev)
    x = torch.randn(rows, k_in, generator=g).to(dev)
    ref = dy.double().t() @ x.double()
    c1 = _gemm(2, dy, x, None, n_out, k_in, rows, precision=1)
    c2 = _gemm(2, dy, x, None, n_out, k_in, rows, precision=1)
    assert torch.equal(c1, c2)                                           # fixed-order reduction
    torch.testing.assert_close(c1.double(), ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    base = torch.ones(n_out, k_in, device=dev)
    c3 = _gemm(2, dy, x, None, n_out, k_in, rows, precision=1, accumulate=True, c=base.clone())
    torch.testing.assert_close(c3, c1 + 1.0, rtol=1e-6, atol=1e-6)


def test_gemm_tc_strided_operands_fused_qkv_slices():
    """Operands with a leading dimension larger than the row (column slices of a fused buffer)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    m, k, n = 3000, 128, 256
    big = torch.randn(m, 3 * k, generator=g).to(dev)
    a = big[:, k:2 * k]                                   # ld = 3k
    w = torch.randn(n, k, generator=g).to(dev)
    out = _gemm(0, a, w, None, m, n, k, precision=1)
    torch.testing.assert_close(out.double(), a.double() @ w.double().t(), rtol=1e-4, atol=2e-4)


def test_linear_autograd_with_tensor_core_gemm():
    from tagan_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    old = ops.GEMM_PRECISION
    ops.GEMM_PRECISION = 1
    try:
        m, k, n = 20000, 128, 384
        x = torch.randn(m, k, generator=g, requires_grad=True)
        wt = (torch.randn(n, k, generator=g) / k ** 0.5).requires_grad_(True)
        b = torch.randn(n, generator=g, requires_grad=True)
        wo = torch.randn(m, n, generator=g) / m ** 0.5
        ref = torch.nn.functional.linear(x.double(), wt.double(), b.double())
        (ref * wo.double()).sum().backward()
        xd, wd, bd = (t.detach().to(dev).requires_grad_(True) for t in (x, wt, b))
        out = ops.linear(xd, wd, bd)
        (out * wo.to(dev)).sum().backward()
        torch.testing.assert_close(out.detach().cpu(), ref.detach().float(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(xd.grad.cpu(), x.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(wd.grad.cpu(), wt.grad, rtol=1e-4, atol=2e-5)
    finally:
        ops.GEMM_PRECISION = old


@pytest.mark.parametrize("m,n,k", [(384, 128, 100_000), (128, 128, 3000), (200, 72, 1031), (512, 256, 40_000), (16, 16, 50)])
def test_gemm_tn_colsum(m, n, k):
    """dW = A^T B with the column sums of A (the bias gradient) accumulated in the same pass: both outputs against
    fp64 torch (dW to the 3xTF32 tolerance of the other GEMM tests, column sums to fp32 summation-order tolerance),
    and bit-identical reruns (fixed-order partial reductions)."""
    from tagan_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(m + n + k)
    a = torch.randn(k, m, device=dev)
    b = torch.randn(k, n, device=dev)
    c = torch.empty(m, n, device=dev)
    cs = ops.gemm_tn_colsum(m, n, k, a, m, b, n, c, n)
    ref = (a.double().t() @ b.double())
    scale = float(ref.abs().max())
    torch.testing.assert_close(c.double(), ref, rtol=1e-4, atol=1e-5 * max(1.0, scale))      # the file's parity bar
    cref = a.double().sum(0)
    assert float((cs.double() - cref).abs().max()) <= 1e-5 * max(1.0, float(cref.abs().max()))
    c2 = torch.empty_like(c)
    cs2 = ops.gemm_tn_colsum(m, n, k, a, m, b, n, c2, n)
    assert torch.equal(c, c2) and torch.equal(cs, cs2)
