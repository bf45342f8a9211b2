"""Thin host wrappers over the C ABI plus the autograd glue.

Each wrapper checks that tensors live on a CUDA device (no CPU fallback), hands raw pointers
and torch's current stream to libtagan_b200.so and raises ``RuntimeError`` on a non-zero
return code (SURVEY.md section 8b error convention).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

METRICS = ["scaled_dot_product", "dot_product", "cosine_similarity", "euclidean", "squared_euclidean",
           "manhattan", "cosine_distance", "gaussian_kernel", "rbf_kernel"]
METRIC_ID = {m: i for i, m in enumerate(METRICS)}

# 0 = fp32 FFMA, 1 = 3xTF32 tcgen05 (fp32-accurate), 2 = 1xTF32 tcgen05
GEMM_PRECISION = 1

# stage-level fusion (tagan_b200/fused.py): fused GEMM epilogues, fused row passes, in-place residual gradients.
# False = the op-by-op composition of the stand-alone kernels (kept for dropout > 0 in training and as a test reference)
FUSION = True

# number of libtagan_b200 kernels launched (bench.py reports `gpu_launches` from this)
CALLS = {"n": 0}

# When a dict, the named C calls are bracketed with CUDA events on the launching stream
# (bench.py's live per-kernel timing); None = off.
PROFILE = None


class _timed:
    """Bracket a library call with CUDA events on the launching stream when ``PROFILE`` is a dict; ``nbytes`` = the
    call's algorithmic bytes (recorded beside the events so bench.py can quote GB/s per kernel family)."""

    def __init__(self, name, nbytes=0):
        self.name = name
        self.nbytes = nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e.record(torch.cuda.current_stream())
            PROFILE.setdefault(self.name, []).append((self.s, self.e, self.nbytes))


def _ptr(t):
    """Device pointer of a tensor (or a raw ``c_void_p`` passed through)."""
    if t is None or isinstance(t, C.c_void_p):
        return t
    if not t.is_cuda:
        raise RuntimeError("tagan_b200 has no CPU path: tensor must live on a CUDA device")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, last dim contiguous."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 0 or t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _rows(t: torch.Tensor):
    """View a [..., C] tensor as (rows, C, ld); copies only if rows are not uniformly strided."""
    t = _f32c(t)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() > 2:
        if not t.is_contiguous():
            t = t.contiguous()
        t = t.view(-1, t.shape[-1])
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else t.shape[1]
    return t, t.shape[0], t.shape[1], ld


_WS = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream); stream order makes reuse safe."""
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# ----------------------------------------------------------------------------------------
# (a1) CSR
# ----------------------------------------------------------------------------------------
@dataclass
class CSR:
    num_nodes: int
    num_edges: int
    rowptr: torch.Tensor
    col: torch.Tensor
    row: torch.Tensor
    rowptr_t: Optional[torch.Tensor]
    row_t: Optional[torch.Tensor]
    perm_t: Optional[torch.Tensor]
    status: torch.Tensor
    ready: Optional[torch.cuda.Event] = None    # set when the CSR was built on a side stream (see build_csr_async)

    @property
    def nnz(self) -> int:          # host sync; only for tests / attention-weight export
        return int(self.rowptr[-1].item())


def build_csr(edge_index: torch.Tensor, num_nodes: int, transpose: bool = True, validate: bool = False) -> CSR:
    """Device CSR of ``adj[ei[0],ei[1]]=1; adj+=eye`` (reference graph_attention.py:98-102)."""
    lib = _lib.load()
    if not edge_index.is_cuda:
        raise RuntimeError("tagan_b200 has no CPU path: edge_index must live on a CUDA device")
    ei = edge_index
    if ei.dtype != torch.int64:
        ei = ei.long()
    ei = ei.contiguous()
    e = ei.shape[1] if ei.numel() else 0
    dev = ei.device
    cap = e + num_nodes
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr = torch.empty(num_nodes + 1, **i32)
    col = torch.empty(max(cap, 1), **i32)
    row = torch.empty(max(cap, 1), **i32)
    status = torch.empty(1, **i32)
    rowptr_t = row_t = perm_t = None
    if transpose:
        rowptr_t = torch.empty(num_nodes + 1, **i32)
        row_t = torch.empty(max(cap, 1), **i32)
        perm_t = torch.empty(max(cap, 1), **i32)
    nbytes = lib.tagan_csr_workspace_bytes(e, num_nodes)
    ws = workspace(nbytes, dev)
    with _timed("csr_build"):
        rc = lib.tagan_csr_build(_ptr(ei) if e else None, e, num_nodes, _ptr(rowptr), _ptr(col), _ptr(row),
                                 _ptr(rowptr_t), _ptr(row_t), _ptr(perm_t), _ptr(status), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "tagan_csr_build")
    CALLS["n"] += 22 if transpose else 12
    if validate and int(status.item()) != 0:
        raise IndexError("edge_index out of range for num_nodes=%d" % num_nodes)
    return CSR(num_nodes, e, rowptr, col, row, rowptr_t, row_t, perm_t, status)


MAX_CSR_BATCH = 128
BATCHED_CSR = True          # forward_seq / TAGANModel: one block-diagonal CSR + one kernel-(a) launch per pass for all snapshots


def build_csr_batched(edge_indices, sizes, transpose: bool = True) -> CSR:
    """ONE block-diagonal CSR for T <= 128 snapshots (``edge_indices[t]``: contiguous int64 ``[2,E_t]`` on the device,
    ``sizes[t]`` nodes): rows / columns of snapshot t are offset by ``sum(sizes[:t])``, so it equals the per-snapshot CSRs
    concatenated and kernel (a) runs once over the stacked rows.  One launch set instead of T (21 launches each)."""
    lib = _lib.load()
    t_steps = len(edge_indices)
    if not 0 < t_steps <= MAX_CSR_BATCH:
        raise ValueError("build_csr_batched: 1..%d snapshots" % MAX_CSR_BATCH)
    eis = []
    for ei in edge_indices:
        if not ei.is_cuda:
            raise RuntimeError("tagan_b200 has no CPU path: edge_index must live on a CUDA device")
        if ei.dtype != torch.int64:
            ei = ei.long()
        if ei.numel() and (ei.dim() != 2 or ei.shape[0] != 2):
            raise ValueError("edge_index must be [2,E]")
        eis.append(ei if (ei.numel() == 0 or ei.stride(1) == 1) else ei.contiguous())   # column slices of one [2,E] are fine
    dev = eis[0].device
    ecounts = [int(ei.shape[1]) if ei.numel() else 0 for ei in eis]
    e_tot, n_tot = sum(ecounts), int(sum(sizes))
    if e_tot + n_tot >= 2 ** 31:
        raise NotImplementedError("build_csr_batched: sum(E_t) + sum(N_t) must stay below 2^31")
    src = (C.c_void_p * t_steps)(*[ei[0].data_ptr() if ec else None for ei, ec in zip(eis, ecounts)])
    dst = (C.c_void_p * t_steps)(*[ei[1].data_ptr() if ec else None for ei, ec in zip(eis, ecounts)])
    ec_arr = (C.c_int64 * t_steps)(*ecounts)
    nc_arr = (C.c_int32 * t_steps)(*[int(s) for s in sizes])
    cap = e_tot + n_tot
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr = torch.empty(n_tot + 1, **i32)
    col = torch.empty(max(cap, 1), **i32)
    row = torch.empty(max(cap, 1), **i32)
    status = torch.empty(1, **i32)
    rowptr_t = row_t = perm_t = None
    if transpose:
        rowptr_t = torch.empty(n_tot + 1, **i32)
        row_t = torch.empty(max(cap, 1), **i32)
        perm_t = torch.empty(max(cap, 1), **i32)
    ws = workspace(lib.tagan_csr_workspace_bytes(e_tot, n_tot), dev)
    with _timed("csr_build"):
        rc = lib.tagan_csr_build_batched(src, dst, ec_arr, nc_arr, t_steps, _ptr(rowptr), _ptr(col), _ptr(row), _ptr(rowptr_t),
                                         _ptr(row_t), _ptr(perm_t), _ptr(status), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "tagan_csr_build_batched")
    CALLS["n"] += 22 if transpose else 12
    csr = CSR(n_tot, e_tot, rowptr, col, row, rowptr_t, row_t, perm_t, status)
    csr._keepalive = eis                      # the kernels read the edge lists asynchronously
    return csr


def build_csr_batched_async(edge_indices, sizes, device, transpose: bool = True) -> CSR:
    """``build_csr_batched`` on the side stream (it depends only on the edge lists, so it runs under LN1 + the QKV projection);
    the returned CSR carries a ``ready`` event."""
    main = torch.cuda.current_stream()
    side = side_stream(device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        csr = build_csr_batched([ei.to(device) for ei in edge_indices], sizes, transpose=transpose)
        for t in (csr.rowptr, csr.col, csr.row, csr.rowptr_t, csr.row_t, csr.perm_t, csr.status):
            if t is not None:
                t.record_stream(main)
        csr.ready = torch.cuda.Event()
        csr.ready.record(side)
    return csr


_SIDE = {}


def side_stream(device) -> torch.cuda.Stream:
    """One auxiliary stream per device for work that is independent of the main chain (CSR builds)."""
    idx = torch.device(device).index
    st = _SIDE.get(idx)
    if st is None:
        st = _SIDE[idx] = torch.cuda.Stream(device=device)
    return st


def build_csr_async(edge_indices, num_nodes: int, device, transpose: bool = True, validate: bool = False):
    """Build the CSRs of several snapshots on the side stream.  They depend only on ``edge_index``, so they run under
    whatever the main stream does next (LN1 + QKV projection of the batched geometric stage, then the attention of
    earlier snapshots); each CSR carries a ``ready`` event that its consumer waits for."""
    main = torch.cuda.current_stream()
    side = side_stream(device)
    side.wait_stream(main)
    out = []
    with torch.cuda.stream(side):
        for ei in edge_indices:
            if isinstance(ei, CSR):
                out.append(ei)
                continue
            csr = build_csr(ei.to(device), num_nodes, transpose=transpose, validate=validate)
            for t in (csr.rowptr, csr.col, csr.row, csr.rowptr_t, csr.row_t, csr.perm_t, csr.status):
                if t is not None:
                    t.record_stream(main)
            csr.ready = torch.cuda.Event()
            csr.ready.record(side)
            out.append(csr)
    return out


def wait_csr(csr: CSR) -> None:
    if csr.ready is not None:
        torch.cuda.current_stream().wait_event(csr.ready)


# ----------------------------------------------------------------------------------------
# GEMM / Linear
# ----------------------------------------------------------------------------------------
def gemm(op: int, m: int, n: int, k: int, a, lda, b, ldb, bias, c, ldc, accumulate=False):
    lib = _lib.load()
    nbytes = lib.tagan_gemm_workspace_bytes(op, m, n, k)
    dev = c.device if isinstance(c, torch.Tensor) else torch.device("cuda", torch.cuda.current_device())
    ws = workspace(nbytes, dev) if nbytes else None
    with _timed("gemm", 4 * (m * k + n * k + m * n * (2 if accumulate else 1))):
        rc = lib.tagan_gemm(op, m, n, k, _ptr(a), lda, _ptr(b), ldb, _ptr(bias), _ptr(c), ldc, int(accumulate),
                            GEMM_PRECISION, _ptr(ws), ws.numel() if ws is not None else 0, _stream())
    _lib.check(rc, "tagan_gemm")
    CALLS["n"] += 1


def colsum(x2d: torch.Tensor, rows: int, cols: int, ld: int) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    nbytes = lib.tagan_colsum_workspace_bytes(rows, cols)
    ws = workspace(nbytes, x2d.device)
    rc = lib.tagan_colsum(_ptr(x2d), ld, _ptr(out), _ptr(ws), ws.numel(), rows, cols, _stream())
    _lib.check(rc, "tagan_colsum")
    CALLS["n"] += 1
    return out


def gemm_tn_colsum(m: int, n: int, k: int, a, lda, b, ldb, c, ldc) -> torch.Tensor:
    """C[m,n] = A[k,m]^T . B[k,n] and returns colsum(A) [m]  (dW and db of a Linear in one pass over dY)."""
    lib = _lib.load()
    dev = c.device
    out = torch.empty(m, dtype=torch.float32, device=dev)
    nbytes = lib.tagan_gemm_tn_colsum_workspace_bytes(m, n, k)
    ws = workspace(nbytes, dev)
    with _timed("gemm", 4 * (m * k + n * k + m * n)):
        rc = lib.tagan_gemm_tn_colsum(m, n, k, _ptr(a), lda, _ptr(b), ldb, _ptr(c), ldc, _ptr(out), GEMM_PRECISION,
                                      _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "tagan_gemm_tn_colsum")
    CALLS["n"] += 3
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b  (nn.Linear); dX = dY W, dW = dY^T X, db = colsum(dY)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2, m, k, ldx = _rows(x)
        w = _f32c(weight)
        n = w.shape[0]
        y = torch.empty(m, n, dtype=torch.float32, device=x.device)
        gemm(0, m, n, k, x2, ldx, w, w.stride(0), bias, y, n)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], n)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2, m, n, ldy = _rows(dy)
        k = w.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(m, k, dtype=torch.float32, device=dy.device)
            gemm(1, m, k, n, dy2, ldy, w, w.stride(0), None, dx, k)
            dx = dx.view(ctx.in_shape)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            dw = torch.empty(n, k, dtype=torch.float32, device=dy.device)
            ldx = x2.stride(0) if m > 1 else k
            if want_db:
                db = gemm_tn_colsum(n, k, m, dy2, ldy, x2, ldx, dw, k)
            else:
                gemm(2, n, k, m, dy2, ldy, x2, ldx, None, dw, k)
        elif want_db:
            db = colsum(dy2, m, n, ldy)
        return dx, dw, db


def linear(x, weight, bias=None):
    return _LinearFn.apply(x, weight, bias)


# ----------------------------------------------------------------------------------------
# LayerNorm with fused residual / row scale
# ----------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    """y = LN(x [+ res]) * rowscale; gamma=None means plain (x+res)*rowscale."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, rowscale):
        lib = _lib.load()
        x2, rows, cols, ldx = _rows(x)
        r2 = ldres = None
        if res is not None:
            r2, _, _, ldres = _rows(res)
        y = torch.empty(rows, cols, dtype=torch.float32, device=x.device)
        xsum = torch.empty_like(y) if res is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device) if gamma is not None else None
        rstd = torch.empty_like(mean) if gamma is not None else None
        rc = lib.tagan_layernorm_fwd(_ptr(x2), ldx, _ptr(r2), ldres or 0, _ptr(gamma), _ptr(beta), _ptr(rowscale),
                                     _ptr(y), cols, _ptr(xsum), _ptr(mean), _ptr(rstd), rows, cols, _stream())
        _lib.check(rc, "tagan_layernorm_fwd")
        CALLS["n"] += 1
        saved_x = xsum if xsum is not None else x2
        ctx.save_for_backward(saved_x, gamma, rowscale, mean, rstd)
        ctx.has_res = res is not None
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        xs, gamma, rowscale, mean, rstd = ctx.saved_tensors
        dy2, rows, cols, lddy = _rows(dy)
        dx = torch.empty(rows, cols, dtype=torch.float32, device=dy.device)
        dgamma = dbeta = None
        ws = None
        if gamma is not None:
            dgamma = torch.empty(cols, dtype=torch.float32, device=dy.device)
            dbeta = torch.empty_like(dgamma)
            ws = workspace(lib.tagan_layernorm_bwd_workspace_bytes(rows, cols), dy.device)
        ldx = xs.stride(0) if rows > 1 else cols
        rc = lib.tagan_layernorm_bwd(_ptr(dy2), lddy, _ptr(xs), ldx, _ptr(gamma), _ptr(rowscale), _ptr(mean),
                                     _ptr(rstd), _ptr(dx), cols, 0, _ptr(dgamma), _ptr(dbeta), _ptr(ws),
                                     ws.numel() if ws is not None else 0, rows, cols, _stream())
        _lib.check(rc, "tagan_layernorm_bwd")
        CALLS["n"] += 1
        dx = dx.view(ctx.shape)
        return dx, (dx if ctx.has_res else None), dgamma, dbeta, None


def layer_norm(x, gamma=None, beta=None, res=None, rowscale=None):
    """``LN(x + res)``; with gamma None just ``x + res`` (use_layer_norm=False paths)."""
    if gamma is None and res is None and rowscale is None:
        return x
    return _LayerNormFn.apply(x, res, gamma, beta, rowscale)


# ----------------------------------------------------------------------------------------
# (a2-a4) geometric attention core
# ----------------------------------------------------------------------------------------
class _GeoAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, metric_param, csr: CSR, heads: int, metric: int, want_attn: bool):
        lib = _lib.load()
        qkv2, n, three_h, ld = _rows(qkv)
        h = three_h // 3
        ctxv = torch.empty(n, h, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(n, heads, dtype=torch.float32, device=qkv.device)
        attn = None
        if want_attn:
            attn = torch.zeros(max(csr.num_edges + csr.num_nodes, 1), heads, dtype=torch.float32, device=qkv.device)
        base = qkv2.data_ptr()
        q, k, v = (C.c_void_p(base + i * h * 4) for i in range(3))
        with _timed("geo_attn_fwd"):
            rc = lib.tagan_geo_attn_fwd(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), n, h, heads, metric,
                                        _ptr(metric_param), _ptr(ctxv), _ptr(lse), _ptr(attn), _stream())
        _lib.check(rc, "tagan_geo_attn_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(qkv2, metric_param, ctxv, lse)
        ctx.csr, ctx.heads, ctx.metric = csr, heads, metric
        if want_attn:
            ctx.mark_non_differentiable(attn)
            return ctxv, attn
        return ctxv, None

    @staticmethod
    def backward(ctx, dctx, _dattn):
        lib = _lib.load()
        qkv2, metric_param, ctxv, lse = ctx.saved_tensors
        csr = ctx.csr
        if csr.rowptr_t is None:
            raise RuntimeError("CSR was built without its transpose; backward needs it")
        n, three_h = qkv2.shape
        h = three_h // 3
        ld = qkv2.stride(0) if n > 1 else three_h
        dctx = _f32c(dctx).contiguous()
        dqkv = torch.empty(n, three_h, dtype=torch.float32, device=dctx.device)
        delta = torch.empty(n, ctx.heads, dtype=torch.float32, device=dctx.device)
        want_dp = metric_param is not None and ctx.metric in (7, 8)
        dp_ws = torch.empty(n, ctx.heads, dtype=torch.float32, device=dctx.device) if want_dp else None
        dparam = torch.empty(ctx.heads, dtype=torch.float32, device=dctx.device) if want_dp else None
        base, dbase = qkv2.data_ptr(), dqkv.data_ptr()
        q, k, v = (C.c_void_p(base + i * h * 4) for i in range(3))
        dq, dk, dv = (C.c_void_p(dbase + i * h * 4) for i in range(3))
        with _timed("geo_attn_bwd"):
            rc = lib.tagan_geo_attn_bwd(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.rowptr_t), _ptr(csr.row_t),
                                        n, h, ctx.heads, ctx.metric, _ptr(metric_param), _ptr(ctxv), _ptr(lse), _ptr(dctx),
                                        dq, dk, dv, three_h, _ptr(delta), _ptr(dp_ws), _ptr(dparam), _stream())
        _lib.check(rc, "tagan_geo_attn_bwd")
        CALLS["n"] += 3 if want_dp else 2
        return dqkv, dparam, None, None, None, None


def geo_shape_supported(hidden: int, heads: int) -> bool:
    """Shapes kernel (a) is instantiated for (csrc/geo_attn.cu pick_shape): hidden in {32,64,128,256,512} and a power-of-two
    number of lanes per head -- every named configuration; e.g. 5 heads x 8 is not."""
    if hidden not in (32, 64, 128, 256, 512) or heads <= 0 or hidden % heads:
        return False
    vec = {32: 1, 64: 2}.get(hidden, 4)
    d = hidden // heads
    if d % vec:
        return False
    group = d // vec
    return 1 <= group <= 32 and (group & (group - 1)) == 0


def geo_attention_core(qkv, csr: CSR, heads: int, metric: str, metric_param=None, want_attn=False):
    """qkv ``[N,3H]`` (fused projection) -> ctx ``[N,H]`` (and per-entry weights ``[cap,h]``)."""
    return _GeoAttnFn.apply(qkv, metric_param, csr, heads, METRIC_ID[metric], want_attn)


class _GeoAttnSeqFn(torch.autograd.Function):
    """Kernel (a) over T snapshots that share one stacked projection ``qkv [T*N,3H]``: one launch pair per
    snapshot on row slices, one autograd node for the whole stage."""

    @staticmethod
    def forward(ctx, qkv, metric_param, csrs, heads: int, metric: int):
        lib = _lib.load()
        qkv2, rows, three_h, ld = _rows(qkv)
        t_steps = len(csrs)
        n = csrs[0].num_nodes
        if rows != t_steps * n or any(c.num_nodes != n for c in csrs):
            raise ValueError("geo_attention_seq: qkv rows must be T * N with the same N in every snapshot")
        h = three_h // 3
        ctxv = torch.empty(rows, h, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(rows, heads, dtype=torch.float32, device=qkv.device)
        base, cbase, lbase = qkv2.data_ptr(), ctxv.data_ptr(), lse.data_ptr()
        for t, csr in enumerate(csrs):
            off = base + t * n * ld * 4
            q, k, v = (C.c_void_p(off + i * h * 4) for i in range(3))
            wait_csr(csr)
            with _timed("geo_attn_fwd"):
                rc = lib.tagan_geo_attn_fwd(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), n, h, heads, metric,
                                            _ptr(metric_param), C.c_void_p(cbase + t * n * h * 4),
                                            C.c_void_p(lbase + t * n * heads * 4), None, _stream())
            _lib.check(rc, "tagan_geo_attn_fwd")
        CALLS["n"] += t_steps
        ctx.save_for_backward(qkv2, metric_param, ctxv, lse)
        ctx.csrs, ctx.heads, ctx.metric = csrs, heads, metric
        return ctxv

    @staticmethod
    def backward(ctx, dctx):
        lib = _lib.load()
        qkv2, metric_param, ctxv, lse = ctx.saved_tensors
        csrs, heads, metric = ctx.csrs, ctx.heads, ctx.metric
        if any(c.rowptr_t is None for c in csrs):
            raise RuntimeError("CSR was built without its transpose; backward needs it")
        rows, three_h = qkv2.shape
        h = three_h // 3
        t_steps = len(csrs)
        n = csrs[0].num_nodes
        ld = qkv2.stride(0) if rows > 1 else three_h
        dev = dctx.device
        dctx = _f32c(dctx).contiguous()
        dqkv = torch.empty(rows, three_h, dtype=torch.float32, device=dev)
        delta = torch.empty(n, heads, dtype=torch.float32, device=dev)
        want_dp = metric_param is not None and metric in (7, 8)
        dp_ws = torch.empty(n, heads, dtype=torch.float32, device=dev) if want_dp else None
        dparam_t = torch.empty(t_steps, heads, dtype=torch.float32, device=dev) if want_dp else None
        base, dbase = qkv2.data_ptr(), dqkv.data_ptr()
        for t, csr in enumerate(csrs):
            off, doff = base + t * n * ld * 4, dbase + t * n * three_h * 4
            q, k, v = (C.c_void_p(off + i * h * 4) for i in range(3))
            dq, dk, dv = (C.c_void_p(doff + i * h * 4) for i in range(3))
            with _timed("geo_attn_bwd"):
                rc = lib.tagan_geo_attn_bwd(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.rowptr_t), _ptr(csr.row_t),
                                            n, h, heads, metric, _ptr(metric_param),
                                            C.c_void_p(ctxv.data_ptr() + t * n * h * 4),
                                            C.c_void_p(lse.data_ptr() + t * n * heads * 4),
                                            C.c_void_p(dctx.data_ptr() + t * n * h * 4), dq, dk, dv, three_h, _ptr(delta),
                                            _ptr(dp_ws), C.c_void_p(dparam_t.data_ptr() + t * heads * 4) if want_dp else None,
                                            _stream())
            _lib.check(rc, "tagan_geo_attn_bwd")
        CALLS["n"] += (3 if want_dp else 2) * t_steps
        dparam = dparam_t.sum(0) if want_dp else None
        return dqkv, dparam, None, None, None


def geo_attention_seq(qkv, csrs, heads: int, metric: str, metric_param=None):
    """qkv ``[T*N,3H]`` (stacked fused projection) + T CSRs over the same N nodes -> ctx ``[T*N,H]``."""
    return _GeoAttnSeqFn.apply(qkv, metric_param, list(csrs), heads, METRIC_ID[metric])


def stack_rows(xs) -> torch.Tensor:
    """``torch.stack(xs, 0)`` that costs nothing when the tensors already are consecutive slices of one
    allocation (what ``GraphedStep`` hands out) and none of them carries a gradient."""
    xs = list(xs)
    x0 = xs[0]
    if isinstance(x0, torch.Tensor) and x0.is_contiguous() and not any(x.requires_grad for x in xs):
        numel, base, off0 = x0.numel(), x0.untyped_storage().data_ptr(), x0.storage_offset()
        if numel and all(x.shape == x0.shape and x.dtype == x0.dtype and x.is_contiguous()
                         and x.untyped_storage().data_ptr() == base and x.storage_offset() == off0 + i * numel
                         for i, x in enumerate(xs)):
            return torch.as_strided(x0, (len(xs),) + tuple(x0.shape), (numel,) + tuple(x0.stride()), off0)
    return torch.stack(xs, 0)


# ----------------------------------------------------------------------------------------
# element-wise helpers
# ----------------------------------------------------------------------------------------
class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        a2, b2 = _f32c(a).contiguous(), _f32c(b).contiguous()
        out = torch.empty_like(a2)
        rc = lib.tagan_axpby(_ptr(a2), 1.0, _ptr(b2), 1.0, _ptr(out), a2.numel(), _stream())
        _lib.check(rc, "tagan_axpby")
        CALLS["n"] += 1
        return out

    @staticmethod
    def backward(ctx, d):
        return d, d


def add(a, b):
    return _AddFn.apply(a, b)


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x2 = _f32c(x).contiguous()
        y = torch.empty_like(x2)
        _lib.check(lib.tagan_gelu_fwd(_ptr(x2), _ptr(y), x2.numel(), _stream()), "tagan_gelu_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(x2)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (x2,) = ctx.saved_tensors
        dy2 = _f32c(dy).contiguous()
        dx = torch.empty_like(x2)
        _lib.check(lib.tagan_gelu_bwd(_ptr(dy2), _ptr(x2), _ptr(dx), x2.numel(), _stream()), "tagan_gelu_bwd")
        CALLS["n"] += 1
        return dx


def gelu(x):
    return _GeluFn.apply(x)


# ----------------------------------------------------------------------------------------
# (b1,b2) temporal attention core
# ----------------------------------------------------------------------------------------
@dataclass
class TemporalMask:
    """Resolved mask of one AsymmetricTemporalAttention call (see layers.py for the rules)."""
    flags: int = 0                              # bit0 causal, bit1 band, bit2 allones => causal
    band: float = 10.0
    ts: Optional[torch.Tensor] = None           # [B,T] fp32
    mask: Optional[torch.Tensor] = None         # uint8 [mask_b, mask_h, T, T]
    allones_flag: Optional[torch.Tensor] = None  # int32 [1] device


def mask_allones_flag(ts, mask_u8, batch: int, t: int, band: float, device) -> torch.Tensor:
    lib = _lib.load()
    flag = torch.empty(1, dtype=torch.int32, device=device)
    rc = lib.tagan_tattn_mask_allones(_ptr(ts), batch, t, band, _ptr(mask_u8),
                                      mask_u8.numel() if mask_u8 is not None else 0, _ptr(flag), _stream())
    _lib.check(rc, "tagan_tattn_mask_allones")
    CALLS["n"] += 2
    return flag


class _TAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, bias, tmask: TemporalMask, batch: int, t: int, heads: int, time_major: bool,
                want_attn: bool):
        lib = _lib.load()
        qkv2, rows, three_h, ld = _rows(qkv)
        h = three_h // 3
        assert rows == batch * t
        dev = qkv.device
        ctxv = torch.empty(rows, h, dtype=torch.float32, device=dev)
        lse = torch.empty(batch, heads, t, dtype=torch.float32, device=dev)
        attn = torch.empty(batch, heads, t, t, dtype=torch.float32, device=dev) if want_attn else None
        bias_c = bias_t = None
        bstride = 0
        if bias is not None:
            bias_c = _f32c(bias).contiguous()
            bias_t = bias_c.transpose(-1, -2).contiguous()
            bstride = heads * t * t if bias_c.dim() == 4 else 0
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        base = qkv2.data_ptr()
        q, k, v = (C.c_void_p(base + i * h * 4) for i in range(3))
        with _timed("tattn_fwd"):
            rc = lib.tagan_tattn_fwd(q, k, v, ld, batch, t, h, heads, int(time_major), _ptr(bias_c), _ptr(bias_t), bstride,
                                     _ptr(tmask.ts), tmask.flags, tmask.band, _ptr(tmask.allones_flag), _ptr(m), mb, mh,
                                     _ptr(ctxv), _ptr(lse), _ptr(attn), _stream())
        _lib.check(rc, "tagan_tattn_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(qkv2, bias_c, bias_t, ctxv, lse)
        ctx.tmask, ctx.dims, ctx.bstride = tmask, (batch, t, heads, time_major), bstride
        ctx.bias_shape = bias.shape if bias is not None else None
        if want_attn:
            ctx.mark_non_differentiable(attn)
        return ctxv, attn

    @staticmethod
    def backward(ctx, dctx, _dattn):
        lib = _lib.load()
        qkv2, bias_c, bias_t, ctxv, lse = ctx.saved_tensors
        tmask = ctx.tmask
        batch, t, heads, time_major = ctx.dims
        rows, three_h = qkv2.shape
        h = three_h // 3
        ld = qkv2.stride(0) if rows > 1 else three_h
        dev = dctx.device
        dctx = _f32c(dctx).contiguous()
        dqkv = torch.empty(rows, three_h, dtype=torch.float32, device=dev)
        dbias = ws = None
        if bias_c is not None and ctx.needs_input_grad[1]:
            dbias = torch.empty_like(bias_c)
            if ctx.bstride == 0:
                ws = workspace(lib.tagan_tattn_bwd_workspace_bytes(batch, t, heads), dev)
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        base, dbase = qkv2.data_ptr(), dqkv.data_ptr()
        q, k, v = (C.c_void_p(base + i * h * 4) for i in range(3))
        dq, dk, dv = (C.c_void_p(dbase + i * h * 4) for i in range(3))
        with _timed("tattn_bwd"):
            rc = lib.tagan_tattn_bwd(q, k, v, ld, batch, t, h, heads, int(time_major), _ptr(bias_c), _ptr(bias_t),
                                     ctx.bstride, _ptr(tmask.ts), tmask.flags, tmask.band, _ptr(tmask.allones_flag),
                                     _ptr(m), mb, mh, _ptr(ctxv), _ptr(lse), _ptr(dctx), dq, dk, dv, three_h,
                                     _ptr(dbias), _ptr(ws), ws.numel() if ws is not None else 0, _stream())
        _lib.check(rc, "tagan_tattn_bwd")
        CALLS["n"] += 2 if ws is not None else 1
        if dbias is not None:
            dbias = dbias.view(ctx.bias_shape)
        return dqkv, dbias, None, None, None, None, None, None


def temporal_attention_core(qkv, bias, tmask: TemporalMask, batch: int, t: int, heads: int,
                            time_major: bool = False, want_attn: bool = False):
    """qkv rows ``[B*T,3H]`` -> ctx rows ``[B*T,H]`` (+ ``attn[B,h,T,T]``)."""
    return _TAttnFn.apply(qkv, bias, tmask, batch, t, heads, time_major, want_attn)


class _TAttnPerNodeFn(torch.autograd.Function):
    """Temporal attention core with PER-NODE timestamps: the RBF time bias (reference temporal_attention.py:792-871) is
    produced on device for a chunk of nodes at a time (``tagan_time_bias_fwd``), consumed by the attention kernel through
    its per-node bias path and discarded; backward recomputes the tile, takes the chunk's dBias and reduces it to the
    parameter gradients at once (``tagan_time_bias_bwd``).  No ``[B,T,T,nb]`` or whole-batch ``[B,h,T,T]`` tensor exists."""

    CHUNK_BYTES = 16 << 20          # bias tile per chunk (three such tiles live at once in backward): L2-resident

    @staticmethod
    def forward(ctx, qkv, pos_bias, ts, mu, sigma, wc, bc, tmask: "TemporalMask", batch: int, t: int, heads: int, time_major: bool,
                want_attn: bool):
        lib = _lib.load()
        qkv2, rows, three_h, ld = _rows(qkv)
        h = three_h // 3
        assert rows == batch * t
        dev = qkv.device
        nb = mu.numel()
        ts_c = _f32c(ts).contiguous()
        ctxv = torch.empty(rows, h, dtype=torch.float32, device=dev)
        lse = torch.empty(batch, heads, t, dtype=torch.float32, device=dev)
        attn = torch.empty(batch, heads, t, t, dtype=torch.float32, device=dev) if want_attn else None
        rng = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.check(lib.tagan_ts_range(_ptr(ts_c), batch, t, _ptr(rng), _stream()), "tagan_ts_range")
        pos_c = _f32c(pos_bias).contiguous() if pos_bias is not None else None
        mu_c, sg_c, wc_c, bc_c = (_f32c(v).contiguous() for v in (mu, sigma, wc, bc))
        htt = heads * t * t
        chunk = max(1, min(batch, _TAttnPerNodeFn.CHUNK_BYTES // (htt * 4)))
        bias = torch.empty(chunk, heads, t, t, dtype=torch.float32, device=dev)
        bias_t = torch.empty_like(bias)
        rsb, rst = (1, batch) if time_major else (t, 1)
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        base, cbase = qkv2.data_ptr(), ctxv.data_ptr()
        for b0 in range(0, batch, chunk):
            bc_n = min(chunk, batch - b0)
            _lib.check(lib.tagan_time_bias_fwd(_ptr(ts_c), b0, bc_n, t, heads, nb, _ptr(rng), _ptr(mu_c), _ptr(sg_c), _ptr(wc_c),
                                               _ptr(bc_c), _ptr(pos_c), _ptr(bias), _ptr(bias_t), _stream()), "tagan_time_bias_fwd")
            off = base + b0 * rsb * ld * 4
            q, k, v = (C.c_void_p(off + i * h * 4) for i in range(3))
            mptr = C.c_void_p(m.data_ptr() + b0 * mh * t * t) if (m is not None and mb > 1) else _ptr(m)
            with _timed("tattn_fwd"):
                rc = lib.tagan_tattn_fwd_strided(q, k, v, ld, bc_n, t, h, heads, rsb, rst, _ptr(bias), _ptr(bias_t), htt,
                                                 C.c_void_p(ts_c.data_ptr() + b0 * t * 4), tmask.flags, tmask.band,
                                                 _ptr(tmask.allones_flag), mptr, mb if mb == 1 else bc_n, mh,
                                                 C.c_void_p(cbase + b0 * rsb * h * 4), C.c_void_p(lse.data_ptr() + b0 * heads * t * 4),
                                                 C.c_void_p(attn.data_ptr() + b0 * htt * 4) if attn is not None else None, _stream())
            _lib.check(rc, "tagan_tattn_fwd_strided")
            CALLS["n"] += 2
        ctx.save_for_backward(qkv2, pos_c, ts_c, mu_c, sg_c, wc_c, bc_c, ctxv, lse, rng)
        ctx.tmask, ctx.dims, ctx.chunk = tmask, (batch, t, heads, time_major, nb), chunk
        ctx.pos_shape = pos_bias.shape if pos_bias is not None else None
        if want_attn:
            ctx.mark_non_differentiable(attn)
        return ctxv, attn

    @staticmethod
    def backward(ctx, dctx, _dattn):
        lib = _lib.load()
        qkv2, pos_c, ts_c, mu_c, sg_c, wc_c, bc_c, ctxv, lse, rng = ctx.saved_tensors
        tmask = ctx.tmask
        batch, t, heads, time_major, nb = ctx.dims
        chunk = ctx.chunk
        rows, three_h = qkv2.shape
        h = three_h // 3
        ld = qkv2.stride(0) if rows > 1 else three_h
        dev = dctx.device
        dctx = _f32c(dctx).contiguous()
        dqkv = torch.empty(rows, three_h, dtype=torch.float32, device=dev)
        htt = heads * t * t
        bias = torch.empty(chunk, heads, t, t, dtype=torch.float32, device=dev)
        bias_t = torch.empty_like(bias)
        dbias = torch.empty_like(bias)
        dparams = torch.zeros(heads * nb + heads + 2 * nb, dtype=torch.float32, device=dev)
        dpos = torch.zeros(heads, t, t, dtype=torch.float32, device=dev) if pos_c is not None else None
        ws = workspace(lib.tagan_time_bias_bwd_workspace_bytes(heads, nb), dev)
        rsb, rst = (1, batch) if time_major else (t, 1)
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        base, dbase = qkv2.data_ptr(), dqkv.data_ptr()
        for b0 in range(0, batch, chunk):
            bc_n = min(chunk, batch - b0)
            _lib.check(lib.tagan_time_bias_fwd(_ptr(ts_c), b0, bc_n, t, heads, nb, _ptr(rng), _ptr(mu_c), _ptr(sg_c), _ptr(wc_c),
                                               _ptr(bc_c), _ptr(pos_c), _ptr(bias), _ptr(bias_t), _stream()), "tagan_time_bias_fwd")
            off, doff = base + b0 * rsb * ld * 4, dbase + b0 * rsb * three_h * 4
            q, k, v = (C.c_void_p(off + i * h * 4) for i in range(3))
            dq, dk, dv = (C.c_void_p(doff + i * h * 4) for i in range(3))
            mptr = C.c_void_p(m.data_ptr() + b0 * mh * t * t) if (m is not None and mb > 1) else _ptr(m)
            with _timed("tattn_bwd"):
                rc = lib.tagan_tattn_bwd_strided(q, k, v, ld, bc_n, t, h, heads, rsb, rst, _ptr(bias), _ptr(bias_t), htt,
                                                 C.c_void_p(ts_c.data_ptr() + b0 * t * 4), tmask.flags, tmask.band,
                                                 _ptr(tmask.allones_flag), mptr, mb if mb == 1 else bc_n, mh,
                                                 C.c_void_p(ctxv.data_ptr() + b0 * rsb * h * 4),
                                                 C.c_void_p(lse.data_ptr() + b0 * heads * t * 4),
                                                 C.c_void_p(dctx.data_ptr() + b0 * rsb * h * 4), dq, dk, dv, three_h, _ptr(dbias),
                                                 None, 0, _stream())
            _lib.check(rc, "tagan_tattn_bwd_strided")
            rc = lib.tagan_time_bias_bwd(_ptr(ts_c), b0, bc_n, t, heads, nb, _ptr(rng), _ptr(mu_c), _ptr(sg_c), _ptr(wc_c),
                                         _ptr(dbias), _ptr(dparams), _ptr(dpos), 1, _ptr(ws), ws.numel(), _stream())
            _lib.check(rc, "tagan_time_bias_bwd")
            CALLS["n"] += 5
        dwc = dparams[:heads * nb].view(heads, nb)
        dbc = dparams[heads * nb:heads * nb + heads]
        dmu = dparams[heads * nb + heads:heads * nb + heads + nb]
        dsg = dparams[heads * nb + heads + nb:]
        if dpos is not None:
            dpos = dpos.view(ctx.pos_shape)
        return dqkv, dpos, None, dmu, dsg, dwc, dbc, None, None, None, None, None, None


def temporal_attention_core_per_node(qkv, pos_bias, ts, mu, sigma, wc, bc, tmask: "TemporalMask", batch: int, t: int, heads: int,
                                     time_major: bool = False, want_attn: bool = False):
    """qkv rows ``[B*T,3H]`` + per-node timestamps ``ts [B,T]`` -> ctx rows ``[B*T,H]`` (+ ``attn[B,h,T,T]``); the RBF time
    bias parameters (``mu``, clamped ``sigma`` ``[nb]``, ``wc [h,nb]``, ``bc [h]``) receive their gradients directly."""
    return _TAttnPerNodeFn.apply(qkv, pos_bias, ts, mu, sigma, wc, bc, tmask, batch, t, heads, time_major, want_attn)


# ----------------------------------------------------------------------------------------
# (b3-b6) propagation element-wise stages
# ----------------------------------------------------------------------------------------
class _GatesFn(torch.autograd.Function):
    """g [rows,2H] = [reset|update] pre-activations, second [rows,H] -> (r*second, z)."""

    @staticmethod
    def forward(ctx, g, second):
        lib = _lib.load()
        g2, rows, two_h, _ = _rows(g)
        g2 = g2.contiguous()
        h = two_h // 2
        s2, _, _, lds = _rows(second)
        r = torch.empty(rows, h, dtype=torch.float32, device=g.device)
        z = torch.empty_like(r)
        rs = torch.empty_like(r)
        _lib.check(lib.tagan_gates_fwd(_ptr(g2), 2 * h, _ptr(s2), lds, _ptr(r), _ptr(z), _ptr(rs), h, rows, h, _stream()),
                   "tagan_gates_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(r, z, s2)
        return rs, z

    @staticmethod
    def backward(ctx, drs, dz):
        lib = _lib.load()
        r, z, s2 = ctx.saved_tensors
        rows, h = r.shape
        lds = s2.stride(0) if rows > 1 else h
        drs = _f32c(drs).contiguous()
        dz = _f32c(dz).contiguous()
        dg = torch.empty(rows, 2 * h, dtype=torch.float32, device=r.device)
        dsecond = torch.empty(rows, h, dtype=torch.float32, device=r.device)
        _lib.check(lib.tagan_gates_bwd(_ptr(drs), h, _ptr(dz), _ptr(r), _ptr(z), _ptr(s2), lds, _ptr(dg), 2 * h,
                                       _ptr(dsecond), h, 0, rows, h, _stream()), "tagan_gates_bwd")
        CALLS["n"] += 1
        return dg, dsecond


class _BlendFn(torch.autograd.Function):
    """out = (1-z)*base + z*tanh(cand_pre) (+ base)."""

    @staticmethod
    def forward(ctx, cand_pre, z, base, residual: bool):
        lib = _lib.load()
        c2, rows, h, _ = _rows(cand_pre)
        c2 = c2.contiguous()
        z2 = _f32c(z).contiguous()
        b2, _, _, ldb = _rows(base)
        cand = torch.empty(rows, h, dtype=torch.float32, device=c2.device)
        out = torch.empty_like(cand)
        _lib.check(lib.tagan_blend_fwd(_ptr(c2), h, _ptr(z2), _ptr(b2), ldb, _ptr(cand), _ptr(out), int(residual), rows, h,
                                       _stream()), "tagan_blend_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(z2, cand, b2)
        ctx.residual = residual
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        z2, cand, b2 = ctx.saved_tensors
        rows, h = cand.shape
        ldb = b2.stride(0) if rows > 1 else h
        dout = _f32c(dout).contiguous()
        dc = torch.empty_like(cand)
        dz = torch.empty_like(cand)
        db = torch.empty_like(cand)
        _lib.check(lib.tagan_blend_bwd(_ptr(dout), _ptr(z2), _ptr(cand), _ptr(b2), ldb, _ptr(dc), h, _ptr(dz), _ptr(db), h,
                                       0, int(ctx.residual), rows, h, _stream()), "tagan_blend_bwd")
        CALLS["n"] += 1
        return dc, dz, db, None


def gru_like_cell(first, second, w_rz, b_rz, w_c, b_c, blend_with_first: bool, residual: bool):
    """Shared body of TemporalGRUCell (:531-539) and TemporalGatingUnit (:1043-1060).

    first/second: the two normalised halves of the concatenation; w_rz = [W_reset; W_update]."""
    g = linear(torch.cat([first, second], dim=-1), w_rz, b_rz)
    rs, z = _GatesFn.apply(g, second)
    cand_pre = linear(torch.cat([first, rs], dim=-1), w_c, b_c)
    return _BlendFn.apply(cand_pre, z, first if blend_with_first else second, residual)


class _WindowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, window: int, agg: int):
        lib = _lib.load()
        p2 = _f32c(p).contiguous()
        t = p2.shape[0]
        inner = p2.numel() // max(t, 1)
        out = torch.empty_like(p2)
        _lib.check(lib.tagan_skip_window_fwd(_ptr(p2), _ptr(out), t, inner, window, agg, _stream()),
                   "tagan_skip_window_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(*( (p2, out) if agg == 1 else () ))
        ctx.cfg = (t, inner, window, agg)
        return out

    @staticmethod
    def backward(ctx, dg):
        lib = _lib.load()
        t, inner, window, agg = ctx.cfg
        p2 = out = None
        if agg == 1:
            p2, out = ctx.saved_tensors
        dg = _f32c(dg).contiguous()
        dp = torch.empty_like(dg)
        _lib.check(lib.tagan_skip_window_bwd(_ptr(dg), _ptr(p2), _ptr(out), _ptr(dp), t, inner, window, agg, _stream()),
                   "tagan_skip_window_bwd")
        CALLS["n"] += 1
        return dp, None, None


AGG_ID = {"mean": 0, "max": 1, "sum": 2}


def skip_window(p, window: int, aggregation: str):
    """p ``[T, ...]`` -> sliding-window aggregate over the leading (snapshot) axis."""
    return _WindowFn.apply(p, window, AGG_ID.get(aggregation, 2))


def decay_scale(ts: torch.Tensor, t: int) -> torch.Tensor:
    """exp(-clamp(ts[:,t]-ts[:,t-1], 0, 10)) per row (TemporalGRUCell :509-514)."""
    lib = _lib.load()
    ts2 = _f32c(ts)
    out = torch.empty(ts2.shape[0], dtype=torch.float32, device=ts2.device)
    _lib.check(lib.tagan_decay_scale(_ptr(ts2), ts2.stride(0), t, _ptr(out), ts2.shape[0], _stream()), "tagan_decay_scale")
    CALLS["n"] += 1
    return out
