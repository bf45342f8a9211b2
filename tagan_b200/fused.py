"""Stage-level autograd Functions: the hot path with its element-wise / LayerNorm work fused into the
projections (``tagan_gemm_fused`` epilogues) or into single row passes (``csrc/fused_rows.cu``), and with every
residual-gradient accumulation done in place by the backward kernels instead of by autograd's ``add``.

One Function per stage of the reference's forward; each cites the reference lines it covers:

* ``geo_layer``        LN1 -> q/k/v Linear -> kernel (a) -> output_proj -> + identity -> LN2
                       (src/tagan/layers/geometric_attention.py:538-596), T snapshots batched
* ``evolution``        TemporalEvolutionLayer.forward (temporal_propagation.py:648-755): GRU scan with the
                       concatenations of :531-538 replaced by two-source GEMMs, gates / blend as GEMM epilogues,
                       the LN_out -> LN_h hand-over between steps as one kernel, then output_projection + x -> LN
* ``skip_connection``  TemporalSkipConnection.forward (:846-946)
* ``proj_ln``          ``layer_norm(dropout(output_proj(.)))`` of TemporalPropagation.forward (:1487-1500)
* ``tattn_layer``      AsymmetricTemporalAttention.forward (temporal_attention.py:985-1200) around kernel (b)
* ``mse``              mean of squares (the benchmark / trainer objective)

They apply when LayerNorm is on and dropout is inactive (eval() or p = 0: the parity setting, where the reference's
dropout layers are identities); the layer classes fall back to the op-by-op composition otherwise.  ``FUSED_GEMM``
False keeps the stage structure but composes plain ``tagan_gemm`` with the stand-alone kernels (used by the tests to
check the fused epilogues against the unfused arithmetic).
"""
import ctypes as C

import torch

from . import _lib, ops
from .gru import _ln_bwd, _ln_fwd, _off
from .ops import CALLS, _ptr, _stream, _timed, gemm, gemm_tn_colsum, workspace

FUSED_GEMM = True          # GEMM epilogue fusion (tests flip it to compare with the unfused composition)
EPI_FAST_MATH = False      # sigmoid / tanh of the GEMM epilogues through MUFU exp / reciprocal (rel. error ~1e-6)
# output_proj + residual + LayerNorm as a GEMM epilogue.  Correct (tests/test_gpu_fused.py) but OFF by default: measured on
# B200 (profiles/r02_fused_gemm_micro.jsonl) the 1.6M x 128 x 128 projection takes 1.56 ms with the epilogue against
# 0.83 ms for GEMM + LayerNorm kernel -- the four epilogue warps keep only ~8 KB of residual loads in flight per SM
# (ncu: long-scoreboard stalls 16.8 vs 3.8 per issue), far below what HBM latency needs.
FUSED_RES_LN = False

_F32 = torch.float32


def _e(*shape, dev):
    return torch.empty(*shape, dtype=_F32, device=dev)


def _al(*ts) -> bool:
    return all(t is None or (t.data_ptr() % 16 == 0) for t in ts)


def _use_fused_gemm() -> bool:
    return FUSED_GEMM and ops.FUSION and ops.GEMM_PRECISION in (1, 2, 3)


def _epi(mode, split=0, in0=None, ld_in0=0, in1=None, ld_in1=0, out0=None, ld_out0=0, out1=None, ld_out1=0,
         out2=None, ld_out2=0, gamma=None, beta=None, mean=None, rstd=None) -> _lib.Epilogue:
    def p(t):
        if t is None:
            return None
        return t.value if isinstance(t, C.c_void_p) else t.data_ptr()
    e = _lib.Epilogue()
    e.mode, e.split = mode, split
    e.in0, e.ld_in0, e.in1, e.ld_in1 = p(in0), ld_in0, p(in1), ld_in1
    e.out0, e.ld_out0, e.out1, e.ld_out1, e.out2, e.ld_out2 = p(out0), ld_out0, p(out1), ld_out1, p(out2), ld_out2
    e.gamma, e.beta, e.mean, e.rstd = p(gamma), p(beta), p(mean), p(rstd)
    return e


_EPI_STREAMS = {0: 1, 1: 3, 2: 4, 3: 4, 4: 5, 5: 0.5}     # [M,N]-sized tensors an epilogue reads + writes (algorithmic bytes)


def gemm_fused(op, m, n, k, a, lda, a2, lda2, k1, b, ldb, bias, epi, dev):
    lib = _lib.load()
    nbytes = lib.tagan_gemm_fused_workspace_bytes(op, m, n, k)
    ws = workspace(nbytes, dev) if nbytes else None
    with _timed("gemm", 4 * (m * k + n * k + m * n * _EPI_STREAMS[epi.mode])):
        rc = lib.tagan_gemm_fused(op, m, n, k, _ptr(a), lda, _ptr(a2), lda2, k1, _ptr(b), ldb, _ptr(bias), C.byref(epi),
                                  ops.GEMM_PRECISION | (8 if EPI_FAST_MATH else 0), _ptr(ws),
                                  ws.numel() if ws is not None else 0, _stream())
    _lib.check(rc, "tagan_gemm_fused")
    CALLS["n"] += 2


def linear_res_ln(x2, w, bias, res, gamma, beta, need_sum=True):
    """``LN(x2 . w^T + bias (+ res))`` -> (y, xsum, mean, rstd); xsum = the pre-LayerNorm sum (None unless need_sum).
    One GEMM with the residual add and the LayerNorm in its epilogue when the row fits one tile (N <= 128)."""
    lib = _lib.load()
    m, k = x2.shape
    n = w.shape[0]
    dev = x2.device
    ldx = x2.stride(0) if m > 1 else k
    y = _e(m, n, dev=dev)
    mean, rstd = _e(m, dev=dev), _e(m, dev=dev)
    xsum = _e(m, n, dev=dev) if need_sum else None
    if (_use_fused_gemm() and FUSED_RES_LN and n <= 128 and n % 4 == 0 and ldx % 4 == 0 and w.stride(0) % 4 == 0 and m > 0
            and _al(x2, w, bias, res, gamma, beta)):
        epi = _epi(_lib.EPI_RES_LN, in0=res, ld_in0=n, out0=y, ld_out0=n, out1=xsum, ld_out1=n, gamma=gamma, beta=beta,
                   mean=mean, rstd=rstd)
        gemm_fused(0, m, n, k, x2, ldx, None, 0, 0, w, w.stride(0), bias, epi, dev)
        return y, xsum, mean, rstd
    o = _e(m, n, dev=dev)
    gemm(0, m, n, k, x2, ldx, w, w.stride(0), bias, o, n)
    if res is None:
        _ln_fwd(lib, _ptr(o), n, gamma, beta, None, _ptr(y), n, _ptr(mean), _ptr(rstd), m, n)
        xsum = o if need_sum else None
    else:
        _ln_fwd_res(lib, o, res, gamma, beta, y, xsum, mean, rstd, m, n)
    return y, xsum, mean, rstd


def _ln_fwd_res(lib, o, res, gamma, beta, y, xsum, mean, rstd, m, n):
    rc = lib.tagan_layernorm_fwd(_ptr(o), n, _ptr(res), n, _ptr(gamma), _ptr(beta), None, _ptr(y), n, _ptr(xsum),
                                 _ptr(mean), _ptr(rstd), m, n, _stream())
    _lib.check(rc, "tagan_layernorm_fwd")
    CALLS["n"] += 1


def _ln_plain(lib, x2, gamma, beta):
    rows, cols = x2.shape
    y = torch.empty_like(x2)
    mean, rstd = _e(rows, dev=x2.device), _e(rows, dev=x2.device)
    _ln_fwd(lib, _ptr(x2), x2.stride(0) if rows > 1 else cols, gamma, beta, None, _ptr(y), cols, _ptr(mean), _ptr(rstd),
            rows, cols)
    return y, mean, rstd


def _ln_backward(lib, dy, xs, gamma, mean, rstd, dx, accumulate):
    """dx (+)= dLN(dy); returns (dgamma, dbeta)."""
    rows, cols = xs.shape
    dev = xs.device
    dg, db = _e(cols, dev=dev), _e(cols, dev=dev)
    _ln_bwd(lib, _ptr(dy), dy.stride(0) if rows > 1 else cols, _ptr(xs), xs.stride(0) if rows > 1 else cols, gamma, None,
            _ptr(mean), _ptr(rstd), _ptr(dx), cols, accumulate, _ptr(dg), _ptr(db), rows, cols, dev)
    return dg, db


def _rows2(x):
    x = x if x.dtype == _F32 else x.float()
    x = x.contiguous()
    return x.view(-1, x.shape[-1])


# ------------------------------------------------------------------------------------------
# geometric layer (rows a2-a4), T snapshots batched
# ------------------------------------------------------------------------------------------
class _GeoLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln1w, ln1b, wq, bq, wk, bk, wv, bv, wo, bo, ln2w, ln2b, metric_param, csrs, heads, metric, bf16=False):
        lib = _lib.load()
        rows = _rows2(x)
        r, hdim = rows.shape
        dev = rows.device
        t_steps = len(csrs)
        offs = [0]
        for c in csrs:                                                            # snapshots may have different node counts
            offs.append(offs[-1] + c.num_nodes)
        if r != offs[-1]:
            raise ValueError("geo_layer: the packed rows must be the concatenation of every snapshot's nodes")
        xn, mean1, rstd1 = _ln_plain(lib, rows, ln1w, ln1b)                        # :538-542
        w_qkv = torch.cat([wq, wk, wv], 0)
        b_qkv = torch.cat([bq, bk, bv], 0)
        es = 2 if bf16 else 4                                                     # bytes per stored q/k/v element
        if bf16:
            # bf16-STORAGE mode: the projection is computed as always (3xTF32, fp32 accumulate) and rounded to bf16 (RNE) by
            # the GEMM epilogue; kernel (a) gathers half the bytes and works in fp32
            qkv = torch.empty(r, 3 * hdim, dtype=torch.bfloat16, device=dev)
            if _use_fused_gemm() and r > 0 and _al(xn, w_qkv, b_qkv):
                gemm_fused(0, r, 3 * hdim, hdim, xn, hdim, None, 0, 0, w_qkv, hdim, b_qkv,
                           _epi(_lib.EPI_STORE_BF16, out0=qkv, ld_out0=3 * hdim), dev)
            else:
                tmp = _e(r, 3 * hdim, dev=dev)
                gemm(0, r, 3 * hdim, hdim, xn, hdim, w_qkv, hdim, b_qkv, tmp, 3 * hdim)
                qkv.copy_(tmp)
        else:
            qkv = _e(r, 3 * hdim, dev=dev)
            gemm(0, r, 3 * hdim, hdim, xn, hdim, w_qkv, hdim, b_qkv, qkv, 3 * hdim)     # :546-548
        ctxv = _e(r, hdim, dev=dev)
        lse = _e(r, heads, dev=dev)
        ld = 3 * hdim
        fwd_fn = lib.tagan_geo_attn_fwd_bf16 if bf16 else lib.tagan_geo_attn_fwd
        base, cbase, lbase = qkv.data_ptr(), ctxv.data_ptr(), lse.data_ptr()
        for t, csr in enumerate(csrs):
            off = base + offs[t] * ld * es
            q, k, v = (C.c_void_p(off + i * hdim * es) for i in range(3))
            ops.wait_csr(csr)
            with _timed("geo_attn_fwd"):
                rc = fwd_fn(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), csr.num_nodes, hdim, heads, metric,
                                            _ptr(metric_param), C.c_void_p(cbase + offs[t] * hdim * 4),
                                            C.c_void_p(lbase + offs[t] * heads * 4), None, _stream())
            _lib.check(rc, "tagan_geo_attn_fwd")
        CALLS["n"] += t_steps
        need = any(ctx.needs_input_grad)
        out, xsum, mean2, rstd2 = linear_res_ln(ctxv, wo, bo, rows, ln2w, ln2b, need_sum=need)   # :586-596
        ctx.save_for_backward(rows, xn, mean1, rstd1, qkv, ctxv, lse, xsum, mean2, rstd2, ln1w, w_qkv, wo, ln2w, metric_param)
        ctx.csrs, ctx.heads, ctx.metric, ctx.shape, ctx.offs, ctx.bf16 = csrs, heads, metric, x.shape, offs, bf16
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        rows, xn, mean1, rstd1, qkv, ctxv, lse, xsum, mean2, rstd2, ln1w, w_qkv, wo, ln2w, metric_param = ctx.saved_tensors
        csrs, heads, metric = ctx.csrs, ctx.heads, ctx.metric
        if any(c.rowptr_t is None for c in csrs):
            raise RuntimeError("CSR was built without its transpose; backward needs it")
        r, hdim = rows.shape
        dev = rows.device
        t_steps, offs = len(csrs), ctx.offs
        n = max(c.num_nodes for c in csrs)
        dout2 = _rows2(dout)
        d_o = _e(r, hdim, dev=dev)                                    # gradient of the pre-LN2 sum = of o AND of identity
        dln2w, dln2b = _ln_backward(lib, dout2, xsum, ln2w, mean2, rstd2, d_o, False)
        dwo = _e(hdim, hdim, dev=dev)
        dbo = gemm_tn_colsum(hdim, hdim, r, d_o, hdim, ctxv, hdim, dwo, hdim)
        dctx = _e(r, hdim, dev=dev)
        gemm(1, r, hdim, hdim, d_o, hdim, wo, hdim, None, dctx, hdim)
        dqkv = _e(r, 3 * hdim, dev=dev)
        delta = _e(n, heads, dev=dev)
        want_dp = metric_param is not None and metric in (7, 8)
        dp_ws = _e(n, heads, dev=dev) if want_dp else None
        dparam_t = _e(t_steps, heads, dev=dev) if want_dp else None
        ld = 3 * hdim
        es = 2 if ctx.bf16 else 4
        bwd_fn = lib.tagan_geo_attn_bwd_bf16 if ctx.bf16 else lib.tagan_geo_attn_bwd
        base, dbase = qkv.data_ptr(), dqkv.data_ptr()
        for t, csr in enumerate(csrs):
            off, doff = base + offs[t] * ld * es, dbase + offs[t] * ld * 4
            q, k, v = (C.c_void_p(off + i * hdim * es) for i in range(3))
            dq, dk, dv = (C.c_void_p(doff + i * hdim * 4) for i in range(3))
            with _timed("geo_attn_bwd"):
                rc = bwd_fn(q, k, v, ld, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.rowptr_t), _ptr(csr.row_t),
                                            csr.num_nodes, hdim, heads, metric, _ptr(metric_param),
                                            C.c_void_p(ctxv.data_ptr() + offs[t] * hdim * 4),
                                            C.c_void_p(lse.data_ptr() + offs[t] * heads * 4),
                                            C.c_void_p(dctx.data_ptr() + offs[t] * hdim * 4), dq, dk, dv, ld, _ptr(delta),
                                            _ptr(dp_ws), C.c_void_p(dparam_t.data_ptr() + t * heads * 4) if want_dp else None,
                                            _stream())
            _lib.check(rc, "tagan_geo_attn_bwd")
        CALLS["n"] += (3 if want_dp else 2) * t_steps
        dw_qkv = _e(3 * hdim, hdim, dev=dev)
        db_qkv = gemm_tn_colsum(3 * hdim, hdim, r, dqkv, ld, xn, hdim, dw_qkv, hdim)
        dxn = dctx                                                    # reuse: dctx is dead after the attention backward
        gemm(1, r, hdim, 3 * hdim, dqkv, ld, w_qkv, hdim, None, dxn, hdim)
        dln1w, dln1b = _ln_backward(lib, dxn, rows, ln1w, mean1, rstd1, d_o, True)    # d_o += dLN1: the residual sum, in place
        dparam = dparam_t.sum(0) if want_dp else None
        h = hdim
        return (d_o.view(ctx.shape), dln1w, dln1b, dw_qkv[:h], db_qkv[:h], dw_qkv[h:2 * h], db_qkv[h:2 * h],
                dw_qkv[2 * h:], db_qkv[2 * h:], dwo, dbo, dln2w, dln2b, dparam, None, None, None, None)


def geo_layer(ga, x, csrs):
    """``GeometricAttention`` module ``ga`` over stacked snapshots: x ``[T*N,H]`` / ``[T,N,H]`` + T CSRs, or the PACKED
    rows ``[sum N_t, H]`` of snapshots with different node counts (snapshot t's nodes are rows ``off[t]:off[t+1]``)."""
    return _GeoLayerFn.apply(x, ga.layer_norm1.weight, ga.layer_norm1.bias, ga.q_linear.weight, ga.q_linear.bias,
                             ga.k_linear.weight, ga.k_linear.bias, ga.v_linear.weight, ga.v_linear.bias,
                             ga.output_proj.weight, ga.output_proj.bias, ga.layer_norm2.weight, ga.layer_norm2.bias,
                             getattr(ga, "distance_param", None), list(csrs), ga.num_heads,
                             ops.METRIC_ID[ga.distance_metric], getattr(ga, "qkv_storage", "fp32") == "bf16")


# ------------------------------------------------------------------------------------------
# evolution layer (rows b3-b4)
# ------------------------------------------------------------------------------------------
def evolution_supported(din: int, hd: int) -> bool:
    """Two-source GEMMs need the split point on a 32-column k-block; float4 kernels need hidden % 4 == 0."""
    return _use_fused_gemm() and din % 32 == 0 and hd % 4 == 0 and hd <= 512 and din <= 512


class _EvolutionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x3, ts, lnx_w, lnx_b, lnh_w, lnh_b, lno_w, lno_b, w_r, b_r, w_z, b_z, w_c, b_c, w_o, b_o, ln_w, ln_b,
                residual: bool):
        lib = _lib.load()
        x3 = x3.contiguous().float()
        t_steps, n, din = x3.shape
        hd = w_r.shape[0]
        rows = t_steps * n
        dev = x3.device
        xrows = x3.view(rows, din)
        xhat, mean_x, rstd_x = _ln_plain(lib, xrows, lnx_w, lnx_b)               # LN_x of every step at once (:499-501)
        w_rz = torch.cat([w_r, w_z], 0).contiguous()                              # [2H, in+H], columns [x^ | h^] as in the reference
        b_rz = torch.cat([b_r, b_z], 0).contiguous()
        w_c = w_c.contiguous()
        kk = din + hd
        hhat = _e(t_steps, n, hd, dev=dev)
        r, z, rs, cand, s = (_e(t_steps, n, hd, dev=dev) for _ in range(5))
        mean_o, rstd_o, mean_h, rstd_h = (_e(t_steps, n, dev=dev) for _ in range(4))
        ts_c = ts.contiguous().float() if ts is not None else None
        decay = _e(t_steps, n, dev=dev) if ts_c is not None else None
        g_rz, g_c = _e(n, 2 * hd, dev=dev), _e(n, hd, dev=dev)                     # per-step pre-activations (scratch)
        nh = n * hd
        hhat[0].zero_()                                                           # h is None -> zeros, no LayerNorm (:503-504)
        for t in range(t_steps):
            x_t, hh_t = _off(xhat, t * n * din), _off(hhat, t * nh)
            # [r_pre | z_pre] = [x^ | h^] . W_rz^T + b: the concatenation of :531 as a two-source A operand
            gemm_fused(0, n, 2 * hd, kk, x_t, din, hh_t, hd, din, w_rz, kk, b_rz, _epi(_lib.EPI_STORE, out0=g_rz, ld_out0=2 * hd), dev)
            _lib.check(lib.tagan_gates_fwd(_ptr(g_rz), 2 * hd, hh_t, hd, _off(r, t * nh), _off(z, t * nh), _off(rs, t * nh), hd,
                                           n, hd, _stream()), "tagan_gates_fwd")                       # :531-535
            gemm_fused(0, n, hd, kk, x_t, din, _off(rs, t * nh), hd, din, w_c, kk, b_c, _epi(_lib.EPI_STORE, out0=g_c, ld_out0=hd), dev)
            # cand = tanh(.); hn = (1-z) h^ + z cand (:538-542); s_t = LN_out(hn) (:545-546);
            # h^_{t+1} = LN_h(s_t) * exp(-clamp(dt, 0, 10)) (:505-514) -- one pass, hn never stored
            last = t == t_steps - 1
            rc = lib.tagan_gru_blend_ln_fwd(_ptr(g_c), hd, _off(z, t * nh), hh_t, hd, _off(cand, t * nh), _ptr(lno_w), _ptr(lno_b),
                                            _ptr(lnh_w), _ptr(lnh_b), _ptr(ts_c) if not last else None,
                                            ts_c.stride(0) if ts_c is not None else 0, t + 1, _off(s, t * nh), hd,
                                            None if last else _off(hhat, (t + 1) * nh), hd, _off(mean_o, t * n), _off(rstd_o, t * n),
                                            None if last else _off(mean_h, (t + 1) * n), None if last else _off(rstd_h, (t + 1) * n),
                                            None if (last or decay is None) else _off(decay, (t + 1) * n), n, hd, _stream())
            _lib.check(rc, "tagan_gru_blend_ln_fwd")
            CALLS["n"] += 2
        res = xrows if (residual and din == w_o.shape[0]) else None
        need = any(ctx.needs_input_grad)
        e, xsum, mean_e, rstd_e = linear_res_ln(s.view(rows, hd), w_o, b_o, res, ln_w, ln_b, need_sum=need)   # :738-753
        ctx.save_for_backward(xrows, xhat, mean_x, rstd_x, hhat, r, z, rs, cand, s, mean_o, rstd_o, mean_h, rstd_h, decay,
                              lnx_w, lnh_w, lno_w, lno_b, w_rz, w_c, w_o, ln_w, xsum, mean_e, rstd_e)
        ctx.dims, ctx.has_res = (t_steps, n, din, hd), res is not None
        return e.view(t_steps, n, w_o.shape[0])

    @staticmethod
    def backward(ctx, de):
        lib = _lib.load()
        (xrows, xhat, mean_x, rstd_x, hhat, r, z, rs, cand, s, mean_o, rstd_o, mean_h, rstd_h, decay, lnx_w, lnh_w,
         lno_w, lno_b, w_rz, w_c, w_o, ln_w, xsum, mean_e, rstd_e) = ctx.saved_tensors
        t_steps, n, din, hd = ctx.dims
        rows, nh, kk = t_steps * n, n * hd, din + hd
        dev = xrows.device
        hout = w_o.shape[0]
        de2 = _rows2(de)
        d_o = _e(rows, hout, dev=dev)
        dln_w, dln_b = _ln_backward(lib, de2, xsum, ln_w, mean_e, rstd_e, d_o, False)
        dw_o = _e(hout, hd, dev=dev)
        db_o = gemm_tn_colsum(hout, hd, rows, d_o, hout, s, hd, dw_o, hd)
        ds = _e(rows, hd, dev=dev)                                                # external gradient of every s_t
        gemm(1, rows, hd, hout, d_o, hout, w_o, hd, None, ds, hd)
        dg = _e(rows, 3 * hd, dev=dev)                                            # [d r_pre | d z_pre | d cand_pre]
        dhh_a, dhh_b = _e(n, hd, dev=dev), _e(n, hd, dev=dev)                     # gradient of h^: next step's (in) / this step's (out)
        drs = _e(n, hd, dev=dev)
        daff = torch.zeros(4, hd, dtype=_F32, device=dev)                         # d gamma_o, d beta_o, d gamma_h, d beta_h
        ws_pair = workspace(lib.tagan_ln_pair_bwd_workspace_bytes(n, hd), dev)
        w_c_h = C.c_void_p(w_c.data_ptr() + din * 4)                              # W_c[:, in:]  ([H, H], ld in+H)
        w_rz_h = C.c_void_p(w_rz.data_ptr() + din * 4)                            # W_rz[:, in:] ([2H, H], ld in+H)
        for t in range(t_steps - 1, -1, -1):
            last = t == t_steps - 1
            dg_t = t * n * 3 * hd
            # LN_h' of step t+1, LN_out' and the blend of step t in one pass: d z_pre, d cand_pre into their slices, d h^_t
            rc = lib.tagan_gru_blend_ln_bwd(_off(ds, t * nh), hd, None if last else _ptr(dhh_a), hd, _off(cand, t * nh),
                                            _off(z, t * nh), _off(hhat, t * nh), hd, _off(dg, dg_t + hd), _off(dg, dg_t + 2 * hd),
                                            3 * hd, _ptr(dhh_b), hd, _ptr(lno_w), _ptr(lno_b), _ptr(lnh_w), _off(mean_o, t * n),
                                            _off(rstd_o, t * n), None if last else _off(mean_h, (t + 1) * n),
                                            None if last else _off(rstd_h, (t + 1) * n),
                                            None if (last or decay is None) else _off(decay, (t + 1) * n), _ptr(daff), 1,
                                            _ptr(ws_pair), ws_pair.numel(), n, hd, _stream())
            _lib.check(rc, "tagan_gru_blend_ln_bwd")
            CALLS["n"] += 2
            if t > 0:
                # d(rs) = d cand_pre . W_c[:, in:];  d r_pre = d(rs) h^ r(1-r);  dhh += d(rs) r;  dhh += [d r_pre | d z_pre] . W_rz[:, in:]
                gemm(1, n, hd, hd, _off(dg, dg_t + 2 * hd), 3 * hd, w_c_h, kk, None, drs, hd)
                _lib.check(lib.tagan_gru_reset_bwd(_ptr(drs), _off(r, t * nh), _off(hhat, t * nh), hd, _off(dg, dg_t), 3 * hd,
                                                   _ptr(dhh_b), hd, n, hd, _stream()), "tagan_gru_reset_bwd")
                CALLS["n"] += 1
                gemm(1, n, hd, 2 * hd, _off(dg, dg_t), 3 * hd, w_rz_h, kk, None, dhh_b, hd, accumulate=True)
            else:
                dg.view(t_steps, n, 3 * hd)[0, :, :hd].zero_()                    # h^_0 = 0: no gradient through r
            dhh_a, dhh_b = dhh_b, dhh_a
        # weight gradients over all T*N rows; step 0 contributes zeros through h^ = r*h^ = 0
        w_x = torch.cat([w_rz[:, :din], w_c[:, :din]], 0).contiguous()             # [3H, in]
        dxhat = _e(rows, din, dev=dev)
        gemm(1, rows, din, 3 * hd, dg, 3 * hd, w_x, din, None, dxhat, din)
        dw_rz = _e(2 * hd, kk, dev=dev)
        dw_c = _e(hd, kk, dev=dev)
        dw_x = _e(3 * hd, din, dev=dev)
        db = gemm_tn_colsum(3 * hd, din, rows, dg, 3 * hd, xhat, din, dw_x, din)
        dw_rz[:, :din].copy_(dw_x[:2 * hd])
        dw_c[:, :din].copy_(dw_x[2 * hd:])
        gemm(2, 2 * hd, hd, rows, dg, 3 * hd, hhat, hd, None, C.c_void_p(dw_rz.data_ptr() + din * 4), kk)
        gemm(2, hd, hd, rows, _off(dg, 2 * hd), 3 * hd, rs, hd, None, C.c_void_p(dw_c.data_ptr() + din * 4), kk)
        if ctx.has_res:
            dx = d_o                                                              # identity branch; LN_x' accumulates onto it
            dlnx_w, dlnx_b = _ln_backward(lib, dxhat, xrows, lnx_w, mean_x, rstd_x, dx, True)
        else:
            dx = _e(rows, din, dev=dev)
            dlnx_w, dlnx_b = _ln_backward(lib, dxhat, xrows, lnx_w, mean_x, rstd_x, dx, False)
        return (dx.view(t_steps, n, din), None, dlnx_w, dlnx_b, daff[2], daff[3], daff[0], daff[1],
                dw_rz[:hd], db[:hd], dw_rz[hd:], db[hd:2 * hd], dw_c, db[2 * hd:], dw_o, db_o, dln_w, dln_b, None)


def evolution(layer, x3, ts):
    """``TemporalEvolutionLayer`` ``layer`` (unidirectional, LayerNorm on) on x3 ``[T,N,in]`` -> ``[T,N,hidden]``."""
    cell = layer.forward_cell
    return _EvolutionFn.apply(
        x3, ts if layer.time_aware else None, cell.layer_norm_x.weight, cell.layer_norm_x.bias, cell.layer_norm_h.weight,
        cell.layer_norm_h.bias, cell.layer_norm_out.weight, cell.layer_norm_out.bias, cell.reset_gate.weight,
        cell.reset_gate.bias, cell.update_gate.weight, cell.update_gate.bias, cell.candidate.weight, cell.candidate.bias,
        layer.output_projection.weight, layer.output_projection.bias, layer.layer_norm.weight, layer.layer_norm.bias,
        bool(layer.residual))


# ------------------------------------------------------------------------------------------
# skip connection (row b5)
# ------------------------------------------------------------------------------------------
def skip_supported(mod, t_steps: int, n: int) -> bool:
    return (ops.FUSION and mod.use_layer_norm and mod.apply_activation and mod.aggregation in ("mean", "sum")
            and 1 <= mod.window_size <= 4 and mod.hidden_dim % 4 == 0 and mod.hidden_dim <= 512 and mod.input_dim <= 512)


class _SkipFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e3, w_in, b_in, ln1_w, ln1_b, w_out, b_out, ln2_w, ln2_b, window: int, agg: int, residual: bool):
        lib = _lib.load()
        e3 = e3.contiguous().float()
        t_steps, n, din = e3.shape
        hs = w_in.shape[0]
        rows = t_steps * n
        dev = e3.device
        erows = e3.view(rows, din)
        a = _e(rows, hs, dev=dev)
        gemm(0, rows, hs, din, erows, din, w_in, w_in.stride(0), b_in, a, hs)                       # :866
        p = _e(rows, hs, dev=dev)
        mean1, rstd1 = _e(rows, dev=dev), _e(rows, dev=dev)
        _lib.check(lib.tagan_gelu_ln_fwd(_ptr(a), hs, _ptr(ln1_w), _ptr(ln1_b), _ptr(p), hs, _ptr(mean1), _ptr(rstd1), rows, hs,
                                         _stream()), "tagan_gelu_ln_fwd")                            # :869-877
        gg = _e(rows, hs, dev=dev)
        _lib.check(lib.tagan_window_gelu_fwd(_ptr(p), _ptr(gg), t_steps, n * hs, window, agg, _stream()),
                   "tagan_window_gelu_fwd")                                                           # :880-894, :929-933
        CALLS["n"] += 2
        need = any(ctx.needs_input_grad)
        y, xsum, mean2, rstd2 = linear_res_ln(gg, w_out, b_out, erows if residual else None, ln2_w, ln2_b, need_sum=need)
        ctx.save_for_backward(erows, a, p, gg, xsum, mean1, rstd1, mean2, rstd2, w_in, ln1_w, w_out, ln2_w)
        ctx.cfg = (t_steps, n, din, hs, window, agg, residual)
        return y.view(t_steps, n, din)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        erows, a, p, gg, xsum, mean1, rstd1, mean2, rstd2, w_in, ln1_w, w_out, ln2_w = ctx.saved_tensors
        t_steps, n, din, hs, window, agg, residual = ctx.cfg
        rows = t_steps * n
        dev = erows.device
        dy2 = _rows2(dy)
        d_o = _e(rows, din, dev=dev)
        dln2_w, dln2_b = _ln_backward(lib, dy2, xsum, ln2_w, mean2, rstd2, d_o, False)
        dw_out = _e(din, hs, dev=dev)
        db_out = gemm_tn_colsum(din, hs, rows, d_o, din, gg, hs, dw_out, hs)
        dgg = _e(rows, hs, dev=dev)
        gemm(1, rows, hs, din, d_o, din, w_out, w_out.stride(0), None, dgg, hs)
        dp = gg                                                                   # gg is dead: reuse its storage
        _lib.check(lib.tagan_window_gelu_bwd(_ptr(dgg), _ptr(p), _ptr(dp), t_steps, n * hs, window, agg, _stream()),
                   "tagan_window_gelu_bwd")
        da = dgg
        daff = _e(2, hs, dev=dev)
        ws = workspace(lib.tagan_gelu_ln_bwd_workspace_bytes(rows, hs), dev)
        _lib.check(lib.tagan_gelu_ln_bwd(_ptr(dp), hs, _ptr(a), hs, _ptr(ln1_w), _ptr(mean1), _ptr(rstd1), _ptr(da), hs,
                                         _ptr(daff), _ptr(ws), ws.numel(), rows, hs, _stream()), "tagan_gelu_ln_bwd")
        CALLS["n"] += 3
        dw_in = _e(hs, din, dev=dev)
        db_in = gemm_tn_colsum(hs, din, rows, da, hs, erows, din, dw_in, din)
        if residual:
            de = d_o
            gemm(1, rows, din, hs, da, hs, w_in, w_in.stride(0), None, de, din, accumulate=True)
        else:
            de = _e(rows, din, dev=dev)
            gemm(1, rows, din, hs, da, hs, w_in, w_in.stride(0), None, de, din)
        return (de.view(t_steps, n, din), dw_in, db_in, daff[0], daff[1], dw_out, db_out, dln2_w, dln2_b, None, None, None)


def skip_connection(mod, e3):
    return _SkipFn.apply(e3, mod.input_proj.weight, mod.input_proj.bias, mod.layer_norm1.weight, mod.layer_norm1.bias,
                         mod.output_proj.weight, mod.output_proj.bias, mod.layer_norm2.weight, mod.layer_norm2.bias,
                         int(mod.window_size), ops.AGG_ID[mod.aggregation], bool(mod.residual))


# ------------------------------------------------------------------------------------------
# projection + LayerNorm (TemporalPropagation tail, :1487-1500)
# ------------------------------------------------------------------------------------------
class _ProjLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, ln_w, ln_b):
        rows = _rows2(x)
        y, xsum, mean, rstd = linear_res_ln(rows, w, b, None, ln_w, ln_b, need_sum=any(ctx.needs_input_grad))
        ctx.save_for_backward(rows, w, ln_w, xsum, mean, rstd)
        ctx.shape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        rows, w, ln_w, xsum, mean, rstd = ctx.saved_tensors
        m, k = rows.shape
        n = w.shape[0]
        dev = rows.device
        d_o = _e(m, n, dev=dev)
        dln_w, dln_b = _ln_backward(lib, _rows2(dy), xsum, ln_w, mean, rstd, d_o, False)
        dw = _e(n, k, dev=dev)
        db = gemm_tn_colsum(n, k, m, d_o, n, rows, k, dw, k)
        dx = _e(m, k, dev=dev)
        gemm(1, m, k, n, d_o, n, w, w.stride(0), None, dx, k)
        return dx.view(ctx.shape), dw, db, dln_w, dln_b


def proj_ln(x, linear, ln):
    return _ProjLNFn.apply(x, linear.weight, linear.bias, ln.weight, ln.bias)


# ------------------------------------------------------------------------------------------
# temporal attention layer (rows b1-b2) around kernel (b)
# ------------------------------------------------------------------------------------------
class _TAttnLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, ln1w, ln1b, wq, bq, wk, bk, wv, bv, wo, bo, ln2w, ln2b, tmask, batch, t, heads, time_major):
        lib = _lib.load()
        rows = _rows2(x)
        r, hdim = rows.shape
        dev = rows.device
        assert r == batch * t
        xn, mean1, rstd1 = _ln_plain(lib, rows, ln1w, ln1b)                        # temporal_attention.py:985-990
        w_qkv = torch.cat([wq, wk, wv], 0)
        b_qkv = torch.cat([bq, bk, bv], 0)
        qkv = _e(r, 3 * hdim, dev=dev)
        gemm(0, r, 3 * hdim, hdim, xn, hdim, w_qkv, hdim, b_qkv, qkv, 3 * hdim)
        ctxv = _e(r, hdim, dev=dev)
        lse = _e(batch, heads, t, dev=dev)
        bias_c = bias_t = None
        bstride = 0
        if bias is not None:
            bias_c = ops._f32c(bias).contiguous()
            bias_t = bias_c.transpose(-1, -2).contiguous()
            bstride = heads * t * t if bias_c.dim() == 4 else 0
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        ld = 3 * hdim
        base = qkv.data_ptr()
        q, k, v = (C.c_void_p(base + i * hdim * 4) for i in range(3))
        with _timed("tattn_fwd"):
            rc = lib.tagan_tattn_fwd(q, k, v, ld, batch, t, hdim, heads, int(time_major), _ptr(bias_c), _ptr(bias_t), bstride,
                                     _ptr(tmask.ts), tmask.flags, tmask.band, _ptr(tmask.allones_flag), _ptr(m), mb, mh,
                                     _ptr(ctxv), _ptr(lse), None, _stream())
        _lib.check(rc, "tagan_tattn_fwd")
        CALLS["n"] += 1
        need = any(ctx.needs_input_grad)
        out, xsum, mean2, rstd2 = linear_res_ln(ctxv, wo, bo, rows, ln2w, ln2b, need_sum=need)     # :1186-1200
        ctx.save_for_backward(rows, xn, mean1, rstd1, qkv, ctxv, lse, xsum, mean2, rstd2, ln1w, w_qkv, wo, ln2w, bias_c, bias_t)
        ctx.tmask, ctx.dims, ctx.bstride = tmask, (batch, t, heads, time_major), bstride
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.shape = x.shape
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (rows, xn, mean1, rstd1, qkv, ctxv, lse, xsum, mean2, rstd2, ln1w, w_qkv, wo, ln2w, bias_c, bias_t) = ctx.saved_tensors
        tmask = ctx.tmask
        batch, t, heads, time_major = ctx.dims
        r, hdim = rows.shape
        dev = rows.device
        d_o = _e(r, hdim, dev=dev)
        dln2w, dln2b = _ln_backward(lib, _rows2(dout), xsum, ln2w, mean2, rstd2, d_o, False)
        dwo = _e(hdim, hdim, dev=dev)
        dbo = gemm_tn_colsum(hdim, hdim, r, d_o, hdim, ctxv, hdim, dwo, hdim)
        dctx = _e(r, hdim, dev=dev)
        gemm(1, r, hdim, hdim, d_o, hdim, wo, hdim, None, dctx, hdim)
        ld = 3 * hdim
        dqkv = _e(r, ld, dev=dev)
        dbias = ws = None
        if bias_c is not None and ctx.needs_input_grad[1]:
            dbias = torch.empty_like(bias_c)
            if ctx.bstride == 0:
                ws = workspace(lib.tagan_tattn_bwd_workspace_bytes(batch, t, heads), dev)
        m = tmask.mask
        mb, mh = (m.shape[0], m.shape[1]) if m is not None else (1, 1)
        base, dbase = qkv.data_ptr(), dqkv.data_ptr()
        q, k, v = (C.c_void_p(base + i * hdim * 4) for i in range(3))
        dq, dk, dv = (C.c_void_p(dbase + i * hdim * 4) for i in range(3))
        with _timed("tattn_bwd"):
            rc = lib.tagan_tattn_bwd(q, k, v, ld, batch, t, hdim, heads, int(time_major), _ptr(bias_c), _ptr(bias_t),
                                     ctx.bstride, _ptr(tmask.ts), tmask.flags, tmask.band, _ptr(tmask.allones_flag),
                                     _ptr(m), mb, mh, _ptr(ctxv), _ptr(lse), _ptr(dctx), dq, dk, dv, ld,
                                     _ptr(dbias), _ptr(ws), ws.numel() if ws is not None else 0, _stream())
        _lib.check(rc, "tagan_tattn_bwd")
        CALLS["n"] += 2 if ws is not None else 1
        dw_qkv = _e(ld, hdim, dev=dev)
        db_qkv = gemm_tn_colsum(ld, hdim, r, dqkv, ld, xn, hdim, dw_qkv, hdim)
        dxn = dctx
        gemm(1, r, hdim, ld, dqkv, ld, w_qkv, hdim, None, dxn, hdim)
        dln1w, dln1b = _ln_backward(lib, dxn, rows, ln1w, mean1, rstd1, d_o, True)
        if dbias is not None:
            dbias = dbias.view(ctx.bias_shape)
        h = hdim
        return (d_o.view(ctx.shape), dbias, dln1w, dln1b, dw_qkv[:h], db_qkv[:h], dw_qkv[h:2 * h], db_qkv[h:2 * h],
                dw_qkv[2 * h:], db_qkv[2 * h:], dwo, dbo, dln2w, dln2b, None, None, None, None, None)


def tattn_layer(mod, x_rows, bias, tmask, batch, t, time_major):
    """``AsymmetricTemporalAttention`` module ``mod`` on physical rows ``[B*T,H]`` (``b*T+t`` or ``t*B+b`` order)."""
    return _TAttnLayerFn.apply(x_rows, bias, mod.layer_norm1.weight, mod.layer_norm1.bias, mod.q_linear.weight,
                               mod.q_linear.bias, mod.k_linear.weight, mod.k_linear.bias, mod.v_linear.weight,
                               mod.v_linear.bias, mod.output_proj.weight, mod.output_proj.bias, mod.layer_norm2.weight,
                               mod.layer_norm2.bias, tmask, batch, t, mod.num_heads, time_major)


# ------------------------------------------------------------------------------------------
# mean of squares
# ------------------------------------------------------------------------------------------
class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        xc = x if (x.is_contiguous() and x.dtype == _F32) else x.contiguous().float()
        loss = _e((), dev=xc.device)
        ws = workspace(lib.tagan_mse_workspace_bytes(), xc.device)
        _lib.check(lib.tagan_mse_fwd(_ptr(xc), xc.numel(), _ptr(loss), _ptr(ws), ws.numel(), _stream()), "tagan_mse_fwd")
        CALLS["n"] += 2
        ctx.save_for_backward(xc)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lib = _lib.load()
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        g = dloss.contiguous().float()
        _lib.check(lib.tagan_mse_bwd(_ptr(xc), xc.numel(), _ptr(g), _ptr(dx), _stream()), "tagan_mse_bwd")
        CALLS["n"] += 1
        return dx


def mean_square(x: torch.Tensor) -> torch.Tensor:
    """``x.square().mean()`` in one read of x (and one read + one write for the gradient); x must be dense (any
    permutation of a contiguous tensor gives the same value: pass the storage-order view)."""
    return _MSEFn.apply(x)
