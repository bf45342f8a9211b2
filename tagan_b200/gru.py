"""Fused GRU scan over the snapshot axis (hot-path rows b3-b4).

One autograd Function for the whole recurrence of ``TemporalEvolutionLayer`` (reference
src/tagan/layers/temporal_propagation.py:675-688 calling ``TemporalGRUCell.forward`` :475-551 per step):

* ``LN_x`` and the input halves of the three gate Linears do not depend on the state, so they run ONCE
  for all T steps: ``G[T*N,3H] = LN_x(X) . [W_r|W_z|W_c][:, :in]^T + b`` (one large tcgen05 GEMM);
* per step only the state halves are added with accumulate-GEMMs into column slices of ``G``
  (``G[:, :2H] += h^ . W_rz_h^T``, ``G[:, 2H:] += (r*h^) . W_c_h^T``), followed by the fused gate / blend /
  LayerNorm kernels.  No concatenation, no stack, no zeros tensors are materialised;
* backward walks the steps in reverse with the same kernels, carrying dL/dh in place, and forms every
  weight gradient with three large reductions over all T*N rows at the end.
"""
import ctypes as C

import torch

from . import _lib
from .ops import CALLS, _ptr, _stream, colsum, gemm, gemm_tn_colsum, workspace


def _off(t: torch.Tensor, elems: int):
    return C.c_void_p(t.data_ptr() + elems * 4)


def _ln_fwd(lib, x, ldx, gamma, beta, rowscale, y, ldy, mean, rstd, rows, cols):
    rc = lib.tagan_layernorm_fwd(x, ldx, None, 0, _ptr(gamma), _ptr(beta), rowscale, y, ldy, None, mean, rstd, rows, cols,
                                 _stream())
    _lib.check(rc, "tagan_layernorm_fwd")
    CALLS["n"] += 1


def _ln_bwd(lib, dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx, accumulate, dgamma, dbeta, rows, cols, dev):
    ws = None
    if gamma is not None:
        ws = workspace(lib.tagan_layernorm_bwd_workspace_bytes(rows, cols), dev)
    rc = lib.tagan_layernorm_bwd(dy, lddy, xsum, ldx, _ptr(gamma), rowscale, mean, rstd, dx, lddx, int(accumulate),
                                 dgamma, dbeta, _ptr(ws), ws.numel() if ws is not None else 0, rows, cols, _stream())
    _lib.check(rc, "tagan_layernorm_bwd")
    CALLS["n"] += 3 if gamma is not None else 1


class _GRUScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x3, ts, lnx_w, lnx_b, lnh_w, lnh_b, lno_w, lno_b, w_r, b_r, w_z, b_z, w_c, b_c, reverse: bool):
        lib = _lib.load()
        x3 = x3.contiguous().float()
        t_steps, n, din = x3.shape
        hd = w_r.shape[0]
        rows = t_steps * n
        dev = x3.device
        f32 = dict(dtype=torch.float32, device=dev)
        use_ln = lnx_w is not None
        xhat = torch.empty(rows, din, **f32) if use_ln else x3.view(rows, din)
        mean_x = torch.empty(rows, **f32) if use_ln else None
        rstd_x = torch.empty(rows, **f32) if use_ln else None
        if use_ln:
            _ln_fwd(lib, _ptr(x3), din, lnx_w, lnx_b, None, _ptr(xhat), din, _ptr(mean_x), _ptr(rstd_x), rows, din)
        w_x = torch.cat([w_r[:, :din], w_z[:, :din], w_c[:, :din]], 0).contiguous()       # [3H, in]
        b_x = torch.cat([b_r, b_z, b_c], 0).contiguous()
        w_h_rz = torch.cat([w_r[:, din:], w_z[:, din:]], 0).contiguous()                   # [2H, H]
        w_h_c = w_c[:, din:].contiguous()                                                  # [H, H]
        g = torch.empty(rows, 3 * hd, **f32)
        gemm(0, rows, 3 * hd, din, xhat, din, w_x, din, b_x, g, 3 * hd)
        hhat = torch.empty(t_steps, n, hd, **f32)
        r = torch.empty(t_steps, n, hd, **f32)
        z = torch.empty_like(r)
        rs = torch.empty_like(r)
        cand = torch.empty_like(r)
        hn = torch.empty_like(r)
        s = torch.empty_like(r)
        mean_h = torch.empty(t_steps, n, **f32) if use_ln else None
        rstd_h = torch.empty(t_steps, n, **f32) if use_ln else None
        mean_o = torch.empty(t_steps, n, **f32) if use_ln else None
        rstd_o = torch.empty(t_steps, n, **f32) if use_ln else None
        decay = torch.empty(t_steps, n, **f32) if ts is not None else None
        ts_c = ts.contiguous().float() if ts is not None else None
        order = list(range(t_steps - 1, -1, -1)) if reverse else list(range(t_steps))
        nh = n * hd
        prev = None
        for idx, t in enumerate(order):
            g_t = _off(g, t * n * 3 * hd)
            hh_t = _off(hhat, t * nh)
            if idx == 0:
                hhat[t].zero_()                                           # h is None -> zeros, no LayerNorm (:503-504)
            else:
                dec = None
                if ts_c is not None:                                      # exp(-clamp(dt,0,10)) (:509-514)
                    tt = t + 1 if reverse else t
                    _lib.check(lib.tagan_decay_scale(_ptr(ts_c), ts_c.stride(0), tt, _off(decay, t * n), n, _stream()),
                               "tagan_decay_scale")
                    CALLS["n"] += 1
                    dec = _off(decay, t * n)
                _ln_fwd(lib, _off(s, prev * nh), hd, lnh_w, lnh_b, dec, hh_t, hd,
                        _off(mean_h, t * n) if use_ln else None, _off(rstd_h, t * n) if use_ln else None, n, hd)
                gemm(0, n, 2 * hd, hd, hh_t, hd, w_h_rz, hd, None, g_t, 3 * hd, accumulate=True)
            _lib.check(lib.tagan_gates_fwd(g_t, 3 * hd, hh_t, hd, _off(r, t * nh), _off(z, t * nh), _off(rs, t * nh), hd,
                                           n, hd, _stream()), "tagan_gates_fwd")
            gc_t = _off(g, t * n * 3 * hd + 2 * hd)
            if idx > 0:
                gemm(0, n, hd, hd, _off(rs, t * nh), hd, w_h_c, hd, None, gc_t, 3 * hd, accumulate=True)
            _lib.check(lib.tagan_blend_fwd(gc_t, 3 * hd, _off(z, t * nh), hh_t, hd, _off(cand, t * nh), _off(hn, t * nh),
                                           0, n, hd, _stream()), "tagan_blend_fwd")
            CALLS["n"] += 2
            _ln_fwd(lib, _off(hn, t * nh), hd, lno_w, lno_b, None, _off(s, t * nh), hd,
                    _off(mean_o, t * n) if use_ln else None, _off(rstd_o, t * n) if use_ln else None, n, hd)
            prev = t
        ctx.save_for_backward(x3, xhat, mean_x, rstd_x, hhat, r, z, rs, cand, hn, s, mean_h, rstd_h, mean_o, rstd_o,
                              decay, lnx_w, lnh_w, lno_w, w_x, w_h_rz, w_h_c)
        ctx.order, ctx.dims, ctx.use_ln = order, (t_steps, n, din, hd), use_ln
        return s

    @staticmethod
    def backward(ctx, ds):
        lib = _lib.load()
        (x3, xhat, mean_x, rstd_x, hhat, r, z, rs, cand, hn, s, mean_h, rstd_h, mean_o, rstd_o, decay, lnx_w, lnh_w,
         lno_w, w_x, w_h_rz, w_h_c) = ctx.saved_tensors
        t_steps, n, din, hd = ctx.dims
        order, use_ln = ctx.order, ctx.use_ln
        rows, nh = t_steps * n, n * hd
        dev = ds.device
        f32 = dict(dtype=torch.float32, device=dev)
        dsa = ds.contiguous().float().clone()                    # dL/dS, the recurrent part is accumulated in place
        dg = torch.empty(rows, 3 * hd, **f32)
        dhn = torch.empty(n, hd, **f32)
        dz = torch.empty(n, hd, **f32)
        dhh = torch.empty(n, hd, **f32)
        drs = torch.zeros(n, hd, **f32)
        dlno = torch.zeros(2, t_steps, hd, **f32) if use_ln else None
        dlnh = torch.zeros(2, t_steps, hd, **f32) if use_ln else None
        for idx in range(t_steps - 1, -1, -1):
            t = order[idx]
            dg_t = _off(dg, t * n * 3 * hd)
            dgc_t = _off(dg, t * n * 3 * hd + 2 * hd)
            hh_t = _off(hhat, t * nh)
            _ln_bwd(lib, _off(dsa, t * nh), hd, _off(hn, t * nh), hd, lno_w, None,
                    _off(mean_o, t * n) if use_ln else None, _off(rstd_o, t * n) if use_ln else None, _ptr(dhn), hd, False,
                    _off(dlno, t * hd) if use_ln else None, _off(dlno, (t_steps + t) * hd) if use_ln else None, n, hd, dev)
            _lib.check(lib.tagan_blend_bwd(_ptr(dhn), _off(z, t * nh), _off(cand, t * nh), hh_t, hd, dgc_t, 3 * hd, _ptr(dz),
                                           _ptr(dhh), hd, 0, 0, n, hd, _stream()), "tagan_blend_bwd")
            if idx > 0:                                          # d(r*h^) = dcand_pre . W_c_h
                gemm(1, n, hd, hd, dgc_t, 3 * hd, w_h_c, hd, None, drs, hd)
            elif t_steps > 1:
                drs.zero_()
            _lib.check(lib.tagan_gates_bwd(_ptr(drs), hd, _ptr(dz), _off(r, t * nh), _off(z, t * nh), hh_t, hd, dg_t, 3 * hd,
                                           _ptr(dhh), hd, 1, n, hd, _stream()), "tagan_gates_bwd")
            CALLS["n"] += 2
            if idx > 0:
                prev = order[idx - 1]
                gemm(1, n, hd, 2 * hd, dg_t, 3 * hd, w_h_rz, hd, None, dhh, hd, accumulate=True)
                _ln_bwd(lib, _ptr(dhh), hd, _off(s, prev * nh), hd, lnh_w, _off(decay, t * n) if decay is not None else None,
                        _off(mean_h, t * n) if use_ln else None, _off(rstd_h, t * n) if use_ln else None,
                        _off(dsa, prev * nh), hd, True,
                        _off(dlnh, t * hd) if use_ln else None, _off(dlnh, (t_steps + t) * hd) if use_ln else None, n, hd, dev)
        # weight gradients: three reductions over all T*N rows (step 0 contributes zeros through h^ = r*h^ = 0)
        dxhat = torch.empty(rows, din, **f32)
        gemm(1, rows, din, 3 * hd, dg, 3 * hd, w_x, din, None, dxhat, din)
        dw_x = torch.empty(3 * hd, din, **f32)
        db = gemm_tn_colsum(3 * hd, din, rows, dg, 3 * hd, xhat, din, dw_x, din)     # dW_x and the bias gradients
        dw_h_rz = torch.empty(2 * hd, hd, **f32)
        gemm(2, 2 * hd, hd, rows, dg, 3 * hd, hhat, hd, None, dw_h_rz, hd)
        dw_h_c = torch.empty(hd, hd, **f32)
        gemm(2, hd, hd, rows, _off(dg, 2 * hd), 3 * hd, rs, hd, None, dw_h_c, hd)
        dlnx_w = dlnx_b = None
        if use_ln:
            dx = torch.empty(rows, din, **f32)
            dlnx_w = torch.empty(din, **f32)
            dlnx_b = torch.empty(din, **f32)
            _ln_bwd(lib, _ptr(dxhat), din, _ptr(x3), din, lnx_w, None, _ptr(mean_x), _ptr(rstd_x), _ptr(dx), din, False,
                    _ptr(dlnx_w), _ptr(dlnx_b), rows, din, dev)
        else:
            dx = dxhat
        dw_r = torch.cat([dw_x[:hd], dw_h_rz[:hd]], 1)
        dw_z = torch.cat([dw_x[hd:2 * hd], dw_h_rz[hd:]], 1)
        dw_c = torch.cat([dw_x[2 * hd:], dw_h_c], 1)
        dlnh_w = dlnh[0].sum(0) if use_ln else None
        dlnh_b = dlnh[1].sum(0) if use_ln else None
        dlno_w = dlno[0].sum(0) if use_ln else None
        dlno_b = dlno[1].sum(0) if use_ln else None
        return (dx.view(t_steps, n, din), None, dlnx_w, dlnx_b, dlnh_w, dlnh_b, dlno_w, dlno_b,
                dw_r, db[:hd], dw_z, db[hd:2 * hd], dw_c, db[2 * hd:], None)


def gru_scan(x3, ts, cell, reverse: bool = False):
    """x3 ``[T,N,in]`` -> states ``[T,N,hidden]`` of ``cell`` (a ``TemporalGRUCell``); ts ``[N,T]`` or None."""
    ln = cell.use_layer_norm
    return _GRUScanFn.apply(
        x3, ts,
        cell.layer_norm_x.weight if ln else None, cell.layer_norm_x.bias if ln else None,
        cell.layer_norm_h.weight if ln else None, cell.layer_norm_h.bias if ln else None,
        cell.layer_norm_out.weight if ln else None, cell.layer_norm_out.bias if ln else None,
        cell.reset_gate.weight, cell.reset_gate.bias, cell.update_gate.weight, cell.update_gate.bias,
        cell.candidate.weight, cell.candidate.bias, reverse)
