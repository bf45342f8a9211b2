"""GPU parity tests of the round-2 fusions (tagan_b200/fused.py, csrc/fused_rows.cu, the fused epilogues of
csrc/gemm_tma.cu): each fused kernel against an fp64 torch restatement of the reference arithmetic it replaces,
each stage Function against the reference golden vectors (hidden 32 / 64, where the fused paths apply) and against
the unfused op-by-op composition on the same inputs.  fp32: rtol 1e-4 / atol 1e-5 for outputs; gradients that are
sums over rows take the atol relative to their magnitude."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-5)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _gclose(got, ref, msg=None):
    scale = max(1.0, float(ref.abs().max())) if ref.numel() else 1.0
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5 * scale, msg=msg)


def _d(t):
    return t.detach().double().cpu()


# ------------------------------------------------------------------------------------------
# fused GEMM epilogues
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,res", [(1, 128, 128, True), (37, 64, 32, True), (300, 128, 256, False),
                                       (1000, 128, 128, True), (129, 32, 96, True), (4096, 128, 128, True)])
def test_gemm_res_ln_epilogue(dev, m, n, k, res):
    from tagan_b200 import fused
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev)
    r = (3.0 + 2.0 * torch.randn(m, n, device=dev)) if res else None           # non-zero row mean: exercises the shifted moments
    g = 1.0 + 0.3 * torch.randn(n, device=dev)
    be = 0.2 * torch.randn(n, device=dev)
    fused.FUSED_RES_LN = True
    try:
        y, xsum, mean, rstd = fused.linear_res_ln(x, w, b, r, g, be, need_sum=True)
    finally:
        fused.FUSED_RES_LN = False
    v = _d(x) @ _d(w).t() + _d(b) + (_d(r) if res else 0.0)
    ref = F.layer_norm(v, (n,), _d(g), _d(be), 1e-5)
    torch.testing.assert_close(xsum.double().cpu(), v, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(y.double().cpu(), ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(mean.double().cpu(), v.mean(1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rstd.double().cpu(), 1.0 / torch.sqrt(v.var(1, unbiased=False) + 1e-5), rtol=1e-4, atol=1e-5)
    # the unfused composition (GEMM, then the LayerNorm kernel: the default) gives the same numbers
    y2, xsum2, _, _ = fused.linear_res_ln(x, w, b, r, g, be, need_sum=True)
    torch.testing.assert_close(y, y2, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,din,hd", [(5, 32, 32), (257, 64, 64), (1000, 128, 128), (300, 256, 256), (130, 32, 64)])
def test_gemm_gru_epilogues(dev, rows, din, hd):
    """GATES / BLEND / GATES_BWD with the two-source A operand against the concatenated fp64 arithmetic
    (reference temporal_propagation.py:531-542 and its autograd)."""
    from tagan_b200 import _lib, fused
    torch.manual_seed(rows + hd)
    kk = din + hd
    xh = torch.randn(rows, din, device=dev)
    hh = torch.randn(rows, hd, device=dev)
    w_rz = torch.randn(2 * hd, kk, device=dev) / kk ** 0.5
    b_rz = torch.randn(2 * hd, device=dev)
    w_c = torch.randn(hd, kk, device=dev) / kk ** 0.5
    b_c = torch.randn(hd, device=dev)
    r, z, rs, cand, hn = (torch.empty(rows, hd, device=dev) for _ in range(5))
    epi = fused._epi(_lib.EPI_GATES, split=hd, in0=hh, ld_in0=hd, out0=r, ld_out0=hd, out1=rs, ld_out1=hd, out2=z, ld_out2=hd)
    fused.gemm_fused(0, rows, 2 * hd, kk, xh, din, hh, hd, din, w_rz, kk, b_rz, epi, dev)
    epi = fused._epi(_lib.EPI_BLEND, in0=z, ld_in0=hd, in1=hh, ld_in1=hd, out0=cand, ld_out0=hd, out1=hn, ld_out1=hd)
    fused.gemm_fused(0, rows, hd, kk, xh, din, rs, hd, din, w_c, kk, b_c, epi, dev)
    pre = torch.cat([_d(xh), _d(hh)], 1) @ _d(w_rz).t() + _d(b_rz)
    r_ref, z_ref = torch.sigmoid(pre[:, :hd]), torch.sigmoid(pre[:, hd:])
    rs_ref = r_ref * _d(hh)
    cand_ref = torch.tanh(torch.cat([_d(xh), rs_ref], 1) @ _d(w_c).t() + _d(b_c))
    hn_ref = (1 - z_ref) * _d(hh) + z_ref * cand_ref
    for got, ref, nm in ((r, r_ref, "r"), (z, z_ref, "z"), (rs, rs_ref, "rs"), (cand, cand_ref, "cand"), (hn, hn_ref, "hn")):
        torch.testing.assert_close(got.double().cpu(), ref, rtol=1e-4, atol=1e-5, msg=lambda s, nm=nm: f"{nm}: {s}")
    # backward epilogue: d(rs) = dgc . W_c[:, in:];  dg_r = d(rs) * hh * r(1-r);  dhh += d(rs) * r
    dg = torch.randn(rows, 3 * hd, device=dev)
    dhh0 = torch.randn(rows, hd, device=dev)
    dhh = dhh0.clone()
    w_c_h = C.c_void_p(w_c.data_ptr() + din * 4)
    epi = fused._epi(_lib.EPI_GATES_BWD, in0=r, ld_in0=hd, in1=hh, ld_in1=hd, out0=dg, ld_out0=3 * hd, out1=dhh, ld_out1=hd)
    dgc = dg[:, 2 * hd:].clone()
    fused.gemm_fused(1, rows, hd, hd, C.c_void_p(dg.data_ptr() + 2 * hd * 4), 3 * hd, None, 0, 0, w_c_h, kk, None, epi, dev)
    drs = _d(dgc) @ _d(w_c)[:, din:]
    torch.testing.assert_close(dg[:, :hd].double().cpu(), drs * _d(hh) * _d(r) * (1 - _d(r)), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dhh.double().cpu(), _d(dhh0) + drs * _d(r), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dg[:, 2 * hd:], dgc, rtol=0, atol=0)            # the A operand slice is untouched


# ------------------------------------------------------------------------------------------
# fused row kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(1, 16), (37, 32), (300, 64), (1000, 128), (257, 256), (50, 96), (9, 512)])
def test_ln_pair_fwd_bwd(dev, rows, cols):
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(rows * 7 + cols)
    p, st = ops._ptr, ops._stream
    hn = (2.0 + torch.randn(rows, cols)).to(dev)
    go, bo, gh, bh = ((1.0 + 0.3 * torch.randn(cols)).to(dev), (0.2 * torch.randn(cols)).to(dev),
                      (1.0 + 0.3 * torch.randn(cols)).to(dev), (0.2 * torch.randn(cols)).to(dev))
    ts = torch.cumsum(torch.rand(rows, 4) * 4.0, 1).to(dev)
    ts[0, 2] = ts[0, 1] + 20.0
    s, hh = torch.empty(rows, cols, device=dev), torch.empty(rows, cols, device=dev)
    mo, ro, mh, rh, dec = (torch.empty(rows, device=dev) for _ in range(5))
    _lib.check(lib.tagan_ln_pair_fwd(p(hn), cols, p(go), p(bo), p(gh), p(bh), p(ts), 4, 2, p(s), cols, p(hh), cols, p(mo), p(ro),
                                     p(mh), p(rh), p(dec), rows, cols, st()), "ln_pair_fwd")
    hn64 = _d(hn).requires_grad_(True)
    prm = [_d(t).requires_grad_(True) for t in (go, bo, gh, bh)]
    s_ref = F.layer_norm(hn64, (cols,), prm[0], prm[1], 1e-5)
    dec_ref = torch.exp(-torch.clamp(_d(ts)[:, 2] - _d(ts)[:, 1], 0.0, 10.0))
    hh_ref = F.layer_norm(s_ref, (cols,), prm[2], prm[3], 1e-5) * dec_ref[:, None]
    torch.testing.assert_close(s.double().cpu(), s_ref.detach(), **TOL)
    torch.testing.assert_close(hh.double().cpu(), hh_ref.detach(), **TOL)
    torch.testing.assert_close(dec.double().cpu(), dec_ref, **TOL)
    ds_ext, dhh = torch.randn(rows, cols, device=dev), torch.randn(rows, cols, device=dev)
    (s_ref * _d(ds_ext)).sum().backward(retain_graph=True)
    (hh_ref * _d(dhh)).sum().backward()
    dhn = torch.empty(rows, cols, device=dev)
    daff = torch.zeros(4, cols, device=dev)
    ws = ops.workspace(lib.tagan_ln_pair_bwd_workspace_bytes(rows, cols), dev)
    _lib.check(lib.tagan_ln_pair_bwd(p(ds_ext), cols, p(dhh), cols, p(hn), cols, p(go), p(bo), p(gh), p(mo), p(ro), p(mh), p(rh),
                                     p(dec), p(dhn), cols, p(daff), 1, p(ws), ws.numel(), rows, cols, st()), "ln_pair_bwd")
    _gclose(dhn.double().cpu(), hn64.grad)
    for i, nm in enumerate(("gamma_o", "beta_o", "gamma_h", "beta_h")):
        _gclose(daff[i].double().cpu(), prm[i].grad, msg=lambda m, nm=nm: f"d{nm}: {m}")
    # no next step: only LN_out
    s2 = torch.empty_like(s)
    _lib.check(lib.tagan_ln_pair_fwd(p(hn), cols, p(go), p(bo), None, None, None, 0, 0, p(s2), cols, None, 0, p(mo), p(ro),
                                     None, None, None, rows, cols, st()), "ln_pair_fwd")
    torch.testing.assert_close(s2, s, rtol=0, atol=0)


@pytest.mark.parametrize("rows,cols", [(2, 16), (300, 64), (1000, 128), (257, 256), (33, 72)])
def test_gru_blend_ln_fwd_bwd(dev, rows, cols):
    """tanh / blend / LN_out / LN_h(+decay) of one GRU step in one pass, and its backward, against fp64 autograd of the
    reference arithmetic (temporal_propagation.py:538-546 and :505-514)."""
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(rows * 3 + cols)
    p, st = ops._ptr, ops._stream
    gc = torch.randn(rows, cols).to(dev)
    z = torch.rand(rows, cols).to(dev)
    hh = torch.randn(rows, cols).to(dev)
    go, bo, gh, bh = ((1.0 + 0.3 * torch.randn(cols)).to(dev), (0.2 * torch.randn(cols)).to(dev),
                      (1.0 + 0.3 * torch.randn(cols)).to(dev), (0.2 * torch.randn(cols)).to(dev))
    ts = torch.cumsum(torch.rand(rows, 3) * 2.0, 1).to(dev)
    cand, s, hn_next = (torch.empty(rows, cols, device=dev) for _ in range(3))
    mo, ro, mh, rh, dec = (torch.empty(rows, device=dev) for _ in range(5))
    _lib.check(lib.tagan_gru_blend_ln_fwd(p(gc), cols, p(z), p(hh), cols, p(cand), p(go), p(bo), p(gh), p(bh), p(ts), 3, 1, p(s), cols,
                                          p(hn_next), cols, p(mo), p(ro), p(mh), p(rh), p(dec), rows, cols, st()), "blend_ln_fwd")
    gc64, z64, hh64 = (_d(t).requires_grad_(True) for t in (gc, z, hh))
    prm = [_d(t).requires_grad_(True) for t in (go, bo, gh, bh)]
    cand_ref = torch.tanh(gc64)
    hn_ref = (1 - z64) * hh64 + z64 * cand_ref
    s_ref = F.layer_norm(hn_ref, (cols,), prm[0], prm[1], 1e-5)
    dec_ref = torch.exp(-torch.clamp(_d(ts)[:, 1] - _d(ts)[:, 0], 0.0, 10.0))
    nx_ref = F.layer_norm(s_ref, (cols,), prm[2], prm[3], 1e-5) * dec_ref[:, None]
    torch.testing.assert_close(cand.double().cpu(), cand_ref.detach(), **TOL)
    torch.testing.assert_close(s.double().cpu(), s_ref.detach(), **TOL)
    torch.testing.assert_close(hn_next.double().cpu(), nx_ref.detach(), **TOL)
    ds_ext, dnx = torch.randn(rows, cols, device=dev), torch.randn(rows, cols, device=dev)
    ((s_ref * _d(ds_ext)).sum() + (nx_ref * _d(dnx)).sum()).backward()
    dg = torch.zeros(rows, 3 * cols, device=dev)
    dhh = torch.empty(rows, cols, device=dev)
    daff = torch.zeros(4, cols, device=dev)
    ws = ops.workspace(lib.tagan_ln_pair_bwd_workspace_bytes(rows, cols), dev)
    off = lambda t, e: C.c_void_p(t.data_ptr() + 4 * e)
    _lib.check(lib.tagan_gru_blend_ln_bwd(p(ds_ext), cols, p(dnx), cols, p(cand), p(z), p(hh), cols, off(dg, cols), off(dg, 2 * cols),
                                          3 * cols, p(dhh), cols, p(go), p(bo), p(gh), p(mo), p(ro), p(mh), p(rh), p(dec), p(daff), 1,
                                          p(ws), ws.numel(), rows, cols, st()), "blend_ln_bwd")
    # d z_pre = dz * z(1-z) with dz the gradient of z itself; d cand_pre = gradient of gc
    _gclose(dg[:, cols:2 * cols].double().cpu(), z64.grad * _d(z) * (1 - _d(z)))
    _gclose(dg[:, 2 * cols:].double().cpu(), gc64.grad)
    _gclose(dhh.double().cpu(), hh64.grad)
    assert float(dg[:, :cols].abs().max()) == 0.0
    for i, nm in enumerate(("gamma_o", "beta_o", "gamma_h", "beta_h")):
        _gclose(daff[i].double().cpu(), prm[i].grad, msg=lambda m, nm=nm: f"d{nm}: {m}")
    # reset-gate backward
    drs, r = torch.randn(rows, cols, device=dev), torch.rand(rows, cols, device=dev)
    dhh0 = dhh.clone()
    _lib.check(lib.tagan_gru_reset_bwd(p(drs), p(r), p(hh), cols, p(dg), 3 * cols, p(dhh), cols, rows, cols, st()), "reset_bwd")
    torch.testing.assert_close(dg[:, :cols].double().cpu(), _d(drs) * _d(hh) * _d(r) * (1 - _d(r)), **TOL)
    torch.testing.assert_close(dhh.double().cpu(), _d(dhh0) + _d(drs) * _d(r), **TOL)


@pytest.mark.parametrize("rows,cols", [(3, 16), (300, 64), (1000, 128), (130, 256), (40, 200)])
def test_gelu_ln_fwd_bwd(dev, rows, cols):
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(rows + cols)
    p, st = ops._ptr, ops._stream
    a = (1.5 * torch.randn(rows, cols)).to(dev)
    g, b = (1.0 + 0.3 * torch.randn(cols)).to(dev), (0.2 * torch.randn(cols)).to(dev)
    y = torch.empty_like(a)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    _lib.check(lib.tagan_gelu_ln_fwd(p(a), cols, p(g), p(b), p(y), cols, p(mean), p(rstd), rows, cols, st()), "gelu_ln_fwd")
    a64, g64, b64 = (_d(t).requires_grad_(True) for t in (a, g, b))
    ref = F.layer_norm(F.gelu(a64), (cols,), g64, b64, 1e-5)
    torch.testing.assert_close(y.double().cpu(), ref.detach(), **TOL)
    dy = torch.randn(rows, cols, device=dev)
    (ref * _d(dy)).sum().backward()
    da = torch.empty_like(a)
    daff = torch.empty(2, cols, device=dev)
    ws = ops.workspace(lib.tagan_gelu_ln_bwd_workspace_bytes(rows, cols), dev)
    _lib.check(lib.tagan_gelu_ln_bwd(p(dy), cols, p(a), cols, p(g), p(mean), p(rstd), p(da), cols, p(daff), p(ws), ws.numel(),
                                     rows, cols, st()), "gelu_ln_bwd")
    _gclose(da.double().cpu(), a64.grad)
    _gclose(daff[0].double().cpu(), g64.grad)
    _gclose(daff[1].double().cpu(), b64.grad)


@pytest.mark.parametrize("t,inner,window,agg", [(1, 64, 3, 0), (5, 128, 1, 0), (16, 1024, 3, 0), (7, 4096, 2, 2), (32, 256, 4, 0),
                                                (3, 8, 3, 2)])
def test_window_gelu_fwd_bwd(dev, t, inner, window, agg):
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(t * 31 + inner)
    p_, st = ops._ptr, ops._stream
    x = torch.randn(t, inner, device=dev)
    out = torch.empty_like(x)
    _lib.check(lib.tagan_window_gelu_fwd(p_(x), p_(out), t, inner, window, agg, st()), "window_gelu_fwd")
    x64 = _d(x).requires_grad_(True)
    rows = []
    for i in range(t):
        seg = x64[max(0, i - window):min(t, i + window + 1)]
        rows.append(seg.mean(0) if agg == 0 else seg.sum(0))
    ref = F.gelu(torch.stack(rows))
    torch.testing.assert_close(out.double().cpu(), ref.detach(), **TOL)
    dg = torch.randn(t, inner, device=dev)
    (ref * _d(dg)).sum().backward()
    dp = torch.empty_like(x)
    _lib.check(lib.tagan_window_gelu_bwd(p_(dg), p_(x), p_(dp), t, inner, window, agg, st()), "window_gelu_bwd")
    torch.testing.assert_close(dp.double().cpu(), x64.grad, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("shape", [(7,), (1000, 3), (100000, 16, 8)])
def test_mean_square(dev, shape):
    from tagan_b200 import fused
    torch.manual_seed(len(shape))
    x = torch.randn(*shape, device=dev, requires_grad=True)
    loss = fused.mean_square(x)
    (loss * 3.0).backward()
    x64 = _d(x).requires_grad_(True)
    ref = x64.square().mean()
    (ref * 3.0).backward()
    torch.testing.assert_close(loss.double().cpu(), ref.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(x.grad.double().cpu(), x64.grad, rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------------------------------
# stage Functions against the reference goldens at hidden 32 / 64
# ------------------------------------------------------------------------------------------
def _check_param_grads(module, grads, tag):
    params = dict(module.named_parameters())
    for k, gref in grads.items():
        g = params[k].grad
        if gref is None:
            assert g is None or float(g.abs().max()) == 0.0, (tag, k)
        else:
            assert g is not None, (tag, k)
            _gclose(g.cpu(), gref, msg=lambda m, k=k: f"{tag} d{k}: {m}")


def _seq_case(dev, c, module, call, tag):
    module.load_state_dict(c["sd"])
    module.zero_grad(set_to_none=True)
    xs = [t.to(dev).requires_grad_(True) for t in c["xs"]]
    ys = call(module, xs)
    for y, yr in zip(ys, c["outs"]):
        torch.testing.assert_close(y.detach().cpu(), yr, **TOL, msg=lambda m: f"{tag} out: {m}")
    sum((y * w.to(dev)).sum() for y, w in zip(ys, c["wout"])).backward()
    for a, b in zip(xs, c["dxs"]):
        _gclose(a.grad.cpu(), b, msg=lambda m: f"{tag} dx: {m}")
    _check_param_grads(module, c["grads"], tag)


@pytest.mark.parametrize("hd", [32, 64])
@pytest.mark.parametrize("fusion", [True, False])
def test_fused_stages_golden(dev, golden, hd, fusion):
    """The reference's own outputs and autograd gradients (oracle/make_golden_r02.py) at sizes where the two-source
    GEMMs, the GEMM epilogues and the fused row passes are the code that runs (fusion=True) -- and the op-by-op
    composition on the same vectors (fusion=False)."""
    import tagan_b200
    from tagan_b200 import fused, ops
    g = golden("propagation_h32.pt")
    ops.FUSION = fusion
    try:
        c = g[f"evolution_h{hd}"]
        ts = c["ts"].to(dev)
        if fusion:
            assert fused.evolution_supported(hd, hd)
        _seq_case(dev, c, tagan_b200.TemporalEvolutionLayer(hd, hd, dropout=0.0).to(dev), lambda m, xs: m(xs, ts), "evolution")
        _seq_case(dev, g[f"evolution_no_ts_h{hd}"], tagan_b200.TemporalEvolutionLayer(hd, hd, dropout=0.0).to(dev),
                  lambda m, xs: m(xs, None), "evolution_no_ts")
        for agg in ("mean", "sum"):
            c = g[f"skip_{agg}_h{hd}"]
            _seq_case(dev, c, tagan_b200.TemporalSkipConnection(hd, window_size=c["window"], aggregation=agg,
                                                                dropout=0.0).to(dev), lambda m, xs: m(xs), "skip_" + agg)
        tp = tagan_b200.TemporalPropagation(hd, hd, dropout=0.0).to(dev)
        _seq_case(dev, g[f"propagation_core_h{hd}"], tp, lambda m, xs: list(m.forward_core(xs, ts).unbind(0)), "core")
    finally:
        ops.FUSION = True


@pytest.mark.parametrize("n,t,hidden,heads", [(500, 8, 128, 8), (130, 5, 64, 4), (64, 4, 256, 8)])
def test_layer_fused_equals_unfused(dev, n, t, hidden, heads):
    """Whole TAGAN layer (geometric + propagation + temporal attention): stage-fused path vs the op-by-op
    composition of round 1 (itself pinned to the goldens) on the same weights and inputs."""
    import tagan_b200
    from tagan_b200 import ops
    torch.manual_seed(n)
    layer = tagan_b200.TAGANLayer(hidden, heads, "euclidean").to(dev)
    xs = [torch.randn(n, hidden, device=dev) for _ in range(t)]
    eis = [torch.randint(0, n, (2, 6 * n), device=dev) for _ in range(t)]
    ts = torch.arange(t, dtype=torch.float32, device=dev).expand(n, t)
    wout = torch.randn(n, t, hidden, device=dev)
    res = {}
    for fusion in (True, False):
        ops.FUSION = fusion
        try:
            layer.zero_grad(set_to_none=True)
            xin = [x.clone().requires_grad_(True) for x in xs]
            out = layer(xin, eis, ts)
            (out * wout).sum().backward()
            res[fusion] = (out.detach().clone(), [x.grad.clone() for x in xin],
                           {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None})
        finally:
            ops.FUSION = True
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-4, atol=2e-5)
    for a, b in zip(res[True][1], res[False][1]):
        _gclose(a, b)
    assert res[True][2].keys() == res[False][2].keys()
    # analytically zero by softmax shift invariance: both sides hold only the rounding noise of a sum over n*t*heads terms
    # of magnitude ~1 (measured 1.2e-4 for one element once kernel (a) moved to the SFU exp / sqrt): atol 5e-4
    zero = ("temporal_attention.k_linear.bias", "temporal_attention.time_q_proj.bias",
            "temporal_attention.time_encoding.basis_proj.bias")
    for k in res[True][2]:
        if k in zero:
            torch.testing.assert_close(res[True][2][k], res[False][2][k], rtol=1e-4, atol=5e-4, msg=lambda m, k=k: f"d{k}: {m}")
        else:
            _gclose(res[True][2][k], res[False][2][k], msg=lambda m, k=k: f"d{k}: {m}")
