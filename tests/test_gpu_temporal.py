"""GPU parity tests for hot-path rows b1-b2 (AsymmetricTemporalAttention) through the C ABI,
against the reference golden vectors and the CPU oracle.  fp32, rtol 1e-4 / atol 1e-5
(parameter gradients: atol scaled by the gradient's max magnitude, stated below)."""
import pytest
import torch

from parity_util import close as _pclose

from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-5)


def _close(got, ref, msg=None, rtol=1e-4, atol=1e-5):
    """Outputs / attention weights: the literal north-star tolerance rtol 1e-4 / atol 1e-5 (recorded, see parity_util)."""
    _pclose(got, ref, rtol=rtol, atol=atol, scaled=False, msg=msg)


def _gclose(got, ref, msg=None, rtol=1e-4, atol=1e-5):
    """Gradients (sums over nodes / entries): atol relative to the gradient's magnitude when that exceeds 1."""
    _pclose(got, ref, rtol=rtol, atol=atol, scaled=True, msg=msg)



@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# Analytically ZERO gradients (softmax shift-invariance: these biases move every score of a row
# equally); both sides hold rounding noise only, compared with atol 1e-4.
ZERO_GRADS = {"k_linear.bias", "time_q_proj.bias", "time_encoding.basis_proj.bias"}


def _gtol(gref, name=""):
    if name in ZERO_GRADS:
        return dict(rtol=1e-4, atol=1e-4)
    return dict(rtol=1e-4, atol=1e-5 * max(1.0, float(gref.abs().max())))


def _to_dev(v, dev):
    if v is None:
        return None
    if isinstance(v, list):
        return [t.to(dev) for t in v]
    return v.to(dev)


def test_temporal_attention_vs_reference_golden(dev, golden):
    import tagan_b200
    for c in golden("temporal_attention.pt"):
        layer = tagan_b200.AsymmetricTemporalAttention(
            c["hidden"], num_heads=c["heads"], dropout=0.0, causal=c["causal"],
            asymmetric_window_size=c["window"], relative_position_bias=c["rel_bias"]).to(dev)
        layer.load_state_dict(c["sd"], strict=True)
        if "x_list" in c:
            xin = [t.to(dev).requires_grad_(True) for t in c["x_list"]]
        else:
            xin = c["x"].to(dev).requires_grad_(True)
        out, attn = layer(xin, time_stamps=_to_dev(c["ts"], dev), attention_mask=_to_dev(c["mask"], dev),
                          return_attention_weights=True)
        name = c["name"]
        _close(out.detach().cpu(), c["out"], msg=lambda m: f"{name} out: {m}")
        torch.testing.assert_close(attn.cpu(), c["attn"], **TOL, msg=lambda m: f"{name} attn: {m}")
        (out * c["wout"].to(dev)).sum().backward()
        if "x_list" in c:
            for a, b in zip(xin, c["dx_list"]):
                _gclose(a.grad.cpu(), b, msg=lambda m: f"{name} dx: {m}")
        else:
            _gclose(xin.grad.cpu(), c["dx"], msg=lambda m: f"{name} dx: {m}")
        params = dict(layer.named_parameters())
        for k, gref in c["grads"].items():
            g = params[k].grad
            if gref is None:
                assert g is None or float(g.abs().max()) == 0.0, (name, k)
            else:
                assert g is not None, (name, k)
                torch.testing.assert_close(g.cpu(), gref, **_gtol(gref, k), msg=lambda m, k=k: f"{name} d{k}: {m}")


@pytest.mark.parametrize("b,t,hidden,heads", [(33, 16, 128, 8), (7, 32, 128, 4), (5, 48, 64, 4), (3, 128, 128, 8),
                                              (9, 5, 64, 4), (4, 16, 256, 8), (6, 11, 40, 5), (6, 11, 64, 4), (5, 9, 256, 8)])
@pytest.mark.parametrize("mode", ["shared_ts", "no_ts_causal", "per_node_ts", "mask3d", "unsorted_ts", "boundary_ts"])
def test_temporal_attention_vs_oracle_shapes(dev, b, t, hidden, heads, mode):
    import tagan_b200
    torch.manual_seed(b * 1000 + t)
    layer = tagan_b200.AsymmetricTemporalAttention(hidden, num_heads=heads, dropout=0.0,
                                                   causal=(mode == "no_ts_causal")).to(dev)
    with torch.no_grad():
        for n_, p in layer.named_parameters():
            if p.dim() == 1 and "basis" not in n_:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(b, t, hidden) * 0.7
    wout = torch.randn(b, t, hidden)
    ts = mask = None
    if mode == "shared_ts":
        ts = torch.arange(t).float().repeat(b, 1)
    elif mode == "per_node_ts":
        ts = torch.cumsum(torch.rand(b, t) * 2.5, dim=1)
    elif mode == "unsorted_ts":                      # band mask on non-monotone times: no key window can be derived
        ts = (torch.rand(1, t) * 40.0).repeat(b, 1)
    elif mode == "boundary_ts":                      # sorted, duplicates, gaps of exactly the band width (|dt| == 10 is valid)
        ts = (torch.div(torch.arange(t), 3, rounding_mode="floor").float() * 10.0).repeat(b, 1)
    elif mode == "mask3d":
        mask = torch.maximum((torch.rand(b, t, t) > 0.5).float(), torch.eye(t).unsqueeze(0))
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in layer.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    ref, aref = R.asym_temporal_attention(xr, sd, heads, time_stamps=ts, attention_mask=mask,
                                          causal=(mode == "no_ts_causal"), return_attention_weights=True)
    (ref * wout).sum().backward()
    xd = x.to(dev).requires_grad_(True)
    out, attn = layer(xd, time_stamps=_to_dev(ts, dev), attention_mask=_to_dev(mask, dev), return_attention_weights=True)
    (out * wout.to(dev)).sum().backward()
    _close(out.detach().cpu(), ref.detach())
    torch.testing.assert_close(attn.cpu(), aref.detach(), **TOL)
    _gclose(xd.grad.cpu(), xr.grad)
    for k, p in layer.named_parameters():
        gref = sd[k].grad
        if gref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            torch.testing.assert_close(p.grad.cpu(), gref, **_gtol(gref, k), msg=lambda m, k=k: f"d{k}: {m}")


def test_temporal_attention_time_major_matches_batch_major(dev):
    """List-of-snapshots input ([T,N,H] in memory) and the stacked [N,T,H] tensor give identical results."""
    import tagan_b200
    torch.manual_seed(0)
    n, t, hidden, heads = 50, 16, 128, 8
    layer = tagan_b200.AsymmetricTemporalAttention(hidden, num_heads=heads, dropout=0.0).to(dev)
    xs = [torch.randn(n, hidden, device=dev) for _ in range(t)]
    a = layer(xs, attention_mask=torch.ones(t, t, device=dev))
    b = layer(torch.stack(xs, 1), attention_mask=torch.ones(t, t, device=dev))
    assert a.shape == b.shape == (n, t, hidden)
    torch.testing.assert_close(a, b, rtol=0, atol=0)


def test_temporal_attention_full_size_properties(dev):
    """Config-3 size (100k nodes x 16 snapshots, H=128, h=8): rows of the attention matrix sum to 1,
    band/causal structure holds, run-to-run bit-identical (deterministic backward)."""
    from tagan_b200 import ops
    g = torch.Generator().manual_seed(1)
    b, t, hidden, heads = 100_000, 16, 128, 8
    qkv = (torch.randn(b * t, 3 * hidden, generator=g) * 0.5).to(dev)
    bias = torch.randn(heads, t, t, generator=g).to(dev)
    tm = ops.TemporalMask(flags=1)
    ctx1, _ = ops.temporal_attention_core(qkv, bias, tm, b, t, heads)
    ctx2, _ = ops.temporal_attention_core(qkv, bias, tm, b, t, heads)
    assert torch.equal(ctx1, ctx2)
    sub = 2000
    _, attn = ops.temporal_attention_core(qkv[:sub * t], bias, tm, sub, t, heads, want_attn=True)
    torch.testing.assert_close(attn.sum(-1), torch.ones(sub, heads, t, device=dev), rtol=1e-4, atol=1e-5)
    assert float(attn.triu(1).abs().max()) == 0.0                      # causal
    q = qkv.clone().requires_grad_(True)
    bb = bias.clone().requires_grad_(True)
    d = torch.randn(b * t, hidden, generator=g).to(dev)
    c, _ = ops.temporal_attention_core(q, bb, tm, b, t, heads)
    c.backward(d)
    q2 = qkv.clone().requires_grad_(True)
    bb2 = bias.clone().requires_grad_(True)
    c, _ = ops.temporal_attention_core(q2, bb2, tm, b, t, heads)
    c.backward(d)
    assert torch.equal(q.grad, q2.grad) and torch.equal(bb.grad, bb2.grad)
