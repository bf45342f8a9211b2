"""TEST INFRASTRUCTURE ONLY -- round-2 additions to tests/golden/, again produced by RUNNING THE UNMODIFIED
REFERENCE (imported from /root/reference in the build container):

    python -m oracle.make_golden_r02

* ``propagation_h32.pt``  TemporalEvolutionLayer / TemporalSkipConnection / propagation core at hidden 32 and 64 --
  sizes at which the fused two-source GEMMs of tagan_b200.fused apply (the round-1 vectors use hidden 16).
* ``geo_c2_powerlaw.pt``  one config-2-shaped snapshot (N = 10 000, ~200 k power-law edges, H = 128, 4 heads) through
  the reference's DENSE ``TAGANGraphAttention`` (scaled_dot_product; 5.5 GB, a few seconds), forward and backward,
  plus a 1 200-node power-law snapshot for the default ``euclidean`` metric (whose dense backward is a Python loop).
  Outputs and the edge list are stored; x is re-created by an integer hash (bit-identical on any machine).
* ``tattn_per_node.pt``   AsymmetricTemporalAttention with PER-NODE timestamps (the in-kernel RBF time bias).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, _grads, _randomize, _sd  # noqa: E402


def hashed_uniform(shape, salt: int) -> torch.Tensor:
    """Deterministic values in [-1, 1) from INTEGER arithmetic only (bit-identical on every CPU; torch.randn is not: its
    vectorised and scalar paths differ between machines), so fixtures can re-create large inputs instead of storing them."""
    n = 1
    for s_ in shape:
        n *= s_
    i = torch.arange(n, dtype=torch.int64) + salt * 1_000_003
    i = (i * 2654435761) % 4294967296
    i = ((i ^ (i >> 15)) * 2246822519) % 4294967296
    i = ((i ^ (i >> 13)) * 3266489917) % 4294967296
    i = i ^ (i >> 16)
    return ((i % 65536).float() / 32768.0 - 1.0).reshape(shape)


def powerlaw_inputs(n, hidden, seed):
    """x and the output weighting of the power-law cases (shared by the generator and the tests); the edge list is stored."""
    return 1.5 * hashed_uniform((n, hidden), seed), hashed_uniform((n, hidden), seed + 1)


def powerlaw_edges(n, e, seed):
    from tagan_b200.synth import random_edges
    return random_edges(n, e, torch.Generator().manual_seed(seed), "powerlaw")


def prop_cases(ref):
    out = {}
    for hd, n, t in ((32, 9, 5), (64, 6, 7)):
        g = torch.Generator().manual_seed(500 + hd)
        xs = [torch.randn(n, hd, generator=g) for _ in range(t)]
        ts = torch.cumsum(torch.rand(n, t, generator=g) * 3.0, dim=1)
        ts[:, 2] = ts[:, 1] + 15.0                              # clamp(.,0,10)
        ts[:, 3] = ts[:, 2] - 0.5                               # negative diff clamps to 0

        def seq_case(module, call, key, **extra):
            module.eval()
            module.zero_grad(set_to_none=True)
            xin = [x_.clone().requires_grad_(True) for x_ in xs]
            wo = [torch.randn(n, hd, generator=g) for _ in range(t)]
            with ref_loader.quiet():
                ys = call(module, xin)
                sum((y * w).sum() for y, w in zip(ys, wo)).backward()
            out[key] = dict(hidden=hd, xs=[x_.detach().clone() for x_ in xin], ts=ts.clone(), wout=wo, sd=_sd(module),
                            outs=[y.detach().clone() for y in ys], dxs=[x_.grad.clone() for x_ in xin],
                            grads=_grads(module), **extra)

        torch.manual_seed(40 + hd)
        ev = ref.TemporalEvolutionLayer(hd, hd, dropout=0.0)
        _randomize(ev, 16)
        seq_case(ev, lambda m, xi: m(xi, ts), f"evolution_h{hd}")
        seq_case(ev, lambda m, xi: m(xi, None), f"evolution_no_ts_h{hd}")
        for agg, w in (("mean", 3), ("sum", 2)):
            torch.manual_seed(41 + hd)
            sk = ref.TemporalSkipConnection(hd, window_size=w, aggregation=agg, dropout=0.0)
            _randomize(sk, 18)
            seq_case(sk, lambda m, xi: m(xi), f"skip_{agg}_h{hd}", window=w, aggregation=agg)
        torch.manual_seed(42 + hd)
        tp = ref.TemporalPropagation(hd, hd, dropout=0.0)
        _randomize(tp, 20)
        tp.eval()

        def core(m, xi):
            e = m.evolution_layer(xi, ts)
            e = m.skip_connection(e)
            return [m.layer_norm(m.dropout_layer(m.output_proj(f))) for f in e]
        seq_case(tp, core, f"propagation_core_h{hd}")
    return out


def geo_powerlaw_cases(ref):
    cases = []
    for (n, e, hidden, heads, metric, seed) in ((10_000, 200_000, 128, 4, "scaled_dot_product", 7001),
                                                (1_200, 24_000, 128, 4, "euclidean", 7002)):
        x, wout = powerlaw_inputs(n, hidden, seed)
        ei = powerlaw_edges(n, e, seed)
        torch.manual_seed(seed)
        layer = ref.TAGANGraphAttention(hidden, num_heads=heads, dropout=0.0, distance_metric=metric)
        _randomize(layer, seed % 97)
        layer.eval()
        xr = x.clone().requires_grad_(True)
        with ref_loader.quiet():
            out = layer(xr, ei)
            (out * wout).sum().backward()
        deg = torch.bincount(ei[0], minlength=n)
        cases.append(dict(n=n, e=e, hidden=hidden, heads=heads, metric=metric, seed=seed, sd=_sd(layer),
                          edge_index=ei.to(torch.int32), out=out.detach().clone(), dx=xr.grad.clone(), grads=_grads(layer),
                          max_degree=int(deg.max()), x_checksum=float(x.double().sum())))
        print("geo powerlaw", n, metric, "max raw degree", int(deg.max()))
    return cases


def tattn_per_node_cases(ref):
    cases = []
    g = torch.Generator().manual_seed(31337)
    for name, b, t, hidden, heads, causal in (("pn_small", 5, 6, 32, 4, False), ("pn_t16", 7, 16, 64, 8, False),
                                              ("pn_causal", 4, 9, 32, 2, True), ("pn_t40", 3, 40, 32, 4, False)):
        torch.manual_seed(900 + b)
        m = ref.AsymmetricTemporalAttention(hidden, num_heads=heads, dropout=0.0, causal=causal)
        _randomize(m, 77 + t)
        m.eval()
        x = torch.randn(b, t, hidden, generator=g, requires_grad=True)
        ts = torch.cumsum(torch.rand(b, t, generator=g) * 2.5, dim=1)          # different per node
        ts[0, t // 2:] += 12.0                                                 # a gap beyond the +-10 band
        wout = torch.randn(b, t, hidden, generator=g)
        with ref_loader.quiet():
            out, attn = m(x, time_stamps=ts, return_attention_weights=True)
            (out * wout).sum().backward()
        cases.append(dict(name=name, b=b, t=t, hidden=hidden, heads=heads, causal=causal, x=x.detach().clone(), ts=ts,
                          wout=wout, sd=_sd(m), out=out.detach().clone(), attn=attn.detach().clone(), dx=x.grad.clone(),
                          grads=_grads(m)))
    return cases


def model_cases(ref):
    """Whole ``TAGAN.forward`` with T == num_heads == 4 (the all-ones temporal mask then broadcasts and becomes CAUSAL,
    model.py:336-361 + temporal_attention.py:1142-1170) and with equal-sized snapshots; complements tagan_model.pt, whose
    T == heads case uses 5 heads (a head count the geometric kernel does not support: power-of-two head dims only)."""
    import numpy as np
    out = []
    for name, t_steps, sizes, heads, hidden, out_dim in (("t4_h4_ragged", 4, None, 4, 64, 1), ("t6_h8_equal", 6, 12, 8, 64, 1),
                                                         ("t4_h4_multiclass", 4, None, 4, 32, 3)):
        torch.manual_seed(3)
        np.random.seed(3)
        cfg = dict(node_feature_dim=16, edge_feature_dim=8, hidden_dim=hidden, num_heads=heads, num_layers=2, output_dim=out_dim,
                   dropout=0.0, loss_type="bce", use_edge_features=True, learnable_distance=False, temporal_window_size=3)
        with ref_loader.quiet():
            model = ref.TAGAN(ref.TAGANConfig(**cfg))
        _randomize(model, 31)
        model.eval()
        seq = []
        for _ in range(t_steps):
            nt = sizes if sizes is not None else int(np.random.randint(5, 11))
            x = torch.randn(nt, 16)
            ei = torch.randint(0, nt, (2, 3 * nt))
            seq.append((x, ei, torch.randn(3 * nt, 8), list(range(nt))))
        labels = torch.tensor([1]) if out_dim > 1 else torch.tensor([[1.0]])
        with ref_loader.quiet():
            res = model(seq, labels)
            res["loss"].backward()
        out.append(dict(name=name, cfg=cfg, seq=seq, labels=labels, sd=_sd(model), logits=res["logits"].detach().clone(),
                        loss=res["loss"].detach().clone(), predictions=res["predictions"].detach().clone(), grads=_grads(model)))
        print("model case", name, "loss", float(res["loss"]))
    return out


def main():
    ref = ref_loader.load()
    torch.set_num_threads(8)
    which = set(sys.argv[1:]) or {"prop", "geo", "tattn", "model"}
    if "model" in which:
        torch.save(model_cases(ref), os.path.join(OUT, "tagan_model_r02.pt"))
    if "prop" in which:
        torch.save(prop_cases(ref), os.path.join(OUT, "propagation_h32.pt"))
    if "tattn" in which:
        torch.save(tattn_per_node_cases(ref), os.path.join(OUT, "tattn_per_node.pt"))
    if "geo" in which:
        torch.save(geo_powerlaw_cases(ref), os.path.join(OUT, "geo_c2_powerlaw.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
