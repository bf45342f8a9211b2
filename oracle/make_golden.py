"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

The reference has no golden vectors of its own, so these files are the pin:
inputs, the reference module's ``state_dict``, its outputs, its attention weights and its
autograd gradients, all produced by the reference classes imported from /root/reference.
They travel with the repo; nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.restate import METRICS  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _grads(m):
    return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}


def _randomize(m, seed):
    """Move biases / LN params / tables off their zero/one init so every term is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() == 1 or "table" in name or "kernel" in name:
                if "basis_sigma" in name:
                    p.copy_(0.05 + 0.2 * torch.rand(p.shape, generator=g))
                elif "basis_mu" in name:
                    p.add_(0.05 * torch.randn(p.shape, generator=g))
                elif "distance_param" in name:
                    p.copy_(0.3 + torch.rand(p.shape, generator=g))
                else:
                    p.add_(0.2 * torch.randn(p.shape, generator=g))


def geo_cases(ref):
    cases = []
    g = torch.Generator().manual_seed(1234)
    shapes = [(37, 150, 32, 4), (23, 60, 64, 4)]
    for si, (n, e, hdim, heads) in enumerate(shapes):
        for metric in METRICS:
            for learnable in ([False, True] if metric in ("gaussian_kernel", "rbf_kernel") else [False]):
                torch.manual_seed(100 + si)
                layer = ref.TAGANGraphAttention(hdim, num_heads=heads, dropout=0.0, distance_metric=metric,
                                                use_layer_norm=True, learnable_distance=learnable)
                _randomize(layer, 7 + si)
                layer.eval()
                x = torch.randn(n, hdim, generator=g, requires_grad=True)
                ei = torch.randint(0, n, (2, e), generator=g)
                ei[:, :5] = ei[:, 5:10]            # duplicates
                ei[1, 10:13] = ei[0, 10:13]        # explicit self edges
                ei[0, 13] = -1                     # negative index wraps (torch advanced indexing)
                wout = torch.randn(n, hdim, generator=g)
                captured = {}
                ga = layer.geometric_attention
                orig = ga._get_attention_weights

                def hook(q, k, mask=None, _orig=orig, _c=captured):
                    a = _orig(q, k, mask)
                    _c["attn"] = a.detach().clone()
                    return a
                ga._get_attention_weights = hook
                with ref_loader.quiet():
                    out = layer(x, ei, None)
                    (out * wout).sum().backward()
                ga._get_attention_weights = orig
                cases.append(dict(n=n, hidden=hdim, heads=heads, metric=metric, learnable=learnable,
                                  x=x.detach().clone(), edge_index=ei, wout=wout,
                                  sd=_sd(layer.geometric_attention), out=out.detach().clone(),
                                  attn_dense=captured["attn"][0], dx=x.grad.clone(),
                                  grads=_grads(layer.geometric_attention)))
    # no-layer-norm variant
    torch.manual_seed(5)
    layer = ref.TAGANGraphAttention(32, num_heads=2, dropout=0.0, distance_metric="euclidean", use_layer_norm=False)
    layer.eval()
    x = torch.randn(11, 32, generator=g, requires_grad=True)
    ei = torch.randint(0, 11, (2, 30), generator=g)
    wout = torch.randn(11, 32, generator=g)
    with ref_loader.quiet():
        out = layer(x, ei, None)
        (out * wout).sum().backward()
    cases.append(dict(n=11, hidden=32, heads=2, metric="euclidean", learnable=False, no_ln=True,
                      x=x.detach().clone(), edge_index=ei, wout=wout, sd=_sd(layer.geometric_attention),
                      out=out.detach().clone(), attn_dense=None, dx=x.grad.clone(),
                      grads=_grads(layer.geometric_attention)))
    return cases


def tattn_cases(ref):
    cases = []
    g = torch.Generator().manual_seed(4321)

    def run(name, b, t, hdim, heads, ts=None, mask=None, causal=False, as_list=False, rel_bias=True,
            window=5, ragged=None):
        torch.manual_seed(11)
        layer = ref.AsymmetricTemporalAttention(hdim, num_heads=heads, dropout=0.0, causal=causal,
                                                asymmetric_window_size=window, relative_position_bias=rel_bias)
        _randomize(layer, 3)
        layer.eval()
        if as_list:
            sizes = ragged or [b] * t
            xs = [torch.randn(sz, hdim, generator=g, requires_grad=True) for sz in sizes]
            xin = xs
        else:
            x = torch.randn(b, t, hdim, generator=g, requires_grad=True)
            xin = x
        wout = torch.randn(b, t, hdim, generator=g)
        with ref_loader.quiet():
            out, attn = layer(xin, time_stamps=ts, attention_mask=mask, return_attention_weights=True)
            (out * wout).sum().backward()
        c = dict(name=name, b=b, t=t, hidden=hdim, heads=heads, causal=causal, rel_bias=rel_bias, window=window,
                 ts=ts, mask=mask, wout=wout, sd=_sd(layer), out=out.detach().clone(), attn=attn.detach().clone(),
                 grads=_grads(layer))
        if as_list:
            c["x_list"] = [t_.detach().clone() for t_ in xs]
            c["dx_list"] = [t_.grad.clone() for t_ in xs]
        else:
            c["x"] = x.detach().clone()
            c["dx"] = x.grad.clone()
        cases.append(c)

    b, t, hd, hh = 3, 5, 16, 2
    run("plain_nomask", b, t, hd, hh)
    run("ts_uniform_allones_causal", b, t, hd, hh, ts=torch.arange(t).float().repeat(b, 1))
    ts_gap = torch.tensor([[0., 1., 2., 30., 31.], [0., 5., 11., 12., 40.], [3., 4., 5., 6., 7.]])
    run("ts_gaps_band", b, t, hd, hh, ts=ts_gap)
    m3 = (torch.rand(b, t, t, generator=g) > 0.4).float()
    m3 = torch.maximum(m3, torch.eye(t).unsqueeze(0))
    run("mask3d", b, t, hd, hh, mask=m3)
    run("mask3d_ts", b, t, hd, hh, ts=ts_gap, mask=m3)
    run("mask_wrong_shape", b, t, hd, hh, mask=torch.ones(b, t + 1, t + 1))
    run("mask_list", b, t, hd, hh, mask=[torch.ones(4) for _ in range(t)])
    run("mask2d_ones_T_ne_h", b, t, hd, hh, mask=torch.ones(t, t))
    run("mask2d_ones_T_eq_h", 4, 4, 16, 4, mask=torch.ones(4, 4))
    run("mask3d_ones", b, t, hd, hh, mask=torch.ones(b, t, t))
    run("causal_flag", b, t, hd, hh, causal=True)
    run("causal_flag_ts", b, t, hd, hh, causal=True, ts=ts_gap)
    run("T12_ts_band_noncausal", 2, 12, 16, 2, ts=torch.arange(12).float().repeat(2, 1))
    run("T40_relclamp", 2, 40, 32, 4, ts=torch.arange(40).float().repeat(2, 1) * 0.7)
    run("list_ragged", 6, 5, hd, hh, as_list=True, ragged=[6, 4, 5, 6, 3], mask=torch.ones(5, 5))
    run("no_relbias_w2", b, t, hd, hh, rel_bias=False, window=2, ts=ts_gap)
    run("ts_constant", b, t, hd, hh, ts=torch.full((b, t), 2.0))
    return cases


def prop_cases(ref):
    g = torch.Generator().manual_seed(99)
    out = {}
    n, t, hd = 7, 6, 16
    xs = [torch.randn(n, hd, generator=g) for _ in range(t)]
    ts = torch.cumsum(torch.rand(n, t, generator=g) * 3.0, dim=1)
    ts[:, 3] = ts[:, 2] + 20.0          # exercises clamp(.,0,10)
    ts[:, 4] = ts[:, 3] - 1.0           # negative diff clamps to 0
    ts[:, 5] = ts[:, 4] + 0.5

    # GRU cell
    torch.manual_seed(21)
    cell = ref.TemporalGRUCell(hd, hd, dropout=0.0)
    _randomize(cell, 5)
    cell.eval()
    x = xs[0].clone().requires_grad_(True)
    h = xs[1].clone().requires_grad_(True)
    td = ts[:, 1] - ts[:, 0]
    wout = torch.randn(n, hd, generator=g)
    with ref_loader.quiet():
        o0 = cell(x, None, None)
        o1 = cell(x, h, td)
        (o1 * wout).sum().backward()
    out["gru_cell"] = dict(x=x.detach().clone(), h=h.detach().clone(), td=td, wout=wout, sd=_sd(cell),
                           out_h_none=o0.detach().clone(), out=o1.detach().clone(), dx=x.grad.clone(),
                           dh=h.grad.clone(), grads=_grads(cell))

    def seq_case(module, call, key, **extra):
        module.eval()
        module.zero_grad(set_to_none=True)
        xin = [x_.clone().requires_grad_(True) for x_ in xs]
        wo = [torch.randn(n, hd, generator=g) for _ in range(t)]
        with ref_loader.quiet():
            ys = call(module, xin)
            sum((y * w).sum() for y, w in zip(ys, wo)).backward()
        out[key] = dict(xs=[x_.detach().clone() for x_ in xin], ts=ts, wout=wo, sd=_sd(module),
                        outs=[y.detach().clone() for y in ys], dxs=[x_.grad.clone() for x_ in xin],
                        grads=_grads(module), **extra)

    torch.manual_seed(22)
    ev = ref.TemporalEvolutionLayer(hd, hd, dropout=0.0)
    _randomize(ev, 6)
    seq_case(ev, lambda m, xi: m(xi, ts), "evolution")
    seq_case(ev, lambda m, xi: m(xi, None), "evolution_no_ts")
    for agg in ("mean", "max", "sum"):
        torch.manual_seed(23)
        sk = ref.TemporalSkipConnection(hd, window_size=2 if agg == "max" else 3, aggregation=agg, dropout=0.0)
        _randomize(sk, 8)
        seq_case(sk, lambda m, xi: m(xi), "skip_" + agg, window=sk.window_size, aggregation=agg)

    torch.manual_seed(24)
    gu = ref.TemporalGatingUnit(hd, dropout=0.0)
    _randomize(gu, 9)
    gu.eval()
    cur = xs[2].clone().requires_grad_(True)
    prev = xs[3].clone().requires_grad_(True)
    with ref_loader.quiet():
        o = gu(cur, prev)
        (o * wout).sum().backward()
    out["gating"] = dict(cur=cur.detach().clone(), prev=prev.detach().clone(), wout=wout, sd=_sd(gu),
                         out=o.detach().clone(), dcur=cur.grad.clone(), dprev=prev.grad.clone(), grads=_grads(gu))

    # Propagation core: the reference's forward never completes, so run its
    # sub-modules in the order forward() would (temporal_propagation.py:1343-1349, 1487-1500).
    torch.manual_seed(25)
    tp = ref.TemporalPropagation(hd, hd, dropout=0.0)
    _randomize(tp, 10)
    tp.eval()

    def core(m, xi):
        e = m.evolution_layer(xi, ts)
        e = m.skip_connection(e)
        return [m.layer_norm(m.dropout_layer(m.output_proj(f))) for f in e]
    seq_case(tp, core, "propagation_core")
    return out


def bank_cases(ref):
    out = {}

    def dump(bank, cap):
        st = np.zeros((cap, bank.hidden_dim), np.float32)
        valid = np.zeros(cap, np.uint8)
        inact = np.zeros(cap, np.int32)
        last = np.full(cap, -1, np.int32)
        freq = np.zeros(cap, np.int32)
        for k, v in bank.node_states.items():
            st[k] = v.numpy()
            valid[k] = 1
        for k, v in bank.inactivity_counter.items():
            inact[k] = v
        for k, v in bank.last_seen.items():
            last[k] = v
        for k, v in bank.frequency.items():
            freq[k] = v
        return dict(states=st, valid=valid, inactivity=inact, last_seen=last, frequency=freq, size=bank.size)

    # KAT from SURVEY.md section 3.5
    bank = ref.NodeMemoryBank(2, decay_factor=0.8, max_inactivity=3)
    steps = [([7, 3, 9], [[1, 1], [2, 2], [3, 3]]), ([3], [[10, 10]]), ([3], [[20, 20]]),
             ([7, 3], [[5, 5], [30, 30]]), ([3], [[40, 40]])]
    trace = []
    for t, (ids, st) in enumerate(steps):
        bank.update(ids, torch.tensor(st, dtype=torch.float32), t)
        trace.append(dict(op="update", ids=ids, states=np.array(st, np.float32), t=t, after=dump(bank, 10)))
    out["kat"] = dict(hidden=2, decay=0.8, max_inactivity=3, cap=10, trace=trace)

    # random op sequence with duplicates, get_states insertions, update_state, decay_all
    rng = np.random.RandomState(0)
    cap, hd = 40, 8
    bank = ref.NodeMemoryBank(hd, decay_factor=0.7, max_inactivity=2)
    trace = []
    tstep = 0
    for it in range(60):
        r = rng.rand()
        if r < 0.6:
            m = rng.randint(1, 12)
            ids = rng.randint(0, cap, size=m).tolist()          # duplicates allowed
            st = rng.randn(m, hd).astype(np.float32)
            bank.update(ids, torch.from_numpy(st), tstep)
            trace.append(dict(op="update", ids=ids, states=st, t=tstep, after=dump(bank, cap)))
            tstep += int(rng.randint(1, 3))
        elif r < 0.8:
            m = rng.randint(1, 8)
            ids = rng.randint(0, cap, size=m).tolist()
            got = bank.get_states(ids).numpy().copy()
            trace.append(dict(op="get_states", ids=ids, got=got, after=dump(bank, cap)))
        elif r < 0.9:
            nid = int(rng.randint(0, cap))
            st = rng.randn(hd).astype(np.float32)
            bank.update_state(nid, torch.from_numpy(st), tstep)
            trace.append(dict(op="update_state", ids=[nid], states=st, t=tstep, after=dump(bank, cap)))
        else:
            bank.decay_all()
            trace.append(dict(op="decay_all", after=dump(bank, cap)))
    out["random"] = dict(hidden=hd, decay=0.7, max_inactivity=2, cap=cap, trace=trace)
    return out


def model_case(ref):
    """Whole ``TAGAN.forward`` on the example.py shapes (config 1) incl. its fallback behaviour."""
    out = []
    for learnable, heads in ((False, 4), (True, 4), (False, 5)):
        torch.manual_seed(0)
        np.random.seed(0)
        hidden = 64 if heads == 4 else 40
        cfg = ref.TAGANConfig(node_feature_dim=16, edge_feature_dim=8, hidden_dim=hidden, num_heads=heads,
                              num_layers=2, output_dim=1, dropout=0.0, loss_type="bce", use_edge_features=True,
                              learnable_distance=learnable, temporal_window_size=3)
        with ref_loader.quiet():
            model = ref.TAGAN(cfg)
        model.eval()
        seq = []
        for _ in range(5):                                       # example.py:23-65
            nt = int(np.random.randint(5, 11))
            x = torch.randn(nt, 16)
            ei = torch.randint(0, nt, (2, 2 * nt))
            ea = torch.randn(2 * nt, 8)
            ids = np.random.choice(10, nt, replace=False).tolist()
            seq.append((x, ei, ea, ids))
        labels = torch.tensor([[1.0]])
        with ref_loader.quiet():
            res = model(seq, labels)
            res["loss"].backward()
        out.append(dict(cfg=dict(node_feature_dim=16, edge_feature_dim=8, hidden_dim=hidden, num_heads=heads,
                                 num_layers=2, output_dim=1, dropout=0.0, loss_type="bce", use_edge_features=True,
                                 learnable_distance=learnable, temporal_window_size=3),
                        seq=seq, labels=labels, sd=_sd(model), logits=res["logits"].detach().clone(),
                        loss=res["loss"].detach().clone(), grads=_grads(model)))
    return out


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    torch.save(geo_cases(ref), os.path.join(OUT, "geo_attention.pt"))
    torch.save(tattn_cases(ref), os.path.join(OUT, "temporal_attention.pt"))
    torch.save(prop_cases(ref), os.path.join(OUT, "propagation.pt"))
    torch.save(bank_cases(ref), os.path.join(OUT, "memory_bank.pt"))
    torch.save(model_case(ref), os.path.join(OUT, "tagan_model.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
