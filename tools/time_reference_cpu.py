#!/usr/bin/env python
"""Time the UNMODIFIED reference on its own CPU-runnable cases -- build container only
(needs /root/reference; torch CPU, all cores of this container, stdout of the reference's prints suppressed):

  * config 1: full ``TAGAN.forward`` + backward on the example.py shapes (T=5, N<=10, hidden 64, 4 heads, 2 layers);
  * config 2: one snapshot (10 000 nodes, ~200k power-law edges, hidden 128, 4 heads) through the dense
    ``TAGANGraphAttention``, scaled_dot_product (forward + backward) and euclidean (forward only: its backward is a Python loop
    over N x heads slices, > 100 s at N = 2000).

1 warm-up + median of 3.  Writes profiles/r02_reference_cpu_c1_c2.json."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402


def med(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def main():
    ref = ref_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"host": {"cores": torch.get_num_threads(), "torch": torch.__version__}, "protocol": "1 warm-up + median of 3"}
    # ---- config 1
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = ref.TAGANConfig(node_feature_dim=16, edge_feature_dim=8, hidden_dim=64, num_heads=4, num_layers=2, output_dim=1,
                          dropout=0.0, loss_type="bce", use_edge_features=True)
    with ref_loader.quiet():
        model = ref.TAGAN(cfg)
    seq = []
    for _ in range(5):
        nt = int(np.random.randint(5, 11))
        seq.append((torch.randn(nt, 16), torch.randint(0, nt, (2, 2 * nt)), torch.randn(2 * nt, 8),
                    np.random.choice(10, nt, replace=False).tolist()))
    labels = torch.tensor([[1.0]])
    units = sum(int(s[1].shape[1]) for s in seq)

    def c1():
        with ref_loader.quiet():
            model.zero_grad()
            model(seq, labels)["loss"].backward()
    t = med(c1)
    out["c1_full_model_fwd_bwd"] = {"seconds": t, "edge_snapshots": units, "edge_snapshots_per_s": units / t}
    # ---- config 2 snapshot, dense reference
    from tagan_b200.synth import random_edges
    g = torch.Generator().manual_seed(1)
    n, e, hidden, heads = 10_000, 200_000, 128, 4
    ei = random_edges(n, e, g, "powerlaw")
    x = torch.randn(n, hidden, generator=g)
    for metric, bwd in (("scaled_dot_product", True), ("euclidean", False)):
        layer = ref.TAGANGraphAttention(hidden, num_heads=heads, dropout=0.0, distance_metric=metric)

        def c2():
            with ref_loader.quiet():
                xr = x.clone().requires_grad_(bwd)
                o = layer(xr, ei)
                if bwd:
                    o.square().sum().backward()
        t = med(c2, reps=3 if bwd else 1)
        out[f"c2_snapshot_dense_{metric}_{'fwd_bwd' if bwd else 'fwd_only'}"] = {"seconds": t, "edge_snapshots": e,
                                                                                 "edge_snapshots_per_s": e / t}
        print(metric, t, flush=True)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_reference_cpu_c1_c2.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
