#!/usr/bin/env python
"""Kernel-(a) microbenchmark: device CSR build + fused geometric attention fwd + bwd on ONE snapshot
of a named config, timed with CUDA events (L2 flushed between iterations).  Also the command that is
run under ncu for profiles/ (short: a few launches).

    python tools/profile_geo.py --workload c4 [--iters 10] [--metric euclidean]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tagan_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--metric", default="euclidean")
    a = ap.parse_args()
    w = synth.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n, e, hdim, h = w.num_nodes, w.num_edges, w.hidden, w.heads
    ei = synth.random_edges(n, e, g, w.graph).to(dev)
    qkv = (torch.randn(n, 3 * hdim, generator=g) * 0.5).to(dev).requires_grad_(True)
    dctx = torch.randn(n, hdim, generator=g).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    t = {"csr": [], "fwd": [], "bwd": []}
    nnz = None
    for it in range(a.iters + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        flush.zero_()
        ev[0].record()
        csr = ops.build_csr(ei, n)
        ev[1].record()
        flush.zero_()
        ev[2].record()
        ctx, _ = ops.geo_attention_core(qkv, csr, h, a.metric)
        ev[3].record()
        flush.zero_()
        qkv.grad = None
        ev[4].record()
        ctx.backward(dctx)
        ev[5].record()
        torch.cuda.synchronize()
        if nnz is None:
            nnz = csr.nnz
        if it >= 3:
            t["csr"].append(ev[0].elapsed_time(ev[1]))
            t["fwd"].append(ev[2].elapsed_time(ev[3]))
            t["bwd"].append(ev[4].elapsed_time(ev[5]))
    med = {k: sorted(v)[len(v) // 2] for k, v in t.items()}
    b_fwd = nnz * (2 * hdim * 4 + 4) + n * (2 * hdim * 4 + 8 * h + 8)
    b_bwd = nnz * (4 * hdim * 4 + 12) + n * (6 * hdim * 4 + 8 * h + 8)
    b_csr = e * 16 * 2 + (e + n) * 4 * 10        # two passes over int64 pairs + ~10 int32 passes over the entries
    out = {"workload": w.name, "metric": a.metric, "nnz": nnz, "ms": med,
           "fwd_GBs": b_fwd / med["fwd"] / 1e6, "bwd_GBs": b_bwd / med["bwd"] / 1e6, "csr_GBs": b_csr / med["csr"] / 1e6,
           "fwd_frac": b_fwd / med["fwd"] / 1e6 / peak, "bwd_frac": b_bwd / med["bwd"] / 1e6 / peak,
           "edge_snapshots_per_s_kernel_a": e / ((med["fwd"] + med["bwd"]) * 1e-3), "peak_GBs": peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
