#!/usr/bin/env python
"""Time the temporal-attention kernels (a3) alone at a named workload's shape: B = nodes, T = snapshots.

    python tools/profile_tattn.py --workload c3 [--time-major 1]

Prints one JSON line: per-launch ms (CUDA events, L2 flushed between launches) and the fraction of the measured
HBM peak given the algorithmic bytes (fwd: read q,k,v + write ctx,lse; bwd: read q,k,v,ctx,dctx,lse + write dq,dk,dv).
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
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--time-major", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--nodes", type=int, default=0, help="override the node count (config 5 is benchmarked on a 25k-node sample)")
    ap.add_argument("--mask", default="causal", choices=["causal", "band"],
                    help="band: shared integer timestamps 0..T-1 with the reference's +-10 band (what the bench layer resolves to)")
    a = ap.parse_args()
    w = synth.WORKLOADS[a.workload]
    b, t, h, heads = a.nodes or w.num_nodes, w.snapshots, w.hidden, w.heads
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    qkv = (torch.randn(b * t, 3 * h, device=dev) * 0.5).requires_grad_(True)
    bias = (torch.randn(heads, t, t, device=dev) * 0.1).requires_grad_(True)
    tmask = ops.TemporalMask(flags=1)
    if a.mask == "band":
        ts = torch.arange(t, dtype=torch.float32, device=dev).repeat(b, 1).contiguous()
        tmask = ops.TemporalMask(flags=2 | 4, ts=ts, allones_flag=ops.mask_allones_flag(ts, None, b, t, 10.0, dev))
    dctx = torch.randn(b * t, h, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf, tb = [], []
    for it in range(a.iters + 2):
        qkv.grad = None
        bias.grad = None
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        ctx, _ = ops.temporal_attention_core(qkv, bias, tmask, b, t, heads, bool(a.time_major), False)
        e[1].record()
        flush.zero_()
        e[2].record()
        ctx.backward(dctx)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[2].elapsed_time(e[3]))
    rows = b * t
    bytes_f = rows * h * 4 * 4 + rows * heads * 4
    bytes_b = rows * h * 4 * 8 + rows * heads * 4
    peak = 6458.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    mf, mb = min(tf), min(tb)
    print(json.dumps({"workload": w.name, "B": b, "T": t, "H": h, "heads": heads, "time_major": a.time_major,
                      "fwd_ms": mf, "bwd_ms": mb, "fwd_frac_hbm": bytes_f / (mf * 1e-3) / 1e9 / peak,
                      "bwd_frac_hbm": bytes_b / (mb * 1e-3) / 1e9 / peak,
                      "note": "bwd includes autograd glue (dqkv allocation, bias-gradient reduce)"}))


if __name__ == "__main__":
    main()
