#!/usr/bin/env python
"""Per-stage fwd+bwd timing of the TAGAN layer at a bench workload, for the three code paths:
   unfused (ops.FUSION = False), stage-fused with plain GEMMs (fused.FUSED_GEMM = False), stage-fused with GEMM epilogues.
CUDA events, median of `--reps`, one JSON line per (stage, mode) into gpurun_out/stages.jsonl."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--snapshots", type=int, default=0)
    ap.add_argument("--nodes", type=int, default=0, help="override the node count (edges scale with it)")
    ap.add_argument("--only", default="", help="comma-separated stage names")
    ap.add_argument("--out", default="gpurun_out/stages.jsonl")
    args = ap.parse_args()
    import tagan_b200
    from tagan_b200 import fused, ops, synth
    dev = torch.device("cuda:0")
    w = synth.WORKLOADS[args.workload]
    t_steps = args.snapshots or w.snapshots
    n, hdim = args.nodes or w.num_nodes, w.hidden
    n_edges = int(w.num_edges * n / w.num_nodes)
    torch.manual_seed(0)
    layer = tagan_b200.TAGANLayer(hdim, w.heads, "euclidean").to(dev)
    layer.geometric.validate_indices = False
    gen = torch.Generator().manual_seed(0)
    x3 = torch.randn(t_steps, n, hdim, device=dev)
    eis = [synth.random_edges(n, n_edges, gen, w.graph).to(dev) for _ in range(t_steps)]
    ts = torch.arange(t_steps, dtype=torch.float32, device=dev).expand(n, t_steps)
    csrs = [ops.build_csr(ei, n) for ei in eis]
    go = torch.randn(t_steps, n, hdim, device=dev)
    prop = layer.propagation

    stages = {
        "geometric": lambda x: layer.geometric.forward_seq(x, csrs),
        "evolution": lambda x: prop.evolution_layer.forward_stacked(x, ts),
        "skip": lambda x: prop.skip_connection.forward_stacked(x),
        "prop_tail": lambda x: prop._tail(x),
        "temporal_attention": lambda x: layer.temporal_attention(x, time_stamps=ts, time_major=True).permute(1, 0, 2),
        "loss": lambda x: fused.mean_square(x) if ops.FUSION else x.square().mean(),
    }
    modes = {"unfused": (False, False, False), "stage_fused_plain_gemm": (True, False, False),
             "stage_fused_gemm_epilogues": (True, True, False), "stage_fused_gemm_epilogues_fastmath": (True, True, True)}
    out = open(args.out, "w")
    only = [x for x in args.only.split(",") if x]
    for sname, fn in stages.items():
        if only and sname not in only:
            continue
        for mname, (fus, fg, fm) in modes.items():
            if mname in ("stage_fused_plain_gemm", "stage_fused_gemm_epilogues_fastmath"):
                continue
            ops.FUSION, fused.FUSED_GEMM, fused.EPI_FAST_MATH = fus, fg, fm
            times_f, times_b = [], []
            try:
                for rep in range(args.reps + 2):
                    x = x3.clone().requires_grad_(True)
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    y = fn(x)
                    e1.record()
                    if y.dim() == 0:
                        y.backward()
                    else:
                        y.backward(go.view(y.shape) if y.numel() == go.numel() else torch.ones_like(y))
                    e2.record()
                    torch.cuda.synchronize()
                    if rep >= 2:
                        times_f.append(e0.elapsed_time(e1))
                        times_b.append(e1.elapsed_time(e2))
                    del x, y
                rec = {"stage": sname, "mode": mname, "fwd_ms": statistics.median(times_f), "bwd_ms": statistics.median(times_b)}
            except Exception as exc:  # noqa: BLE001
                rec = {"stage": sname, "mode": mname, "error": str(exc).splitlines()[0][:200]}
            finally:
                ops.FUSION, fused.FUSED_GEMM, fused.EPI_FAST_MATH = True, True, False
            rec["total_ms"] = rec.get("fwd_ms", 0) + rec.get("bwd_ms", 0)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
