#!/usr/bin/env python
"""Node-partitioned geometric layer over NCCL (torchrun, one rank per GPU).

1. parity: every rank compares its slice of the partitioned layer (fwd + input/weight gradients) with the
   unpartitioned layer computed locally on the same seeded inputs;
2. timing on one snapshot of a named config (default c4: 1M nodes, 10M edges, H=256): CSR build,
   K|V all-gather, fused kernel fwd, bwd (row+col pass), dK|dV reduce-scatter -- CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_partitioned.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tagan_b200  # noqa: E402
from tagan_b200 import ops, partitioned, synth  # noqa: E402
from tagan_b200.dist import GradBucket, NodePartition  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--stage-snapshots", type=int, default=8, help="snapshots of the pipelined geometric-stage timing")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---------------- parity ----------------
    torch.manual_seed(0)
    n, e, hdim, heads = 4096, 60000, 128, 8
    layer = tagan_b200.GeometricAttention(hdim, heads, dropout=0.0, distance_metric="euclidean").to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, hdim, generator=g).to(dev)
    ei = torch.randint(0, n, (2, e), generator=g).to(dev)
    wout = torch.randn(n, hdim, generator=g).to(dev)
    part = NodePartition(n, world)
    lo, hi = part.bounds(rank)
    xf = x.clone().requires_grad_(True)
    out_full = layer.forward_csr(xf, ops.build_csr(ei, n))
    (out_full * wout).sum().backward()
    gfull = {k: p.grad.clone() for k, p in layer.named_parameters()}
    layer.zero_grad()
    comm = partitioned.TorchDistComm(part, rank)
    xl = x[lo:hi].clone().requires_grad_(True)
    csr = partitioned.build_csr_part(ei, part, rank)
    out_loc = partitioned.geometric_layer_part(layer, xl, csr, comm, n)
    (out_loc * wout[lo:hi]).sum().backward()
    bucket = GradBucket(list(layer.parameters()))
    bucket.all_reduce(world)                                   # mean over ranks
    ok_fwd = bool(torch.equal(out_loc.detach(), out_full.detach()[lo:hi]))
    err_dx = float((xl.grad - xf.grad[lo:hi]).abs().max())
    err_dw = max(float((p.grad * world - gfull[k]).abs().max() / max(1.0, float(gfull[k].abs().max())))
                 for k, p in layer.named_parameters())
    flags = torch.tensor([1.0 if ok_fwd else 0.0, err_dx, err_dw], device=dev)
    dist.all_reduce(flags[0:1], op=dist.ReduceOp.MIN)
    dist.all_reduce(flags[1:], op=dist.ReduceOp.MAX)
    parity = {"forward_bit_identical": bool(flags[0].item() == 1.0), "max_abs_err_dx": float(flags[1]),
              "max_rel_err_dparams": float(flags[2])}
    assert parity["forward_bit_identical"] and parity["max_abs_err_dx"] < 2e-5 and parity["max_rel_err_dparams"] < 1e-4, parity

    # ---------------- timing ----------------
    w = synth.WORKLOADS[a.workload]
    n, e, hdim, heads = w.num_nodes, w.num_edges, w.hidden, w.heads
    part = NodePartition(n, world)
    lo, hi = part.bounds(rank)
    g = torch.Generator().manual_seed(7)
    ei = synth.random_edges(n, e, g, w.graph).to(dev)          # same edge list on every rank
    qkv = (torch.randn(hi - lo, 3 * hdim, generator=torch.Generator().manual_seed(rank)) * 0.5).to(dev).requires_grad_(True)
    dctx = torch.randn(hi - lo, hdim, device=dev)
    comm = partitioned.TorchDistComm(part, rank)
    names = ["csr", "allgather", "fwd", "bwd_kernels", "reduce_scatter"]
    acc = {k: [] for k in names}
    lib_metric = "euclidean"
    for it in range(a.iters + 2):
        ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in names}
        dist.barrier()
        torch.cuda.synchronize()
        ev["csr"][0].record()
        csr = partitioned.build_csr_part(ei, part, rank)
        ev["csr"][1].record()
        ev["allgather"][0].record()
        kv = comm.all_gather_rows(qkv.detach()[:, hdim:])
        ev["allgather"][1].record()
        ev["fwd"][0].record()
        ctx = partitioned.geo_attention_core_part(qkv, csr, comm, heads, lib_metric, n)   # includes its own all-gather
        ev["fwd"][1].record()
        qkv.grad = None
        ev["bwd_kernels"][0].record()
        ctx.backward(dctx)                                                                # kernels + reduce-scatter
        ev["bwd_kernels"][1].record()
        full = torch.empty(n, 2 * hdim, device=dev)
        ev["reduce_scatter"][0].record()
        comm.reduce_scatter_rows(full)
        ev["reduce_scatter"][1].record()
        torch.cuda.synchronize()
        if it >= 2:
            for k in names:
                acc[k].append(ev[k][0].elapsed_time(ev[k][1]))
    med = torch.tensor([sorted(acc[k])[len(acc[k]) // 2] for k in names], device=dev)
    dist.all_reduce(med, op=dist.ReduceOp.MAX)
    del ctx, kv, full, csr, qkv, dctx
    mem = lambda tag: print(f"[mem r{rank}] {tag}: alloc {torch.cuda.memory_allocated() / 2**30:.1f} GiB, "
                            f"peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", file=sys.stderr, flush=True)
    mem("after micro timings")
    # ---------------- pipelined geometric stage over several snapshots (fwd + bwd, weights replicated) ----------------
    ts_n = a.stage_snapshots
    layer = tagan_b200.GeometricAttention(hdim, heads, dropout=0.0, distance_metric="euclidean").to(dev)
    gen = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(hi - lo, hdim, generator=gen).to(dev).requires_grad_(True) for _ in range(ts_n)]
    eis = [synth.random_edges(n, e, torch.Generator().manual_seed(50 + s_), w.graph).to(dev) for s_ in range(ts_n)]
    stage = {}
    keep = {}
    for mode in ("sequential", "pipelined"):
        times = []
        for it in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            csrs = [partitioned.build_csr_part(ei_, part, rank) for ei_ in eis]
            if mode == "pipelined":
                outs = partitioned.geometric_stage_part(layer, xs, csrs, comm, n)
            else:
                outs = [partitioned.geometric_layer_part(layer, x_, c_, comm, n) for x_, c_ in zip(xs, csrs)]
            loss = sum(o.square().mean() for o in outs)
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
            if it >= 1:
                times.append(e0.elapsed_time(e1))
            if it == 2:                                       # pipelined == sequential: outputs and input gradients
                keep[mode] = ([o.detach()[:4096].clone() for o in outs], [x_.grad[:4096].clone() for x_ in xs])
            layer.zero_grad()
            for x_ in xs:
                x_.grad = None
            del outs, loss, csrs
            mem(f"{mode} it{it}")
        tt = torch.tensor([min(times)], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stage[mode] = float(tt)
    same = all(torch.equal(p_, s_) for k_ in (0, 1) for p_, s_ in zip(keep["pipelined"][k_], keep["sequential"][k_]))
    flag = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    parity["pipelined_equals_sequential"] = bool(flag.item() == 1.0)
    assert parity["pipelined_equals_sequential"]
    if rank == 0:
        t = dict(zip(names, [float(v) for v in med]))
        # "fwd" includes one all-gather, "bwd_kernels" one reduce-scatter
        total = t["csr"] + t["fwd"] + t["bwd_kernels"]
        print(json.dumps({"world": world, "workload": w.name, "parity": parity, "ms": t,
                          "kernel_a_step_ms": total, "edge_snapshots_per_s": e / (total * 1e-3),
                          "halo_bytes_per_rank": (n - (hi - lo)) * 2 * hdim * 4,
                          "geometric_stage": {"snapshots": ts_n, "ms": stage,
                                              "edge_snapshots_per_s_pipelined": e * ts_n / (stage["pipelined"] * 1e-3),
                                              "edge_snapshots_per_s_sequential": e * ts_n / (stage["sequential"] * 1e-3)}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
