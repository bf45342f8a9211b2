#!/usr/bin/env python
"""Benchmark of the TAGAN hot path: edge-snapshots/s of one "TAGAN layer" forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = geometric attention layer over all T snapshots (device CSR build included) + propagation
core + temporal attention + memory-bank gather/update per snapshot, forward and backward, fp32,
dropout 0, on synthetic temporal graphs of the named config.  Multi-GPU is
data-parallel: one graph sequence per GPU, replicated weights, one NCCL all-reduce of the gradients
per step (weak scaling).  Prints ONE JSON line (rank 0).

`--impl reference` times the CPU oracle port of the same path (oracle/restate.py, pinned against the
unmodified reference by tests/golden) on all host cores, on a bounded sample of the same workload:
the reference's own dense N x N implementation cannot allocate this config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "edge-snapshots/sec TAGAN layer fwd+bwd"
UNIT = "edge-snapshots/s"
CPU_SAMPLE_SNAPSHOTS = 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--metric", default="euclidean", help="DistanceMetric of the geometric layer (reference default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the kernel arm with eager launches instead of graph replays")
    ap.add_argument("--e2e-eager", action="store_true", help="e2e arm without the CUDA-graph capture")
    ap.add_argument("--e2e-no-prefetch", action="store_true",
                    help="e2e arm: H2D copies inside the step's graph instead of prefetching the next step's inputs")
    ap.add_argument("--no-bank", action="store_true")
    ap.add_argument("--qkv-storage", default="fp32", choices=["fp32", "bf16"],
                    help="bf16: the geometric layer stores its projected q/k/v rows in bf16 (fp32 arithmetic; separate tolerance)")
    ap.add_argument("--c5-nodes", type=int, default=25_000, help="node count of the config-5 sample (T stays 128)")
    ap.add_argument("--no-partitioned", action="store_true",
                    help="N >= 2: skip the node-partitioned config-4 block that follows the data-parallel measurement")
    ap.add_argument("--cuda-profiler", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------
def cpu_sample_step_factory(w, metric, seed=0):
    """Bounded sample of the workload for the CPU port: same N, E, H, heads; T = 2 snapshots."""
    import contextlib
    import io
    from oracle import restate as R
    from tagan_b200 import synth
    import tagan_b200
    torch.set_num_threads(os.cpu_count() or 1)
    xs, eis, ts = synth.make_sequence(w, seed=seed, snapshots=CPU_SAMPLE_SNAPSHOTS)
    torch.manual_seed(0)
    layer = tagan_b200.TAGANLayer(w.hidden, w.heads, metric)
    sdg = {k: v.detach().clone().requires_grad_(True) for k, v in layer.geometric.geometric_attention.state_dict().items()}
    sdp = {k: v.detach().clone().requires_grad_(True) for k, v in layer.propagation.state_dict().items()}
    sdt = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in layer.temporal_attention.state_dict().items()}

    def step():
        with contextlib.redirect_stdout(io.StringIO()):
            xr = [x.clone().requires_grad_(True) for x in xs]
            csrs = [R.build_csr(e, w.num_nodes) for e in eis]
            out = R.tagan_layer(xr, eis, ts, sdg, sdp, sdt, w.heads, metric, csrs=csrs)
            out.square().mean().backward()
    units = w.num_edges * CPU_SAMPLE_SNAPSHOTS
    desc = (f"{w.name}: same N={w.num_nodes}, E={w.num_edges}/snapshot, H={w.hidden}, h={w.heads}, "
            f"T={CPU_SAMPLE_SNAPSHOTS} of {w.snapshots} snapshots; CSR build + layer fwd+bwd, oracle/restate.py (sparse "
            f"torch CPU port of the reference semantics)")
    return step, units, desc


def run_reference(args):
    from tagan_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c1":
        print(json.dumps({"impl": "reference", "unavailable": "config 1 is the whole-model case; the CPU arm of this bench times the "
                          "layer-level oracle port only (the reference's own TAGAN on config 1, timed in the build container, is in "
                          "profiles/r02_reference_cpu_c1_c2.json)"}), flush=True)
        return
    w = synth.WORKLOADS[args.workload]
    if args.workload == "c5":
        import dataclasses
        w = dataclasses.replace(w, num_nodes=args.c5_nodes, num_edges=int(w.num_edges * args.c5_nodes / w.num_nodes),
                                name=w.name + f" (node sample: {args.c5_nodes} of {w.num_nodes} nodes, same average degree)")
    step, units, desc = cpu_sample_step_factory(w, args.metric)
    for _ in range(min(args.warmup, 1)):          # CPU path: one warm-up is enough to fault pages in
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    cores = torch.get_num_threads()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": w.name, "distance_metric": args.metric, "sample": desc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"



# ------------------------------------------------------------------------------------------
# single large graph, node-partitioned (config 4): snapshot-parallel geometric stage + node-parallel temporal stages
# ------------------------------------------------------------------------------------------
def partitioned_snapshots(w, world):
    """Largest T <= w.snapshots with T % world == 0 that fits 180 GB per rank: the saved activations are about
    8.3 [N,H] tensors per local snapshot (geometric) + 35 [T,N/world,H] tensors (temporal stages)."""
    t = w.snapshots
    while t > world:
        per_rank = (t / world * 8.3 + t / world * 35) * w.num_nodes * w.hidden * 4 / 1e9
        if t % world == 0 and per_rank < 120:
            return t
        t -= 1
    return world


def run_partitioned(w, metric, world, rank, dev, steps, warmup, no_bank=False, want_e2e=True, qkv_storage="fp32"):
    """One TAGAN layer fwd+bwd on ONE graph of w.num_nodes nodes over `world` GPUs (tagan_b200.partitioned.
    forward_snapshot_parallel).  Inputs are generated on the device (16 GB of host randn would dominate the run)."""
    import torch.distributed as dist
    import tagan_b200
    from tagan_b200 import fused, ops, partitioned
    from tagan_b200.dist import FlatGradBucket, NodePartition
    n, e, hdim, heads = w.num_nodes, w.num_edges, w.hidden, w.heads
    t_steps = partitioned_snapshots(w, world)
    t_loc = t_steps // world
    part = NodePartition(n, world)
    lo, hi = part.bounds(rank)
    n_loc = hi - lo
    comm = partitioned.AllToAllComm(world)

    # parity on a small graph first: partitioned == unpartitioned (forward bit-identical on every rank)
    torch.manual_seed(0)
    par = None
    if world > 1:
        pn, pe, ph, pheads, pt = 64 * world, 3000, 64, 4, 2 * world
        small = tagan_b200.TAGANLayer(ph, pheads, metric).to(dev)
        g = torch.Generator().manual_seed(5)
        pxs = torch.randn(pt, pn, ph, generator=g).to(dev)
        peis = [torch.randint(0, pn, (2, pe), generator=g).to(dev) for _ in range(pt)]
        pts = torch.arange(pt, dtype=torch.float32, device=dev).expand(pn, pt)
        full = small(list(pxs.unbind(0)), peis, pts)
        plo, phi = NodePartition(pn, world).bounds(rank)
        loc = partitioned.forward_snapshot_parallel(small, pxs[:, plo:phi].contiguous(), peis[rank * 2:(rank + 1) * 2],
                                                    NodePartition(pn, world), rank, comm, pts[plo:phi])
        same = torch.tensor([1.0 if torch.equal(loc, full[plo:phi]) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        par = {"forward_bit_identical_to_unpartitioned": bool(same.item() == 1.0), "nodes": pn, "snapshots": pt}
        del small, full, loc

    torch.manual_seed(0)
    layer = tagan_b200.TAGANLayer(hdim, heads, metric).to(dev)
    layer.geometric.validate_indices = False
    layer.geometric.geometric_attention.qkv_storage = qkv_storage
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    x_loc = torch.randn(t_steps, n_loc, hdim, device=dev, generator=gen)
    my_eis = []
    for t in range(rank * t_loc, (rank + 1) * t_loc):                # the global edge list of snapshot t is a function of t only
        ge = torch.Generator(device=dev).manual_seed(7000 + t)
        my_eis.append(torch.randint(0, n, (2, e), device=dev, generator=ge))
    ts_loc = torch.arange(t_steps, dtype=torch.float32, device=dev).expand(n_loc, t_steps)
    bank = None
    if not no_bank:
        bank = tagan_b200.NodeMemoryBank(hdim, 0.8, 3, device=dev, capacity=n_loc)
        bank.check_range = False
    bucket = FlatGradBucket(list(layer.parameters()))

    def step(x, eis):
        bucket.zero()
        out = partitioned.forward_snapshot_parallel(layer, x, eis, part, rank, comm, ts_loc, bank)
        loss = fused.mean_square(out.permute(1, 0, 2))
        loss.backward()
        if world > 1:
            bucket.all_reduce(world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step(x_loc, my_eis)
    barrier()
    torch.cuda.reset_peak_memory_stats()
    ops.CALLS["n"] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step(x_loc, my_eis)
    ev1.record()
    barrier()
    launches = ops.CALLS["n"]
    tms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / steps
    units = e * t_steps
    res = {"workload": w.name, "nodes": n, "edges_per_snapshot": e, "snapshots": t_steps, "snapshots_of": w.snapshots,
           "hidden": hdim, "heads": heads, "ms_per_step": ms_step, "value": units / (ms_step * 1e-3), "unit": UNIT,
           "mode": "snapshot-parallel geometric stage (%d snapshots per GPU, full CSR, no halo) + 2 all-to-alls per direction "
                   "+ node-parallel GRU / skip / temporal attention / bank (%d nodes per GPU); weight-gradient all-reduce; "
                   "eager launches" % (t_loc, n_loc),
           "a2a_bytes_per_rank_per_exchange": (world - 1) * t_loc * n_loc * hdim * 4,
           "peak_mem_gb_per_rank": torch.cuda.max_memory_allocated() / 1e9, "gpu_launches": launches, "parity": par}
    # whole-layer byte budget of SURVEY.md section 8d (same formula as the data-parallel line), summed over all ranks
    nnz = e + n                                                       # upper bound (duplicates are rare on a uniform graph)
    bytes_a = nnz * (6 * hdim * 4 + 16) + n * (8 * hdim * 4 + 16 * heads + 16)
    wl = t_steps * bytes_a + 73 * hdim * 4 * n * t_steps + 11 * t_steps * hdim * 4 * n
    if not no_bank:
        wl += t_steps * n * ((2 * hdim * 4 + 4) + (3 * hdim * 4 + 16) + (2 * hdim * 4 + 8))
    peak, _ = peaks()
    res["whole_layer_roofline"] = {"algorithmic_gb_per_step": wl / 1e9, "achieved_gbs_all_ranks": wl / 1e9 / (ms_step * 1e-3),
                                   "frac_of_n_gpus_hbm": wl / 1e9 / (ms_step * 1e-3) / (peak * world)}
    if want_e2e:
        # end to end: this rank's inputs start in pinned host memory every step (H2D inside the timed region), loss read back
        xh = torch.empty(x_loc.shape, dtype=torch.float32).pin_memory()
        xh.copy_(x_loc)
        eh = [torch.empty(ei.shape, dtype=torch.int64).pin_memory() for ei in my_eis]
        for a, b in zip(eh, my_eis):
            a.copy_(b)
        xd, ed = torch.empty_like(x_loc), [torch.empty_like(ei) for ei in my_eis]

        def e2e_step():
            xd.copy_(xh, non_blocking=True)
            for a, b in zip(ed, eh):
                a.copy_(b, non_blocking=True)
            return float(step(xd, ed).item())
        e2e_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(steps):
            last = e2e_step()
        g1.record()
        barrier()
        ems = torch.tensor([g0.elapsed_time(g1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = xh.numel() * 4 + sum(a.numel() * 8 for a in eh)
        res["e2e"] = {"value": units / (float(ems.item()) / steps * 1e-3), "unit": UNIT, "ms_per_step": float(ems.item()) / steps,
                      "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world, "loss": last,
                      "mode": "eager: per-rank H2D of its node slice (all snapshots) and of its snapshots' edge lists, step, loss D2H"}
    del layer, x_loc, my_eis, bank
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------
# config 1: the reference's own example.py case -- the WHOLE model (TAGAN.forward + loss + backward + clip + Adam)
# ------------------------------------------------------------------------------------------
def example_sequence(seed=0):
    """example.py:23-65: T = 5 snapshots of 5..10 nodes, 16 node features, 2 N_t edges, 8 edge features."""
    import numpy as np
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    seq = []
    for _ in range(5):
        nt = int(rng.randint(5, 11))
        seq.append((torch.randn(nt, 16, generator=g), torch.randint(0, nt, (2, 2 * nt), generator=g), torch.randn(2 * nt, 8, generator=g),
                    rng.choice(10, nt, replace=False).tolist()))
    return seq


def run_c1(args, rank, world, dev):
    """One training step of the whole model on the example.py shapes (example.py:131-159: hidden 64, 4 heads, 2 layers,
    output_dim 1, BCE), captured as one CUDA graph (tagan_b200.head.TrainStep).  Units = raw edges of the sequence."""
    import tagan_b200
    from tagan_b200 import ops
    from tagan_b200.head import TrainStep
    cfg = dict(node_feature_dim=16, edge_feature_dim=8, hidden_dim=64, num_heads=4, num_layers=2, output_dim=1, dropout=0.0,
               loss_type="bce", use_edge_features=True, learnable_distance=False, temporal_window_size=3)
    seq = example_sequence(rank)
    units = sum(int(s[1].shape[1]) for s in seq)
    torch.manual_seed(0)
    model = tagan_b200.TAGANModel(cfg).to(dev).eval()
    opt = tagan_b200.FusedAdam(list(model.parameters()), lr=1e-3, max_grad_norm=1.0)
    host = tagan_b200.PackedSequence.from_snapshots([s[0] for s in seq], [s[1] for s in seq], pin=True)
    packed = host.to(dev)
    labels = torch.tensor([[1.0]], device=dev)
    step = TrainStep(model, opt)
    ops.CALLS["n"] = 0
    step.eager(packed, labels)
    launches = ops.CALLS["n"]
    step.capture(packed, labels)
    for _ in range(max(args.warmup, 3)):
        step.replay()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step.replay()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    # e2e: the packed sequence starts in pinned host memory every step (4 copies), the loss is read back
    loss_h = torch.zeros((), dtype=torch.float32).pin_memory()

    def e2e_step():
        for d, h in ((packed.x, host.x), (packed.edges, host.edges)):
            d.copy_(h, non_blocking=True)
        loss_h.copy_(step.replay(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_h)
    e2e_step()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        last = e2e_step()
    g1.record()
    torch.cuda.synchronize()
    ems = g0.elapsed_time(g1) / args.steps
    return {"metric": METRIC, "value": units * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c1-example-toy (whole model: TAGAN.forward + BCE + backward + clip_grad_norm + Adam)",
                       "snapshots": 5, "nodes": [int(s[0].shape[0]) for s in seq], "edges_per_sequence": units, "hidden": 64,
                       "heads": 4, "layers": 2, "parallelism": f"dp{world}", "l2": "working set of a few hundred KB: launch-latency "
                       "bound, L2-resident by construction (the reference's own CPU-runnable case)",
                       "timing": "one CUDA graph replay per step"},
            "roofline": {"bound": "launch latency (tiny tensors)", "achieved": None, "peak": None, "unit": "GB/s", "frac": None,
                         "traffic": None, "kernels_per_step": launches},
            "cpu_baseline": None,
            "e2e": {"value": units * world / (ems * 1e-3), "unit": UNIT, "ms_per_step": ems, "h2d_bytes_per_step": host.x.numel() * 4 +
                    host.edges.numel() * 8, "d2h_bytes_per_step": 4, "loss": last,
                    "mode": "pinned packed sequence -> 2 H2D copies -> graph replay (fwd + loss + bwd + clip + Adam) -> loss D2H"},
            "gpu_launches": launches * args.steps}


def run_ours(args):
    import torch.distributed as dist
    import tagan_b200
    from tagan_b200 import fused, ops, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the TAGAN hot path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "c1":
        line = run_c1(args, rank, world, dev)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    w = synth.WORKLOADS[args.workload]
    if args.workload == "c5":
        # long horizon (T = 128): [T,N,H] activations of the full 250k-node graph are 16 GB each, the layer keeps ~35 of them
        # for backward -- run the largest node count that fits (same degree, same T) and say so
        import dataclasses
        full_n = w.num_nodes
        n_fit = args.c5_nodes
        w = dataclasses.replace(w, num_nodes=n_fit, num_edges=int(w.num_edges * n_fit / full_n),
                                name=w.name + f" (node sample: {n_fit} of {full_n} nodes, same average degree)")
    if args.workload == "c4":
        # the north-star configuration: ONE graph of 1M nodes, node-partitioned (does not fit one GPU at T=16: runs the
        # largest snapshot count that does, and says so)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        res = run_partitioned(w, args.metric, world, rank, dev, args.steps, args.warmup, args.no_bank, not args.no_e2e,
                              args.qkv_storage)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32" if args.qkv_storage == "fp32" else "f32 arithmetic, q/k/v rows stored bf16",
                    "data": "synthetic (generated on device)",
                    "config": {"workload": w.name, "nodes": res["nodes"], "edges_per_snapshot": res["edges_per_snapshot"],
                               "snapshots": res["snapshots"], "snapshots_of": res["snapshots_of"], "hidden": res["hidden"],
                               "heads": res["heads"], "distance_metric": args.metric, "parallelism": res["mode"],
                               "memory_bank": not args.no_bank,
                               "l2": "inputs larger than L2 (K and V of one snapshot are 1 GB each)"},
                    "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peaks()[0], "achieved": res["whole_layer_roofline"]["achieved_gbs_all_ranks"] / world,
                                 "frac": res["whole_layer_roofline"]["frac_of_n_gpus_hbm"], "traffic": None,
                                 "kernel": "whole layer, per GPU (SURVEY 8d byte budget / step time / n_gpus)",
                                 "whole_layer": res["whole_layer_roofline"]},
                    "cpu_baseline": None, "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"], "clocks": clocks,
                    "node_partitioned": res}
            print(json.dumps(line), flush=True)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        return
    t_steps, n, e, hdim = w.snapshots, w.num_nodes, w.num_edges, w.hidden

    # synthetic inputs: pinned host copies (e2e arm) and resident device copies (kernel arm)
    xs_h, eis_h, ts_h = synth.make_sequence(w, seed=rank, pin=True)
    xs_d = [x.to(dev, non_blocking=True) for x in xs_h]
    eis_d = [ei.to(dev, non_blocking=True) for ei in eis_h]
    ts_d = torch.arange(t_steps, dtype=torch.float32, device=dev).expand(n, t_steps)   # shared timestamps (stride 0)
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    torch.manual_seed(0)
    layer = tagan_b200.TAGANLayer(hdim, w.heads, args.metric).to(dev)
    layer.geometric.geometric_attention.qkv_storage = args.qkv_storage
    bank = None
    if not args.no_bank:
        bank = tagan_b200.NodeMemoryBank(hdim, 0.8, 3, device=dev, capacity=n)
        bank.check_range = False
    from tagan_b200.dist import FlatGradBucket
    # every p.grad is a view of one flat buffer: backward accumulates in place, the all-reduce needs no packing
    bucket = FlatGradBucket(list(layer.parameters()))

    def fwd_bwd(xs, eis):
        bucket.zero()
        out = layer(xs, eis, ts_d, bank=bank, node_ids=[ids] * t_steps)       # [N,T,H] view of time-major storage
        # mean of squares, taken in the storage order of `out` (a permutation of the same elements) so that the
        # loss and its gradient are contiguous element-wise passes instead of strided ones
        loss = fused.mean_square(out.permute(1, 0, 2))
        loss.backward()
        return loss

    def step(xs, eis):
        loss = fwd_bwd(xs, eis)
        if world > 1:                       # data-parallel: one flat NCCL all-reduce of the gradients
            bucket.all_reduce(world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(xs_d, eis_d)
    barrier()

    # ---- kernel arm: inputs resident in HBM --------------------------------------------------
    nnz = [int(ops.build_csr(ei, n, transpose=False).rowptr[-1].item()) for ei in eis_d]

    # (1) eager pass of the K steps with every library call bracketed by CUDA events: per-kernel durations for the
    #     roofline block and the kernel shares (the bracketing and ~1000 Python-issued launches per step make this
    #     pass host-bound on slow host cores, so it is not the throughput measurement)
    ops.PROFILE = {}
    ops.CALLS["n"] = 0
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        step(xs_d, eis_d)
    ev1.record()
    barrier()
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    prof, ops.PROFILE = ops.PROFILE, None
    launches = ops.CALLS["n"]
    eager_ms_step = ev0.elapsed_time(ev1) / args.steps

    # (2) timed region: the same K steps with the same resident inputs, each step replayed as one CUDA graph (the
    #     sync-free step captured once; the data-parallel all-reduce stays an eager NCCL call after each replay)
    graph, timing_mode = None, "eager launches"
    if not args.eager:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fwd_bwd(xs_d, eis_d)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fwd_bwd(xs_d, eis_d)
            timing_mode = "cuda graph replay (one graph launch per step)"
        except Exception as exc:
            graph, timing_mode = None, "eager launches (graph capture failed: %s)" % (str(exc).splitlines()[0][:120],)
            torch.cuda.synchronize()

    def timed_step():
        if graph is None:
            step(xs_d, eis_d)
        else:
            graph.replay()
            if world > 1:
                bucket.all_reduce(world)

    timed_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        timed_step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    if graph is not None:
        graph.reset()
        graph = None
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    units_per_step = e * t_steps * world
    value = units_per_step / (ms_step * 1e-3)

    # ---- live per-kernel timing and roofline of the dominant kernel ---------------------------
    kt = {k: [s.elapsed_time(t) for s, t, _ in v] for k, v in prof.items()}
    kbytes = {k: sum(b for _, _, b in v) for k, v in prof.items()}
    share = {k: sum(v) / (ms_step * args.steps) for k, v in kt.items()}
    h = w.heads
    mean_nnz = sum(nnz) / len(nnz)
    bytes_fwd = mean_nnz * (2 * hdim * 4 + 4) + n * (2 * hdim * 4 + 8 * h + 8)
    bytes_bwd = mean_nnz * (4 * hdim * 4 + 12) + n * (6 * hdim * 4 + 8 * h + 8)
    peak, peak_src = peaks()
    t_fwd = statistics.mean(kt["geo_attn_fwd"]) * 1e-3
    t_bwd = statistics.mean(kt["geo_attn_bwd"]) * 1e-3
    # snapshots one launch covers: T with the block-diagonal CSR (one launch per pass), 1 with per-snapshot launches
    snaps_fwd = t_steps * args.steps / max(1, len(kt["geo_attn_fwd"]))
    snaps_bwd = t_steps * args.steps / max(1, len(kt["geo_attn_bwd"]))
    launch_bytes_fwd, launch_bytes_bwd = bytes_fwd * snaps_fwd, bytes_bwd * snaps_bwd
    ach_bwd = launch_bytes_bwd / t_bwd / 1e9
    ach_fwd = launch_bytes_fwd / t_fwd / 1e9
    roofline = {"kernel": "geo_attn_bwd (row pass + column pass)", "bound": "hbm", "achieved": ach_bwd, "peak": peak,
                "unit": "GB/s", "frac": ach_bwd / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": launch_bytes_bwd, "avg_launch_ms": t_bwd * 1e3,
                "snapshots_per_launch": snaps_bwd, "algorithmic_bytes_per_snapshot": bytes_bwd,
                "geo_attn_fwd": {"achieved": ach_fwd, "frac": ach_fwd / peak, "algorithmic_bytes_per_launch": launch_bytes_fwd,
                                 "avg_launch_ms": t_fwd * 1e3, "snapshots_per_launch": snaps_fwd},
                "share_of_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
                "timed_in": "eager pass of the same K steps (CUDA events around every library call) run immediately "
                            "before the timed region; shares are relative to the timed step; csr_build runs on a side "
                            "stream concurrently with other kernels, so its bracket over-states its share"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file)).get(args.workload, {})
            per_snap = tr.get("geo_attn_bwd_per_snapshot")
            roofline["traffic"] = per_snap * snaps_bwd if per_snap is not None else tr.get("geo_attn_bwd")
            if tr.get("geo_attn_fwd_per_snapshot") is not None:
                roofline["geo_attn_fwd"]["traffic"] = tr["geo_attn_fwd_per_snapshot"] * snaps_fwd
            roofline["traffic_source"] = tr.get("source", "profiles/traffic.json (ncu --set full, dram__bytes_read+write per launch)")
        except Exception:
            pass
    # all projections (tcgen05 GEMMs, fused epilogues included): algorithmic bytes of every call / summed event time
    if kt.get("gemm"):
        g_ms = sum(kt["gemm"]) / args.steps
        g_gb = kbytes["gemm"] / args.steps / 1e9
        roofline["gemm"] = {"bound": "hbm (K <= 256 projections stream their operands)", "calls_per_step": len(kt["gemm"]) // args.steps,
                            "algorithmic_gb_per_step": g_gb, "ms_per_step": g_ms, "achieved": g_gb / (g_ms * 1e-3),
                            "frac": g_gb / (g_ms * 1e-3) / peak, "unit": "GB/s"}
    # whole layer against the SURVEY section 8d byte budget: kernel (a) + node stream (73*H*4 B per node-snapshot)
    # + kernel (b) core (11*T*H*4 B per node) + memory bank (gather, update, sweep), per step
    s4 = 4
    wl_bytes = (t_steps * (bytes_fwd + bytes_bwd) + 73 * hdim * s4 * n * t_steps + 11 * t_steps * hdim * s4 * n)
    if not args.no_bank:
        wl_bytes += t_steps * n * ((2 * hdim * s4 + 4) + (3 * hdim * s4 + 16) + (2 * hdim * s4 + 8))
    roofline["whole_layer"] = {"algorithmic_gb_per_step": wl_bytes / 1e9, "ms_per_step": ms_step,
                               "achieved": wl_bytes / 1e9 / (ms_step * 1e-3), "frac": wl_bytes / 1e9 / (ms_step * 1e-3) / peak,
                               "unit": "GB/s", "budget": "SURVEY.md section 8d: kernel (a) fwd+bwd + 73*H*4 B node stream per "
                               "node-snapshot + 11*T*H*4 B kernel (b) per node + bank"}

    # ---- e2e arm: host buffers, H2D copies and D2H loss read inside the timed region ----------
    e2e = None
    if not args.no_e2e:
        h2d = sum(x.numel() * 4 for x in xs_h) + sum(ei.numel() * 8 for ei in eis_h)

        copy_stream = torch.cuda.Stream(device=dev)

        def e2e_step():
            # per-snapshot H2D copies run on a copy stream; the compute stream waits for snapshot t only when the
            # geometric layer reaches it, so the 1.3 GB of input copies overlap with the kernels of earlier snapshots
            main = torch.cuda.current_stream()
            copy_stream.wait_stream(main)
            xs, eis, evs = [], [], []
            with torch.cuda.stream(copy_stream):
                for x, ei in zip(xs_h, eis_h):
                    xs.append(x.to(dev, non_blocking=True))
                    eis.append(ei.to(dev, non_blocking=True))
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    evs.append(ev)
            for x, ei in zip(xs, eis):
                x.record_stream(main)
                ei.record_stream(main)

            class _Lazy:
                """list whose items make the compute stream wait for their copy on first access"""
                def __init__(self, items):
                    self.items = items

                def __iter__(self):
                    for t, it in enumerate(self.items):
                        main.wait_event(evs[t])
                        yield it

                def __len__(self):
                    return len(self.items)
            return float(step(_Lazy(xs), _Lazy(eis)).item())
        mode = "eager (copy stream + per-snapshot events)"
        run_step = e2e_step
        gstep = None
        if not args.e2e_eager:
            # the sync-free step captured once as a CUDA graph: H2D copies, forward, loss, backward, D2H of the loss
            try:
                gstep = tagan_b200.GraphedStep(fwd_bwd, xs_h, eis_h, dev, prefetch=not args.e2e_no_prefetch,
                                               after_replay=(lambda: bucket.all_reduce(world)) if world > 1 else None)
                run_step = gstep
                mode = ("cuda graph (fwd + bwd + D2H loss); every step copies its inputs H2D from pinned memory into a "
                        "staging set on a copy stream while the previous step's graph runs (loader-style prefetch), "
                        "then moves them device-to-device into the graph's static inputs"
                        if not args.e2e_no_prefetch else
                        "cuda graph (H2D copies + fwd + bwd + D2H loss in one graph launch per step, no prefetch)")
            except Exception as exc:            # capture unsupported in this environment: measure the eager path
                mode = "eager (graph capture failed: %s)" % (str(exc).splitlines()[0][:120],)
                torch.cuda.synchronize()
        run_step()
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            last_loss = run_step()
        g1.record()
        barrier()
        ems = torch.tensor([g0.elapsed_time(g1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": units_per_step / (float(ems.item()) / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": float(ems.item()) / args.steps,
               "wall_ms_per_step": (time.perf_counter() - t0) / args.steps * 1e3, "mode": mode, "loss": last_loss}
        if gstep is not None:
            gstep.release()
            gstep = run_step = None

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on this box's host cores ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cstep, cunits, cdesc = cpu_sample_step_factory(w, args.metric)
        cstep()                                                  # warm-up (page faults, thread pool)
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            cstep()
            times.append(time.perf_counter() - t0)
        cdt = statistics.median(times)
        cpu = {"value": cunits / cdt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": cdesc,
               "seconds": cdt, "protocol": "1 warm-up + median of 3"}

    # ---- N >= 2: the node-partitioned single-graph mode (config 4) on the snapshot count that fits, in the same line ----
    node_part = None
    if world > 1 and not args.no_partitioned and args.workload == "c3":
        del layer, xs_d, eis_d, xs_h, eis_h, bank
        torch.cuda.empty_cache()
        try:
            node_part = run_partitioned(synth.WORKLOADS["c4"], args.metric, world, rank, dev, min(args.steps, 3), 3,
                                        args.no_bank, want_e2e=False)
        except Exception as exc:  # noqa: BLE001
            node_part = {"error": str(exc).splitlines()[0][:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.qkv_storage == "fp32" else "f32 arithmetic, q/k/v rows stored bf16",
                "data": "synthetic",
                "config": {"workload": w.name, "nodes": n, "edges_per_snapshot": e, "snapshots": t_steps,
                           "hidden": hdim, "heads": h, "distance_metric": args.metric,
                           "parallelism": f"dp{world} (one sequence per GPU, NCCL grad all-reduce)",
                           "memory_bank": not args.no_bank,
                           "l2": ("inputs larger than L2: one step streams %.0f GB of activations through a 126 MB L2 "
                                  "(per-snapshot K/V of %.0f MB %s)" % (
                                      wl_bytes / 1e9, n * hdim * 4 / 1e6,
                                      "fit L2, so kernel (a)'s gathers are L2 hits" if n * hdim * 8 < 100e6
                                      else "do not fit L2"))},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "timing": {"mode": timing_mode, "eager_ms_per_step": eager_ms_step}, "node_partitioned": node_part}
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
