"""Synthetic temporal graphs for the BASELINE.json configs.

All generators are seeded and produce CPU tensors (pinned on request); node ids are 0..N-1.
"""
from dataclasses import dataclass
from typing import List

import torch


@dataclass
class Workload:
    name: str
    num_nodes: int
    num_edges: int      # raw edges per snapshot
    snapshots: int
    hidden: int
    heads: int
    graph: str          # "uniform" | "powerlaw"


WORKLOADS = {
    # config 1: example.py toy (N<=10, E=2N, T=5, H=64, h=4)
    "c1": Workload("c1-example-toy", 10, 20, 5, 64, 4, "uniform"),
    # config 2: social graph 10k nodes, ~200k edges, 32 snapshots, 4 heads, hidden 128
    "c2": Workload("c2-social-10k-200k-T32-H128-h4", 10_000, 200_000, 32, 128, 4, "powerlaw"),
    # config 3: 100k nodes, 2M edges, 16 snapshots, 8 heads, hidden 128 (one sequence per GPU, data-parallel)
    "c3": Workload("c3-100k-2M-T16-H128-h8", 100_000, 2_000_000, 16, 128, 8, "uniform"),
    # config 4: 1M nodes, 10M edges, 16 snapshots, hidden 256 (node-partitioned across GPUs)
    "c4": Workload("c4-1M-10M-T16-H256-h8", 1_000_000, 10_000_000, 16, 256, 8, "uniform"),
    # config 5: 250k nodes, 5M edges, 128 snapshots, hidden 128
    "c5": Workload("c5-250k-5M-T128-H128-h8", 250_000, 5_000_000, 128, 128, 8, "uniform"),
}


def random_edges(num_nodes: int, num_edges: int, gen: torch.Generator, graph: str = "uniform") -> torch.Tensor:
    """[2,E] int64.  "powerlaw": Barabasi-Albert-like degree skew -- the expected degree of the node of
    rank i is proportional to i^(-1/2) (inverse-CDF sampling hub = N*u^2), i.e. max degree ~ m*sqrt(N) --
    made bidirectional as src/tagan/utils/data_utils.py:69-75 does."""
    if graph == "uniform":
        return torch.randint(0, num_nodes, (2, num_edges), generator=gen, dtype=torch.int64)
    half = num_edges // 2
    u = torch.rand(half, generator=gen)
    hub = (u.pow(2.0) * num_nodes).long().clamp_(0, num_nodes - 1)       # skewed towards low ids
    other = torch.randint(0, num_nodes, (half,), generator=gen, dtype=torch.int64)
    src = torch.cat([hub, other])
    dst = torch.cat([other, hub])
    return torch.stack([src, dst])


def make_sequence(w: Workload, seed: int = 0, snapshots: int = None, resample: float = 0.1, pin: bool = False):
    """Returns (xs: list of [N,H] fp32, edge_indices: list of [2,E] int64, time_stamps [N,T] fp32).
    10 % of the edges are re-sampled per snapshot."""
    gen = torch.Generator().manual_seed(seed)
    t = snapshots or w.snapshots
    ei = random_edges(w.num_nodes, w.num_edges, gen, w.graph)
    xs: List[torch.Tensor] = []
    eis: List[torch.Tensor] = []
    for _ in range(t):
        k = int(w.num_edges * resample)
        if k > 0:
            idx = torch.randint(0, w.num_edges, (k,), generator=gen)
            ei = ei.clone()
            ei[:, idx] = random_edges(w.num_nodes, k, gen, w.graph)
        x = torch.randn(w.num_nodes, w.hidden, generator=gen)
        if pin:
            x, e = x.pin_memory(), ei.pin_memory()
        else:
            e = ei
        xs.append(x)
        eis.append(e)
    ts = torch.arange(t, dtype=torch.float32).repeat(w.num_nodes, 1)
    return xs, eis, ts
