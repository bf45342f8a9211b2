"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

Two ways the path shards:
  * data-parallel over graph sequences: replicated weights, ONE flat all-reduce of the gradients
    per step (``GradBucket``) -- 0.5-2 M fp32 parameters, latency-bound on NVSwitch;
  * node partition of one large graph (``NodePartition``): contiguous id ranges, every temporal
    stage is per node (no traffic); the geometric layer needs the projected K|V rows of remote
    neighbours -> an all-gather forward and a reduce-scatter of dK|dV backward
    (``tagan_b200.partitioned``).
The helpers here are pure host logic and are also exercised on CPU with the gloo backend.
"""
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 bucket for the data-parallel gradient all-reduce (mean over ranks)."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append((off, p.numel()))
            off += p.numel()

    def all_reduce(self, world: int):
        """Pack grads (missing grad = zeros), all-reduce once, write the mean back into ``p.grad``."""
        for p, (off, n) in zip(self.params, self.offsets):
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
        if world > 1:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)
        for p, (off, n) in zip(self.params, self.offsets):
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)


class FlatGradBucket:
    """Data-parallel gradient bucket WITHOUT per-parameter copies: every ``p.grad`` is a view into one flat fp32 buffer
    (autograd accumulates into the views in place), so the all-reduce is ONE collective on ``flat`` and nothing is packed
    or unpacked.  Use ``bucket.zero()`` instead of ``zero_grad(set_to_none=True)`` (which would drop the views)."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = self.flat[off:off + k].view_as(p)
            off += (k + 3) // 4 * 4

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, world: int):
        """Mean over ranks, in place; the parameters' ``.grad`` views see the result immediately."""
        if world > 1:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)


@dataclass
class NodePartition:
    """Contiguous 1-D partition of node ids ``0..N-1`` over ``world`` ranks (equal blocks, the last
    ranks one shorter when N % world != 0), so the all-gather of K|V rows is a plain concatenation."""
    num_nodes: int
    world: int

    def bounds(self, rank: int) -> Tuple[int, int]:
        base, rem = divmod(self.num_nodes, self.world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def sizes(self) -> List[int]:
        return [self.bounds(r)[1] - self.bounds(r)[0] for r in range(self.world)]

    def owner(self, node: int) -> int:
        base, rem = divmod(self.num_nodes, self.world)
        cut = rem * (base + 1)
        return node // (base + 1) if node < cut else rem + (node - cut) // max(base, 1)

    def max_rows(self) -> int:
        return max(self.sizes())
