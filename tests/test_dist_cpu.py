"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (gradient bucket for data
parallelism, node partition arithmetic for the halo exchange).  Kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tagan_b200.dist import GradBucket, NodePartition
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    extra = torch.nn.Parameter(torch.ones(4))            # never receives a gradient (like time_k_proj)
    params = list(lin.parameters()) + [extra]
    x = torch.full((2, 5), float(rank + 1))
    lin(x).sum().backward()
    local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    GradBucket(params).all_reduce(world)
    # expected mean of the two ranks' gradients
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, out in zip(local, gathered):
        dist.all_gather(out, g)
    ok = all(torch.allclose(p.grad, sum(out) / world) for p, out in zip(params, gathered))
    # all-gather of equal-size K|V blocks == concatenation in node order
    part = NodePartition(11, world)
    lo, hi = part.bounds(rank)
    rows = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 2)
    pad = torch.zeros(part.max_rows(), 2)
    pad[: hi - lo] = rows
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    full = torch.cat([o[:s] for o, s in zip(outs, part.sizes())])
    ok = ok and torch.equal(full[:, 0], torch.arange(11, dtype=torch.float32))
    # flat-view bucket: grads accumulate into one buffer, one all-reduce, no copies
    from tagan_b200.dist import FlatGradBucket
    lin2 = torch.nn.Linear(5, 3)
    with torch.no_grad():
        for p_, q_ in zip(lin2.parameters(), lin.parameters()):
            p_.copy_(q_)
    fb = FlatGradBucket(list(lin2.parameters()))
    fb.zero()
    lin2(x).sum().backward()
    assert all(p_.grad.untyped_storage().data_ptr() == fb.flat.untyped_storage().data_ptr() for p_ in lin2.parameters())
    fb.all_reduce(world)
    ok = ok and all(torch.allclose(p_.grad, q_.grad) for p_, q_ in zip(lin2.parameters(), lin.parameters()))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradbucket_and_partition_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_node_partition_arithmetic():
    from tagan_b200.dist import NodePartition
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (16, 4)]:
        part = NodePartition(n, w)
        b = [part.bounds(r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert sum(part.sizes()) == n and max(part.sizes()) - min(part.sizes()) <= 1
        for node in {0, n // 2, n - 1}:
            lo, hi = part.bounds(part.owner(node))
            assert lo <= node < hi


def _a2a_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tagan_b200 import partitioned
    t, n, h = 2 * world, 3 * world, 4
    n_loc, t_loc = n // world, t // world
    # global tensor G[t, node, c] = 1000 t + 10 node + c; every rank holds its node slice
    full = (1000.0 * torch.arange(t).view(t, 1, 1) + 10.0 * torch.arange(n).view(1, n, 1) + torch.arange(h).view(1, 1, h))
    x_loc = full[:, rank * n_loc:(rank + 1) * n_loc].clone().requires_grad_(True)
    comm = partitioned.AllToAllComm(world)
    snap = partitioned.node_to_snapshot(x_loc, comm)                   # [t_loc, n, h]: every node, my snapshots
    ok = torch.equal(snap.detach(), full[rank * t_loc:(rank + 1) * t_loc])
    back = partitioned.snapshot_to_node(snap * 2.0, comm)              # [t, n_loc, h]
    ok = ok and torch.equal(back.detach(), 2.0 * full[:, rank * n_loc:(rank + 1) * n_loc])
    wgt = torch.arange(t * n_loc * h, dtype=torch.float32).view(t, n_loc, h) + rank
    (back * wgt).sum().backward()                                      # gradient returns through both exchanges
    ok = ok and torch.equal(x_loc.grad, 2.0 * wgt)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_snapshot_parallel_all_to_all_world2_gloo():
    """node-partitioned <-> snapshot-partitioned exchange of tagan_b200.partitioned (forward values and the gradient
    path through both all-to-alls) on two gloo ranks."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
