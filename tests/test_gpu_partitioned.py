"""GPU (one device): the node-partitioned geometric layer, with the two collectives emulated in-process,
must reproduce the unpartitioned layer bit-for-bit in the forward pass and to fp32 rounding in the
backward pass (the reduce-scatter only changes the summation order of dK|dV partials).  The partitioned CSR
is checked bit-exactly against slices of the full CSR."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class EmuComm:
    """Single-process stand-in for all_gather / reduce_scatter over `world` emulated ranks."""

    def __init__(self, part, rank, kv_blocks, dkv_store):
        self.part, self.rank, self.kv_blocks, self.dkv_store = part, rank, kv_blocks, dkv_store

    def all_gather_rows(self, local):
        return torch.cat(self.kv_blocks, 0)

    def reduce_scatter_rows(self, full):
        self.dkv_store[self.rank] = full.clone()
        lo, hi = self.part.bounds(self.rank)
        return torch.zeros(hi - lo, full.shape[1], device=full.device)      # patched after all ranks ran


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_layer_matches_full(world):
    import tagan_b200
    from tagan_b200 import ops, partitioned
    from tagan_b200.dist import NodePartition
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    n, e, hdim, heads = 1024, 12000, 128, 8
    layer = tagan_b200.GeometricAttention(hdim, heads, dropout=0.0, distance_metric="euclidean").to(dev)
    x = torch.randn(n, hdim, device=dev)
    ei = torch.randint(0, n, (2, e), device=dev)
    wout = torch.randn(n, hdim, device=dev)
    part = NodePartition(n, world)

    # reference: unpartitioned
    xf = x.clone().requires_grad_(True)
    csr = ops.build_csr(ei, n)
    out_full = layer.forward_csr(xf, csr)
    (out_full * wout).sum().backward()
    gfull = {k: p.grad.clone() for k, p in layer.named_parameters()}
    layer.zero_grad()

    # partitioned CSR == row slices of the full CSR (bit-exact)
    nnz = csr.nnz
    for r in range(world):
        lo, hi = part.bounds(r)
        c = partitioned.build_csr_part(ei, part, r)
        b, en = int(csr.rowptr[lo]), int(csr.rowptr[hi])
        assert torch.equal(c.rowptr, csr.rowptr[lo:hi + 1] - csr.rowptr[lo])
        assert torch.equal(c.col[:en - b], csr.col[b:en])
        assert torch.equal(c.row[:en - b], csr.row[b:en] - lo)
        assert int(c.rowptr_t[-1]) == en - b

    # emulated ranks: K|V blocks come from each rank's own projection
    import torch.nn.functional as F
    kv_blocks = []
    for r in range(world):
        lo, hi = part.bounds(r)
        with torch.no_grad():
            xn = ops.layer_norm(x[lo:hi], layer.layer_norm1.weight, layer.layer_norm1.bias)
            w = torch.cat([layer.k_linear.weight, layer.v_linear.weight], 0)
            bb = torch.cat([layer.k_linear.bias, layer.v_linear.bias], 0)
            kv_blocks.append(ops.linear(xn, w, bb))
    dkv_store = {}
    outs, xs = [], []
    for r in range(world):
        lo, hi = part.bounds(r)
        xl = x[lo:hi].clone().requires_grad_(True)
        c = partitioned.build_csr_part(ei, part, r)
        comm = EmuComm(part, r, kv_blocks, dkv_store)
        o = partitioned.geometric_layer_part(layer, xl, c, comm, n)
        outs.append(o)
        xs.append(xl)
    out_part = torch.cat(outs, 0)
    assert torch.equal(out_part.detach(), out_full.detach())                       # forward bit-identical
    # backward: each rank's own dK|dV contribution is zeroed by the emulation; add the reduced partials by hand
    (out_part * wout).sum().backward()
    dkv_total = sum(dkv_store[r] for r in range(world))                             # what reduce_scatter would deliver
    assert dkv_total.shape == (n, 2 * hdim)
    # full-graph dK|dV from the unpartitioned call for comparison
    qkv = None
    with torch.no_grad():
        xn = ops.layer_norm(x, layer.layer_norm1.weight, layer.layer_norm1.bias)
        w_qkv = torch.cat([layer.q_linear.weight, layer.k_linear.weight, layer.v_linear.weight], 0)
        b_qkv = torch.cat([layer.q_linear.bias, layer.k_linear.bias, layer.v_linear.bias], 0)
        qkv = ops.linear(xn, w_qkv, b_qkv)
    qkv = qkv.requires_grad_(True)
    ctx_full, _ = ops.geo_attention_core(qkv, csr, heads, "euclidean")
    o_full = ops.linear(ctx_full, layer.output_proj.weight, layer.output_proj.bias)
    o_full = ops.layer_norm(o_full, layer.layer_norm2.weight, layer.layer_norm2.bias, res=x)
    (o_full * wout).sum().backward()
    torch.testing.assert_close(dkv_total, qkv.grad[:, hdim:], rtol=1e-5, atol=1e-5)


class OneRankComm:
    """world = 1: both collectives are the identity; exercises the async hooks of the pipelined stage."""

    def __init__(self):
        self.log = []

    def all_gather_rows_async(self, local):
        self.log.append("ag")
        out = local.clone()
        return out, (lambda: self.log.append("ag_wait"))

    def reduce_scatter_rows_async(self, full):
        self.log.append("rs")
        out = full.clone()
        return out, (lambda: self.log.append("rs_wait"))


def test_pipelined_stage_matches_per_snapshot_layer():
    """geometric_stage_part (projection/gather of t+1 issued before attention of t) == the plain layer per
    snapshot, bit-exact forward, and the same input / weight gradients (fp32 sums in a different order only
    for the weights: tolerance 1e-5 relative)."""
    import tagan_b200
    from tagan_b200 import ops, partitioned
    from tagan_b200.dist import NodePartition
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    n, e, hdim, heads, t_steps = 512, 5000, 64, 4, 4
    layer = tagan_b200.GeometricAttention(hdim, heads, dropout=0.0, distance_metric="scaled_dot_product").to(dev)
    xs = [torch.randn(n, hdim, device=dev) for _ in range(t_steps)]
    eis = [torch.randint(0, n, (2, e), device=dev) for _ in range(t_steps)]
    wout = [torch.randn(n, hdim, device=dev) for _ in range(t_steps)]
    part = NodePartition(n, 1)

    xa = [x.clone().requires_grad_(True) for x in xs]
    ref = [layer.forward_csr(x, ops.build_csr(ei, n)) for x, ei in zip(xa, eis)]
    sum((o * w).sum() for o, w in zip(ref, wout)).backward()
    gref = {k: p.grad.clone() for k, p in layer.named_parameters()}
    layer.zero_grad()

    comm = OneRankComm()
    xb = [x.clone().requires_grad_(True) for x in xs]
    csrs = [partitioned.build_csr_part(ei, part, 0) for ei in eis]
    outs = partitioned.geometric_stage_part(layer, xb, csrs, comm, n)
    # the gather of snapshot t+1 is started before the attention of snapshot t waits for its own
    assert comm.log[:4] == ["ag", "ag", "ag_wait", "ag"], comm.log[:6]
    sum((o * w).sum() for o, w in zip(outs, wout)).backward()
    assert comm.log.count("rs") == t_steps and comm.log.count("rs_wait") == t_steps
    for o, r in zip(outs, ref):
        assert torch.equal(o, r)
    for a, b in zip(xb, xa):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
    for k, p in layer.named_parameters():
        torch.testing.assert_close(p.grad, gref[k], rtol=1e-5, atol=1e-5 * max(1.0, float(gref[k].abs().max())))


def test_forward_node_partitioned_world1_matches_layer():
    """forward_node_partitioned (whole TAGAN layer on a node-partitioned graph) with a single rank == TAGANLayer.forward:
    same CSR rows, same kernels, identity collectives -> bit-identical forward, gradients to summation order."""
    import tagan_b200
    from tagan_b200.dist import NodePartition
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    n, e, hdim, heads, t_steps = 640, 5000, 64, 4, 5
    layer = tagan_b200.TAGANLayer(hdim, heads, "euclidean").to(dev)
    xs = [torch.randn(n, hdim, device=dev) for _ in range(t_steps)]
    eis = [torch.randint(0, n, (2, e), device=dev) for _ in range(t_steps)]
    ts = torch.arange(t_steps, dtype=torch.float32, device=dev).expand(n, t_steps)
    wout = torch.randn(n, t_steps, hdim, device=dev)
    ref = layer(xs, eis, ts)
    (ref * wout).sum().backward()
    gref = {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}
    layer.zero_grad(set_to_none=True)
    out = tagan_b200.forward_node_partitioned(layer, xs, eis, NodePartition(n, 1), 0, OneRankComm(), ts)
    (out * wout).sum().backward()
    assert torch.equal(out, ref)
    for k, p in layer.named_parameters():
        if k in gref:
            torch.testing.assert_close(p.grad, gref[k], rtol=1e-4, atol=1e-4 * max(1.0, float(gref[k].abs().max())),
                                       msg=lambda m, k=k: f"{k}: {m}")


def test_snapshot_parallel_world1_matches_layer():
    """forward_snapshot_parallel with one rank (identity exchanges) == TAGANLayer.forward bit for bit, same gradients."""
    import tagan_b200
    from tagan_b200 import partitioned
    from tagan_b200.dist import NodePartition
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    n, e, hdim, heads, t_steps = 384, 4000, 64, 4, 4
    layer = tagan_b200.TAGANLayer(hdim, heads, "euclidean").to(dev)
    xs = torch.randn(t_steps, n, hdim, device=dev)
    eis = [torch.randint(0, n, (2, e), device=dev) for _ in range(t_steps)]
    ts = torch.arange(t_steps, dtype=torch.float32, device=dev).expand(n, t_steps)
    wout = torch.randn(n, t_steps, hdim, device=dev)
    ref = layer(list(xs.unbind(0)), eis, ts)
    (ref * wout).sum().backward()
    gref = {k: p.grad.clone() for k, p in layer.named_parameters() if p.grad is not None}
    layer.zero_grad(set_to_none=True)
    x_in = xs.clone().requires_grad_(True)
    out = partitioned.forward_snapshot_parallel(layer, x_in, eis, NodePartition(n, 1), 0, partitioned.AllToAllComm(1), ts)
    (out * wout).sum().backward()
    assert torch.equal(out, ref)
    assert x_in.grad is not None and bool(torch.isfinite(x_in.grad).all())
    for k, p in layer.named_parameters():
        if k in gref:
            assert torch.equal(p.grad, gref[k]), k
