"""Node-partitioned geometric attention for one large graph.

Rank r owns the contiguous node range ``NodePartition.bounds(r)``: its query rows, its slice of the
features and of every per-node temporal stage (those need no communication).  The geometric layer needs
the projected K|V rows of remote neighbours:

    forward : K|V_local [n_loc,2H]  --all_gather-->  K|V_global [N,2H]  -> fused kernel over local rows
    backward: column pass over ALL source nodes gives partial dK|dV [N,2H]  --reduce_scatter(sum)-->  local

On a uniform random graph ~(world-1)/world of the neighbours are remote, so the halo is the whole K|V
matrix and a plain all-gather is the right collective (NVSwitch: every peer at full bandwidth).
``comm`` abstracts the two collectives so the same code runs under torch.distributed (NCCL) and under the
single-process emulation used by the 1-GPU tests.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .dist import NodePartition
from .ops import CALLS, CSR, _f32c, _ptr, _stream, workspace


def build_csr_part(edge_index: torch.Tensor, part: NodePartition, rank: int, transpose: bool = True) -> CSR:
    """CSR of this rank's rows (local ids) against global columns, built on device from the FULL edge list."""
    lib = _lib.load()
    ei = edge_index.long().contiguous()
    e = ei.shape[1]
    n = part.num_nodes
    lo, hi = part.bounds(rank)
    r = hi - lo
    dev = ei.device
    i32 = dict(dtype=torch.int32, device=dev)
    cap = e + r
    rowptr = torch.empty(r + 1, **i32)
    col = torch.empty(max(cap, 1), **i32)
    row = torch.empty(max(cap, 1), **i32)
    status = torch.empty(1, **i32)
    rowptr_t = row_t = perm_t = None
    if transpose:
        rowptr_t = torch.empty(n + 1, **i32)
        row_t = torch.empty(max(cap, 1), **i32)
        perm_t = torch.empty(max(cap, 1), **i32)
    ws = workspace(lib.tagan_csr_workspace_bytes(e, n), dev)
    rc = lib.tagan_csr_build_part(_ptr(ei) if e else None, e, n, lo, r, _ptr(rowptr), _ptr(col), _ptr(row),
                                  _ptr(rowptr_t), _ptr(row_t), _ptr(perm_t), _ptr(status), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "tagan_csr_build_part")
    CALLS["n"] += 22 if transpose else 12
    return CSR(r, e, rowptr, col, row, rowptr_t, row_t, perm_t, status)


class TorchDistComm:
    """all_gather / reduce_scatter of equal-size blocks over a torch.distributed group (NCCL).

    The ``*_async`` forms enqueue the collective on a side stream of this object and return ``(out, wait)``:
    ``wait()`` makes the *caller's current* stream wait for the collective.  The caller decides when that has
    to happen, which is what lets the exchange of snapshot t+1 overlap the kernels of snapshot t (forward) and
    the reduce-scatter of snapshot t overlap the backward of snapshot t-1.  The input is kept alive by the
    ``wait`` closure, so drop the closure after calling it."""

    def __init__(self, part: NodePartition, rank: int, group=None):
        self.part, self.rank, self.group = part, rank, group
        assert part.num_nodes % part.world == 0, "equal blocks required (pad the graph to a multiple of world)"
        self.stream = torch.cuda.Stream()

    def _side(self, fn, out, inp):
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)                       # inp (and the allocation of out) are ready
        with torch.cuda.stream(self.stream):
            fn(out, inp, group=self.group)                 # blocking form: the side stream waits for NCCL's stream
            done = self.stream.record_event()

        def wait(_keep=(inp, out)):
            torch.cuda.current_stream().wait_event(done)
        return out, wait

    def all_gather_rows_async(self, local: torch.Tensor):
        out = torch.empty(self.part.num_nodes, local.shape[1], dtype=local.dtype, device=local.device)
        return self._side(dist.all_gather_into_tensor, out, local.contiguous())

    def reduce_scatter_rows_async(self, full: torch.Tensor):
        n_loc = self.part.num_nodes // self.part.world
        out = torch.empty(n_loc, full.shape[1], dtype=full.dtype, device=full.device)
        return self._side(dist.reduce_scatter_tensor, out, full.contiguous())

    def all_gather_rows(self, local):
        out, wait = self.all_gather_rows_async(local)
        wait()
        return out

    def reduce_scatter_rows(self, full):
        out, wait = self.reduce_scatter_rows_async(full)
        wait()
        return out


def _async(comm, name, arg):
    """Use the comm's async collective if it has one, else the blocking one with a no-op wait."""
    fn = getattr(comm, name + "_async", None)
    if fn is not None:
        return fn(arg)
    return getattr(comm, name)(arg), (lambda: None)


class _Halo:
    """Per-snapshot rendezvous between the gather node and the attention node of the autograd graph."""
    __slots__ = ("gather_wait", "rs_out", "rs_wait", "dqkv")

    def __init__(self):
        self.gather_wait = self.rs_out = self.rs_wait = self.dqkv = None


class _HaloGatherFn(torch.autograd.Function):
    """forward: start the all-gather of this rank's K|V rows.  backward: wait for the reduce-scatter that the
    attention node started, and hand back the complete dqkv = [dQ | dK|dV]."""

    @staticmethod
    def forward(ctx, qkv_loc, comm, halo: _Halo):
        h = qkv_loc.shape[1] // 3
        kv, halo.gather_wait = _async(comm, "all_gather_rows", qkv_loc.detach()[:, h:])
        ctx.halo, ctx.h = halo, h
        return kv

    @staticmethod
    def backward(ctx, _dkv_placeholder):
        halo, h = ctx.halo, ctx.h
        halo.rs_wait()
        halo.dqkv[:, h:] = halo.rs_out
        out = halo.dqkv
        halo.dqkv = halo.rs_out = halo.rs_wait = None
        return out, None, None


class _PartGeoAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv_loc, kv, metric_param, csr: CSR, comm, halo: _Halo, heads: int, metric: int, n_src: int):
        lib = _lib.load()
        qkv2 = _f32c(qkv_loc).contiguous()
        n_loc, three_h = qkv2.shape
        h = three_h // 3
        halo.gather_wait()                                           # compute stream waits for the halo only now
        halo.gather_wait = None
        ctxv = torch.empty(n_loc, h, dtype=torch.float32, device=qkv2.device)
        lse = torch.empty(n_loc, heads, dtype=torch.float32, device=qkv2.device)
        k_ptr, v_ptr = C.c_void_p(kv.data_ptr()), C.c_void_p(kv.data_ptr() + h * 4)
        rc = lib.tagan_geo_attn_fwd_part(_ptr(qkv2), three_h, k_ptr, v_ptr, 2 * h, _ptr(csr.rowptr), _ptr(csr.col), n_loc,
                                         h, heads, metric, _ptr(metric_param), _ptr(ctxv), _ptr(lse), None, _stream())
        _lib.check(rc, "tagan_geo_attn_fwd_part")
        CALLS["n"] += 2
        ctx.save_for_backward(qkv2, kv, metric_param, ctxv, lse)
        ctx.csr, ctx.comm, ctx.halo, ctx.heads, ctx.metric, ctx.n_src = csr, comm, halo, heads, metric, n_src
        return ctxv

    @staticmethod
    def backward(ctx, dctx):
        lib = _lib.load()
        qkv2, kv, metric_param, ctxv, lse = ctx.saved_tensors
        csr, comm, halo, heads, metric, n_src = ctx.csr, ctx.comm, ctx.halo, ctx.heads, ctx.metric, ctx.n_src
        n_loc, three_h = qkv2.shape
        h = three_h // 3
        dev = qkv2.device
        dctx = _f32c(dctx).contiguous()
        dqkv = torch.empty(n_loc, three_h, dtype=torch.float32, device=dev)
        dkv_part = torch.empty(n_src, 2 * h, dtype=torch.float32, device=dev)
        delta = torch.empty(n_loc, heads, dtype=torch.float32, device=dev)
        want_dp = metric_param is not None and metric in (7, 8)
        dp_ws = torch.empty(n_loc, heads, dtype=torch.float32, device=dev) if want_dp else None
        dparam = torch.empty(heads, dtype=torch.float32, device=dev) if want_dp else None
        k_ptr, v_ptr = C.c_void_p(kv.data_ptr()), C.c_void_p(kv.data_ptr() + h * 4)
        dk_ptr, dv_ptr = C.c_void_p(dkv_part.data_ptr()), C.c_void_p(dkv_part.data_ptr() + h * 4)
        rc = lib.tagan_geo_attn_bwd_part(_ptr(qkv2), three_h, k_ptr, v_ptr, 2 * h, _ptr(csr.rowptr), _ptr(csr.col),
                                         _ptr(csr.rowptr_t), _ptr(csr.row_t), n_loc, n_src, h, heads, metric,
                                         _ptr(metric_param), _ptr(ctxv), _ptr(lse), _ptr(dctx), _ptr(dqkv), three_h,
                                         dk_ptr, dv_ptr, 2 * h, _ptr(delta), _ptr(dp_ws), _ptr(dparam), _stream())
        _lib.check(rc, "tagan_geo_attn_bwd_part")
        CALLS["n"] += 4
        # every rank's partial dK|dV -> owner; started now, awaited in _HaloGatherFn.backward
        halo.rs_out, halo.rs_wait = _async(comm, "reduce_scatter_rows", dkv_part)
        halo.dqkv = dqkv
        return None, dkv_part, dparam, None, None, None, None, None, None


def geo_attention_core_part(qkv_loc, csr: CSR, comm, heads: int, metric: str, n_src: int, metric_param=None):
    halo = _Halo()
    kv = _HaloGatherFn.apply(qkv_loc, comm, halo)
    return _PartGeoAttnFn.apply(qkv_loc, kv, metric_param, csr, comm, halo, heads, ops.METRIC_ID[metric], n_src)


def _project(layer, x_loc):
    ln = layer.use_layer_norm
    xn = ops.layer_norm(x_loc, layer.layer_norm1.weight, layer.layer_norm1.bias) if ln else x_loc
    w_qkv = torch.cat([layer.q_linear.weight, layer.k_linear.weight, layer.v_linear.weight], 0)
    b_qkv = torch.cat([layer.q_linear.bias, layer.k_linear.bias, layer.v_linear.bias], 0)
    return ops.linear(xn, w_qkv, b_qkv)


def _finish(layer, ctxv, x_loc):
    o = ops.linear(ctxv, layer.output_proj.weight, layer.output_proj.bias)
    if layer.use_layer_norm:
        return ops.layer_norm(o, layer.layer_norm2.weight, layer.layer_norm2.bias, res=x_loc)
    return ops.add(o, x_loc)


def geometric_layer_part(layer, x_loc: torch.Tensor, csr: CSR, comm, n_src: int) -> torch.Tensor:
    """``GeometricAttention.forward`` (reference geometric_attention.py:518-598) on this rank's node slice.
    ``layer`` is a ``tagan_b200.GeometricAttention`` with replicated weights."""
    qkv = _project(layer, x_loc)
    ctxv = geo_attention_core_part(qkv, csr, comm, layer.num_heads, layer.distance_metric, n_src,
                                   getattr(layer, "distance_param", None))
    return _finish(layer, ctxv, x_loc)


def geometric_stage_part(layer, xs_loc, csrs, comm, n_src: int):
    """The geometric layer over all T snapshots of a node-partitioned graph, software-pipelined: the projection
    and the halo all-gather of snapshot t+1 are issued before the attention kernel of snapshot t, so the NVLink
    transfer runs under the kernels (and, by the construction of the autograd graph, the reduce-scatter of
    snapshot t runs under the backward kernels of snapshot t-1)."""
    t_steps = len(xs_loc)
    mid = ops.METRIC_ID[layer.distance_metric]
    param = getattr(layer, "distance_param", None)
    halos = [_Halo() for _ in range(t_steps)]
    qkv = [None] * t_steps
    kv = [None] * t_steps
    outs = []
    qkv[0] = _project(layer, xs_loc[0])
    kv[0] = _HaloGatherFn.apply(qkv[0], comm, halos[0])
    for t in range(t_steps):
        if t + 1 < t_steps:
            qkv[t + 1] = _project(layer, xs_loc[t + 1])
            kv[t + 1] = _HaloGatherFn.apply(qkv[t + 1], comm, halos[t + 1])
        ctxv = _PartGeoAttnFn.apply(qkv[t], kv[t], param, csrs[t], comm, halos[t], layer.num_heads, mid, n_src)
        outs.append(_finish(layer, ctxv, xs_loc[t]))
        qkv[t] = kv[t] = None
    return outs


# ------------------------------------------------------------------------------------------
# Snapshot-parallel geometric stage (T % world == 0): no per-snapshot halo at all
# ------------------------------------------------------------------------------------------
class AllToAllComm:
    """The one exchange of the snapshot-parallel mode: equal blocks over a torch.distributed group (NCCL over
    NVLink; gloo on CPU for the host-logic tests).  ``world == 1`` is the identity."""

    def __init__(self, world: int, group=None):
        self.world, self.group = world, group

    def all_to_all(self, send: torch.Tensor) -> torch.Tensor:
        """send ``[world, ...]``: block r goes to rank r; returns ``[world, ...]`` with block s received from rank s."""
        if self.world == 1:
            return send
        send = send.contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv


class _NodeToSnapshotFn(torch.autograd.Function):
    """``[T, n_loc, H]`` (this rank's nodes, every snapshot) -> ``[T/world, N, H]`` (every node, this rank's snapshots).
    Snapshot t belongs to rank t // (T/world), so the send buffer is the input itself; backward is the inverse."""

    @staticmethod
    def forward(ctx, x, comm: AllToAllComm):
        t, n_loc, h = x.shape
        w = comm.world
        t_loc = t // w
        recv = comm.all_to_all(x.contiguous().view(w, t_loc, n_loc, h))          # [src rank, t_loc, n_loc, H]
        ctx.comm, ctx.shape = comm, (t, n_loc, h)
        return recv.permute(1, 0, 2, 3).reshape(t_loc, w * n_loc, h)              # node order = rank order (contiguous ranges)

    @staticmethod
    def backward(ctx, d):
        t, n_loc, h = ctx.shape
        w = ctx.comm.world
        send = d.reshape(t // w, w, n_loc, h).permute(1, 0, 2, 3).contiguous()     # [dst rank, t_loc, n_loc, H]
        return ctx.comm.all_to_all(send).view(t, n_loc, h), None


class _SnapshotToNodeFn(torch.autograd.Function):
    """Inverse of ``_NodeToSnapshotFn``: ``[T/world, N, H]`` -> ``[T, n_loc, H]``."""

    @staticmethod
    def forward(ctx, y, comm: AllToAllComm):
        t_loc, n, h = y.shape
        w = comm.world
        n_loc = n // w
        send = y.reshape(t_loc, w, n_loc, h).permute(1, 0, 2, 3).contiguous()
        ctx.comm, ctx.shape = comm, (t_loc, n, h)
        return comm.all_to_all(send).view(w * t_loc, n_loc, h)

    @staticmethod
    def backward(ctx, d):
        t_loc, n, h = ctx.shape
        w = ctx.comm.world
        recv = ctx.comm.all_to_all(d.contiguous().view(w, t_loc, n // w, h))
        return recv.permute(1, 0, 2, 3).reshape(t_loc, n, h), None


def node_to_snapshot(x, comm):
    return _NodeToSnapshotFn.apply(x, comm)


def snapshot_to_node(y, comm):
    return _SnapshotToNodeFn.apply(y, comm)


def snapshot_parallel_supported(t_steps: int, num_nodes: int, world: int) -> bool:
    return world >= 1 and t_steps % world == 0 and num_nodes % world == 0


def forward_snapshot_parallel(layer, x_loc: torch.Tensor, my_edge_indices, part: NodePartition, rank: int, comm: AllToAllComm,
                              time_stamps=None, bank=None) -> torch.Tensor:
    """The whole TAGAN layer on ONE large graph over ``world`` GPUs without a per-snapshot halo (config 4).

    The geometric stage has no dependency across snapshots, the temporal stages none across nodes.  So:

      x_loc [T, n_loc, H]  --all-to-all-->  [T/world, N, H]: this rank runs the UNPARTITIONED geometric layer (LN1, QKV,
      kernel (a) over the full CSR, output_proj, LN2) on its T/world snapshots  --all-to-all-->  [T, n_loc, H]
      -> GRU scan, skip connection, temporal attention, memory bank on this rank's nodes (purely local).

    One exchange of ``(world-1)/world * T * n_loc * H`` floats each way (1.75 GB per rank at config 4 on 8 GPUs) replaces
    the 2 * T all-gathers / reduce-scatters of the whole ``[N, 2H]`` K|V matrix of the halo path (57 GB per rank), and
    the geometric kernels are exactly the single-GPU ones, so the forward is bit-identical to the unpartitioned layer.
    ``my_edge_indices``: the GLOBAL edge lists of this rank's snapshots ``rank*T/world ... (rank+1)*T/world - 1``.
    Weights are replicated; every rank holds a PARTIAL weight gradient (its snapshots / its nodes): all-reduce(sum)."""
    t_steps, n_loc, _ = x_loc.shape
    if not snapshot_parallel_supported(t_steps, part.num_nodes, comm.world):
        raise ValueError("snapshot-parallel mode needs T % world == 0 and N % world == 0 (use forward_node_partitioned)")
    assert len(my_edge_indices) == t_steps // comm.world
    x_snap = node_to_snapshot(x_loc, comm)                                           # [T/world, N, H]
    geo_snap = layer.geometric.forward_seq(x_snap, my_edge_indices)                  # single-GPU geometric layer
    geo = snapshot_to_node(geo_snap, comm)                                           # [T, n_loc, H]
    prop = layer.propagation.forward_core(geo, time_stamps)
    if bank is not None:
        ids = torch.arange(n_loc, dtype=torch.int32, device=prop.device)             # bank slots are local row ids
        for t in range(t_steps):
            bank.get_states(ids)
            bank.update(ids, prop[t].detach(), t)
    return layer.temporal_attention(prop, time_stamps=time_stamps, time_major=True)
