"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's hot path (the oracle).

Every function cites the reference lines it follows (paths relative to /root/reference).
Plain torch on CPU (fp32 by default, fp64-capable) plus numpy for the integer parts.  All
functions are autograd-transparent, so gradients of the oracle are the gradient oracle.
Pinned against the live reference by ``oracle/make_golden.py`` -> ``tests/golden/*.pt`` and
``tests/test_oracle_golden.py``.

State dicts use the reference's own parameter names, so a reference
module's ``state_dict()`` can be passed straight in.
"""
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# Metric ids shared with include/tagan_b200.h (enum tagan_metric).
METRICS = [
    "scaled_dot_product", "dot_product", "cosine_similarity", "euclidean",
    "squared_euclidean", "manhattan", "cosine_distance", "gaussian_kernel", "rbf_kernel",
]
METRIC_ID = {m: i for i, m in enumerate(METRICS)}


# --------------------------------------------------------------------------------------
# (a1) edge_index -> de-duplicated adjacency with self loops, as CSR
# --------------------------------------------------------------------------------------
def build_csr(edge_index, num_nodes: int) -> Dict[str, np.ndarray]:
    """Sparse form of ``adj[ei[0], ei[1]] = 1; adj += eye`` (src/tagan/layers/graph_attention.py:98-102).

    The dense mask only tests ``== 0`` (geometric_attention.py:505), so duplicates and explicit
    self edges collapse: the entry set is ``unique(edges) U {(i,i)}``.  Row = ``edge_index[0]``
    (the aggregating/query node), col = ``edge_index[1]`` (the gathered key/value node).
    Negative indices wrap like torch advanced indexing does.  Entries are row-major sorted.

    Returns int32 ``rowptr[N+1], col[nnz]`` and the transposed CSR ``rowptr_t[N+1], row_t[nnz],
    perm_t[nnz]`` (entries sorted by (col,row); ``perm_t[k]`` = CSR position of transposed entry k).
    """
    ei = np.asarray(edge_index.cpu() if isinstance(edge_index, torch.Tensor) else edge_index).astype(np.int64)
    n = int(num_nodes)
    r = ei[0].copy()
    c = ei[1].copy()
    if r.size and (r.min() < -n or r.max() >= n or c.min() < -n or c.max() >= n):
        raise IndexError("edge_index out of range")
    r[r < 0] += n
    c[c < 0] += n
    diag = np.arange(n, dtype=np.int64)
    keys = np.unique(np.concatenate([r * n + c, diag * n + diag]))
    row = keys // max(n, 1)
    col = keys % max(n, 1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    rowptr = np.cumsum(rowptr)
    perm_t = np.argsort(col, kind="stable")
    rowptr_t = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr_t, col + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    return dict(rowptr=rowptr.astype(np.int32), col=col.astype(np.int32),
                row=row.astype(np.int32), rowptr_t=rowptr_t.astype(np.int32),
                row_t=row[perm_t].astype(np.int32), perm_t=perm_t.astype(np.int32))


# --------------------------------------------------------------------------------------
# (a2) DistanceMetric, per (edge, head)
# --------------------------------------------------------------------------------------
def _cos(q, k):
    # src/tagan/layers/geometric_attention.py:66-90
    qn = torch.norm(q, p=2, dim=-1, keepdim=True)
    kn = torch.norm(k, p=2, dim=-1, keepdim=True)
    qn = torch.where(qn == 0, torch.ones_like(qn) * 1e-8, qn)
    kn = torch.where(kn == 0, torch.ones_like(kn) * 1e-8, kn)
    sim = torch.sum(q * k, dim=-1) / (qn * kn).squeeze(-1)
    return torch.clamp(sim, -1.0, 1.0)


def edge_scores(q: torch.Tensor, k: torch.Tensor, metric: str,
                metric_param: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Attention score of ``q[e,h,:]`` against ``k[e,h,:]`` -> ``[E,h]``.

    Follows ``DistanceMetric`` (geometric_attention.py:24-193) and the sign conventions of
    ``GeometricAttention._get_attention_weights`` (:351-354 sdp; :364-376 similarities used
    as-is; :378-401 distances negated; :403-434 kernels used as-is, per-head sigma/gamma =
    ``distance_param[h]`` only when learnable, else the function default 1.0).
    """
    d = q.shape[-1]
    if metric == "scaled_dot_product":
        return torch.sum(q * k, dim=-1) / math.sqrt(d)
    if metric == "dot_product":
        return torch.sum(q * k, dim=-1)
    if metric == "cosine_similarity":
        return _cos(q, k)
    if metric == "euclidean":
        return -torch.sqrt(torch.sum((q - k) ** 2, dim=-1) + 1e-8)
    if metric == "squared_euclidean":
        return -torch.sum((q - k) ** 2, dim=-1)
    if metric == "manhattan":
        return -torch.sum(torch.abs(q - k), dim=-1)
    if metric == "cosine_distance":
        return -(1.0 - _cos(q, k))
    sq = torch.sum((q - k) ** 2, dim=-1)
    if metric == "gaussian_kernel":
        sigma = metric_param if metric_param is not None else torch.ones((), dtype=q.dtype)
        return torch.exp(-sq / (2 * sigma ** 2))
    if metric == "rbf_kernel":
        gamma = metric_param if metric_param is not None else torch.ones((), dtype=q.dtype)
        return torch.exp(-gamma * sq)
    raise ValueError(f"Unknown distance metric: {metric}")


def _ln(x, sd, prefix, enabled=True):
    if not enabled:
        return x
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def segment_softmax_aggregate(s, v_col, row, num_rows):
    """softmax of s[E,h] over entries sharing ``row`` and weighted sum of v_col[E,h,d]."""
    h = s.shape[1]
    idx = row[:, None].expand(-1, h)
    m = torch.full((num_rows, h), -float("inf"), dtype=s.dtype).scatter_reduce(0, idx, s, "amax", include_self=True)
    p = torch.exp(s - m[row])
    l = torch.zeros((num_rows, h), dtype=s.dtype).index_add_(0, row, p)
    a = p / l[row]
    ctx = torch.zeros((num_rows,) + tuple(v_col.shape[1:]), dtype=s.dtype).index_add_(0, row, a[..., None] * v_col)
    return a, ctx


# --------------------------------------------------------------------------------------
# (a1,a3,a4) TAGANGraphAttention.forward, sparse
# --------------------------------------------------------------------------------------
def geo_attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], edge_index, num_heads: int,
                  metric: str = "scaled_dot_product", use_layer_norm: bool = True,
                  learnable_distance: bool = False, csr: Optional[dict] = None,
                  return_attn: bool = False, qkv_round: Optional[torch.dtype] = None):
    """``TAGANGraphAttention.forward`` (graph_attention.py:61-133) ->
    ``GeometricAttention.forward`` (geometric_attention.py:518-598) on the entry set of
    :func:`build_csr`.  ``sd`` keys: ``{q,k,v}_linear.*, output_proj.*, layer_norm{1,2}.*``
    and optionally ``distance_param``.  dropout = 0 (eval).  Returns ``out[N,H]`` and, if
    asked, per-entry weights ``attn[nnz,h]`` == dense ``attn[0,h,row,col]`` (:511).
    """
    n, hdim = x.shape
    d = hdim // num_heads
    if csr is None:
        csr = build_csr(edge_index, n)
    row = torch.from_numpy(csr["row"].astype(np.int64))
    col = torch.from_numpy(csr["col"].astype(np.int64))
    xn = _ln(x, sd, "layer_norm1", use_layer_norm)                       # :541-542
    q = _lin(xn, sd, "q_linear").view(n, num_heads, d)                   # :546-560
    k = _lin(xn, sd, "k_linear").view(n, num_heads, d)
    v = _lin(xn, sd, "v_linear").view(n, num_heads, d)
    if qkv_round is not None:
        # restatement of the bf16-STORAGE mode of the product (not a reference feature): q, k, v are rounded to `qkv_round`
        # (round-to-nearest-even) where they are stored, everything else stays in the working precision; the rounding is a
        # straight-through op for the gradient, as in the product
        q, k, v = (t + (t.to(qkv_round).to(t.dtype) - t).detach() for t in (q, k, v))
    param = sd.get("distance_param") if (learnable_distance and metric in ("gaussian_kernel", "rbf_kernel")) else None
    s = edge_scores(q[row], k[col], metric, param)                       # :332-503
    a, ctx = segment_softmax_aggregate(s, v[col], row, n)                # :505-511, :579
    out = _lin(ctx.reshape(n, hdim), sd, "output_proj") + x              # :586-592 (residual on pre-LN x)
    out = _ln(out, sd, "layer_norm2", use_layer_norm)                    # :595-596
    return (out, a) if return_attn else out


def geo_layer_with_skip(x, sd_layers: Sequence[dict], skip_ln: Optional[dict], edge_index, num_heads,
                        metric, use_layer_norm=True, learnable_distance=False):
    """Per-snapshot stack of ``model.py:242-262``: L geometric layers, after layer 0 only
    ``x = x + skip_layer_norm(embedded_x)`` (row a5)."""
    skip = x
    csr = build_csr(edge_index, x.shape[0])
    for i, sd in enumerate(sd_layers):
        x = geo_attention(x, sd, edge_index, num_heads, metric, use_layer_norm, learnable_distance, csr=csr)
        if i == 0:
            x = x + (_ln(skip, skip_ln, "skip_layer_norm") if skip_ln is not None else skip)
    return x


# --------------------------------------------------------------------------------------
# (b1,b2) AsymmetricTemporalAttention.forward
# --------------------------------------------------------------------------------------
def time_bias(time_stamps: torch.Tensor, sd, num_heads: int) -> torch.Tensor:
    """RBF time bias ``[B,h,T,T]``: ``_compute_time_based_attention`` (temporal_attention.py:792-871)
    -> ``TimeEncoding._get_basis_encoding`` (:122-220; *global* min/max normalisation :142-152,
    sigma clamp :173-179, exponent clamp :189) -> ``time_q_proj`` (:848).  ``time_k_proj`` is unused."""
    b, t = time_stamps.shape
    diffs = time_stamps.unsqueeze(2) - time_stamps.unsqueeze(1)           # :1033  Delta[b,i,j] = t_i - t_j
    if diffs.numel() == 0:
        return torch.zeros(b, num_heads, t, t, dtype=time_stamps.dtype)
    tv = torch.nan_to_num(diffs, nan=0.0) if torch.isnan(diffs).any() else diffs
    tmin, tmax = tv.min(), tv.max()
    if bool(tmax > tmin) and bool((tmax - tmin) > 1e-7):
        tn = (tv - tmin) / (tmax - tmin)
    else:
        tn = torch.zeros_like(tv)
    mu = sd["time_encoding.basis_mu"]
    sigma = sd["time_encoding.basis_sigma"]
    if float(sigma.detach().min()) < 1e-7:
        sigma = torch.clamp(sigma, min=1e-7)
    expo = -((tn.unsqueeze(-1) - mu) ** 2 / (2 * sigma ** 2))
    expo = torch.clamp(expo, min=-88.0, max=88.0)
    enc = _lin(torch.exp(expo), sd, "time_encoding.basis_proj")           # [B,T,T,H]; scale 1.0, dropout off
    return _lin(enc, sd, "time_q_proj").permute(0, 3, 1, 2)


def asym_temporal_attention(x, sd, num_heads: int, time_stamps: Optional[torch.Tensor] = None,
                            attention_mask=None, causal: bool = False, time_aware: bool = True,
                            use_layer_norm: bool = True, asymmetric_window_size: int = 5,
                            relative_position_bias: bool = True, max_relative_position: int = 32,
                            use_time_masks: bool = True, return_attention_weights: bool = False):
    """``AsymmetricTemporalAttention.forward`` (temporal_attention.py:904-1205), dropout = 0.

    Keeps the reference's data-dependent mask rules: wrong-shaped mask => causal (:1118-1132);
    right-shaped mask gets ``unsqueeze(1)`` and, if ALL ones, is multiplied by ``tril`` (:1142-1148);
    a masked_fill that fails to broadcast leaves the scores unmasked (:1164-1170).
    """
    if isinstance(x, list):                                               # :928-976 pad + stack + permute
        cur = [t[0] if isinstance(t, list) and len(t) > 0 else t for t in x]
        mx = max(t.shape[0] for t in cur)
        cur = [torch.cat([t, torch.zeros(mx - t.shape[0], t.shape[1], dtype=t.dtype)], 0) if t.shape[0] < mx else t
               for t in cur]
        x = torch.stack(cur, 0).permute(1, 0, 2)
    b, t, hdim = x.shape
    d = hdim // num_heads
    identity = x
    xn = _ln(x, sd, "layer_norm1", use_layer_norm)
    q = _lin(xn, sd, "q_linear").view(b, t, num_heads, d).transpose(1, 2)
    k = _lin(xn, sd, "k_linear").view(b, t, num_heads, d).transpose(1, 2)
    v = _lin(xn, sd, "v_linear").view(b, t, num_heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)               # :1008
    pos = torch.arange(t)
    rel = pos.unsqueeze(1) - pos.unsqueeze(0)                             # i - j
    if relative_position_bias:                                            # :1011-1021
        idx = torch.clamp(rel + max_relative_position, 0, 2 * max_relative_position)
        s = s + sd["relative_pos_table"][idx].permute(2, 0, 1).unsqueeze(0)
    w = asymmetric_window_size                                            # :756-790, :1024-1027
    within = (rel >= -w) & (rel <= w)
    kv = sd["asymmetric_kernel"][torch.clamp(rel + w, 0, 2 * w)] * within.unsqueeze(-1).to(x.dtype)
    s = s + kv.permute(2, 0, 1).unsqueeze(0)
    if time_aware and time_stamps is not None:                            # :1030-1069
        s = s + time_bias(time_stamps, sd, num_heads)
        if use_time_masks:
            tm = (torch.abs(time_stamps.unsqueeze(2) - time_stamps.unsqueeze(1)) <= 10.0).to(x.dtype)  # :873-903
            if attention_mask is not None:
                if not isinstance(attention_mask, list):
                    if attention_mask.shape[-2:] == tm.shape[-2:]:
                        attention_mask = attention_mask * tm
            else:
                attention_mask = tm
    if causal:                                                            # :1073-1076
        s = s.masked_fill(torch.tril(torch.ones(t, t)).unsqueeze(0) == 0, float("-inf"))
    if attention_mask is not None:                                        # :1079-1172
        if isinstance(attention_mask, list):
            ln = len(attention_mask)
            attention_mask = torch.ones(b, ln, ln)
        if attention_mask.shape[-1] != t or attention_mask.shape[-2] != t:
            expanded = torch.tril(torch.ones(t, t)).unsqueeze(0).unsqueeze(0).expand(b, num_heads, t, t)
        else:
            expanded = attention_mask.unsqueeze(1)
            if bool(torch.all(expanded == 1.0)):
                expanded = expanded * torch.tril(torch.ones(t, t))
        if expanded.numel() != 0:
            try:
                s = s.masked_fill(expanded == 0, float("-inf"))
            except Exception:
                pass
    a = F.softmax(s, dim=-1)
    ctx = torch.matmul(a, v).transpose(1, 2).reshape(b, t, hdim)
    out = _lin(ctx, sd, "output_proj") + identity
    out = _ln(out, sd, "layer_norm2", use_layer_norm)
    return (out, a) if return_attention_weights else out


# --------------------------------------------------------------------------------------
# (b3,b4) TemporalGRUCell / TemporalEvolutionLayer (unidirectional)
# --------------------------------------------------------------------------------------
def gru_cell(x, h, time_diff, sd, prefix="", use_layer_norm=True):
    """``TemporalGRUCell.forward`` (temporal_propagation.py:475-551), dropout = 0."""
    x = _ln(x, sd, prefix + "layer_norm_x", use_layer_norm)               # :499-500
    if h is None:
        h = torch.zeros(x.shape[0], sd[prefix + "candidate.weight"].shape[0], dtype=x.dtype)   # :503-504
    else:
        h = _ln(h, sd, prefix + "layer_norm_h", use_layer_norm)           # :505-506
    if time_diff is not None:                                             # :509-514
        h = h * torch.exp(-torch.clamp(time_diff, min=0.0, max=10.0)).unsqueeze(1)
    xh = torch.cat([x, h], dim=-1)
    r = torch.sigmoid(_lin(xh, sd, prefix + "reset_gate"))                # :531-532
    z = torch.sigmoid(_lin(xh, sd, prefix + "update_gate"))
    ht = torch.tanh(_lin(torch.cat([x, r * h], dim=-1), sd, prefix + "candidate"))  # :535-536
    hn = (1 - z) * h + z * ht                                             # :539
    return _ln(hn, sd, prefix + "layer_norm_out", use_layer_norm)         # :545-546


def evolution_layer(xs: List[torch.Tensor], time_stamps, sd, time_aware=True, use_layer_norm=True,
                    residual=True):
    """``TemporalEvolutionLayer.forward`` (temporal_propagation.py:648-755), bidirectional=False."""
    h = None
    states = []
    for t in range(len(xs)):                                              # :675-688
        td = None
        if time_stamps is not None and time_aware and t > 0:
            td = time_stamps[:, t] - time_stamps[:, t - 1]
        h = gru_cell(xs[t], h, td, sd, "forward_cell.", use_layer_norm)
        states.append(h)
    outs = []
    for t in range(len(xs)):                                              # :735-753
        o = _lin(states[t], sd, "output_projection")
        if residual and xs[t].shape[-1] == o.shape[-1]:
            o = o + xs[t]
        outs.append(_ln(o, sd, "layer_norm", use_layer_norm))
    return outs


# --------------------------------------------------------------------------------------
# (b5) TemporalSkipConnection
# --------------------------------------------------------------------------------------
def skip_connection(xs: List[torch.Tensor], sd, window_size=3, aggregation="mean", use_layer_norm=True,
                    residual=True):
    """``TemporalSkipConnection.forward`` (temporal_propagation.py:846-946), dropout = 0.
    Two-sided window ``[max(0,t-w), min(T,t+w+1))`` (:880-894); 'max' (:896-910); else sum (:912-926)."""
    tt = len(xs)
    proj = [_ln(F.gelu(_lin(x, sd, "input_proj")), sd, "layer_norm1", use_layer_norm) for x in xs]   # :866-877
    outs = []
    for t in range(tt):
        win = torch.stack(proj[max(0, t - window_size):min(tt, t + window_size + 1)], 0)
        if aggregation == "mean":
            g = win.mean(0)
        elif aggregation == "max":
            g = win.max(0)[0]
        else:
            g = win.sum(0)
        o = _lin(F.gelu(g), sd, "output_proj")                            # :929-931
        if residual:
            o = o + xs[t]
        outs.append(_ln(o, sd, "layer_norm2", use_layer_norm))
    return outs


# --------------------------------------------------------------------------------------
# (b6) TemporalGatingUnit
# --------------------------------------------------------------------------------------
def gating_unit(cur, prev, sd, use_layer_norm=True, residual=True):
    """``TemporalGatingUnit.forward`` (temporal_propagation.py:1022-1067), dropout = 0."""
    c = _ln(cur, sd, "layer_norm_in1", use_layer_norm)
    p = _ln(prev, sd, "layer_norm_in2", use_layer_norm)
    comb = torch.cat([c, p], dim=1)
    u = torch.sigmoid(_lin(comb, sd, "update_gate"))
    r = torch.sigmoid(_lin(comb, sd, "reset_gate"))
    cand = torch.tanh(_lin(torch.cat([c, r * p], dim=1), sd, "output_gate"))
    o = (1 - u) * c + u * cand
    if residual:
        o = o + c
    return _ln(o, sd, "layer_norm_out", use_layer_norm)


# --------------------------------------------------------------------------------------
# (b7) runnable core of TemporalPropagation.forward
# --------------------------------------------------------------------------------------
def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def propagation_core(xs: List[torch.Tensor], time_stamps, sd, time_aware=True, use_layer_norm=True,
                     use_skip_connection=True, window_size=3, aggregation="mean", residual=True):
    """evolution -> skip -> ``LN(output_proj(.))`` (temporal_propagation.py:1343-1349, 1487-1500).
    The per-node memory-bank loop between them (:1357-1485) never completes in the reference
     and is not part of this function."""
    ev = evolution_layer(xs, time_stamps, _sub(sd, "evolution_layer."), time_aware, use_layer_norm, residual)
    if use_skip_connection:
        ev = skip_connection(ev, _sub(sd, "skip_connection."), window_size, aggregation, use_layer_norm, residual)
    return [_ln(_lin(e, sd, "output_proj"), sd, "layer_norm", use_layer_norm) for e in ev]


# --------------------------------------------------------------------------------------
# (c1-c3) NodeMemoryBank -- dense-table restatement, ids in [0, capacity)
# --------------------------------------------------------------------------------------
class BankOracle:
    """``NodeMemoryBank`` (src/tagan/utils/memory_bank.py:14-360) on dense numpy tables.

    ``valid``       <-> ``node_id in node_states``
    ``has_seen``    <-> ``node_id in last_seen``     (``get_states`` inserts without it, :205-209)
    Scalars ``w``, ``1-w`` and ``decay**k`` are computed in double and rounded to fp32 once,
    exactly what ``python_float * float32_tensor`` does in torch, so states are bit-exact.
    """

    def __init__(self, hidden_dim, capacity, decay_factor=0.8, max_inactivity=5):
        self.h, self.cap = hidden_dim, capacity
        self.decay_factor, self.max_inactivity = decay_factor, max_inactivity
        self.states = np.zeros((capacity, hidden_dim), np.float32)
        self.valid = np.zeros(capacity, np.uint8)
        self.has_seen = np.zeros(capacity, np.uint8)
        self.last_seen = np.zeros(capacity, np.int32)
        self.inactivity = np.zeros(capacity, np.int32)
        self.frequency = np.zeros(capacity, np.int32)

        self.size = 0      # like the reference, refreshed only at the end of update() (:169)

    def update(self, node_ids, states, timestep=0):
        """memory_bank.py:65-173."""
        ids = np.asarray(node_ids, np.int64)
        st = np.asarray(states, np.float32)
        m = min(len(ids), st.shape[0])                                   # bounds check :95
        ids, st = ids[:m], st[:m]
        self.inactivity[self.valid.astype(bool)] += 1                    # :89-90
        for i in range(m):                                               # :93-141 (order matters for duplicates)
            n = ids[i]
            self.frequency[n] += 1
            reappearing = bool(self.valid[n] and self.has_seen[n] and self.last_seen[n] < timestep - 1)
            cur = st[i].copy()
            if np.isnan(cur).any() and self.valid[n]:                    # :109-113 (rand recovery :116 not restated)
                cur = self.states[n].copy()
            if reappearing:                                              # :121-129
                w = max(0.4, self.decay_factor ** min(timestep - int(self.last_seen[n]), 3))
                self.states[n] = np.float32(w) * self.states[n] + np.float32(1 - w) * cur
            else:
                self.states[n] = cur                                     # :135
            self.valid[n] = 1
            self.inactivity[n] = 0                                       # :138
            self.last_seen[n] = timestep                                 # :141
            self.has_seen[n] = 1
        absent = self.valid.astype(bool)                                 # :149-153
        absent[ids] = False
        for n in np.nonzero(absent)[0]:
            self.states[n] = self.states[n] * np.float32(self.decay_factor ** int(self.inactivity[n]))
        prune = self.valid.astype(bool) & (self.inactivity > self.max_inactivity)   # :156-166
        self.valid[prune] = 0
        self.has_seen[prune] = 0
        self.inactivity[prune] = 0
        self.last_seen[prune] = 0
        self.states[prune] = 0
        self.size = int(self.valid.sum())                                # :169

    def get_states(self, node_ids):
        """memory_bank.py:187-211 (unknown ids -> zeros AND inserted, inactivity 0, no last_seen)."""
        ids = np.asarray(node_ids, np.int64)
        out = np.zeros((len(ids), self.h), np.float32)
        for i, n in enumerate(ids):
            if self.valid[n]:
                out[i] = self.states[n]
            else:
                self.states[n] = 0
                self.valid[n] = 1
                self.inactivity[n] = 0
        return out

    def get_state(self, node_id):
        return self.states[node_id].copy() if self.valid[node_id] else None

    def update_state(self, node_id, state, timestep=0):
        self.update([node_id], np.asarray(state, np.float32)[None], timestep)   # :235-244

    def decay_all(self):
        v = self.valid.astype(bool)
        self.states[v] = self.states[v] * np.float32(self.decay_factor)         # :222-225


# --------------------------------------------------------------------------------------
# "TAGAN layer" used by bench.py's CPU baseline
# --------------------------------------------------------------------------------------
def tagan_layer(xs: List[torch.Tensor], edge_indices, time_stamps_bt, sd_geo, sd_prop, sd_tattn,
                num_heads, metric, csrs=None):
    """One geometric layer per snapshot + propagation core + temporal attention (fwd).
    ``time_stamps_bt`` is ``[N,T]``."""
    geo = [geo_attention(x, sd_geo, ei, num_heads, metric, csr=(csrs[i] if csrs else None))
           for i, (x, ei) in enumerate(zip(xs, edge_indices))]
    prop = propagation_core(geo, time_stamps_bt, sd_prop)
    return asym_temporal_attention(torch.stack(prop, 1), sd_tattn, num_heads, time_stamps=time_stamps_bt)
