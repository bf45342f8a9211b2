"""The "TAGAN layer" the benchmark times and the hook that swaps the B200
layers into a reference ``TAGAN`` model.

``TAGANLayer`` = one geometric attention layer applied to every snapshot (CSR built on device per
snapshot) + the runnable propagation core (evolution -> skip -> projection) + temporal attention
over the snapshot axis + memory-bank gather/update per snapshot.  Everything runs in
libtagan_b200; torch only carries tensors between the calls and drives autograd.
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .layers import AsymmetricTemporalAttention, LayerNorm, TAGANGraphAttention, TemporalPropagation
from .memory_bank import NodeMemoryBank


class TAGANLayer(nn.Module):
    def __init__(self, hidden_dim: int, num_heads: int, distance_metric: str = "euclidean", dropout: float = 0.0,
                 temporal_window_size: int = 3, window_size: int = 5, causal_attention: bool = False):
        super().__init__()
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.geometric = TAGANGraphAttention(hidden_dim, num_heads, dropout, distance_metric)
        self.geometric.validate_indices = False
        self.propagation = TemporalPropagation(hidden_dim, hidden_dim, dropout, window_size=temporal_window_size)
        self.temporal_attention = AsymmetricTemporalAttention(hidden_dim, num_heads, dropout, causal=causal_attention,
                                                              asymmetric_window_size=window_size)
        # False: the reference's observable behaviour (bank gathered/updated, gating unit unused); True: the intended
        # per-snapshot gating pass (TemporalPropagation.forward_with_memory)
        self.gated_memory = False
        # batch LN / QKV / out-projection of the geometric layer over all snapshots when they share the node set
        self.batched_geometric = True

    def forward(self, xs: Sequence[torch.Tensor], edge_indices: Sequence, time_stamps: Optional[torch.Tensor] = None,
                bank: Optional[NodeMemoryBank] = None, node_ids: Optional[Sequence[torch.Tensor]] = None):
        """xs: T tensors ``[N,H]``; edge_indices: T ``[2,E]`` int64 tensors (or prebuilt ``ops.CSR``);
        time_stamps ``[N,T]``.  Returns ``[N,T,H]``."""
        xs_l = list(xs)
        uniform = all(isinstance(x, torch.Tensor) and x.shape == xs_l[0].shape for x in xs_l)
        if uniform and self.batched_geometric:
            # same node set in every snapshot: LN / projections of the geometric layer batched over T (one GEMM each)
            geo = self.geometric.forward_seq(xs_l, edge_indices)                      # [T,N,H]
        else:
            geo = [self.geometric(x, ei) for x, ei in zip(xs_l, edge_indices)]
        if bank is not None and self.gated_memory:
            n = geo[0].shape[0]
            ids_seq = node_ids if node_ids is not None else [torch.arange(n, dtype=torch.int32, device=geo[0].device)] * len(geo)
            prop = self.propagation.forward_with_memory(geo, ids_seq, bank, time_stamps)
            return self.temporal_attention(prop, time_stamps=time_stamps, time_major=True)
        prop = self.propagation.forward_core(geo, time_stamps)                      # [T,N,H]
        if bank is not None:
            t_steps, n = prop.shape[0], prop.shape[1]
            for t in range(t_steps):
                ids = node_ids[t] if node_ids is not None else torch.arange(n, dtype=torch.int32, device=prop.device)
                bank.get_states(ids)                                                 # previous states (gather)
                bank.update(ids, prop[t].detach(), t)                                # scatter-update / decay / prune
        return self.temporal_attention(prop, time_stamps=time_stamps, time_major=True)


def forward_node_partitioned(layer: "TAGANLayer", xs_loc: Sequence[torch.Tensor], edge_indices: Sequence[torch.Tensor],
                             part, rank: int, comm, time_stamps: Optional[torch.Tensor] = None,
                             bank: Optional[NodeMemoryBank] = None) -> torch.Tensor:
    """``TAGANLayer.forward`` for ONE large graph whose nodes are partitioned over the ranks (config 4).

    ``xs_loc``: this rank's rows ``[n_loc,H]`` of every snapshot; ``edge_indices``: the GLOBAL edge lists (every
    rank builds only its own rows of the CSR).  The geometric layer exchanges K|V halos (software-pipelined
    all-gather / reduce-scatter, ``partitioned.geometric_stage_part``); everything after it -- GRU scan, skip
    connection, temporal attention, memory bank -- is per node and therefore purely local.  Weights are replicated:
    all-reduce their gradients after backward (``dist.GradBucket``).  Returns this rank's ``[n_loc,T,H]``."""
    from . import partitioned
    ga = layer.geometric.geometric_attention
    csrs = [partitioned.build_csr_part(ei, part, rank) for ei in edge_indices]
    geo = torch.stack(partitioned.geometric_stage_part(ga, list(xs_loc), csrs, comm, part.num_nodes), 0)
    prop = layer.propagation.forward_core(geo, time_stamps)                          # [T,n_loc,H], local
    if bank is not None:
        n_loc = prop.shape[1]
        ids = torch.arange(n_loc, dtype=torch.int32, device=prop.device)             # bank slots are local row ids
        for t in range(prop.shape[0]):
            bank.get_states(ids)
            bank.update(ids, prop[t].detach(), t)
    return layer.temporal_attention(prop, time_stamps=time_stamps, time_major=True)


def patch(model: nn.Module) -> nn.Module:
    """Swap the hot-path layers of a reference ``TAGAN`` for the B200 ones, in place.

    Keeps every parameter (state_dict keys are identical) and the reference's own ``forward``:
    ``model.geometric_attention_layers[i]``, ``model.temporal_propagation``, ``model.temporal_attention``
    and ``model.memory_bank`` (model.py:57-61, 73-113 of the reference) are replaced.  Every swapped module takes the
    train()/eval() mode of the module it replaces, so ``model.eval(); patch(model)`` stays deterministic.
    """
    cfg = model.config
    dev = next(model.parameters()).device
    metric = "scaled_dot_product" if cfg.learnable_distance else "euclidean"       # reference model.py:80
    new_layers = nn.ModuleList()
    for old in model.geometric_attention_layers:
        new = TAGANGraphAttention(cfg.hidden_dim, cfg.num_heads, cfg.dropout, metric, cfg.use_layer_norm,
                                  cfg.learnable_distance)
        new.load_state_dict(old.state_dict())
        new_layers.append(new.to(dev).train(old.training))
    model.geometric_attention_layers = new_layers
    tp = TemporalPropagation(cfg.hidden_dim, cfg.hidden_dim, cfg.dropout, cfg.time_aware, cfg.bidirectional,
                             cfg.use_layer_norm, cfg.use_skip_connection, cfg.use_gating, cfg.temporal_window_size,
                             cfg.aggregation_method, cfg.use_residual)
    tp.load_state_dict(model.temporal_propagation.state_dict())
    model.temporal_propagation = tp.to(dev).train(model.temporal_propagation.training)
    ta = AsymmetricTemporalAttention(cfg.hidden_dim, cfg.num_heads, cfg.dropout, causal=cfg.causal_attention,
                                     time_aware=True, use_layer_norm=cfg.use_layer_norm,
                                     asymmetric_window_size=cfg.window_size,
                                     relative_position_bias=cfg.asymmetric_temporal_bias)
    ta.load_state_dict(model.temporal_attention.state_dict())
    model.temporal_attention = ta.to(dev).train(model.temporal_attention.training)
    if getattr(model, "skip_layer_norm", None) is not None:                          # row a5 (model.py:258-262)
        ln = LayerNorm(cfg.hidden_dim)
        ln.load_state_dict(model.skip_layer_norm.state_dict())
        model.skip_layer_norm = ln.to(dev).train(model.skip_layer_norm.training)
    if dev.type == "cuda":      # (on CPU only the module swap is done; the kernels need a CUDA device to run)
        old_bank = getattr(model, "memory_bank", None)
        if old_bank is not None and len(getattr(old_bank, "node_states", {})) > 0:
            import warnings
            warnings.warn("tagan_b200.patch: the model's memory bank holds %d node states; they are NOT migrated to the "
                          "device bank (the reference model never reads them: SURVEY.md fact 5)" % len(old_bank.node_states))
        model.memory_bank = NodeMemoryBank(cfg.hidden_dim, decay_factor=0.8, max_inactivity=cfg.temporal_window_size,
                                           device=dev)
    return model


# ------------------------------------------------------------------------------------------
# TAGAN.forward as a packed device pipeline
# ------------------------------------------------------------------------------------------
class TAGANModel(nn.Module):
    """Stand-alone mirror of the reference ``TAGAN`` (src/tagan/model.py:22-473) with the SAME parameter names -- a
    reference ``state_dict`` loads unchanged -- whose ``forward`` reproduces the observable behaviour of the shipped
    ``TAGAN.forward`` as one packed device pipeline:

    * all T snapshots' nodes are ONE packed matrix ``[sum N_t, .]``: node embedding, every LayerNorm / projection of the
      geometric layers and the layer-0 skip (model.py:233-262) run once over all rows; kernel (a) runs per snapshot on row
      ranges (snapshots may have different node counts);
    * ``temporal_propagation`` is NOT called: the reference's call raises and falls back to the geometric outputs
      (model.py:291-309); ``edge_embedding`` is ignored as in the reference (fact 3);
    * temporal attention sees the zero-padded ``[T, maxN, H]`` stack (one pad/stack launch) with the reference's all-ones
      mask (model.py:336-361: causal only when ``T == num_heads``, otherwise silently unmasked);
    * node pooling with the reference's ``view`` scrambling (model.py:377-427), classification head and loss are one launch
      each (``tagan_b200.head``).

    ``config``: anything with the ``TAGANConfig`` attributes used here (a dict works)."""

    def __init__(self, config):
        super().__init__()
        from .head import ClassificationModule
        from types import SimpleNamespace
        cfg = SimpleNamespace(**config) if isinstance(config, dict) else config
        g = lambda k, d: getattr(cfg, k, d)  # noqa: E731
        self.config = cfg
        hd, heads, drop = cfg.hidden_dim, cfg.num_heads, g("dropout", 0.1)
        self.hidden_dim, self.output_dim = hd, g("output_dim", 1)
        use_ln = g("use_layer_norm", True)
        learnable = g("learnable_distance", False)
        self.node_embedding = nn.Linear(cfg.node_feature_dim, hd)
        efd = g("edge_feature_dim", 0) if g("use_edge_features", True) else 0
        self.edge_embedding = nn.Linear(efd, hd) if efd > 0 else None
        metric = "scaled_dot_product" if learnable else "euclidean"                    # model.py:80
        self.geometric_attention_layers = nn.ModuleList(
            [TAGANGraphAttention(hd, heads, drop, metric, use_ln, learnable) for _ in range(g("num_layers", 2))])
        self.temporal_propagation = TemporalPropagation(hd, hd, drop, g("time_aware", True), g("bidirectional", False), use_ln,
                                                        g("use_skip_connection", True), g("use_gating", True),
                                                        g("temporal_window_size", 3), g("aggregation_method", "mean"),
                                                        g("use_residual", True))
        self.temporal_attention = AsymmetricTemporalAttention(hd, heads, drop, causal=g("causal_attention", False), time_aware=True,
                                                              use_layer_norm=use_ln, asymmetric_window_size=g("window_size", 5),
                                                              relative_position_bias=g("asymmetric_temporal_bias", True))
        self.classification_head = ClassificationModule(hd, self.output_dim, drop, use_ln)
        self.skip_layer_norm = LayerNorm(hd) if use_ln else None
        nn.init.xavier_uniform_(self.node_embedding.weight)
        nn.init.zeros_(self.node_embedding.bias)
        if self.edge_embedding is not None:
            nn.init.xavier_uniform_(self.edge_embedding.weight)
            nn.init.zeros_(self.edge_embedding.bias)

    @staticmethod
    def _unpack(snapshot):
        if isinstance(snapshot, dict):
            return snapshot["x"], snapshot["edge_index"]
        if isinstance(snapshot, (tuple, list)) and len(snapshot) >= 4:
            return snapshot[0], snapshot[1]
        raise ValueError("snapshot must be a dict with keys x / edge_index / edge_attr / node_ids or a 4-tuple")

    def forward(self, graph_sequence, labels=None):
        """``graph_sequence``: list of T snapshots (dicts or ``(x, edge_index, edge_attr, node_ids)`` tuples, model.py:168-171),
        or a ``tagan_b200.loader.PackedSequence`` already on the device.  Returns the reference's dict
        (``logits``, ``predictions``, ``loss``)."""
        from . import fused
        from .head import pack_padded, pool_blocks
        from .loader import PackedSequence
        dev = self.node_embedding.weight.device
        if isinstance(graph_sequence, PackedSequence):
            seq = graph_sequence
        else:
            pairs = [self._unpack(s) for s in graph_sequence]
            seq = PackedSequence.from_snapshots([p[0] for p in pairs], [p[1] for p in pairs]).to(dev)
        t_steps, sizes, maxn = seq.num_snapshots, seq.sizes, seq.max_nodes
        batched = (ops.BATCHED_CSR and t_steps <= ops.MAX_CSR_BATCH
                   and all(l.geometric_attention._fused_ok() for l in self.geometric_attention_layers))
        if batched:                                       # ONE block-diagonal graph over the packed rows
            csrs = [ops.build_csr_batched([seq.edge_index(t) for t in range(t_steps)], sizes, transpose=torch.is_grad_enabled())]
        else:
            csrs = [ops.build_csr(seq.edge_index(t), sizes[t], transpose=torch.is_grad_enabled(), validate=False)
                    for t in range(t_steps)]
        x = ops.linear(seq.x, self.node_embedding.weight, self.node_embedding.bias)                      # model.py:233
        skip = x
        for i, layer in enumerate(self.geometric_attention_layers):                                         # :244-262
            ga = layer.geometric_attention
            ga._check_shape()
            if ga._fused_ok():
                x = fused.geo_layer(ga, x, csrs)
            else:
                x = torch.cat([ga.forward_csr(x[seq.offsets_host[t]:seq.offsets_host[t + 1]], csrs[t]) for t in range(t_steps)], 0)
            if i == 0:
                sk = self.skip_layer_norm(skip) if self.skip_layer_norm is not None else skip
                x = ops.add(x, sk)
        # temporal propagation: the reference's call never completes and falls back to x (model.py:291-309)
        if len(set(sizes)) == 1:
            phys = x.view(t_steps, maxn, self.hidden_dim)
        else:
            phys = pack_padded(x, seq.offsets, t_steps, maxn)                                                # temporal_attention.py:928-976
        # model.py:336-361 passes ones(T,T): resolved on the host from the shapes alone (no device read, graph-capturable)
        out = self.temporal_attention(phys, time_major=True,
                                      resolved_mask=self.temporal_attention.all_ones_mask_spec(t_steps))   # [maxN,T,H] view of [T,maxN,H]
        gf = pool_blocks(out.permute(1, 0, 2), maxn, t_steps, True)                                          # model.py:377-427
        bsz = 1
        if labels is not None and labels.dim() > 0:
            bsz = labels.shape[0]
        if bsz > 1:                                   # graph_features rows beyond the first stay zero in the reference
            gf3 = torch.cat([gf.unsqueeze(0), torch.zeros(bsz - 1, t_steps, self.hidden_dim, device=dev)], 0)
        else:
            gf3 = gf.unsqueeze(0)
        head = self.classification_head.classification_head
        lab = cls = None
        if labels is not None:
            labels = labels.to(dev)
            if labels.dtype == torch.bool:
                labels = labels.long()
            if self.output_dim > 1 and labels.dim() == 1:                                                    # model.py:436-438
                cls = labels.long().contiguous()
            else:                                                                                            # TemporalLossFunction :420-456
                lab = labels.float().reshape(-1, self.output_dim) if labels.numel() % self.output_dim == 0 else None
                if lab is None or (lab.shape[0] != bsz and bsz != 1):
                    raise ValueError(f"Predictions shape {(bsz, self.output_dim)} does not match targets shape {tuple(labels.shape)}")
        logits, loss = head.forward_loss(gf3, lab, cls)
        predictions = torch.sigmoid(logits) if self.output_dim == 1 else torch.softmax(logits, dim=1)        # :447-459
        return {"logits": logits, "predictions": predictions, "loss": loss}
