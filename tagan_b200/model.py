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
