"""Drop-in ``nn.Module`` replacements for the reference's hot-path layers.

Same constructor arguments, parameter names (``state_dict`` keys) and forward signatures as the
reference classes, so reference checkpoints load unchanged and ``tagan_b200.patch(model)`` can
swap them into a reference ``TAGAN``.  All arithmetic runs in libtagan_b200.so; dropout is the
only op left to torch (its RNG cannot be matched, parity is defined at dropout=0 / eval()).
"""
import math
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class GeometricAttention(nn.Module):
    """Mirror of reference ``GeometricAttention`` (src/tagan/layers/geometric_attention.py:228-598)
    operating on a CSR instead of a dense ``[1,N,N]`` mask."""

    def __init__(self, hidden_dim: int, num_heads: int = 8, dropout: float = 0.1,
                 distance_metric: str = "scaled_dot_product", use_layer_norm: bool = True,
                 learnable_distance: bool = False):
        super().__init__()
        if distance_metric not in ops.METRIC_ID:
            # same failure as DistanceMetric.get_metric (:196-225); "mahalanobis" is unreachable there too
            raise ValueError(f"Unknown distance metric: {distance_metric}")
        assert hidden_dim % num_heads == 0, "Hidden dimension must be divisible by number of heads"
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.dropout_prob = dropout
        self.distance_metric = distance_metric
        self.use_layer_norm = use_layer_norm
        self.learnable_distance = learnable_distance
        self.head_dim = hidden_dim // num_heads
        self.q_linear = nn.Linear(hidden_dim, hidden_dim)
        self.k_linear = nn.Linear(hidden_dim, hidden_dim)
        self.v_linear = nn.Linear(hidden_dim, hidden_dim)
        self.output_proj = nn.Linear(hidden_dim, hidden_dim)
        if use_layer_norm:
            self.layer_norm1 = nn.LayerNorm(hidden_dim)
            self.layer_norm2 = nn.LayerNorm(hidden_dim)
        self.attn_dropout = nn.Dropout(dropout)
        self.output_dropout = nn.Dropout(dropout)
        if learnable_distance and distance_metric in ("gaussian_kernel", "rbf_kernel"):
            self.distance_param = nn.Parameter(torch.ones(num_heads))
        self._init_parameters()

    def _init_parameters(self):
        for lin in (self.q_linear, self.k_linear, self.v_linear, self.output_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
        if hasattr(self, "distance_param"):
            nn.init.constant_(self.distance_param, 1.0 if self.distance_metric == "gaussian_kernel" else 0.1)

    def forward_csr(self, x: torch.Tensor, csr: ops.CSR, return_attention_weights: bool = False):
        """x ``[N,H]`` -> ``[N,H]`` (+ per-entry weights ``[nnz,h]`` aligned with ``csr.row/col``)."""
        ln = self.use_layer_norm
        xn = ops.layer_norm(x, self.layer_norm1.weight, self.layer_norm1.bias) if ln else x
        w_qkv = torch.cat([self.q_linear.weight, self.k_linear.weight, self.v_linear.weight], 0)
        b_qkv = torch.cat([self.q_linear.bias, self.k_linear.bias, self.v_linear.bias], 0)
        qkv = ops.linear(xn, w_qkv, b_qkv)
        param = getattr(self, "distance_param", None)
        ctx, attn = ops.geo_attention_core(qkv, csr, self.num_heads, self.distance_metric, param,
                                           want_attn=return_attention_weights)
        o = ops.linear(ctx, self.output_proj.weight, self.output_proj.bias)
        o = self.output_dropout(o)
        if ln:
            out = ops.layer_norm(o, self.layer_norm2.weight, self.layer_norm2.bias, res=x)
        else:
            out = ops.add(o, x)
        return (out, attn) if return_attention_weights else out

    def extra_repr(self) -> str:
        return (f"hidden_dim={self.hidden_dim}, num_heads={self.num_heads}, distance_metric={self.distance_metric}, "
                f"learnable_distance={self.learnable_distance}, use_layer_norm={self.use_layer_norm}, "
                f"dropout={self.dropout_prob}")


class TAGANGraphAttention(nn.Module):
    """Mirror of reference ``TAGANGraphAttention`` (src/tagan/layers/graph_attention.py:15-136).

    ``forward(x, edge_index, edge_attr=None, return_attention_weights=False)`` has the
    reference's signature.  ``edge_attr`` is accepted and ignored exactly as in the reference
    (:107-112).  With ``return_attention_weights`` the reference returns the placeholder
    ``{"node_attention": None}`` (:121); we keep that key and add the real per-entry weights under
    ``"edge_attention"`` / ``"edge_row"`` / ``"edge_col"``.
    A prebuilt CSR may be passed as ``edge_index`` (``ops.CSR``) to amortise the build across
    layers and across forward/backward.
    """

    def __init__(self, hidden_dim: int, num_heads: int = 8, dropout: float = 0.1,
                 distance_metric: str = "scaled_dot_product", use_layer_norm: bool = True,
                 learnable_distance: bool = False):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.geometric_attention = GeometricAttention(hidden_dim, num_heads, dropout, distance_metric,
                                                      use_layer_norm, learnable_distance)
        self.validate_indices = True

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights: bool = False):
        if isinstance(edge_index, ops.CSR):
            csr = edge_index
        else:
            if edge_index is None:
                # the reference then attends over ALL node pairs (no mask, graph_attention.py:95-96);
                # that dense O(N^2) mode is out of scope for the sparse kernel
                raise NotImplementedError("edge_index=None (dense all-pairs attention) is not supported")
            csr = ops.build_csr(edge_index.to(x.device), x.shape[0], transpose=torch.is_grad_enabled(),
                                validate=self.validate_indices)
        if return_attention_weights:
            out, attn = self.geometric_attention.forward_csr(x, csr, True)
            nnz = csr.nnz
            weights = {"node_attention": None, "edge_attention": attn[:nnz], "edge_row": csr.row[:nnz],
                       "edge_col": csr.col[:nnz]}
            return out, weights
        return self.geometric_attention.forward_csr(x, csr)

    def extra_repr(self) -> str:
        return f"hidden_dim={self.hidden_dim}"
