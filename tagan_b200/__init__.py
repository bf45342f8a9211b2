"""tagan_b200 -- B200 (sm_100a) kernels behind TAGAN's per-snapshot attention-and-propagation core.

Host side of the C ABI in ``include/tagan_b200.h``: drop-in ``nn.Module`` mirrors of the
reference layers (``layers``), thin ctypes wrappers + autograd glue (``ops``) and synthetic
workloads (``synth``).  The package name is ``tagan_b200`` because the repository's long name
(``temporal-asymmetric-graph-attention-network_b200``) is not an importable identifier.
"""
from . import _lib, ops  # noqa: F401
from .layers import (AsymmetricTemporalAttention, GeometricAttention, LayerNorm, TAGANGraphAttention,  # noqa: F401
                     TemporalEvolutionLayer, TemporalGatingUnit, TemporalGRUCell, TemporalPropagation,
                     TemporalSkipConnection, TimeEncoding)

from .memory_bank import NodeMemoryBank  # noqa: F401,E402
from .model import TAGANLayer, TAGANModel, forward_node_partitioned, patch  # noqa: F401,E402
from .head import ClassificationModule, FusedAdam, TemporalClassificationHead  # noqa: F401,E402
from .loader import PackedSequence, SequenceLoader  # noqa: F401,E402
from .graphed import GraphedStep  # noqa: F401,E402

__all__ = ["NodeMemoryBank", "TAGANLayer", "TAGANModel", "ClassificationModule", "TemporalClassificationHead", "FusedAdam",
           "PackedSequence", "SequenceLoader", "patch", "GraphedStep", "forward_node_partitioned", "ops", "GeometricAttention", "TAGANGraphAttention", "AsymmetricTemporalAttention", "TimeEncoding",
           "TemporalGRUCell", "TemporalEvolutionLayer", "TemporalSkipConnection", "TemporalGatingUnit",
           "TemporalPropagation"]
