"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference modules (build container only).

``/root/reference`` exists only in the build container, never on the GPU box; this loader is
used by ``oracle/make_golden.py`` (fixture generation) and by CPU tests that are skipped when
the reference is absent.  Recipe from SURVEY.md section 8c: stub the plotting modules the
reference imports at module scope (``src/tagan/utils/metrics.py:16``), put the checkout on
``sys.path`` and silence the unconditional ``print`` calls inside its hot path
(``geometric_attention.py:476-507``, ``temporal_attention.py:133-208``).
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TAGAN_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
    "seaborn", "plotly", "plotly.graph_objects", "plotly.express", "plotly.subplots",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "tagan"))


@contextlib.contextmanager
def quiet():
    """Swallow the reference's unconditional debug prints."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


_loaded = None


def load():
    """Return a namespace holding the reference classes on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for m in _STUBS:
        sys.modules.setdefault(m, types.ModuleType(m))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with quiet():
        from src.tagan.layers.graph_attention import TAGANGraphAttention
        from src.tagan.layers.geometric_attention import GeometricAttention, DistanceMetric
        from src.tagan.layers.temporal_attention import AsymmetricTemporalAttention, TimeEncoding
        from src.tagan.layers.temporal_propagation import (
            TemporalGRUCell, TemporalEvolutionLayer, TemporalSkipConnection,
            TemporalGatingUnit, TemporalPropagation)
        from src.tagan.utils.memory_bank import NodeMemoryBank
        from src.tagan.utils.config import TAGANConfig
        from src.tagan.model import TAGAN
    ns = types.SimpleNamespace(
        TAGANGraphAttention=TAGANGraphAttention, GeometricAttention=GeometricAttention,
        DistanceMetric=DistanceMetric, AsymmetricTemporalAttention=AsymmetricTemporalAttention,
        TimeEncoding=TimeEncoding, TemporalGRUCell=TemporalGRUCell,
        TemporalEvolutionLayer=TemporalEvolutionLayer, TemporalSkipConnection=TemporalSkipConnection,
        TemporalGatingUnit=TemporalGatingUnit, TemporalPropagation=TemporalPropagation,
        NodeMemoryBank=NodeMemoryBank, TAGANConfig=TAGANConfig, TAGAN=TAGAN)
    _loaded = ns
    return ns
