"""CPU, build container only (skipped where /root/reference is absent): `tagan_b200.patch` swaps the
hot-path layers of an unmodified reference TAGAN and keeps every state_dict key, shape and value, so
reference checkpoints load into the patched model and vice versa."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.mark.parametrize("learnable", [False, True])
def test_patch_preserves_state_dict(learnable):
    import tagan_b200
    ref = ref_loader.load()
    cfg = ref.TAGANConfig(node_feature_dim=16, edge_feature_dim=8, hidden_dim=64, num_heads=4, num_layers=2,
                          output_dim=1, dropout=0.0, loss_type="bce", use_edge_features=True,
                          learnable_distance=learnable)
    with ref_loader.quiet():
        model = ref.TAGAN(cfg)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tagan_b200.patch(model)
    after = model.state_dict()
    assert list(after.keys()) == list(before.keys())
    for k in before:
        assert after[k].shape == before[k].shape and torch.equal(after[k], before[k]), k
    assert isinstance(model.geometric_attention_layers[0], tagan_b200.TAGANGraphAttention)
    assert isinstance(model.temporal_attention, tagan_b200.AsymmetricTemporalAttention)
    assert isinstance(model.temporal_propagation, tagan_b200.TemporalPropagation)
    # a reference checkpoint loads into the patched model
    with ref_loader.quiet():
        fresh = ref.TAGAN(cfg)
    model.load_state_dict(fresh.state_dict(), strict=True)


def test_module_constructor_signatures_match_reference():
    import inspect
    import tagan_b200
    ref = ref_loader.load()
    pairs = [(tagan_b200.TAGANGraphAttention, ref.TAGANGraphAttention),
             (tagan_b200.GeometricAttention, ref.GeometricAttention),
             (tagan_b200.AsymmetricTemporalAttention, ref.AsymmetricTemporalAttention),
             (tagan_b200.TemporalGRUCell, ref.TemporalGRUCell),
             (tagan_b200.TemporalEvolutionLayer, ref.TemporalEvolutionLayer),
             (tagan_b200.TemporalSkipConnection, ref.TemporalSkipConnection),
             (tagan_b200.TemporalGatingUnit, ref.TemporalGatingUnit),
             (tagan_b200.TemporalPropagation, ref.TemporalPropagation)]
    for mine, theirs in pairs:
        a = inspect.signature(mine.__init__).parameters
        b = inspect.signature(theirs.__init__).parameters
        assert list(a.keys()) == list(b.keys()), (mine.__name__, list(a.keys()), list(b.keys()))
        for k in b:
            assert a[k].default == b[k].default, (mine.__name__, k)
    bank_a = inspect.signature(tagan_b200.NodeMemoryBank.__init__).parameters
    bank_b = inspect.signature(ref.NodeMemoryBank.__init__).parameters
    assert list(bank_b.keys()) == list(bank_a.keys())[:len(bank_b)]
    for name in ("update", "get_state", "get_states", "get_active_nodes", "decay_all", "reset", "update_state",
                 "save", "load", "get_memory_stats"):
        assert hasattr(tagan_b200.NodeMemoryBank, name), name
