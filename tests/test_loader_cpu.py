"""CPU: the packed wire / on-disk format of tagan_b200.loader (host logic only)."""
import torch


def test_packed_sequence_roundtrip(tmp_path):
    from tagan_b200 import PackedSequence
    torch.manual_seed(0)
    sizes = [5, 9, 1]
    xs = [torch.randn(n, 6) for n in sizes]
    es = [torch.randint(0, n, (2, 2 * n)) for n in sizes]
    ids = [list(range(10, 10 + n)) for n in sizes]
    seq = PackedSequence.from_snapshots(xs, es, node_ids=ids, labels=torch.tensor([[1.0]]))
    assert seq.num_snapshots == 3 and seq.sizes == sizes and seq.max_nodes == 9
    assert seq.offsets.tolist() == [0, 5, 14, 15] and seq.eoffsets.tolist() == [0, 10, 28, 30]
    for t in range(3):
        assert torch.equal(seq.snapshot_x(t), xs[t]) and torch.equal(seq.edge_index(t), es[t])
    back = seq.to_snapshots()
    assert all(torch.equal(b[0], x) and torch.equal(b[1], e) and b[3] == i for b, x, e, i in zip(back, xs, es, ids))
    p = str(tmp_path / "s.tagan")
    seq.save(p)
    again = PackedSequence.load(p)
    assert torch.equal(again.x, seq.x) and torch.equal(again.edges, seq.edges) and again.offsets_host == seq.offsets_host
    assert torch.equal(again.node_ids, seq.node_ids) and torch.equal(again.labels, seq.labels)
    assert seq.nbytes() == seq.x.numel() * 4 + seq.edges.numel() * 8 + 4 * 4 + 4 * 8


def test_model_state_dict_keys_match_reference_golden(golden):
    """TAGANModel carries exactly the reference's parameter names (a reference checkpoint loads unchanged)."""
    import tagan_b200
    for c in golden("tagan_model.pt"):
        m = tagan_b200.TAGANModel(dict(c["cfg"]))
        res = m.load_state_dict(c["sd"])
        assert not res.missing_keys and not res.unexpected_keys
        assert set(dict(m.named_parameters())) == set(c["grads"])


def test_batched_csr_has_no_cpu_path():
    """The block-diagonal CSR build, like every product entry point, refuses host tensors instead of falling back."""
    import pytest
    import torch
    from tagan_b200 import ops
    eis = [torch.randint(0, 5, (2, 7)), torch.randint(0, 4, (2, 3))]
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.build_csr_batched(eis, [5, 4])
    with pytest.raises(ValueError):
        ops.build_csr_batched([], [])
