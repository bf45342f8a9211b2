"""GPU parity tests for hot-path rows c1-c3 (NodeMemoryBank) through the C ABI: the traces recorded
from the unmodified reference class (tests/golden/memory_bank.pt) are replayed on the device bank and
every integer array AND every state must be BIT-EXACT; larger random traces are checked against the
numpy oracle; full-size behaviour through size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _equal(bank, valid, inactivity, last_seen_valid, has_seen, frequency, states, tag):
    v = valid.astype(bool)
    cap = len(v)
    assert np.array_equal(bank.valid[:cap].cpu().numpy().astype(bool), v), tag
    assert np.array_equal(bank.inactivity[:cap].cpu().numpy()[v], inactivity[v]), tag
    hs = has_seen.astype(bool)
    assert np.array_equal(bank.has_seen[:cap].cpu().numpy().astype(bool) & v, hs & v), tag
    assert np.array_equal(bank.last_seen[:cap].cpu().numpy()[hs & v], last_seen_valid[hs & v]), tag
    assert np.array_equal(bank.frequency_t[:cap].cpu().numpy(), frequency), tag
    got = bank.table[:cap].cpu().numpy()
    assert np.array_equal(got[v].view(np.uint32), states[v].view(np.uint32)), tag     # bit-exact states


@pytest.mark.parametrize("key", ["kat", "random"])
def test_bank_replays_reference_trace_bit_exact(dev, golden, key):
    import tagan_b200
    c = golden("memory_bank.pt")[key]
    bank = tagan_b200.NodeMemoryBank(c["hidden"], c["decay"], c["max_inactivity"], device=dev, capacity=c["cap"])
    for i, op in enumerate(c["trace"]):
        if op["op"] == "update":
            bank.update(op["ids"], torch.from_numpy(op["states"]).to(dev), op["t"])
        elif op["op"] == "update_state":
            bank.update_state(op["ids"][0], torch.from_numpy(op["states"]).to(dev), op["t"])
        elif op["op"] == "get_states":
            got = bank.get_states(op["ids"]).cpu().numpy()
            assert np.array_equal(got.view(np.uint32), op["got"].view(np.uint32)), i
        else:
            bank.decay_all()
        a = op["after"]
        _equal(bank, a["valid"], a["inactivity"], a["last_seen"], (a["last_seen"] >= 0), a["frequency"], a["states"],
               (key, i, op["op"]))
        if op["op"] in ("update", "update_state"):
            assert bank.size == a["size"], (key, i)
    bank.check_ids()


def test_bank_random_trace_vs_oracle(dev):
    import tagan_b200
    rng = np.random.RandomState(7)
    cap, hd = 5000, 32
    bank = tagan_b200.NodeMemoryBank(hd, 0.8, 3, device=dev, capacity=cap)
    ora = R.BankOracle(hd, cap, 0.8, 3)
    t = 0
    for it in range(25):
        m = int(rng.randint(1, 2500))
        ids = rng.randint(0, cap, size=m)                       # duplicates guaranteed
        st = rng.randn(m, hd).astype(np.float32)
        bank.update(torch.from_numpy(ids.astype(np.int32)).to(dev), torch.from_numpy(st).to(dev), t)
        ora.update(ids, st, t)
        if it % 4 == 1:
            q = rng.randint(0, cap, size=300)
            got = bank.get_states(q.tolist()).cpu().numpy()
            exp = ora.get_states(q)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
        if it % 7 == 3:
            bank.decay_all()
            ora.decay_all()
        _equal(bank, ora.valid, ora.inactivity, ora.last_seen, ora.has_seen, ora.frequency, ora.states, it)
        assert bank.size == ora.size
        t += int(rng.randint(1, 4))


def test_bank_generic_hashable_ids(dev):
    import tagan_b200
    bank = tagan_b200.NodeMemoryBank(4, 0.5, 2, device=dev)
    bank.update(["a", ("b", 1), "c"], torch.arange(12, dtype=torch.float32).view(3, 4).to(dev), 0)
    assert sorted(map(str, bank.get_active_nodes())) == sorted(["a", "('b', 1)", "c"])
    assert torch.equal(bank.get_state(("b", 1)).cpu(), torch.tensor([4., 5., 6., 7.]))
    assert bank.get_state("zzz") is None
    bank.update(["a"], torch.ones(1, 4, device=dev), 1)
    assert torch.equal(bank.get_state("c").cpu(), torch.tensor([8., 9., 10., 11.]) * 0.5)


def test_bank_save_load_roundtrip(dev, tmp_path):
    import tagan_b200
    bank = tagan_b200.NodeMemoryBank(8, 0.8, 5, device=dev)
    bank.update([3, 9, 4], torch.randn(3, 8, device=dev), 0)
    bank.update([3], torch.randn(1, 8, device=dev), 1)
    f = str(tmp_path / "bank.pkl")
    bank.save(f)
    b2 = tagan_b200.NodeMemoryBank.load(f, device=dev)
    assert b2.inactivity_counter == bank.inactivity_counter
    for k, v in bank.node_states.items():
        assert torch.equal(b2.node_states[k], v)


def test_bank_full_size_properties(dev):
    """Config-3 node count (100k ids, H=128): k missed calls leave decay^(1+..+k) (triangular decay,
    SURVEY 3.5), pruning happens exactly after max_inactivity, and update is idempotent on the integer
    state when replayed with the same timestep."""
    import tagan_b200
    n, hd = 100_000, 128
    bank = tagan_b200.NodeMemoryBank(hd, 0.8, 3, device=dev, capacity=n)
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    x = torch.randn(n, hd, device=dev)
    bank.update(ids, x, 0)
    assert bank.size == n
    assert torch.equal(bank.table, x)
    half = ids[: n // 2]
    for step, expo in ((1, 1), (2, 3), (3, 6)):
        bank.update(half, x[: n // 2], step)
        d = np.float32(1.0)
        for k in range(1, step + 1):
            d = np.float32(d * np.float32(0.8 ** k))
        torch.testing.assert_close(bank.table[n // 2:], x[n // 2:] * float(0.8 ** expo), rtol=1e-6, atol=0)
        assert bank.size == n
    bank.update(half, x[: n // 2], 4)                         # inactivity 4 > 3 => pruned
    assert bank.size == n // 2
    assert int(bank.valid[n // 2:].sum().item()) == 0
    assert torch.equal(bank.table[: n // 2], x[: n // 2])
    assert torch.equal(bank.get_states(ids[n // 2: n // 2 + 10]), torch.zeros(10, hd, device=dev))
