"""Drop-in ``NodeMemoryBank`` on dense device tables (hot-path rows c1-c3).

Mirrors the reference class (src/tagan/utils/memory_bank.py:14-360): same constructor, methods and
observable bookkeeping, but the python dicts become ``table[cap,H]`` + ``valid/has_seen/last_seen/
inactivity/frequency`` arrays on the GPU, updated by libtagan_b200 kernels.  Integer bookkeeping and
states are bit-exact with the reference (tests/test_gpu_bank.py).

Node ids: non-negative ints are used as slots directly (the synthetic configs use ``0..N-1``);
any other hashable id is mapped to a fresh slot through a host dict.  Capacity grows by doubling.
Not restated: the NaN-recovery branch (:109-118) -- its ``rand``-based arm cannot be reproduced.
"""
import ctypes as C
import os
import pickle
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .ops import CALLS, _f32c, _ptr, _stream


class NodeMemoryBank:
    POW_TABLE = 4096

    def __init__(self, hidden_dim: int, decay_factor: float = 0.8, max_inactivity: int = 5,
                 device: Optional[torch.device] = None, capacity: int = 1024):
        self.hidden_dim = hidden_dim
        self.decay_factor = decay_factor
        self.max_inactivity = max_inactivity
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tagan_b200.NodeMemoryBank has no CPU path: device must be a CUDA device")
        self._slot_of: Dict[Any, int] = {}      # only for ids that are not plain non-negative ints
        self._id_of: Dict[int, Any] = {}
        self._next_free = 0
        self.check_range = True      # False: trust tensor ids to be < capacity (no host sync per call)
        self._alloc(max(int(capacity), 1))
        self._size_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._consts()

    # -- storage ---------------------------------------------------------------------------
    def _alloc(self, cap: int):
        d = self.device
        self.capacity = cap
        self.table = torch.zeros(cap, self.hidden_dim, dtype=torch.float32, device=d)
        self.valid = torch.zeros(cap, dtype=torch.uint8, device=d)
        self.has_seen = torch.zeros(cap, dtype=torch.uint8, device=d)
        self.last_seen = torch.zeros(cap, dtype=torch.int32, device=d)
        self.inactivity = torch.zeros(cap, dtype=torch.int32, device=d)
        self.frequency_t = torch.zeros(cap, dtype=torch.int32, device=d)
        self._marks = torch.empty(2 * cap, dtype=torch.int32, device=d)

    def _grow(self, need: int):
        if need <= self.capacity:
            return
        cap = self.capacity
        while cap < need:
            cap *= 2
        old = (self.table, self.valid, self.has_seen, self.last_seen, self.inactivity, self.frequency_t)
        n = self.capacity
        self._alloc(cap)
        for new, o in zip((self.table, self.valid, self.has_seen, self.last_seen, self.inactivity, self.frequency_t), old):
            new[:n].copy_(o)

    def _consts(self):
        """python-double scalars rounded to fp32 once, as ``float * float32_tensor`` does in torch."""
        d = self.decay_factor
        w2 = max(0.4, d ** 2)                               # memory_bank.py:124
        w3 = max(0.4, d ** 3)
        self._w23 = (C.c_float * 4)(w2, 1 - w2, w3, 1 - w3)
        n = min(self.POW_TABLE, self.max_inactivity + 2)
        self._pow = torch.tensor([np.float32(d ** k) for k in range(max(n, 2))], dtype=torch.float32, device=self.device)

    def _slots(self, node_ids) -> torch.Tensor:
        if isinstance(node_ids, torch.Tensor):
            ids = node_ids.to(self.device, dtype=torch.int32).contiguous()
            if ids.numel() and self.check_range:
                self._grow(int(ids.max().item()) + 1)
            return ids
        plain = all(isinstance(i, (int, np.integer)) and not isinstance(i, bool) and 0 <= i < 2 ** 31 - 1 for i in node_ids)
        if plain and not self._slot_of:
            arr = np.asarray(list(node_ids), dtype=np.int32)
            if arr.size:
                self._grow(int(arr.max()) + 1)
            return torch.from_numpy(arr).to(self.device)
        # generic hashable ids -> fresh slots
        if not self._slot_of and bool(self.valid.any().item()):
            raise RuntimeError("cannot mix plain integer ids with non-integer ids in one bank")
        out = []
        for i in node_ids:
            s = self._slot_of.get(i)
            if s is None:
                s = self._next_free
                self._next_free += 1
                self._slot_of[i] = s
                self._id_of[s] = i
            out.append(s)
        if out:
            self._grow(max(out) + 1)
        return torch.tensor(out, dtype=torch.int32, device=self.device)

    def _key(self, slot: int):
        return self._id_of.get(slot, slot) if self._slot_of else slot

    def _slot_of_key(self, node_id) -> Optional[int]:
        if self._slot_of:
            return self._slot_of.get(node_id)
        if isinstance(node_id, (int, np.integer)) and 0 <= node_id < self.capacity:
            return int(node_id)
        return None

    # -- reference API ---------------------------------------------------------------------
    @property
    def size(self) -> int:
        """Refreshed only by ``update`` (reference :169); reading it syncs the stream."""
        return int(self._size_dev.item())

    def update(self, node_ids, states: torch.Tensor, timestep: int = 0, verbose: bool = False):
        """memory_bank.py:65-173 in one enqueue (4 kernels), no host sync."""
        lib = _lib.load()
        states = _f32c(states.to(self.device))
        if states.dim() == 1:
            states = states.unsqueeze(0)
        if states.dim() != 2 or states.shape[1] != self.hidden_dim:
            raise ValueError(f"NodeMemoryBank.update: states must be [M, {self.hidden_dim}], got {tuple(states.shape)}")
        if states.untyped_storage().data_ptr() == self.table.untyped_storage().data_ptr():
            states = states.clone()            # a view of the table (e.g. an old get_state row): the kernels assume no aliasing
        ids = self._slots(node_ids)
        m = ids.numel()
        lds = states.stride(0) if states.shape[0] > 1 else self.hidden_dim
        rc = lib.tagan_bank_update(_ptr(self.table), _ptr(self.valid), _ptr(self.has_seen), _ptr(self.last_seen),
                                   _ptr(self.inactivity), _ptr(self.frequency_t), _ptr(ids) if m else None, m,
                                   _ptr(states) if states.numel() else None, lds, states.shape[0], self.hidden_dim,
                                   self.capacity, int(timestep), self._w23, _ptr(self._pow), self._pow.numel(),
                                   float(self.decay_factor), int(self.max_inactivity), _ptr(self._marks),
                                   _ptr(self._size_dev), _ptr(self._status), _stream())
        _lib.check(rc, "tagan_bank_update")
        CALLS["n"] += 4

    def get_states(self, node_ids) -> torch.Tensor:
        """memory_bank.py:187-211."""
        lib = _lib.load()
        ids = self._slots(node_ids)
        m = ids.numel()
        out = torch.empty(m, self.hidden_dim, dtype=torch.float32, device=self.device)
        rc = lib.tagan_bank_gather(_ptr(self.table), _ptr(self.valid), _ptr(self.inactivity), _ptr(ids) if m else None,
                                   _ptr(out) if m else None, m, self.hidden_dim, self.capacity, _ptr(self._status),
                                   _stream())
        _lib.check(rc, "tagan_bank_gather")
        CALLS["n"] += 2
        return out

    def known_mask(self, node_ids) -> torch.Tensor:
        """bool [M]: which ids currently have a stored state (``node_id in node_states``), on device."""
        ids = self._slots(node_ids).long()
        return self.valid[ids].bool()

    def get_state(self, node_id) -> Optional[torch.Tensor]:
        s = self._slot_of_key(node_id)
        if s is None or not bool(self.valid[s].item()):
            return None
        return self.table[s].clone()           # the reference hands out a tensor that later updates replace, never mutate

    def update_state(self, node_id, state: torch.Tensor, timestep: int = 0):
        self.update([node_id], state.unsqueeze(0), timestep)          # :235-244

    def get_active_nodes(self) -> List[Any]:
        return [self._key(int(s)) for s in torch.nonzero(self.valid).flatten().tolist()]

    def decay_all(self):
        lib = _lib.load()
        rc = lib.tagan_bank_decay_all(_ptr(self.table), _ptr(self.valid), float(np.float32(self.decay_factor)),
                                      self.hidden_dim, self.capacity, _stream())
        _lib.check(rc, "tagan_bank_decay_all")
        CALLS["n"] += 1

    def reset(self):
        for t in (self.table, self.valid, self.has_seen, self.last_seen, self.inactivity, self.frequency_t, self._size_dev):
            t.zero_()
        self._slot_of.clear()
        self._id_of.clear()
        self._next_free = 0

    # dict-style views of the reference attributes (host copies; for inspection / pickling only)
    @property
    def node_states(self) -> Dict[Any, torch.Tensor]:
        snap = self.table.clone()              # detached snapshot: later updates / decay must not mutate what was handed out
        return {self._key(int(s)): snap[s] for s in torch.nonzero(self.valid).flatten().tolist()}

    @property
    def inactivity_counter(self) -> Dict[Any, int]:
        ina = self.inactivity.cpu()
        return {self._key(int(s)): int(ina[s]) for s in torch.nonzero(self.valid).flatten().tolist()}

    @property
    def frequency(self) -> Dict[Any, int]:
        f = self.frequency_t.cpu()
        return {self._key(int(s)): int(f[s]) for s in torch.nonzero(f).flatten().tolist()}

    def save(self, filepath: str):
        """Same pickle layout as the reference (:246-272)."""
        os.makedirs(os.path.dirname(os.path.abspath(filepath)), exist_ok=True)
        sd = {"hidden_dim": self.hidden_dim, "decay_factor": self.decay_factor, "max_inactivity": self.max_inactivity,
              "node_states": {k: v.cpu() for k, v in self.node_states.items()},
              "inactivity_counter": self.inactivity_counter}
        with open(filepath, "wb") as f:
            pickle.dump(sd, f)

    @classmethod
    def load(cls, filepath: str, device: Optional[torch.device] = None):
        """Classmethod form (:299-333; it shadows the instance method in the reference too)."""
        with open(filepath, "rb") as f:
            sd = pickle.load(f)
        bank = cls(sd["hidden_dim"], sd["decay_factor"], sd["max_inactivity"], device=device)
        keys = list(sd["node_states"].keys())
        if keys:
            slots = bank._slots(keys).long()
            bank.table[slots] = torch.stack([sd["node_states"][k] for k in keys]).to(bank.device)
            bank.valid[slots] = 1
            ina = torch.tensor([sd["inactivity_counter"].get(k, 0) for k in keys], dtype=torch.int32, device=bank.device)
            bank.inactivity[slots] = ina
        return bank

    def get_memory_stats(self) -> Dict[str, Any]:
        v = self.valid.bool()
        n = int(v.sum().item())
        avg = float(self.inactivity[v].float().mean().item()) if n else 0
        return {"num_nodes": n, "avg_inactivity": avg, "max_inactivity_limit": self.max_inactivity,
                "decay_factor": self.decay_factor, "hidden_dim": self.hidden_dim}

    def check_ids(self):
        """Raise if any id handed to update/get_states was out of range (host sync)."""
        if int(self._status.item()) != 0:
            raise IndexError("node id out of range for the memory bank")

    def __repr__(self) -> str:
        return (f"NodeMemoryBank(hidden_dim={self.hidden_dim}, decay_factor={self.decay_factor}, "
                f"max_inactivity={self.max_inactivity}, active_nodes={int(self.valid.sum().item())})")
