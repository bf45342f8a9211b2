"""Input format + loader of the device pipeline.

The reference's data package (``src/tagan/data``: TemporalGraphDataset / DataLoader) is missing from its repository; what
its call sites hand to ``TAGAN.forward`` is a list of T snapshots, each a dict ``{x, edge_index, edge_attr, node_ids,
timestep}`` (debug_tagan.py:56-61) or a tuple ``(x, edge_index, edge_attr, node_ids)`` (model.py:168-171), moved to the
device one snapshot at a time inside the forward loop (model.py:216-228).  ``PackedSequence`` is that same information as
FOUR flat arrays -- the wire / on-disk format of this framework:

    x          [sum N_t, F] fp32      node features of all snapshots, snapshot-major
    edges      [2, sum E_t] int64     edge lists, node ids LOCAL to their snapshot
    offsets    [T+1] int32            row range of snapshot t in ``x``
    eoffsets   [T+1] int64            column range of snapshot t in ``edges``
    (+ optional node_ids [sum N_t] int64, labels)

so a sequence crosses PCIe as four ``cudaMemcpyAsync`` from pinned memory instead of 3T small copies, and
``SequenceLoader`` keeps the next sequence's copy in flight on a side stream while the current one trains.
``save`` / ``load`` write the arrays as one ``.npz``-style file (numpy; no pickle of tensors)."""
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch


@dataclass
class PackedSequence:
    x: torch.Tensor
    edges: torch.Tensor
    offsets: torch.Tensor
    eoffsets: torch.Tensor
    offsets_host: List[int]
    eoffsets_host: List[int]
    node_ids: Optional[torch.Tensor] = None
    labels: Optional[torch.Tensor] = None
    ready: Optional[torch.cuda.Event] = field(default=None, repr=False)

    # -- construction -----------------------------------------------------------------------
    @staticmethod
    def from_snapshots(xs: Sequence[torch.Tensor], edge_indices: Sequence[torch.Tensor], node_ids: Optional[Sequence] = None,
                       labels: Optional[torch.Tensor] = None, pin: bool = False) -> "PackedSequence":
        """Pack the reference's per-snapshot tensors (CPU or CUDA) into the flat format."""
        sizes = [int(x.shape[0]) for x in xs]
        esizes = [int(e.shape[1]) for e in edge_indices]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        eoffs = [0]
        for s in esizes:
            eoffs.append(eoffs[-1] + s)
        x = torch.cat([t.float() for t in xs], 0)
        edges = torch.cat([e.long() for e in edge_indices], 1) if sum(esizes) else torch.zeros(2, 0, dtype=torch.long, device=x.device)
        dev = x.device
        ids = None
        if node_ids is not None:
            ids = torch.cat([torch.as_tensor(i, dtype=torch.long) for i in node_ids]).to(dev)
        seq = PackedSequence(x, edges.to(dev), torch.tensor(offs, dtype=torch.int32, device=dev),
                             torch.tensor(eoffs, dtype=torch.int64, device=dev), offs, eoffs, ids, labels)
        return seq.pin_memory() if pin and dev.type == "cpu" else seq

    # -- views ------------------------------------------------------------------------------
    @property
    def num_snapshots(self) -> int:
        return len(self.offsets_host) - 1

    @property
    def sizes(self) -> List[int]:
        o = self.offsets_host
        return [o[i + 1] - o[i] for i in range(len(o) - 1)]

    @property
    def max_nodes(self) -> int:
        return max(self.sizes) if self.num_snapshots else 0

    def edge_index(self, t: int) -> torch.Tensor:
        return self.edges[:, self.eoffsets_host[t]:self.eoffsets_host[t + 1]]

    def snapshot_x(self, t: int) -> torch.Tensor:
        return self.x[self.offsets_host[t]:self.offsets_host[t + 1]]

    def to_snapshots(self):
        """Back to the reference's list of ``(x, edge_index, edge_attr=None, node_ids)`` tuples."""
        out = []
        for t in range(self.num_snapshots):
            ids = None
            if self.node_ids is not None:
                ids = self.node_ids[self.offsets_host[t]:self.offsets_host[t + 1]].tolist()
            out.append((self.snapshot_x(t), self.edge_index(t), None, ids if ids is not None else list(range(self.sizes[t]))))
        return out

    # -- movement ---------------------------------------------------------------------------
    def _map(self, fn) -> "PackedSequence":
        return PackedSequence(fn(self.x), fn(self.edges), fn(self.offsets), fn(self.eoffsets), self.offsets_host, self.eoffsets_host,
                              fn(self.node_ids) if self.node_ids is not None else None,
                              fn(self.labels) if self.labels is not None else None)

    def pin_memory(self) -> "PackedSequence":
        return self._map(lambda t: t.pin_memory())

    def to(self, device, non_blocking: bool = False) -> "PackedSequence":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.edges, self.offsets, self.eoffsets))

    # -- on-disk format -----------------------------------------------------------------------
    def save(self, path: str) -> None:
        arrays = {"x": self.x.cpu().numpy(), "edges": self.edges.cpu().numpy(), "offsets": np.asarray(self.offsets_host, np.int32),
                  "eoffsets": np.asarray(self.eoffsets_host, np.int64)}
        if self.node_ids is not None:
            arrays["node_ids"] = self.node_ids.cpu().numpy()
        if self.labels is not None:
            arrays["labels"] = self.labels.cpu().numpy()
        with open(path, "wb") as f:
            np.savez(f, **arrays)

    @staticmethod
    def load(path: str, pin: bool = False) -> "PackedSequence":
        with np.load(path, allow_pickle=False) as z:
            offs, eoffs = z["offsets"].tolist(), z["eoffsets"].tolist()
            seq = PackedSequence(torch.from_numpy(z["x"]), torch.from_numpy(z["edges"]), torch.tensor(offs, dtype=torch.int32),
                                 torch.tensor(eoffs, dtype=torch.int64), offs, eoffs,
                                 torch.from_numpy(z["node_ids"]) if "node_ids" in z.files else None,
                                 torch.from_numpy(z["labels"]) if "labels" in z.files else None)
        return seq.pin_memory() if pin else seq


class SequenceLoader:
    """Iterate device-resident ``PackedSequence``s with the NEXT one's host-to-device copy in flight on a copy stream
    (what ``model.py:216-228``'s per-snapshot ``.to(device)`` becomes).  ``source`` yields pinned ``PackedSequence``s
    (or paths written by ``PackedSequence.save``)."""

    def __init__(self, source: Iterable, device, prefetch: bool = True):
        self.source, self.device, self.prefetch = source, torch.device(device), prefetch
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def _stage(self, item) -> PackedSequence:
        seq = PackedSequence.load(item, pin=True) if isinstance(item, str) else item
        if self.copy_stream is None:
            return seq.to(self.device)
        with torch.cuda.stream(self.copy_stream):
            dev_seq = seq.to(self.device, non_blocking=True)
            dev_seq.ready = torch.cuda.Event()
            dev_seq.ready.record(self.copy_stream)
        dev_seq._host_keepalive = seq                 # the pinned source must outlive the asynchronous copy
        return dev_seq

    def __iter__(self) -> Iterator[PackedSequence]:
        it = iter(self.source)
        nxt = None
        for item in it:
            cur, nxt = nxt, self._stage(item)
            if not self.prefetch:
                cur, nxt = nxt, None
            if cur is not None:
                yield self._finish(cur)
        if nxt is not None:
            yield self._finish(nxt)

    def _finish(self, seq: PackedSequence) -> PackedSequence:
        if seq.ready is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(seq.ready)
            for t in (seq.x, seq.edges, seq.offsets, seq.eoffsets, seq.node_ids, seq.labels):
                if t is not None:
                    t.record_stream(cur)
        return seq
