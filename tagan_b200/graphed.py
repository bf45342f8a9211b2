"""Whole-step CUDA graph for the TAGAN hot path.

The step is sync-free (device-built CSR, no host reads), so the H2D copies of the T snapshots, the layer forward,
the loss, the backward and the D2H copy of the loss are captured ONCE into a CUDA graph and replayed per step:
about three thousand kernel launches turn into one ``cudaGraphLaunch``, and the input copies (a forked branch of
the graph) overlap the kernels of earlier snapshots exactly as the eager copy-stream version does.

Static buffers: the device copies of ``xs`` / ``edge_indices`` and every activation / gradient live in the graph's
private pool; new inputs are supplied by writing into the pinned host tensors handed to the constructor.
"""
from typing import Callable, Optional, Sequence

import torch

__all__ = ["GraphedStep"]


class _WaitingList:
    """list whose item t makes the current stream wait for event t on first access"""

    def __init__(self, items, events):
        self.items, self.events = items, events

    def __iter__(self):
        cur = torch.cuda.current_stream()
        for it, ev in zip(self.items, self.events):
            cur.wait_event(ev)
            yield it

    def __len__(self):
        return len(self.items)


class GraphedStep:
    """``step = GraphedStep(fn, xs_host, eis_host, device); loss = step()``.

    ``fn(xs_dev, eis_dev) -> scalar loss tensor`` runs forward AND backward (and, if wanted, the optimizer and the
    gradient all-reduce); ``xs_host`` / ``eis_host`` are pinned host tensors.  ``warmup`` eager runs on a side
    stream precede the capture (lazy initialisation of the library -- function attributes, workspaces -- must not
    happen under capture).

    ``prefetch=False``: the H2D copies are a forked branch INSIDE the graph; snapshot t's kernels wait for
    snapshot t's copy only.  One ``cudaGraphLaunch`` per step, but the step cannot finish faster than its own
    input copy.

    ``prefetch=True`` (default; what a training loop with a prefetching loader does): the graph holds the
    compute only.  The inputs of step k+1 are copied H2D into a staging set on a copy stream WHILE step k's
    graph runs; step k+1 begins with a device-to-device move staging -> static inputs (sub-millisecond), after
    which the staging set is free again.  Every step still performs exactly one H2D copy of its inputs and one D2H
    read of its loss; the host tensors must hold step k+1's data when ``step()`` for step k is called (the first
    prefetch is issued by the constructor).  Ownership: the host tensors are read asynchronously by the copy stream
    after ``launch()``; ``__call__`` waits for that read (``wait_prefetch``) before it returns, so a caller may
    refill them as soon as it has the loss.  Callers of the asynchronous ``launch()`` must call ``wait_prefetch()``
    themselves before touching the host tensors."""

    def __init__(self, fn: Callable, xs_host: Sequence[torch.Tensor], eis_host: Sequence[torch.Tensor],
                 device: torch.device, warmup: int = 2, prefetch: bool = True,
                 before_capture: Optional[Callable] = None, after_replay: Optional[Callable] = None):
        for t in list(xs_host) + list(eis_host):
            if not t.is_pinned():
                raise ValueError("GraphedStep: host inputs must be pinned (the graph holds their addresses)")
        self.xs_host, self.eis_host = list(xs_host), list(eis_host)
        self.device, self.prefetch = device, prefetch
        # eager work enqueued right after every replay (e.g. the data-parallel gradient all-reduce: NCCL collectives
        # are kept OUT of the captured graph so that communicator teardown never waits on a graph that still holds them)
        self.after_replay = after_replay
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.xs_host + self.eis_host)
        side = torch.cuda.Stream(device=device)
        self._copy = copy = torch.cuda.Stream(device=device)
        self._xs = self._alloc_like(self.xs_host)
        self._eis = self._alloc_like(self.eis_host)

        def body_with_copies():
            cur = torch.cuda.current_stream()
            copy.wait_stream(cur)                                       # fork
            events = []
            with torch.cuda.stream(copy):
                for xd, xh, ed, eh in zip(self._xs, self.xs_host, self._eis, self.eis_host):
                    xd.copy_(xh, non_blocking=True)
                    ed.copy_(eh, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                    events.append(ev)
            loss = fn(_WaitingList(self._xs, events), _WaitingList(self._eis, events))
            self.loss_host.copy_(loss.detach().reshape(()), non_blocking=True)
            cur.wait_stream(copy)                                       # join
            return loss

        def body_compute_only():
            loss = fn(self._xs, self._eis)
            self.loss_host.copy_(loss.detach().reshape(()), non_blocking=True)
            return loss

        body = body_compute_only if prefetch else body_with_copies
        if prefetch:
            self._sx = [torch.empty_like(t) for t in self._xs]
            self._se = [torch.empty_like(t) for t in self._eis]
            self._h2d_ready = torch.cuda.Event()
            self._staged = torch.cuda.Event()
            for xd, xh, ed, eh in zip(self._xs, self.xs_host, self._eis, self.eis_host):   # inputs of the warm-up runs
                xd.copy_(xh, non_blocking=True)
                ed.copy_(eh, non_blocking=True)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(device)
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._loss_dev = body()
        if prefetch:
            self._staged.record(torch.cuda.current_stream())
            self._issue_prefetch()

    def _alloc_like(self, hosts):
        """Device twins of the host tensors; one allocation sliced per snapshot when the shapes agree, so that the
        model's ``ops.stack_rows`` is a view."""
        h0 = hosts[0]
        if all(h.shape == h0.shape and h.dtype == h0.dtype for h in hosts):
            return list(torch.empty((len(hosts),) + tuple(h0.shape), dtype=h0.dtype, device=self.device).unbind(0))
        return [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in hosts]

    def _issue_prefetch(self) -> None:
        """H2D copy of the host tensors' current contents into the staging set (copy stream, asynchronous)."""
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(self._staged)                         # staging set has been consumed
            for sd, xh, se, eh in zip(self._sx, self.xs_host, self._se, self.eis_host):
                sd.copy_(xh, non_blocking=True)
                se.copy_(eh, non_blocking=True)
            self._h2d_ready.record(self._copy)

    def launch(self) -> None:
        """Enqueue one step (asynchronous)."""
        if self.prefetch:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._h2d_ready)
            torch._foreach_copy_(self._xs, self._sx)
            torch._foreach_copy_(self._eis, self._se)
            self._staged.record(cur)
            self.graph.replay()
            self._issue_prefetch()                                      # next step's inputs, under this step's kernels
        else:
            self.graph.replay()
        if self.after_replay is not None:
            self.after_replay()

    def release(self) -> None:
        """Drop the captured graph and its memory pool (call before tearing down process groups)."""
        torch.cuda.synchronize(self.device)
        self.graph.reset()

    def wait_prefetch(self) -> None:
        """Block until the H2D prefetch issued by the last ``launch()`` has read the pinned host tensors.  They belong
        to ``GraphedStep`` from ``launch()`` until this returns; refill them (step k+2's data) only afterwards."""
        if self.prefetch:
            self._h2d_ready.synchronize()

    def __call__(self) -> float:
        """One step; returns the loss.  On return the pinned host tensors are free to be overwritten."""
        self.launch()
        torch.cuda.current_stream().synchronize()
        self.wait_prefetch()
        return float(self.loss_host)
