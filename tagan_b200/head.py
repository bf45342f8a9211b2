"""After the hot path: ragged-snapshot pad/stack, node pooling, the classification head
with its loss, and the optimizer step -- each ONE launch of libtagan_b200, so a whole training step (forward, loss,
backward, gradient clipping, Adam) is sync-free and captures into a single CUDA graph.

``ClassificationModule`` mirrors the reference module tree so its ``state_dict`` keys match
(``classification_head.classification_head.{attention.0,attention.2,classifier.0,classifier.1,classifier.4}.*``,
reference src/tagan/layers/classification.py:743-975, 1069-1231)."""
import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .ops import CALLS, _ptr, _stream, workspace

_F32 = torch.float32


# ------------------------------------------------------------------------------------------
# ragged snapshots: packed rows <-> zero-padded stack
# ------------------------------------------------------------------------------------------
class _PackPaddedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, offsets, t_steps: int, maxn: int):
        lib = _lib.load()
        packed = packed.contiguous().float()
        h = packed.shape[1]
        out = torch.empty(t_steps, maxn, h, dtype=_F32, device=packed.device)
        _lib.check(lib.tagan_pack_padded_fwd(_ptr(packed), _ptr(offsets), _ptr(out), t_steps, maxn, h, _stream()),
                   "tagan_pack_padded_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(offsets)
        ctx.cfg = (packed.shape[0], t_steps, maxn, h)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (offsets,) = ctx.saved_tensors
        rows, t_steps, maxn, h = ctx.cfg
        dpacked = torch.empty(rows, h, dtype=_F32, device=dout.device)
        _lib.check(lib.tagan_pack_padded_bwd(_ptr(dout.contiguous()), _ptr(offsets), _ptr(dpacked), t_steps, maxn, h, _stream()),
                   "tagan_pack_padded_bwd")
        CALLS["n"] += 1
        return dpacked, None, None, None


def pack_padded(packed: torch.Tensor, offsets: torch.Tensor, t_steps: int, maxn: int) -> torch.Tensor:
    """packed ``[sum N_t, H]`` + device ``offsets[T+1]`` (int32) -> ``[T, maxn, H]``, rows beyond N_t zero: the pad + stack
    of the reference's list input (temporal_attention.py:928-976) in one launch, no per-snapshot ``F.pad`` / ``stack``."""
    return _PackPaddedFn.apply(packed, offsets, t_steps, maxn)


# ------------------------------------------------------------------------------------------
# node pooling of TAGAN.forward
# ------------------------------------------------------------------------------------------
class _PoolBlocksFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_phys, batch: int, t_steps: int, time_major: bool):
        lib = _lib.load()
        x_phys = x_phys.contiguous().float()
        h = x_phys.shape[-1]
        out = torch.empty(t_steps, h, dtype=_F32, device=x_phys.device)
        ws = workspace(lib.tagan_pool_blocks_workspace_bytes(t_steps, h), x_phys.device)
        _lib.check(lib.tagan_pool_blocks_fwd(_ptr(x_phys), batch, t_steps, h, int(time_major), _ptr(out), _ptr(ws), ws.numel(),
                                             _stream()), "tagan_pool_blocks_fwd")
        CALLS["n"] += 2
        ctx.cfg = (batch, t_steps, h, time_major, x_phys.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        batch, t_steps, h, time_major, shape = ctx.cfg
        dx = torch.empty(shape, dtype=_F32, device=dout.device)
        _lib.check(lib.tagan_pool_blocks_bwd(_ptr(dout.contiguous()), batch, t_steps, h, int(time_major), _ptr(dx), _stream()),
                   "tagan_pool_blocks_bwd")
        CALLS["n"] += 1
        return dx, None, None, None


def pool_blocks(x_phys: torch.Tensor, batch: int, t_steps: int, time_major: bool) -> torch.Tensor:
    """``graph_features[t]`` of TAGAN.forward (model.py:377-427) from the temporal attention output ``x[B,T,H]``
    (``time_major``: its storage is ``[T,B,H]``): the mean of the t-th block of B consecutive rows of ``x.view(B*T, H)``."""
    return _PoolBlocksFn.apply(x_phys, batch, t_steps, time_major)


# ------------------------------------------------------------------------------------------
# classification head + loss
# ------------------------------------------------------------------------------------------
def _head_struct(tensors) -> _lib.HeadWeights:
    hw = _lib.HeadWeights()
    for name, t in zip(("attn0_weight", "attn0_bias", "attn2_weight", "fc0_weight", "fc0_bias", "ln_weight", "ln_bias",
                        "fc1_weight", "fc1_bias"), tensors):
        setattr(hw, name, None if t is None else t.data_ptr())
    return hw


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gf, wa1, ba1, wa2, w1, b1, lng, lnb, w2, b2, labels, class_index, loss_type: int):
        lib = _lib.load()
        gf = gf.contiguous().float()
        bsz, t, h = gf.shape
        o = w2.shape[0]
        dev = gf.device
        e = lambda *s: torch.empty(*s, dtype=_F32, device=dev)  # noqa: E731
        u, alpha, pooled, h1, hn, stats, logits = e(bsz, t, h), e(bsz, t), e(bsz, h), e(bsz, h), e(bsz, h), e(bsz, 2), e(bsz, o)
        want_loss = labels is not None or class_index is not None
        loss = e(()) if want_loss else None
        lab = labels.contiguous().float() if labels is not None else None
        rows = lab.shape[0] if lab is not None else 0
        weights = [t_.contiguous() if t_ is not None else None for t_ in (wa1, ba1, wa2, w1, b1, lng, lnb, w2, b2)]
        hw = _head_struct(weights)
        rc = lib.tagan_head_fwd(C.byref(hw), _ptr(gf), bsz, t, h, o, loss_type, _ptr(lab), rows, _ptr(class_index), _ptr(u), _ptr(alpha),
                                _ptr(pooled), _ptr(h1), _ptr(hn), _ptr(stats), _ptr(logits), _ptr(loss), _stream())
        _lib.check(rc, "tagan_head_fwd")
        CALLS["n"] += 1
        ctx.save_for_backward(gf, u, alpha, pooled, h1, hn, stats, logits, lab, class_index, *[w for w in weights if w is not None])
        ctx.has_ln, ctx.cfg = lng is not None, (bsz, t, h, o, loss_type, rows)
        if want_loss:
            return logits, loss
        return logits, None

    @staticmethod
    def backward(ctx, dlogits, dloss):
        lib = _lib.load()
        saved = ctx.saved_tensors
        gf, u, alpha, pooled, h1, hn, stats, logits, lab, class_index = saved[:10]
        ws_ = list(saved[10:])
        if not ctx.has_ln:
            ws_ = ws_[:5] + [None, None] + ws_[5:]
        bsz, t, h, o, loss_type, rows = ctx.cfg
        dev = gf.device
        hw = _head_struct(ws_)
        grads = [torch.empty_like(w) if w is not None else None for w in ws_]
        dw = _head_struct(grads)
        dgf = torch.empty_like(gf)
        ws = workspace(lib.tagan_head_bwd_workspace_bytes(bsz, t, h, o), dev)
        dl = dloss.contiguous().float() if dloss is not None else None
        dlg = dlogits.contiguous().float() if dlogits is not None else None
        rc = lib.tagan_head_bwd(C.byref(hw), _ptr(gf), bsz, t, h, o, loss_type, _ptr(lab), rows, _ptr(class_index), _ptr(u), _ptr(alpha),
                                _ptr(pooled), _ptr(h1), _ptr(hn), _ptr(stats), _ptr(logits), _ptr(dl), _ptr(dlg), _ptr(dgf), C.byref(dw),
                                _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "tagan_head_bwd")
        CALLS["n"] += 1
        return (dgf, *grads, None, None, None)


class TemporalClassificationHead(nn.Module):
    """Mirror of the reference ``TemporalClassificationHead`` in the configuration ``TAGAN`` builds (attention pooling,
    two layers, ReLU; classification.py:743-975).  ``forward(x[Bsz,T,H])`` -> logits ``[Bsz,O]``; ``forward_loss`` also
    returns the loss TAGAN.forward computes (BCE with logits, or cross entropy for class-index labels)."""

    def __init__(self, hidden_dim: int, num_classes: int, dropout: float = 0.1, use_layer_norm: bool = True):
        super().__init__()
        self.hidden_dim, self.num_classes, self.dropout, self.use_layer_norm = hidden_dim, num_classes, dropout, use_layer_norm
        self.attention = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, 1, bias=False))
        layers = [nn.Linear(hidden_dim, hidden_dim)]
        if use_layer_norm:
            layers.append(nn.LayerNorm(hidden_dim))
        layers += [nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden_dim, num_classes)]
        self.classifier = nn.Sequential(*layers)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _weights(self):
        ln = self.classifier[1] if self.use_layer_norm else None
        return (self.attention[0].weight, self.attention[0].bias, self.attention[2].weight.reshape(-1), self.classifier[0].weight,
                self.classifier[0].bias, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                self.classifier[-1].weight, self.classifier[-1].bias)

    def forward_loss(self, x, labels=None, class_index=None):
        if self.training and self.dropout > 0:
            raise NotImplementedError("the fused head has no dropout: use eval() or dropout=0 (the parity setting)")
        loss_type = 1 if class_index is not None else 0
        return _HeadFn.apply(x, *self._weights(), labels, class_index, loss_type)

    def forward(self, x, mask=None, labels=None):
        if mask is not None:
            raise NotImplementedError("temporal mask in the classification head (TAGAN.forward never passes one)")
        return self.forward_loss(x)[0]


class ClassificationModule(nn.Module):
    """Mirror of the reference ``ClassificationModule`` (single task; classification.py:1069-1231)."""

    def __init__(self, hidden_dim: int, output_dim: int, dropout: float = 0.1, use_layer_norm: bool = True):
        super().__init__()
        self.classification_head = TemporalClassificationHead(hidden_dim, output_dim, dropout, use_layer_norm)

    def forward(self, x, mask=None, labels=None):
        return self.classification_head(x, mask, labels)


# ------------------------------------------------------------------------------------------
# optimizer step
# ------------------------------------------------------------------------------------------
class FusedAdam:
    """``torch.optim.Adam`` + ``clip_grad_norm_`` (reference trainer.py:295-311) as three launches on FLAT buffers.

    The parameters are re-pointed to views of one flat fp32 buffer and their ``.grad`` to views of a flat gradient buffer
    (autograd then accumulates in place), so the gradient norm is one reduction, the update one kernel, and a data-parallel
    all-reduce needs no packing (``flat_grad`` is the bucket).  The step counter lives on the device: replaying a captured
    graph advances it.  Use ``opt.zero_grad()`` (NOT ``model.zero_grad(set_to_none=True)``, which would drop the views).
    Parameters that never receive a gradient hold a zero gradient: with ``weight_decay = 0`` their update is exactly zero,
    as for torch's ``grad is None``; with ``weight_decay > 0`` they decay, where torch would skip them."""

    def __init__(self, params: Sequence[nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no parameters"
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("tagan_b200 has no CPU path: FusedAdam needs CUDA parameters")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)                   # every view 16-byte aligned
        self.flat = torch.zeros(n, dtype=_F32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=_F32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=_F32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=_F32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grad_msq = torch.zeros((), dtype=_F32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                view = self.flat[off:off + k].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += (k + 3) // 4 * 4

    def zero_grad(self):
        self.flat_grad.zero_()                       # keeps the views: backward accumulates into the flat buffer

    def step(self):
        lib = _lib.load()
        n = self.flat.numel()
        if self.max_grad_norm > 0:
            ws = workspace(lib.tagan_mse_workspace_bytes(), self.flat.device)
            _lib.check(lib.tagan_mse_fwd(_ptr(self.flat_grad), n, _ptr(self.grad_msq), _ptr(ws), ws.numel(), _stream()), "tagan_mse_fwd")
            CALLS["n"] += 2
        rc = lib.tagan_adam_clip_step(_ptr(self.flat), _ptr(self.flat_grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq), n, self.lr,
                                      self.betas[0], self.betas[1], self.eps, self.weight_decay, self.max_grad_norm,
                                      _ptr(self.grad_msq) if self.max_grad_norm > 0 else None, _ptr(self.step_dev), _stream())
        _lib.check(rc, "tagan_adam_clip_step")
        CALLS["n"] += 2

    def grad_norm(self) -> float:
        """Total gradient norm seen by the last ``step()`` (host sync)."""
        return float(torch.sqrt(self.grad_msq * self.flat.numel()))


class TrainStep:
    """One training step of the reference trainer (trainer.py:295-311: forward with labels, ``zero_grad``, backward,
    ``clip_grad_norm_``, ``Adam.step``) on a device-resident ``PackedSequence``, optionally captured ONCE as a CUDA graph and
    replayed: ``TrainStep(model, opt).capture(seq, labels)`` then ``loss = step.replay()`` after refilling ``seq`` /
    ``labels`` in place (``step.seq.x.copy_(...)``).  Sync-free: the only host read is the caller's ``float(loss)``."""

    def __init__(self, model: nn.Module, opt: FusedAdam):
        self.model, self.opt = model, opt
        self.graph = None
        self.seq = self.labels = self.loss = None

    def eager(self, seq, labels) -> torch.Tensor:
        self.opt.zero_grad()
        loss = self.model(seq, labels)["loss"]
        loss.backward()
        self.opt.step()
        return loss.detach()

    def capture(self, seq, labels, warmup: int = 2):
        """Warm up on a side stream (lazy library initialisation must not happen under capture), restore the parameters
        and optimizer state touched by the warm-up, then capture."""
        self.seq, self.labels = seq, labels
        opt = self.opt
        snap = [t.clone() for t in (opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_dev)]
        side = torch.cuda.Stream(device=seq.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.eager(seq, labels)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(seq.x.device)
        for dst, src in zip((opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_dev), snap):
            dst.copy_(src)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self.eager(seq, labels)
        return self

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
