"""Fused AdamW for the SVOL head (train.py:71-78: ``torch.optim.AdamW(lr=1e-4, weight_decay=1e-4)``).

All parameters of the module are re-pointed into ONE flat fp32 buffer (same order and 4-element alignment as the
training engine's flat gradient buffer), so an optimizer step is a single ``svol_adamw`` launch and the data-parallel
gradient exchange is a single NCCL all-reduce over ``TrainEngine.grad_flat`` instead of one
bucket per parameter.  ``step()`` takes the gradients from ``param.grad`` like any torch optimizer;
``step(from_engine=True)`` reads the engine's flat gradient buffer directly (valid when the step's only backward was
the CUDA head backward -- the normal training loop), skipping the per-parameter copies.
"""
from __future__ import annotations

import torch

from . import _lib


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FusedAdamW:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.params = list(model.parameters())
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs the parameters on a CUDA device; there is no CPU fallback")
        total = sum(_round_up(p.numel(), 4) for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self._views, off = [], 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                                   # parameters now live in the flat buffer
                self._views.append((off, p.numel()))
                off += _round_up(p.numel(), 4)
        self.step_count = 0
        self._invalidate()

    def _invalidate(self):
        """The CUDA head packs bf16 copies of its weights; in-place updates do not bump tensor versions."""
        for mod in self.model.modules():
            eng = getattr(mod, "_engine", None)
            if eng is not None:
                eng._wstate = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, from_engine: bool = False, grad_scale: float = 1.0):
        if from_engine:
            eng = None
            for mod in self.model.modules():
                eng = getattr(mod, "_train_engine", None) or eng
            if eng is None or eng.grad_flat is None or eng.grad_flat.numel() != self.flat.numel():
                raise RuntimeError("step(from_engine=True) needs the model's TrainEngine gradients (run a training backward first)")
            g = eng.grad_flat
        else:
            g = self.grad
            g.zero_()
            for p, (off, n) in zip(self.params, self._views):
                if p.grad is not None:
                    g[off:off + n].view_as(p).copy_(p.grad)
        self.step_count += 1
        P = _lib.ptr
        _lib.check(_lib.get_lib().svol_adamw(P(self.flat), P(g), P(self.m), P(self.v), self.flat.numel(), self.lr, self.betas[0],
                                             self.betas[1], self.eps, self.weight_decay, self.step_count, grad_scale,
                                             _lib.stream_ptr()), "adamw")
        self._invalidate()
