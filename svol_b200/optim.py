"""Fused AdamW for the SVOL head (train.py:71-96: ``torch.optim.AdamW(param_dicts, lr=args.lr, weight_decay=args.wd)``).

A ``torch.optim.Optimizer`` -- ``param_groups`` (the lr schedulers of train.py:129-137 rewrite ``group['lr']``; read at
every step), ``state_dict()`` / ``load_state_dict()`` in torch.optim.AdamW's own format (train.py:149,270 checkpoint
them; a checkpoint written by ``torch.optim.AdamW`` loads here and vice versa), ``zero_grad()``, ``add_param_group`` is
not supported after construction -- whose parameters are re-pointed into ONE flat fp32 buffer (4-element aligned
segments, the layout of the training engine's flat gradient buffer).  An optimizer step is then a single
``svol_adamw_segments`` launch, and the data-parallel gradient exchange a single NCCL all-reduce over
``TrainEngine.grad_flat`` instead of one bucket per parameter.

Like ``torch.optim.AdamW``, a parameter without a gradient is left untouched (no weight decay, no moment update):
``step()`` skips parameters whose ``.grad`` is None; ``step(from_engine=True)`` (gradients read straight from the engine's
flat buffer, valid when the step's only backward was the CUDA head backward) skips the parameters the head's forward
never reaches (``class_head.*``, the dead ``sketch_video_cross_attn.out_proj.*``: ``SVANet.params_without_grad()``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import _lib


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4):
        """``params``: an ``nn.Module`` (all its parameters, one group) or what torch optimizers take (an iterable of
        parameters or of ``{'params': [...], 'lr': ...}`` dicts)."""
        self.model = params if isinstance(params, torch.nn.Module) else None
        if self.model is not None:
            params = [p for p in self.model.parameters()]
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True)
        super().__init__(params, defaults)
        if len(self.param_groups) > 8:
            raise NotImplementedError("FusedAdamW supports up to 8 parameter groups")
        self.params: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        self._group_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs the parameters on a CUDA device; there is no CPU fallback")
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise NotImplementedError("FusedAdamW takes fp32 parameters on one device")
        total = sum(_round_up(p.numel(), 4) for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self._views, ends, off = [], [], 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                                   # parameters now live in the flat buffer
                self._views.append((off, p.numel()))
                off += _round_up(p.numel(), 4)
                ends.append(off)
        self._seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
        self._seg_group = torch.empty(len(ends), dtype=torch.int32, device=dev)
        self._active_key = None
        self._steps = [0] * len(self.params)                    # per-parameter step counts (torch keeps one per parameter)
        self._group_steps = [0] * len(self.param_groups)
        self._bind_state()
        self._engine_skip = None
        self._invalidate()

    # ------------------------------------------------------------------ torch.optim state
    def _bind_state(self):
        """``self.state[p]`` in torch.optim.AdamW's layout, exp_avg / exp_avg_sq being views of the flat moment buffers."""
        for p, (off, n), st in zip(self.params, self._views, self._steps):
            self.state[p] = {"step": torch.tensor(float(st)), "exp_avg": self.m[off:off + n].view_as(p),
                             "exp_avg_sq": self.v[off:off + n].view_as(p)}

    def state_dict(self):
        for p, st in zip(self.params, self._steps):
            self.state[p]["step"] = torch.tensor(float(st))
        sd = super().state_dict()
        # torch.optim.AdamW has no state entry for a parameter that never received a gradient
        never = {i for i, st in enumerate(self._steps) if st == 0}
        sd["state"] = {k: v for k, v in sd["state"].items() if k not in never}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                      # fills self.state[p] with copies, updates param_groups
        with torch.no_grad():
            for i, (p, (off, n)) in enumerate(zip(self.params, self._views)):
                st = self.state.get(p, {})
                if "exp_avg" in st:
                    self.m[off:off + n].view_as(p).copy_(st["exp_avg"])
                    self.v[off:off + n].view_as(p).copy_(st["exp_avg_sq"])
                    self._steps[i] = int(float(st["step"]))
                else:
                    self.m[off:off + n].zero_()
                    self.v[off:off + n].zero_()
                    self._steps[i] = 0
        for gi in range(len(self.param_groups)):
            self._group_steps[gi] = max([s for s, g in zip(self._steps, self._group_of) if g == gi] or [0])
        self._bind_state()
        self._invalidate()

    def add_param_group(self, param_group):
        if hasattr(self, "flat"):
            raise NotImplementedError("FusedAdamW lays its parameters out once, at construction")
        super().add_param_group(param_group)

    # ------------------------------------------------------------------ CUDA head coupling
    def _engines(self):
        mods = self.model.modules() if self.model is not None else []
        return [m for m in mods if getattr(m, "_engine", None) is not None]

    def _invalidate(self):
        """The CUDA head packs bf16 copies of its weights; updates through the flat buffer's raw pointer do not bump
        tensor versions, so the engines are told explicitly (HeadEngine repacks and bumps ``weights_generation``, which the
        training engine's transposed copies follow)."""
        for mod in self._engines():
            mod._engine._wstate = None

    # ------------------------------------------------------------------ step
    def _set_active(self, active: List[bool]):
        key = tuple(active)
        if key != self._active_key:
            seg = [g if a else -1 for g, a in zip(self._group_of, active)]
            self._seg_group.copy_(torch.tensor(seg, dtype=torch.int32), non_blocking=False)
            self._active_key = key

    @torch.no_grad()
    def step(self, closure=None, from_engine: bool = False, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if from_engine:
            heads = self._engines()
            eng = heads[-1]._train_engine if heads else None
            if eng is None or eng.grad_flat is None or eng.grad_flat.numel() != self.flat.numel() or \
                    len(eng._params) != len(self.params) or any(a is not b for a, b in zip(eng._params, self.params)):
                raise RuntimeError("step(from_engine=True) needs an optimizer built on exactly the head module's parameters "
                                   "and a training backward of that head (TrainEngine.grad_flat)")
            g = eng.grad_flat
            if self._engine_skip is None:
                skip = {id(p) for p in heads[-1].params_without_grad()}
                self._engine_skip = [id(p) not in skip for p in self.params]
            active = self._engine_skip
        else:
            g = self.grad
            active = [p.grad is not None for p in self.params]
            for p, (off, n), a in zip(self.params, self._views, active):
                if a:
                    g[off:off + n].view_as(p).copy_(p.grad)
        self._set_active(active)
        live_groups = {gi for gi, a in zip(self._group_of, active) if a}
        for gi in live_groups:
            self._group_steps[gi] += 1
        for i, a in enumerate(active):
            if a:
                self._steps[i] = self._group_steps[self._group_of[i]]
        groups = (_lib.AdamwGroup * len(self.param_groups))()
        for gi, grp in enumerate(self.param_groups):
            if grp.get("amsgrad") or grp.get("maximize"):
                raise NotImplementedError("FusedAdamW: amsgrad / maximize are not implemented")
            groups[gi].lr, groups[gi].beta1, groups[gi].beta2 = float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1])
            groups[gi].eps, groups[gi].weight_decay = float(grp["eps"]), float(grp["weight_decay"])
            groups[gi].step = max(self._group_steps[gi], 1)
        P = _lib.ptr
        _lib.check(_lib.get_lib().svol_adamw_segments(P(self.flat), P(g), P(self.m), P(self.v), self.flat.numel(),
                                                      P(self._seg_end), P(self._seg_group), len(self.params), groups,
                                                      len(self.param_groups), float(grad_scale), _lib.stream_ptr()),
                   "adamw_segments")
        self._invalidate()
        return loss

    @property
    def step_count(self) -> int:
        return max(self._group_steps)

    # kept for callers of the round-1 interface
    @property
    def lr(self) -> float:
        return self.param_groups[0]["lr"]
