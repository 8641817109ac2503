"""TitanAdamW for B200: the reference's interface (titan.py:8-296) on top of the same one-launch update kernel.

The reference's Titan offloads every gradient to a CPU fp32 buffer from a post-accumulate-grad hook
(titan.py:93-131) so that a 6 GB card can train.  On a 180 GB B200 the offload target is HBM: the hook moves
the gradient into a persistent fp32 device buffer (accumulating across micro-steps, as titan.py:121-128 does on
the CPU) and clears ``p.grad`` so that ``p.grad is None`` after backward, exactly like the reference.
``clip_grad_norm`` (titan.py:162-184) works on those fp32 buffers (fp32 norm and coefficient, ``+1e-6``), and
``step`` runs the shared Raven/Titan math (titan.py:270-291 == raven.py:126-143) in one multi-tensor launch.
"""
from __future__ import annotations

import weakref

import torch

from .. import _lib
from .raven import RavenAdamW, _DT, _ptr_table


class TitanAdamW(RavenAdamW):
    _name = "TitanAdamW"

    def __init__(self, params, lr: float = 1e-4, betas: tuple[float, float] = (0.9, 0.999), weight_decay: float = 0.01,
                 eps: float = 1e-8, debias_strength: float = 1.0, momentum_dtype: torch.dtype = torch.bfloat16):
        super().__init__(params, lr=lr, betas=betas, weight_decay=weight_decay, eps=eps,
                         debias_strength=debias_strength, momentum_dtype=momentum_dtype)
        self._dev_grads = {}            # parameter -> persistent fp32 gradient buffer in HBM (the reference's CPU buffer)
        self._grad_ready = set()
        self._hook_handles = []
        self._closed = False
        self._pending_clip = None
        me = weakref.ref(self)          # hooks and ownership marks must not keep a discarded optimizer alive
        for p in (q for group in self.param_groups for q in group["params"] if q.requires_grad):
            self._claim(p, me)

    def _claim(self, p, me):
        """Take ownership of one trainable parameter: ``p._titan_optimizer_owner`` (a weak reference, the attribute the
        reference uses, titan.py:62-100) marks it so that two live Titan optimizers never offload the same gradient."""
        register = getattr(p, "register_post_accumulate_grad_hook", None)
        if register is None:
            raise RuntimeError("TitanAdamW needs Tensor.register_post_accumulate_grad_hook (PyTorch >= 2.0)")
        mark = getattr(p, "_titan_optimizer_owner", None)
        other = mark() if callable(mark) else None
        if other is not None and other is not self:
            self.close()
            raise RuntimeError("parameter already belongs to a live TitanAdamW: close() that optimizer before building another")
        self._dev_grads[p] = torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
        p._titan_optimizer_owner = me

        def after_accumulate(param, me=me):
            live = me()
            if live is not None:
                live._offload_gradient(param)

        self._hook_handles.append(register(after_accumulate))

    def _offload_gradient(self, param):
        if param.grad is None:
            return
        buf = self._dev_grads[param]
        if param in self._grad_ready:
            buf.add_(param.grad.detach().to(torch.float32))
        else:
            buf.copy_(param.grad.detach())
            self._grad_ready.add(param)
        param.grad = None

    def _grad_of(self, p):
        if p in self._grad_ready:
            return self._dev_grads[p]
        return None if p.grad is None else p.grad.float()

    def close(self):
        """Detach from the parameters: hooks removed, ownership marks cleared (only ours), gradient buffers freed
        (what titan.py:133-146 promises; idempotent, also called from ``__del__``)."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        while self._hook_handles:
            self._hook_handles.pop().remove()
        for p in list(self._dev_grads):
            mark = getattr(p, "_titan_optimizer_owner", None)
            if callable(mark) and mark() is self:
                del p._titan_optimizer_owner
        self._dev_grads.clear()
        self._grad_ready.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def zero_grad(self, set_to_none: bool = True):
        self._pending_clip = None
        if set_to_none:
            self._grad_ready.clear()
        else:
            for p in self._grad_ready:
                self._dev_grads[p].zero_()
        super().zero_grad(set_to_none)

    @torch.no_grad()
    def clip_grad_norm(self, max_norm, norm_type=2.0):
        """titan.py:162-184: norm over the offloaded fp32 gradients; scale in place when the coefficient is < 1."""
        params = [p for group in self.param_groups for p in group["params"] if p in self._grad_ready]
        if not params:
            return torch.tensor(0.0)
        if float(norm_type) != 2.0:
            raise _lib.AozoraError("TitanAdamW.clip_grad_norm: only norm_type=2 is implemented on the B200 path")
        items = [(None, p, self._dev_grads[p]) for p in params]
        out = self._gradnorm(items, max_norm if max_norm > 0 else 3.0e38, emulate_bf16=False)
        total_norm = out[0].clone()
        if max_norm > 0:
            # coefficient = min(1, max_norm / (norm + 1e-6)); it is applied to the fp32 gradients inside the next
            # step() launch (bit-identical to scaling the buffers in place first, and saves one pass over them)
            self._pending_clip = out[1:2].clone()
        return total_norm

    @torch.no_grad()
    def step(self, closure=None, clip_coef=None):
        if clip_coef is None and self._pending_clip is not None:
            clip_coef = self._pending_clip
        self._pending_clip = None
        return super().step(closure, clip_coef=clip_coef)
