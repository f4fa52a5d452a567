"""Fused optimiser step on B200: ``RAdam`` with the reference's constructor (Radam.py:15-19) plus
``FusedAdamW``; both can fold the ``clip_grad_norm_`` of Train.py:154-159 into the same two launches.

The reference's ``RAdam.step`` is a Python loop of ~10 small kernels per tensor (Radam.py:31-88);
here one kernel computes the global gradient norm and one applies clip + update to every tensor
(``spk_optim_step``).  State layout (``exp_avg``, ``exp_avg_sq``, ``step`` per parameter) matches
the reference so optimizer ``state_dict``s are interchangeable.
"""
import ctypes

import torch
from torch.optim.optimizer import Optimizer

from . import _native as N


class _FusedBase(Optimizer):
    KIND = 0

    def __init__(self, params, lr, betas, eps, weight_decay, max_grad_norm=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)   # > 0 folds clip_grad_norm_(max_norm) into the step
        self._scratch = None

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            dev = plist[0].device
            if self._scratch is None or self._scratch.device != dev:
                self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
            step = None
            for start in range(0, len(plist), 64):
                chunk = plist[start:start + 64]
                if len(plist) > 64 and self.max_grad_norm > 0:
                    raise RuntimeError("fused clip supports at most 64 tensors per group")
                tab = N.OptimTensors()
                tab.count = len(chunk)
                for i, p in enumerate(chunk):
                    N.require_cuda(p, "parameter")
                    if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                        raise RuntimeError("fused optimiser needs contiguous fp32 parameters and gradients")
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = 0
                        st["exp_avg"] = torch.zeros_like(p)
                        st["exp_avg_sq"] = torch.zeros_like(p)
                    st["step"] += 1
                    step = st["step"]
                    tab.param[i], tab.grad[i] = p.data_ptr(), p.grad.data_ptr()
                    tab.exp_avg[i], tab.exp_avg_sq[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    tab.numel[i] = p.numel()
                b1, b2 = group["betas"]
                with torch.cuda.device(dev):
                    N.check(N.lib().spk_optim_step(ctypes.byref(tab), self.KIND, int(step), float(group["lr"]),
                                                   float(b1), float(b2), float(group["eps"]),
                                                   float(group["weight_decay"]), self.max_grad_norm,
                                                   float(grad_scale), N.ptr(self._scratch), N.stream_ptr(dev)),
                            "spk_optim_step")
        return loss

    def grad_norm(self):
        """Global gradient norm measured by the last fused step (device tensor; no sync)."""
        return self._scratch[0].sqrt()


class RAdam(_FusedBase):
    """Rectified Adam, same signature and update rule as the reference (Radam.py:15-90)."""
    KIND = 0

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, max_grad_norm=0.0):
        super().__init__(params, lr, betas, eps, weight_decay, max_grad_norm)


class FusedAdamW(_FusedBase):
    """torch.optim.AdamW semantics (Train.py:122-127 at HEAD), fused."""
    KIND = 1

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        super().__init__(params, lr, betas, eps, weight_decay, max_grad_norm)
