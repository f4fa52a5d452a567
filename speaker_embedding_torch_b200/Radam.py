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
        """One fused step over ALL parameter groups.  The clip norm is global (every parameter of every group, as
        ``clip_grad_norm_(model.parameters())`` at Train.py:154-159) and deterministic; tensors are launched in chunks
        of 64 that share one (group, step count), so per-parameter step counts (a parameter whose gradient was None on
        some steps) get their own bias correction, as in the reference's per-tensor loop (Radam.py:31-88)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        chunks, dev = [], None
        for group in self.param_groups:
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                N.require_cuda(p, "parameter")
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("fused optimiser needs contiguous fp32 parameters and gradients")
                dev = dev or p.device
                if p.device != dev:
                    raise RuntimeError("fused optimiser: all parameters must live on one device")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                buckets.setdefault(st["step"] + 1, []).append(p)
            for step, plist in buckets.items():
                for start in range(0, len(plist), 64):
                    chunks.append((group, step, plist[start:start + 64]))
        if not chunks:
            return loss
        if len(chunks) > N.OPTIM_MAX_CHUNKS:
            raise RuntimeError("fused optimiser: %d launch chunks exceed the %d the norm scratch holds"
                               % (len(chunks), N.OPTIM_MAX_CHUNKS))
        need = 1 + 296 * N.OPTIM_MAX_CHUNKS
        if self._scratch is None or self._scratch.device != dev or self._scratch.numel() < need:
            self._scratch = torch.zeros(need, dtype=torch.float32, device=dev)
        tables = []
        for group, step, plist in chunks:
            tab = N.OptimTensors()
            tab.count = len(plist)
            for i, p in enumerate(plist):
                st = self.state[p]
                tab.param[i], tab.grad[i] = p.data_ptr(), p.grad.data_ptr()
                tab.exp_avg[i], tab.exp_avg_sq[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                tab.numel[i] = p.numel()
            tables.append(tab)

        def launch(phase, c):
            group, step, _ = chunks[c]
            b1, b2 = group["betas"]
            N.check(N.lib().spk_optim_step(ctypes.byref(tables[c]), self.KIND, int(step), float(group["lr"]),
                                           float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                           self.max_grad_norm, float(grad_scale), N.ptr(self._scratch), phase, c,
                                           len(chunks), N.stream_ptr(dev)), "spk_optim_step")

        with torch.cuda.device(dev):
            if len(chunks) == 1:
                launch(0, 0)
            else:
                if self.max_grad_norm > 0:
                    for c in range(len(chunks)):
                        launch(1, c)          # norm partials of every chunk first: the clip is global
                for c in range(len(chunks)):
                    launch(2, c)
        for _, _, plist in chunks:            # only after every launch was accepted
            for p in plist:
                self.state[p]["step"] += 1
            # the kernels wrote the parameters through raw pointers: tell torch (version counters are what autograd's
            # saved-tensor checks and the inference weight cache of Modules.py look at)
            torch.autograd.graph.increment_version(plist)
        return loss

    def grad_norm(self):
        """Global gradient norm measured by the last fused step (device tensor; no sync).  Only measured when
        ``max_grad_norm > 0`` (the norm kernel is skipped otherwise)."""
        if self.max_grad_norm <= 0 or self._scratch is None:
            raise RuntimeError("grad_norm() is measured by the fused clip: construct the optimiser with max_grad_norm > 0")
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
