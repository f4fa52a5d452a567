"""Learning-rate schedules behind the reference's two class names (Noam_Scheduler.py:5-29).

Both schedules are a pure function of the step count, so they share one scheduler class that multiplies every
group's base learning rate by ``factor(step)``; the factor functions are module-level so the fused optimiser step
(``Radam.RAdam`` / ``csrc/optim.cu``) and the CPU oracle (``oracle.ge2e_oracle.modified_noam_lr``) can be checked
against exactly the same closed forms:

    noam_factor(s, W)          = sqrt(W) * min(1/sqrt(s), s / W**1.5)      warm-up W steps, then 1/sqrt decay
    modified_noam_factor(s, B) = sqrt(B / (s + B))                         no warm-up, 1/sqrt decay with offset B

with ``s = max(1, steps taken so far)`` as in the reference.
"""
import math

from torch.optim.lr_scheduler import _LRScheduler


def noam_factor(step: int, warmup_steps: float) -> float:
    s = float(max(1, step))
    rise = s / (warmup_steps ** 1.5)
    decay = 1.0 / math.sqrt(s)
    return math.sqrt(warmup_steps) * (rise if rise < decay else decay)


def modified_noam_factor(step: int, base: float) -> float:
    s = float(max(1, step))
    return math.sqrt(base / (s + base))


class _Closed_Form_Scheduler(_LRScheduler):
    """``lr_g(step) = base_lr_g * factor(step)`` for every parameter group g."""

    def _factor(self, step):
        raise NotImplementedError

    def get_lr(self):
        k = self._factor(self.last_epoch)
        return [k * group_lr for group_lr in self.base_lrs]


class Noam_Scheduler(_Closed_Form_Scheduler):
    def __init__(self, optimizer, warmup_steps):
        self.warmup_steps = warmup_steps          # attribute name kept: it is part of the scheduler's state_dict
        super().__init__(optimizer)

    def _factor(self, step):
        return noam_factor(step, self.warmup_steps)


class Modified_Noam_Scheduler(_Closed_Form_Scheduler):
    def __init__(self, optimizer, base):
        self.base = base                          # state_dict key of the reference's checkpoints
        super().__init__(optimizer)

    def _factor(self, step):
        return modified_noam_factor(step, self.base)
