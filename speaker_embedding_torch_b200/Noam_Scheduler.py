"""LR schedules with the reference's interface (Noam_Scheduler.py:5-29): host-side scalar math."""
from torch.optim.lr_scheduler import _LRScheduler


class Noam_Scheduler(_LRScheduler):
    def __init__(self, optimizer, warmup_steps):
        self.warmup_steps = warmup_steps
        super().__init__(optimizer)

    def get_lr(self):
        step = max(1, self.last_epoch)
        scale = self.warmup_steps ** 0.5 * min(step ** -0.5, step * self.warmup_steps ** -1.5)
        return [lr * scale for lr in self.base_lrs]


class Modified_Noam_Scheduler(_LRScheduler):
    """No-warm-up variant: lr = base_lr * sqrt(base / (step + base))."""

    def __init__(self, optimizer, base):
        self.base = base
        super().__init__(optimizer)

    def get_lr(self):
        step = max(1, self.last_epoch)
        scale = self.base ** 0.5 * (step + self.base) ** -0.5
        return [lr * scale for lr in self.base_lrs]
