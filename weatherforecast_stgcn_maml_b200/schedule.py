"""Host-side training-loop logic of the meta-trainer: the outer learning-rate schedule, the
adaptive task sampler and the task-difficulty update of train_hybrid_maml_v5.py:245-294.  Plain scalar logic -- it
drives ``MetaTrainer.set_lr`` / the task subset of a meta-step; nothing here touches the device."""
from __future__ import annotations

import math

import numpy as np


class CosineWarmRestarts:
    """``torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(T_0, T_mult, eta_min)`` as the reference configures it
    (T_0=10, T_mult=2, eta_min=1e-6, stepped once per epoch: train_hybrid_maml_v5.py:250-252,294), for an optimiser
    that is not a ``torch.optim.Optimizer`` (the fused AdamW keeps its state in flat device buffers)."""

    def __init__(self, base_lr, T_0=10, T_mult=2, eta_min=1e-6):
        if T_0 <= 0 or T_mult < 1:
            raise ValueError("T_0 must be positive and T_mult >= 1")
        self.base_lr, self.T_0, self.T_mult, self.eta_min = float(base_lr), int(T_0), int(T_mult), float(eta_min)
        self.T_i, self.T_cur, self.last_epoch = int(T_0), 0, 0
        self.lr = float(base_lr)

    def step(self):
        """Advance one epoch and return the new learning rate (same recurrence as torch's ``step()`` without an epoch)."""
        self.last_epoch += 1
        self.T_cur += 1
        if self.T_cur >= self.T_i:
            self.T_cur -= self.T_i
            self.T_i *= self.T_mult
        self.lr = self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.T_cur / self.T_i)) / 2
        return self.lr

    def get_last_lr(self):
        return [self.lr]


class AdaptiveTaskSampler:
    """The reference's "adaptive task sampling" (train_hybrid_maml_v5.py:264-292): draw ``batch_size`` of the tasks
    without replacement, with probabilities proportional to a per-task difficulty that is an EMA of the META loss --
    every task receives the same value, so the draw degenerates to uniform; reproduced as is,
    including the use of numpy's global RNG (``np.random.seed(42)`` at train_hybrid_maml_v5.py:22)."""

    def __init__(self, num_tasks, batch_size):
        self.num_tasks, self.batch_size = int(num_tasks), int(batch_size)
        self.task_losses = []

    def sample(self):
        n, b = self.num_tasks, self.batch_size
        if n > b and self.task_losses:
            total = sum(self.task_losses)
            probs = np.array(self.task_losses) / total if total > 0 else None
            return [int(i) for i in np.random.choice(n, b, replace=False, p=probs)]
        if n > b:
            return [int(i) for i in np.random.choice(n, b, replace=False)]
        return list(range(n))

    def update(self, loss):
        loss = float(loss)
        if len(self.task_losses) < self.num_tasks:
            self.task_losses.extend([loss] * (self.num_tasks - len(self.task_losses)))
        else:
            self.task_losses = [0.9 * t + 0.1 * loss for t in self.task_losses]
