"""Learning-rate schedule of the reference trainer: ``LinearWarmupCosineAnnealingLR`` stepped once per EPOCH
(train.py:71-84: warmup_epochs = 10 % of max_epochs, eta_min = 1e-6, interval 'epoch'; utils/schedulers.py:239-346).

Closed form of the reference's chainable recursion (utils/schedulers.py:328-346 gives the same closed form): linear from
``warmup_start_lr`` to ``base_lr`` over ``warmup_epochs - 1`` steps — so, stepped per epoch, the FIRST epoch trains at
``warmup_start_lr`` = 0, exactly as the reference warns (utils/schedulers.py:242-245) — then half a cosine down to
``eta_min`` at ``max_epochs``.  Host arithmetic only: the trainer reads ``trainer.lr`` at every step.
"""
from __future__ import annotations

import math


def warmup_cosine_lr(epoch: int, base_lr: float, warmup_epochs: int, max_epochs: int, warmup_start_lr: float = 0.0,
                     eta_min: float = 0.0) -> float:
    """lr used during 0-based ``epoch`` (== scheduler.last_epoch after ``epoch`` calls of ``scheduler.step()``)."""
    if warmup_epochs == 0:
        # max_epochs < 10 with train.py's int(0.1 * epochs): the reference's chainable recursion (utils/schedulers.py:
        # 303-326) returns warmup_start_lr at epoch 0 and never passes through its `last_epoch == warmup_epochs` reset,
        # so the cosine factor is applied to (lr - eta_min) starting from warmup_start_lr — the lr stays within
        # eta_min of 0 for the whole run.  Reproduced literally, quirk included.
        lr = warmup_start_lr
        for e in range(1, epoch + 1):
            if (e - 1 - max_epochs) % (2 * max_epochs) == 0:
                lr = lr + (base_lr - eta_min) * (1.0 - math.cos(math.pi / max_epochs)) / 2.0
            else:
                lr = ((1.0 + math.cos(math.pi * e / max_epochs)) / (1.0 + math.cos(math.pi * (e - 1) / max_epochs))
                      * (lr - eta_min) + eta_min)
        return lr
    if epoch < warmup_epochs:
        if warmup_epochs <= 1:
            return warmup_start_lr
        return warmup_start_lr + epoch * (base_lr - warmup_start_lr) / (warmup_epochs - 1)
    return eta_min + 0.5 * (base_lr - eta_min) * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / (max_epochs - warmup_epochs)))


class EpochSchedule:
    """``configure_optimizers`` of the reference (train.py:68-84) for the CUDA trainer:

        sched = EpochSchedule(trainer, base_lr=opt.lr, max_epochs=opt.epochs)
        for epoch in range(opt.epochs):
            sched.begin_epoch(epoch)          # sets trainer.lr
            for batch in loader: trainer.train_step(...)
    """

    def __init__(self, trainer, base_lr: float, max_epochs: int, warmup_epochs: int | None = None, eta_min: float = 1e-6):
        self.trainer, self.base_lr, self.max_epochs = trainer, base_lr, max_epochs
        self.warmup_epochs = int(0.1 * max_epochs) if warmup_epochs is None else warmup_epochs
        self.eta_min = eta_min

    def lr_at(self, epoch: int) -> float:
        return warmup_cosine_lr(epoch, self.base_lr, self.warmup_epochs, self.max_epochs, 0.0, self.eta_min)

    def begin_epoch(self, epoch: int) -> float:
        self.trainer.lr = self.lr_at(epoch)
        return self.trainer.lr
