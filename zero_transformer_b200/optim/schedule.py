"""Learning-rate schedule: linear warmup from 0 to ``peak`` over
``warmup_steps`` then cosine decay to ``end_value`` at ``decay_steps``
(counted from step 0, warmup included), constant afterwards.

Same curve as the schedule the reference builds in ``main_zero.py:207-213``;
the reference hard-codes ``decay_steps=143000`` -- here it is
``training.decay_steps`` (default 143000).  ``step`` is the number of optimizer
updates already applied, so the very first update uses ``lr(0) == 0`` exactly
like an optax schedule driven by ``ScaleByScheduleState.count``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass(frozen=True)
class WarmupCosine:
    peak_value: float
    warmup_steps: int
    decay_steps: int = 143000
    end_value: float = 0.0
    init_value: float = 0.0

    def __call__(self, step: int) -> float:
        step = max(int(step), 0)
        if self.warmup_steps > 0 and step < self.warmup_steps:
            frac = step / float(self.warmup_steps)
            return self.init_value + frac * (self.peak_value - self.init_value)
        span = max(self.decay_steps - self.warmup_steps, 1)
        t = min(step - self.warmup_steps, span) / float(span)
        cos = 0.5 * (1.0 + math.cos(math.pi * t))
        return self.end_value + (self.peak_value - self.end_value) * cos


def schedule_from_config(training) -> WarmupCosine:
    return WarmupCosine(
        peak_value=float(training.peak_learning_rate),
        warmup_steps=int(training.warmup_steps),
        decay_steps=int(training.get("decay_steps", 143000)),
        end_value=float(training.end_learning_rate),
    )
