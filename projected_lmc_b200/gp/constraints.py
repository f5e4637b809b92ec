"""Parameter constraints with gpytorch's transform (softplus) and buffer names, so
``state_dict()`` keys match the reference's (``raw_*_constraint.lower_bound`` ...)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def inv_softplus(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(-torch.expm1(-x))


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound)))


class GreaterThan(Interval):
    """value = softplus(raw) + lower_bound."""

    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)

    def transform(self, raw: torch.Tensor) -> torch.Tensor:
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value: torch.Tensor) -> torch.Tensor:
        return inv_softplus(value - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
