"""Parameter holders: carry the reference's state_dict names/shapes/init, never compute."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class ConvParams(nn.Module):
    """Parameters of a Conv2d (weight [co, ci/groups, k, k], optional bias) with torch's default Conv2d init."""

    def __init__(self, c_in: int, c_out: int, k: int, bias: bool, groups: int = 1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c_out, c_in // groups, k, k))
        self.bias = nn.Parameter(torch.empty(c_out)) if bias else None
        fan_in = (c_in // groups) * k * k
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, *a, **k):  # pragma: no cover - holders are not callable layers
        raise RuntimeError("parameter holder: the forward runs in the fused CUDA path of the owning model")


class AffineParams(nn.Module):
    """weight (ones) and optional bias (zeros) of shape [n] (LayerNorm body)."""

    def __init__(self, n: int, bias: bool):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        if bias:
            self.bias = nn.Parameter(torch.zeros(n))


class Holder(nn.Module):
    """Plain namespace module."""


def ordered_tensors(module: nn.Module):
    """Tensors in state_dict() order (parameters and buffers interleaved as registered)."""
    return list(module.state_dict(keep_vars=True).values())
