"""DIN — drop-in for the reference's din.py (variant A): local-activation unit with ReLU scores,
zero masking from `seq_length` and weighted-sum pooling (din.py:6-47)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import cabi
from .functional import DinFn


def _glorot(fan_in, fan_out, device):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return torch.empty(fan_in, fan_out, device=device).uniform_(-lim, lim)


class DIN(nn.Module):
    """DIN(**kwargs); call(queries [B,H], keys [B,T,H], values [B,T,H], seq_length [B]) -> [B,H].
    Dense layers `din_nn_0` (3H -> 16, relu) and `din_nn_1` (16 -> 1, relu) as in din.py:12-16."""

    def __init__(self, **kwargs):
        super().__init__()
        self.built = False

    def build(self, H, device):
        self.din_nn_0_kernel = nn.Parameter(_glorot(3 * H, 16, device))
        self.din_nn_0_bias = nn.Parameter(torch.zeros(16, device=device))
        self.din_nn_1_kernel = nn.Parameter(_glorot(16, 1, device))
        self.din_nn_1_bias = nn.Parameter(torch.zeros(1, device=device))
        self.built = True

    def forward(self, queries, keys, values, seq_length=None):
        if seq_length is None:
            # tf.sequence_mask(None) fails in the reference too (din.py:24)
            raise ValueError("DIN needs seq_length (tf.sequence_mask, din.py:24)")
        if not self.built:
            self.build(queries.shape[-1], queries.device)
        return DinFn.apply(cabi.DIN_A, queries, keys, values, seq_length.to(torch.int32).contiguous(), None,
                           self.din_nn_0_kernel, self.din_nn_0_bias, self.din_nn_1_kernel, self.din_nn_1_bias)
