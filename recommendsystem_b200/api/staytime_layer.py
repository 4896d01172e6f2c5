"""Drop-ins for staytime/layer.py: DIN (variant B, softmax attention), DeepCrossLayer, FMLayer."""
from __future__ import annotations

import torch
from torch import nn

from .. import cabi
from .din import _glorot
from .functional import CrossFn, DinFn


class DIN(nn.Module):
    """staytime/layer.py:6-41 — call(query [B,H], facts [B,T,H'], mask bool [B,T]) -> [B,H]:
    sigmoid Dense(16) -> Dense(1), masked positions -> -2**32+1, softmax over T, weights @ facts.
    `facts` may be a column slice of a wider sequence embedding (e.g. seq[:, :, 0:16],
    staytime/VideoDnn.py:68): it is read in place through its row stride."""

    def __init__(self, **kwargs):
        super().__init__()
        self.built = False

    def build(self, H, device):
        self.layer_1_kernel = nn.Parameter(_glorot(4 * H, 16, device))
        self.layer_1_bias = nn.Parameter(torch.zeros(16, device=device))
        self.layer_2_kernel = nn.Parameter(_glorot(16, 1, device))
        self.layer_2_bias = nn.Parameter(torch.zeros(1, device=device))
        self.built = True

    def forward(self, query, facts, mask):
        if not self.built:
            self.build(query.shape[-1], query.device)
        m = None
        if mask is not None:
            m = mask[:, :facts.shape[1]].to(torch.uint8).contiguous()
        return DinFn.apply(cabi.DIN_B, query, facts, None, None, m, self.layer_1_kernel, self.layer_1_bias,
                           self.layer_2_kernel, self.layer_2_bias)


class DeepCrossLayer(nn.Module):
    """staytime/layer.py:44-80 — DCN cross layers: cross <- inputs * (cross @ w_i) + b_i + cross.
    HBM-bound: rs_cross_fwd / rs_cross_bwd read each row once forward, twice backward."""

    def __init__(self, num_layer=3, **kwargs):
        super().__init__()
        self.num_layer = num_layer
        self.built = False

    def build(self, dim, device):
        self.W = nn.ParameterList([nn.Parameter(_glorot(dim, 1, device)) for _ in range(self.num_layer)])   # w_i
        self.b = nn.ParameterList([nn.Parameter(torch.zeros(dim, device=device)) for _ in range(self.num_layer)])
        self.built = True

    def forward(self, inputs):
        if not self.built:
            self.build(inputs.shape[-1], inputs.device)
        # cross <- inputs * (cross @ w_i) + b_i + cross (:66-72) for all layers in ONE fused row kernel (csrc/cross.cu)
        W = torch.stack([w.reshape(-1) for w in self.W])
        b = torch.stack([v.reshape(-1) for v in self.b])
        return CrossFn.apply(inputs, W, b)


class FMLayer(nn.Module):
    """staytime/layer.py:83-116 — FM second-order term over [B, fields, k]: 0.5 * ((sum e)^2 - sum e^2)."""

    def forward(self, inputs):
        if inputs.dim() != 3:
            raise ValueError("Unexpected inputs dimensions %d, expect to be 3 dimensions" % inputs.dim())
        s = inputs.sum(dim=1, keepdim=True)
        cross = s * s - (inputs * inputs).sum(dim=1, keepdim=True)
        return 0.5 * cross.sum(dim=2)
