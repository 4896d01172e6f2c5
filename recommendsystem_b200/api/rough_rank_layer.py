"""Drop-ins for rough_rank/layer.py: DNN, MMOE, PLE, CrossNet, Similarity, KDLoss.  The Dense
contractions (the part that costs time) run through the C-ABI GEMM kernels via DenseFn — tcgen05
for bf16 activations, FFMA for fp32; the small gate arithmetic (softmax over <= 8 experts,
gate-weighted sum, mat-vec cross terms) is memory-bound glue composed from torch CUDA ops."""
from __future__ import annotations

import math

import torch
from torch import nn

from .functional import CrossFn, dense


def _glorot_normal(fan_in, fan_out, device, gen=None):
    # tf.keras.initializers.GlorotNormal: truncated normal, stddev = sqrt(2 / (fan_in + fan_out))
    std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    w = torch.empty(fan_in, fan_out, device=device)
    nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return w


class Similarity(nn.Module):
    """rough_rank/layer.py:6-30 — row-wise dot product of (user_emb, item_emb), optional sigmoid."""

    def __init__(self, use_sigmoid=False, **kwargs):
        super().__init__()
        self.use_sigmoid = use_sigmoid

    def forward(self, inputs):
        user_emb, item_emb = inputs
        out = (user_emb * item_emb).sum(dim=-1, keepdim=True)
        return torch.sigmoid(out) if self.use_sigmoid else out

    def get_config(self):
        return {"use_sigmoid": self.use_sigmoid}


class DNN(nn.Module):
    """rough_rank/layer.py:33-117 — x <- act_i(x @ kernel_i + bias_i) [+ dropout]; GlorotNormal kernels
    (`kernel{i}`), zero biases (`bias{i}`); all activations `activation` except the last, which is
    `output_activation` when given (:88-96).  `l2_reg` is exposed as `regularization_loss()`."""

    def __init__(self, hidden_units, activation="relu", l2_reg=0, dropout_rate=0, use_bn=False,
                 output_activation=None, seed=None, **kwargs):
        super().__init__()
        self.hidden_units = list(hidden_units)
        self.activation, self.l2_reg, self.dropout_rate = activation, l2_reg, dropout_rate
        self.use_dropout = dropout_rate > 0
        self.use_bn, self.output_activation, self.seed = use_bn, output_activation, seed
        if use_bn:
            raise NotImplementedError("use_bn=True (BatchNormalization) is not built; no reference call site uses it")
        self.built = False

    def build(self, input_size, device):
        units = [int(input_size)] + self.hidden_units
        gen = None
        if self.seed is not None:
            gen = torch.Generator(device=device).manual_seed(int(self.seed))
        self.kernels = nn.ParameterList([nn.Parameter(_glorot_normal(units[i], units[i + 1], device, gen))
                                         for i in range(len(self.hidden_units))])
        self.bias = nn.ParameterList([nn.Parameter(torch.zeros(units[i + 1], device=device))
                                      for i in range(len(self.hidden_units))])
        n = len(self.hidden_units)
        self.acts = [self.activation] * n
        if n and self.output_activation:
            self.acts[-1] = self.output_activation
        self.built = True

    def forward(self, inputs, training=None):
        if not self.built:
            self.build(inputs.shape[-1], inputs.device)
        training = self.training if training is None else training
        x = inputs
        for i in range(len(self.hidden_units)):
            x = dense(x, self.kernels[i], self.bias[i], self.acts[i])        # :101-105
            if self.use_dropout and training:
                x = torch.nn.functional.dropout(x, self.dropout_rate, True)   # :106-107
        return x

    def regularization_loss(self):
        return self.l2_reg * sum((k * k).sum() for k in self.kernels) if self.built and self.l2_reg else 0.0

    def get_config(self):
        return {"activation": self.activation, "hidden_units": self.hidden_units, "l2_reg": self.l2_reg,
                "use_bn": self.use_bn, "dropout_rate": self.dropout_rate,
                "output_activation": self.output_activation, "seed": self.seed}


def _mix(expert_outs, gate_out):
    """sum_e gate[b,e] * expert[b,e,:]  (tf.stack / expand_dims / multiply / reduce_sum)."""
    return (torch.stack(expert_outs, dim=-2) * gate_out.unsqueeze(-1)).sum(dim=-2)


class MMOE(nn.Module):
    """rough_rank/layer.py:120-171."""

    def __init__(self, num_tasks, num_experts=2, expert_dnn_units=(32,), gate_dnn_units=(), expert_dnn_params=None,
                 gate_dnn_params=None, **kwargs):
        super().__init__()
        self.num_tasks, self.num_experts = num_tasks, num_experts
        self.expert_dnn_units = expert_dnn_units
        self.gate_dnn_units = list(gate_dnn_units) + [num_experts]
        self.expert_dnn_params = dict(expert_dnn_params or {})
        self.gate_dnn_params = {"output_activation": "softmax", **(gate_dnn_params or {})}
        self.expert_nets = nn.ModuleList([DNN(expert_dnn_units, **self.expert_dnn_params) for _ in range(num_experts)])
        self.gate_nets = nn.ModuleList([DNN(self.gate_dnn_units, **self.gate_dnn_params) for _ in range(num_tasks)])

    def forward(self, inputs, training=None):
        expert_outs = [net(inputs, training=training) for net in self.expert_nets]
        return [_mix(expert_outs, self.gate_nets[i](inputs, training=training)) for i in range(self.num_tasks)]

    def get_config(self):
        return {"num_tasks": self.num_tasks, "num_experts": self.num_experts,
                "expert_dnn_units": self.expert_dnn_units, "gate_dnn_units": self.gate_dnn_units,
                "expert_dnn_params": self.expert_dnn_params, "gate_dnn_params": self.gate_dnn_params}


class PLE(nn.Module):
    """rough_rank/layer.py:174-233 — shared + task-specific experts, one softmax gate per task."""

    def __init__(self, num_tasks, num_shared_experts=2, num_specific_experts=2, expert_dnn_units=(32,),
                 gate_dnn_units=(), expert_dnn_params=None, gate_dnn_params=None, **kwargs):
        super().__init__()
        self.num_tasks = num_tasks
        self.num_shared_experts, self.num_specific_experts = num_shared_experts, num_specific_experts
        self.expert_dnn_units = expert_dnn_units
        self.gate_dnn_units = list(gate_dnn_units) + [num_shared_experts + num_specific_experts]
        self.expert_dnn_params = dict(expert_dnn_params or {})
        self.gate_dnn_params = {"output_activation": "softmax", **(gate_dnn_params or {})}
        mk = lambda: DNN(expert_dnn_units, **self.expert_dnn_params)
        self.shared_expert_nets = nn.ModuleList([mk() for _ in range(num_shared_experts)])
        self.specific_expert_nets = nn.ModuleList([nn.ModuleList([mk() for _ in range(num_specific_experts)])
                                                   for _ in range(num_tasks)])
        self.gate_nets = nn.ModuleList([DNN(self.gate_dnn_units, **self.gate_dnn_params) for _ in range(num_tasks)])

    def forward(self, inputs, training=None):
        shared = [net(inputs, training=training) for net in self.shared_expert_nets]
        outs = []
        for i in range(self.num_tasks):
            specific = [net(inputs, training=training) for net in self.specific_expert_nets[i]]
            outs.append(_mix(shared + specific, self.gate_nets[i](inputs, training=training)))
        return outs


class CrossNet(nn.Module):
    """rough_rank/layer.py:236-270 — DCN-v1: x_{l+1} = x0 (x_l . w_l) + b_l + x_l."""

    def __init__(self, layer_num=2, l2_reg=0, seed=1024, **kwargs):
        super().__init__()
        self.layer_num, self.l2_reg, self.seed = layer_num, l2_reg, seed
        self.built = False

    def build(self, dim, device):
        gen = torch.Generator(device=device).manual_seed(int(self.seed))
        self.kernels = nn.ParameterList([nn.Parameter(_glorot_normal(dim, 1, device, gen)) for _ in range(self.layer_num)])
        self.bias = nn.ParameterList([nn.Parameter(torch.zeros(dim, 1, device=device)) for _ in range(self.layer_num)])
        self.built = True

    def forward(self, inputs):
        if not self.built:
            self.build(inputs.shape[-1], inputs.device)
        # x_{l+1} = x0 (x_l . kernel_l) + bias_l + x_l for all layers in ONE fused row kernel (csrc/cross.cu)
        W = torch.stack([k.reshape(-1) for k in self.kernels])
        b = torch.stack([v.reshape(-1) for v in self.bias])
        return CrossFn.apply(inputs, W, b)

    def get_config(self):
        return {"layer_num": self.layer_num, "l2_reg": self.l2_reg, "seed": self.seed}


class KDLoss(nn.Module):
    """rough_rank/layer.py:272-279 — per-sample MSE (reduction NONE) between teacher and student."""

    def forward(self, student_predictions, teacher_predictions):
        return ((teacher_predictions - student_predictions) ** 2).mean(dim=-1)
