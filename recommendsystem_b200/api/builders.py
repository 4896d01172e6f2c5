"""Model builders of the reference's CTR path, wired from the drop-in layers.

    AutoInt(model_config).run()                     autoint:11-60 (+ rank/ctr/base_model.py output_layer)
    AUTOINT(linear_slots, bucket_size, ...)         rank/multi_head/multidnn.py:14-259 (AutoInt + DNN +
                                                    7-expert / 7-gate MMoE, 7 sigmoid heads)
    cross_entropy                                   rank/multi_head/model.py:18-22, rank/ctr/base_model.py:7-12

AutoInt is backed by the captured single-launch-stream engine (recommendsystem_b200.autoint); AUTOINT
is an nn.Module over the autograd wrappers (every Dense / InteractingLayer forward+backward is a C-ABI
kernel) whose sparse side is api.embedding.EmbeddingFeatures.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
from torch import nn

from ..autoint import AutoIntConfig, AutoIntTrainer
from .embedding import Adam, EmbeddingFeatures, category_column, embedding_column
from .functional import dense
from .interacting_layer import InteractingLayer


def cross_entropy(y_true, y_pred, a=1.0, reduce_mean=False):
    """-y log(p+1e-6) - (a-y) log(1-p+1e-6), summed over the label axis (keepdims);
    rank/multi_head/model.py:18-22.  reduce_mean=True: mean over the batch (rank/ctr/base_model.py:7-12)."""
    l = -y_true * torch.log(y_pred + 1e-6) - (a - y_true) * torch.log(1 - y_pred + 1e-6)
    l = l.sum(dim=1, keepdim=True)
    return l.mean() if reduce_mean else l


class AutoInt:
    """AutoInt(model_config) — autoint:11-60.  model_config['model_param'] follows the keys the reference
    reads (:30-41,49-50): interact{layer_num,unit_num,head_num,use_dropout,dropout_rate,use_res},
    mlp{hidden_units,activation}, logits{hidden_units,activation}; plus 'feature' {num_fields,
    rows_per_field, embed_dim} and optional 'batch', 'dtype', 'ln_eps' for the synthetic setting (the
    reference's JSON slot bookkeeping, rank/ctr/base_model.py:35-158, is a "next" row in SURVEY §8f)."""

    def __init__(self, model_config: Dict, device="cuda:0", **kwargs):
        self.model_config = model_config
        mp, ft = model_config["model_param"], model_config["feature"]
        it = mp["interact"]
        if mp["mlp"].get("activation", "relu") != "relu":
            raise NotImplementedError("the fused AutoInt engine implements the relu tower")
        lg = mp.get("logits", {"hidden_units": [1], "activation": "sigmoid"})
        if list(lg["hidden_units"]) != [1] or lg["activation"] != "sigmoid":
            raise NotImplementedError("logits head: Dense(1, sigmoid) (autoint:49-52 clips to a probability)")
        if it.get("use_dropout", False):
            raise NotImplementedError("the captured AutoInt engine runs without attention dropout; use "
                                      "api.InteractingLayer(use_dropout=True) in a module graph")
        self.cfg = AutoIntConfig(num_fields=ft["num_fields"], rows_per_field=ft["rows_per_field"],
                                 embed_dim=ft["embed_dim"], layer_num=it["layer_num"], unit_num=it["unit_num"],
                                 head_num=it["head_num"], use_res=it.get("use_res", True),
                                 ln_eps=model_config.get("ln_eps", 1e-3), mlp_hidden=tuple(mp["mlp"]["hidden_units"]),
                                 batch=model_config.get("batch", 8192), dtype=model_config.get("dtype", "f32"))
        self.device = device

    def run(self):
        trainer = AutoIntTrainer(self.cfg, self.device)
        return {"train": trainer, "predict": trainer.predict}


class _KerasDense(nn.Module):
    def __init__(self, units, activation=None, init_std=None):
        super().__init__()
        self.units, self.activation, self.init_std = units, activation, init_std
        self.kernel = None

    def forward(self, x):
        if self.kernel is None:
            fan_in = x.shape[-1]
            w = torch.empty(fan_in, self.units, device=x.device)
            if self.init_std is None:
                lim = (6.0 / (fan_in + self.units)) ** 0.5
                w.uniform_(-lim, lim)
            else:
                nn.init.trunc_normal_(w, std=self.init_std, a=-2 * self.init_std, b=2 * self.init_std)
            self.kernel = nn.Parameter(w)
            self.bias = nn.Parameter(torch.zeros(self.units, device=x.device))
        return dense(x, self.kernel, self.bias, self.activation)


class AutoIntSubModel(nn.Module):
    """create_autoint_sub_model (rank/multi_head/multidnn.py:14-212): InteractingLayer(1, 8, 2 heads) ||
    Dense stack -> concat -> 7 (of 8 built) experts Dense(32, relu) -> 7 softmax gates -> 7 sigmoid heads."""

    NUM_LABELS = 7

    def __init__(self, deep_hidden_units: Sequence[int] = (32, 16), expert_num=7, expert_units=32):
        super().__init__()
        # multidnn.py:54: attention dropout 0.2 in training (fused into the kernel, counter-based mask)
        self.interact = InteractingLayer(layer_num=1, unit_num=8, head_num=2, use_dropout=True, dropout_rate=0.2,
                                         use_res=True)
        self.deep = nn.ModuleList([_KerasDense(u, "relu") for u in deep_hidden_units])
        self.experts = nn.ModuleList([_KerasDense(expert_units, "relu", 0.001) for _ in range(expert_num + 1)])
        self.gates = nn.ModuleList([_KerasDense(expert_num, "softmax", 0.001) for _ in range(self.NUM_LABELS)])
        self.heads = nn.ModuleList([_KerasDense(1, "sigmoid") for _ in range(self.NUM_LABELS)])
        self.expert_num = expert_num

    def forward(self, embs: Sequence[torch.Tensor]):
        all_inputs = torch.stack(list(embs), dim=1)                       # :25-27,50  [B,F,8]
        autoint = self.interact(all_inputs).flatten(1)                    # :54-56
        deep = all_inputs.flatten(1)                                      # :60
        for layer in self.deep:
            deep = layer(deep)                                            # :62-63
        result = torch.cat([deep, autoint], dim=1)                        # :72
        experts = torch.stack([e(result) for e in self.experts][: self.expert_num], dim=1)   # :80-92
        preds = []
        for g, h in zip(self.gates, self.heads):
            gate = g(result).unsqueeze(-1)                                # :97-104
            preds.append(h((experts * gate).sum(dim=1)))                  # :106-116, heads :118-206
        return torch.cat(preds, dim=1)                                    # [B,7]


class AUTOINT:
    """AUTOINT(linear_features, ...) — rank/multi_head/multidnn.py:214-259: 8-d embeddings (combiner mean,
    sparse Adam lr 5e-5) feeding AutoIntSubModel.  `linear_features` is a list of slot names;
    `bucket_size` replaces the missing src.* Config.  train_step() runs forward, summed BCE over the 7
    labels, backward, dense Adam (lr 1e-5, rank/multi_head/model.py:53) and the sparse push."""

    def __init__(self, linear_features: Sequence[str], bucket_size=100_000, dnn_hidden_units=(32, 16),
                 device="cuda:0", seed=0):
        cols = [embedding_column(category_column(s, bucket_size), dimension=8, combiner="mean") for s in linear_features]
        self.slots = list(linear_features)
        self.emb = EmbeddingFeatures(cols, Adam(5e-5, 0.9, 0.999, 1e-8), "linear", device=device, seed=seed)
        self.sub_model = AutoIntSubModel(dnn_hidden_units).to(device)
        self.opt = None

    def predict(self, inputs: Dict[str, torch.Tensor]):
        with torch.no_grad():
            embs = self.emb(inputs)
            return self.sub_model([embs[s] for s in self.slots])

    def train_step(self, inputs: Dict[str, torch.Tensor], labels: torch.Tensor):
        embs = self.emb(inputs)
        leaves = [embs[s].detach().requires_grad_(True) for s in self.slots]
        pred = self.sub_model(leaves)
        if self.opt is None:
            self.opt = torch.optim.Adam(self.sub_model.parameters(), lr=1e-5, betas=(0.9, 0.999), eps=1e-8, capturable=True)
        loss = cross_entropy(labels, pred).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.emb.backward({s: l.grad for s, l in zip(self.slots, leaves)})
        return loss.detach(), pred.detach()
