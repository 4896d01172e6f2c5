"""oracle_torch.py — second, independent CPU restatement (torch, op-for-op, unfused)
of the reference's TensorFlow graph for the CTR hot path, with autograd.

TEST INFRASTRUCTURE ONLY (see oracle/oracle_np.py header; PARITY UNPINNED for the
same reasons).  Two uses: (1) cross-check of oracle_np (two restatements written
from the same reference lines must agree before either is trusted); (2) the CPU
baseline timed by bench.py (`cpu_baseline`, `--impl reference`): TensorFlow is not
installable offline, so the reference's TF CPU path is represented by this
op-for-op restatement — every intermediate is materialised like the TF graph does
(tile / concat / split / matmul / softmax), nothing is fused.

Never imported by recommendsystem_b200/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


def layer_norm(x, gamma, beta, eps):
    mean = x.mean(-1, keepdim=True)
    var = (x - mean).pow(2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def interacting_layer(x, Wq, bq, Wk, bk, Wv, bv, Wr, br, gamma, beta, ln_eps, head_num, layer_num,
                      use_res=True, dropout=None):
    """InteractingLayer.call, InteractingLayer.py:37-61, written with the same
    split/concat ops as the reference."""
    if x.dim() != 3:
        raise ValueError("The rank of input of InteractingLayer must be 3, but now is %d" % x.dim())
    output = x
    for it in range(layer_num):                                           # :41
        query = torch.relu(output @ Wq + bq)                              # :42
        key = torch.relu(output @ Wk + bk)                                # :43
        value = torch.relu(output @ Wv + bv)                              # :44
        if use_res:
            res = torch.relu(output @ Wr + br)                            # :45-46
        query = torch.cat(torch.chunk(query, head_num, dim=2), dim=0)     # :47
        key = torch.cat(torch.chunk(key, head_num, dim=2), dim=0)         # :48
        value = torch.cat(torch.chunk(value, head_num, dim=2), dim=0)     # :49
        weight = torch.matmul(query, key.transpose(1, 2))                 # :50
        weight = weight / (key.shape[-1] ** 0.5)                          # :51
        weight = torch.softmax(weight, dim=-1)                            # :52
        if dropout is not None:                                           # :53-54 (training); dropout = (rate, seed):
            from . import oracle_np as onp                                # the kernels' counter-based mask, [H,B,F,F]
            B_, F_ = x.shape[0], x.shape[1]
            sc = onp.dropout_scale(dropout[1], it, B_, head_num, F_, dropout[0]).reshape(head_num * B_, F_, F_)
            weight = weight * torch.from_numpy(sc).to(weight.dtype)
        output = torch.matmul(weight, value)                              # :55
        output = torch.cat(torch.chunk(output, head_num, dim=0), dim=2)   # :56
        if use_res:
            output = output + res                                         # :57-58
        output = torch.relu(output)                                       # :59
        output = layer_norm(output, gamma, beta, ln_eps)                  # :60
    return output


def din_a(queries, keys, values, seq_length, W1, b1, W2, b2):
    """din.py:18-47."""
    queries = queries.unsqueeze(1)                                        # :19
    from_seq_len = queries.shape[1]
    to_seq_len = keys.shape[1]
    masks = torch.arange(to_seq_len)[None, :] < seq_length[:, None]       # tf.sequence_mask :24
    queries = queries.unsqueeze(2).repeat(1, 1, to_seq_len, 1)            # :26,28
    keys4 = keys.unsqueeze(1).repeat(1, from_seq_len, 1, 1)               # :27,29
    deep = torch.cat([queries, keys4, queries * keys4], dim=-1)           # :31
    deep = torch.relu(deep @ W1 + b1)                                     # :33-34
    deep = torch.relu(deep @ W2 + b2)
    deep = deep.squeeze(-1)                                               # :37
    masks = masks.unsqueeze(1).repeat(1, from_seq_len, 1)                 # :40-41
    deep = torch.where(masks, deep, torch.zeros_like(deep))               # :42
    output = torch.matmul(deep, values)                                   # :44
    return output.squeeze(1)                                              # :45


def din_b(query, facts, mask, W1, b1, W2, b2):
    """staytime/layer.py:16-41."""
    seq_len = facts.shape[1]
    queries = query.repeat(1, seq_len).reshape(facts.shape)               # :20-21
    din_all = torch.cat([queries, facts, queries - facts, queries * facts], dim=-1)  # :22-23
    d1 = torch.sigmoid(din_all @ W1 + b1)                                 # :24
    d2 = d1 @ W2 + b2                                                     # :25
    scores = d2.reshape(-1, 1, seq_len)                                   # :26
    if mask is not None:
        key_masks = mask[:, :seq_len].bool().unsqueeze(1)                 # :30-31
        paddings = torch.ones_like(scores) * (-2 ** 32 + 1)               # :32
        scores = torch.where(key_masks, scores, paddings)                 # :34
    scores = torch.softmax(scores, dim=-1)                                # :35
    output = torch.matmul(scores, facts)                                  # :36
    return output.squeeze(1)                                              # :39


def multi_layer_dense(x, weights, biases, activation):
    """MultiLayerDense (autoint:40-41,49-50) / DNN.call (rough_rank/layer.py:100-109)."""
    for W, b in zip(weights, biases):
        x = x @ W + b
        if activation == "relu":
            x = torch.relu(x)
        elif activation == "sigmoid":
            x = torch.sigmoid(x)
    return x


def cross_entropy(y_true, y_pred, a=1.0):
    """rank/ctr/base_model.py:7-12."""
    l = -y_true * torch.log(y_pred + 1e-6) - (a - y_true) * torch.log(1 - y_pred + 1e-6)
    return torch.mean(torch.sum(l, dim=1))


class AutoIntCPU:
    """Whole AutoInt train step on the CPU (embedding lookup -> InteractingLayer ||
    MLP -> logits -> clip -> BCE -> backward -> Adam on dense params and touched
    rows), autoint:18-56 + rank/ctr/base_model.py:160-201.  Used as the CPU
    baseline ("CPU restatement of reference graph (TensorFlow unavailable offline)")."""

    def __init__(self, table, params, head_num, layer_num, ln_eps, lr_dense=5e-5, lr_sparse=5e-5,
                 beta1=0.9, beta2=0.999, eps=1e-8):
        self.table = table                      # [R_total, d] fp32
        self.m = torch.zeros_like(table)
        self.v = torch.zeros_like(table)
        self.params = {k: (v.clone().requires_grad_(True)) for k, v in params.items()}
        self.H, self.L, self.ln_eps = head_num, layer_num, ln_eps
        self.dense_opt = torch.optim.Adam(list(self.params.values()), lr=lr_dense, betas=(beta1, beta2), eps=eps)
        self.lr_sparse, self.b1, self.b2, self.eps = lr_sparse, beta1, beta2, eps
        self.step_no = 0

    def forward(self, X):
        P = self.params
        B = X.shape[0]
        U = P["Wqkvr"].shape[1] // 4
        W, b = P["Wqkvr"], P["bqkvr"]
        A = interacting_layer(X, W[:, :U], b[:U], W[:, U:2 * U], b[U:2 * U], W[:, 2 * U:3 * U],
                              b[2 * U:3 * U], W[:, 3 * U:], b[3 * U:], P["gamma"], P["beta"],
                              self.ln_eps, self.H, self.L)
        A = A.reshape(B, -1)
        n_mlp = len([k for k in P if k.startswith("mlp_W")])
        deep = multi_layer_dense(X.reshape(B, -1), [P[f"mlp_W{i}"] for i in range(n_mlp)],
                                 [P[f"mlp_b{i}"] for i in range(n_mlp)], "relu")
        Z = torch.cat([deep, A], dim=1)
        out = torch.sigmoid(Z @ P["out_W"] + P["out_b"])
        return torch.clamp(out, 1e-6, 1.0)

    def train_step(self, rowidx, y):
        """rowidx int64 [B,F] arena rows, y [B,1]."""
        self.step_no += 1
        X = self.table[rowidx].detach().requires_grad_(True)   # embedding lookup
        p = self.forward(X)
        loss = cross_entropy(y, p)
        self.dense_opt.zero_grad(set_to_none=True)
        loss.backward()
        self.dense_opt.step()
        # sparse Adam on touched rows (sum duplicate lookups first)
        flat = rowidx.reshape(-1)
        uniq, inv = torch.unique(flat, return_inverse=True)
        g = torch.zeros(len(uniq), X.shape[-1], dtype=X.dtype).index_add_(0, inv, X.grad.reshape(len(flat), -1))
        t = self.step_no
        corr = (1 - self.b2 ** t) ** 0.5 / (1 - self.b1 ** t)
        m = self.b1 * self.m[uniq] + (1 - self.b1) * g
        v = self.b2 * self.v[uniq] + (1 - self.b2) * g * g
        self.m[uniq] = m
        self.v[uniq] = v
        self.table[uniq] -= self.lr_sparse * corr * m / (v.sqrt() + self.eps)
        return float(loss)
