"""TEST INFRASTRUCTURE — CPU restatement of the reference's two composed dense graphs, never imported
by the product package (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use oracle/).

    video_dnn_fwd(xp, ...)   create_moe_sub_model, staytime/VideoDnn.py:27-215 (+ staytime/config.py)
    dssm_fwd(xp, ...)        DSSM dense graph, rough_rank/model.py:16-86,130-170 (+ rough_rank/layer.py)

Each function is written once against a tiny array namespace `xp` so that the SAME statement of the
reference runs in numpy float64 (forward oracle) and in torch float64 with autograd (gradient oracle);
`NP` / `TH` below are the two namespaces.  Parameters are passed in a flat dict keyed by the reference's
Keras layer names (the state_dict keys of recommendsystem_b200.api.video_dnn / rough_rank_model).

Parity unpinned against TensorFlow itself (no TF / tensornet offline, the reference has no golden
vectors): see DESIGN.md §5.  The layer-level pieces (din_b, deep_cross, cross_net, _interacting) are pinned against
the reference's own layer code executed under a numpy stand-in of the TF ops (tests/test_oracle_reference_pin.py).
"""
import numpy as np

USER_SLOTS = ["1568", "1589", "2039", "1570"]
ITEM_SLOTS = ["1591", "1593", "1737", "1614"]
BIAS_SLOTS = ["3051", "1570", "2039", "2544", "1568", "3376", "3365", "3369", "2597", "1737", "1593", "1591",
              "1589", "1614"]
BIN_LIST = [-19.0 + 0.5 * i for i in range(400)]            # staytime/config.py bin_list
PAD = float(-2 ** 32 + 1)


class NP:
    """numpy float64 namespace"""
    @staticmethod
    def cat(xs, axis=-1): return np.concatenate(xs, axis=axis)
    @staticmethod
    def stack(xs, axis=0): return np.stack(xs, axis=axis)
    @staticmethod
    def relu(x): return np.maximum(x, 0)
    @staticmethod
    def sigmoid(x): return 1.0 / (1.0 + np.exp(-x))
    @staticmethod
    def softmax(x):
        e = np.exp(x - x.max(-1, keepdims=True))
        return e / e.sum(-1, keepdims=True)
    @staticmethod
    def where(c, a, b): return np.where(c, a, b)
    @staticmethod
    def detach(x): return x
    @staticmethod
    def const(v, like): return np.asarray(v, dtype=like.dtype)
    @staticmethod
    def sum(x, axis, keepdims=False): return x.sum(axis=axis, keepdims=keepdims)
    @staticmethod
    def mean(x, axis): return x.mean(axis=axis)
    @staticmethod
    def tile_rows(q, T): return np.broadcast_to(q[:, None, :], (q.shape[0], T, q.shape[1]))
    @staticmethod
    def expand(x, axis): return np.expand_dims(x, axis)


class TH:
    """torch float64 namespace (autograd)"""
    @staticmethod
    def cat(xs, axis=-1):
        import torch
        return torch.cat(list(xs), dim=axis)
    @staticmethod
    def stack(xs, axis=0):
        import torch
        return torch.stack(list(xs), dim=axis)
    @staticmethod
    def relu(x):
        import torch
        return torch.relu(x)
    @staticmethod
    def sigmoid(x):
        import torch
        return torch.sigmoid(x)
    @staticmethod
    def softmax(x):
        import torch
        return torch.softmax(x, dim=-1)
    @staticmethod
    def where(c, a, b):
        import torch
        return torch.where(c, a, b)
    @staticmethod
    def detach(x): return x.detach()
    @staticmethod
    def const(v, like):
        import torch
        return torch.as_tensor(v, dtype=like.dtype)
    @staticmethod
    def sum(x, axis, keepdims=False): return x.sum(dim=axis, keepdim=keepdims)
    @staticmethod
    def mean(x, axis): return x.mean(dim=axis)
    @staticmethod
    def tile_rows(q, T): return q[:, None, :].expand(q.shape[0], T, q.shape[1])
    @staticmethod
    def expand(x, axis): return x.unsqueeze(axis)


def _dense(xp, x, P, name, act=None):
    """tf.keras.layers.Dense: act(x @ kernel[in,out] + bias)."""
    y = x @ P[name + ".kernel"] + P[name + ".bias"]
    if act == "relu":
        return xp.relu(y)
    if act == "sigmoid":
        return xp.sigmoid(y)
    if act == "softmax":
        return xp.softmax(y)
    return y


def din_b(xp, query, facts, mask, P, name):
    """staytime/layer.py:16-41."""
    T = facts.shape[1]
    q = xp.tile_rows(query, T)                                                   # :18-21
    z = xp.cat([q, facts, q - facts, q * facts], -1)                             # :22-23
    h = xp.sigmoid(z @ P[name + ".layer_1_kernel"] + P[name + ".layer_1_bias"])  # :24
    s = (h @ P[name + ".layer_2_kernel"] + P[name + ".layer_2_bias"])[..., 0]    # :25-26
    s = xp.where(mask[:, :T], s, xp.const(PAD, s) * (s * 0 + 1))                 # :29-34
    p = xp.softmax(s)                                                            # :35
    return xp.sum(xp.expand(p, -1) * facts, 1)                                   # :36-39


def deep_cross(xp, x, P, name, num_layer=3):
    """staytime/layer.py:66-72: cross <- inputs * (cross @ w_i) + b_i + cross."""
    cross = x
    for i in range(num_layer):
        cross = x * (cross @ P["%s.W.%d" % (name, i)]) + P["%s.b.%d" % (name, i)] + cross
    return cross


def video_dnn_fwd(xp, embs, seqs, P, slots, seq_slots, units=(256, 128), num_experts=3, num_tasks=3,
                  task_names=("staytime_pred", "shortplay_pred", "longplay_pred")):
    """create_moe_sub_model (staytime/VideoDnn.py:27-215).  embs: slot -> [B,32]; seqs: slot -> ([B,T,32], mask)."""
    slots, seq_slots = sorted(slots), sorted(seq_slots)
    general = {s: embs[s][:, 0:16] for s in slots}                               # :47-48
    gi = [general[s] for s in slots]
    bias_inputs = [embs[s][:, 16:] for s in slots if s in BIAS_SLOTS]            # :45-46
    din_embs = []
    for s in seq_slots:                                                          # :66-77
        q = general["1591"] if s == "2125" else (general["1593"] if s == "2128" else general["1737"])
        seq, mask = seqs[s]
        din_embs.append(din_b(xp, q, seq[:, :, 0:16], mask, P, "din.din_%s" % s))
    n = len(slots)
    squeeze = xp.detach(xp.cat(gi, -1))                                          # :84-86
    w = 2 * _dense(xp, _dense(xp, squeeze, P, "senet_squeeze_layer1", "relu"), P, "senet_extract_layer2", "sigmoid")
    rw = [g * w[:, i:i + 1] for i, g in enumerate(gi)]                           # :93-96
    mult = xp.relu(xp.cat([general[s] for s in USER_SLOTS], -1) * xp.cat([general[s] for s in ITEM_SLOTS], -1))  # :98-105
    st = xp.stack(rw, 0)
    sum_embs = xp.sum(st, 0)
    cross_term = sum_embs * sum_embs - xp.sum(st * st, 0)                        # :107-112
    fm_logit = 0.5 * xp.sum(cross_term, -1, keepdims=True)                       # :114
    ffm = []
    for x in USER_SLOTS:                                                         # :117-120, 11-25
        for y in ITEM_SLOTS:
            ffm.append(_dense(xp, general[x], P, "ffm.ffm_x_%s_%s_8" % (x, y)) *
                       _dense(xp, general[y], P, "ffm.ffm_y_%s_%s_8" % (x, y)))
    concated = xp.cat(rw + [cross_term, mult, xp.cat(ffm, -1)] + din_embs, -1)   # :122-123
    gate_input = xp.cat(bias_inputs, -1)                                         # :126
    expert_outs = []
    for i in range(num_experts):                                                 # :130-148
        deep = concated
        for j in range(len(units)):
            g = _dense(xp, _dense(xp, gate_input, P, "experts.gate_%d_%d_1" % (i, j), "relu"), P,
                       "experts.gate_%d_%d_2" % (i, j), "sigmoid") * 2
            deep = g * _dense(xp, deep, P, "experts.expert_output_%d_%d" % (i, j), "relu")
        expert_outs.append(deep)
    ec = xp.stack(expert_outs, 1)                                                # :150
    mmoe = []
    for i in range(num_tasks):                                                   # :153-164
        go = concated
        for j in range(2):
            go = _dense(xp, go, P, "task_gates.gate_%d_%d" % (i, j), "relu")
        go = _dense(xp, go, P, "task_gates.gate_output_%d" % i, "softmax")
        mmoe.append(xp.sum(ec * xp.expand(go, -1), 1))
    ext = xp.cat([mmoe[0], deep_cross(xp, concated, P, "cross")], -1)            # :167-168
    p = xp.softmax(_dense(xp, ext, P, "staytime_output"))                        # :169-170
    bins = xp.const(np.asarray(BIN_LIST).reshape(-1, 1), p)
    pred = p @ bins                                                              # :176
    pred = xp.where(pred < 0.0, pred * 0, pred)                                  # :177
    out = {"staytime": xp.cat([p, pred], -1), "staytime_pred": pred}             # :179
    for k in (1, 2):                                                             # :181-191
        t = task_names[k]
        dl = _dense(xp, mmoe[k], P, "tower_deep.tower_deep_%s" % t, "relu")
        out[t.replace("_pred", "")] = _dense(xp, xp.cat([fm_logit, dl], 1), P, "tower_out.%s" % t, "sigmoid")
    return out


# ------------------------------------------------------------------------------------- rough_rank
def _dnn(xp, x, P, name, n_layers, act="relu", out_act=None):
    """DNN.call (rough_rank/layer.py:100-109)."""
    for i in range(n_layers):
        a = out_act if (i == n_layers - 1 and out_act) else act
        y = x @ P["%s.kernels.%d" % (name, i)] + P["%s.bias.%d" % (name, i)]
        x = xp.relu(y) if a == "relu" else (xp.softmax(y) if a == "softmax" else (xp.sigmoid(y) if a == "sigmoid" else y))
    return x


def ple(xp, x, P, name, num_tasks, n_shared=4, n_specific=4):
    """PLE.call (rough_rank/layer.py:211-224): experts = shared + task-specific, softmax gate per task."""
    shared = [_dnn(xp, x, P, "%s.shared_expert_nets.%d" % (name, e), 1) for e in range(n_shared)]
    outs = []
    for t in range(num_tasks):
        spec = [_dnn(xp, x, P, "%s.specific_expert_nets.%d.%d" % (name, t, e), 1) for e in range(n_specific)]
        gate = _dnn(xp, x, P, "%s.gate_nets.%d" % (name, t), 1, out_act="softmax")
        outs.append(xp.sum(xp.stack(shared + spec, -2) * xp.expand(gate, -1), -2))
    return outs


def cross_net(xp, x, P, name, layer_num=2):
    """CrossNet.call (rough_rank/layer.py:256-264): x_{l+1} = x0 (x_l . w_l) + b_l + x_l."""
    x0 = x
    xl = x
    for i in range(layer_num):
        xw = xl @ P["%s.kernels.%d" % (name, i)]                                  # [B,1]
        xl = x0 * xw + P["%s.bias.%d" % (name, i)][:, 0] + xl
    return xl


def dssm_fwd(xp, embs, dense_mask, P, user_ids, item_ids):
    """DSSM dense graph (rough_rank/model.py:130-170)."""
    srt = lambda ids: xp.cat([embs[k] for k in sorted(ids)], -1)                 # dict_to_sorted_list :10-13
    xu, xi, xt = srt(user_ids), srt(item_ids), srt(list(user_ids) + list(item_ids))
    pu = ple(xp, xu, P, "user.ple", 2)                                           # :46-55
    ou = [_dnn(xp, pu[t], P, "user.heads.%d" % t, 1, out_act="linear") for t in range(2)]
    sel = (dense_mask == 1).reshape(-1, 1)
    user = xp.where(sel, ou[1], ou[0])
    pi = ple(xp, xi, P, "item.ple", 1)                                           # :57-62
    item = _dnn(xp, pi[0], P, "item.heads.0", 1, out_act="linear")
    deep = _dense(xp, _dense(xp, xt, P, "teacher.dense0", "relu"), P, "teacher.dense1", "relu")   # :26-27
    merge = xp.cat([deep, cross_net(xp, xt, P, "teacher.cross")], -1)            # :25,28
    teacher_logit = _dense(xp, _dense(xp, merge, P, "teacher.dense2"), P, "teacher.pred")          # :29-30
    h = _dense(xp, xp.cat([user, item], -1), P, "shallow.shallow_dnn_0", "relu")  # :75-78
    student_logit = _dense(xp, h, P, "shallow.logit_shallow")                    # :80
    kd = xp.mean((xp.detach(teacher_logit) - student_logit) ** 2, -1)            # KDLoss layer.py:277-279
    return {"student": xp.sigmoid(student_logit), "teacher": xp.sigmoid(teacher_logit), "distill": kd,
            "user_emb": user, "item_emb": item}


# ------------------------------------------------------------------------------------- rank/ctr
PPNET_SPLIT = [256, 64, 8, 256, 64, 8, 32, 16]
GATE_FEATURE_LIST = ["1568", "1570", "1578", "1591", "1593", "1614", "1736", "1737", "2039", "2599", "3051", "3303",
                     "3389", "1576", "1577", "1578"]


def rank_ctr_layout(model_config):
    """SingleSlot + BaseModel.__init__ bookkeeping (rank/ctr/base_model.py:14-27,35-86,132-158), restated with
    the reference's own variable roles (last_start / last_end per slot).  Returns (max_embed_size, structure,
    bias, gate) with slices as [slot, start, end]."""
    fs = model_config["feature_slot"]
    slot = {}          # slot -> dict(intervals, last_start, last_end, total)
    bias = {}

    def update(s, emb_size, is_single):                                  # SingleSlot.update_intervals :22-27
        st = slot.setdefault(s, dict(intervals=[], last_start=-1, last_end=-1, total=0))
        st["last_start"] = st["last_end"] + 1
        st["last_end"] = st["last_start"] + emb_size - 1
        if is_single:
            st["intervals"].append([st["last_start"], st["last_end"] + 1])
        st["total"] += emb_size

    for k, ft in fs["sparse_feature"].items():                           # :40-56
        s = ft["slot_id"][0]
        update(s, ft["emb_size"], "bias" not in ft)
        if "bias" in ft:
            if "bias_type" not in ft:
                raise Exception("bias_type could not be null")
            bias.setdefault(s, {})[ft["bias_type"]] = [slot[s]["last_start"], slot[s]["last_end"] + 1]
    for k, ft in fs["sequence_feature"].items():                         # :63-71
        s = ft["slot_id"][0]
        if s in slot:
            raise Exception("sequence feature " + s + "has been defined more than once")
        update(s, ft["emb_size"], True)
    max_embed = max(st["total"] for st in slot.values())                 # :82-86
    structure, gate = [], []
    for s, st in slot.items():                                           # :136-143
        for iv in st["intervals"]:
            structure.append([s, iv[0], iv[1]])
            if s in GATE_FEATURE_LIST:
                gate.append([s, iv[0], iv[1]])
    b = {}
    for s in sorted(bias):                                               # :146-154
        for t, iv in bias[s].items():
            b.setdefault(t, []).append([s, iv[0], iv[1]])
    return max_embed, structure, b, gate


def _interacting(xp, x, P, name, H, L, eps=1e-3, dropout=None):
    """InteractingLayer.call (InteractingLayer.py:37-61) with the layer's own Dense / LayerNorm parameters.
    dropout = (rate, seed): training-mode attention dropout (:53-54) with the kernels' counter-based mask."""
    g = lambda k: P["%s.%s" % (name, k)]
    if xp is NP:
        from . import oracle_np as onp
        W = np.concatenate([g("query_dense_kernel"), g("key_dense_kernel"), g("value_dense_kernel"), g("res_dense_kernel")], 1)
        b = np.concatenate([g("query_dense_bias"), g("key_dense_bias"), g("value_dense_bias"), g("res_dense_bias")])
        return onp.interacting_fwd(x, W, b, g("layer_norm_gamma"), g("layer_norm_beta"), eps, H, L, True, dropout=dropout)
    from . import oracle_torch as ot
    return ot.interacting_layer(x, g("query_dense_kernel"), g("query_dense_bias"), g("key_dense_kernel"), g("key_dense_bias"),
                                g("value_dense_kernel"), g("value_dense_bias"), g("res_dense_kernel"), g("res_dense_bias"),
                                g("layer_norm_gamma"), g("layer_norm_beta"), eps, H, L, True, dropout=dropout)


# rank/multi_head/multidnn.py:207 MultiLabelInfo.label_list — the order of the 7 outputs
AUTOINT_LABELS = ["like_pred", "click_comment_pred", "comment_pred", "click_sharing_pred", "follow_pred",
                  "click_avatar_pred", "unlike_pred"]


def autoint_multihead_fwd(xp, embs, P, deep_hidden_units=(32, 16), dropout=None, eps=1e-3):
    """create_autoint_sub_model (rank/multi_head/multidnn.py:14-212).  embs: list of [B, 8] slot embeddings in
    `linear_features` order; P keyed by the Keras layer names (interacting_layer.*, dnn_{i}, expert_{i}_fc1,
    gate_{i}_fc2, <label>_pred); returns [B, 7] in MultiLabelInfo.label_list order (:206-209).
    dropout = (rate 0.2, seed) in training (:54), None at inference.  The `dense_weight_*` inputs (:30-31) are
    declared by the reference but feed nothing (the sub-model is built on emb_inputs only, :211)."""
    all_inputs = xp.stack([e for e in embs], 1)                                     # :25-27,50  [B, F, 8]
    autoint = _interacting(xp, all_inputs, P, "interacting_layer", 2, 1, eps, dropout)   # :54  (1 iteration, 8 units, 2 heads)
    B = all_inputs.shape[0]
    autoint = autoint.reshape(B, -1)                                                # :56 Flatten
    deep = all_inputs.reshape(B, -1)                                                # :60 Flatten
    for i in range(len(deep_hidden_units)):
        deep = _dense(xp, deep, P, "dnn_%d" % i, "relu")                            # :62-63 (regularizers add no forward term)
    result = xp.cat([deep, autoint], 1)                                             # :72
    experts = xp.stack([_dense(xp, result, P, "expert_%d_fc1" % i, "relu") for i in range(7)], 1)   # :80-92: 8 built, [0:7] used
    preds = []
    for i, label in enumerate(AUTOINT_LABELS):
        gate = _dense(xp, result, P, "gate_%d_fc2" % i, "softmax")                  # :97-99
        mixed = xp.sum(experts * xp.expand(gate, -1), 1)                            # :104-108
        preds.append(_dense(xp, mixed, P, label, "sigmoid"))                        # :118-205
    return xp.cat(preds, 1)


def rank_ctr_fwd(xp, emb, P, structure, bias, gate):
    """Model.model_layer (rank/ctr/model_init.py:19-162).  emb: slot -> [B, max_embed_size]."""
    take = lambda sl: emb[sl[0]][:, sl[1]:sl[2]]
    st = [take(s) for s in structure]
    n = len(st)
    squeeze = xp.detach(xp.cat([xp.sum(e, 1, keepdims=True) / e.shape[1] for e in st], 1))       # :22-31
    w = 2 * _dense(xp, _dense(xp, squeeze, P, "senet_squeeze_layer", "relu"), P, "senet_extract_layer", "sigmoid")
    rw = [e * w[:, i:i + 1] for i, e in enumerate(st)]                                            # :39-41
    fields = xp.stack([_dense(xp, e, P, "emb_linear_map.%d" % i) for i, e in enumerate(rw)], 1)   # :44-49
    auto = _interacting(xp, fields, P, "interact", 2, 1)                                          # :54-59
    auto = auto.reshape(auto.shape[0], -1)                                                        # :60
    ppnet = 2 * _dense(xp, xp.cat([take(s) for s in bias["ppnet"]], 1), P, "dnn_ppnet_gate", "sigmoid")   # :63-65
    offs = np.cumsum([0] + PPNET_SPLIT)
    gates = [ppnet[:, offs[i]:offs[i + 1]] for i in range(len(PPNET_SPLIT))]                      # :66
    deep = xp.cat(rw, 1)                                                                          # :70
    for i in range(2):                                                                            # :72-77
        deep = xp.relu(_dense(xp, deep, P, "dnn.%d" % i) * gates[i + 6])
    mult = xp.relu(xp.cat([take(s) for s in bias["multiply_user"]], 1) *
                   xp.cat([take(s) for s in bias["multiply_item"]], 1))                           # :80-84
    result = xp.cat([deep, auto, mult], 1)                                                        # :88
    can = _dense(xp, xp.cat([take(s) for s in bias["can"]], 1), P, "dnn_can")                     # :92-93
    B = can.shape[0]
    w1, b1 = can[:, 0:48].reshape(B, 8, 6), can[:, 48:54].reshape(B, 1, 6)                        # :94-98
    w2, b2 = can[:, 54:78].reshape(B, 6, 4), can[:, 78:82].reshape(B, 1, 4)
    gate_input = xp.cat([take(s) for s in gate], 1)                                               # :105
    ex = []
    for i in range(3):                                                                            # :106-115
        x = result
        for j in range(2):
            g = 2 * _dense(xp, _dense(xp, gate_input, P, "experts.gate_%d_%d_1" % (i, j), "relu"), P,
                           "experts.gate_%d_%d_2" % (i, j), "sigmoid")
            x = g * _dense(xp, x, P, "experts.expert_output_%d_%d" % (i, j), "relu")
        ex.append(x)
    ec = xp.stack(ex, 1)
    out = {}
    for i in range(2):                                                                            # :123-161
        go = result
        for j in range(2):
            go = _dense(xp, go, P, "task_gates.gate_%d_%d" % (i, j), "relu")
        go = _dense(xp, go, P, "task_gates.gate_output_%d" % i, "softmax")
        r = xp.sum(ec * xp.expand(go, -1), 1)
        for j in range(2):
            if j == 0:
                r = xp.relu(r * gates[i * 3])
            r = xp.relu(_dense(xp, r, P, "task_dnn2.task%d_dnn2_%d" % (i, j)) * gates[i * 3 + j + 1])
        c = xp.relu(xp.expand(r, 1) @ w1 + b1)
        c = xp.relu(c @ w2 + b2)[:, 0, :]
        p = _dense(xp, xp.cat([r, c], 1), P, "task_out.%d" % i, "sigmoid")
        out["task%d" % i] = xp.where(p < 1e-6, p * 0 + 1e-6, p)                                    # clip_by_value(1e-6, 1.0)
    return out
