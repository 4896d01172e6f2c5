"""Random parameter dicts (keyed like the API modules' state_dict) and inputs for the two composed graphs."""
import numpy as np

from util import glorot_uniform

USER4 = ["1568", "1589", "2039", "1570"]
ITEM4 = ["1591", "1593", "1737", "1614"]


def _dense(P, rng, name, fan_in, units, bias_scale=0.1):
    P[name + ".kernel"] = glorot_uniform(rng, fan_in, units)
    P[name + ".bias"] = (bias_scale * rng.standard_normal(units)).astype(np.float32)


def video_dnn_params(rng, slots, seq_slots, units=(256, 128)):
    n = len(slots)
    P = {}
    for s in seq_slots:
        P["din.din_%s.layer_1_kernel" % s] = glorot_uniform(rng, 64, 16)
        P["din.din_%s.layer_1_bias" % s] = (0.1 * rng.standard_normal(16)).astype(np.float32)
        P["din.din_%s.layer_2_kernel" % s] = glorot_uniform(rng, 16, 1)
        P["din.din_%s.layer_2_bias" % s] = (0.1 * rng.standard_normal(1)).astype(np.float32)
    _dense(P, rng, "senet_squeeze_layer1", 16 * n, int(n / 4))
    _dense(P, rng, "senet_extract_layer2", int(n / 4), n)
    for x in USER4:
        for y in ITEM4:
            _dense(P, rng, "ffm.ffm_x_%s_%s_8" % (x, y), 16, 8)
            _dense(P, rng, "ffm.ffm_y_%s_%s_8" % (x, y), 16, 8)
    width = 16 * n + 16 + 64 + 128 + 16 * len(seq_slots)
    for i in range(3):
        w = width
        for j, u in enumerate(units):
            _dense(P, rng, "experts.gate_%d_%d_1" % (i, j), 224, u)
            _dense(P, rng, "experts.gate_%d_%d_2" % (i, j), u, u)
            _dense(P, rng, "experts.expert_output_%d_%d" % (i, j), w, u)
            w = u
    for i in range(3):
        _dense(P, rng, "task_gates.gate_%d_0" % i, width, 64)
        _dense(P, rng, "task_gates.gate_%d_1" % i, 64, 32)
        _dense(P, rng, "task_gates.gate_output_%d" % i, 32, 3)
    for i in range(3):
        P["cross.W.%d" % i] = (glorot_uniform(rng, width, 1) * 0.3).astype(np.float32)
        P["cross.b.%d" % i] = (0.05 * rng.standard_normal(width)).astype(np.float32)
    _dense(P, rng, "staytime_output", units[-1] + width, 400)
    for t in ("shortplay_pred", "longplay_pred"):
        _dense(P, rng, "tower_deep.tower_deep_%s" % t, units[-1], 1)
        _dense(P, rng, "tower_out.%s" % t, 2, 1)
    return P


def video_dnn_inputs(rng, B, T, slots, seq_slots, scale=0.3):
    embs = {s: (scale * rng.standard_normal((B, 32))).astype(np.float32) for s in slots}
    seqs = {}
    for s in seq_slots:
        lens = rng.integers(0, T + 1, size=B)
        lens[0] = T
        if B > 1:
            lens[1] = 0                                   # an all-masked row: softmax over equal pads = mean
        mask = np.arange(T)[None, :] < lens[:, None]
        seqs[s] = ((scale * rng.standard_normal((B, T, 32))).astype(np.float32), mask)
    return embs, seqs


def _dnn(P, rng, name, dims):
    for i in range(len(dims) - 1):
        P["%s.kernels.%d" % (name, i)] = glorot_uniform(rng, dims[i], dims[i + 1])
        P["%s.bias.%d" % (name, i)] = (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32)


def dssm_params(rng, user_ids, item_ids, d=16):
    P = {}
    for tower, ids, nt in (("user", user_ids, 2), ("item", item_ids, 1)):
        w = d * len(ids)
        for e in range(4):
            _dnn(P, rng, "%s.ple.shared_expert_nets.%d" % (tower, e), [w, 32])
        for t in range(nt):
            for e in range(4):
                _dnn(P, rng, "%s.ple.specific_expert_nets.%d.%d" % (tower, t, e), [w, 32])
            _dnn(P, rng, "%s.ple.gate_nets.%d" % (tower, t), [w, 8])
            _dnn(P, rng, "%s.heads.%d" % (tower, t), [32, 16])
    w = d * (len(user_ids) + len(item_ids))
    for i in range(2):
        P["teacher.cross.kernels.%d" % i] = (glorot_uniform(rng, w, 1) * 0.3).astype(np.float32)
        P["teacher.cross.bias.%d" % i] = (0.05 * rng.standard_normal((w, 1))).astype(np.float32)
    _dense(P, rng, "teacher.dense0", w, 128)
    _dense(P, rng, "teacher.dense1", 128, 64)
    _dense(P, rng, "teacher.dense2", 64 + w, 16)
    _dense(P, rng, "teacher.pred", 16, 1)
    _dense(P, rng, "shallow.shallow_dnn_0", 32, 32)
    _dense(P, rng, "shallow.logit_shallow", 32, 1)
    return P


def rank_ctr_config(rng, n_common=14, n_extra_slots=3):
    """A small synthetic model_config with the structure of rank/ctr/model_parameter.json: common features of
    width 8 / 16, slots shared by several features, all four bias types, gate slots, a sequence feature."""
    sparse = {}
    gate_slots = ["1568", "1570", "1578", "1591"]
    slots = gate_slots + [str(5000 + i) for i in range(n_common - len(gate_slots))]
    for i, s in enumerate(slots):
        sparse["f_common_%d" % i] = {"emb_size": int(rng.choice([8, 16])), "slot_id": [s]}
    for i in range(n_extra_slots):                      # a second common feature on an existing slot
        sparse["f_second_%d" % i] = {"emb_size": 8, "slot_id": [slots[i]]}
    for i in range(4):
        sparse["f_ppnet_%d" % i] = {"emb_size": 8, "bias": True, "bias_type": "ppnet", "slot_id": [slots[i]]}
        sparse["f_can_%d" % i] = {"emb_size": 8, "bias": True, "bias_type": "can", "slot_id": [slots[i + 2]]}
    for i in range(2):
        sparse["f_mu_%d" % i] = {"emb_size": 16, "bias": True, "bias_type": "multiply_user", "slot_id": [slots[i + 5]]}
        sparse["f_mi_%d" % i] = {"emb_size": 16, "bias": True, "bias_type": "multiply_item", "slot_id": [slots[i + 7]]}
    sparse["f_bias_only"] = {"emb_size": 8, "bias": True, "bias_type": "ppnet", "slot_id": ["7000"]}
    return {"feature_slot": {"sparse_feature": sparse,
                             "sequence_feature": {"seq_0": {"emb_size": 16, "slot_id": ["8000"]}},
                             "dense_feature": {}}}


def rank_ctr_params(rng, structure, bias, gate):
    n = len(structure)
    w = lambda sl: sl[2] - sl[1]
    P = {}
    _dense(P, rng, "senet_squeeze_layer", n, n // 4)
    _dense(P, rng, "senet_extract_layer", n // 4, n)
    for i, sl in enumerate(structure):
        _dense(P, rng, "emb_linear_map.%d" % i, w(sl), 8)
    for nm in ("query", "key", "value", "res"):
        P["interact.%s_dense_kernel" % nm] = glorot_uniform(rng, 8, 8)
        P["interact.%s_dense_bias" % nm] = (0.1 * rng.standard_normal(8)).astype(np.float32)
    P["interact.layer_norm_gamma"] = (1 + 0.1 * rng.standard_normal(8)).astype(np.float32)
    P["interact.layer_norm_beta"] = (0.1 * rng.standard_normal(8)).astype(np.float32)
    _dense(P, rng, "dnn_ppnet_gate", sum(w(s) for s in bias["ppnet"]), 704)
    tot = sum(w(s) for s in structure)
    _dense(P, rng, "dnn.0", tot, 32)
    _dense(P, rng, "dnn.1", 32, 16)
    _dense(P, rng, "dnn_can", sum(w(s) for s in bias["can"]), 82)
    res_w = 16 + 8 * n + sum(w(s) for s in bias["multiply_user"])
    gw = sum(w(s) for s in gate)
    for i in range(3):
        x = res_w
        for j, u in enumerate((512, 256)):
            _dense(P, rng, "experts.gate_%d_%d_1" % (i, j), gw, u)
            _dense(P, rng, "experts.gate_%d_%d_2" % (i, j), u, u)
            _dense(P, rng, "experts.expert_output_%d_%d" % (i, j), x, u)
            x = u
    for i in range(2):
        _dense(P, rng, "task_gates.gate_%d_0" % i, res_w, 256)
        _dense(P, rng, "task_gates.gate_%d_1" % i, 256, 32)
        _dense(P, rng, "task_gates.gate_output_%d" % i, 32, 3)
        _dense(P, rng, "task_dnn2.task%d_dnn2_0" % i, 256, 64)
        _dense(P, rng, "task_dnn2.task%d_dnn2_1" % i, 64, 8)
        _dense(P, rng, "task_out.%d" % i, 12, 1)
    return P
