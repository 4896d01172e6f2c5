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
