"""Generates tests/golden/ref_layers.npz by EXECUTING the reference's own layer code, unmodified, from /root/reference:

    InteractingLayer.py::InteractingLayer.call           (AutoInt self-attention, the headline hot path)
    din.py::DIN.call                                      (DIN variant A)
    staytime/layer.py::DIN.call, DeepCrossLayer.call, FMLayer.call
    rough_rank/layer.py::DNN.call, CrossNet.call, PLE.call
    staytime/VideoDnn.py::create_moe_sub_model            (the whole dense graph of BASELINE configs[4], Keras functional
                                                           code run eagerly on seeded inputs)
    rank/multi_head/multidnn.py::create_autoint_sub_model (BASELINE configs[3]; with rank/multi_head/interacting_layer.py)
    rank/ctr/base_model.py::BaseModel.__init__ + model_init.py::Model.model_layer  (the production rank/ctr model)
    rough_rank/model.py::DSSM                             (user / item / teacher / shallow towers + distillation)
    staytime/parse.py::parse_input_func                   (the label transform; tf.io.parse_example handed in ready)
    autoint::AutoInt.model_layer                          (THE HEADLINE MODEL, BASELINE configs[0] / [1]; MultiLayerDense,
                                                           a file missing from the reference tree, restated as a Dense stack)
    rank/ctr/base_model.py::cross_entropy, staytime/model.py::custom_kl_loss / cross_entropy / mse_loss / huber_loss

TensorFlow is not installable offline, so `tensorflow` is replaced by oracle/tf_numpy_shim.py — a numpy fp64 stand-in
for the individual TF ops those files call (each with its documented semantics).  What is pinned is therefore the
reference's COMPOSITION of those ops (split / concat order of the heads, scaling, masks, residual, the loop over
layer_num re-using the same Dense layers, ...), which is exactly what oracle/oracle_np.py restates by hand.
`.layer_normalization.LayerNormalization` is imported by the reference but is not in its tree; the shim's restatement
is used (mean / variance over the last axis, gamma, beta, eps).

Golden inputs: seeded activations and the weights the layers created (Keras [in, out] kernels); golden outputs: what
the reference code returned.  tests/test_oracle_reference_pin.py compares the oracle against them (CPU, no reference tree
needed at test time).

    python tools/gen_reference_layer_golden.py        (only in the container that has /root/reference)
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_numpy_shim as shim  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_layers.npz")


def load(path, name, package=None):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def manifest(keys_by_id):
    """[[oracle key | "", shape, offset], ...] of the weights drawn since the last shim.seed, in creation order."""
    import json
    rows = []
    for w in shim.WEIGHT_LOG:
        key, off = keys_by_id.get(id(w), ("", 0.0))
        rows.append([key, list(w.shape), off])
    return np.asarray(json.dumps(rows))


def main():
    shim.install()
    # InteractingLayer.py does `from .layer_normalization import LayerNormalization`: give it a parent package
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [REF]
    sys.modules["refpkg"] = pkg
    ln = types.ModuleType("refpkg.layer_normalization")
    ln.LayerNormalization = shim.LayerNormalization
    sys.modules["refpkg.layer_normalization"] = ln
    inter = load("InteractingLayer.py", "refpkg.InteractingLayer", package="refpkg")
    din_a = load("din.py", "ref_din")
    stay = load("staytime/layer.py", "ref_staytime_layer")
    rough = load("rough_rank/layer.py", "ref_rough_layer")

    rng = np.random.default_rng(20240)
    out = {}

    # ---- InteractingLayer: (tag, B, F, D, layer_num, unit_num, head_num, use_res)
    for k, (tag, B, F, D, L, U, H, res) in enumerate([("cfg1", 5, 39, 16, 3, 16, 2, True), ("rankctr", 4, 11, 8, 1, 8, 2, True),
                                                      ("nores", 3, 7, 16, 2, 16, 4, False), ("h1", 2, 5, 16, 1, 16, 1, True)]):
        shim.seed(100 + k)
        layer = inter.InteractingLayer(layer_num=L, unit_num=U, head_num=H, use_dropout=False, use_res=res)
        x = rng.standard_normal((B, F, D))
        y = layer(shim.T(x))
        zeros_w, zeros_b = np.zeros((D, U)), np.zeros(U)
        Wr = layer.res_dense.kernel if res else zeros_w
        br = layer.res_dense.bias if res else zeros_b
        out[f"inter_{tag}_x"] = x
        out[f"inter_{tag}_W"] = np.concatenate([layer.query_dense.kernel, layer.key_dense.kernel, layer.value_dense.kernel, Wr], 1)
        out[f"inter_{tag}_b"] = np.concatenate([layer.query_dense.bias, layer.key_dense.bias, layer.value_dense.bias, br])
        out[f"inter_{tag}_gamma"] = np.asarray(layer.layer_norm.gamma)
        out[f"inter_{tag}_beta"] = np.asarray(layer.layer_norm.beta)
        out[f"inter_{tag}_cfg"] = np.asarray([H, L, int(res)], np.int64)
        out[f"inter_{tag}_eps"] = np.asarray(layer.layer_norm.eps)
        out[f"inter_{tag}_y"] = np.asarray(y)

    # ---- DIN variant A (din.py): relu MLP scores, sequence mask, scores @ values
    shim.seed(11)
    B, T, H = 6, 9, 16
    layer = din_a.DIN()
    q, keys, values = rng.standard_normal((B, H)), rng.standard_normal((B, T, H)), rng.standard_normal((B, T, H))
    seq_len = np.array([9, 1, 4, 0, 7, 9], np.int64)      # tf.sequence_mask: maxlen = max(lengths) must equal T
    y = layer(shim.T(q), shim.T(keys), shim.T(values), seq_len)
    out.update(dina_q=q, dina_keys=keys, dina_values=values, dina_seq_len=seq_len, dina_W1=layer.nn[0].kernel,
               dina_b1=layer.nn[0].bias, dina_W2=layer.nn[1].kernel, dina_b2=layer.nn[1].bias, dina_y=np.asarray(y))

    # ---- DIN variant B (staytime/layer.py): sigmoid hidden, -2^32+1 padding, softmax, mask wider than the sequence
    shim.seed(12)
    B, T, H, Tm = 5, 7, 16, 10
    layer = stay.DIN()
    q, facts = rng.standard_normal((B, H)), rng.standard_normal((B, T, H))
    mask = rng.random((B, Tm)) < 0.6
    mask[0, :] = True
    mask[1, :T] = False                                    # a fully masked row: uniform softmax over the paddings
    y = layer(shim.T(q), shim.T(facts), shim.T(mask))
    out.update(dinb_q=q, dinb_facts=facts, dinb_mask=mask, dinb_W1=layer.layer_1.kernel, dinb_b1=layer.layer_1.bias,
               dinb_W2=layer.layer_2.kernel, dinb_b2=layer.layer_2.bias, dinb_y=np.asarray(y))

    # ---- DeepCrossLayer / FMLayer (staytime/layer.py), CrossNet / DNN (rough_rank/layer.py)
    shim.seed(13)
    x = rng.standard_normal((7, 24))
    layer = stay.DeepCrossLayer(num_layer=3)
    y = layer(shim.T(x))
    out.update(dcross_x=x, dcross_W=np.stack([np.asarray(w) for w in layer.W]), dcross_b=np.stack([np.asarray(b) for b in layer.bias]),
               dcross_y=np.asarray(y))
    x3 = rng.standard_normal((4, 9, 16))
    out.update(fm_x=x3, fm_y=np.asarray(stay.FMLayer()(shim.T(x3))))
    shim.seed(14)
    x = rng.standard_normal((6, 20))
    layer = rough.CrossNet(layer_num=2)
    y = layer(shim.T(x))
    out.update(cnet_x=x, cnet_k=np.stack([np.asarray(k) for k in layer.kernels]), cnet_b=np.stack([np.asarray(b) for b in layer.bias]),
               cnet_y=np.asarray(y))
    shim.seed(15)
    x = rng.standard_normal((5, 12))
    layer = rough.DNN((8, 6, 3), activation="relu", output_activation="sigmoid")
    y = layer(shim.T(x))
    for i in range(3):
        out[f"dnn_k{i}"], out[f"dnn_b{i}"] = np.asarray(layer.kernels[i]), np.asarray(layer.bias[i])
    out.update(dnn_x=x, dnn_y=np.asarray(y))

    # ---- PLE (rough_rank/layer.py:174-224), as rough_rank/model.py::create_tower instantiates it (2 tasks, 4 + 4 experts)
    shim.seed(18)
    x = rng.standard_normal((6, 20))
    layer = rough.PLE(name="ple_user", num_tasks=2, num_shared_experts=4, num_specific_experts=4, expert_dnn_units=(32,),
                      gate_dnn_units=(), expert_dnn_params=dict(), gate_dnn_params=dict())
    ys = layer(shim.T(x))
    out.update(ple_x=x, ple_y=np.stack([np.asarray(y_) for y_ in ys]))
    for e_, net in enumerate(layer.shared_expert_nets):
        out[f"ple_P_p.shared_expert_nets.{e_}.kernels.0"], out[f"ple_P_p.shared_expert_nets.{e_}.bias.0"] = net.kernels[0], net.bias[0]
    for t_ in range(2):
        for e_, net in enumerate(layer.specific_expert_nets[t_]):
            out[f"ple_P_p.specific_expert_nets.{t_}.{e_}.kernels.0"] = net.kernels[0]
            out[f"ple_P_p.specific_expert_nets.{t_}.{e_}.bias.0"] = net.bias[0]
        out[f"ple_P_p.gate_nets.{t_}.kernels.0"], out[f"ple_P_p.gate_nets.{t_}.bias.0"] = layer.gate_nets[t_].kernels[0], layer.gate_nets[t_].bias[0]

    # ---- the composed dense graph of BASELINE configs[4]: staytime/VideoDnn.py::create_moe_sub_model, executed eagerly
    # (tn.layers.Input hands back the seeded arrays), on a reduced slot list that keeps every user / item / bias slot
    pk = types.ModuleType("video_id_rank_staytime_mtl_ppnet_v7")
    pk.__path__ = []
    pm = types.ModuleType("video_id_rank_staytime_mtl_ppnet_v7.model")
    pm.__path__ = []
    sys.modules["video_id_rank_staytime_mtl_ppnet_v7"], sys.modules["video_id_rank_staytime_mtl_ppnet_v7.model"] = pk, pm
    cfgm = load("staytime/config.py", "video_id_rank_staytime_mtl_ppnet_v7.model.config")
    sys.modules["video_id_rank_staytime_mtl_ppnet_v7.model.layers"] = stay
    vd = load("staytime/VideoDnn.py", "ref_videodnn")
    bias = ['3051', '1570', '2039', '2544', '1568', '3376', '3365', '3369', '2597', '1737', '1593', '1591', '1589', '1614']
    slots = sorted(set(bias) | {'1571', '1574', '2040', '4500'})           # 18 slots (len / 4 is not an integer: :82)
    seq_slots = sorted(cfgm.Config.SEQ_SLOTS)
    Bv, Tv = 6, 5
    shim.seed(16)
    del shim.LAYERS[:], shim.WEIGHT_LOG[:]
    feats = {s: types.SimpleNamespace(feature_id=s) for s in slots + seq_slots}
    for s_ in slots:
        shim.FEEDS[s_] = 0.3 * rng.standard_normal((Bv, 32))
        out[f"vd_emb_{s_}"] = shim.FEEDS[s_]
    for s_ in seq_slots:
        lens = np.array([Tv, 0, 3, 1, 5, 2])
        shim.FEEDS[f"seq_emb_{s_}"] = 0.3 * rng.standard_normal((Bv, Tv, 32))
        shim.FEEDS[f"seq_mask_{s_}"] = np.arange(Tv)[None, :] < lens[:, None]
        out[f"vd_seq_{s_}"], out[f"vd_mask_{s_}"] = shim.FEEDS[f"seq_emb_{s_}"], shim.FEEDS[f"seq_mask_{s_}"]
    models = vd.create_moe_sub_model([(feats[s_], (None, 32)) for s_ in slots],
                                     [(feats[s_], (None, Tv, 32), (None, Tv)) for s_ in seq_slots], (16, 8))
    for k_, v_ in models["sub_model_train"].outputs.items():
        out["vd_train_" + k_] = np.asarray(v_)
    for k_, v_ in models["sub_model_predict"].outputs.items():
        out["vd_predict_" + k_] = np.asarray(v_)
    # weights: re-drawn by the test from (seed, manifest); keyed by the layer names the reference gave them (the
    # oracle's / the product's state_dict keys)
    ids = {}
    for layer in shim.LAYERS:
        n_ = layer.name
        if isinstance(layer, stay.DIN):
            for a_, d_ in (("layer_1", layer.layer_1), ("layer_2", layer.layer_2)):
                ids[id(d_.kernel)], ids[id(d_.bias)] = (f"din.{n_}.{a_}_kernel", 0.0), (f"din.{n_}.{a_}_bias", 0.0)
        elif isinstance(layer, stay.DeepCrossLayer):
            for i in range(layer.num_layer):
                ids[id(layer.W[i])], ids[id(layer.bias[i])] = (f"cross.W.{i}", 0.0), (f"cross.b.{i}", 0.0)
        elif isinstance(layer, shim.Dense) and n_:
            if n_.startswith("ffm_"):
                key = "ffm." + n_
            elif n_.startswith("expert_output_") or (n_.startswith("gate_") and n_.count("_") == 3):
                key = "experts." + n_
            elif n_.startswith("gate_"):
                key = "task_gates." + n_
            elif n_.startswith("tower_deep_"):
                key = "tower_deep." + n_
            elif n_ in cfgm.Config.task_names:
                key = "tower_out." + n_
            else:
                key = n_
            ids[id(layer.kernel)], ids[id(layer.bias)] = (key + ".kernel", 0.0), (key + ".bias", 0.0)
    out["vd_seed"], out["vd_manifest"] = np.asarray(16), manifest(ids)
    out["vd_slots"], out["vd_seq_slots"] = np.asarray(slots), np.asarray(seq_slots)

    # ---- the dense graph of BASELINE configs[3]: rank/multi_head/multidnn.py::create_autoint_sub_model (InteractingLayer
    # from rank/multi_head/interacting_layer.py, DNN, 8 experts of which 7 are mixed, 7 softmax gates, 7 sigmoid heads)
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k_, v_ in attrs.items():
            setattr(m, k_, v_)
        sys.modules[name] = m
        return m
    for n_ in ("src", "src.pipeline", "src.model", "src.util"):
        stub(n_)
    stub("src.pipeline.model_result", ModelResult=object)
    stub("src.model.feature_column", tn_category_columns_builder=None, embedding_columns_builder=None, create_emb_model=None)
    stub("src.pipeline.multi_sparse_table", MultiSparseTableInfo=types.SimpleNamespace(
        add_sparse_table_name_mapping_to_input_tensor_prefix=lambda a, b: None))
    stub("src.util.tools")
    stub("src.pipeline.multi_label", MultiLabelInfo=types.SimpleNamespace(label_list=None))
    stub("interact_multihead_autoint_alllabel")
    stub("interact_multihead_autoint_alllabel.model")
    stub("interact_multihead_autoint_alllabel.model.layer_normalization", LayerNormalization=shim.LayerNormalization)
    il = load("rank/multi_head/interacting_layer.py", "interact_multihead_autoint_alllabel.model.interacting_layer")
    md = load("rank/multi_head/multidnn.py", "ref_multidnn")
    shim.seed(17)
    del shim.LAYERS[:], shim.WEIGHT_LOG[:]
    Ba, Fa = 5, 39
    embs_a = [0.5 * rng.standard_normal((Ba, 8)) for _ in range(Fa)]
    for i_, e_ in enumerate(embs_a):
        shim.FEEDS["emb_%d" % i_] = e_
    model = md.create_autoint_sub_model([(i_, e_) for i_, e_ in enumerate(embs_a)], {}, (32, 16), False)
    out["ai_embs"] = np.stack(embs_a)
    out["ai_y"] = np.concatenate([np.asarray(o) for o in model.outputs], 1)
    out["ai_labels"] = np.asarray(sys.modules["src.pipeline.multi_label"].MultiLabelInfo.label_list)
    ids = {}
    for layer in shim.LAYERS:
        if isinstance(layer, il.InteractingLayer):
            for nm in ("query", "key", "value", "res"):
                d_ = getattr(layer, nm + "_dense")
                ids[id(d_.kernel)], ids[id(d_.bias)] = (f"interacting_layer.{nm}_dense_kernel", 0.0), (f"interacting_layer.{nm}_dense_bias", 0.0)
            ids[id(layer.layer_norm._w["gamma"])] = ("interacting_layer.layer_norm_gamma", 1.0)
            ids[id(layer.layer_norm._w["beta"])] = ("interacting_layer.layer_norm_beta", 0.0)
            out["ai_eps"] = np.asarray(layer.layer_norm.eps)
        elif isinstance(layer, shim.Dense) and layer.name:
            ids[id(layer.kernel)], ids[id(layer.bias)] = (layer.name + ".kernel", 0.0), (layer.name + ".bias", 0.0)
    out["ai_seed"], out["ai_manifest"] = np.asarray(17), manifest(ids)

    # ---- the production model of rank/ctr: base_model.py::BaseModel.__init__ (slot slicing) + model_init.py::Model.
    # model_layer (SENet, per-field linear maps, InteractingLayer, PPNet gates, CAN co-action, gated experts, MMoE, two
    # towers), on the small model_config of the GPU parity test (tests/util_models.py::rank_ctr_config)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import json
    import util_models as um
    cfg_rc = um.rank_ctr_config(np.random.default_rng(2))
    Brc = 4

    class Feature:
        def __init__(self, feature_id=None, feature_slot=None, sparse=True, **kw):
            self.feature_id = feature_id

        def __lt__(self, other):
            return self.feature_id < other.feature_id

    class EmbeddingFeatures:
        def __init__(self, cols, opt, name=None):
            self.dim = cols[0][1]

        def __call__(self, inputs):
            embs_ = {}
            for fea_id in inputs:
                embs_[fea_id] = shim.T(0.3 * rng.standard_normal((Brc, self.dim)))
                shim.FEEDS["emb_%s" % fea_id] = embs_[fea_id]
            return embs_
    tnm = sys.modules["tensornet"]
    tnm.feature_column = types.SimpleNamespace(FeatureSlot=lambda s_: s_, Feature=Feature,
                                               category_column=lambda key, bucket_size: key)
    tnm.core = types.SimpleNamespace(Adam=lambda **kw: None)
    tnm.layers.EmbeddingFeatures = EmbeddingFeatures
    sys.modules["tensorflow"].feature_column = types.SimpleNamespace(
        embedding_column=lambda col, dimension, combiner: (col, dimension))
    stub("rcpkg")
    stub("rcpkg.common_module")
    stub("rcpkg.common_module.interacting_layer", InteractingLayer=inter.InteractingLayer)
    stub("rcpkg.common_module.multi_dense_layer", MultiLayerDense=None)       # imported, its only use is commented out
    load("rank/ctr/base_model.py", "rcpkg.base_model", package="rcpkg")
    mi = load("rank/ctr/model_init.py", "rcpkg.model_init", package="rcpkg")
    shim.seed(19)
    del shim.LAYERS[:], shim.WEIGHT_LOG[:]
    shim.FEEDS.clear()
    model = mi.Model(cfg_rc)
    model.model_layer()
    out["rc_config"] = np.asarray(json.dumps(cfg_rc))
    for k_, v_ in shim.FEEDS.items():
        out["rc_" + k_] = np.asarray(v_)
    for i_, t_ in enumerate(['video_id_rank_hp_ctr_addfeasetwo_click', 'video_id_rank_hp_ctr_addfeasetwo_effect_click']):
        out["rc_task%d" % i_] = np.asarray(model.output[t_])
    ids = {}
    unnamed = [l_ for l_ in shim.LAYERS if isinstance(l_, shim.Dense) and not l_.name and l_.units == 1]
    for layer in shim.LAYERS:
        n_ = layer.name
        if isinstance(layer, inter.InteractingLayer):
            for nm in ("query", "key", "value", "res"):
                d_ = getattr(layer, nm + "_dense")
                ids[id(d_.kernel)], ids[id(d_.bias)] = (f"interact.{nm}_dense_kernel", 0.0), (f"interact.{nm}_dense_bias", 0.0)
            ids[id(layer.layer_norm._w["gamma"])] = ("interact.layer_norm_gamma", 1.0)
            ids[id(layer.layer_norm._w["beta"])] = ("interact.layer_norm_beta", 0.0)
            out["rc_eps"] = np.asarray(layer.layer_norm.eps)
        elif isinstance(layer, shim.Dense) and n_:
            if n_.startswith("emb_linear_map_"):
                key = "emb_linear_map." + n_[len("emb_linear_map_"):]
            elif n_ in ("dnn_0", "dnn_1"):
                key = "dnn." + n_[4:]
            elif n_.startswith("expert_output_") or (n_.startswith("gate_") and n_.count("_") == 3):
                key = "experts." + n_
            elif n_.startswith("gate_"):
                key = "task_gates." + n_
            elif n_.startswith("task") and "_dnn2_" in n_:
                key = "task_dnn2." + n_
            else:
                key = n_
            ids[id(layer.kernel)], ids[id(layer.bias)] = (key + ".kernel", 0.0), (key + ".bias", 0.0)
    for i_, layer in enumerate(unnamed):                                    # the two unnamed Dense(1, sigmoid) heads (:156)
        ids[id(layer.kernel)], ids[id(layer.bias)] = ("task_out.%d.kernel" % i_, 0.0), ("task_out.%d.bias" % i_, 0.0)
    out["rc_seed"], out["rc_manifest"] = np.asarray(19), manifest(ids)

    # ---- THE HEADLINE MODEL: /root/reference/autoint::AutoInt.model_layer on BaseModel's field list (39 fields x 16,
    # InteractingLayer x3 / 2 heads, DNN 256-128, Dense(1, sigmoid), clip) — BASELINE configs[0] / configs[1].
    # `.common_module.multi_dense_layer.MultiLayerDense` is imported by the reference but NOT in its tree: restated here as
    # what its call sites say (units list + one activation: a stack of Dense layers).
    class MultiLayerDense(shim.Layer):
        def __init__(self, units, activation=None, **kw):
            super().__init__()
            self.stack = [shim.Dense(u, activation=activation, name="mld_%d" % i) for i, u in enumerate(units)]

        def build(self, input_shape):
            self.built = True

        def call(self, x):
            for d_ in self.stack:
                x = d_(x)
            return x
    sys.modules["rcpkg.common_module.multi_dense_layer"].MultiLayerDense = MultiLayerDense
    import importlib.machinery
    loader = importlib.machinery.SourceFileLoader("rcpkg.autoint", os.path.join(REF, "autoint"))
    spec = importlib.util.spec_from_loader("rcpkg.autoint", loader)
    am = importlib.util.module_from_spec(spec)
    am.__package__ = "rcpkg"
    sys.modules["rcpkg.autoint"] = am
    loader.exec_module(am)
    Fh, Bh = 39, 4
    cfg_h = {"feature_slot": {"sparse_feature": {"f%02d" % i: {"emb_size": 16, "slot_id": [str(1000 + i)]} for i in range(Fh)},
                              "sequence_feature": {}, "dense_feature": {}},
             "model_param": {"interact": {"layer_num": 3, "unit_num": 16, "head_num": 2, "use_dropout": False,
                                          "dropout_rate": 0.0, "use_res": True},
                             "mlp": {"hidden_units": [256, 128], "activation": "relu"},
                             "logits": {"hidden_units": [1], "activation": "sigmoid"}}}
    shim.seed(20)
    del shim.LAYERS[:], shim.WEIGHT_LOG[:]
    shim.FEEDS.clear()
    model = am.AutoInt(cfg_h)
    model.model_layer()
    out["hl_X"] = np.stack([np.asarray(shim.FEEDS["emb_%d" % (1000 + i)]) for i in range(Fh)], 1)     # [B, F, 16]
    out["hl_p"] = np.asarray(model.output)
    ids = {}
    mlds = [l_ for l_ in shim.LAYERS if isinstance(l_, MultiLayerDense)]
    for layer in shim.LAYERS:
        if isinstance(layer, inter.InteractingLayer):
            for nm in ("query", "key", "value", "res"):
                d_ = getattr(layer, nm + "_dense")
                ids[id(d_.kernel)], ids[id(d_.bias)] = (f"{nm}_kernel", 0.0), (f"{nm}_bias", 0.0)
            ids[id(layer.layer_norm._w["gamma"])] = ("gamma", 1.0)
            ids[id(layer.layer_norm._w["beta"])] = ("beta", 0.0)
            out["hl_eps"] = np.asarray(layer.layer_norm.eps)
    for i_, d_ in enumerate(mlds[0].stack):
        ids[id(d_.kernel)], ids[id(d_.bias)] = ("mlp_W%d" % i_, 0.0), ("mlp_b%d" % i_, 0.0)
    ids[id(mlds[1].stack[0].kernel)], ids[id(mlds[1].stack[0].bias)] = ("out_W", 0.0), ("out_b", 0.0)
    out["hl_seed"], out["hl_manifest"] = np.asarray(20), manifest(ids)

    # ---- DSSM (rough_rank/model.py::DSSM): user tower (PLE, 2 tasks, selected per sample by the dense feature 4575),
    # item tower (PLE), teacher (CrossNet + DNN), shallow student tower on the two tower outputs, distillation loss —
    # on the reference's own feature lists (rough_rank/config)
    stub("rrpkg")
    stub("rrpkg.config").__path__ = [os.path.join(REF, "rough_rank", "config")]
    stub("rrpkg.sub")
    load("rough_rank/config/feature_id.py", "rrpkg.config.feature_id", package="rrpkg.config")
    rcfg = load("rough_rank/config/config.py", "rrpkg.config.config", package="rrpkg.config")
    sys.modules["rrpkg.config"].config = rcfg
    sys.modules["rrpkg.sub.layer"] = rough
    Bd = 5
    dense_mask = np.array([[1.0], [0.0], [1.0], [0.0], [0.0]])

    class Feature2:
        def __init__(self, feature_id=None, feature_slot=None, sparse=True, feature_name=None, **kw):
            self.feature_id, self.feature_name = feature_id, feature_name

    class EmbeddingFeatures2:
        def __init__(self, cols, opt, name=None):
            self.cols = cols

        def __call__(self, inputs):
            embs_ = {}
            for key, dim in self.cols:
                embs_[key] = shim.T(0.3 * rng.standard_normal((Bd, dim)))
                for tower in ("user", "item", "teacher"):
                    shim.FEEDS["emb_%s_%s" % (tower, key)] = embs_[key]
            return embs_
    tnm.feature_column = types.SimpleNamespace(FeatureSlot=lambda s_: s_, Feature=Feature2,
                                               category_column=lambda key, bucket_size: key)
    tnm.layers.EmbeddingFeatures = EmbeddingFeatures2
    shim.seed(21)
    del shim.LAYERS[:], shim.WEIGHT_LOG[:]
    shim.FEEDS.clear()
    shim.FEEDS["dense_weight_4575"] = dense_mask
    dm = load("rough_rank/model.py", "rrpkg.sub.model", package="rrpkg.sub")
    models = dm.DSSM()
    outs = models["train"].outputs
    user_ids, item_ids = [str(v) for v in rcfg.USER_FEATURE_IDS], [str(v) for v in rcfg.ITEM_FEATURE_IDS]
    out["ds_user_ids"], out["ds_item_ids"], out["ds_mask"] = np.asarray(user_ids), np.asarray(item_ids), dense_mask
    for k_ in rcfg.USER_FEATURE_IDS + rcfg.ITEM_FEATURE_IDS:
        out["ds_emb_%s" % k_] = np.asarray(shim.FEEDS["emb_teacher_%s" % k_])
    out["ds_student"], out["ds_teacher"], out["ds_distill"] = (np.asarray(outs[k_]) for k_ in ("student", "teacher", "distill"))
    out["ds_user_emb"] = np.asarray(shim.FEEDS["shallow_user_emb_output"])
    out["ds_item_emb"] = np.asarray(shim.FEEDS["shallow_item_emb_output"])
    ids = {}

    def dnn_keys(net, key):
        for i_ in range(len(net.kernels)):
            ids[id(net.kernels[i_])], ids[id(net.bias[i_])] = ("%s.kernels.%d" % (key, i_), 0.0), ("%s.bias.%d" % (key, i_), 0.0)
    teacher_dense = [l_ for l_ in shim.LAYERS if isinstance(l_, shim.Dense) and not l_.name]
    for i_, d_ in enumerate(teacher_dense):                     # Dense(128), Dense(64), Dense(16) of create_tower_teacher
        ids[id(d_.kernel)], ids[id(d_.bias)] = ("teacher.dense%d.kernel" % i_, 0.0), ("teacher.dense%d.bias" % i_, 0.0)
    for layer in shim.LAYERS:
        n_ = layer.name
        if isinstance(layer, rough.PLE):
            tower = n_[len("ple_"):]
            for e_, net in enumerate(layer.shared_expert_nets):
                dnn_keys(net, "%s.ple.shared_expert_nets.%d" % (tower, e_))
            for t_ in range(layer.num_tasks):
                for e_, net in enumerate(layer.specific_expert_nets[t_]):
                    dnn_keys(net, "%s.ple.specific_expert_nets.%d.%d" % (tower, t_, e_))
                dnn_keys(layer.gate_nets[t_], "%s.ple.gate_nets.%d" % (tower, t_))
        elif isinstance(layer, rough.DNN) and n_ in ("td_user_emb", "hpld_user_emb", "item_emb"):
            dnn_keys(layer, {"td_user_emb": "user.heads.0", "hpld_user_emb": "user.heads.1", "item_emb": "item.heads.0"}[n_])
        elif isinstance(layer, rough.CrossNet):
            for i_ in range(layer.layer_num):
                ids[id(layer.kernels[i_])], ids[id(layer.bias[i_])] = ("teacher.cross.kernels.%d" % i_, 0.0), ("teacher.cross.bias.%d" % i_, 0.0)
        elif isinstance(layer, shim.Dense) and n_:
            key = {"pred_teacher": "teacher.pred", "shallow_dnn_0": "shallow.shallow_dnn_0", "logit_shallow": "shallow.logit_shallow"}[n_]
            ids[id(layer.kernel)], ids[id(layer.bias)] = (key + ".kernel", 0.0), (key + ".bias", 0.0)
    out["ds_seed"], out["ds_manifest"] = np.asarray(21), manifest(ids)

    # ---- the losses: rank/ctr/base_model.py:7-12, rank/multi_head/model.py:18-22, staytime/model.py:20-60
    base = load("rank/ctr/base_model.py", "ref_rank_ctr_base_model")
    yb = (rng.random((9, 2)) < 0.3).astype(np.float64)
    pb = np.clip(rng.random((9, 2)), 1e-6, 1.0)
    pb[0, 0], pb[1, 1] = 1e-6, 1.0
    out.update(bce_y=yb, bce_p=pb, bce_loss=np.asarray(base.cross_entropy(shim.T(yb), shim.T(pb))))
    stub("video_id_rank_staytime_mtl_ppnet_v7.model.custom_metrics", CustomAccuracy=object, CustomMAE=object, CustomMSE=object)
    stub("video_id_rank_staytime_mtl_ppnet_v7.model.VideoDNN", mtl_net=vd.mtl_net)
    for n_ in ("src.util.util",):
        stub(n_, read_dataset=None, trained_delta_days=None, dump_predict=None)
    smod = load("staytime/model.py", "ref_staytime_model")
    yt = np.zeros((7, 401))
    yt[np.arange(7), rng.integers(0, 400, 7)] = 1.0             # one-hot bins + the stay-time column
    yt[:, 400] = rng.random(7) * 50
    yp = rng.random((7, 400))
    yp = np.concatenate([yp / yp.sum(1, keepdims=True), rng.random((7, 1)) * 50], 1)
    out.update(kl_y=yt, kl_p=yp, kl_loss=np.asarray(smod.custom_kl_loss(shim.T(yt), shim.T(yp))))
    yc, pc = (rng.random((7, 1)) < 0.4).astype(np.float64), rng.random((7, 1))
    out.update(ce_y=yc, ce_p=pc, ce_loss=np.asarray(smod.cross_entropy(shim.T(yc), shim.T(pc))))
    ym, pm = rng.random((7, 1)) * 4, rng.random((7, 1)) * 3
    out.update(mse_y=ym, mse_p=pm, mse_loss=np.asarray(smod.mse_loss(shim.T(ym), shim.T(pm))),
               huber_loss=np.asarray(smod.huber_loss(shim.T(ym), shim.T(pm))))

    # ---- rank/multi_head/model.py::cross_entropy (:18-22): per-sample sum over the 7 labels, keepdims
    stub("interact_multihead_autoint_alllabel.model.config", Config=types.SimpleNamespace(LINEAR_SLOTS=[], DENSE_SLOTS=[]))
    stub("interact_multihead_autoint_alllabel.model.mutiDnnAutointOrigin", AUTOINT=md.AUTOINT)
    mh = load("rank/multi_head/model.py", "ref_multi_head_model")
    y7 = (rng.random((6, 7)) < 0.3).astype(np.float64)
    p7 = np.clip(rng.random((6, 7)), 1e-6, 1.0)
    out.update(ce7_y=y7, ce7_p=p7, ce7_loss=np.asarray(mh.cross_entropy(shim.T(y7), shim.T(p7))))

    # ---- the label transform: staytime/parse.py::parse_input_func (watch time -> short / long play labels, the 400-bin
    # gaussian stay-time label + clipped watch time, landing-page sample weight) on a hand-made parsed example
    pr = load("staytime/parse.py", "ref_staytime_parse")
    watch = np.array([0, 6999, 7000, 7001, 18000, 18001, 65432, 159999, 160000, 160001, 400000, 12345], np.int64)
    extra = np.array(["label", "x_video_homepage_landing_y", "video_homepage_landing", "other", "label", "label",
                      "a video_homepage_landing", "label", "label", "label", "zzz", "label"])
    shim.PARSED.clear()
    shim.PARSED.update(extra_info=shim.T(extra), video_duration=shim.T(watch), watch_duration=shim.T(watch))
    _, y_lab, w_lab = pr.parse_input_func(None)
    pre_ = "video_id_rank_staytime_mtl_ppnet_v7_"
    out.update(lab_watch=watch, lab_landing=np.array(["video_homepage_landing" in e_ for e_ in extra]),
               lab_staytime=np.asarray(y_lab[pre_ + "staytime"]), lab_short=np.asarray(y_lab[pre_ + "shortplay"]),
               lab_long=np.asarray(y_lab[pre_ + "longplay"]), lab_weight=np.asarray(w_lab))

    np.savez_compressed(OUT, **{k: np.asarray(v) for k, v in out.items()})
    print("wrote", OUT, len(out), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
