"""Pins the oracle to the reference's own code: tests/golden/ref_layers.npz holds what the reference's layer files
(InteractingLayer.py, din.py, staytime/layer.py, rough_rank/layer.py) RETURNED when executed unmodified on seeded fp64
inputs / weights, with `tensorflow` replaced by the numpy stand-in oracle/tf_numpy_shim.py
(tools/gen_reference_layer_golden.py, run where /root/reference exists).  oracle/oracle_np.py and oracle/oracle_models.py
— the checkers of every GPU parity test — must reproduce those outputs to fp64 round-off."""
import os

import numpy as np
import pytest

from oracle import oracle_models as om
from oracle import oracle_np as onp

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_layers.npz"))
TOL = 1e-12


def weights(prefix):
    """The weights the reference's layers drew in that run: re-drawn from (seed, creation-order manifest) instead of
    being stored (oracle/tf_numpy_shim.py::replay_weights), keyed by the oracle's parameter names."""
    import json
    from oracle.tf_numpy_shim import replay_weights
    return replay_weights(int(G[prefix + "_seed"]), json.loads(str(G[prefix + "_manifest"])))


def close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.max(np.abs(got - want) / (1.0 + np.abs(want)))
    assert err <= TOL, (what, err)


@pytest.mark.parametrize("tag", ["cfg1", "rankctr", "nores", "h1"])
def test_interacting_layer_matches_reference_code(tag):
    """InteractingLayer.call (InteractingLayer.py:37-61): head split / concat order, 1/sqrt(d_head) scaling, softmax,
    residual, relu, LayerNorm, the SAME four Dense layers re-applied layer_num times; with and without use_res."""
    H, L, res = (int(v) for v in G[f"inter_{tag}_cfg"])
    y = onp.interacting_fwd(G[f"inter_{tag}_x"], G[f"inter_{tag}_W"], G[f"inter_{tag}_b"], G[f"inter_{tag}_gamma"],
                            G[f"inter_{tag}_beta"], float(G[f"inter_{tag}_eps"]), H, L, use_res=bool(res))
    close(y, G[f"inter_{tag}_y"], f"InteractingLayer {tag}")


def test_interacting_layer_models_restatement_matches_reference_code():
    """The second restatement (oracle_models._interacting, used by the composed-model oracles) against the same run."""
    tag = "rankctr"
    H, L, _ = (int(v) for v in G[f"inter_{tag}_cfg"])
    W, b = G[f"inter_{tag}_W"], G[f"inter_{tag}_b"]
    U = W.shape[1] // 4
    P = {}
    for j, n in enumerate(["query", "key", "value", "res"]):
        P[f"il.{n}_dense_kernel"], P[f"il.{n}_dense_bias"] = W[:, j * U:(j + 1) * U], b[j * U:(j + 1) * U]
    P["il.layer_norm_gamma"], P["il.layer_norm_beta"] = G[f"inter_{tag}_gamma"], G[f"inter_{tag}_beta"]
    try:
        y = om._interacting(om.NP, G[f"inter_{tag}_x"], P, "il", H, L, eps=float(G[f"inter_{tag}_eps"]))
    except KeyError as e:       # parameter naming of the restatement differs: report it instead of passing silently
        pytest.fail(f"oracle_models._interacting parameter names changed: {e}")
    close(y, G[f"inter_{tag}_y"], "oracle_models._interacting")


def test_din_a_matches_reference_code():
    """din.py:18-47 incl. tf.sequence_mask lengths 0 / 1 / T."""
    y = onp.din_a_fwd(G["dina_q"], G["dina_keys"], G["dina_values"], G["dina_seq_len"], G["dina_W1"], G["dina_b1"],
                      G["dina_W2"], G["dina_b2"])
    close(y, G["dina_y"], "DIN A")


def test_din_b_matches_reference_code():
    """staytime/layer.py:16-41 incl. a mask wider than the sequence (tf.slice) and a fully masked row."""
    T = G["dinb_facts"].shape[1]
    y = onp.din_b_fwd(G["dinb_q"], G["dinb_facts"], G["dinb_mask"][:, :T], G["dinb_W1"], G["dinb_b1"], G["dinb_W2"],
                      G["dinb_b2"])
    close(y, G["dinb_y"], "DIN B (oracle_np)")
    P = {"d.layer_1_kernel": G["dinb_W1"], "d.layer_1_bias": G["dinb_b1"], "d.layer_2_kernel": G["dinb_W2"],
         "d.layer_2_bias": G["dinb_b2"]}
    close(om.din_b(om.NP, G["dinb_q"], G["dinb_facts"], G["dinb_mask"], P, "d"), G["dinb_y"], "DIN B (oracle_models)")


def test_cross_layers_match_reference_code():
    """DeepCrossLayer.call (staytime/layer.py:66-72) and CrossNet.call (rough_rank/layer.py:256-264)."""
    P = {}
    for i in range(3):
        P[f"c.W.{i}"], P[f"c.b.{i}"] = G["dcross_W"][i], G["dcross_b"][i]
    close(om.deep_cross(om.NP, G["dcross_x"], P, "c", 3), G["dcross_y"], "DeepCrossLayer")
    P = {}
    for i in range(2):
        P[f"n.kernels.{i}"], P[f"n.bias.{i}"] = G["cnet_k"][i], G["cnet_b"][i]
    close(om.cross_net(om.NP, G["cnet_x"], P, "n", 2), G["cnet_y"], "CrossNet")


def test_fm_and_dnn_match_reference_code():
    """FMLayer.call (staytime/layer.py:100-111) and DNN.call (rough_rank/layer.py:100-109, output_activation)."""
    x = G["fm_x"]
    s = x.sum(1)
    close(0.5 * (s * s - (x * x).sum(1)).sum(-1, keepdims=True), G["fm_y"], "FMLayer")
    h = G["dnn_x"]
    for i, act in enumerate(["relu", "relu", "sigmoid"]):
        h = onp.dense(h, G[f"dnn_k{i}"], G[f"dnn_b{i}"], act)
    close(h, G["dnn_y"], "DNN")


def test_video_dnn_sub_model_matches_reference_code():
    """The whole dense graph of BASELINE configs[4]: staytime/VideoDnn.py::create_moe_sub_model (DIN x3, SENet on a
    stop-gradient copy, FM, FFM, PPNet-gated experts, MMoE gates, DeepCross, 400-way stay-time head, two towers) as the
    reference's own Keras functional code computed it on seeded inputs — against oracle_models.video_dnn_fwd, the checker
    of the VideoDnn GPU parity tests.  18 slots (every user / item / bias slot), T = 5, units (16, 8)."""
    slots, seq_slots = [str(s) for s in G["vd_slots"]], [str(s) for s in G["vd_seq_slots"]]
    embs = {s: G[f"vd_emb_{s}"] for s in slots}
    seqs = {s: (G[f"vd_seq_{s}"], G[f"vd_mask_{s}"]) for s in seq_slots}
    P = weights("vd")
    out = om.video_dnn_fwd(om.NP, embs, seqs, P, slots, seq_slots, units=(16, 8))
    pre = "video_id_rank_staytime_mtl_ppnet_v7_"
    close(out["staytime"], G["vd_train_" + pre + "staytime"], "VideoDnn stay-time distribution + expectation")
    close(out["staytime_pred"], G["vd_predict_" + pre + "staytime"], "VideoDnn stay-time prediction")
    close(out["shortplay"], G["vd_train_" + pre + "shortplay"], "VideoDnn shortplay")
    close(out["longplay"], G["vd_train_" + pre + "longplay"], "VideoDnn longplay")
    p = np.asarray(G["vd_train_" + pre + "staytime"])[:, :400]
    assert p.max() < 0.9 and p.min() > 1e-12          # the fixture's softmax is not saturated


def test_autoint_multihead_sub_model_matches_reference_code():
    """BASELINE configs[3]: rank/multi_head/multidnn.py::create_autoint_sub_model (its own InteractingLayer copy, the
    DNN, 8 experts of which the first 7 are mixed, 7 softmax gates, 7 sigmoid heads in MultiLabelInfo.label_list order)
    as the reference's code computed it (dropout layers at inference) — against oracle_models.autoint_multihead_fwd."""
    assert [str(v) for v in G["ai_labels"]] == om.AUTOINT_LABELS
    P = weights("ai")
    y = om.autoint_multihead_fwd(om.NP, list(G["ai_embs"]), P, deep_hidden_units=(32, 16), dropout=None,
                                 eps=float(G["ai_eps"]))
    close(y, G["ai_y"], "AUTOINT sub-model")
    assert "expert_7_fc1.kernel" in P          # the eighth expert the reference builds and never uses (:80-92)


def test_ple_matches_reference_code():
    """PLE.call (rough_rank/layer.py:211-224) as create_tower builds it (2 tasks, 4 shared + 4 specific experts of one
    Dense(32, relu) each, softmax gates over the 8 experts) — against oracle_models.ple (the DSSM oracle's towers)."""
    P = {k[len("ple_P_"):]: G[k] for k in G.files if k.startswith("ple_P_")}
    ys = om.ple(om.NP, G["ple_x"], P, "p", 2, 4, 4)
    close(np.stack(ys), G["ple_y"], "PLE")


def test_losses_match_reference_code():
    """cross_entropy (rank/ctr/base_model.py:7-12) against oracle_np.bce_loss (the checker of the fused head kernel), and
    the product's own stay-time losses (api/video_dnn.py, plain torch: they run on CPU tensors) against
    staytime/model.py:20-60 as the reference's code evaluated them."""
    loss, _ = onp.bce_loss(G["bce_p"], G["bce_y"])
    close(loss, G["bce_loss"], "cross_entropy (rank/ctr)")
    torch = pytest.importorskip("torch")
    from recommendsystem_b200.api import video_dnn as V
    t = lambda k: torch.from_numpy(np.asarray(G[k], np.float64))
    close(V.custom_kl_loss(t("kl_y"), t("kl_p")).numpy(), G["kl_loss"], "custom_kl_loss")
    ce = -t("ce_y") * torch.log(t("ce_p") + 1e-6) - (1 - t("ce_y")) * torch.log(1.0 - t("ce_p") + 1e-6)
    close(ce.numpy(), G["ce_loss"], "cross_entropy (staytime) formula")
    got = V.cross_entropy(t("ce_y"), t("ce_p")).numpy()           # the product casts labels to fp32 like the reference
    assert np.max(np.abs(got - G["ce_loss"])) <= 1e-6
    assert abs(float(V.mse_loss(t("mse_y"), t("mse_p"))) - float(G["mse_loss"])) <= 1e-6       # labels cast to fp32
    close(V.huber_loss(t("mse_y"), t("mse_p")).numpy(), G["huber_loss"], "huber_loss")
    # the 7-label loss of rank/multi_head/model.py:18-22 (sum over the labels, keepdims) and its batch-mean form of
    # rank/ctr/base_model.py:7-12, as api.builders.cross_entropy (what AUTOINT.train_step / RankCtrNet minimise)
    from recommendsystem_b200.api.builders import cross_entropy
    close(cross_entropy(t("ce7_y"), t("ce7_p")).numpy(), G["ce7_loss"], "cross_entropy (multi_head)")
    close(float(cross_entropy(t("bce_y"), t("bce_p"), reduce_mean=True)), G["bce_loss"], "cross_entropy (rank/ctr form)")


def test_rank_ctr_production_model_matches_reference_code():
    """rank/ctr: base_model.py::BaseModel.__init__ (SingleSlot slicing of the config's features into structure / bias /
    gate columns) + model_init.py::Model.model_layer (SENet on a stop-gradient copy, one linear map per field,
    InteractingLayer, PPNet gates split 256|64|8|256|64|8|32|16, CAN co-action with per-sample 8x6 and 6x4 matrices,
    gated experts, MMoE, two towers, clip) executed as they are on the synthetic config of the GPU parity test — against
    oracle_models.rank_ctr_layout + rank_ctr_fwd."""
    import json
    cfg = json.loads(str(G["rc_config"]))
    me, st, b, gt = om.rank_ctr_layout(cfg)
    emb = {k[len("rc_emb_"):]: G[k] for k in G.files if k.startswith("rc_emb_")}
    assert all(v.shape[1] == me for v in emb.values())
    P = weights("rc")
    out = om.rank_ctr_fwd(om.NP, emb, P, st, b, gt)
    close(out["task0"], G["rc_task0"], "rank/ctr click")
    close(out["task1"], G["rc_task1"], "rank/ctr effect_click")
    assert 1e-3 < float(np.min(G["rc_task0"])) and float(np.max(G["rc_task0"])) < 1 - 1e-3      # not saturated / clipped


def test_headline_autoint_model_matches_reference_code():
    """THE benchmarked model (BASELINE configs[0] / [1]): /root/reference/autoint::AutoInt.model_layer — 39 fields x 16
    stacked on axis 1, InteractingLayer(layer_num=3, unit_num=16, head_num=2), Flatten, DNN 256-128 on the flattened
    fields, concat [deep | autoint], Dense(1, sigmoid), clip_by_value(1e-6, 1) — executed on BaseModel's own field list,
    against oracle_np.autoint_fwd_bwd (the checker of the trainer's parity tests and of smoke())."""
    W = weights("hl")
    P = {"Wqkvr": np.concatenate([W[n + "_kernel"] for n in ("query", "key", "value", "res")], 1),
         "bqkvr": np.concatenate([W[n + "_bias"] for n in ("query", "key", "value", "res")]),
         "gamma": W["gamma"], "beta": W["beta"], "mlp_W": [W["mlp_W0"], W["mlp_W1"]], "mlp_b": [W["mlp_b0"], W["mlp_b1"]],
         "out_W": W["out_W"], "out_b": W["out_b"]}
    X = G["hl_X"]
    y = np.zeros((X.shape[0], 1))
    res = onp.autoint_fwd_bwd(X, P, y, 2, 3, float(G["hl_eps"]))
    p_raw = res["p_raw"] if "p_raw" in res else res["logits"]
    close(np.clip(p_raw, 1e-6, 1.0), G["hl_p"], "AutoInt.model_layer")


def test_dssm_matches_reference_code():
    """rough_rank/model.py::DSSM on the reference's own feature lists (rough_rank/config): user tower (PLE with two tasks,
    one of the two heads selected per sample by the dense feature 4575), item tower, teacher (CrossNet + DNN), the
    shallow student tower on the two tower embeddings and the distillation loss — against oracle_models.dssm_fwd."""
    user_ids, item_ids = [str(v) for v in G["ds_user_ids"]], [str(v) for v in G["ds_item_ids"]]
    embs = {k: G["ds_emb_" + k] for k in user_ids + item_ids}
    out = om.dssm_fwd(om.NP, embs, G["ds_mask"], weights("ds"), user_ids, item_ids)
    close(out["user_emb"], G["ds_user_emb"], "DSSM user tower")
    close(out["item_emb"], G["ds_item_emb"], "DSSM item tower")
    close(out["teacher"], G["ds_teacher"], "DSSM teacher")
    close(out["student"], G["ds_student"], "DSSM student")
    close(out["distill"], G["ds_distill"], "DSSM distillation loss")
    sel = G["ds_mask"].ravel() == 1
    assert sel.any() and (~sel).any()                 # both user heads are exercised
    # the product's resolved feature lists (api/rough_rank_model.py::config) are what rough_rank/config evaluates to
    from recommendsystem_b200.api.rough_rank_model import config as C
    assert list(C.USER_FEATURE_IDS) == user_ids and list(C.ITEM_FEATURE_IDS) == item_ids


def test_staytime_labels_match_reference_code():
    """staytime/parse.py::parse_input_func (:30-68) executed on a hand-made parsed example: short / long play thresholds
    at 7000 / 18000 ms (strict), the 160 s cap, the 400-bin gaussian label scaled by the bin width, the clipped watch
    time in column 400 and the x5 landing-page sample weight — against oracle_metrics.staytime_labels (fp64), the
    checker of rs_staytime_labels."""
    from oracle import oracle_metrics as omet
    lab, sh, lo, w = omet.staytime_labels(G["lab_watch"], G["lab_landing"].astype(np.int64), dtype=np.float64)
    close(lab, G["lab_staytime"], "stay-time label")
    assert np.array_equal(sh, G["lab_short"]) and np.array_equal(lo, G["lab_long"])
    close(w, G["lab_weight"], "sample weight")


def test_reference_arm_model_matches_reference_code():
    """oracle_torch.AutoIntCPU — the model that `bench.py --impl reference` and `cpu_baseline` TIME on the host cores —
    computes what the reference's own autoint::AutoInt.model_layer computed (and its loss is the reference's
    cross_entropy): the baseline the speed-up is quoted against is the reference's graph, not a lighter stand-in."""
    torch = pytest.importorskip("torch")
    from oracle import oracle_torch as ot
    W = weights("hl")
    t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    params = {"Wqkvr": t(np.concatenate([W[n + "_kernel"] for n in ("query", "key", "value", "res")], 1)),
              "bqkvr": t(np.concatenate([W[n + "_bias"] for n in ("query", "key", "value", "res")])),
              "gamma": t(W["gamma"]), "beta": t(W["beta"]), "mlp_W0": t(W["mlp_W0"]), "mlp_b0": t(W["mlp_b0"]),
              "mlp_W1": t(W["mlp_W1"]), "mlp_b1": t(W["mlp_b1"]), "out_W": t(W["out_W"]), "out_b": t(W["out_b"])}
    X = t(G["hl_X"])
    model = ot.AutoIntCPU(torch.zeros(1, X.shape[-1], dtype=torch.float64), params, 2, 3, float(G["hl_eps"]))
    close(model.forward(X).detach().numpy(), G["hl_p"], "AutoIntCPU.forward")
    close(float(ot.cross_entropy(t(G["bce_y"]), t(G["bce_p"]))), G["bce_loss"], "oracle_torch.cross_entropy")


def test_oracle_gradients_are_the_derivative_of_the_pinned_forward():
    """The reference has no backward code (TensorFlow differentiates its graph); the oracle's analytic InteractingLayer
    backward — the checker of the CUDA backward kernels — must therefore be THE derivative of the forward that is pinned
    above: central differences of the pinned forward along random directions in x and in every parameter, fp64, on the
    reference-run configuration (39 fields x 16, 3 layers, 2 heads)."""
    tag = "cfg1"
    H, L, res = (int(v) for v in G[f"inter_{tag}_cfg"])
    x, W, b = G[f"inter_{tag}_x"], G[f"inter_{tag}_W"], G[f"inter_{tag}_b"]
    gamma, beta, eps = G[f"inter_{tag}_gamma"], G[f"inter_{tag}_beta"], float(G[f"inter_{tag}_eps"])
    rng = np.random.default_rng(5)
    dy = rng.standard_normal(G[f"inter_{tag}_y"].shape)
    f = lambda x_, W_, b_, g_, bt_: float((onp.interacting_fwd(x_, W_, b_, g_, bt_, eps, H, L, use_res=bool(res)) * dy).sum())
    dx, dW, db, dg, dbt = onp.interacting_bwd(x, W, b, gamma, beta, eps, H, L, dy, use_res=bool(res))
    args, grads = [x, W, b, gamma, beta], [dx, dW, db, dg, dbt]
    h = 1e-6
    for i, (a, g) in enumerate(zip(args, grads)):
        v = rng.standard_normal(a.shape)
        plus = [q + h * v if j == i else q for j, q in enumerate(args)]
        minus = [q - h * v if j == i else q for j, q in enumerate(args)]
        num = (f(*plus) - f(*minus)) / (2 * h)
        ana = float((np.asarray(g).reshape(a.shape) * v).sum())
        assert abs(num - ana) <= 1e-6 * max(1.0, abs(ana)), (i, num, ana)


def test_gradient_oracles_forward_matches_reference_code():
    """The GPU tests take their GRADIENT references from torch autograd through the same model restatements evaluated in
    the torch namespace (oracle_models.TH, fp64).  Their forward — the function autograd differentiates — reproduces the
    reference-code outputs too, for all four composed graphs."""
    torch = pytest.importorskip("torch")
    import json
    t = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    tP = lambda prefix: {k: t(v) for k, v in weights(prefix).items()}
    pre = "video_id_rank_staytime_mtl_ppnet_v7_"
    slots, seq_slots = [str(s) for s in G["vd_slots"]], [str(s) for s in G["vd_seq_slots"]]
    out = om.video_dnn_fwd(om.TH, {s: t(G[f"vd_emb_{s}"]) for s in slots},
                           {s: (t(G[f"vd_seq_{s}"]), torch.from_numpy(G[f"vd_mask_{s}"])) for s in seq_slots}, tP("vd"),
                           slots, seq_slots, units=(16, 8))
    close(out["staytime"].numpy(), G["vd_train_" + pre + "staytime"], "VideoDnn (torch)")
    close(out["longplay"].numpy(), G["vd_train_" + pre + "longplay"], "VideoDnn longplay (torch)")
    y = om.autoint_multihead_fwd(om.TH, [t(e) for e in G["ai_embs"]], tP("ai"), (32, 16), None, eps=float(G["ai_eps"]))
    close(y.numpy(), G["ai_y"], "AUTOINT (torch)")
    cfg = json.loads(str(G["rc_config"]))
    me, st, b, gt = om.rank_ctr_layout(cfg)
    emb = {k[len("rc_emb_"):]: t(G[k]) for k in G.files if k.startswith("rc_emb_")}
    out = om.rank_ctr_fwd(om.TH, emb, tP("rc"), st, b, gt)
    close(out["task0"].numpy(), G["rc_task0"], "rank/ctr (torch)")
    uid, iid = [str(v) for v in G["ds_user_ids"]], [str(v) for v in G["ds_item_ids"]]
    out = om.dssm_fwd(om.TH, {k: t(G["ds_emb_" + k]) for k in uid + iid}, t(G["ds_mask"]), tP("ds"), uid, iid)
    close(out["student"].numpy(), G["ds_student"], "DSSM student (torch)")
    close(out["distill"].numpy(), G["ds_distill"], "DSSM distill (torch)")
