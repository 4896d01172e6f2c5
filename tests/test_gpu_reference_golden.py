"""The CUDA path against what the REFERENCE'S OWN CODE computed: tests/golden/ref_layers.npz holds the outputs of the
reference's layer files and model builders executed unmodified on seeded inputs (tools/gen_reference_layer_golden.py,
`tensorflow` replaced by the numpy stand-in oracle/tf_numpy_shim.py).  Here the kernels (through the C-ABI) and the
drop-in modules are fed the same inputs and weights (by Keras layer name) — no oracle in between.  fp32, 1e-5."""
import json
import os

import numpy as np
import pytest

from util import REL_F32, assert_close

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_layers.npz"))


def weights(prefix):
    from oracle.tf_numpy_shim import replay_weights
    return replay_weights(int(G[prefix + "_seed"]), json.loads(str(G[prefix + "_manifest"])))


def dev(a, d, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(d, dtype)


def load(module, P, d):
    sd = module.state_dict()
    diff = set(k for k in sd if k != "wt_bins") ^ set(P)
    assert not diff, sorted(diff)[:10]
    module.load_state_dict({**{k: dev(v, d).reshape(sd[k].shape) for k, v in P.items()},
                            **({"wt_bins": sd["wt_bins"]} if "wt_bins" in sd else {})})


@pytest.mark.parametrize("tag", ["cfg1", "rankctr", "nores", "h1"])
def test_interacting_layer_kernel_vs_reference_code(cuda_dev, tag):
    """rs_interacting_fwd (fp32 kernels) on the inputs / weights of InteractingLayer.call as the reference ran it."""
    from recommendsystem_b200 import ops
    H, L, res = (int(v) for v in G[f"inter_{tag}_cfg"])
    y, _ = ops.interacting_fwd(dev(G[f"inter_{tag}_x"], cuda_dev), dev(G[f"inter_{tag}_W"], cuda_dev),
                               dev(G[f"inter_{tag}_b"], cuda_dev), dev(G[f"inter_{tag}_gamma"], cuda_dev),
                               dev(G[f"inter_{tag}_beta"], cuda_dev), float(G[f"inter_{tag}_eps"]), H, L, bool(res))
    assert_close(y.cpu().numpy(), G[f"inter_{tag}_y"], REL_F32, f"InteractingLayer {tag}")


def test_din_kernels_vs_reference_code(cuda_dev):
    """rs_din_fwd, both variants, on din.py::DIN.call / staytime/layer.py::DIN.call as the reference ran them."""
    from recommendsystem_b200 import cabi, ops
    d = cuda_dev
    ya = ops.din_fwd(cabi.DIN_A, dev(G["dina_q"], d), dev(G["dina_keys"], d), dev(G["dina_values"], d),
                     dev(G["dina_seq_len"], d, torch.int32), None, dev(G["dina_W1"], d), dev(G["dina_b1"], d),
                     dev(G["dina_W2"], d), dev(G["dina_b2"], d))
    assert_close(ya.cpu().numpy(), G["dina_y"], REL_F32, "DIN A")
    T = G["dinb_facts"].shape[1]
    yb = ops.din_fwd(cabi.DIN_B, dev(G["dinb_q"], d), dev(G["dinb_facts"], d), None, None,
                     dev(G["dinb_mask"][:, :T].astype(np.uint8), d, torch.uint8), dev(G["dinb_W1"], d), dev(G["dinb_b1"], d),
                     dev(G["dinb_W2"], d), dev(G["dinb_b2"], d))
    assert_close(yb.cpu().numpy(), G["dinb_y"], REL_F32, "DIN B")


def test_cross_kernels_and_labels_vs_reference_code(cuda_dev):
    """rs_cross_fwd on DeepCrossLayer.call / CrossNet.call, rs_staytime_labels on parse_input_func."""
    from oracle import oracle_metrics as omet
    from recommendsystem_b200 import ops
    d = cuda_dev
    y = ops.cross_fwd(dev(G["dcross_x"], d), dev(G["dcross_W"][:, :, 0], d), dev(G["dcross_b"], d))
    assert_close(y.cpu().numpy(), G["dcross_y"], REL_F32, "DeepCrossLayer")
    y = ops.cross_fwd(dev(G["cnet_x"], d), dev(G["cnet_k"][:, :, 0], d), dev(G["cnet_b"][:, :, 0], d))
    assert_close(y.cpu().numpy(), G["cnet_y"], REL_F32, "CrossNet")
    bins = torch.tensor(omet.BIN_LIST, dtype=torch.float32, device=d)
    lab, sh, lo, w = ops.staytime_labels(dev(G["lab_watch"], d, torch.int64), bins, dev(G["lab_landing"], d, torch.uint8))
    assert_close(lab.cpu().numpy(), G["lab_staytime"], REL_F32, "stay-time label")
    assert np.array_equal(sh.cpu().numpy(), G["lab_short"]) and np.array_equal(lo.cpu().numpy(), G["lab_long"])
    assert np.array_equal(w.cpu().numpy().reshape(-1), G["lab_weight"].reshape(-1))


def test_video_dnn_module_vs_reference_code(cuda_dev):
    """api.video_dnn.VideoDnnSubModel with the weights of the reference's create_moe_sub_model run (by Keras layer name)."""
    from recommendsystem_b200.api.video_dnn import TASK_KEYS, VideoDnnSubModel
    d = cuda_dev
    slots, seq = [str(s) for s in G["vd_slots"]], [str(s) for s in G["vd_seq_slots"]]
    model = VideoDnnSubModel(slots, seq, (16, 8)).to(d)
    embs = {s: dev(G[f"vd_emb_{s}"], d) for s in slots}
    seqs = {s: (dev(G[f"vd_seq_{s}"], d), dev(G[f"vd_mask_{s}"], d, torch.bool)) for s in seq}
    with torch.no_grad():
        model(embs, seqs)                                       # lazy build
        load(model, weights("vd"), d)
        train, predict = model(embs, seqs)
    pre = "video_id_rank_staytime_mtl_ppnet_v7_"
    assert_close(train[TASK_KEYS[0]].cpu().numpy(), G["vd_train_" + pre + "staytime"], REL_F32, "stay-time [B,401]")
    assert_close(predict[TASK_KEYS[0]].cpu().numpy(), G["vd_predict_" + pre + "staytime"], REL_F32, "stay-time prediction")
    assert_close(train[TASK_KEYS[1]].cpu().numpy(), G["vd_train_" + pre + "shortplay"], REL_F32, "shortplay")
    assert_close(train[TASK_KEYS[2]].cpu().numpy(), G["vd_train_" + pre + "longplay"], REL_F32, "longplay")


def test_autoint_multihead_module_vs_reference_code(cuda_dev):
    """api.builders.AUTOINT's sub-model (inference: no attention dropout) on the reference's create_autoint_sub_model run."""
    from recommendsystem_b200.api.builders import AUTOINT
    d = cuda_dev
    F = G["ai_embs"].shape[0]
    model = AUTOINT([str(3000 + i) for i in range(F)], ["d0"], False, dnn_hidden_units=(32, 16), bucket_size=100,
                    device=d).sub_model.eval()
    embs = [dev(e, d) for e in G["ai_embs"]]
    with torch.no_grad():
        model(embs)
        load(model, weights("ai"), d)
        y = model(embs)
    assert_close(y.cpu().numpy(), G["ai_y"], REL_F32, "AUTOINT [B,7]")


def test_rank_ctr_module_vs_reference_code(cuda_dev):
    """api.rank_ctr.RankCtrSubModel on the reference's BaseModel.__init__ + Model.model_layer run."""
    from recommendsystem_b200.api.rank_ctr import TASK_NAMES, RankCtrSubModel, parse_feature_slots
    d = cuda_dev
    cfg = json.loads(str(G["rc_config"]))
    model = RankCtrSubModel(parse_feature_slots(cfg)).to(d).eval()
    emb = {k[len("rc_emb_"):]: dev(G[k], d) for k in G.files if k.startswith("rc_emb_")}
    with torch.no_grad():
        model(emb)
        load(model, weights("rc"), d)
        out = model(emb)
    for i, t in enumerate(TASK_NAMES):
        assert_close(out[t].cpu().numpy(), G["rc_task%d" % i], REL_F32, t)


def test_dssm_module_vs_reference_code(cuda_dev):
    """api.rough_rank_model.DssmSubModel on the reference's DSSM() run; the product's resolved feature lists are the
    ones rough_rank/config evaluates to."""
    from recommendsystem_b200.api.rough_rank_model import DssmSubModel, config as C
    d = cuda_dev
    uid, iid = [str(v) for v in G["ds_user_ids"]], [str(v) for v in G["ds_item_ids"]]
    assert list(C.USER_FEATURE_IDS) == uid and list(C.ITEM_FEATURE_IDS) == iid
    model = DssmSubModel(C.USER_FEATURE_IDS, C.ITEM_FEATURE_IDS).to(d)
    embs = {k: dev(G["ds_emb_" + k], d) for k in uid + iid}
    mask = dev(G["ds_mask"], d)
    with torch.no_grad():
        model(embs, mask)
        load(model, weights("ds"), d)
        out = model(embs, mask)
    assert_close(out["student"].cpu().numpy(), G["ds_student"], REL_F32, "student")
    assert_close(out["teacher"].cpu().numpy(), G["ds_teacher"], REL_F32, "teacher")
    assert_close(out["distill"].cpu().numpy().reshape(-1), G["ds_distill"].reshape(-1), REL_F32, "distill")
