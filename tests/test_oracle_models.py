"""CPU: the restatement of the two composed dense graphs (oracle/oracle_models.py) evaluated in numpy
float64 and in torch float64 agree, reproduces the reference's documented shapes / invariants, and its
torch twin is differentiable (the gradient oracle of tests/test_gpu_models.py)."""
import numpy as np
import pytest

from util import rel_err
import util_models as um

torch = pytest.importorskip("torch")

SLOTS = sorted(set(um.USER4 + um.ITEM4 + ["3051", "2544", "3376", "3365", "3369", "2597", "2125", "2128", "2130", "1571"]))
SEQ = ["2125", "2128", "2130"]


def _f64(d):
    return {k: np.asarray(v, np.float64) for k, v in d.items()}


def test_video_dnn_np_vs_torch():
    from oracle import oracle_models as om
    rng = np.random.default_rng(3)
    B, T = 6, 7
    P = _f64(um.video_dnn_params(rng, SLOTS, SEQ, units=(32, 16)))
    embs, seqs = um.video_dnn_inputs(rng, B, T, SLOTS, SEQ)
    e64 = _f64(embs)
    s64 = {k: (v[0].astype(np.float64), v[1]) for k, v in seqs.items()}
    out = om.video_dnn_fwd(om.NP, e64, s64, P, SLOTS, SEQ, units=(32, 16))
    tP = {k: torch.from_numpy(v) for k, v in P.items()}
    te = {k: torch.from_numpy(v).requires_grad_(True) for k, v in e64.items()}
    ts = {k: (torch.from_numpy(v[0]), torch.from_numpy(v[1])) for k, v in s64.items()}
    tout = om.video_dnn_fwd(om.TH, te, ts, tP, SLOTS, SEQ, units=(32, 16))
    for k in out:
        assert rel_err(tout[k].detach().numpy(), out[k]) < 1e-12, k
    assert out["staytime"].shape == (B, 401) and out["shortplay"].shape == (B, 1)
    assert np.allclose(out["staytime"][:, :400].sum(-1), 1.0)               # 400-way softmax (VideoDnn.py:170)
    assert np.allclose(out["staytime"][:, 400:], np.maximum(out["staytime"][:, :400] @ np.asarray(om.BIN_LIST)[:, None], 0))
    (tout["staytime"].sum() + tout["shortplay"].sum() + tout["longplay"].sum()).backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in te.values())


def test_video_dnn_concat_width_matches_reference():
    """91 slots, 3 sequences: concated_input is [B,1712] and the PPNet gate input [B,224]."""
    from recommendsystem_b200.api.staytime_config import Config as C
    assert len(C.SLOTS) == 91 and len(C.bin_list) == 400 and C.bin_list[0] == -19.0 and C.bin_list[-1] == 180.5
    n = len(C.SLOTS)
    assert 16 * n + 16 + 64 + 128 + 16 * len(C.SEQ_SLOTS) == 1712
    assert int(n / 4) == 22


def test_dssm_np_vs_torch():
    from oracle import oracle_models as om
    rng = np.random.default_rng(5)
    user_ids, item_ids = ["2597", "2", "1568", "2125"], ["1591", "1593", "2049"]
    B = 9
    P = _f64(um.dssm_params(rng, user_ids, item_ids))
    embs = {k: 0.3 * rng.standard_normal((B, 16)) for k in user_ids + item_ids}
    mask = (rng.random((B, 1)) < 0.5).astype(np.float64)
    out = om.dssm_fwd(om.NP, embs, mask, P, user_ids, item_ids)
    tP = {k: torch.from_numpy(v) for k, v in P.items()}
    te = {k: torch.from_numpy(v).requires_grad_(True) for k, v in embs.items()}
    tout = om.dssm_fwd(om.TH, te, torch.from_numpy(mask), tP, user_ids, item_ids)
    for k in out:
        assert rel_err(tout[k].detach().numpy(), out[k]) < 1e-12, k
    assert out["student"].shape == (B, 1) and out["distill"].shape == (B,)
    # the distillation loss does not back-propagate into the teacher (stop_gradient, model.py:163)
    tout["distill"].sum().backward()
    g_teacher_only = sum(float(te[k].grad.abs().sum()) for k in te)
    assert g_teacher_only > 0


def test_rough_rank_config_ids():
    from recommendsystem_b200.api.rough_rank_model import config as C
    assert len(C.USER_FEATURE_IDS) == 33 and len(C.ITEM_FEATURE_IDS) == 19 and len(C.ALL_FEATURE_ID_2_SLOT) == 52
    assert C.get_feature_id("2597") == "2597"
    with pytest.raises(ValueError, match="feature: nope not found"):
        C.get_feature_id("nope")


def test_autoint_multihead_oracle_np_matches_torch():
    """The two statements of create_autoint_sub_model (numpy forward oracle, torch autograd gradient oracle) agree
    to round-off, with and without the counter-based attention dropout."""
    import torch
    from oracle import oracle_models as om
    rng = np.random.default_rng(5)
    B, F = 9, 11
    embs = [0.5 * rng.standard_normal((B, 8)) for _ in range(F)]
    P = {}
    for nm in ("query", "key", "value", "res"):
        P["interacting_layer.%s_dense_kernel" % nm] = rng.standard_normal((8, 8)) * 0.4
        P["interacting_layer.%s_dense_bias" % nm] = rng.standard_normal(8) * 0.1
    P["interacting_layer.layer_norm_gamma"] = 1 + 0.1 * rng.standard_normal(8)
    P["interacting_layer.layer_norm_beta"] = 0.1 * rng.standard_normal(8)
    width = 8 * F
    for i, u in enumerate((32, 16)):
        P["dnn_%d.kernel" % i] = rng.standard_normal((width, u)) / np.sqrt(width)
        P["dnn_%d.bias" % i] = 0.1 * rng.standard_normal(u)
        width = u
    cat = 16 + 8 * F
    for i in range(8):
        P["expert_%d_fc1.kernel" % i] = rng.standard_normal((cat, 32)) / np.sqrt(cat)
        P["expert_%d_fc1.bias" % i] = 0.1 * rng.standard_normal(32)
    for i in range(7):
        P["gate_%d_fc2.kernel" % i] = rng.standard_normal((cat, 7)) / np.sqrt(cat)
        P["gate_%d_fc2.bias" % i] = 0.1 * rng.standard_normal(7)
    for lab in om.AUTOINT_LABELS:
        P[lab + ".kernel"] = rng.standard_normal((32, 1)) * 0.3
        P[lab + ".bias"] = 0.1 * rng.standard_normal(1)
    for dropout in (None, (0.2, 12345)):
        a = om.autoint_multihead_fwd(om.NP, embs, P, (32, 16), dropout)
        b = om.autoint_multihead_fwd(om.TH, [torch.from_numpy(e) for e in embs], {k: torch.from_numpy(v) for k, v in P.items()},
                                     (32, 16), dropout)
        assert a.shape == (B, 7) and np.all((a > 0) & (a < 1))
        assert np.max(np.abs(a - b.numpy())) < 1e-12
    assert np.max(np.abs(om.autoint_multihead_fwd(om.NP, embs, P, (32, 16), (0.2, 1)) - a)) > 1e-6   # the mask matters
