"""rank/ctr production model.

CPU: the JSON-driven slot slicing is pinned against tests/golden/rank_ctr_layout.json — the layout the
REFERENCE's own SingleSlot / BaseModel.__init__ code computes for its shipped model_parameter.json
(tools/gen_rank_ctr_golden.py runs that code with tensorflow / tensornet stubbed) — for both the product parser
(api.rank_ctr.parse_feature_slots) and the oracle restatement; the dense-graph restatement agrees between
numpy and torch.  GPU: RankCtrSubModel against the oracle, forward and gradients."""
import json
import os

import numpy as np
import pytest

from util import REL_F32, assert_close, rel_err
import util_models as um

torch = pytest.importorskip("torch")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rank_ctr_layout.json")


def _golden_config():
    g = json.load(open(GOLD))
    sparse = {}
    for name, slots, emb, btype, bias_without_type in g["input"]["sparse_feature"]:
        ft = {"emb_size": emb, "slot_id": slots}
        if btype is not None or bias_without_type:
            ft["bias"] = True
        if btype is not None:
            ft["bias_type"] = btype
        sparse[name] = ft
    cfg = {"feature_slot": {"sparse_feature": sparse,
                            "sequence_feature": {n: {"emb_size": e, "slot_id": s} for n, s, e in g["input"]["sequence_feature"]},
                            "dense_feature": {n: {"slot_id": s} for n, s in g["input"]["dense_feature"]}}}
    return g, cfg


def test_slot_layout_matches_reference_golden():
    from oracle import oracle_models as om
    from recommendsystem_b200.api.rank_ctr import parse_feature_slots
    g, cfg = _golden_config()
    lay = parse_feature_slots(cfg)
    assert lay.max_embed_size == g["max_embed_size"] == 96
    assert lay.sparse_slots == g["sparse_slots"] and len(lay.sparse_slots) == 176
    assert [list(s) for s in lay.structure] == g["structure"] and len(lay.structure) == 175
    assert {k: [list(s) for s in v] for k, v in lay.bias.items()} == g["bias"]
    assert {k: len(v) for k, v in lay.bias.items()} == {"multiply_user": 6, "ppnet": 14, "can": 14, "multiply_item": 3}
    assert [list(s) for s in lay.gate] == g["gate"]
    me, st, b, gt = om.rank_ctr_layout(cfg)
    assert (me, st, b, gt) == (g["max_embed_size"], g["structure"], g["bias"], g["gate"])


def test_slot_layout_errors_and_synthetic_config():
    from oracle import oracle_models as om
    from recommendsystem_b200.api.rank_ctr import parse_feature_slots
    cfg = um.rank_ctr_config(np.random.default_rng(0))
    lay = parse_feature_slots(cfg)
    me, st, b, gt = om.rank_ctr_layout(cfg)
    assert lay.max_embed_size == me and [list(s) for s in lay.structure] == st and [list(s) for s in lay.gate] == gt
    assert {k: [list(s) for s in v] for k, v in lay.bias.items()} == b
    bad = {"feature_slot": {"sparse_feature": {"x": {"emb_size": 8, "bias": True, "slot_id": ["1"]}},
                            "sequence_feature": {}, "dense_feature": {}}}
    with pytest.raises(Exception, match="bias_type could not be null"):
        parse_feature_slots(bad)
    dup = {"feature_slot": {"sparse_feature": {"x": {"emb_size": 8, "slot_id": ["1"]}},
                            "sequence_feature": {"s": {"emb_size": 8, "slot_id": ["1"]}}, "dense_feature": {}}}
    with pytest.raises(Exception, match="has been defined more than once"):
        parse_feature_slots(dup)


def _setup(rng, B):
    from oracle import oracle_models as om
    cfg = um.rank_ctr_config(rng)
    me, st, b, gt = om.rank_ctr_layout(cfg)
    P = um.rank_ctr_params(rng, st, b, gt)
    slots = sorted({s[0] for s in st} | {s[0] for v in b.values() for s in v})
    emb = {s: (0.3 * rng.standard_normal((B, me))).astype(np.float32) for s in slots}
    return cfg, (me, st, b, gt), P, emb


def test_rank_ctr_np_vs_torch():
    from oracle import oracle_models as om
    rng = np.random.default_rng(2)
    cfg, (me, st, b, gt), P, emb = _setup(rng, 5)
    f64 = lambda d: {k: np.asarray(v, np.float64) for k, v in d.items()}
    out = om.rank_ctr_fwd(om.NP, f64(emb), f64(P), st, b, gt)
    te = {k: torch.from_numpy(v.astype(np.float64)).requires_grad_(True) for k, v in emb.items()}
    tout = om.rank_ctr_fwd(om.TH, te, {k: torch.from_numpy(v.astype(np.float64)) for k, v in P.items()}, st, b, gt)
    for k in out:
        assert rel_err(tout[k].detach().numpy(), out[k]) < 1e-12
        assert out[k].shape == (5, 1) and (out[k] >= 1e-6).all() and (out[k] <= 1).all()
    (tout["task0"].sum() + tout["task1"].sum()).backward()
    assert all(torch.isfinite(v.grad).all() for v in te.values())


@pytest.mark.gpu
def test_rank_ctr_sub_model_gpu(cuda_dev):
    from oracle import oracle_models as om
    from recommendsystem_b200.api.rank_ctr import TASK_NAMES, RankCtrSubModel, parse_feature_slots
    rng = np.random.default_rng(4)
    B = 48
    cfg, (me, st, b, gt), P, emb = _setup(rng, B)
    model = RankCtrSubModel(parse_feature_slots(cfg)).to(cuda_dev).eval()       # eval: no attention dropout
    te = {k: torch.from_numpy(v).to(cuda_dev).requires_grad_(True) for k, v in emb.items()}
    model(te)
    sd = model.state_dict()
    assert set(sd) == set(P), sorted(set(sd) ^ set(P))[:8]
    model.load_state_dict({k: torch.from_numpy(v).to(cuda_dev) for k, v in P.items()})
    out = model(te)
    f64 = lambda d: {k: np.asarray(v, np.float64) for k, v in d.items()}
    ref = om.rank_ctr_fwd(om.NP, f64(emb), f64(P), st, b, gt)
    for i, t in enumerate(TASK_NAMES):
        assert_close(out[t].detach().cpu().numpy(), ref["task%d" % i], REL_F32, t)
    (out[TASK_NAMES[0]].sum() - 2.0 * out[TASK_NAMES[1]].sum()).backward()
    tP = {k: torch.from_numpy(v.astype(np.float64)).requires_grad_(True) for k, v in P.items()}
    re = {k: torch.from_numpy(v.astype(np.float64)).requires_grad_(True) for k, v in emb.items()}
    ro = om.rank_ctr_fwd(om.TH, re, tP, st, b, gt)
    (ro["task0"].sum() - 2.0 * ro["task1"].sum()).backward()
    g = np.stack([te[k].grad.cpu().numpy() for k in sorted(te)])
    r = np.stack([re[k].grad.numpy() for k in sorted(re)])
    assert_close(g, r, 5 * REL_F32, "d/d embedding rows")
    named = dict(model.named_parameters())
    for name in ("senet_extract_layer.kernel", "emb_linear_map.3.kernel", "interact.query_dense_kernel",
                 "interact.layer_norm_gamma", "dnn_ppnet_gate.kernel", "dnn_can.kernel", "experts.gate_2_1_2.kernel",
                 "task_gates.gate_output_1.kernel", "task_dnn2.task1_dnn2_1.kernel", "task_out.0.bias"):
        assert_close(named[name].grad.cpu().numpy(), tP[name].grad.numpy(), 5 * REL_F32, "d/d " + name)


@pytest.mark.gpu
def test_rank_ctr_full_config_train_step(cuda_dev):
    """The shipped configuration end to end: 176 slots x 96-wide rows, 175 fields through the InteractingLayer."""
    from recommendsystem_b200.api.rank_ctr import TASK_NAMES, Model
    g, cfg = _golden_config()
    net = Model(cfg, bucket_size=2000, device=str(cuda_dev)).run()["net"]
    gen = torch.Generator().manual_seed(0)
    B = 64
    inputs = {s: torch.randint(0, 10 ** 9, (B,), generator=gen) for s in net.layout.sparse_slots}
    labels = {t: (torch.rand(B, 1, generator=gen) < 0.3).float().to(cuda_dev) for t in TASK_NAMES}
    pred = net.predict(inputs)
    assert pred[TASK_NAMES[0]].shape == (B, 1)
    losses = [float(net.train_step(inputs, labels)[0]) for _ in range(5)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
