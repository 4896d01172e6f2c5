"""GPU parity of the label transform (rs_staytime_labels) and the streaming binary metrics
(rs_binary_metrics_update / _result) against oracle/oracle_metrics.py.
Bars: integer outputs (short / long labels, confusion histograms, counts) bit-exact; fp32 label values
within 1e-6 relative of the float32 restatement (expf vs numpy exp differ by <= 2 ulp) and of the
float64 one (1e-4: the exp argument is O(100)); metric values within 1e-12 of the oracle (same integer counts, double arithmetic)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("B", [1, 3, 7, 1000, 16384])
def test_staytime_labels(cuda_dev, B):
    from oracle import oracle_metrics as om
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B)
    watch = rng.integers(0, 200_000, size=B).astype(np.int64)
    watch[0] = 7000
    if B > 6:
        watch[1:7] = [7001, 18000, 18001, 160_000, 159_999, 2 ** 40]
    landing = (rng.random(B) < 0.3).astype(np.uint8)
    bins = torch.tensor(om.BIN_LIST, dtype=torch.float32, device=cuda_dev)
    lab, sh, lo, w = ops.staytime_labels(torch.from_numpy(watch).to(cuda_dev), bins,
                                         torch.from_numpy(landing).to(cuda_dev))
    torch.cuda.synchronize()
    r32 = om.staytime_labels(watch, landing, dtype=np.float32)
    r64 = om.staytime_labels(watch, landing, dtype=np.float64)
    assert lab.shape == (B, 401)
    np.testing.assert_array_equal(sh.cpu().numpy(), r32[1])
    np.testing.assert_array_equal(lo.cpu().numpy(), r32[2])
    np.testing.assert_array_equal(w.cpu().numpy(), r32[3])
    got = lab.cpu().numpy()
    np.testing.assert_array_equal(got[:, 400], r32[0][:, 400])           # capped watch time: exact
    np.testing.assert_allclose(got, r32[0], rtol=1e-6, atol=1e-36)
    np.testing.assert_allclose(got, r64[0], rtol=1e-4, atol=1e-36)     # exp argument O(100): fp32 rounding x100


def test_staytime_labels_optional_outputs_and_errors(cuda_dev):
    from recommendsystem_b200 import cabi, ops
    bins = torch.tensor([0.0, 1.0, 2.0], device=cuda_dev)
    wt = torch.tensor([1500, 0], dtype=torch.int64, device=cuda_dev)
    lab, sh, lo, w = ops.staytime_labels(wt, bins, None, want_weight=False, sigma=1.0, left=0.0, right=2.0)
    assert w is None and lab.shape == (2, 4)
    exp = np.exp(-np.square(np.array([0, 1, 2.0]) - 1.5) / 2) / np.sqrt(2 * np.pi) * 1.0
    np.testing.assert_allclose(lab[0, :3].cpu().numpy(), exp, rtol=1e-6)
    with pytest.raises(ValueError):
        ops.staytime_labels(wt.int(), bins)
    with pytest.raises(cabi.RsError):
        ops.staytime_labels(wt, bins, sigma=0.0)


def _synth(n, seed, pos=0.3):
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < pos).astype(np.float32)
    p = np.clip(0.3 + 0.25 * (y - 0.3) + 0.2 * rng.standard_normal(n), 0, 1).astype(np.float32)
    return y, p


def _oracle_all(y, p):
    from oracle import oracle_metrics as om
    return dict(auc=om.keras_auc(y, p), binary_accuracy=om.binary_accuracy(y, p), ctr=om.ctr(y), copc=om.copc(y, p))


@pytest.mark.parametrize("n", [1, 33, 1000, 16384, 300_001])
def test_binary_metrics_vs_oracle(cuda_dev, n):
    from oracle import oracle_metrics as om
    from recommendsystem_b200.api.metrics import BinaryMetrics
    y, p = _synth(n, n)
    # predictions sitting exactly on thresholds and on the ends
    t = om.keras_thresholds()
    k = min(n, 64)
    p[:k] = np.resize(np.concatenate([t[1:20], [0.0, 1.0, 0.5]]), k).astype(np.float32)
    m = BinaryMetrics(device=cuda_dev)
    m.update_state(torch.from_numpy(y).to(cuda_dev), torch.from_numpy(p).to(cuda_dev))
    T = 200
    st = m.state.cpu().numpy()
    tp, fp, tn, fn = om.confusion(y, p, t)
    pos_h, neg_h = st[:T + 1], st[T + 1:2 * T + 2]
    # tp_i = positives above threshold i = sum_{k > i} pos_hist[k]: bit-exact counts
    np.testing.assert_array_equal(np.cumsum(pos_h[::-1])[::-1][1:], tp.astype(np.int64))
    np.testing.assert_array_equal(np.cumsum(neg_h[::-1])[::-1][1:], fp.astype(np.int64))
    assert st[2 * T + 2] == n
    from test_dist_cpu import _metric_state_np          # the accumulator layout the gloo merge test restates
    want = _metric_state_np(y, p)
    np.testing.assert_array_equal(st[:2 * T + 4], want[:2 * T + 4])
    np.testing.assert_allclose(st[2 * T + 4:].view(np.float64), want[2 * T + 4:].view(np.float64), rtol=1e-13)
    r, o = m.result(), _oracle_all(y, p)
    assert r["count"] == n
    for key in ("auc", "binary_accuracy", "ctr", "copc"):
        assert r[key] == pytest.approx(o[key], rel=1e-12, abs=1e-12), key
    assert r["mean_prediction"] == pytest.approx(float(p.astype(np.float64).mean()), rel=1e-12)


def test_binary_metrics_streaming_equals_one_pass_and_is_deterministic(cuda_dev):
    from recommendsystem_b200.api.metrics import BinaryMetrics
    y, p = _synth(50_000, 7)
    yt, pt = torch.from_numpy(y).to(cuda_dev), torch.from_numpy(p).to(cuda_dev)
    a, b, c = (BinaryMetrics(device=cuda_dev) for _ in range(3))
    a.update_state(yt, pt)
    c.update_state(yt, pt)
    assert torch.equal(a.state, c.state)                         # bitwise, double sums included
    for lo in range(0, 50_000, 12_500):
        b.update_state(yt[lo:lo + 12_500].view(-1, 1), pt[lo:lo + 12_500].view(-1, 1))
    T = 200
    assert torch.equal(a.state[:2 * T + 4], b.state[:2 * T + 4])  # integer words: exact
    ra, rb = a.result(), b.result()
    assert ra["auc"] == rb["auc"] and ra["binary_accuracy"] == rb["binary_accuracy"]
    assert ra["copc"] == pytest.approx(rb["copc"], rel=1e-13)
    b.reset_states()
    assert int(b.state.abs().sum()) == 0


def test_binary_metrics_bf16_and_degenerate(cuda_dev):
    from oracle import oracle_metrics as om
    from recommendsystem_b200.api.metrics import BinaryMetrics
    y, p = _synth(4096, 3)
    pb = torch.from_numpy(p).to(cuda_dev).bfloat16()
    m = BinaryMetrics(device=cuda_dev)
    m.update_state(torch.from_numpy(y).to(cuda_dev), pb)
    assert m.result()["auc"] == pytest.approx(om.keras_auc(y, pb.float().cpu().numpy()), abs=1e-12)
    ones = BinaryMetrics(device=cuda_dev)
    ones.update_state(torch.ones(100, device=cuda_dev), torch.rand(100, device=cuda_dev))
    assert ones.result()["auc"] == 0.0 and ones.result()["ctr"] == 1.0          # div_no_nan, as Keras
    empty = BinaryMetrics(device=cuda_dev)
    empty.update_state(torch.zeros(0, device=cuda_dev), torch.zeros(0, device=cuda_dev))
    assert empty.result() == dict(auc=0.0, binary_accuracy=0.0, ctr=0.0, copc=0.0, count=0.0, mean_prediction=0.0)
    with pytest.raises(ValueError):
        m.update_state(torch.ones(3, device=cuda_dev), torch.ones(4, device=cuda_dev))


def test_metric_classes_and_parse_input_func(cuda_dev):
    from oracle import oracle_metrics as om
    from recommendsystem_b200.api.metrics import AUC, COPC, CTR, BinaryAccuracy, BinaryMetrics
    from recommendsystem_b200.api.staytime_parse import TASK_PREFIX, parse_input_func
    y, p = _synth(2000, 5)
    yt, pt = torch.from_numpy(y).to(cuda_dev), torch.from_numpy(p).to(cuda_dev)
    auc = AUC(device=cuda_dev)
    auc.update_state(yt, pt)
    assert auc.result() == pytest.approx(om.keras_auc(y, p), abs=1e-12)
    core = BinaryMetrics(device=cuda_dev)
    four = [BinaryAccuracy(shared=core), AUC(shared=core), CTR(shared=core), COPC(shared=core)]
    core.update_state(yt, pt)
    for mm in four:
        mm.update_state(yt, pt)                                   # shared: no second pass
    o = _oracle_all(y, p)
    assert [mm.result() for mm in four] == pytest.approx(
        [o["binary_accuracy"], o["auc"], o["ctr"], o["copc"]], rel=1e-12)
    assert int(core.state[2 * 200 + 2]) == 2000

    batch = {"watch_duration": np.array([6000, 20000, 9000], np.int64),
             "extra_info": ["x", "a video_homepage_landing b", b"label"], "2125": "kept"}
    feats, yd, w = parse_input_func(batch, device=cuda_dev)
    assert feats["2125"] == "kept" and "watch_duration" not in feats and feats["example_id"] == batch["extra_info"]
    assert yd[TASK_PREFIX + "shortplay"].tolist() == [0, 1, 1] and yd[TASK_PREFIX + "longplay"].tolist() == [0, 1, 0]
    assert w.reshape(-1).tolist() == [1.0, 5.0, 1.0]
    ref = om.staytime_labels(batch["watch_duration"], [0, 1, 0])[0]
    np.testing.assert_allclose(yd[TASK_PREFIX + "staytime"].cpu().numpy(), ref, rtol=1e-6, atol=1e-36)
