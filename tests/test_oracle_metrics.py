"""Oracle for labels / metrics (CPU): internal consistency and pins against independent definitions."""
import numpy as np
import pytest

from oracle import oracle_metrics as om


def test_staytime_label_shape_and_mass():
    watch = np.array([0, 500, 7000, 7001, 18000, 18001, 60_000, 159_999, 160_000, 400_000], np.int64)
    lab, sh, lo, w = om.staytime_labels(watch, dtype=np.float64)
    assert lab.shape == (10, 401) and sh.tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 1, 1]
    assert lo.tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1, 1]
    assert lab[-1, -1] == 160.0 and lab[1, -1] == 0.5
    # a Gaussian of sigma 4 sampled every 0.5 s times the bin width: sums to ~1 away from the edges
    assert abs(lab[6, :400].sum() - 1.0) < 1e-3
    assert np.argmax(lab[6, :400]) == om.BIN_LIST.index(60.0)
    lab32 = om.staytime_labels(watch, dtype=np.float32)[0]
    np.testing.assert_allclose(lab32, lab, rtol=2e-5, atol=1e-30)
    w5 = om.staytime_labels(watch, np.arange(10) % 2)[3]
    assert w5.reshape(-1).tolist() == [1, 5] * 5


def test_keras_thresholds():
    t = om.keras_thresholds()
    assert t.shape == (200,) and t[0] < 0 and t[-1] > 1 and np.all(np.diff(t) > 0)
    assert t[1] == np.float32(1 / 199)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_keras_auc_tracks_exact_auc(seed):
    rng = np.random.default_rng(seed)
    y = (rng.random(5000) < 0.3).astype(np.float32)
    p = np.clip(0.3 + 0.25 * (y - 0.3) + 0.2 * rng.standard_normal(5000), 0, 1).astype(np.float32)
    a, e = om.keras_auc(y, p), om.exact_roc_auc(y, p)
    assert abs(a - e) < 2e-3                       # 200-threshold trapezoid vs the pairwise statistic
    assert om.keras_auc(y, y) == pytest.approx(1.0, abs=1e-12)
    assert om.keras_auc(y, 1 - y) == pytest.approx(0.0, abs=1e-12)
    assert om.keras_auc(y, np.full_like(y, 0.5)) == pytest.approx(0.5, abs=1e-12)


def test_degenerate_labels_use_div_no_nan():
    p = np.linspace(0, 1, 50, dtype=np.float32)
    assert om.keras_auc(np.zeros(50), p) == 0.0
    assert om.keras_auc(np.ones(50), p) == 0.0


def test_simple_metrics():
    y = np.array([1, 0, 1, 0], np.float32)
    p = np.array([0.9, 0.6, 0.5, 0.1], np.float32)
    assert om.binary_accuracy(y, p) == 0.5           # 0.5 is not > 0.5
    assert om.ctr(y) == 0.5
    assert om.copc(y, p) == pytest.approx(2 / 2.1, rel=1e-6)


def test_product_thresholds_and_regex_match_the_oracle():
    """Host logic of the product side (no GPU): the AUC thresholds it uploads are the oracle's, bit for bit, and the
    landing-page match follows RE2 full-match semantics ('.' does not cross a newline, parse.py:66)."""
    from recommendsystem_b200.api.metrics import keras_auc_thresholds
    from recommendsystem_b200.api.staytime_parse import landing_mask
    for T in (2, 3, 50, 200, 1000):
        np.testing.assert_array_equal(keras_auc_thresholds(T), om.keras_thresholds(T))
    with pytest.raises(ValueError):
        keras_auc_thresholds(1)
    got = landing_mask(["video_homepage_landing", "x video_homepage_landing y", b"video_homepage_landing",
                        "video_homepage_landin", "", "a\nvideo_homepage_landing", "label"])
    assert got.tolist() == [1, 1, 1, 0, 0, 0, 0]


def test_no_cpu_fallback_for_labels_and_metrics():
    import torch
    from recommendsystem_b200.api.metrics import BinaryMetrics
    from recommendsystem_b200.api.staytime_parse import staytime_labels
    with pytest.raises(RuntimeError):
        BinaryMetrics(device="cpu")
    with pytest.raises(RuntimeError):
        staytime_labels(torch.zeros(4, dtype=torch.int64))


def test_staytime_label_is_a_sampled_normal_pdf():
    """Independent statement of parse.py:53-63: label[b, j] = N(bin_j; wt_b, sigma = 4) * bin width (scipy)."""
    from scipy.stats import norm
    watch = np.array([0, 12_345, 60_000, 159_000, 999_999], np.int64)
    lab = om.staytime_labels(watch, dtype=np.float64)[0]
    wt = np.minimum(watch / 1000.0, 160.0)
    width = (180.5 + 19) / 399
    want = norm.pdf(np.asarray(om.BIN_LIST)[None, :], loc=wt[:, None], scale=4.0) * width
    np.testing.assert_allclose(lab[:, :400], want, rtol=1e-12, atol=1e-300)
    np.testing.assert_array_equal(lab[:, 400], wt)


def test_auc_against_sklearn():
    """Second, independent pin: scikit-learn's ROC AUC equals the rank-statistic restatement exactly and the
    Keras 200-threshold AUC to within its discretisation error."""
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(9)
    for n, pos in ((2000, 0.5), (20000, 0.05)):
        y = (rng.random(n) < pos).astype(np.float32)
        p = np.clip(0.4 * y + 0.3 + 0.25 * rng.standard_normal(n), 0, 1).astype(np.float32)
        sk = roc_auc_score(y, p)
        assert om.exact_roc_auc(y, p) == pytest.approx(sk, abs=1e-12)
        assert abs(om.keras_auc(y, p) - sk) < 3e-3
