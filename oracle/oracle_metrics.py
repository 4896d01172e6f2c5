"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the steps either side of the reference's train step;
never imported by the product package (only tests/, __graft_entry__.smoke() and bench.py may use oracle/).

    staytime_labels(...)        label half of parse_input_func, staytime/parse.py:30-68 (+ config.py bin_list)
    keras_auc(...)              tf.keras.metrics.AUC() as compiled in rough_rank/model.py:216-217 and
                                staytime/model.py:80-81: update_confusion_matrix_variables + AUC.result()
    binary_accuracy / ctr / copc  the other metrics of the same lists (tn.metric.* restated, see below)

TensorFlow and TensorNet are not installable offline and the reference has no golden vectors, so these
follow the published Keras algorithm (metrics.AUC: thresholds [-eps] + i/(T-1) + [1+eps], eps = 1e-7,
`pred > threshold`, labels cast to bool, ROC with 'interpolation' summation, div_no_nan) operation by
operation; tn.metric.CTR / COPC are restated from their meaning (mean label; sum(label)/sum(pred)).
Parity unpinned against TF itself; `tests/test_oracle_metrics.py` pins the AUC restatement against the
rank-statistic definition of ROC AUC (exact pairwise count) within the 200-threshold discretisation error.
"""
import math

import numpy as np

BIN_LIST = [-19.0 + 0.5 * i for i in range(400)]            # staytime/config.py:18-58
MULTICLASS_NUM = 400                                         # staytime/config.py:17


def staytime_labels(watch_ms, extra_info_landing=None, bin_list=BIN_LIST, dtype=np.float32):
    """staytime/parse.py:30-68 in the reference's operation order; dtype=float32 reproduces TF's arithmetic,
    float64 gives the exact value the tolerance is measured against."""
    wt_i = np.asarray(watch_ms, np.int64)
    short_label = np.where(wt_i > 7000, 1, 0).astype(np.int64)               # :30-34
    long_label = np.where(wt_i > 18000, 1, 0).astype(np.int64)               # :36-38
    wt = wt_i.astype(dtype)                                                   # :40
    wt = wt / dtype(1000.0)                                                   # :41
    wt = np.where(wt > dtype(160.0), dtype(160.0), wt)                        # :42
    n = wt.shape[0]
    bins = np.repeat(np.asarray([bin_list], dtype), n, axis=0)                # :45-46
    wt = wt.reshape(n, 1)                                                     # :48
    wt_ext = np.repeat(wt, len(bin_list), axis=1)                             # :50
    dist = bins - wt_ext                                                      # :52
    sq = np.square(np.abs(dist))                                              # :53
    left, right = -19, 180.5                                                  # :55-56
    width = (right - left) / (len(bin_list) - 1)                              # :57
    sigma = 4
    div_num = dtype(math.sqrt(2 * math.pi) * sigma)                           # :59
    label = np.exp(sq / dtype(-2 * math.pow(sigma, 2))) / div_num             # :60
    label = (label * dtype(width)).astype(dtype)                              # :61
    staytime_label = np.concatenate([label, wt], -1)                          # :62
    if extra_info_landing is None:
        weight = np.ones_like(wt)
    else:
        weight = np.where(np.asarray(extra_info_landing).reshape(n, 1) != 0, dtype(5), dtype(1)).astype(dtype)  # :64
    return staytime_label, short_label, long_label, weight


def keras_thresholds(num_thresholds=200, eps=1e-7):
    t = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    return np.asarray([0.0 - eps] + t + [1.0 + eps], np.float32)


def confusion(y_true, y_pred, thresholds):
    """update_confusion_matrix_variables: per threshold tp, fp, tn, fn (counts), float32 compare."""
    p = np.asarray(y_pred, np.float32).reshape(-1)
    y = np.asarray(y_true).reshape(-1).astype(bool)
    above = p[None, :] > np.asarray(thresholds, np.float32)[:, None]          # [T, n]
    tp = (above & y[None, :]).sum(1).astype(np.float64)
    fp = (above & ~y[None, :]).sum(1).astype(np.float64)
    fn = (~above & y[None, :]).sum(1).astype(np.float64)
    tn = (~above & ~y[None, :]).sum(1).astype(np.float64)
    return tp, fp, tn, fn


def _div_no_nan(a, b):
    return np.where(b != 0, a / np.where(b != 0, b, 1), 0.0)


def keras_auc_from_counts(tp, fp, tn, fn):
    """AUC.result() for curve='ROC', summation_method='interpolation'."""
    T = len(tp)
    recall = _div_no_nan(tp, tp + fn)
    fp_rate = _div_no_nan(fp, fp + tn)
    x, y = fp_rate, recall
    heights = (y[:T - 1] + y[1:]) / 2.0
    return float(np.sum((x[:T - 1] - x[1:]) * heights))


def keras_auc(y_true, y_pred, num_thresholds=200):
    return keras_auc_from_counts(*confusion(y_true, y_pred, keras_thresholds(num_thresholds)))


def binary_accuracy(y_true, y_pred, threshold=0.5):
    p = np.asarray(y_pred, np.float32).reshape(-1)
    y = np.asarray(y_true, np.float32).reshape(-1)
    return float(np.mean((p > np.float32(threshold)).astype(np.float32) == y))


def ctr(y_true, y_pred=None):
    return float(np.mean(np.asarray(y_true, np.float64)))


def copc(y_true, y_pred):
    return float(np.sum(np.asarray(y_true, np.float64)) / np.sum(np.asarray(y_pred, np.float32).astype(np.float64)))


def exact_roc_auc(y_true, y_pred):
    """Rank-statistic ROC AUC (ties count 1/2): the quantity Keras' thresholded AUC approximates."""
    p = np.asarray(y_pred, np.float64).reshape(-1)
    y = np.asarray(y_true).reshape(-1).astype(bool)
    order = np.argsort(p, kind="mergesort")
    ranks = np.empty(len(p), np.float64)
    sp = p[order]
    i = 0
    while i < len(sp):
        j = i
        while j + 1 < len(sp) and sp[j + 1] == sp[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    n1, n0 = y.sum(), (~y).sum()
    return float((ranks[y].sum() - n1 * (n1 + 1) / 2.0) / (n1 * n0))
