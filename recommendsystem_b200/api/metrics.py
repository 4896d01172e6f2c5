"""Streaming metrics of the reference's `compile(metrics=...)` lists, on the GPU
(rough_rank/model.py:215-219: BinaryAccuracy, AUC, tn.metric.CTR, tn.metric.COPC;
staytime/model.py:78-82: BinaryAccuracy, AUC).  Keras metric protocol: `update_state(y_true, y_pred)`,
`result()`, `reset_states()`.

One `BinaryMetrics` pass (rs_binary_metrics_update) feeds all four: exact int64 confusion histograms over
the Keras AUC thresholds, sample / correct counts and ordered double sums of labels and predictions.
The per-metric classes below either own a state or share one (`shared=`), so a head that reports all four
reads its predictions once.  `all_reduce()` merges the additive states of the ranks (the reference's
TensorNet metrics are summed over shards the same way).

tf.keras.metrics.AUC defaults restated: num_thresholds=200, curve='ROC', summation_method='interpolation',
thresholds = [-1e-7] + [i/199 for i in 1..198] + [1 + 1e-7] (float32), pred > threshold, label cast to bool.
tn.metric.CTR / COPC live in un-vendored TensorNet: restated from their published meaning
(CTR = clicks / shows = mean label; COPC = clicks / predicted clicks = sum(label) / sum(pred)) - parity unpinned.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

KERAS_EPSILON = 1e-7


def keras_auc_thresholds(num_thresholds: int = 200) -> np.ndarray:
    """metrics.AUC.__init__: evenly spaced interior thresholds plus the two epsilon-padded ends, as float32."""
    if num_thresholds <= 1:
        raise ValueError("`num_thresholds` must be > 1.")
    t = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    return np.asarray([0.0 - KERAS_EPSILON] + t + [1.0 + KERAS_EPSILON], np.float32)


class BinaryMetrics:
    """Shared accumulator: AUC, accuracy, CTR, COPC, count, mean prediction."""
    NAMES = ("auc", "binary_accuracy", "ctr", "copc", "count", "mean_prediction")

    def __init__(self, num_thresholds: int = 200, threshold: float = 0.5, device="cuda:0"):
        self.num_thresholds = int(num_thresholds)
        self.threshold = float(threshold)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BinaryMetrics runs on a CUDA device only (no CPU fallback)")
        self.thresholds = torch.from_numpy(keras_auc_thresholds(num_thresholds)).to(self.device)
        self.state = ops.binary_metrics_state(self.num_thresholds, self.device)
        self._out = torch.empty(6, dtype=torch.float64, device=self.device)

    def update_state(self, y_true, y_pred, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("weighted metrics: the reference passes metrics=, not weighted_metrics=")
        y_true = y_true.to(self.device, torch.float32).reshape(-1)
        y_pred = y_pred.to(self.device).reshape(-1)
        if y_pred.dtype not in (torch.float32, torch.bfloat16):
            y_pred = y_pred.float()
        ops.binary_metrics_update(self.state, y_pred, y_true, self.thresholds, self.threshold)

    def result_tensor(self) -> torch.Tensor:
        """float64[6] on the device (no synchronisation)."""
        return ops.binary_metrics_result(self.state, self.num_thresholds, self._out)

    def result(self) -> dict:
        r = self.result_tensor().cpu().numpy()
        return {k: float(v) for k, v in zip(self.NAMES, r)}

    def reset_states(self):
        self.state.zero_()

    reset_state = reset_states

    def all_reduce(self, group=None):
        """Sum the states of all ranks in place (integer words and the two double sums separately)."""
        merge_metric_states(self.state, self.num_thresholds, group)


def merge_metric_states(state: torch.Tensor, num_thresholds: int, group=None):
    """state: int64[2T+6]; words [0, 2T+4) are integer counts, the last two hold float64 bit patterns."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return state
    n_int = 2 * num_thresholds + 4
    counts = state[:n_int].clone()
    sums = state[n_int:].view(torch.float64).clone()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    state[:n_int].copy_(counts)
    state[n_int:].copy_(sums.view(torch.int64))
    return state


class _One:
    KEY = ""

    def __init__(self, name=None, shared: BinaryMetrics | None = None, device="cuda:0", **kw):
        self.name = name or self.KEY
        self.core = shared if shared is not None else BinaryMetrics(device=device, **kw)
        self._owns = shared is None

    def update_state(self, y_true, y_pred, sample_weight=None):
        if self._owns:
            self.core.update_state(y_true, y_pred, sample_weight)

    def result(self) -> float:
        return self.core.result()[self.KEY]

    def reset_states(self):
        if self._owns:
            self.core.reset_states()

    reset_state = reset_states


class AUC(_One):
    """tf.keras.metrics.AUC() with its defaults (ROC, 200 thresholds, interpolation)."""
    KEY = "auc"

    def __init__(self, num_thresholds=200, name=None, shared=None, device="cuda:0"):
        if shared is None:
            super().__init__(name, None, device, num_thresholds=num_thresholds)
        else:
            super().__init__(name, shared)


class BinaryAccuracy(_One):
    """tf.keras.metrics.BinaryAccuracy(threshold=0.5)."""
    KEY = "binary_accuracy"

    def __init__(self, name=None, threshold=0.5, shared=None, device="cuda:0"):
        if shared is None:
            super().__init__(name, None, device, threshold=threshold)
        else:
            super().__init__(name, shared)


class CTR(_One):
    """tn.metric.CTR(): mean label (clicks / shows)."""
    KEY = "ctr"


class COPC(_One):
    """tn.metric.COPC(): sum(label) / sum(prediction)."""
    KEY = "copc"
