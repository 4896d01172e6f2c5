"""Label half of the reference's input function (`staytime/parse.py:16-71`) on the GPU.

TFRecord decoding (`tf.io.parse_example`) and file sharding are outside the hot path; what follows the
decode - short / long play labels, the Gaussian-smoothed 400-bin staytime distribution with the capped
watch time appended, and the landing-page sample weights - is one kernel here (rs_staytime_labels).
`parse_input_func` keeps the reference's contract: it takes the decoded batch (a dict holding
'watch_duration' int64 [B], 'extra_info' [B] strings / bytes, and the slot features) and returns
`(feature_dict, y, sample_weight)` with the same keys.
"""
from __future__ import annotations

import re
from typing import Sequence

import numpy as np
import torch

from .. import ops
from .staytime_config import Config as C

TASK_PREFIX = "video_id_rank_staytime_mtl_ppnet_v7_"
_LANDING = re.compile(r".*video_homepage_landing.*")   # tf.strings.regex_full_match, RE2: "." excludes \n (parse.py:66)

_bins_cache = {}


def _bins(device) -> torch.Tensor:
    key = str(device)
    if key not in _bins_cache:
        _bins_cache[key] = torch.tensor(C.bin_list, dtype=torch.float32, device=device)
    return _bins_cache[key]


def landing_mask(extra_info: Sequence) -> np.ndarray:
    """uint8 [B]: which `extra_info` strings fully match '.*video_homepage_landing.*' (host string work)."""
    out = np.zeros(len(extra_info), np.uint8)
    for i, s in enumerate(extra_info):
        if isinstance(s, (bytes, bytearray)):
            s = s.decode("utf-8", "replace")
        out[i] = 1 if _LANDING.fullmatch(str(s)) else 0
    return out


def staytime_labels(watch_duration: torch.Tensor, landing: torch.Tensor | None = None):
    """watch_duration int64 [B] (ms, on the GPU) -> (y dict, sample_weight [B,1]) as parse.py:30-72."""
    if watch_duration.device.type != "cuda":
        raise RuntimeError("staytime_labels runs on a CUDA device only (no CPU fallback)")
    wt = watch_duration.reshape(-1)
    label, short, long_, weight = ops.staytime_labels(wt, _bins(wt.device), landing, short_ms=7000, long_ms=18000,
                                                      cap_s=160.0, sigma=4.0, left=-19.0, right=180.5,
                                                      landing_weight=5.0)
    y = {
        TASK_PREFIX + "staytime": label,          # [B, multiclass_num + 1]
        TASK_PREFIX + "shortplay": short,         # [B] int64
        TASK_PREFIX + "longplay": long_,
    }
    return y, weight


def parse_input_func(batch: dict, device="cuda:0"):
    """Decoded batch -> (feature_dict, y, sample_weight); same keys as the reference (parse.py:16-72)."""
    feature_dict = dict(batch)
    wt = feature_dict.pop("watch_duration")
    extra = feature_dict.pop("extra_info", None)
    if extra is None:
        extra = ["label"] * len(wt)                                   # FixedLenFeature default_value (parse.py:18)
    feature_dict["example_id"] = extra                                # parse.py:27
    wt = torch.as_tensor(wt, dtype=torch.int64).to(device, non_blocking=True)
    landing = torch.from_numpy(landing_mask(extra)).to(device, non_blocking=True)
    y, weight = staytime_labels(wt, landing)
    return feature_dict, y, weight
