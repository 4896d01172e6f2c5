"""Developer tool (GPU box): label transform and metrics kernels against the HBM roofline.
    python tools/labels_bench.py > gpurun_out/labels_bench.log"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from embed_sweep_util import timeit
from recommendsystem_b200 import ops
from recommendsystem_b200.api.metrics import BinaryMetrics

dev = torch.device("cuda:0")
bins = torch.tensor([-19.0 + 0.5 * i for i in range(400)], device=dev)
for B in (16384, 262144, 1 << 20):
    wt = torch.randint(0, 200_000, (B,), device=dev, dtype=torch.int64)
    landing = (torch.rand(B, device=dev) < 0.3).to(torch.uint8)
    t = timeit(lambda: ops.staytime_labels(wt, bins, landing))
    alg = B * (8 + 1 + 401 * 4 + 8 + 8 + 4)
    print(json.dumps({"op": "staytime_labels", "B": B, "us": round(t, 2), "GBps": round(alg / t / 1e3, 1)}))
for n in (16384, 1 << 20, 1 << 24):
    y = (torch.rand(n, device=dev) < 0.3).float()
    p = torch.rand(n, device=dev)
    m = BinaryMetrics(device=dev)
    t = timeit(lambda: m.update_state(y, p))
    print(json.dumps({"op": "binary_metrics_update", "n": n, "us": round(t, 2), "GBps": round(n * 8 / t / 1e3, 1)}))
