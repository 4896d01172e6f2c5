"""Developer tool (GPU box): train-step time of the composed models built on the hot-path kernels
(SURVEY §8 rows a13 VideoDnn MTL = cfg5, a14 DSSM, f1 rank/ctr) through the drop-in API, eager launches,
inputs resident on the GPU.  Prints one JSON line per model: ms per step, samples/s, library launches per step.
    python tools/models_bench.py > gpurun_out/models_bench.log"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import cabi

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def run(name, step, B, warm=3, reps=8, net=None, inputs=None, labels=None):
    if net is not None:
        _run(name + " [eager]", step, B, warm, reps)
        from recommendsystem_b200.api.graph import GraphedTrainStep
        gs = GraphedTrainStep(net, inputs, labels)
        _run(name + " [CUDA graph]", lambda: gs(inputs, labels), B, warm, reps * 4)
    else:
        _run(name, step, B, warm, reps)


def _run(name, step, B, warm=3, reps=8):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    lib = cabi.load()
    l0 = lib.rs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"model": name, "B": B, "ms_per_step": round(ms, 3), "host_ms_per_step": round(wall, 3),
                      "samples_per_s": round(B / ms * 1e3), "launches_per_step": (lib.rs_launch_count() - l0) // reps}),
          flush=True)


def video_dnn(B=16384, T=50):
    from recommendsystem_b200.api.staytime_config import Config as C
    from recommendsystem_b200.api.video_dnn import TASK_KEYS, mtl_net
    net = mtl_net(C.SLOTS, C.SEQ_SLOTS, T, dnn_hidden_units=(256, 128), bucket_size=81920, device=str(dev))["net"]
    inputs = {s: torch.randint(0, 10 ** 9, (B,), generator=g).to(dev) for s in C.SLOTS}
    for s in C.SEQ_SLOTS:
        ids = torch.randint(0, 10 ** 9, (B, T), generator=g)
        lens = torch.randint(0, T + 1, (B,), generator=g)
        ids[torch.arange(T)[None, :] >= lens[:, None]] = -1
        inputs[s] = ids.to(dev)
    y0 = torch.softmax(torch.randn(B, 400, generator=g), -1)
    labels = {TASK_KEYS[0]: torch.cat([y0, torch.zeros(B, 1)], 1).to(dev),
              TASK_KEYS[1]: (torch.rand(B, 1, generator=g) < 0.3).float().to(dev),
              TASK_KEYS[2]: (torch.rand(B, 1, generator=g) < 0.3).float().to(dev)}
    run("VideoDnn mtl_net (cfg5: 91 slots x 32, 3 seq slots T=50)", lambda: net.train_step(inputs, labels), B, net=net, inputs=inputs, labels=labels)


def dssm(B=16384):
    from recommendsystem_b200.api.rough_rank_model import DSSM, config as RC
    net = DSSM(bucket_size=25600, device=str(dev))["net"]
    din = {f: torch.randint(0, 10 ** 9, (B,), generator=g).to(dev) for f in RC.USER_FEATURE_IDS + RC.ITEM_FEATURE_IDS}
    din[RC.DENSE_MASK_ID] = (torch.rand(B, 1, generator=g) < 0.5).float().to(dev)
    dl = {"student": (torch.rand(B, 1, generator=g) < 0.3).float().to(dev),
          "teacher": (torch.rand(B, 1, generator=g) < 0.3).float().to(dev)}
    run("DSSM rough_rank (52 features x 16)", lambda: net.train_step(din, dl), B, net=net, inputs=din, labels=dl)


def rank_ctr(B=4096):
    from recommendsystem_b200.api.rank_ctr import TASK_NAMES, Model
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, "tests"))
    from test_rank_ctr import _golden_config          # the reference's shipped model_parameter.json, as a fixture
    _, cfg = _golden_config()
    net = Model(cfg, bucket_size=265000, device=str(dev)).run()["net"]
    inputs = {s: torch.randint(0, 10 ** 9, (B,), generator=g).to(dev) for s in net.layout.sparse_slots}
    labels = {t: (torch.rand(B, 1, generator=g) < 0.3).float().to(dev) for t in TASK_NAMES}
    run("rank/ctr production model (176 slots x 96, InteractingLayer F=175)", lambda: net.train_step(inputs, labels), B, net=net, inputs=inputs, labels=labels)


if __name__ == "__main__":
    which = sys.argv[1:] or ["video_dnn", "dssm", "rank_ctr"]
    for w in which:
        try:
            globals()[w]()
        except Exception as e:                      # keep going: one line per model
            print(json.dumps({"model": w, "error": repr(e)[:300]}), flush=True)
