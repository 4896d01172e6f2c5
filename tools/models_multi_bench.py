"""bench.py --config cfg4 | cfg5: the composed models of BASELINE configs[3] / configs[4] on N GPUs (one process per
GPU, torch.distributed.run), train step as ONE CUDA graph per rank (api.graph.GraphedTrainStep), synthetic data.

  cfg4  rank/multi_head AUTOINT (rank/multi_head/multidnn.py:214-259): 39 slots x 8-d, tables of 200 M rows in total
        (+ Adam moments: 19.2 GB) row-sharded over the ranks, batch 8192 per GPU
  cfg5  staytime VideoDnn mtl_net (staytime/VideoDnn.py:266-302): 91 slots x 32-d + 3 sequence slots (T = 50),
        bucket 81 920, batch 16384 per GPU

N = 1 runs the unsharded EmbeddingFeatures; N > 1 the row-sharded ShardedEmbeddingFeatures (peer-memory gather /
scatter over NVLink, NCCL only for the id all-to-all and the dense all-reduce).  value = samples/s of the whole job,
CUDA events, max over ranks.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(args):
    import torch
    import torch.distributed as dist
    from recommendsystem_b200 import cabi
    from recommendsystem_b200.api.graph import GraphedTrainStep
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    kw = {}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from recommendsystem_b200.api.sharded_embedding import ShardedEmbeddingFeatures
        kw = dict(embedding_cls=ShardedEmbeddingFeatures, group=None)
    g = torch.Generator(device=dev).manual_seed(20261018 + rank)
    if args.config == "cfg4":
        from recommendsystem_b200.api.builders import AUTOINT, AUTOINT_LABELS
        B, F = 8192, 39
        bucket = int(os.environ.get("RS_CFG4_ROWS", 200_000_000)) // F
        slots = [str(1000 + i) for i in range(F)]
        net = AUTOINT(slots, [], True, dnn_hidden_units=(32, 16), bucket_size=bucket, device=str(dev), seed=1, **kw).model
        inputs = {s: torch.randint(0, 2 ** 40, (B,), device=dev, generator=g) for s in slots}
        labels = {k: (torch.rand(B, 1, device=dev, generator=g) < 0.1).float() for k in AUTOINT_LABELS}
        workload = (f"rank/multi_head AUTOINT: InteractingLayer(1, 8 units, 2 heads, dropout 0.2) || DNN(32,16) -> 7-expert / "
                    f"7-gate MMoE -> 7 heads; {F} slots x 8-d, {bucket * F / 1e6:.0f} M rows (+ Adam state) row-sharded over "
                    f"{world} GPU(s), batch {B}/GPU (BASELINE configs[3])")
        metric = "autoint_multihead_train_samples_per_s"
    else:
        from recommendsystem_b200.api.staytime_config import Config as C
        from recommendsystem_b200.api.video_dnn import TASK_KEYS, create_model_func
        B, T = 16384, 50
        net = create_model_func(device=str(dev), seed=1, **kw)["net"]
        inputs = {s: torch.randint(0, 2 ** 40, (B,), device=dev, generator=g) for s in C.SLOTS}
        for s in C.SEQ_SLOTS:
            ids = torch.randint(0, 2 ** 40, (B, T), device=dev, generator=g)
            lens = torch.randint(0, T + 1, (B,), device=dev, generator=g)
            ids[torch.arange(T, device=dev)[None, :] >= lens[:, None]] = -1
            inputs[s] = ids
        y0 = torch.softmax(torch.randn(B, 400, device=dev, generator=g), -1)
        labels = {TASK_KEYS[0]: torch.cat([y0, torch.zeros(B, 1, device=dev)], 1),
                  TASK_KEYS[1]: (torch.rand(B, 1, device=dev, generator=g) < 0.3).float(),
                  TASK_KEYS[2]: (torch.rand(B, 1, device=dev, generator=g) < 0.3).float()}
        workload = (f"staytime VideoDnn mtl_net: 91 slots x 32-d + 3 sequence slots (T=50), DIN x3, SENet, FM, FFM, PPNet-gated "
                    f"3-expert MMoE, DCN, 400-bin head; bucket 81920, batch {B}/GPU on {world} GPU(s) (BASELINE configs[4])")
        metric = "videodnn_train_samples_per_s"
    torch.manual_seed(11)
    n0 = cabi.launch_count()
    step = GraphedTrainStep(net, inputs, labels, warmup=2)
    launches = (cabi.launch_count() - n0) // 3
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        loss, _ = step(inputs, labels)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss, _ = step(inputs, labels)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if hasattr(net.emb, "check_overflow"):
            net.emb.check_overflow()
    if rank == 0:
        print(json.dumps({
            "metric": metric, "value": B * world * K / (ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload, "batch_per_gpu": B,
                                                            "parallelism": f"dp{world}" + ("+row-sharded tables (peer memory)" if world > 1 else "")},
            "gpu_launches": int(launches * K), "launches_per_step": int(launches), "loss": float(loss)}), flush=True)
    torch.cuda.synchronize()
    if world > 1:
        sys.stdout.flush()
        os._exit(0)
