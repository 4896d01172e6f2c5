"""Whole-step CUDA graphs for the composed models (VideoDnn MTL, DSSM, rank/ctr, AUTOINT).

The models' train steps are hundreds to a thousand small launches (one Dense / gate / slice per reference
layer) driven from Python autograd: on a B200 the step is bound by the host, not by the GPU.  Every launch in
them is capture-safe (the C-ABI kernels run on the current stream with caller-owned memory, the dense optimizer
is `capturable`, no host read-back), so the whole step - embedding gather, forward, backward, dense Adam, key
sort + sorted-segment sparse update - is captured ONCE and replayed:

    step = GraphedTrainStep(net, example_inputs, example_labels)      # warm-up + capture
    loss, outputs = step(inputs, labels)                               # copy into the static buffers, replay

The warm-up steps (they allocate workspaces and create the optimizer state) are UNDONE before the capture: the
embedding rows they touched, the dense parameters and both optimizers' state are restored bit for bit, so capturing
after loading a checkpoint does not change the loaded weights and the first replay is the first train step.

`net` is any object with `train_step(inputs: dict, labels: dict) -> (loss, outputs: dict)` (api.video_dnn.MtlNet,
api.rough_rank_model.DssmNet, api.rank_ctr.RankCtrNet, ...).  Shapes and dtypes are fixed by the example batch;
a different batch size needs its own GraphedTrainStep.  Inputs may live on the host (pinned or not): the copy
into the static device buffers is the only per-step host work besides the replay.
"""
from __future__ import annotations

from typing import Dict

import torch


def _static(d: Dict, dev):
    out = {}
    for k, v in d.items():
        out[k] = v.to(dev).clone() if isinstance(v, torch.Tensor) else v
    return out


class GraphedTrainStep:
    def __init__(self, net, inputs: Dict, labels: Dict, device=None, warmup: int = 3):
        dev = torch.device(device) if device is not None else next(
            v.device for v in list(labels.values()) + list(inputs.values()) if isinstance(v, torch.Tensor) and v.is_cuda)
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs a CUDA device")
        self.net, self.dev = net, dev
        self.inputs, self.labels = _static(inputs, dev), _static(labels, dev)
        snap = self._snapshot(net, self.inputs)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # allocates workspaces, creates the optimizer state
                net.train_step(self.inputs, self.labels)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(net, snap)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.outputs = net.train_step(self.inputs, self.labels)
        self.warmup_steps = 0                        # the warm-up steps were undone: replay 1 is train step 1

    @staticmethod
    def _snapshot(net, inputs):
        """Everything the warm-up steps change.  Lazily built layers are built first (a forward has no side effects)
        so that there are initial weights to return to."""
        if not hasattr(net, "emb") or not hasattr(net.emb, "snapshot"):
            return None
        with torch.no_grad():
            net.predict(inputs)
        snap = {"emb": net.emb.snapshot(inputs), "params": [p.detach().clone() for p in net.sub_model.parameters()]}
        opt = getattr(net, "opt", None)
        snap["opt"] = opt.snapshot() if opt is not None and hasattr(opt, "snapshot") else None
        if hasattr(net, "extra_state_snapshot"):
            snap["extra"] = net.extra_state_snapshot()
        return snap

    @staticmethod
    def _restore(net, snap):
        if snap is None:
            return
        net.emb.restore(snap["emb"])
        opt = getattr(net, "opt", None)
        if snap["opt"] is not None:
            opt.restore(snap["opt"])
        elif opt is not None and hasattr(opt, "flat_m"):          # created by the warm-up: back to a fresh state
            opt.flat_m.zero_(); opt.flat_v.zero_(); opt.scalars.zero_()
        with torch.no_grad():
            for p, v in zip(net.sub_model.parameters(), snap["params"]):
                p.copy_(v)
        if "extra" in snap:
            net.extra_state_restore(snap["extra"])

    def __call__(self, inputs: Dict, labels: Dict):
        for k, v in inputs.items():
            if isinstance(v, torch.Tensor):
                self.inputs[k].copy_(v, non_blocking=True)
        for k, v in labels.items():
            if isinstance(v, torch.Tensor):
                self.labels[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss, self.outputs
