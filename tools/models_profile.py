"""Developer tool (GPU box): which kernels a composed model's train step spends its time in (torch profiler,
CUPTI kernel times).  python tools/models_profile.py video_dnn|dssm|rank_ctr > gpurun_out/models_profile.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import models_bench as mb
from torch.profiler import ProfilerActivity, profile

steps = {}


def capture(name, step, B, net=None, inputs=None, labels=None, **kw):
    if net is not None and len(sys.argv) > 2 and sys.argv[2] == "graph":
        from recommendsystem_b200.api.graph import GraphedTrainStep
        gs = GraphedTrainStep(net, inputs, labels)
        steps[name + " [CUDA graph]"] = lambda: gs(inputs, labels)
    else:
        steps[name] = step


mb.run = capture  # (name, step, B, **kw)
which = sys.argv[1] if len(sys.argv) > 1 else "video_dnn"
getattr(mb, which)()
(name, step), = steps.items()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(name)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
