"""Developer tool (GPU box): a few launches of the tcgen05 InteractingLayer forward (with the fused embedding lookup) +
backward at the bench shape through the trainer's own calls (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200.autoint import AutoIntConfig, AutoIntTrainer

dev = torch.device("cuda:0")
tr = AutoIntTrainer(AutoIntConfig(batch=8192, dtype="bf16", rows_per_field=1_000_000), dev)
g = torch.Generator(device=dev).manual_seed(0)
for _ in range(3):
    ids = torch.randint(0, 1_000_000, (8192, 39), device=dev, generator=g)
    y = (torch.rand(8192, 1, device=dev, generator=g) < 0.25).float()
    tr.step(ids, y)
torch.cuda.synchronize()
print("ok", float(tr.loss))
