"""Developer tool (GPU box): a few launches of the tcgen05 InteractingLayer forward + backward at the bench shape
(for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, F, D, L = 8192, 39, 16, 3
x = torch.randn(B, F, D, device=dev, generator=g).bfloat16()
dy = torch.randn(B, F, D, device=dev, generator=g).bfloat16()
W = (torch.rand(D, 64, device=dev, generator=g) - 0.5) * 0.8
b = torch.zeros(64, device=dev); gm = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
for _ in range(3):
    y, saved = ops.interacting_fwd(x, W, b, gm, bt, 1e-3, 2, L, True, compute_bf16=True)
    ops.interacting_bwd(x, saved, W, b, gm, bt, 1e-3, 2, L, dy, True, compute_bf16=True)
torch.cuda.synchronize()
print("ok")
