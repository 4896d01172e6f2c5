"""Forward / dgrad / wgrad fp32 GEMMs of the VideoDnn expert layer (16384 x 1712 -> 256) through the 3xTF32
tcgen05 kernel, and one staytime-label + metrics pass (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import ops
from recommendsystem_b200.api.metrics import BinaryMetrics
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, I, O = 16384, 1712, 256
x = torch.randn(B, I, device=dev, generator=g); W = torch.randn(I, O, device=dev, generator=g) / I ** 0.5
dy = torch.randn(B, O, device=dev, generator=g)
for _ in range(2):
    ops.gemm(x, W); ops.gemm(dy, W, transB=True); ops.gemm(x, dy, transA=True)
bins = torch.tensor([-19.0 + 0.5 * i for i in range(400)], device=dev)
wt = torch.randint(0, 200_000, (262144,), device=dev, dtype=torch.int64)
m = BinaryMetrics(device=dev)
y = (torch.rand(1 << 22, device=dev) < 0.3).float(); p = torch.rand(1 << 22, device=dev)
for _ in range(2):
    ops.staytime_labels(wt, bins)
    m.update_state(y, p)
torch.cuda.synchronize()
print("ok")
