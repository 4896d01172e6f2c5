"""Developer tool: what-if timing of the tcgen05 InteractingLayer backward (RS_NVCC_DEFS=-DRS_ITB_PROFILE build).
Each mode removes one ingredient (results are WRONG; only the time matters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import cabi, ops
B, F, D, L = 8192, 39, 16, 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, F, D, device=dev, generator=g).bfloat16()
dy = torch.randn(B, F, D, device=dev, generator=g).bfloat16()
W = (torch.rand(D, 64, device=dev, generator=g) - 0.5) * 0.8
b = torch.zeros(64, device=dev); gm = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
y, saved = ops.interacting_fwd(x, W, b, gm, bt, 1e-3, 2, L, True, compute_bf16=True)
lib = cabi.load()
names = {0: "baseline", 1: "no MMA issue", 2: "no fence.proxy.async", 4: "no exp", 8: "no P/dS stores", 14: "no fence/exp/stores",
         15: "all four"}
for mode, name in names.items():
    lib.rs_debug_itb_exp(mode)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.interacting_bwd(x, saved, W, b, gm, bt, 1e-3, 2, L, dy, True, compute_bf16=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:22s} {sorted(ts)[2]:8.1f} us")
lib.rs_debug_itb_exp(0)
