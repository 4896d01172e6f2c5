"""Developer tool: per-phase clock64 breakdown of the tcgen05 InteractingLayer backward (CTA 0, thread 0).
Build with RS_NVCC_DEFS=-DRS_ITB_PROFILE, run on a GPU box:  python tools/itb_profile.py"""
import ctypes, os, sys
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
lib.rs_debug_itb_exp(256)          # clock64 hooks on
out = (ctypes.c_ulonglong * 32)()
for _ in range(3):
    ops.interacting_bwd(x, saved, W, b, gm, bt, 1e-3, 2, L, dy, True, compute_bf16=True)
lib.rs_debug_itb_profile(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.interacting_bwd(x, saved, W, b, gm, bt, 1e-3, 2, L, dy, True, compute_bf16=True)
e1.record(); torch.cuda.synchronize()
lib.rs_debug_itb_profile(out, 0)
v = list(out)[:12]
tot = sum(v)
print("kernel ms", e0.elapsed_time(e1), "cycles CTA0", tot)
names = ["S", "dP+dV", "dQ+dK", "dX+dW+Z"]
for p in range(4):
    print(f"phase {names[p]:6s} work+sync {v[3*p]:9d} ({v[3*p]/tot:5.1%})  issue {v[3*p+1]:9d} ({v[3*p+1]/tot:5.1%})  mma-wait {v[3*p+2]:9d} ({v[3*p+2]/tot:5.1%})")

w = list(out)
lab = ["ld_window", "max/exp/sum/normalise", "pack + smem store", "fence.proxy.async", "bar.sync wait"]
for name, base in (("warp 0 (one window)", 16), ("warp 1 (two windows)", 24)):
    print(name, {lab[i]: w[base + i] for i in range(5)})
