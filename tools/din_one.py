"""One DIN-A and one DIN-B forward + backward at B = 16384, T = 100, fp32 (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import cabi, ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
B, T, H, Hd = 16384, 100, 16, 16
q = torch.randn(B, H, device=dev, generator=g); keys = torch.randn(B, T, H, device=dev, generator=g)
lens = torch.randint(1, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.uint8).contiguous()
dout = torch.randn(B, H, device=dev, generator=g)
for mode, win in ((cabi.DIN_A, 3 * H), (cabi.DIN_B, 4 * H)):
    W1 = torch.randn(win, Hd, device=dev, generator=g) * 0.2
    b1 = torch.zeros(Hd, device=dev); W2 = torch.randn(Hd, 1, device=dev, generator=g) * 0.2; b2 = torch.zeros(1, device=dev)
    vals = keys if mode == cabi.DIN_A else None
    sl = lens if mode == cabi.DIN_A else None
    mk = None if mode == cabi.DIN_A else mask
    for _ in range(2):
        ops.din_fwd(mode, q, keys, vals, sl, mk, W1, b1, W2, b2)
        ops.din_bwd(mode, q, keys, vals, sl, mk, W1, b1, W2, b2, dout)
torch.cuda.synchronize()
print("ok")
