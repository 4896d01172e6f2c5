"""Developer tool (GPU box): DIN attention unit (BASELINE configs[2]: behaviour sequence length 100, B = 4096,
H = 16) forward / backward bandwidth against the HBM roofline; graph-timed like tools/embed_sweep.py.
Algorithmic bytes: forward B*(T*H*s + H*s + T) + out; backward adds the dkeys write."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import cabi, ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from embed_sweep_util import timeit

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
H, Hd = 16, 16
for B, T in ((4096, 100), (65536, 100), (16384, 50)):
    for dt, s in ((torch.float32, 4), (torch.bfloat16, 2)):
        q = torch.randn(B, H, device=dev, generator=g).to(dt)
        keys = torch.randn(B, T, H, device=dev, generator=g).to(dt)
        lens = torch.randint(1, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
        lens[0] = T
        mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.uint8).contiguous()
        dout = torch.randn(B, H, device=dev, generator=g).to(dt)
        for mode, name, win in ((cabi.DIN_A, "A (din.py)", 3 * H), (cabi.DIN_B, "B (staytime/layer.py)", 4 * H)):
            W1 = torch.randn(win, Hd, device=dev, generator=g) * 0.2
            b1 = torch.zeros(Hd, device=dev); W2 = torch.randn(Hd, 1, device=dev, generator=g) * 0.2; b2 = torch.zeros(1, device=dev)
            vals = keys if mode == cabi.DIN_A else None
            sl = lens if mode == cabi.DIN_A else None
            mk = None if mode == cabi.DIN_A else mask
            tf = timeit(lambda: ops.din_fwd(mode, q, keys, vals, sl, mk, W1, b1, W2, b2))
            tb = timeit(lambda: ops.din_bwd(mode, q, keys, vals, sl, mk, W1, b1, W2, b2, dout))
            fb = B * (T * H * s + H * s + T) + B * H * s
            bb = fb + B * T * H * s + B * H * s
            print(json.dumps({"op": "din", "mode": name, "B": B, "T": T, "dtype": str(dt), "fwd_us": round(tf, 1),
                              "fwd_GBps": round(fb / tf / 1e3, 1), "bwd_us": round(tb, 1), "bwd_GBps": round(bb / tb / 1e3, 1)}))
