"""Developer tool: where does the 3xTF32 GEMM deviate? (per storage order: zero fraction, error map by 32x32 block)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recommendsystem_b200 import ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
M, N, K = 256, 256, 128
for tA, tB in ((False, True), (False, False), (True, False), (True, True)):
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    try:
        C = ops.gemm(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev), transA=tA, transB=tB)
        torch.cuda.synchronize()
    except Exception as e:
        print("tA", tA, "tB", tB, "ERROR", repr(e)[:200]); continue
    got = C.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref)
    print(f"tA={tA} tB={tB}: zero frac {np.mean(got == 0):.3f} max err {err.max():.3e} (max ref {np.abs(ref).max():.2f})")
    blk = err.reshape(M // 32, 32, N // 32, 32).max(axis=(1, 3))
    print(np.array2string(blk, precision=1, max_line_width=200))
    # does got match a K-truncated / partial product?
    for kk in (8, 16, 32, 64, 96):
        r2 = (A.T if tA else A)[:, :kk].astype(np.float64) @ (B.T if tB else B)[:, :kk].T.astype(np.float64) if False else None
