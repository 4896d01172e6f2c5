"""Developer tool: clock64 timeline of CTA (0,0,0) of the tcgen05 GEMM.  Build with RS_NVCC_DEFS=-DRS_GEMM_PROFILE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import cabi, ops
dev = torch.device("cuda:0")
lib = cabi.load()
for (M, N, K) in ((8192, 256, 624), (8192, 128, 256), (8192, 624, 256)):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, B, C=C, transB=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(A, B, C=C, transB=True); e1.record(); torch.cuda.synchronize()
    out = (ctypes.c_longlong * 16)()
    lib.rs_debug_gemm_profile(out)
    t = list(out)[:6]
    print((M, N, K), "event us", round(e0.elapsed_time(e1) * 1e3, 1), "cycles: setup", t[1] - t[0], "first TMA", t[2] - t[1],
          "mainloop", t[3] - t[2], "acc ready", t[4] - t[3], "epilogue", t[5] - t[4], "total", t[5] - t[0])
