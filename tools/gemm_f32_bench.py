"""Developer tool (GPU box): fp32 Dense-layer GEMMs - 3xTF32 tcgen05 kernel vs the FFMA kernel vs cuBLAS fp32
(torch.matmul, TF32 disabled) at the VideoDnn expert / AutoInt tower shapes, forward, dgrad and wgrad.
    python tools/gemm_f32_bench.py > gpurun_out/gemm_f32_bench.log"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from embed_sweep_util import timeit
from recommendsystem_b200 import cabi, ops

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
lib = cabi.load()
for (B, I, O) in ((16384, 1712, 256), (16384, 256, 128), (8192, 624, 256), (8192, 256, 128)):
    x = torch.randn(B, I, device=dev, generator=g)
    W = torch.randn(I, O, device=dev, generator=g) / I ** 0.5
    dy = torch.randn(B, O, device=dev, generator=g)
    cases = {
        "fwd  x[B,in].W[in,out]": (lambda: ops.gemm(x, W), lambda: torch.matmul(x, W), 2 * B * I * O),
        "dgrad dy.W^T": (lambda: ops.gemm(dy, W, transB=True), lambda: torch.matmul(dy, W.t()), 2 * B * I * O),
        "wgrad x^T.dy": (lambda: ops.gemm(x, dy, transA=True), lambda: torch.matmul(x.t(), dy), 2 * B * I * O),
    }
    for name, (mine, cublas, flops) in cases.items():
        lib.rs_set_fp32_gemm_mode(0)
        t3 = timeit(mine)
        lib.rs_set_fp32_gemm_mode(1)
        tf = timeit(mine)
        lib.rs_set_fp32_gemm_mode(0)
        tc = timeit(cublas)
        ref = cublas().double()
        err = float((mine().double() - ref).abs().max() / ref.abs().max())
        print(json.dumps({"shape": [B, I, O], "op": name, "tf32x3_us": round(t3, 1), "ffma_us": round(tf, 1),
                          "cublas_fp32_us": round(tc, 1), "tf32x3_TFs": round(flops / t3 / 1e6, 1),
                          "rel_err_vs_cublas_fp32": float(f"{err:.2e}")}), flush=True)
