"""Developer tool (GPU box): embedding gather / sorted-segment Adam bandwidth against the number of
lookups, and the tcgen05 GEMM against torch.matmul (cuBLAS) at the MLP-tower shapes.
    python tools/embed_sweep.py > gpurun_out/embed_sweep.log"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommendsystem_b200 import ops

dev = torch.device("cuda:0")
F, d, R = 39, 16, 1_000_000
g = torch.Generator(device=dev).manual_seed(1)
table = torch.randn(F * R, d, device=dev, generator=g) * 0.1
m = torch.zeros_like(table); v = torch.zeros_like(table)
row_base = (torch.arange(F, device=dev, dtype=torch.int64) * R)
rows = torch.full((F,), R, device=dev, dtype=torch.int64)
scal = torch.tensor([1.0, 0.9, 0.999, 1.0], device=dev)


sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from embed_sweep_util import timeit


for B in (8192, 32768, 131072):
    ids = torch.randint(0, R, (B, F), device=dev, dtype=torch.int64, generator=g)
    n = B * F
    for odt, ob in ((torch.bfloat16, 2), (torch.float32, 4)):
        t = timeit(lambda: ops.embed_gather(table, ids, row_base, rows, out_dtype=odt, want_keys=True))
        alg = n * (8 + 64 + d * ob + 8)
        print(json.dumps({"op": "gather", "B": B, "n": n, "out": str(odt), "us": round(t, 2), "GBps": round(alg / t / 1e3, 1)}))
    out, keys, _ = ops.embed_gather(table, ids, row_base, rows, out_dtype=torch.bfloat16, want_keys=True)
    ks = ops.sort_keys(keys, ops.row_bits(F * R))
    uniq = int(torch.unique(ks >> 32).numel())
    grad = torch.randn(n, d, device=dev, generator=g).bfloat16()
    t = timeit(lambda: ops.segsum_adam(table, m, v, grad, ks, 1e-3, 0.9, 0.999, 1e-8, scal))
    alg = n * (8 + d * 2) + uniq * 2 * 3 * d * 4
    print(json.dumps({"op": "segsum_adam", "B": B, "n": n, "unique": uniq, "us": round(t, 2), "GBps": round(alg / t / 1e3, 1)}))
    t = timeit(lambda: ops.sort_keys(keys, ops.row_bits(F * R), out=ks))
    print(json.dumps({"op": "sort_keys", "B": B, "us": round(t, 2)}))

# tcgen05 GEMM vs cuBLAS at the tower shapes (bf16, fp32 accumulate)
for (M, N, K) in ((8192, 256, 624), (8192, 128, 256), (8192, 624, 256), (8192, 256, 128)):
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    Bm = torch.randn(N, K, device=dev, generator=g).bfloat16()      # [out, in] K-major
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t0 = timeit(lambda: ops.gemm(A, Bm, C=C, transB=True))
    t1 = timeit(lambda: torch.matmul(A, Bm.t(), out=C))
    print(json.dumps({"op": "gemm", "MNK": [M, N, K], "rs_us": round(t0, 2), "cublas_us": round(t1, 2),
                      "rs_TFs": round(2 * M * N * K / t0 / 1e6, 1), "cublas_TFs": round(2 * M * N * K / t1 / 1e6, 1)}))
