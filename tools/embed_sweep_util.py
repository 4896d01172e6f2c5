"""Shared timing helper of the developer sweep tools."""
import torch


def timeit(fn, reps=10, inner=10):
    """median microseconds per call; `inner` calls are captured in one CUDA graph so that the host
    launch path (ctypes) does not bound small kernels"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(inner):
                fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort()
    return ts[len(ts) // 2]


