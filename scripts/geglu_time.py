"""timing of the GEGLU projections of one UNet CFG step at batch 4 (CUDA-graph replays, burst clocks)"""
import json, math, sys, torch
sys.path.insert(0, ".")
from udifftext_b200 import ops, pack
dev = torch.device("cuda", 0)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e3
for m, k, n in [(32768, 320, 2560), (8192, 640, 5120), (2048, 1280, 10240)]:
    x = torch.randn((m, k), device=dev).half()
    w = torch.randn((n, k)) / math.sqrt(k)
    wp, bp = pack.pack_geglu(w, torch.randn(n))
    wp, bp = wp.to(dev), bp.to(dev)
    y = torch.empty((m, n // 2), device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.linear(x, wp, bp, act=ops.UDT_ACT_GEGLU, out=y))
    print(json.dumps({"m": m, "k": k, "n": n, "us": round(us, 2), "tflops": round(2.0 * m * k * n / us / 1e6, 0)}), flush=True)
