"""LayerNorm timings at the UNet's row shapes (batch 4 and 32), CUDA-graph replays; tuning aid."""
import json, sys, torch
sys.path.insert(0, ".")
from udifftext_b200 import ops
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
for rows, c in [(32768, 320), (16384, 320), (8192, 640), (4096, 640), (2048, 1280), (1024, 1280), (512, 1280), (262144, 320), (65536, 640), (16384, 1280)]:
    x = torch.randn((rows, c), device=dev).half()
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev)
    y = torch.empty_like(x)
    us = timeit(lambda: ops.layernorm(x, g, b, 1e-5, out=y))
    print(json.dumps({"rows": rows, "c": c, "us": round(us, 2), "gbs": round(4.0 * x.numel() / us / 1e3, 1)}), flush=True)
