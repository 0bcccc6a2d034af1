import sys, torch, json
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
gam = torch.ones(1280, device=dev)
for (nb, hw, c, c1, silu) in [(8, 4096, 320, 0, True), (8, 4096, 320, 0, False), (8, 4096, 320, 320, True), (8, 1024, 640, 0, True), (2, 4096, 320, 0, True), (2, 4096, 320, 320, True), (4, 4096, 512, 0, True), (64, 4096, 320, 0, True)]:
    x = torch.randn((nb, hw, c), device=dev).half()
    x1 = torch.randn((nb, hw, c1), device=dev).half() if c1 else None
    ws = torch.zeros(ops.groupnorm_ws_bytes(nb, hw, c + c1) // 8, device=dev, dtype=torch.float64)
    y = torch.empty((nb, hw, c + c1), device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.groupnorm(x, gam[:c + c1], gam[:c + c1], 1e-5, silu, x1=x1, out=y, ws=ws))
    print(json.dumps({"nb": nb, "hw": hw, "c": c + c1, "silu": silu, "us": round(us, 2), "GBs": round(nb * hw * (c + c1) * 4 / us / 1e3, 0)}), flush=True)
