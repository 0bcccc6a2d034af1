import math, sys, torch
sys.path.insert(0, ".")
from udifftext_b200 import ops
dev = torch.device("cuda", 0)
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(20_000_000)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1000
res_mode = sys.argv[1] if len(sys.argv) > 1 else "res"
for m, k, n in [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (32768, 320, 960), (32768, 1280, 320)]:
    x = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    b = torch.randn((n,), device=dev); r = torch.randn((m, n), device=dev).half(); y = torch.empty((m, n), device=dev, dtype=torch.float16)
    for bn in (0, 64, 96, 128, 160, 256):
        if bn > n: continue
        try:
            us = timeit(lambda: ops.linear(x, w, b, residual=(r if res_mode == "res" else None), out=y, bn_hint=bn))
        except Exception as ex:
            print(m, k, n, bn, "ERR", str(ex)[:60]); continue
        print(f"{m}x{k}x{n} bn={bn:3d} {res_mode}: {us:7.1f} us  {2.0*m*k*n/us/1e6:7.1f} TFLOP/s  {(m*k+m*n*(2 if res_mode=='res' else 1))*2/us/1e3:6.0f} GB/s")
