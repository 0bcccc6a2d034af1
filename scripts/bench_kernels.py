"""Micro-benchmarks of the individual kernels (CUDA events, L2-flushing between reps is skipped here: the
shapes are the UNet's and the numbers are for tuning only; bench.py is the contract benchmark)."""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import ops, pack  # noqa: E402


def timeit(fn, iters=20, warm=3):
    """device time per launch inside a CUDA graph of `iters` back-to-back launches (no host launch cost)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda", 0)
    out = []
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    # conv3x3 shapes (cin, cout, hw)
    for cin, cout, hw in [(320, 320, 64), (640, 640, 32), (1280, 1280, 16), (1280, 1280, 8), (2560, 1280, 8), (960, 640, 32)]:
        x = torch.randn((nb, hw, hw, cin), device=dev).half()
        w = (torch.randn((cout, 9 * cin), device=dev) / math.sqrt(9 * cin)).half()
        b = torch.randn((cout,), device=dev)
        y = torch.empty((nb, hw, hw, cout), device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.conv3x3(x, w, b, out=y))
        fl = 2.0 * nb * hw * hw * cout * 9 * cin
        out.append({"op": "conv3x3", "nb": nb, "cin": cin, "cout": cout, "hw": hw, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)})
    for m, k, n in [(nb * 4096, 320, 960), (nb * 4096, 320, 2560), (nb * 4096, 1280, 320), (nb * 1024, 640, 5120), (nb * 256, 1280, 10240), (nb * 256, 5120, 1280)]:
        x = torch.randn((m, k), device=dev).half()
        w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
        y = torch.empty((m, n), device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.linear(x, w, out=y))
        out.append({"op": "linear", "m": m, "k": k, "n": n, "ms": round(ms, 4), "tflops": round(2.0 * m * k * n / ms / 1e9, 1)})
        yy = torch.empty((m, n), device=dev, dtype=torch.float16)
        ms = timeit(lambda: torch.matmul(x, w.t(), out=yy))
        out.append({"op": "cublas", "m": m, "k": k, "n": n, "ms": round(ms, 4), "tflops": round(2.0 * m * k * n / ms / 1e9, 1)})
    for n, heads in [(4096, 5), (1024, 10), (256, 20)]:
        c = heads * 64
        qkv = torch.randn((nb * n, 3 * c), device=dev).half()
        o = torch.empty((nb * n, c), device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, heads, 0.125, out=o))
        fl = 4.0 * nb * heads * n * n * 64
        out.append({"op": "fmha", "nb": nb, "n": n, "heads": heads, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)})
    for hw, c in [(4096, 320), (4096, 640), (1024, 640), (1024, 1280), (256, 1280), (256, 2560), (64, 1280), (64, 2560)]:
        x = torch.randn((nb, hw, c), device=dev).half()
        g = torch.ones(c, device=dev)
        b = torch.zeros(c, device=dev)
        y = torch.empty_like(x)
        ws = torch.empty(ops.groupnorm_ws_bytes(nb, hw, c) // 8, device=dev, dtype=torch.float64)
        ms = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, True, out=y, ws=ws))
        out.append({"op": "groupnorm_silu", "nb": nb, "hw": hw, "c": c, "ms": round(ms, 4), "gbs": round(4.0 * x.numel() / ms / 1e6, 1)})
        if c > 2048:
            continue
        ms = timeit(lambda: ops.layernorm(x, g, b, 1e-5, out=y))
        out.append({"op": "layernorm", "rows": nb * hw, "c": c, "ms": round(ms, 4), "gbs": round(4.0 * x.numel() / ms / 1e6, 1)})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
