"""Role timeline of udt_igemm for one problem shape (tuning aid): per-CTA clock64 stamps written by the kernel
while udt_debug_set_trace() is armed.  usage: python scripts/igemm_trace.py M N K [res] [geglu] [conv NB HW CIN COUT]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import lib, ops  # noqa: E402

EV = ["prod_first", "prod_last", "mma_first_data", "mma_commit", "epi0_acc_ready", "epi0_done", "epi1_acc_ready", "epi1_done"]


def run(fn, label, flops, iters=20):
    dev = torch.device("cuda", 0)
    L = lib.load()
    stride = L.udt_debug_set_trace(None, 0)
    assert stride > 0, "tracing is compiled out: rebuild with `UDT_TRACE=1 python -m udifftext_b200.build --force`"
    # timing without trace
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    buf = torch.zeros((148 * stride,), device=dev, dtype=torch.int64)
    L.udt_debug_set_trace(buf.data_ptr(), buf.numel() * 8)
    fn()
    torch.cuda.synchronize()
    L.udt_debug_set_trace(None, 0)
    t = buf.cpu().view(148, stride)
    used = t[:, 0] != 0
    t = t[used]
    g0 = int(t[:, 0].min())
    print(f"== {label}: {us:.2f} us/launch in a graph of {iters} back-to-back launches ({flops / us / 1e6:.0f} TFLOP/s); "
          f"{int(used.sum())} CTAs; traced launch span {(int(t[:, 1].max()) - g0) / 1e3:.2f} us "
          f"(CTA start spread {(int(t[:, 0].max()) - g0) / 1e3:.2f} us)")
    for cta in (0, len(t) // 2, len(t) - 1):
        row = t[cta]
        c0 = int(row[2])
        print(f"  CTA {cta}: start +{(int(row[0]) - g0) / 1e3:.2f} us, lifetime {int(row[3]) - c0} clk")
        ev = row[4:].view(-1, len(EV))
        for it in range(ev.shape[0]):
            if int(ev[it, 0]) == 0 and int(ev[it, 2]) == 0:
                break
            print("    tile %2d: " % it + "  ".join(f"{n}={int(ev[it, j]) - c0 if int(ev[it, j]) else -1}" for j, n in enumerate(EV)))


def main():
    dev = torch.device("cuda", 0)
    nb = 8
    cases = sys.argv[1:] or ["lin320", "lin320res", "qkv320", "geglu320", "ff2_320", "lin640res", "lin1280res", "conv320", "conv8", "conv1280_16"]
    for c in cases:
        if c.startswith("lin") or c in ("qkv320", "geglu320", "ff2_320"):
            m, k, n, res, act = {"lin320": (nb * 4096, 320, 320, False, 0), "lin320res": (nb * 4096, 320, 320, True, 0),
                                 "qkv320": (nb * 4096, 320, 960, False, 0), "geglu320": (nb * 4096, 320, 2560, False, ops.UDT_ACT_GEGLU),
                                 "ff2_320": (nb * 4096, 1280, 320, True, 0), "lin640res": (nb * 1024, 640, 640, True, 0),
                                 "lin1280res": (nb * 256, 1280, 1280, True, 0)}[c]
            x = torch.randn((m, k), device=dev).half()
            w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
            b = torch.randn((n,), device=dev)
            r = torch.randn((m, n), device=dev).half() if res else None
            y = torch.empty((m, n // 2 if act == ops.UDT_ACT_GEGLU else n), device=dev, dtype=torch.float16)
            run(lambda: ops.linear(x, w, b, residual=r, out=y, act=act), f"{c} M={m} K={k} N={n} res={res}", 2.0 * m * k * n)
        else:
            hw, cin, cout = {"conv320": (64, 320, 320), "conv8": (8, 1280, 1280), "conv1280_16": (16, 1280, 1280)}[c]
            x = torch.randn((nb, hw, hw, cin), device=dev).half()
            w = (torch.randn((cout, 9 * cin), device=dev) / math.sqrt(9 * cin)).half()
            b = torch.randn((cout,), device=dev)
            y = torch.empty((nb, hw, hw, cout), device=dev, dtype=torch.float16)
            run(lambda: ops.conv3x3(x, w, b, out=y), f"{c} {nb}x{hw}x{hw} {cin}->{cout}", 2.0 * nb * hw * hw * cout * 9 * cin)


if __name__ == "__main__":
    main()
