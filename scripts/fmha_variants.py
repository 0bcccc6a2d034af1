"""FMHA variant check (tuning build; run once per UDT_FMHA_W4 / UDT_FMHA_POLY / UDT_FMHA_TAIL setting): accuracy incl. large score spreads and ragged key tiles, then timing."""
import json, os, sys, torch
sys.path.insert(0, ".")
from udifftext_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
tag = "persist=" + os.environ.get("UDT_FMHA_PERSIST", "0") + " pace=" + os.environ.get("UDT_FMHA_PACE_S", "-") + "/" + os.environ.get("UDT_FMHA_PACE_PV", "-") + " dbg=" + os.environ.get("UDT_FMHA_DEBUG", "0") + " ht=" + os.environ.get("UDT_FMHA_HT", "default") + " w4=" + os.environ.get("UDT_FMHA_W4", "default") + " poly=" + os.environ.get("UDT_FMHA_POLY", "default")
ACC = [] if os.environ.get("UDT_FMHA_SKIP_ACC") else [(1, 4096, 2, 1.0), (1, 1024, 1, 6.0), (1, 2048, 1, 6.0), (2, 4096, 5, 6.0), (1, 300, 1, 6.0), (1, 192, 2, 1.0),
                           (2, 64, 20, 6.0), (3, 256, 20, 3.0), (1, 9216, 2, 4.0)]
for (b, n, heads, mul) in ACC:
    c = heads * 64
    d = torch.randn((b * n, 3 * c), generator=g).half().to(dev)
    d[:, :c] *= mul
    y = ops.fmha(d[:, :c], d[:, c:2 * c], d[:, 2 * c:], b, n, n, heads, 0.125)
    torch.cuda.synchronize()
    q, k, v = (d[:, i * c:(i + 1) * c].float().view(b, n, heads, 64).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b * n, c)
    print(tag, (b, n, heads, mul), "rel err", ((y.float() - ref).norm() / ref.norm()).item(), flush=True)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


for (nb, n, heads) in [(8, 4096, 5), (8, 1024, 10), (8, 256, 20), (64, 4096, 5), (16, 9216, 5)]:
    c = heads * 64
    qkv = torch.randn((nb * n, 3 * c), device=dev).half()
    o = torch.empty((nb * n, c), device=dev, dtype=torch.float16)
    ms = timeit(lambda: ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, heads, 0.125, out=o))
    print(json.dumps({"variant": tag, "nb": nb, "n": n, "heads": heads, "us": round(ms * 1e3, 2),
                      "tflops": round(4.0 * nb * heads * n * n * 64 / ms / 1e9, 1)}), flush=True)
