"""Each hot kernel beside the LIBRARY kernel for the same problem on the same GPU (cuBLAS / cuDNN / SDPA backends /
flash-attn / torch's native norms), UNet shapes of a CFG step at batch 4 (UNet batch 8) or `argv[1]`.
Both sides: fp16 in / out, CUDA-graph replays of 20 back-to-back launches, best of 3.  The library side gets the easier
job where the product kernel fuses more (no bias / activation / residual on the cuBLAS and cuDNN calls).
Tuning evidence only (profiles/r02_vs_library.json); bench.py is the contract benchmark."""
import json
import math
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from udifftext_b200 import ops  # noqa: E402

torch.backends.cudnn.benchmark = True


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best * 1e3   # us


def attempt(fn):
    try:
        return round(timeit(fn), 2)
    except Exception as ex:  # a backend that does not support the shape on this build
        torch.cuda.synchronize()
        return "n/a: " + str(ex).split("\n")[0][:80]


def main():
    dev = torch.device("cuda", 0)
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows = []
    # 3x3 convolutions
    for cin, cout, hw in [(320, 320, 64), (640, 640, 32), (1280, 1280, 16), (1280, 1280, 8), (2560, 1280, 8)]:
        x = torch.randn((nb, hw, hw, cin), device=dev).half()
        w4 = (torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half()
        wp = w4.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
        b = torch.randn((cout,), device=dev)
        y = torch.empty((nb, hw, hw, cout), device=dev, dtype=torch.float16)
        ours = attempt(lambda: ops.conv3x3(x, wp, b, out=y))
        xn = x.permute(0, 3, 1, 2)   # NCHW view over channels-last memory
        wcl = w4.contiguous(memory_format=torch.channels_last)
        lib = attempt(lambda: F.conv2d(xn, wcl, None, padding=1))
        ref = F.conv2d(xn, wcl, b.half(), padding=1).permute(0, 2, 3, 1)
        err = float((y.float() - ref.float()).norm() / ref.float().norm())
        rows.append({"op": "conv3x3", "nb": nb, "cin": cin, "cout": cout, "hw": hw, "ours_us": ours, "cudnn_us": lib,
                     "gflop": round(2.0 * nb * hw * hw * cout * 9 * cin / 1e9, 1), "rel_diff": round(err, 5)})
    # linears
    for m, k, n in [(nb * 4096, 320, 320), (nb * 4096, 320, 960), (nb * 4096, 320, 2560), (nb * 4096, 1280, 320),
                    (nb * 1024, 640, 1920), (nb * 1024, 640, 5120), (nb * 1024, 2560, 640),
                    (nb * 256, 1280, 3840), (nb * 256, 1280, 10240), (nb * 256, 5120, 1280)]:
        x = torch.randn((m, k), device=dev).half()
        w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
        y = torch.empty((m, n), device=dev, dtype=torch.float16)
        yy = torch.empty((m, n), device=dev, dtype=torch.float16)
        ours = attempt(lambda: ops.linear(x, w, out=y))
        lib = attempt(lambda: torch.matmul(x, w.t(), out=yy))
        rows.append({"op": "linear", "m": m, "k": k, "n": n, "ours_us": ours, "cublas_us": lib,
                     "gflop": round(2.0 * m * k * n / 1e9, 1)})
    # self-attention, head dim 64
    from torch.nn.attention import SDPBackend, sdpa_kernel
    try:
        from flash_attn import flash_attn_func
    except Exception:
        flash_attn_func = None
    for n, heads in [(4096, 5), (1024, 10), (256, 20)]:
        c = heads * 64
        qkv = torch.randn((nb * n, 3 * c), device=dev).half()
        o = torch.empty((nb * n, c), device=dev, dtype=torch.float16)
        ours = attempt(lambda: ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, heads, 0.125, out=o))
        q, k_, v = (qkv[:, i * c:(i + 1) * c].reshape(nb, n, heads, 64) for i in range(3))
        qh, kh, vh = (t.permute(0, 2, 1, 3).contiguous() for t in (q, k_, v))   # [B, h, N, d] contiguous: the library's best case
        r = {"op": "fmha", "nb": nb, "n": n, "heads": heads, "ours_us": ours,
             "gflop": round(4.0 * nb * heads * n * n * 64 / 1e9, 1)}
        for name, be in [("sdpa_cudnn_us", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash_us", SDPBackend.FLASH_ATTENTION),
                         ("sdpa_efficient_us", SDPBackend.EFFICIENT_ATTENTION)]:
            def run(be=be):
                with sdpa_kernel(be):
                    return F.scaled_dot_product_attention(qh, kh, vh)
            r[name] = attempt(run)
        if flash_attn_func is not None:
            qc, kc, vc = q.contiguous(), k_.contiguous(), v.contiguous()
            r["flash_attn_pkg_us"] = attempt(lambda: flash_attn_func(qc, kc, vc))
        rows.append(r)
    # normalisations
    for hw, c in [(4096, 320), (4096, 640), (1024, 640), (1024, 1280), (256, 1280), (64, 1280)]:
        x = torch.randn((nb, hw, c), device=dev).half()
        g = torch.ones(c, device=dev)
        b = torch.zeros(c, device=dev)
        y = torch.empty_like(x)
        ws = torch.empty(ops.groupnorm_ws_bytes(nb, hw, c) // 8, device=dev, dtype=torch.float64)
        ours = attempt(lambda: ops.groupnorm(x, g, b, 1e-5, True, out=y, ws=ws))
        side = int(math.isqrt(hw))
        xn = x.reshape(nb, side, side, c).permute(0, 3, 1, 2)
        gh, bh = g.half(), b.half()
        lib = attempt(lambda: F.silu(F.group_norm(xn, 32, gh, bh, 1e-5)))
        rows.append({"op": "groupnorm_silu", "nb": nb, "hw": hw, "c": c, "ours_us": ours, "torch_us": lib,
                     "mbytes": round(4.0 * x.numel() / 1e6, 1)})
        ours = attempt(lambda: ops.layernorm(x, g, b, 1e-5, out=y))
        lib = attempt(lambda: F.layer_norm(x, (c,), gh, bh, 1e-5))
        rows.append({"op": "layernorm", "rows": nb * hw, "c": c, "ours_us": ours, "torch_us": lib,
                     "mbytes": round(4.0 * x.numel() / 1e6, 1)})
    for r in rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
