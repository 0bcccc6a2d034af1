"""A handful of representative launches for `ncu --set full` captures (one per kernel family, UNet shapes at batch 4).
usage: python scripts/ncu_targets.py [which ...]   which in {linear, conv, fmha, ln, gn, xattn, conv8}"""
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import ops, pack  # noqa: E402


def main():
    which = sys.argv[1:] or ["linear", "conv", "fmha", "ln", "gn", "xattn", "conv8"]
    dev = torch.device("cuda", 0)
    nb = int(os.environ.get("UDT_NCU_NB", "8"))     # UNet batch (2 x images): 8 = configs[1], 64 = configs[2]
    reps = int(os.environ.get("UDT_NCU_REPS", "3"))
    if "linear" in which:
        m, k, n = nb * 4096, 320, 320
        x = torch.randn((m, k), device=dev).half()
        w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
        b = torch.randn((n,), device=dev)
        r = torch.randn((m, n), device=dev).half()
        y = torch.empty((m, n), device=dev, dtype=torch.float16)
        for _ in range(reps):
            ops.linear(x, w, b, residual=r, out=y)
        m, k, n = nb * 4096, 320, 2560
        w2 = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
        b2 = torch.randn((n,), device=dev)
        for _ in range(reps):
            ops.linear(x, w2, b2, act=ops.UDT_ACT_GEGLU)
    if "conv" in which:
        x = torch.randn((nb, 64, 64, 320), device=dev).half()
        w = (torch.randn((320, 9 * 320), device=dev) / math.sqrt(9 * 320)).half()
        b = torch.randn((320,), device=dev)
        for _ in range(reps):
            ops.conv3x3(x, w, b)
    if "conv8" in which:
        x = torch.randn((nb, 8, 8, 1280), device=dev).half()
        w = (torch.randn((1280, 9 * 1280), device=dev) / math.sqrt(9 * 1280)).half()
        b = torch.randn((1280,), device=dev)
        for _ in range(reps):
            ops.conv3x3(x, w, b)
    if "fmha" in which:
        n, heads = 4096, 5
        c = heads * 64
        qkv = torch.randn((nb * n, 3 * c), device=dev).half()
        for _ in range(reps):
            ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, heads, 0.125)
    if "ln" in which:
        x = torch.randn((nb * 4096, 320), device=dev).half()
        g = torch.ones(320, device=dev)
        for _ in range(reps):
            ops.layernorm(x, g, g, 1e-5)
    if "gn" in which:
        g = torch.ones(1280, device=dev)
        for hw, c in ((4096, 320), (1024, 640), (64, 1280)):   # two-pass, group-owner, group-owner
            x = torch.randn((nb, hw, c), device=dev).half()
            for _ in range(reps):
                ops.groupnorm(x, g[:c], g[:c], 1e-5, True)
    if "xattn" in which:
        q = torch.randn((nb * 4096, 320), device=dev).half()
        kv = torch.randn((nb * 12, 640), device=dev).half()
        for _ in range(reps):
            ops.xattn_small_l(q, kv[:, :320], kv[:, 320:], nb, 4096, 12, 5, 0.125)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
