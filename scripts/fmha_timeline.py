"""Per-tile timeline of the FMHA kernel's softmax warps (build with `UDT_TRACE=1 UDT_STAMPS=1 python -m udifftext_b200.build
--force`): CTA 0 prints clock() stamps of key tiles 8..11; this script runs one launch and prints the stamps as deltas.
softmax slots: 0 before s_full wait, 1 scores ready, 2 first 32 columns in registers, 3 chunk 0 done, 4 last chunk loaded,
5 exponentials done, 6 P stored (tcgen05.st complete), 7 p_full arrived.  (Stamps inside the MMA-issuing warps were used once
to find the issue chain — ~40 clk per tcgen05.mma, ~100 per commit — and removed: they lengthen the chain they measure.)"""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from udifftext_b200 import ops
    n, heads, nb = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    c = heads * 64
    qkv = torch.randn((nb * n, 3 * c), device="cuda").half()
    o = torch.empty((nb * n, c), device="cuda", dtype=torch.float16)
    ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, heads, 0.125, out=o)
    torch.cuda.synchronize()
    sys.exit(0)
n, heads, nb = (sys.argv[1:4] + ["4096", "5", "8"][len(sys.argv) - 1:])[:3] if len(sys.argv) > 1 else ("4096", "5", "8")
env = dict(os.environ, UDT_FMHA_DEBUG=str(int(os.environ.get("UDT_FMHA_DEBUG", "0")) | 64))
out = subprocess.run([sys.executable, __file__, "child", n, heads, nb], env=env, capture_output=True, text=True).stdout
rows = {}
for line in out.splitlines():
    if line.startswith("FSTAMP"):
        f = line.split()
        rows[(int(f[1][1:]), int(f[2][1:]))] = [int(x) for x in f[3:]]
if not rows:
    print("no stamps (build with UDT_TRACE=1 UDT_STAMPS=1)"); sys.exit(1)
t0 = min(v[0] for v in rows.values())
for (w, j), v in sorted(rows.items()):
    rel = [(x - t0) & 0xffffffff for x in v]
    print(f"w{w} j{j} abs0={rel[0]:6d} deltas " + " ".join(f"{(rel[i + 1] - rel[i]):5d}" for i in range(7)) + f" | end {rel[7]:6d}")
