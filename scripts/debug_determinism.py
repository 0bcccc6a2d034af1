"""Run the full UNet twice and report the first kernel call whose output differs between the runs."""
import os, sys
import torch
sys.path.insert(0, ".")
from udifftext_b200 import ops, synth
from udifftext_b200.unet import UNetB200

names = ["linear", "conv3x3", "groupnorm", "layernorm", "fmha", "xattn_small_l", "upsample2x"]
log = []
orig = {n: getattr(ops, n) for n in names}
def wrap(n):
    def f(*a, **k):
        y = orig[n](*a, **k)
        log.append((n, tuple(y.shape), y.clone()))
        return y
    return f
for n in names:
    setattr(ops, n, wrap(n))

dev = torch.device("cuda", 0)
gold = torch.load("tests/golden/full_unet.pt")
man = {k: v for k, v in synth.load_manifest("full").items() if k.startswith("model.diffusion_model.")}
sd = {k[len("model.diffusion_model."):]: v for k, v in synth.synthetic_state_dict(man, 1234).items()}
net = UNetB200(sd, dev, **synth.ARCH["full"]["unet"])
runs = []
for r in range(3):
    log.clear()
    y = net.forward(gold["x"], gold["t"], gold["ctx"])
    torch.cuda.synchronize()
    runs.append(list(log))
for r in (1, 2):
    nd = 0
    for i, (a, b) in enumerate(zip(runs[0], runs[r])):
        d = (a[2].float() - b[2].float()).abs().max().item()
        if d > 0:
            rel = ((a[2].float() - b[2].float()).norm() / a[2].float().norm()).item()
            nfrac = ((a[2] != b[2]).float().mean().item())
            print(f"run0 vs run{r}: op#{i} {a[0]} {a[1]} max|d|={d:.3e} rel={rel:.3e} frac_diff={nfrac:.3e}")
            nd += 1
            if nd >= 8:
                break
    print("total ops", len(runs[0]))
