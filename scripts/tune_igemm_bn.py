"""Measured column-tile (BN) table for udt_igemm: every distinct igemm problem of one UNet CFG step is replayed inside a
CUDA graph with the library's own choice (bn_hint = 0, incl. split-K) and with every legal explicit tile; problems where an
explicit tile beats the cost model by more than 3 % go into udifftext_b200/tuning/igemm_bn_b200.json, which ops.igemm consults.
usage: python scripts/tune_igemm_bn.py [batch ...]   (default 4)"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import ops, synth  # noqa: E402

OUT = os.path.join("udifftext_b200", "tuning", "igemm_bn_b200.json")
CANDS = (256, 224, 192, 160, 128, 96, 64, 32)


def time_graph(fn, args, iters=20):
    rc = fn(*args, torch.cuda.current_stream().cuda_stream)      # validate the configuration outside of stream capture
    if rc != 0:
        raise RuntimeError(rc)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(iters):
            rc = fn(*args, st)
            if rc != 0:
                raise RuntimeError(rc)
    gr.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


def tune_batch(b, unet, dev, table, report):
    from udifftext_b200.host.runner import StepRunner
    from udifftext_b200.host.schedule import DiscreteDenoiser, LegacyDDPMDiscretization
    r = StepRunner(unet, b, 64, 64, 12, 5.0)
    den = DiscreteDenoiser({"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
                           {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"}, 1000,
                           {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"})
    g = torch.Generator().manual_seed(0)
    cond = {"t_crossattn": torch.randn((b, 12, 2048), generator=g).to(dev), "concat": torch.randn((b, 5, 64, 64), generator=g).to(dev)}
    uc = {"t_crossattn": torch.zeros((b, 12, 2048), device=dev), "concat": torch.randn((b, 5, 64, 64), generator=g).to(dev)}
    x = torch.randn((b, 4, 64, 64), generator=g).to(dev) * 14.6
    r.begin(x, cond, uc, den, LegacyDDPMDiscretization()(50))
    r.row.copy_(r.table[0:1])
    for _ in range(2):
        r._body()
    torch.cuda.synchronize()
    keep = []
    orig_empty = torch.empty

    def empty_keep(*a, **k):
        t = orig_empty(*a, **k)
        keep.append(t)
        return t

    torch.empty = empty_keep
    ops.IGEMM_TUNING = {}            # measure the cost model, not an older table
    ops.SHAPE_LOG, ops.CALL_LOG, ops.KEY_LOG = [], [], []
    r._body()
    torch.cuda.synchronize()
    torch.empty = orig_empty
    calls, keys = ops.CALL_LOG, ops.KEY_LOG
    ops.SHAPE_LOG = ops.CALL_LOG = ops.KEY_LOG = None
    groups = {}
    ki = 0
    for name, fn, args, shape in calls:
        if name != "udt_igemm":
            continue
        key = keys[ki]
        ki += 1
        groups.setdefault(key, []).append((fn, args))
    assert ki == len(keys)
    saved = 0.0
    total = 0.0
    for key, lst in groups.items():
        fn, args = lst[0]
        d = args[0]._obj
        if key == "fixed" or d.act == ops.UDT_ACT_GEGLU or d.N_out < 32:
            continue                 # GEGLU tile is fixed by the weight packing; narrow outputs use the 16-column kernel
        d.bn_hint = 0
        base = time_graph(fn, args)
        total += base * len(lst)
        best_bn, best_us = 0, base
        for bn in CANDS:
            if bn > (d.N_out + 31) // 32 * 32:
                continue
            d.bn_hint = bn
            try:
                us = time_graph(fn, args)
            except RuntimeError:
                continue
            if us < best_us:
                best_bn, best_us = bn, us
        d.bn_hint = 0
        gain = (base - best_us) * len(lst)
        if best_bn and best_us < 0.97 * base:
            table[key] = best_bn
            saved += gain
            report.append((gain, key, base, best_us, best_bn, len(lst)))
    print(f"batch {b}: tunable igemm time {total / 1e3:.3f} ms, table saves {saved / 1e3:.3f} ms per step", flush=True)


def main():
    batches = [int(a) for a in sys.argv[1:]] or [4]
    dev = torch.device("cuda", 0)
    sd = synth.synthetic_state_dict({k: s for k, s in synth.load_manifest("full").items()
                                     if k.startswith(("model.diffusion_model.", "denoiser."))}, 1234)
    from udifftext_b200.unet import UNetB200
    unet = UNetB200({k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.")}, dev,
                    **synth.ARCH["full"]["unet"])
    table, report = {}, []
    for b in batches:
        tune_batch(b, unet, dev, table, report)
    report.sort(reverse=True)
    for gain, key, base, best, bn, n in report[:40]:
        print(f"  {gain:7.1f} us/step  x{n:<3d} {base:8.2f} -> {best:8.2f} us  BN {bn:3d}  {key}")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(dict(sorted(table.items())), f, indent=0)
    print(f"{len(table)} entries -> {OUT}")


if __name__ == "__main__":
    main()
