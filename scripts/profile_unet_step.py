"""Per-call device time of one UNet CFG step (eager replay, CUDA events around every C-ABI call), grouped by entry
point and problem shape.  Tuning aid; the contract numbers come from bench.py.
usage: python scripts/profile_unet_step.py [batch=4] [reps=5]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import api, ops, synth  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    sd = {k: v for k, v in synth.synthetic_state_dict(
        {k: s for k, s in synth.load_manifest("full").items() if k.startswith(("model.diffusion_model.", "denoiser."))}, 1234).items()}
    from udifftext_b200.unet import UNetB200
    from udifftext_b200.host.runner import StepRunner
    from udifftext_b200.host.schedule import DiscreteDenoiser, LegacyDDPMDiscretization
    unet = UNetB200({k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.")}, dev,
                    **synth.ARCH["full"]["unet"])
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
    ops.SHAPE_LOG = []
    for _ in range(2):
        r._body()
    torch.cuda.synchronize()
    acc = {}
    for _ in range(reps):
        ops.SHAPE_LOG = []
        ops._prof = []
        torch.cuda._sleep(40_000_000)  # GPU busy while the host enqueues: events then time pure device execution
        r._body()
        torch.cuda.synchronize()
        for (name, e0, e1), shape in zip(ops._prof, ops.SHAPE_LOG):
            a = acc.setdefault((name, shape), [0.0, 0])
            a[0] += e0.elapsed_time(e1) / reps
            a[1] += 1
        ops._prof = None
    rows = sorted(acc.items(), key=lambda kv: -kv[1][0])
    total = sum(v[0] for v in acc.values())
    print(f"batch {b}: eager step sum {total:.3f} ms")
    for (name, shape), (ms, calls) in rows[:60]:
        extra = ""
        if name == "udt_igemm":
            m, n, k = shape[0], shape[1], shape[2]
            extra = f"  {2.0 * m * n * k * (calls // reps) / ms / 1e9:8.1f} TFLOP/s"
        print(f"{ms:8.4f} ms  x{calls // reps:<3d} {name:22s} {shape}{extra}")
    r.step(0)  # capture + replay
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        r.step(i)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"batch": b, "graph_step_ms": e0.elapsed_time(e1) / 10, "launches_per_step": r.launches_per_step}))


if __name__ == "__main__":
    main()
