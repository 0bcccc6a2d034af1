"""True device time of every distinct call of one UNet CFG step: each (entry point, problem shape) is replayed 20x
back to back inside a CUDA graph (PDL active, no host launch cost, no event overhead) and the per-launch time is
multiplied by its call count.  Tuning aid.  usage: python scripts/profile_step_graph.py [batch=4]"""
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import ops, synth  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda", 0)
    sd = synth.synthetic_state_dict({k: s for k, s in synth.load_manifest("full").items()
                                     if k.startswith(("model.diffusion_model.", "denoiser."))}, 1234)
    from udifftext_b200.host.runner import StepRunner
    from udifftext_b200.host.schedule import DiscreteDenoiser, LegacyDDPMDiscretization
    from udifftext_b200.unet import UNetB200
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
    for _ in range(2):
        r._body()
    torch.cuda.synchronize()
    keep = []                        # keep every tensor of the logged step alive: the raw pointers stay valid
    orig_empty = torch.empty

    def empty_keep(*a, **k):
        t = orig_empty(*a, **k)
        keep.append(t)
        return t

    torch.empty = empty_keep
    ops.SHAPE_LOG = []
    ops.CALL_LOG = []
    r._body()
    torch.cuda.synchronize()
    torch.empty = orig_empty
    calls = ops.CALL_LOG
    ops.CALL_LOG = None
    ops.SHAPE_LOG = None
    stream = torch.cuda.current_stream()
    groups = {}
    for name, fn, args, shape in calls:
        groups.setdefault((name, shape), []).append((fn, args))
    rows = []
    iters = 20
    for (name, shape), lst in groups.items():
        fn, args = lst[0]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(iters):
                rc = fn(*args, st)
                assert rc == 0, (name, shape, rc)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        rows.append((us * len(lst), us, len(lst), name, shape))
    rows.sort(reverse=True)
    total = sum(r_[0] for r_ in rows)
    by = {}
    for tot, us, n, name, shape in rows:
        by[name] = by.get(name, 0.0) + tot
    print(f"batch {b}: sum of per-shape graph times {total / 1e3:.3f} ms  " + "  ".join(f"{k}={v / 1e3:.3f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
    for tot, us, n, name, shape in rows[:70]:
        extra = ""
        if name == "udt_igemm":
            m, nn, k = shape[0], shape[1], shape[2]
            extra = f"  {2.0 * m * nn * k / us / 1e6:7.0f} TFLOP/s"
        print(f"{tot / 1e3:8.4f} ms  x{n:<3d} {us:8.2f} us  {name:22s} {shape}{extra}")
    r.step(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"whole-step graph replay: {e0.elapsed_time(e1) / 5:.3f} ms  ({r.launches_per_step} C-ABI calls)")


if __name__ == "__main__":
    main()
