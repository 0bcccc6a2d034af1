"""Replays the captured UNet CFG step graph many times from the same state and checks that every replay produces
bit-identical eps / x (all statistics are reduced in a fixed order, split-K partials are summed in a fixed order): a
mismatch would reveal a race between overlapped (PDL) kernels.  usage: python scripts/stress_determinism.py [batch] [replays]"""
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import synth  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
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
    r.step(0)                      # captures the graph
    torch.cuda.synchronize()
    x0 = x.clone()
    ref_eps = ref_x = None
    bad = 0
    for i in range(n):
        r.x.copy_(x0)
        r._next = -1
        r.set_step(i % 3)          # the captured step loads its row through the device-side step counter
        r.replay()
        if i % 3 == 0:
            torch.cuda.synchronize()
            if ref_eps is None:
                ref_eps, ref_x = r.eps.clone(), r.x.clone()
            elif not (torch.equal(r.eps, ref_eps) and torch.equal(r.x, ref_x)):
                bad += 1
                d = (r.eps - ref_eps).abs()
                print(f"replay {i}: MISMATCH max |d eps| {d.max().item():.3e} in {int((d > 0).sum())} elements")
    print(f"batch {b}: {n} replays, {bad} mismatching of {(n + 2) // 3} compared; eps finite: {bool(torch.isfinite(r.eps).all())}")


if __name__ == "__main__":
    main()
