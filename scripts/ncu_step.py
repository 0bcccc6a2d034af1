"""One eager UNet CFG step (batch 4) bracketed by cudaProfilerStart/Stop — the profiling target for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... python scripts/ncu_step.py
(the launch list with DRAM traffic of every kernel of the step; see profiles/README.md)."""
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import synth  # noqa: E402


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
    torch.cuda.profiler.start()
    r._body()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
