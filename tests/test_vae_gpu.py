"""AutoencoderKL parity on the GPU: the kernel-built encoder / decoder (udifftext_b200.vae) against the fp32
restatement (oracle/restated.py, pinned against the unmodified reference by oracle/make_golden.py) evaluated on
the same device with the same seeded synthetic weights.  Tolerances are fp16-storage tolerances."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _vae_sd(name, prefix):
    from udifftext_b200 import synth
    man = {k: v for k, v in synth.load_manifest(name).items() if k.startswith(prefix)}
    sd = synth.synthetic_state_dict(man, 1234)
    return {k[len(prefix):]: v for k, v in sd.items()}


@pytest.mark.parametrize("name,hw", [("tiny", 64), ("tiny", 96), ("full", 256)])
def test_vae_encode_matches_oracle(udt_lib, name, hw):
    from oracle import restated as R
    from udifftext_b200 import synth
    from udifftext_b200.vae import VAEB200
    dev = torch.device("cuda", 0)
    sd = _vae_sd(name, "conditioner.embedders.2.model.")
    vae = VAEB200(sd, dev, build_decoder=False, **synth.ARCH[name]["vae"])
    g = torch.Generator().manual_seed(hw)
    x = (torch.rand((2, 3, hw, hw), generator=g) * 2 - 1)
    x[:, :, hw // 4: hw // 2] = 0.0  # masked region like the LatentEncoder input
    got = vae.encode_moments(x)
    with torch.no_grad():
        ref = R.vae_encode_moments({k: v.to(dev) for k, v in sd.items()}, x.to(dev))
    torch.cuda.synchronize()
    err = _rel(got, ref)
    print(f"vae encode {name} {hw}: rel-L2 {err:.3e}")
    assert tuple(got.shape) == tuple(ref.shape)
    assert err < 2.2e-3      # fp16 storage through 11 ResnetBlocks: 1.5 x the 1.44e-3 measured on B200


@pytest.mark.parametrize("name,lat", [("tiny", 8), ("tiny", 12), ("full", 32)])
def test_vae_decode_matches_oracle(udt_lib, name, lat):
    from oracle import restated as R
    from udifftext_b200 import synth
    from udifftext_b200.vae import VAEB200
    dev = torch.device("cuda", 0)
    sd = _vae_sd(name, "first_stage_model.")
    vae = VAEB200(sd, dev, build_encoder=False, **synth.ARCH[name]["vae"])
    g = torch.Generator().manual_seed(lat)
    z = torch.randn((3, 4, lat, lat), generator=g) * 0.18215 * 4
    got = vae.decode(z, in_scale=1.0 / 0.18215, out_scale=0.5, out_shift=0.5, clamp01=True, chunk=2)
    with torch.no_grad():
        ref = torch.clamp((R.vae_decode({k: v.to(dev) for k, v in sd.items()}, z.to(dev) / 0.18215) + 1.0) / 2.0, 0.0, 1.0)
    torch.cuda.synchronize()
    err = _rel(got, ref)
    print(f"vae decode {name} {lat}: rel-L2 {err:.3e} max-abs {(got - ref).abs().max().item():.3e}")
    assert tuple(got.shape) == tuple(ref.shape)
    assert err < 8e-4        # 1.5 x the 5.3e-4 measured on B200 (decoded pixels in [0, 1])
