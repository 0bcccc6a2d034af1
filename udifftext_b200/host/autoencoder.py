"""AutoencoderKL wrappers with the reference's surface (sgm/models/autoencoder.py:282-321,
sgm/modules/distributions/distributions.py:24-72); compute is `udifftext_b200.vae.VAEB200`."""
from __future__ import annotations

from typing import Optional

import torch

from ..vae import VAEB200
from . import rng
from .network import _Component


class DiagonalGaussianDistribution:
    """distributions.py:24-72 over NCHW moments [B, 2z, h, w]; `sample()` draws on the CPU generator and moves the
    noise to the device, like the reference (this fixes the RNG stream: SURVEY.md §3.1)."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self) -> torch.Tensor:
        return self.mean + self.std * rng.randn(self.mean.shape, self.parameters.device)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(_Component):
    """ddconfig keys used: ch, ch_mult, num_res_blocks, z_channels, in_channels, out_ch (model.py:482-743).
    `part` restricts the executor to the half that the hot path needs from this instance."""

    def __init__(self, ddconfig=None, embed_dim: int = 4, lossconfig=None, ckpt_path: Optional[str] = None,
                 monitor=None, input_key: str = "jpg", part: str = "both", **unused):
        super().__init__()
        dd = dict(ddconfig or {})
        if dd.get("attn_resolutions"):
            raise NotImplementedError("AutoencoderKL on B200: attn_resolutions must be empty (mid attention only)")
        self.arch = {k: dd[k] for k in ("ch", "ch_mult", "num_res_blocks", "z_channels", "in_channels", "out_ch") if k in dd}
        self.arch["ch_mult"] = list(self.arch.get("ch_mult", (1, 2, 4, 4)))
        self.embed_dim = embed_dim
        self.part = part
        self.exec: Optional[VAEB200] = None
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path)

    def init_from_ckpt(self, path: str) -> None:
        """autoencoder.py:49-72 (strict=False load of a .ckpt / .safetensors autoencoder checkpoint)"""
        if path.endswith("ckpt"):
            sd = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
        elif path.endswith("safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            raise NotImplementedError(path)
        self.load_weights(sd)

    def _invalidate(self):
        self.exec = None

    def _materialise(self, device):
        self.exec = VAEB200(self._require_weights(), device, build_encoder=self.part in ("both", "encoder"),
                            build_decoder=self.part in ("both", "decoder"), **self.arch)

    def _exec(self) -> VAEB200:
        if self.exec is None:
            raise RuntimeError("AutoencoderKL: call .to(cuda device) after loading weights")
        return self.exec

    def moments(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder + quant_conv: fp32 NCHW [B, 2*z, h, w]"""
        return self._exec().encode_moments(x)

    def encode(self, x: torch.Tensor) -> DiagonalGaussianDistribution:
        """autoencoder.py:304-311: returns the posterior"""
        return DiagonalGaussianDistribution(self.moments(x))

    def decode(self, z: torch.Tensor, **decoder_kwargs) -> torch.Tensor:
        """autoencoder.py:313-316"""
        return self._exec().decode(z)


class AutoencoderKLInferenceWrapper(AutoencoderKL):
    """autoencoder.py:319-321: `encode` returns a posterior SAMPLE"""

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return super().encode(x).sample()
