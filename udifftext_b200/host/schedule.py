"""Noise schedule, denoiser parameterisation and guidance — the small host-side pieces of the sampler.

Reference: sgm/modules/diffusionmodules/discretizer.py:10-68 (LegacyDDPMDiscretization), util.py:19-32
(make_beta_schedule "linear"), denoiser.py:6-63 (Denoiser / DiscreteDenoiser), denoiser_scaling.py:16-22
(EpsScaling), denoiser_weighting.py (EpsWeighting), guiders.py:8-54 (VanillaCFG / IdentityGuider),
sampling_utils.py:7-9,39-40.  All of this is float64 / fp32 arithmetic on a handful of scalars per step, done on
the host exactly as the reference does it; the per-pixel math it parameterises runs in the K7 kernels.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .config import default, instantiate_from_config


def make_beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2) -> np.ndarray:
    """'linear' = linear in sqrt(beta) (diffusionmodules/util.py:19-32), float64"""
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule '{schedule}' is not used by the UDiffText inference path")
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def append_zero(x: torch.Tensor) -> torch.Tensor:
    return torch.cat([x, x.new_zeros([1])])


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]


class Discretization:
    """discretizer.py:15-24: callable returning `n` sigmas (descending), optionally with 0 appended / flipped"""

    def __call__(self, n: int, do_append_zero: bool = True, device="cpu", flip: bool = False) -> torch.Tensor:
        sig = self.get_sigmas(n, device=device)
        if do_append_zero:
            sig = append_zero(sig)
        return torch.flip(sig, (0,)) if flip else sig

    def get_sigmas(self, n: int, device="cpu") -> torch.Tensor:
        raise NotImplementedError


class LegacyDDPMDiscretization(Discretization):
    """sigma_t = sqrt((1 - abar_t) / abar_t) of the SD beta schedule, sub-sampled at
    linspace(T-1, 0, n, endpoint=False).astype(int)[::-1] (discretizer.py:10-13,41-68)"""

    def __init__(self, linear_start: float = 0.00085, linear_end: float = 0.0120, num_timesteps: int = 1000):
        self.num_timesteps = num_timesteps
        betas = make_beta_schedule("linear", num_timesteps, linear_start=linear_start, linear_end=linear_end)
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)

    def timesteps(self, n: int) -> np.ndarray:
        """ascending training-timestep indices visited by an n-step schedule"""
        if n < self.num_timesteps:
            return np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
        if n == self.num_timesteps:
            return np.arange(self.num_timesteps)
        raise ValueError(f"{n} sampling steps > {self.num_timesteps} training timesteps")

    def get_sigmas(self, n: int, device="cpu") -> torch.Tensor:
        abar = self.alphas_cumprod[self.timesteps(n)]
        sig = torch.tensor((1 - abar) / abar, dtype=torch.float32, device=device) ** 0.5  # fp32 sqrt, like the reference
        return torch.flip(sig, (0,))


class EpsScaling:
    """denoiser_scaling.py:16-22: c_skip = 1, c_out = -sigma, c_in = 1/sqrt(sigma^2+1), c_noise = sigma"""

    def __call__(self, sigma: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        c_skip = torch.ones_like(sigma, device=sigma.device)
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = sigma.clone()
        return c_skip, c_out, c_in, c_noise


class EpsWeighting:
    """denoiser_weighting.py: w(sigma) = sigma^-2 (training only; kept because the model YAML names it)"""

    def __call__(self, sigma: torch.Tensor) -> torch.Tensor:
        return sigma ** -2.0


class Denoiser:
    """denoiser.py:6-28: D(x; sigma) = net(x * c_in, c_noise, cond) * c_out + x * c_skip"""

    def __init__(self, weighting_config, scaling_config):
        self.weighting = instantiate_from_config(weighting_config)
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def w(self, sigma):
        return self.weighting(sigma)

    def __call__(self, network, input, sigma, cond):
        sigma = self.possibly_quantize_sigma(sigma)
        shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(shape))
        return network(input * c_in, c_noise, cond) * c_out + input * c_skip

    def to(self, device):
        return self


class DiscreteDenoiser(Denoiser):
    """denoiser.py:31-63: sigma is snapped to the nearest of `num_idx` table entries and c_noise becomes the table
    index (the integer timestep the UNet was trained on).  `sigmas` is the registered buffer `denoiser.sigmas`."""

    def __init__(self, weighting_config, scaling_config, num_idx, discretization_config, do_append_zero=False,
                 quantize_c_noise=True, flip=True):
        super().__init__(weighting_config, scaling_config)
        self.sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.quantize_c_noise = quantize_c_noise

    def to(self, device):
        self.sigmas = self.sigmas.to(device)
        return self

    def sigma_to_idx(self, sigma: torch.Tensor) -> torch.Tensor:
        table = self.sigmas.to(sigma.device)
        return (sigma - table[:, None]).abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx: torch.Tensor) -> torch.Tensor:
        return self.sigmas.to(idx.device)[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise


class NoDynamicThresholding:
    """sampling_utils.py:7-9"""

    def __call__(self, uncond, cond, scale):
        return uncond + scale * (cond - uncond)


class VanillaCFG:
    """guiders.py:8-40: classifier-free guidance evaluated as one doubled batch, unconditional half FIRST"""

    CAT_KEYS = ("vector", "t_crossattn", "v_crossattn", "concat")

    def __init__(self, scale, dyn_thresh_config=None):
        self.scale = scale
        self.scale_schedule = lambda sigma: self.scale  # independent of the step
        self.dyn_thresh = instantiate_from_config(default(dyn_thresh_config, {
            "target": "sgm.modules.diffusionmodules.sampling_utils.NoDynamicThresholding"}))

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        return self.dyn_thresh(x_u, x_c, self.scale_schedule(sigma))

    def prepare_inputs(self, x, s, c: Dict, uc: Dict):
        merged = {}
        for k in c:
            if k in self.CAT_KEYS:
                merged[k] = torch.cat((uc[k], c[k]), 0)
            else:
                assert c[k] == uc[k]
                merged[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), merged


class IdentityGuider:
    """guiders.py:43-54"""

    def __call__(self, x, sigma):
        return x

    def prepare_inputs(self, x, s, c, uc):
        return x, s, dict(c)


class DiscreteSampling:
    """sigma_sampling.py (training-time sigma sampler named by the model YAML's loss_fn_config; kept constructible)"""

    def __init__(self, discretization_config, num_idx, do_append_zero=False, flip=True):
        self.num_idx = num_idx
        self.sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero, flip=flip)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def __call__(self, n_samples, rand=None):
        idx = default(rand, torch.randint(0, self.num_idx, (n_samples,)))
        return self.idx_to_sigma(idx)


def to_d(x, sigma, denoised):
    """sampling_utils.py:39-40"""
    return (x - denoised) / append_dims(sigma, x.ndim)


def step_constants(denoiser: DiscreteDenoiser, sigmas: torch.Tensor, s_churn: float = 0.0, s_tmin: float = 0.0,
                   s_tmax: float = float("inf")) -> Dict[str, torch.Tensor]:
    """Per-step scalars of the Euler / eps-parameterisation update, computed on the host in fp32 exactly as the
    reference's tensor code does (sampling.py:324-353 with gamma from :371-375; denoiser.py:22-28):
      idx[i]    = nearest-table index of sigma_hat_i  (the UNet timestep),
      c_in[i]   = 1 / sqrt(sigma_q^2 + 1),  sigma_q = table[idx[i]],
      dsigma[i] = sigma_{i+1} - sigma_hat_i,  eps_scale[i] = sigma_q / sigma_hat_i  (1 when sigma is on the table).
    With these, x_{i+1} = x_i + dsigma[i] * eps_scale[i] * cfg(eps) — algebraically the reference's
    `x + dt * (x - (x - sigma_q * eps)) / sigma_hat`."""
    sig = sigmas.detach().to("cpu", torch.float32)
    n = sig.numel() - 1
    gamma = torch.zeros(n, dtype=torch.float32)
    if s_churn > 0:
        for i in range(n):
            if s_tmin <= float(sig[i]) <= s_tmax:
                gamma[i] = min(s_churn / n, 2 ** 0.5 - 1)
    sigma_hat = sig[:-1] * (gamma + 1.0)
    table = denoiser.sigmas.detach().to("cpu", torch.float32)
    idx = (sigma_hat[None, :] - table[:, None]).abs().argmin(dim=0)
    sigma_q = table[idx]
    c_in = 1 / (sigma_q ** 2 + 1.0) ** 0.5
    return {"idx": idx, "sigma_hat": sigma_hat, "sigma_q": sigma_q, "c_in": c_in, "dsigma": sig[1:] - sigma_hat,
            "eps_scale": sigma_q / sigma_hat, "gamma": gamma}
