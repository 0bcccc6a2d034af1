"""EulerEDMSampler — the "DDIM" of UDiffText: Euler steps in sigma space over the LegacyDDPM schedule with
classifier-free guidance (eta = 0, s_churn = 0).

Reference: sgm/modules/diffusionmodules/sampling.py:28-78 (BaseDiffusionSampler), :89-137 (EDMSampler), :218-420
(EulerEDMSampler: get_init_noise :264-322, sampler_step :324-353, __call__ :355-420).  Same constructor and call
signatures.  `__call__` runs the fused, CUDA-graphed StepRunner when the configuration is the shipped one
(VanillaCFG + DiscreteDenoiser/EpsScaling + UnifiedUNetModel behind OpenAIWrapper) and otherwise falls back to the
generic per-op formulation below (same kernels, more launches) — there is no non-kernel path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import rng
from .config import default, instantiate_from_config
from .network import OpenAIWrapper, UnifiedUNetModel
from .runner import StepRunner
from .schedule import DiscreteDenoiser, EpsScaling, VanillaCFG, append_dims, to_d

DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


class BaseDiffusionSampler:
    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False, device="cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(default(guider_config, DEFAULT_GUIDER))
        self.verbose = verbose
        self.device = device

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device=self.device)
        uc = default(uc, cond)
        x *= torch.sqrt(1.0 + sigmas[0] ** 2.0)  # in place on the caller's tensor, like the reference (:54)
        return x, x.new_ones([x.shape[0]]), sigmas, len(sigmas), cond, uc

    def denoise(self, x, model, sigma, cond, uc):
        denoised = model.denoiser(model.model, *self.guider.prepare_inputs(x, sigma, cond, uc))
        return self.guider(denoised, sigma)

    def get_sigma_gen(self, num_sigmas, init_step=0):
        gen = range(init_step, num_sigmas - 1)
        if self.verbose:
            print(f"Sampler: {type(self).__name__} | Discretization: {type(self.discretization).__name__} | "
                  f"Guider: {type(self.guider).__name__} | {num_sigmas - 1 - init_step} steps")
        return gen


class SingleStepDiffusionSampler(BaseDiffusionSampler):
    def euler_step(self, x, d, dt):
        return x + dt * d


class EDMSampler(SingleStepDiffusionSampler):
    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def _gamma(self, sigma_i, num_sigmas):
        return min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1) if self.s_tmin <= sigma_i <= self.s_tmax else 0.0


class EulerEDMSampler(EDMSampler):
    def save_segment_map(self, attn_maps, tokens=None, save_name=None):
        """sampling.py:254-262: the first len(tokens) per-token maps -> ./temp/seg_map/seg_<name>.npy (demo.py:105)"""
        import os
        section = np.stack([attn_maps[i] for i in range(len(tokens))])
        os.makedirs("./temp/seg_map", exist_ok=True)
        np.save(f"./temp/seg_map/seg_{save_name}.npy", section)

    # ------------------------------------------------------------------------------------------ fused path
    def _fused_ok(self, model) -> bool:
        net = getattr(model.model, "diffusion_model", None)
        return (isinstance(self.guider, VanillaCFG) and isinstance(model.denoiser, DiscreteDenoiser)
                and isinstance(model.denoiser.scaling, EpsScaling) and model.denoiser.quantize_c_noise
                and isinstance(model.model, OpenAIWrapper) and isinstance(net, UnifiedUNetModel)
                and type(self.guider.dyn_thresh).__name__ == "NoDynamicThresholding" and self.s_churn == 0.0)

    MAX_RUNNERS = 4    # cached step graphs per engine (each owns the UNet step's activations): least recently used goes

    def _runner(self, model, x, cond) -> StepRunner:
        """the StepRunner (static buffers + captured step graph) for this request shape.  The guidance scale is NOT part
        of the key — the Euler kernel reads it from the per-step device row — and the cache is a small LRU, so a demo
        whose sliders change scale / num_samples / resolution per request does not accumulate graphs."""
        b, _, h, w = x.shape
        ctx_len = cond["t_crossattn"].shape[1]
        key = (b, h, w, ctx_len)
        r = model._runners.pop(key, None)
        if r is None:
            r = StepRunner(model.model.diffusion_model._exec(), b, h, w, ctx_len, self.guider.scale)
            while len(model._runners) >= self.MAX_RUNNERS:
                model._runners.pop(next(iter(model._runners)))
        model._runners[key] = r          # (re)inserted last = most recently used
        return r

    # ------------------------------------------------------------------------------------------ noise search
    def get_init_noise(self, cfgs, model, cond, batch, uc=None):
        """sampling.py:264-322: draw the initial latent noise; with `noise_iters` > 0 run that many 2-step trial
        samplings and keep the noise whose final textual-attention local loss is lowest.  The reference can only
        score one image (`local_loss.item()`); here every image of the batch keeps its own best noise."""
        hh, ww = batch["target_size_as_tuple"][0]
        dev = torch.device("cuda", index=cfgs.gpu) if not isinstance(self.device, torch.device) else self.device
        shape = (cfgs.batch_size, cfgs.channel, int(hh) // cfgs.factor, int(ww) // cfgs.factor)
        randn = rng.randn(shape, dev)
        iters = int(getattr(cfgs, "noise_iters", 0) or 0)
        if iters == 0:
            return randn
        verbose, self.verbose = self.verbose, False
        fused = self._fused_ok(model)
        best_noise, best_loss = randn.clone(), torch.full((shape[0],), float("inf"), device=dev)
        worst = torch.full((shape[0],), -float("inf"), device=dev)
        trial_losses = []
        for _ in range(iters):
            x = randn.clone()
            x, _, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps=2)
            if fused:
                runner = self._runner(model, x, cond)
                runner.begin(x, cond, uc, model.denoiser, sigmas, self.s_churn, self.s_tmin, self.s_tmax,
                             cfg_scale=self.guider.scale)
                for i in range(num_sigmas - 1):
                    runner.step(i, export_attn_maps=True)
                loss = model.loss_fn.get_min_local_loss(model.model.diffusion_model.attn_map_cache, batch["mask"],
                                                        batch["seg_mask"])
                loss = loss[loss.shape[0] // 2:]                  # conditional half of the CFG batch (:341)
            else:                                                 # any other guider / denoiser / churn: the reference's loop
                s_in = x.new_ones([x.shape[0]])
                for i in range(num_sigmas - 1):
                    x, _, loss = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], model, x, cond, batch, uc,
                                                   self._gamma(sigmas[i], num_sigmas), save_loss=True)
            trial_losses.append(loss)
            better = loss < best_loss
            best_noise[better] = randn[better]
            best_loss = torch.minimum(best_loss, loss)
            worst = torch.maximum(worst, loss)
            randn = rng.randn(shape, dev)                         # a fresh draw per iteration (:311), used or not
        self.verbose = verbose
        self.last_init_losses = torch.stack(trial_losses)         # [noise_iters, B], for inspection / tests
        print(f"Init local loss: Best {best_loss.tolist()} Worst {worst.tolist()}")
        return best_noise

    # ------------------------------------------------------------------------------------------ generic step
    def sampler_step(self, sigma, next_sigma, model, x, cond, batch=None, uc=None, gamma=0.0, alpha=0, iter_enabled=False,
                     thres=None, update=False, name=None, save_loss=False, save_attn=False, save_inter=False):
        """sampling.py:324-353, generic formulation (per-op kernel launches; the fused path is `__call__`)"""
        if update:
            raise NotImplementedError("attend-and-excite needs autograd through the UNet (inference kernels only)")
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            x = x + torch.randn_like(x) * self.s_noise * append_dims(sigma_hat ** 2 - sigma ** 2, x.ndim) ** 0.5
        net = model.model.diffusion_model
        if save_loss or save_attn:
            net._exec().export_attn_maps = True
        try:
            denoised = self.denoise(x, model, sigma_hat, cond, uc)
        finally:
            net._exec().export_attn_maps = False
        inter = model.decode_first_stage(denoised) if save_inter else None
        if save_loss:
            loss = model.loss_fn.get_min_local_loss(net.attn_map_cache, batch["mask"], batch["seg_mask"])
            loss = loss[loss.shape[0] // 2:]
        else:
            loss = torch.zeros(1)
        if save_attn:
            attn_map = net.save_attn_map(save_name=name, tokens=batch["label"][0])
            self.save_segment_map(attn_map, tokens=batch["label"][0], save_name=name)
        d = to_d(x, sigma_hat, denoised)
        return self.euler_step(x, d, append_dims(next_sigma - sigma_hat, x.ndim)), inter, loss

    # ------------------------------------------------------------------------------------------ the hot loop
    def __call__(self, model, x, cond, batch=None, uc=None, num_steps=None, init_step=0, name=None, aae_enabled=False,
                 detailed=False):
        """sampling.py:355-420"""
        if aae_enabled:
            raise NotImplementedError("aae_enabled (attend-and-excite) needs autograd through the UNet; out of scope")
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        if batch is not None and "name" in batch:
            name = batch["name"][0]                   # sampling.py:362
        if not self._fused_ok(model):
            for i in self.get_sigma_gen(num_sigmas, init_step):
                gamma = self._gamma(sigmas[i], num_sigmas)
                x, _, _ = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], model, x, cond, batch, uc, gamma,
                                            name=name, save_attn=detailed and (i == (num_sigmas - 1) // 2))
            return x
        runner = self._runner(model, x, cond)
        runner.begin(x, cond, uc, model.denoiser, sigmas, self.s_churn, self.s_tmin, self.s_tmax,
                     cfg_scale=self.guider.scale)
        net = model.model.diffusion_model
        for i in self.get_sigma_gen(num_sigmas, init_step):
            save_attn = detailed and (i == (num_sigmas - 1) // 2)
            runner.step(i, export_attn_maps=save_attn)
            if save_attn:                             # sampling.py:344-346: the files demo.py:104-105 reads back
                attn_map = net.save_attn_map(save_name=name, tokens=batch["label"][0])
                self.save_segment_map(attn_map, tokens=batch["label"][0], save_name=name)
        self.last_runner = runner
        return runner.result()
