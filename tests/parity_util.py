"""Shared parity harness (TEST INFRASTRUCTURE): runs one `predict` of the product step by step next to the fp32 oracle
(oracle/restated.py) on the same device, seeds and weights, and reports the error of every stage.

Used by tests/test_parity_c2_gpu.py (gates) and tests/parity_report.py (writes profiles/parity_r02.json).  The oracle
is evaluated with TF32 disabled (`fp32_oracle()`): TF32 has fp16's 10-bit mantissa, so a TF32 "fp32" oracle carries an
error of the size being measured.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def quantised_oracle(R, mode: str):
    """context manager: oracle GEMMs / convs with operands rounded to fp16 (emulates fp16 storage, fp32 accumulation)
    mode: 'w' weights only; 'wa' weights + GEMM/conv inputs; 'was' + every GEMM/conv output (all stored activations)"""
    import contextlib

    q = lambda t: t.half().float()

    @contextlib.contextmanager
    def cm():
        lin, conv = R._lin, R._conv

        def _lin(sd, name, x):
            w = q(sd[name + ".weight"])
            xi = q(x) if "a" in mode else x
            y = F.linear(xi, w, sd.get(name + ".bias"))
            return q(y) if "s" in mode else y

        def _conv(sd, name, x, stride=1, padding=1):
            w = q(sd[name + ".weight"])
            xi = q(x) if "a" in mode else x
            y = F.conv2d(xi, w, sd.get(name + ".bias"), stride=stride, padding=padding)
            return q(y) if "s" in mode else y

        R._lin, R._conv = _lin, _conv
        try:
            yield
        finally:
            R._lin, R._conv = lin, conv
    return cm()


@contextlib.contextmanager
def fp32_oracle():
    """true-fp32 cuDNN / cuBLAS for the oracle (no TF32 anywhere)"""
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev[0], prev[1]
        torch.set_float32_matmul_precision(prev[2])


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def maxabs(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.double() - b.double()).abs().max().item()


def guided_eps_product(runner) -> torch.Tensor:
    """guided eps [B,4,h,w] (NCHW fp32) from the runner's last UNet output (fp32 NHWC [2B,h,w,4], uc half first);
    the product's Euler kernel applies x += dsigma * (e_u + s (e_c - e_u)) (guiders.py:25-29, sampling.py:349-351)"""
    e = runner.eps.float()
    b = e.shape[0] // 2
    g = e[:b] + runner.cfg_scale * (e[b:] - e[:b])
    return g.permute(0, 3, 1, 2).contiguous()


def oracle_run(R, sd_dev: Dict[str, torch.Tensor], batch_dev: dict, steps: int, scale: float, seed: int,
               scale_factor: float = 0.18215, phase_ctx: Optional[dict] = None) -> dict:
    """oracle predict with every intermediate kept: c / uc, x before every step, guided eps of every step, z, pixels.
    `phase_ctx` = {"cond" | "unet" | "dec": context-manager factory} wraps single phases (precision studies)."""
    null = contextlib.nullcontext
    pc = lambda name: (phase_ctx or {}).get(name, null)()
    b, _, hh, ww = batch_dev["image"].shape
    dev = batch_dev["image"].device
    lat = (b, 4, hh // 8, ww // 8)
    torch.manual_seed(seed)
    noise_c, noise_uc = torch.randn(lat).to(dev), torch.randn(lat).to(dev)
    with torch.no_grad(), fp32_oracle():
        with pc("cond"):
            c, uc = R.conditioner(R._sub(sd_dev, "conditioner."), batch_dev, noise_c, noise_uc, scale_factor)
        x = torch.randn(lat).to(dev)
        unet_sd = R._sub(sd_dev, "model.diffusion_model.")
        sig = R.sampler_sigmas(steps)
        table = R.denoiser_sigmas()
        x = x * torch.sqrt(1.0 + sig[0] ** 2.0)
        xs, eps_l = [], []
        for i in range(steps):
            xs.append(x.clone())
            with pc("unet"):
                eps = R.cfg_denoise_eps(unet_sd, x, float(sig[i]), table, c, uc, scale)
            eps_l.append(eps)
            x = x + (sig[i + 1] - sig[i]) * eps
        with pc("dec"):
            img = R.vae_decode(R._sub(sd_dev, "first_stage_model."), x / scale_factor)
        pixels = torch.clamp((img + 1.0) / 2.0, 0.0, 1.0)
    return {"c": c, "uc": uc, "xs": xs, "eps": eps_l, "z": x, "pixels": pixels, "sigmas": sig}


def product_run(api, model, batch_host: dict, steps: int, scale: float, seed: int, teacher_xs: Optional[Sequence] = None,
                teacher_steps: Sequence[int] = (), cond_override: Optional[tuple] = None) -> dict:
    """the product's predict (api.predict, test.py:19-40) unrolled so that the guided eps of every step can be read;
    optionally, at `teacher_steps`, the UNet is additionally evaluated on the ORACLE's x (single-step error, no drift)"""
    b = batch_host["image"].shape[0]
    cfgs = api.runtime_config(steps=steps, batch_size=b, scale=[scale, 0.0])
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    torch.manual_seed(seed)
    with torch.no_grad():
        batch, batch_uc = api.prepare_batch(cfgs, dict(batch_host))
        c, uc = model.conditioner.get_unconditional_conditioning(batch, batch_uc=batch_uc,
                                                                 force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
        if cond_override is not None:      # precision study: the ORACLE's conditioning fed to the product's sampler
            c = {k: v.clone() for k, v in cond_override[0].items()}
            uc = {k: v.clone() for k, v in cond_override[1].items()}
        x = sampler.get_init_noise(cfgs, model, cond=c, batch=batch, uc=uc)
        x, s_in, sigmas, num_sigmas, c, uc = sampler.prepare_sampling_loop(x, c, uc, None)
        runner = sampler._runner(model, x, c)
        runner.begin(x, c, uc, model.denoiser, sigmas, 0.0, 0.0, 999.0)
        eps_l, xs, teacher = [], [], {}
        for i in range(num_sigmas - 1):
            if teacher_xs is not None and i in teacher_steps:
                keep = runner.x.clone()
                runner.x.copy_(teacher_xs[i])
                runner.step(i)
                teacher[i] = guided_eps_product(runner)
                runner.x.copy_(keep)
            xs.append(runner.x.clone())
            runner.step(i)
            eps_l.append(guided_eps_product(runner))
        z = runner.result()
        pixels = model.decode_first_stage_clamped(z)
    torch.cuda.synchronize()
    return {"c": c, "uc": uc, "xs": xs, "eps": eps_l, "z": z, "pixels": pixels, "teacher_eps": teacher, "runner": runner}


def compare(prod: dict, ora: dict, check_steps: Sequence[int]) -> dict:
    out = {
        "c_concat_rel": rel(prod["c"]["concat"], ora["c"]["concat"]),
        "uc_concat_rel": rel(prod["uc"]["concat"], ora["uc"]["concat"]),
        "t_crossattn_rel": rel(prod["c"]["t_crossattn"], ora["c"]["t_crossattn"]),
        "latents_rel": rel(prod["z"], ora["z"]), "latents_maxabs": maxabs(prod["z"], ora["z"]),
        "pixels_rel": rel(prod["pixels"], ora["pixels"]), "pixels_maxabs": maxabs(prod["pixels"], ora["pixels"]),
        "eps_rel": {int(i): rel(prod["eps"][i], ora["eps"][i]) for i in check_steps},
        "x_rel": {int(i): rel(prod["xs"][i], ora["xs"][i]) for i in check_steps},
    }
    if prod.get("teacher_eps"):
        out["eps_rel_teacher_forced"] = {int(i): rel(e, ora["eps"][i]) for i, e in prod["teacher_eps"].items()}
    return out
