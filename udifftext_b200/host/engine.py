"""DiffusionEngine — the object `util.init_model` builds from the model YAML and that `predict` drives.

Reference: sgm/models/diffusion.py:22-136 (`__init__`, `init_from_ckpt`, `freeze`, `decode_first_stage`,
`encode_first_stage`).  Same constructor arguments, attribute names (`model`, `denoiser`, `conditioner`,
`first_stage_model`, `loss_fn`, `scale_factor`) and `state_dict` key layout, so a reference checkpoint loads
unchanged; the weights are repacked once into the fp16 K-major layouts of the kernels when the engine is moved
to a CUDA device.  Training members (`training_step`, optimisers, EMA, `log_images`) are out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .config import default, get_obj_from_str, instantiate_from_config

OPENAIUNETWRAPPER = "sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"
UNCONDITIONAL_CONFIG = {"target": "sgm.modules.GeneralConditioner", "params": {"emb_models": []}}


class DiffusionEngine:
    def __init__(self, network_config, denoiser_config, first_stage_config, conditioner_config=None, sampler_config=None,
                 optimizer_config=None, scheduler_config=None, loss_fn_config=None, network_wrapper=None, ckpt_path=None,
                 use_ema=False, ema_decay_rate=0.9999, scale_factor=1.0, disable_first_stage_autocast=False,
                 input_key="jpg", log_keys=None, no_cond_log=False, compile_model=False, opt_keys=None):
        if use_ema:
            raise NotImplementedError("EMA weights are a training feature (out of scope)")
        self.opt_keys, self.log_keys, self.input_key = opt_keys, log_keys, input_key
        net = instantiate_from_config(network_config)
        self.model = get_obj_from_str(default(network_wrapper, OPENAIUNETWRAPPER))(net, compile_model=compile_model)
        self.denoiser = instantiate_from_config(denoiser_config)
        self.sampler = instantiate_from_config(sampler_config) if sampler_config is not None else None
        self.conditioner = instantiate_from_config(default(conditioner_config, UNCONDITIONAL_CONFIG))
        fs_cfg = dict(first_stage_config)
        fs_params = dict(fs_cfg.get("params", {}) or {})
        fs_params.setdefault("part", "decoder")   # the hot path only decodes with this copy (diffusion.py:124-129)
        self.first_stage_model = instantiate_from_config({"target": fs_cfg["target"], "params": fs_params})
        self.loss_fn = instantiate_from_config(loss_fn_config) if loss_fn_config is not None else None
        self.scale_factor = scale_factor
        self.disable_first_stage_autocast = disable_first_stage_autocast
        self.device: Optional[torch.device] = None
        self.training = False
        self._runners: Dict[tuple, object] = {}
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path)

    # ------------------------------------------------------------------------------------------ weights
    def _slots(self) -> List[Tuple[str, object]]:
        slots: List[Tuple[str, object]] = [("model.diffusion_model.", self.model.diffusion_model),
                                           ("first_stage_model.", self.first_stage_model)]
        for i, e in enumerate(getattr(self.conditioner, "embedders", [])):
            slots.append((f"conditioner.embedders.{i}.", e))
        return slots

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """keys as in the reference checkpoint: model.diffusion_model.*, first_stage_model.*,
        conditioner.embedders.{0,2}.*, denoiser.sigmas, loss_fn.g_kernel"""
        used = set()
        missing: List[str] = []
        for prefix, comp in self._slots():
            sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
            used.update(prefix + k for k in sub)
            if sub:
                comp.load_weights(sub)
            elif type(comp).__name__ != "SpatialRescaler":
                missing.append(prefix + "*")
        if "denoiser.sigmas" in sd:
            self.denoiser.sigmas = sd["denoiser.sigmas"].detach().clone().float()
            used.add("denoiser.sigmas")
        if "loss_fn.g_kernel" in sd and self.loss_fn is not None:
            self.loss_fn.g_kernel = sd["loss_fn.g_kernel"].detach().clone().float()
            used.add("loss_fn.g_kernel")
        unexpected = [k for k in sd if k not in used]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing}, unexpected {unexpected[:8]}")
        self._runners.clear()
        if self.device is not None:
            self.to(self.device)
        return missing, unexpected

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        for prefix, comp in self._slots():
            for k, v in comp.weights().items():
                out[prefix + k] = v
        out["denoiser.sigmas"] = self.denoiser.sigmas.detach().cpu()
        if self.loss_fn is not None:
            out["loss_fn.g_kernel"] = self.loss_fn.g_kernel.detach().cpu()
        return out

    def init_from_ckpt(self, path: str) -> None:
        """diffusion.py:87-105 (strict=False)"""
        if path.endswith("ckpt"):
            sd = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
        elif path.endswith("safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            raise NotImplementedError(path)
        missing, unexpected = self.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    # ------------------------------------------------------------------------------------------ module surface
    def to(self, device):
        device = torch.device(device)
        if device != self.device:
            self._runners.clear()      # cached step graphs are bound to the old device's executor and buffers
        self.device = device
        self.model.to(device)
        self.first_stage_model.to(device)
        self.conditioner.to(device)
        self.denoiser.to(device)
        if self.loss_fn is not None:
            self.loss_fn.to(device)
        return self

    def cuda(self, index: int = 0):
        return self.to(torch.device("cuda", index))

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("udifftext_b200 is inference-only (train.py / pretrain.py are out of scope)")
        return self

    def freeze(self):
        return self

    def get_input(self, batch):
        return batch[self.input_key]

    # ------------------------------------------------------------------------------------------ VAE helpers
    @torch.no_grad()
    def decode_first_stage(self, z: torch.Tensor) -> torch.Tensor:
        """diffusion.py:124-129: Decoder(post_quant_conv(z / scale_factor)), fp32 NCHW"""
        return self.first_stage_model._exec().decode(z, in_scale=1.0 / self.scale_factor)

    @torch.no_grad()
    def decode_first_stage_clamped(self, z: torch.Tensor) -> torch.Tensor:
        """decode + test.py:38's clamp((x+1)/2, 0, 1) fused into the decoder's output conversion"""
        vae = self.first_stage_model._exec()
        if z.is_cuda:       # one graph launch instead of ~160 eager calls; the graph's output buffer is static -> hand out a copy
            return vae.graphed("decode")(z.float().contiguous(), in_scale=1.0 / self.scale_factor, out_scale=0.5, out_shift=0.5,
                                         clamp01=True).clone()
        return vae.decode(z, in_scale=1.0 / self.scale_factor, out_scale=0.5, out_shift=0.5, clamp01=True)

    @torch.no_grad()
    def encode_first_stage(self, x: torch.Tensor) -> torch.Tensor:
        """diffusion.py:131-136 (needs a first-stage model built with part='both')"""
        return self.scale_factor * self.first_stage_model.encode(x)
