"""FullLoss — only the inference-time member: `get_min_local_loss`, the attention-map score the noise search and
`detailed` mode read (reference: sgm/modules/diffusionmodules/loss.py:73-129 constructor / Gaussian kernel,
:192-235 get_min_local_loss).  The training losses are out of scope."""
from __future__ import annotations

import math
from typing import List

import torch


class FullLoss:
    def __init__(self, seq_len=12, kernel_size=3, gaussian_sigma=0.5, min_attn_size=16, lambda_local_loss=0.0,
                 lambda_ocr_loss=0.0, lambda_style_loss=0.0, ocr_enabled=False, style_enabled=False,
                 predictor_config=None, sigma_sampler_config=None, type="l2", offset_noise_level=0.0,
                 batch2model_keys=None, **unused):
        # the OCR loss term (loss.py:150-160) belongs to training; the predictor itself is available for evaluation
        # (host/predictor.py) and is instantiated here only to honour `predictor_config` of the model YAML
        self.ocr_enabled = bool(ocr_enabled)
        self.predictor = None
        if ocr_enabled and predictor_config is not None:
            from .config import instantiate_from_config
            self.predictor = instantiate_from_config(predictor_config)
        self.gaussian_kernel_size = kernel_size
        self.min_attn_size = min_attn_size
        self.g_kernel = self.get_gaussian_kernel(kernel_size, gaussian_sigma, seq_len)

    @staticmethod
    def get_gaussian_kernel(kernel_size=3, sigma=1.0, out_channels=3) -> torch.Tensor:
        """normalised 2-D Gaussian, one copy per token channel: [out_channels, 1, k, k] (loss.py:103-129)"""
        ax = torch.arange(kernel_size, dtype=torch.float32)
        yy, xx = torch.meshgrid(ax, ax, indexing="ij")
        mean = (kernel_size - 1) / 2.0
        g = torch.exp(-((xx - mean) ** 2 + (yy - mean) ** 2) / (2.0 * sigma ** 2)) / (2.0 * math.pi * sigma ** 2)
        g = g / g.sum()
        return g.view(1, 1, kernel_size, kernel_size).repeat(out_channels, 1, 1, 1)

    def to(self, device):
        self.g_kernel = self.g_kernel.to(device)
        return self

    def get_min_local_loss(self, attn_map_cache: List[dict], mask: torch.Tensor, seg_mask: torch.Tensor) -> torch.Tensor:
        """-min over valid tokens of max over pixels of (mask * blurred head-mean attention), averaged over the
        `t_attn` layers whose map is at least `min_attn_size` wide (loss.py:192-235).  Returns [B_unet].  One K12
        launch (`udt_attn_local_score`) per layer; CUDA tensors only — there is no CPU path.

        The reference relies on broadcasting the [1, l, n] mask against the CFG-doubled [2, l, n] maps and therefore
        only works for one image (`.item()`, sampling.py:307); the kernel indexes the per-image mask / seg_mask with
        `b % B_images`, which lifts that limit with identical values."""
        layers = [it for it in attn_map_cache
                  if it["name"].endswith("t_attn") and it["size"] is not None and it["size"] >= self.min_attn_size]
        if not layers:
            raise ValueError("get_min_local_loss: no exported t_attn map of at least min_attn_size")
        dev = layers[0]["attn_map"].device
        if dev.type != "cuda":
            raise RuntimeError("FullLoss.get_min_local_loss runs on the CUDA kernels only (no CPU path)")
        from .. import ops
        mask = mask.to(dev, torch.float32).contiguous()
        seg = seg_mask.to(dev, torch.float32).contiguous()
        gk = self.g_kernel[0, 0].to(dev, torch.float32).contiguous()
        b_unet = layers[0]["attn_map"].shape[0] // layers[0]["heads"]
        assert b_unet % mask.shape[0] == 0 and seg.shape[0] == mask.shape[0]
        score = torch.zeros((b_unet,), device=dev, dtype=torch.float32)
        for it in layers:
            assert seg.shape[1] <= it["attn_map"].shape[2]
            ops.attn_local_score(it["attn_map"].contiguous(), mask, seg, gk, score, it["heads"], it["size"])
        return score / len(layers)
