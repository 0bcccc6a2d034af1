"""FullLoss — only the inference-time member: `get_min_local_loss`, the attention-map score the noise search and
`detailed` mode read (reference: sgm/modules/diffusionmodules/loss.py:73-129 constructor / Gaussian kernel,
:192-235 get_min_local_loss).  The training losses are out of scope."""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn.functional as F


class FullLoss:
    def __init__(self, seq_len=12, kernel_size=3, gaussian_sigma=0.5, min_attn_size=16, lambda_local_loss=0.0,
                 lambda_ocr_loss=0.0, lambda_style_loss=0.0, ocr_enabled=False, style_enabled=False,
                 predictor_config=None, sigma_sampler_config=None, type="l2", offset_noise_level=0.0,
                 batch2model_keys=None, **unused):
        if ocr_enabled:
            raise NotImplementedError("OCR loss (PARSeq) is a training / evaluation component and out of scope")
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
        `t_attn` layers whose map is at least `min_attn_size` wide (loss.py:192-235).  Returns [B_unet]."""
        total, count = 0, 0
        for item in attn_map_cache:
            if not item["name"].endswith("t_attn") or item["size"] is None or item["size"] < self.min_attn_size:
                continue
            heads, size, am = item["heads"], item["size"], item["attn_map"]
            seg_l = seg_mask.shape[1]
            _, n, l = am.shape
            assert seg_l <= l
            am = am.reshape(-1, heads, n, l)[..., :seg_l].mean(dim=1).permute(0, 2, 1)            # b, l, n
            am = F.conv2d(am.reshape(-1, seg_l, size, size), self.g_kernel.to(am.device),
                          padding=self.gaussian_kernel_size // 2, groups=seg_l).reshape(-1, seg_l, n)
            mm = F.interpolate(mask.to(am.device), (size, size)).tile((1, seg_l, 1, 1)).reshape(-1, seg_l, n)
            sm = seg_mask.to(am.device)
            if am.shape[0] == 2 * mm.shape[0] and mm.shape[0] > 1:
                # CFG-doubled UNet batch [uc; c]: the reference relies on broadcasting and therefore only works for
                # one image (sampling.py:307); repeating the per-image masks lifts that limit with identical values
                mm, sm = torch.cat([mm, mm]), torch.cat([sm, sm])
            p = (mm * am).max(dim=-1)[0] + (1 - sm)
            total = total + (-p.min(dim=-1)[0])
            count += 1
        return total / count
