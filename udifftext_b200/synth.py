"""Synthetic weights and request batches (there are no checkpoints or datasets in this environment).

`synthetic_state_dict` fills a reference-format `state_dict` from a committed manifest (key -> shape) with
per-key seeded values, so the build container (which can load them into the unmodified reference) and the GPU
box (which cannot see the reference) hold bit-identical weights without shipping gigabytes of tensors.
Zero-initialised reference layers (ResBlock out conv, proj_out, t_attn.to_out, UNet out conv) get non-zero
values too — with the shipped zero init the UNet output is exactly 0 and parity would be vacuous.
"""
from __future__ import annotations

import json
import math
import os
import random
import string
import zlib
from typing import Dict, List, Optional

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
MANIFEST_DIR = os.path.join(HERE, "manifests")
CHARSET = string.printable[:-6]

# network sizes: "full" = configs/test/textdesign_sd_2.yaml of the reference; "tiny" = fast fixture
ARCH = {
    "full": {
        "unet": dict(in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                     channel_mult=[1, 2, 4, 4], num_head_channels=64, transformer_depth=1, t_context_dim=2048),
        "label": dict(max_len=12, emb_dim=2048, n_heads=8, n_trans_layers=12),
        "vae": dict(ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4, in_channels=3, out_ch=3),
        "scale_factor": 0.18215,
    },
    "tiny": {
        "unet": dict(in_channels=9, out_channels=4, model_channels=64, attention_resolutions=[2, 1], num_res_blocks=1,
                     channel_mult=[1, 2], num_head_channels=64, transformer_depth=1, t_context_dim=128),
        "label": dict(max_len=12, emb_dim=128, n_heads=8, n_trans_layers=2),
        "vae": dict(ch=32, ch_mult=[1, 2, 2, 2], num_res_blocks=1, z_channels=4, in_channels=3, out_ch=3),
        "scale_factor": 0.18215,
    },
}


PARSEQ_ARCH = dict(img_size=(32, 128), patch_size=(4, 8), embed_dim=384, enc_num_heads=6, enc_mlp_ratio=4, enc_depth=12,
                   dec_num_heads=12, dec_mlp_ratio=4, dec_depth=1, max_label_length=25, num_tokens=97)


def parseq_manifest(arch: Optional[dict] = None) -> Dict[str, List[int]]:
    """state_dict layout (key -> shape) of the PARSeq checkpoint `parseq-bb5792a6.pt` that ParseqPredictor loads
    (configs/test.yaml:34): timm ViT encoder + one two-stream decoder layer + head (src/parseq/strhub/models/parseq)"""
    a = dict(PARSEQ_ARCH, **(arch or {}))
    d, ph, pw = a["embed_dim"], a["patch_size"][0], a["patch_size"][1]
    n = (a["img_size"][0] // ph) * (a["img_size"][1] // pw)
    m: Dict[str, List[int]] = {"encoder.pos_embed": [1, n, d], "encoder.patch_embed.proj.weight": [d, 3, ph, pw],
                               "encoder.patch_embed.proj.bias": [d]}

    def lin(name, o, i):
        m[name + ".weight"], m[name + ".bias"] = [o, i], [o]

    def ln(name):
        m[name + ".weight"], m[name + ".bias"] = [d], [d]

    for i in range(a["enc_depth"]):
        p = f"encoder.blocks.{i}."
        ln(p + "norm1"); lin(p + "attn.qkv", 3 * d, d); lin(p + "attn.proj", d, d); ln(p + "norm2")
        lin(p + "mlp.fc1", a["enc_mlp_ratio"] * d, d); lin(p + "mlp.fc2", d, a["enc_mlp_ratio"] * d)
    ln("encoder.norm")
    for i in range(a["dec_depth"]):
        p = f"decoder.layers.{i}."
        for att in ("self_attn", "cross_attn"):
            m[p + att + ".in_proj_weight"], m[p + att + ".in_proj_bias"] = [3 * d, d], [3 * d]
            lin(p + att + ".out_proj", d, d)
        lin(p + "linear1", a["dec_mlp_ratio"] * d, d); lin(p + "linear2", d, a["dec_mlp_ratio"] * d)
        for nm in ("norm1", "norm2", "norm_q", "norm_c"):
            ln(p + nm)
    ln("decoder.norm")
    lin("head", a["num_tokens"] - 2, d)
    m["text_embed.embedding.weight"] = [a["num_tokens"], d]
    m["pos_queries"] = [1, a["max_label_length"] + 1, d]
    return m


def load_manifest(name: str) -> Dict[str, List[int]]:
    with open(os.path.join(MANIFEST_DIR, f"{name}.json")) as f:
        return json.load(f)


def _sinusoid_pe(max_len: int, d: int) -> torch.Tensor:
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _ddpm_sigmas(num: int) -> torch.Tensor:
    betas = torch.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, num, dtype=torch.float64) ** 2
    abar = np.cumprod(1.0 - betas.numpy(), axis=0)
    return torch.tensor((1 - abar) / abar, dtype=torch.float32) ** 0.5


def _gauss_kernel(seq_len: int, k: int = 3, sigma: float = 1.0) -> torch.Tensor:
    xs = torch.arange(k).repeat(k).view(k, k)
    grid = torch.stack([xs, xs.t()], dim=-1).float()
    mean = (k - 1) / 2.0
    g = (1.0 / (2.0 * math.pi * sigma ** 2)) * torch.exp(-torch.sum((grid - mean) ** 2.0, dim=-1) / (2 * sigma ** 2))
    g = g / g.sum()
    return g.view(1, 1, k, k).tile(seq_len, 1, 1, 1)


def synthetic_state_dict(manifest: Dict[str, List[int]], seed: int = 1234) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in manifest.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        leaf = key.rsplit(".", 1)[-1]
        if key == "denoiser.sigmas":
            t = _ddpm_sigmas(shape[0])
        elif key.endswith("pos_embedding.pe"):
            t = _sinusoid_pe(shape[0], shape[1])
        elif key == "loss_fn.g_kernel":
            t = _gauss_kernel(shape[0], shape[2])
        elif leaf in ("pos_embed", "pos_queries") or key.endswith("text_embed.embedding.weight"):
            t = torch.randn(shape, generator=g) * (0.5 if leaf != "weight" else 0.05)       # PARSeq tables
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            std = 1.0 if "label_embedding" in key else 1.0 / math.sqrt(fan_in)
            t = torch.randn(shape, generator=g) * std
        elif "norm" in key and leaf == "weight" or key.endswith(("in_layers.0.weight", "out_layers.0.weight", "out.0.weight")):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            t = torch.zeros(shape)
        sd[key] = t.to(torch.float32).contiguous()
    return sd


def synthetic_batch(config_id: int, batch: int, height: int = 512, width: int = 512, label_len: Optional[int] = None
                    ) -> dict:
    """Request batch of SURVEY.md §8(d): seeded image in [-1,1], rectangular mask, random printable labels.
    `label_len=None` draws lengths uniformly from 1..12."""
    g = torch.Generator().manual_seed(config_id)
    image = torch.rand((batch, 3, height, width), generator=g) * 2 - 1
    mask = torch.zeros((batch, 1, height, width))
    mask[:, :, height // 4: height // 2, width // 8: 7 * width // 8] = 1.0
    rng = random.Random(config_id)
    labels = []
    for _ in range(batch):
        n = label_len if label_len is not None else rng.randint(1, 12)
        labels.append("".join(rng.choice(CHARSET[:94]) for _ in range(n)))
    seg = torch.zeros((batch, 12))
    for i, lab in enumerate(labels):
        seg[i, : len(lab)] = 1.0
    size = torch.tensor([[height, width]] * batch)
    return {
        "image": image, "mask": mask, "masked": image * (1 - mask), "seg_mask": seg, "label": labels,
        "txt": [f'"{lab}"' for lab in labels], "name": [str(i) for i in range(batch)],
        "original_size_as_tuple": size.clone(), "target_size_as_tuple": size.clone(),
        "crop_coords_top_left": torch.zeros((batch, 2), dtype=torch.long),
    }
