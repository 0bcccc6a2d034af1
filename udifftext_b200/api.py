"""Public entry points: the reference's `util.py` factories and `test.py`/`demo.py`'s `predict`, same names, same
arguments, same results — on the B200 kernels.

Reference: util.py:7-22 init_model, :24-47 init_sampling, :49-60 deep_copy, :62-77 prepare_batch; test.py:19-40
predict (identical copy in demo.py:15-36).
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops, synth
from .host import rng
from .host.config import AttrDict, instantiate_from_config, load_yaml, to_attr
from .host.sampler import EulerEDMSampler

_SGM = "sgm.modules.diffusionmodules."


def model_config(arch: str = "full") -> AttrDict:
    """The model section of configs/test/textdesign_sd_2.yaml expressed in code (checkpoint paths and the OCR
    predictor — evaluation-only — left out); `arch="tiny"` is the small fixture network of the tests."""
    a = synth.ARCH[arch]
    disc = {"target": _SGM + "discretizer.LegacyDDPMDiscretization"}
    vae = {"target": "sgm.models.autoencoder.AutoencoderKLInferenceWrapper",
           "params": {"embed_dim": 4, "monitor": "val/rec_loss", "lossconfig": {"target": "torch.nn.Identity"},
                      "ddconfig": dict(attn_type="vanilla-xformers", double_z=True, resolution=256, attn_resolutions=[],
                                       dropout=0.0, **a["vae"])}}
    unet = dict(a["unet"], ctrl_channels=0, save_attn_type=["t_attn"],
                save_attn_layers=["output_blocks.6.1" if arch == "full" else "output_blocks.3.1"],
                use_linear_in_transformer=True)
    cfg = {"target": "sgm.models.diffusion.DiffusionEngine", "params": {
        "opt_keys": ["t_attn"], "input_key": "image", "scale_factor": a["scale_factor"], "disable_first_stage_autocast": True,
        "denoiser_config": {"target": _SGM + "denoiser.DiscreteDenoiser", "params": {
            "num_idx": 1000, "weighting_config": {"target": _SGM + "denoiser_weighting.EpsWeighting"},
            "scaling_config": {"target": _SGM + "denoiser_scaling.EpsScaling"}, "discretization_config": disc}},
        "network_config": {"target": _SGM + "openaimodel.UnifiedUNetModel", "params": unet},
        "conditioner_config": {"target": "sgm.modules.GeneralConditioner", "params": {"emb_models": [
            {"is_trainable": False, "emb_key": "t_crossattn", "ucg_rate": 0.1, "input_key": "label",
             "target": "sgm.modules.encoders.modules.LabelEncoder", "params": dict(a["label"])},
            {"is_trainable": False, "input_key": "mask", "target": "sgm.modules.encoders.modules.SpatialRescaler",
             "params": {"in_channels": 1, "multiplier": 0.125}},
            {"is_trainable": False, "input_key": "masked", "target": "sgm.modules.encoders.modules.LatentEncoder",
             "params": {"scale_factor": a["scale_factor"], "config": vae}}]}},
        "first_stage_config": vae,
        "loss_fn_config": {"target": _SGM + "loss.FullLoss", "params": {
            "seq_len": 12, "kernel_size": 3, "gaussian_sigma": 1.0, "min_attn_size": 16 if arch == "full" else 4,
            "lambda_local_loss": 0.01, "lambda_ocr_loss": 0.001, "ocr_enabled": False,
            "sigma_sampler_config": {"target": _SGM + "sigma_sampling.DiscreteSampling",
                                     "params": {"num_idx": 1000, "discretization_config": disc}}}}}}
    return to_attr(cfg)


def runtime_config(**overrides) -> AttrDict:
    """configs/test.yaml's runtime fields that the hot path reads"""
    cfg = dict(type="test", channel=4, factor=8, scale=[5.0, 0.0], noise_iters=0, force_uc_zero_embeddings=["label"],
               aae_enabled=False, detailed=False, steps=50, init_step=0, batch_size=1, gpu=0)
    cfg.update(overrides)
    return to_attr(cfg)


def build_engine(arch: str = "full", device=None, seed: int = 1234, state_dict: Optional[Dict[str, torch.Tensor]] = None):
    """DiffusionEngine with seeded synthetic weights (no checkpoints exist in this environment) or `state_dict`"""
    model = instantiate_from_config(model_config(arch))
    sd = state_dict if state_dict is not None else synth.synthetic_state_dict(synth.load_manifest(arch), seed)
    model.load_state_dict(sd, strict=True)
    if device is not None:
        model.to(device)
    return model.eval()


# ------------------------------------------------------------------------------------------------- util.py mirror
def init_model(cfgs):
    """util.py:7-22"""
    model_cfg = load_yaml(cfgs.model_cfg_path)
    model = instantiate_from_config(model_cfg.model)
    model.init_from_ckpt(cfgs.load_ckpt_path)
    if cfgs.type == "train":
        model.train()
    else:
        model.to(torch.device("cuda", index=cfgs.gpu))
        model.eval()
        model.freeze()
    return model


def init_sampling(cfgs):
    """util.py:24-47"""
    return EulerEDMSampler(
        num_steps=cfgs.steps,
        discretization_config={"target": _SGM + "discretizer.LegacyDDPMDiscretization"},
        guider_config={"target": _SGM + "guiders.VanillaCFG", "params": {"scale": cfgs.scale[0]}},
        s_churn=0.0, s_tmin=0.0, s_tmax=999.0, s_noise=1.0, verbose=True, device=torch.device("cuda", index=cfgs.gpu))


def deep_copy(batch: Dict) -> Dict:
    """util.py:49-60"""
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.clone()
        elif isinstance(v, (tuple, list)):
            out[k] = list(v) if isinstance(v, list) else v
        else:
            out[k] = v
    return out


def prepare_batch(cfgs, batch: Dict) -> Tuple[Dict, Dict]:
    """util.py:62-77: H2D of every tensor; the unconditional batch has empty txt / label.  The uc batch shares the
    device tensors (they are never written) instead of cloning them."""
    dev = torch.device("cuda", index=cfgs.gpu)
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            batch[k] = v.to(dev, non_blocking=True)
    batch_uc = dict(batch)
    batch_uc["txt"] = batch["ntxt"] if "ntxt" in batch else ["" for _ in batch["txt"]]
    if "label" in batch:
        batch_uc["label"] = ["" for _ in batch["label"]]
    return batch, batch_uc


# ------------------------------------------------------------------------------------------------- demo.py:52-101
def request_batch_u8(cfgs, image_u8, mask_u8, text: str, num_samples: int, name: str = "0") -> Dict:
    """demo.py:52-98 with the pixel work on the device: `image_u8` uint8 [H, W, 3] and `mask_u8` uint8 [H, W, 3] (the
    user's brush mask, 0 = keep) already resized to cfgs.H x cfgs.W (cv2.resize stays with the caller) are copied as
    uint8 (1 MB instead of 7.3 MB of fp32 per sample) and expanded by `udt_request_pack_u8`; the returned batch dict
    has the reference's schema with device tensors and goes straight into `predict`."""
    dev = torch.device("cuda", index=cfgs.gpu)
    img = torch.as_tensor(image_u8)
    msk = torch.as_tensor(mask_u8)
    if msk.dim() == 2:
        msk = msk[..., None]
    if img.dtype != torch.uint8 or msk.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or msk.shape[:2] != img.shape[:2]:
        raise ValueError("request_batch_u8: image uint8 [H, W, 3] and mask uint8 [H, W(, C)] of the same size expected")
    seq_len = int(getattr(cfgs, "seq_len", 12))          # configs/demo.yaml:10
    if len(text) > seq_len:
        raise ValueError(f"text longer than seq_len={seq_len}")
    hh, ww = int(img.shape[0]), int(img.shape[1])
    image, mask, masked = ops.request_pack_u8(img.contiguous()[None].to(dev, non_blocking=True),
                                              msk.contiguous()[None].to(dev, non_blocking=True), num_samples)
    seg = torch.cat((torch.ones(len(text)), torch.zeros(seq_len - len(text))))
    tile = lambda t: torch.tile(t[None], (num_samples, 1))
    return {
        "image": image, "mask": mask, "masked": masked, "seg_mask": tile(seg),
        "label": [text] * num_samples, "txt": [f'"{text}"'] * num_samples,
        "original_size_as_tuple": tile(torch.tensor((hh, ww))), "crop_coords_top_left": tile(torch.tensor((0, 0))),
        "target_size_as_tuple": tile(torch.tensor((hh, ww))), "name": [name] * num_samples,
    }


def images_to_u8(samples: torch.Tensor) -> torch.Tensor:
    """demo.py:100-101 / test.py:94: decoded images [B, 3, H, W] in [0, 1] -> uint8 [B, H, W, 3] on the device
    (`(x * 255).astype(uint8)`); the caller's `.cpu()` then moves a quarter of the bytes"""
    return ops.images_to_u8(samples.contiguous())


# ------------------------------------------------------------------------------------------------- test.py:19-40
def predict(cfgs, model, sampler, batch: Dict, shard: Optional[Tuple[int, int, int]] = None):
    """returns (samples [B,3,H,W] fp32 in [0,1], samples_z [B,4,H/8,W/8]) on the device, like the reference.
    `shard=(global_batch, lo, hi)` marks `batch` as rows [lo, hi) of a larger request (multi-GPU): noise is then
    drawn for the whole request and sliced, so results do not depend on the number of ranks."""
    if cfgs.aae_enabled:
        raise NotImplementedError("aae_enabled needs autograd through the UNet; udifftext_b200 is inference-only")
    ctx = rng.batch_shard(*shard) if shard is not None else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        batch, batch_uc = prepare_batch(cfgs, batch)
        c, uc = model.conditioner.get_unconditional_conditioning(
            batch, batch_uc=batch_uc, force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
        x = sampler.get_init_noise(cfgs, model, cond=c, batch=batch, uc=uc)
        samples_z = sampler(model, x, cond=c, batch=batch, uc=uc, init_step=0, aae_enabled=cfgs.aae_enabled,
                            detailed=cfgs.detailed)
        samples = model.decode_first_stage_clamped(samples_z)
    return samples, samples_z


# ------------------------------------------------------------------------------------------------- multi-GPU
def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """rows [lo, hi) of a request batch owned by `rank`: contiguous, sizes differ by at most one"""
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict, lo: int, hi: int) -> Dict:
    """rows [lo, hi) of every per-sample entry of a request batch (tensors and lists)"""
    n = len(batch["label"]) if "label" in batch else None
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and (n is None or v.shape[0] == n):
            out[k] = v[lo:hi]
        elif isinstance(v, list) and (n is None or len(v) == n):
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def all_gather_images(samples: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """The path's only collective: every rank contributes its decoded images [hi-lo, 3, H, W] and receives the whole
    request [global_batch, 3, H, W] in request order (ncclAllGather over NVLink; gloo in the CPU tests).  Equal
    shards use one all_gather_into_tensor; ragged shards are padded to the largest shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    big = max(hi - lo for lo, hi in sizes)
    mine = samples.contiguous()
    if mine.shape[0] < big:
        mine = torch.cat([mine, mine.new_zeros((big - mine.shape[0],) + tuple(mine.shape[1:]))])
    out = mine.new_empty((world * big,) + tuple(mine.shape[1:]))
    dist.all_gather_into_tensor(out, mine, group=group)
    if all(hi - lo == big for lo, hi in sizes):
        return out
    return torch.cat([out[r * big: r * big + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
