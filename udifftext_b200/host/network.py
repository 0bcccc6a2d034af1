"""UNet wrapper classes with the reference's constructor surface; compute is `udifftext_b200.unet.UNetB200`.

Reference: sgm/modules/diffusionmodules/openaimodel.py:275-624 (UnifiedUNetModel), wrappers.py:8-35 (OpenAIWrapper).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from ..unet import UNetB200


class _Component:
    """A sub-tree of the engine's `state_dict`: keeps reference-format fp32 weights on the host until `.to(device)`
    builds the kernel executor."""

    def __init__(self):
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self._device: Optional[torch.device] = None
        self.training = False

    # -- weights -------------------------------------------------------------------------------------
    def load_weights(self, sd: Dict[str, torch.Tensor]) -> None:
        self._sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        self._invalidate()
        if self._device is not None:
            self._materialise(self._device)

    def weights(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd or {})

    def _invalidate(self) -> None:
        pass

    def _materialise(self, device: torch.device) -> None:
        raise NotImplementedError

    def _require_weights(self) -> Dict[str, torch.Tensor]:
        if not self._sd:
            raise RuntimeError(f"{type(self).__name__}: no weights loaded (call init_from_ckpt / load_state_dict first)")
        return self._sd

    # -- nn.Module-like surface ----------------------------------------------------------------------
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("udifftext_b200 runs on sm_100a CUDA devices only (there is no CPU path)")
        if self._device != device:
            self._device = device
            self._materialise(device)
        return self

    def cuda(self, index: int = 0):
        return self.to(torch.device("cuda", index))

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("udifftext_b200 is inference-only (the training path is out of scope)")
        return self.eval()

    def freeze(self):
        return self

    def parameters(self):
        return iter((self._sd or {}).values())


class UnifiedUNetModel(_Component):
    """Constructor kwargs of openaimodel.py:277-335; only the SD-2-inpainting style configuration that the shipped
    model YAML uses is executable (2-D, head dim 64, depth-1 linear transformers with the textual `t_attn`)."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, use_label=None, use_checkpoint=False,
                 use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=True, transformer_depth=1,
                 t_context_dim=None, v_context_dim=None, num_attention_blocks=None, use_linear_in_transformer=False,
                 adm_in_channels=None, save_attn_type=None, save_attn_layers=(), ctrl_channels=0, **unused):
        super().__init__()
        unsupported = []
        if dims != 2: unsupported.append("dims != 2")
        if use_label is not None: unsupported.append("use_label")
        if num_head_channels != 64: unsupported.append("num_head_channels != 64")
        if use_scale_shift_norm or resblock_updown: unsupported.append("scale-shift / updown ResBlocks")
        if not isinstance(transformer_depth, int) or transformer_depth != 1: unsupported.append("transformer_depth != 1")
        if not use_linear_in_transformer: unsupported.append("conv proj_in/proj_out")
        if v_context_dim is not None: unsupported.append("v_context")
        if ctrl_channels: unsupported.append("ctrl_channels")
        if t_context_dim is None: unsupported.append("t_context_dim=None")
        if unsupported:
            raise NotImplementedError("UnifiedUNetModel on B200 supports the UDiffText inference configuration only; "
                                      "unsupported: " + ", ".join(unsupported))
        self.arch = dict(in_channels=in_channels, out_channels=out_channels, model_channels=model_channels,
                         attention_resolutions=list(attention_resolutions), num_res_blocks=num_res_blocks,
                         channel_mult=list(channel_mult), num_head_channels=num_head_channels,
                         transformer_depth=transformer_depth, t_context_dim=t_context_dim)
        self.in_channels, self.out_channels, self.model_channels = in_channels, out_channels, model_channels
        self.attn_type = list(save_attn_type or [])
        self.attn_layers = list(save_attn_layers or [])
        self.exec: Optional[UNetB200] = None

    def _invalidate(self):
        self.exec = None

    def _materialise(self, device):
        self.exec = UNetB200(self._require_weights(), device, **self.arch)

    @property
    def attn_map_cache(self) -> List[dict]:
        """openaimodel.py:542-550: one {name, heads, size, attn_map} item per `t_attn`, refreshed by every forward"""
        return self._exec().attn_map_cache

    def _exec(self) -> UNetB200:
        if self.exec is None:
            raise RuntimeError("UnifiedUNetModel: call .to(cuda device) after loading weights")
        return self.exec

    def clear_attn_map(self):
        for item in self._exec().attn_map_cache:
            item["attn_map"] = None

    def save_attn_map(self, attn_type="t_attn", save_name="temp", tokens=""):
        """openaimodel.py:559-591: mean over the selected layers and heads of the last forward's exported maps; the
        LAST sample's [tokens, h, w] map (numpy) is returned and drawn as the reference's 3 x 4 heat-map figure to
        `temp/attn_map/attn_map_<save_name>.png` (PIL instead of seaborn, see host/attn_viz.py)."""
        maps, heads = [], 1
        for item in self._exec().attn_map_cache:
            name = item["name"]
            if any(name.startswith(b) for b in self.attn_layers) and name.endswith(attn_type):
                if item["attn_map"] is None:
                    raise RuntimeError("attention maps were not exported by the last UNet forward (export_attn_maps)")
                heads = item["heads"]
                maps.append(item["attn_map"].detach().float().cpu())
        am = torch.stack(maps, 0).mean(0)
        bh, n, l = am.shape
        am = am.reshape(-1, heads, n, l).mean(1)
        side = int(n ** 0.5)
        attn_map_i = am.permute(0, 2, 1).reshape(am.shape[0], l, side, side).numpy()[-1]
        from .attn_viz import save_attn_figure
        save_attn_figure(attn_map_i, tokens, f"temp/attn_map/attn_map_{save_name}.png")
        return attn_map_i

    def forward(self, x, timesteps=None, t_context=None, v_context=None, y=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        return self._exec().forward(x, timesteps, t_context)

    __call__ = forward


class IdentityWrapper:
    def __init__(self, diffusion_model, compile_model: bool = False):
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)

    __call__ = forward

    def to(self, device):
        self.diffusion_model.to(device)
        return self


class OpenAIWrapper(IdentityWrapper):
    """wrappers.py:23-35: concat conditioning goes onto the channel axis, `t_crossattn` becomes `t_context`"""

    def forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs) -> torch.Tensor:
        x = torch.cat((x, c.get("concat", torch.empty(0, device=x.device, dtype=x.dtype))), dim=1)
        return self.diffusion_model(x, timesteps=t, t_context=c.get("t_crossattn", None),
                                    v_context=c.get("v_crossattn", None), y=c.get("vector", None), **kwargs)

    __call__ = forward
