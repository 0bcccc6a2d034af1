"""Import the UNMODIFIED reference (`/root/reference`) in the build container — TEST INFRASTRUCTURE ONLY.

The reference imports nine third-party packages that are not installed here.  They are replaced
by in-memory stand-ins *before* `import sgm`; none of them carries arithmetic on the hot path except
`xformers.ops.memory_efficient_attention`, which is defined as softmax(q k^T / sqrt(d)) v and mapped to
`F.scaled_dot_product_attention` (identical math on the 3-D `[B*h, N, d]` inputs the reference passes).
Used by `oracle/make_golden.py` and the container-only tests that pin `oracle/restated.py`; never by the product,
and never on the GPU box (where `/root/reference` does not exist).
"""
from __future__ import annotations

import copy
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("UDT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sgm"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behave like a package so `import a.b` works
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    """dict with attribute access: enough of omegaconf's DictConfig for the reference's constructors"""

    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(obj):
    if isinstance(obj, dict):
        return _AttrDict({k: to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return _ListConfig(to_cfg(v) for v in obj)
    return obj


class _ListConfig(list):
    pass


def install_shims() -> None:
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_udt_shim", False):
        return
    import transformers  # noqa: F401  (must be imported before the stand-ins shadow its optional deps)

    class LightningModule(nn.Module):
        _device = torch.device("cpu")      # pytorch_lightning's device mixin (strhub/models/parseq/system.py reads it)

        @property
        def device(self):
            return self._device

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

        def save_hyperparameters(self, *a, **k):
            pass

    _mod("pytorch_lightning", LightningModule=LightningModule, seed_everything=lambda s: torch.manual_seed(s),
         _udt_shim=True)

    class OmegaConf(dict):
        @staticmethod
        def load(path):
            import yaml
            with open(path) as f:
                return to_cfg(yaml.safe_load(f))

        @staticmethod
        def create(obj):
            return to_cfg(obj)

    _mod("omegaconf", OmegaConf=OmegaConf, ListConfig=_ListConfig, DictConfig=_AttrDict)
    ops = _mod("xformers.ops",
               memory_efficient_attention=lambda q, k, v, attn_bias=None, op=None: F.scaled_dot_product_attention(q, k, v))
    _mod("xformers", ops=ops, __version__="0.0.22.post7")
    for name in ("kornia", "open_clip", "imageio", "seaborn", "matplotlib"):
        _mod(name)
    sys.modules["matplotlib"].pyplot = _mod("matplotlib.pyplot")
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.vision_transformer", VisionTransformer=nn.Module)


def import_reference_parseq():
    """the reference's bundled PARSeq (src/parseq/strhub) with the ABSENT third-party `timm` replaced by the restated ViT
    of oracle/parseq_restated.py (timm's published algorithm); everything else — PARSeq.forward / decode, Decoder,
    DecoderLayer, TokenEmbedding, Tokenizer — is the unmodified reference code.  Returns the PARSeq class."""
    install_shims()
    from oracle import parseq_restated as PR
    sys.modules["pytorch_lightning"].utilities = _mod("pytorch_lightning.utilities")
    _mod("pytorch_lightning.utilities.types", STEP_OUTPUT=object)

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        for child_name, child in module.named_children():
            named_apply(fn, child, ".".join((name, child_name)) if name else child_name, depth_first, True)
        if include_root:
            fn(module=module, name=name)
        return module

    _mod("timm.models.vision_transformer", VisionTransformer=PR.RestatedViT, PatchEmbed=PR.PatchEmbed)
    _mod("timm.models.helpers", named_apply=named_apply)
    _mod("timm.optim", create_optimizer_v2=lambda *a, **k: None)
    root = os.path.join(REFERENCE_ROOT, "src", "parseq")
    if root not in sys.path:
        sys.path.insert(0, root)
    from strhub.models.parseq.system import PARSeq
    return PARSeq


def import_reference():
    """returns the reference's `sgm` package (unmodified sources, imported from REFERENCE_ROOT)"""
    if not reference_available():
        raise RuntimeError(f"reference sources not found at {REFERENCE_ROOT}")
    install_shims()
    # our own repo also ships a package called `sgm` (the drop-in mirror); make sure the reference's wins here
    for k in [k for k in sys.modules if k == "sgm" or k.startswith("sgm.")]:
        if not getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[k]
    if REFERENCE_ROOT not in sys.path or sys.path[0] != REFERENCE_ROOT:
        sys.path.insert(0, REFERENCE_ROOT)
    import sgm  # noqa: E402
    assert sgm.__file__.startswith(REFERENCE_ROOT), sgm.__file__
    return sgm


def load_model_config(small: dict | None = None):
    """configs/test/textdesign_sd_2.yaml with every ckpt_path and the OCR predictor removed (no checkpoints
    exist here); `small` overrides UNet / VAE / LabelEncoder sizes for fast fixtures."""
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "configs/test/textdesign_sd_2.yaml")) as f:
        cfg = yaml.safe_load(f)

    def strip(node):
        if isinstance(node, dict):
            node.pop("ckpt_path", None)
            node.pop("predictor_config", None)
            for v in node.values():
                strip(v)
        elif isinstance(node, list):
            for v in node:
                strip(v)

    strip(cfg)
    params = cfg["model"]["params"]
    if small:
        params["network_config"]["params"].update(small.get("unet", {}))
        params["conditioner_config"]["params"]["emb_models"][0]["params"].update(small.get("label", {}))
        for dd in (params["conditioner_config"]["params"]["emb_models"][2]["params"]["config"]["params"]["ddconfig"],
                   params["first_stage_config"]["params"]["ddconfig"]):
            dd.update(small.get("vae", {}))
    return to_cfg(cfg)


def build_reference_engine(seed: int = 1234, small: dict | None = None):
    """Instantiate the reference DiffusionEngine on CPU with seeded random weights and re-randomise the
    zero-initialised layers (otherwise eps == 0 and every parity check passes vacuously)."""
    sgm = import_reference()
    from sgm.util import instantiate_from_config
    torch.manual_seed(seed)
    cfg = load_model_config(small)
    model = instantiate_from_config(cfg.model)
    model.eval()
    rerandomise_zero_init(model, seed + 1)
    return model


def rerandomise_zero_init(model, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and p.abs().max() == 0:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
