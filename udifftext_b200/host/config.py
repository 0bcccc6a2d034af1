"""Config-driven class injection — the reference's plugin mechanism (sgm/util.py:168-185: a `{target, params}`
mapping names a class by dotted path and its constructor kwargs) — plus a minimal stand-in for the parts of
omegaconf the runtime YAMLs need when omegaconf is not installed."""
from __future__ import annotations

import importlib
from typing import Any, Mapping


def get_obj_from_str(path: str, reload: bool = False) -> Any:
    """'pkg.mod.Name' -> the object (sgm/util.py:178-185)"""
    if path.startswith("sgm."):
        from .registry import TARGETS  # late import: the registry imports the classes that import this module
        if path in TARGETS:
            return TARGETS[path]
        raise NotImplementedError(f"'{path}' is outside the inference hot path that udifftext_b200 implements")
    mod_name, _, attr = path.rpartition(".")
    mod = importlib.import_module(mod_name)
    if reload:
        mod = importlib.reload(mod)
    return getattr(mod, attr)


def instantiate_from_config(config: Mapping) -> Any:
    """sgm/util.py:168-175: `__is_first_stage__` / `__is_unconditional__` sentinels give None, otherwise the
    class named by `target` is called with `params`."""
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**dict(config.get("params", None) or {}))


class AttrDict(dict):
    """dict with attribute access (enough of omegaconf.DictConfig for cfgs.steps / cfgs.scale[0] / cfg.model)"""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value


def to_attr(obj: Any) -> Any:
    if isinstance(obj, Mapping):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def load_yaml(path: str) -> AttrDict:
    """OmegaConf.load replacement: omegaconf when importable, else PyYAML + attribute dicts."""
    try:
        from omegaconf import OmegaConf  # type: ignore
        return OmegaConf.load(path)
    except ImportError:
        import yaml
        with open(path) as f:
            return to_attr(yaml.safe_load(f))


def default(value, fallback):
    return fallback if value is None else value
