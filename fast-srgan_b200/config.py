"""Hydra-less loader for the reference's config schema (configs/config.yaml; train.py:46, inference.py:26).

`load(path, overrides=["training.batch_size=64", ...])` returns an attribute-access namespace like the
OmegaConf DictConfig the reference uses.  PyYAML parses `1e-4` as a *string* (OmegaConf parses a float): numeric
strings are coerced."""
from __future__ import annotations

import os
import types
from typing import Iterable, Optional

import yaml

DEFAULT_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "config.yaml")


def _coerce(v):
    if isinstance(v, str):
        for cast in (int, float):
            try:
                return cast(v)
            except ValueError:
                pass
        if v.lower() in ("true", "false"):
            return v.lower() == "true"
    return v


def _ns(d):
    if isinstance(d, dict):
        return types.SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    return _coerce(d)


def load(path: Optional[str] = None, overrides: Iterable[str] = ()):
    with open(path or DEFAULT_PATH) as f:
        cfg = yaml.safe_load(f)
    for ov in overrides:
        key, _, val = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _coerce(val)
    return _ns(cfg)
