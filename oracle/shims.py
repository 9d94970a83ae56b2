"""Import shims that let the UNMODIFIED reference model files load in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lemevit_b200/`` may import this file.

The reference (``/root/reference/models/lemevit.py:19-22``) imports five symbols from
``timm`` and ``fairscale``, neither of which is installed here; the mmseg copy
(``/root/reference/semantic_segmentation/mmseg/models/backbones/lemevit.py:24-34``)
additionally needs ``mmcv.runner`` / ``mmseg.utils`` and a ``..builder.BACKBONES``
registry.  The stubs below provide exactly those names and nothing that does arithmetic:
every number the reference produces still comes from its own code running on torch.

``load_reference_cls()`` / ``load_reference_mmseg()`` return the reference modules, or
``None`` when ``/root/reference`` is not present (e.g. on the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

from . import ref_copy

# /root/reference in the build container; on the GPU box the verbatim, git-ignored copy under baseline/_ref (oracle/ref_copy.py)
REFERENCE_ROOT = ref_copy.root() or ref_copy.MOUNT


class _Registry:
    """Stand-in for mmcv's Registry: ``@BACKBONES.register_module()``."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls

        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)


class _DropPath(nn.Module):
    """Stochastic depth; identity in eval mode / p == 0 (all the oracle ever uses)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _cfg(url="", **kwargs):
    d = {
        "url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None,
        "crop_pct": 0.9, "interpolation": "bicubic", "fixed_input_size": True,
        "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5),
        "first_conv": "patch_embed.proj", "classifier": "head",
    }
    d.update(kwargs)
    return d


_MODEL_REGISTRY = {}


def _register_model(fn):
    _MODEL_REGISTRY[fn.__name__] = fn
    return fn


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package so sub-imports resolve
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs():
    """Inject the stub modules (idempotent; never overrides a real installed package)."""
    try:  # pragma: no cover - timm is not installed in this image
        import timm  # noqa: F401
        have_timm = not getattr(sys.modules["timm"], "_lemevit_stub", False)
    except Exception:
        have_timm = False
    if not have_timm:
        _mod("timm", _lemevit_stub=True)
        _mod("timm.models", register_model=_register_model)
        _mod("timm.models.layers", DropPath=_DropPath, to_2tuple=_to_2tuple,
             trunc_normal_=nn.init.trunc_normal_, LayerNorm2d=nn.GroupNorm)
        _mod("timm.models.vision_transformer", _cfg=_cfg)
    if "fairscale" not in sys.modules:
        _mod("fairscale")
        _mod("fairscale.nn")
        _mod("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
    import logging

    def _logger(*a, **k):
        return logging.getLogger("lemevit_ref")

    def _load_checkpoint(path, map_location=None, **k):
        return torch.load(path, map_location=map_location)

    if "mmcv" not in sys.modules:
        _mod("mmcv")
        _mod("mmcv.runner", BaseModule=nn.Module, _load_checkpoint=_load_checkpoint)
    for pkg in ("mmseg", "mmdet"):
        if pkg not in sys.modules:
            _mod(pkg)
            _mod(pkg + ".utils", get_root_logger=_logger)
            _mod(pkg + ".models")
            _mod(pkg + ".models.builder", BACKBONES=_Registry("backbone"))
            _mod(pkg + ".models.backbones")


def _load(path, modname):
    if not os.path.isfile(path):
        return None
    install_stubs()
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_cls():
    """The classification model file, untouched (models/lemevit.py)."""
    return _load(os.path.join(REFERENCE_ROOT, "models", "lemevit.py"), "_lemevit_reference_cls")


def load_reference_mmseg():
    """The mmseg backbone copy, untouched."""
    return _load(
        os.path.join(REFERENCE_ROOT, "semantic_segmentation", "mmseg", "models", "backbones", "lemevit.py"),
        "mmseg.models.backbones.lemevit",
    )


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "lemevit.py"))
