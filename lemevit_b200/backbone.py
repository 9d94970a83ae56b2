"""mmseg / mmdet backbone variant of the native LeMeViT and its (conditional) registration.

Mirrors ``@BACKBONES.register_module() class LeMeViT`` of the reference
(semantic_segmentation/mmseg/models/backbones/lemevit.py:660-882 and
object_detection/mmdet/models/backbones/lemevit.py:660-876): same constructor kwargs
(+ ``pretrained`` / ``init_cfg`` / ``frozen_stages``), same state_dict keys minus ``head.*``,
``forward(img) -> [4 x NCHW map]``; 'S' blocks leave the meta tokens untouched (:630-636).
mmcv / mmseg / mmdet are optional: when importable the class registers itself under the name
``LeMeViT``; tests exercise the same code against a stand-in registry.
"""
from __future__ import annotations

import logging
from collections import OrderedDict
from typing import List

import torch
import torch.nn as nn

from .engine import require_cuda
from .model import LeMeViT


class LeMeViTBackbone(LeMeViT):
    backbone_mode = True

    def __init__(self, *args, pretrained=None, init_cfg=None, frozen_stages=-1, out_dtype=None, **kwargs):
        kwargs.setdefault("num_classes", 1000)
        super().__init__(*args, **kwargs)
        assert not (init_cfg and pretrained), 'init_cfg and pretrained cannot be specified at the same time'
        if isinstance(pretrained, str):
            init_cfg = dict(type='Pretrained', checkpoint=pretrained)
        elif pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        self.init_cfg = init_cfg
        self.frozen_stages = frozen_stages
        self.out_dtype = out_dtype
        # the backbone copies have no classifier (reference :786 commented out): drop it so the key set matches
        self.head = nn.Identity()
        self._drop_engine()

    def _build_head(self, num_classes):
        self.head = nn.Identity()

    @staticmethod
    def _clean_state_dict(ckpt) -> "OrderedDict[str, torch.Tensor]":
        """Checkpoint key handling of the reference init_weights (:851-872)."""
        for key in ("state_dict", "state_dict_ema", "model"):
            if key in ckpt:
                ckpt = ckpt[key]
                break
        sd = OrderedDict((k[9:] if k.startswith("backbone.") else k, v) for k, v in ckpt.items())
        if sd and next(iter(sd)).startswith("module."):
            sd = OrderedDict((k[7:], v) for k, v in sd.items())
        return sd

    def init_weights(self, pretrained=None):
        """mmseg: ``init_weights()`` driven by init_cfg (:829-872); mmdet: ``init_weights(pretrained)`` (:845-876)."""
        log = logging.getLogger("lemevit_b200")
        ckpt_path = pretrained if isinstance(pretrained, str) else (self.init_cfg or {}).get("checkpoint")
        if ckpt_path is None:
            log.warning("No pre-trained weights for %s, training start from scratch", self.__class__.__name__)
            self.apply(self._init_weights)
        else:
            ckpt = torch.load(ckpt_path, map_location="cpu")
            log.info(self.load_state_dict(self._clean_state_dict(ckpt), strict=False))
        self._drop_engine()

    def train(self, mode=True):
        # the reference overrides return None (:874-882) — callers must not chain on .eval()/.train()
        super().train(mode)

    def forward(self, x) -> List[torch.Tensor]:
        require_cuda(x)
        eng = self.native_engine(x.device)
        dt = self.out_dtype or (torch.bfloat16 if self.meta_tokens.dtype == torch.bfloat16 else torch.float32)
        return eng.forward_features(x, out_dtype=dt)

    def forward_features(self, x, c=None):
        return self.forward(x)


def register_backbones(verbose: bool = False) -> List[str]:
    """Register ``LeMeViT`` with every importable OpenMMLab BACKBONES registry.  Returns where it landed."""
    done = []
    for pkg in ("mmseg", "mmdet"):
        try:
            builder = __import__(pkg + ".models.builder", fromlist=["BACKBONES"])
            reg = builder.BACKBONES
        except Exception:
            continue
        try:
            reg.register_module(name="LeMeViT", force=True, module=LeMeViTBackbone)
            done.append(pkg)
        except Exception as e:  # pragma: no cover
            if verbose:
                print(f"lemevit_b200: could not register with {pkg}: {e}")
    return done
