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

    def __init__(self, *args, pretrained=None, init_cfg=None, frozen_stages=-1, norm_eval=None, out_dtype=None, **kwargs):
        kwargs.setdefault("num_classes", 1000)
        super().__init__(*args, **kwargs)
        assert not (init_cfg and pretrained), 'init_cfg and pretrained cannot be specified at the same time'
        if isinstance(pretrained, str):
            init_cfg = dict(type='Pretrained', checkpoint=pretrained)
        elif pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        self.init_cfg = init_cfg
        # mmdet / change-detection configs pass a list of stage indices ([-1] = none, object_detection/.../lemevit.py:667,827-831);
        # an int n is read the OpenMMLab way (stages 0..n frozen, -1 = none)
        if isinstance(frozen_stages, int):
            frozen_stages = list(range(frozen_stages + 1)) if frozen_stages >= 0 else [-1]
        self.frozen_stages = list(frozen_stages)
        # mmdet's train() keeps every BatchNorm / LayerNorm in eval mode (freeze_bn = True, :833-842); mmseg's does not (:874-882).
        # None: decided by which registry built the module (register_backbones); False for direct construction.
        self.norm_eval = norm_eval
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

    def _freeze_stages(self):
        # object_detection/mmdet/models/backbones/lemevit.py:827-831
        for i in self.frozen_stages:
            if i >= 0:
                for param in self.stages[i].parameters():
                    param.requires_grad = False

    def train(self, mode=True):
        """mmdet semantics (:833-842): frozen stages lose requires_grad and, with ``norm_eval``, every BatchNorm2d / LayerNorm
        stays in eval mode, so a ``train()``-ed detector backbone computes exactly the eval forward.  Like the reference
        overrides (mmseg :874-882) this returns None — callers must not chain on .eval()/.train()."""
        self._freeze_stages()
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
                    m.eval()

    def _check_inference(self):
        # a backbone whose norms are frozen (mmdet) and that has no stochastic depth runs the same arithmetic in train mode
        if self.training and self.norm_eval and not self.drop_path_rate and not self.drop_rate:
            if not any(m.training for m in self.modules() if isinstance(m, nn.BatchNorm2d)):
                if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                    raise RuntimeError("lemevit_b200 is inference-only: the backbone has trainable parameters but the native forward "
                                       "builds no autograd graph. Freeze it (frozen_stages=[0,1,2,3,4] + torch.no_grad()) or train with "
                                       "the reference implementation.")
                return
        super()._check_inference()

    def forward(self, x) -> List[torch.Tensor]:
        require_cuda(x)
        self._check_inference()
        eng = self.native_engine(x.device)
        dt = self.out_dtype or (torch.bfloat16 if self.meta_tokens.dtype == torch.bfloat16 else torch.float32)
        return eng.forward_features(x, out_dtype=dt)

    def forward_features(self, x, c=None):
        return self.forward(x)


def _mmdet_init(self, *args, **kwargs):
    kwargs.setdefault("norm_eval", True)
    LeMeViTBackbone.__init__(self, *args, **kwargs)


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
            cls = LeMeViTBackbone
            if pkg == "mmdet":      # the mmdet copy freezes its norms in train() (object_detection/.../lemevit.py:833-842)
                cls = type("LeMeViT", (LeMeViTBackbone,), {"__init__": _mmdet_init, "__module__": __name__})
            reg.register_module(name="LeMeViT", force=True, module=cls)
            done.append(pkg)
        except Exception as e:  # pragma: no cover
            if verbose:
                print(f"lemevit_b200: could not register with {pkg}: {e}")
    return done
