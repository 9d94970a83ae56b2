"""Change-detection entrypoints: ``lemevit_tiny`` / ``lemevit_small`` / ``lemevit_base`` of the reference's
``change_detection/models/lemevit.py:874-963`` — the same backbone as the mmseg/mmdet copies (4 NCHW feature maps, 'S' blocks
leave the meta tokens untouched, no classifier), constructed with ``pretrained=<checkpoint path>`` and initialised inside the
constructor (:787, :822-853).  ``change_detection/models/networks.py:365-368`` does
``from .lemevit import lemevit_small; self.backbone = lemevit_small(pretrained=args.pretrained)``; the drop-in is
``from lemevit_b200.change_detection import lemevit_small``.
"""
from __future__ import annotations

from .backbone import LeMeViTBackbone
from .model import _COMMON, _VARIANTS, _cfg

__all__ = ["lemevit_tiny", "lemevit_small", "lemevit_base"]


def _create(variant: str, pretrained, kwargs):
    if isinstance(pretrained, bool):      # timm-style pretrained=False
        pretrained = None
    kw = dict(_COMMON)
    kw.update(_VARIANTS[variant])
    kw.update(kwargs)
    frozen = kw.pop("frozen_stages", [-1])
    model = LeMeViTBackbone(pretrained=pretrained, frozen_stages=frozen, **kw)
    model.default_cfg = _cfg()
    if pretrained is not None:
        model.init_weights()              # the CD copy loads its checkpoint in the constructor (:787)
    return model


def lemevit_tiny(pretrained=None, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_tiny", pretrained, kwargs)


def lemevit_small(pretrained=None, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_small", pretrained, kwargs)


def lemevit_base(pretrained=None, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_base", pretrained, kwargs)
