"""lemevit_b200 — Blackwell-native (sm_100a) LeMeViT backbone forward pass.

Drop-in surface of the reference (ViTAE-Transformer/LeMeViT):
  * ``LeMeViT`` nn.Module + ``lemevit_tiny`` / ``lemevit_small`` / ``lemevit_base`` entrypoints
    (reference models/lemevit.py:663-932); registered with timm when timm is importable, so that
    ``from lemevit_b200 import *`` followed by ``timm.create_model('lemevit_base')`` works exactly like
    the reference's ``from models import *`` (benchmark.py:70, main.py:39, validate.py:32);
  * ``LeMeViTBackbone`` registered as ``LeMeViT`` in mmseg / mmdet ``BACKBONES`` when those are importable.
The forward runs in liblemevit_b200.so (hand-written CUDA, C ABI in include/lemevit_b200.h).
"""
from .model import LeMeViT, lemevit_base, lemevit_small, lemevit_tiny
from .backbone import LeMeViTBackbone, register_backbones

__all__ = ["LeMeViT", "LeMeViTBackbone", "lemevit_tiny", "lemevit_small", "lemevit_base", "register_backbones"]

try:  # timm registry (optional dependency, exactly as the reference uses it: models/lemevit.py:20,845)
    from timm.models import register_model as _register_model

    for _fn in (lemevit_tiny, lemevit_small, lemevit_base):
        try:
            _register_model(_fn)
        except Exception:  # already registered / stub registry
            pass
except Exception:  # timm not installed
    pass

register_backbones()
