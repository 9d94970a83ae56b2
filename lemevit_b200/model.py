"""Drop-in ``LeMeViT`` module: the reference's constructor, attributes, methods and state_dict
schema (reference models/lemevit.py:663-836), with ``forward`` executed by the native sm_100a
library instead of PyTorch ops.

The sub-modules below are PARAMETER CONTAINERS: they exist so that ``state_dict()`` /
``load_state_dict()`` / ``.to()`` / ``parameters()`` / ``deepcopy`` / DDP wrapping behave exactly
as for the reference (569 tensors for Base, same names and shapes — SURVEY.md §8b).  Their own
``forward`` is never called on the hot path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .engine import Engine, require_cuda

_UNSUPPORTED = ("lemevit_b200 implements the live configuration of the published LeMeViT variants only "
                "(pre-norm, no layer scale, cpe_ks=3, no mlp dwconv, attn types C/D/S, qk_dims == embed_dim): {}")


def _cfg(url: str = "", **kwargs):
    """timm-style default_cfg (what ``timm.models.vision_transformer._cfg()`` returns for the
    reference, models/lemevit.py:866); read by ``resolve_data_config`` in benchmark.py:431."""
    d = {
        "url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None,
        "crop_pct": 0.9, "interpolation": "bicubic", "fixed_input_size": True,
        "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5),
        "first_conv": "patch_embed.proj", "classifier": "head",
    }
    d.update(kwargs)
    return d


class _Attn(nn.Module):
    """Holds the projections of CrossAttention / DualCrossAttention / StandardAttention under the
    reference's attribute names (models/lemevit.py:175-178, 241-246, 444-448)."""

    def __init__(self, dim: int, num_heads: int, kind: str):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} not divisible by num_heads {num_heads}"
        self.num_heads = num_heads
        if kind == "C":
            self.q = nn.Linear(dim, dim)
            self.kv = nn.Linear(dim, 2 * dim)
            self.proj = nn.Linear(dim, dim)
        elif kind == "D":
            self.qkv1 = nn.Linear(dim, 3 * dim)
            self.qkv2 = nn.Linear(dim, 3 * dim)
            self.proj_x = nn.Linear(dim, dim)
            self.proj_c = nn.Linear(dim, dim)
        else:
            self.qkv = nn.Linear(dim, 3 * dim)
            self.proj = nn.Linear(dim, dim)


class _Block(nn.Module):
    """Parameter layout of LeMeBlock (models/lemevit.py:500-539)."""

    def __init__(self, dim: int, kind: str, num_heads: int, mlp_ratio: float):
        super().__init__()
        self.attn_type = kind
        self.pos_embed = nn.Conv2d(dim, dim, kernel_size=3, padding=1, groups=dim)
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, num_heads, kind)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        hidden = int(mlp_ratio * dim)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.Identity(), nn.GELU(), nn.Linear(hidden, dim))


def _meta_mlp(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(cin, cin * 4), nn.LayerNorm(cin * 4), nn.GELU(), nn.Linear(cin * 4, cout), nn.LayerNorm(cout))


class LeMeViT(nn.Module):
    """Same signature as the reference constructor (models/lemevit.py:664-687)."""

    backbone_mode = False   # the mmseg/mmdet subclass flips this

    def __init__(self,
                 depth=[2, 3, 4, 8, 3],
                 in_chans=3,
                 num_classes=1000,
                 embed_dim=[64, 64, 128, 320, 512],
                 head_dim=64,
                 mlp_ratios=[4, 4, 4, 4, 4],
                 qkv_bias=True,
                 qk_scale=None,
                 drop_rate=0.,
                 attn_drop=0.,
                 drop_path_rate=0.,
                 attn_type=["C", "D", "D", "S", "S"],
                 queries_len=128,
                 qk_dims=None,
                 cpe_ks=3,
                 pre_norm=True,
                 mlp_dwconv=False,
                 representation_size=None,
                 layer_scale_init_value=-1,
                 use_checkpoint_stages=[],
                 **kwargs):
        super().__init__()
        for cond, what in ((cpe_ks != 3, "cpe_ks != 3"), (not pre_norm, "post-norm"), (mlp_dwconv, "mlp_dwconv"),
                           (representation_size, "representation_size"), (layer_scale_init_value > 0, "layer scale"),
                           (qk_dims is not None and list(qk_dims) != list(embed_dim), "qk_dims != embed_dim"),
                           (not qkv_bias, "qkv_bias=False"), (qk_scale is not None, "qk_scale")):
            if cond:
                raise NotImplementedError(_UNSUPPORTED.format(what))
        for k in attn_type:
            if k not in ("C", "D", "S"):
                raise NotImplementedError("Attention type does not exit")    # reference message, models/lemevit.py:660
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.depth = list(depth)
        self.head_dim = head_dim
        self.mlp_ratios = list(mlp_ratios)
        self.attn_type = list(attn_type)
        self.in_chans = in_chans
        self.num_stages = len(attn_type)
        self.queries_len = queries_len
        self.drop_rate, self.drop_path_rate = drop_rate, drop_path_rate     # training-only knobs, kept for callers

        E = list(embed_dim)
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(
            nn.Conv2d(in_chans, E[0] // 2, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(E[0] // 2), nn.GELU(),
            nn.Conv2d(E[0] // 2, E[0], kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(E[0])))
        for i in range(self.num_stages - 1):
            if attn_type[i] == "C":
                self.downsample_layers.append(nn.Identity())
            else:
                self.downsample_layers.append(nn.Sequential(
                    nn.Conv2d(E[i], E[i + 1], kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(E[i + 1])))

        self.meta_tokens = nn.Parameter(torch.randn(queries_len, E[0]), requires_grad=True)
        self.meta_token_downsample = nn.ModuleList([_meta_mlp(E[0], E[0])])
        for i in range(self.num_stages - 1):
            self.meta_token_downsample.append(_meta_mlp(E[i], E[i + 1]))

        self.stages = nn.ModuleList()
        for i in range(self.num_stages):
            heads = E[i] // head_dim
            self.stages.append(nn.ModuleList([_Block(E[i], attn_type[i], heads, mlp_ratios[i]) for _ in range(depth[i])]))

        self.norm = nn.BatchNorm2d(E[-1])
        self.norm_c = nn.LayerNorm(E[-1])
        self.pre_logits = nn.Identity()
        self._build_head(num_classes)
        self.apply(self._init_weights)
        self._engine: Optional[Engine] = None
        self._engine_sig = None
        self.native_chunk = int(kwargs.pop("native_chunk", 0))

    # ---- reference surface -------------------------------------------------------------------------
    def _build_head(self, num_classes):
        self.head = nn.Linear(self.embed_dim[-1], num_classes) if num_classes > 0 else nn.Identity()

    def _init_weights(self, m):
        # models/lemevit.py:789-796
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        # the reference passes the whole embed_dim list to nn.Linear here (models/lemevit.py:805-807, latent bug);
        # the evident intent — a new head on the last stage width — is what callers (main.py) need.
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim[-1], num_classes) if num_classes > 0 else nn.Identity()
        self._drop_engine()

    # ---- native engine management ------------------------------------------------------------------
    def _signature(self, device):
        sig = [str(device), self.backbone_mode]
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def _drop_engine(self):
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
        self._engine = None
        self._engine_sig = None

    def native_engine(self, device) -> Engine:
        """Pack the current weights (again, if they changed) and return the engine for `device`."""
        sig = self._signature(device)
        if self._engine is None or sig != self._engine_sig:
            self._drop_engine()
            self._engine = Engine(self.state_dict(), depth=self.depth, embed_dim=list(self.embed_dim),
                                  mlp_ratios=self.mlp_ratios, attn_type=self.attn_type, head_dim=self.head_dim,
                                  queries_len=self.queries_len, num_classes=self.num_classes, in_chans=self.in_chans,
                                  backbone=self.backbone_mode, device=device, chunk=self.native_chunk)
            self._engine_sig = sig
        return self._engine

    def __deepcopy__(self, memo):
        # ModelEmaV2 deep-copies the model (reference main.py:316): copy parameters, never the native handle
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engine", "_engine_sig"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engine"] = None
        d["_engine_sig"] = None
        return d

    def _out_dtype(self, x):
        pd = self.meta_tokens.dtype
        return torch.bfloat16 if pd == torch.bfloat16 else torch.float32

    # ---- forward -----------------------------------------------------------------------------------
    def forward_features(self, x, c=None):
        """Pre-head features [B, C_last] (reference :809-829).  ``c`` is accepted for signature
        compatibility; the native path always starts from ``meta_tokens`` like ``forward`` does (:833)."""
        raise NotImplementedError("forward_features() of the classification model is fused into forward(); "
                                  "use forward(x), or the backbone class for multi-scale feature maps")

    def forward(self, x):
        require_cuda(x)
        if self.num_classes <= 0:
            raise NotImplementedError("num_classes == 0 is not implemented on the native path")
        eng = self.native_engine(x.device)
        y = eng.forward_cls(x, out_dtype=torch.float32)
        return y.to(self._out_dtype(x)) if self._out_dtype(x) != torch.float32 else y


# ---- model variants (reference models/lemevit.py:845-932) --------------------------------------------
_VARIANTS = {
    "lemevit_tiny": dict(depth=[1, 2, 2, 8, 2], embed_dim=[64, 64, 128, 192, 320]),
    "lemevit_small": dict(depth=[1, 2, 2, 6, 2], embed_dim=[96, 96, 192, 320, 384]),
    "lemevit_base": dict(depth=[2, 4, 4, 18, 4], embed_dim=[96, 96, 192, 384, 512]),
}
_COMMON = dict(head_dim=32, mlp_ratios=[4, 4, 4, 4, 4], attn_type=["C", "D", "D", "S", "S"], queries_len=16,
               qkv_bias=True, qk_scale=None, attn_drop=0., qk_dims=None, cpe_ks=3, pre_norm=True, mlp_dwconv=False,
               representation_size=None, layer_scale_init_value=-1, use_checkpoint_stages=[])


def _create(variant: str, pretrained, kwargs):
    model = LeMeViT(**_VARIANTS[variant], **_COMMON, **kwargs)
    model.default_cfg = _cfg()
    if pretrained:
        # reference: torch.load(pretrained)["model"] (models/lemevit.py:868-870)
        checkpoint = torch.load(pretrained, map_location="cpu")
        model.load_state_dict(checkpoint["model"])
    return model


def lemevit_tiny(pretrained=False, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_tiny", pretrained, kwargs)


def lemevit_small(pretrained=False, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_small", pretrained, kwargs)


def lemevit_base(pretrained=False, pretrained_cfg=None, pretrained_cfg_overlay=None, **kwargs):
    return _create("lemevit_base", pretrained, kwargs)
