"""Drop-in ``LeMeViT`` module: the reference's constructor, attributes, methods and state_dict
schema (reference models/lemevit.py:663-836), with ``forward`` executed by the native sm_100a
library instead of PyTorch ops.

The sub-modules below are PARAMETER CONTAINERS: they exist so that ``state_dict()`` /
``load_state_dict()`` / ``.to()`` / ``parameters()`` / ``deepcopy`` / DDP wrapping behave exactly
as for the reference (569 tensors for Base, same names and shapes — SURVEY.md §8b).  Their own
``forward`` is never called on the hot path.
"""
from __future__ import annotations

import threading
import warnings
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
        # native engines, one per device: {str(device): (Engine, weight signature)}.  nn.DataParallel replicas (shallow copies
        # of this module on other devices, reference validate.py:260-261) share this dict and its lock.
        self._engines = {}
        self._engines_lock = threading.Lock()
        self._src_sig = None      # replicas: the weight signature of the module they were replicated from
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
    def __setattr__(self, name, value):
        # module surgery (model.head = ..., swapping a stage) changes the tensor set the signature walks
        if isinstance(value, (nn.Module, nn.Parameter)) and "_sig_tensors" in self.__dict__:
            self.__dict__["_sig_tensors"] = None
        super().__setattr__(name, value)

    def _signature(self):
        """(storage address, version counter) of every parameter / buffer: changes on in-place edits (optimizer steps,
        load_state_dict), on ``.to()`` and on re-assignment, which is when the packed weights must be rebuilt.  The tensor
        list itself is cached (walking the module tree of ~570 tensors costs more than reading their versions).
        A DataParallel replica answers with the signature of its source module: its own parameters are fresh broadcast
        copies on every forward, their values are the source's."""
        if self.__dict__.get("_src_sig") is not None:
            return self.__dict__["_src_sig"]
        ts = self.__dict__.get("_sig_tensors")
        if ts is None:
            ts = list(self.parameters()) + list(self.buffers())
            self.__dict__["_sig_tensors"] = ts
        return (self.backbone_mode, len(ts), tuple([(t.data_ptr(), t._version) for t in ts]))

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__["_src_sig"] = self._signature()
        replica.__dict__["_sig_tensors"] = None
        keys = self.__dict__.get("_sd_keys")
        if keys is None:
            keys = self.__dict__["_sd_keys"] = list(self.state_dict().keys())
        replica.__dict__["_sd_keys"] = keys
        return replica

    def _weights(self):
        """state_dict of this module; for a DataParallel replica (whose parameters are plain tensor attributes, not registered
        Parameters, so that ``state_dict()`` is empty) the same keys resolved attribute by attribute."""
        if self.__dict__.get("_src_sig") is None:
            return self.state_dict()
        out = {}
        for key in self.__dict__["_sd_keys"]:
            obj = self
            for part in key.split("."):
                obj = getattr(obj, part)
            out[key] = obj
        return out

    def _drop_engine(self):
        """Close every native engine of this module (weights changed / module surgery)."""
        engines = self.__dict__.get("_engines")
        if engines:
            with self._engines_lock:
                for eng, _ in engines.values():
                    eng.close()
                engines.clear()
        self.__dict__["_sig_tensors"] = None
        self.__dict__["_sd_keys"] = None

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_sig_tensors"] = None      # .to() / .half() / .cuda() may replace Parameter objects
        return super()._apply(fn, *args, **kwargs)

    def native_engine(self, device) -> Engine:
        """Pack the current weights (again, if they changed) and return the engine for `device`."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key, sig = str(device), self._signature()
        ent = self._engines.get(key)
        if ent is not None and ent[1] == sig:
            return ent[0]
        with self._engines_lock:
            ent = self._engines.get(key)
            if ent is not None and ent[1] == sig:
                return ent[0]
            if ent is not None:
                ent[0].close()
                del self._engines[key]
            if self.__dict__.get("_src_sig") is None:
                # the weights of the source module changed: engines of other devices hold stale copies too
                for k in [k for k, (e, sg) in self._engines.items() if sg != sig]:
                    self._engines.pop(k)[0].close()
            eng = Engine(self._weights(), depth=self.depth, embed_dim=list(self.embed_dim),
                         mlp_ratios=self.mlp_ratios, attn_type=self.attn_type, head_dim=self.head_dim,
                         queries_len=self.queries_len, num_classes=self.num_classes, in_chans=self.in_chans,
                         backbone=self.backbone_mode, device=device, chunk=self.native_chunk)
            norm = self.__dict__.get("_input_norm")
            if norm is not None:
                eng.set_input_norm(*norm)
            self._engines[key] = (eng, sig)
        return eng

    def set_input_norm(self, mean, std):
        """Mean / std (per channel, pixel units 0..255) applied to ``torch.uint8`` inputs inside the first stem convolution:
        ``model(x_u8)`` == ``model(((x_u8.float() - mean) / std).to(bfloat16))`` with x_u8 ``[B, 3, H, W]`` (planar or
        channels_last) or ``[B, H, W, 3]``.  This is the GPU half of the timm prefetcher the reference's training and validation
        loaders run (main.py:399-428, mean / std x 255 of ``default_cfg``); default: the ImageNet constants of ``_cfg()``."""
        mean, std = [float(v) for v in mean], [float(v) for v in std]
        if len(mean) != 3 or len(std) != 3:
            raise ValueError("set_input_norm: three channels expected")
        self.__dict__["_input_norm"] = (mean, std)
        for eng, _ in list(self.__dict__.get("_engines", {}).values()):
            eng.set_input_norm(mean, std)

    def __deepcopy__(self, memo):
        # ModelEmaV2 deep-copies the model (reference main.py:316): copy parameters, never the native handles
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_engines":
                new.__dict__[k] = {}
            elif k == "_engines_lock":
                new.__dict__[k] = threading.Lock()
            elif k in ("_sig_tensors", "_src_sig", "_sd_keys"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        d["_engines_lock"] = None
        d["_sig_tensors"] = None
        d["_src_sig"] = None
        d["_sd_keys"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.__dict__["_engines_lock"] = threading.Lock()

    def _out_dtype(self, x):
        pd = self.meta_tokens.dtype
        return torch.bfloat16 if pd == torch.bfloat16 else torch.float32

    _warned_grad = False

    def _check_inference(self):
        """The native path is the eval-mode forward (BatchNorm running statistics folded into the convolutions, DropPath and
        dropout absent, no autograd graph).  In train mode the reference would use batch statistics, update the running
        stats and apply stochastic depth (models/lemevit.py:531,700-703) — refusing is the only answer that cannot silently
        give different numbers."""
        if self.training:
            raise RuntimeError("lemevit_b200 is inference-only: forward() implements the eval-mode LeMeViT forward on native sm_100a "
                               "kernels (no BatchNorm batch statistics, no DropPath, no autograd). Call model.eval() first; "
                               "train with the reference implementation and load the state_dict here.")
        if torch.is_grad_enabled() and not LeMeViT._warned_grad and any(p.requires_grad for p in self.parameters()):
            LeMeViT._warned_grad = True
            warnings.warn("lemevit_b200: forward() runs outside autograd — the output carries no grad_fn even though gradients are "
                          "enabled; wrap inference in torch.no_grad() (this warning is shown once).", stacklevel=3)

    # ---- forward -----------------------------------------------------------------------------------
    def forward_features(self, x, c=None):
        """Pre-head features ``[B, embed_dim[-1]]`` = mean_HW(BN(x)) + mean_M(LN(c)) (reference :809-829).
        ``c``: meta tokens ``[B, queries_len, embed_dim[0]]`` as the reference takes them (they go through
        ``meta_token_downsample[0]`` first, :813); None = ``self.meta_tokens`` broadcast over the batch, which is what
        ``forward`` passes (:833) and is constant-folded at weight-pack time."""
        require_cuda(x)
        self._check_inference()
        eng = self.native_engine(x.device)
        feat = torch.empty((x.shape[0], self.embed_dim[-1]), dtype=torch.bfloat16, device=x.device)
        eng.forward_cls(x, meta_tokens=c, features=feat, want_logits=False)
        dt = self._out_dtype(x)
        return feat if dt == torch.bfloat16 else feat.to(dt)

    def forward(self, x):
        require_cuda(x)
        self._check_inference()
        if self.num_classes <= 0:
            return self.forward_features(x)      # head = nn.Identity (reference :786, reset_classifier(0))
        eng = self.native_engine(x.device)
        return eng.forward_cls(x, out_dtype=self._out_dtype(x))      # the head GEMM writes bf16 / fp32 logits directly


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
