"""CPU oracle for the LeMeViT forward pass — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  ``lemevit_b200/`` never does.

What it is: a functional fp32 (or fp64) restatement, on plain torch CPU tensor ops, of the
eval-mode forward of the reference ``LeMeViT`` (``/root/reference/models/lemevit.py``) and of its
mmseg backbone copy (``/root/reference/semantic_segmentation/mmseg/models/backbones/lemevit.py``).
It consumes a ``state_dict`` with the reference's key names, so the same weights can be fed to
the reference module, to this oracle and to the CUDA path.

Where the arithmetic lives: the reference delegates every number to PyTorch (installed wheel
2.11.0+cu128 here; the reference pins nothing, README suggests torch 2.1): ``nn.Linear``,
``nn.LayerNorm``, ``nn.Conv2d``, ``nn.BatchNorm2d``, ``nn.GELU`` (exact erf) and
``F.scaled_dot_product_attention``.  Their published definitions are restated below, each function
citing the reference call site it follows.

Parity pin: the reference ships NO tests, golden vectors or checkpoints, so
the pin is the reference itself executed in the build container: ``oracle/gen_golden.py`` imports
the untouched reference through ``oracle/shims.py`` and stores its outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this oracle against those fixtures everywhere and against the live
reference whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration of the three published variants  (reference models/lemevit.py:845-932)
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    depth: Sequence[int]
    embed_dim: Sequence[int]
    head_dim: int = 32
    mlp_ratios: Sequence[float] = (4, 4, 4, 4, 4)
    attn_type: Sequence[str] = ("C", "D", "D", "S", "S")
    queries_len: int = 16
    num_classes: int = 1000
    in_chans: int = 3


VARIANTS: Dict[str, OracleConfig] = {
    "lemevit_tiny": OracleConfig(depth=(1, 2, 2, 8, 2), embed_dim=(64, 64, 128, 192, 320)),     # :849-850
    "lemevit_small": OracleConfig(depth=(1, 2, 2, 6, 2), embed_dim=(96, 96, 192, 320, 384)),    # :879-880
    "lemevit_base": OracleConfig(depth=(2, 4, 4, 18, 4), embed_dim=(96, 96, 192, 384, 512)),    # :909-910
    # not a published variant: a small configuration so that golden fixtures stay tiny
    "lemevit_micro": OracleConfig(depth=(1, 1, 1, 2, 1), embed_dim=(32, 32, 64, 96, 128)),
}


# --------------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------------
def linear(t: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear: t @ W^T + b with W[out, in]  (call sites :175,178,241-246,444-448,526-529)."""
    y = t @ w.t()
    return y if b is None else y + b


def layer_norm(t: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (call sites :513,525 eps=1e-6; :731,734,774 eps=1e-5)."""
    mu = t.mean(dim=-1, keepdim=True)
    var = ((t - mu) ** 2).mean(dim=-1, keepdim=True)
    return (t - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(t: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (call sites :528,701,732)."""
    return 0.5 * t * (1.0 + torch.erf(t * (1.0 / math.sqrt(2.0))))


def batch_norm_eval(x: Tensor, w: Tensor, b: Tensor, mean: Tensor, var: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in eval mode on NCHW (call sites :700,703,716,773)."""
    inv = w / torch.sqrt(var + eps)
    return x * inv[None, :, None, None] + (b - mean * inv)[None, :, None, None]


def sdpa(q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
    """softmax(scale * q k^T) v over the key axis; q,k,v are [B,h,L,d]  (:203,297,300,484; fp32 restatement :54-63)."""
    s = (q @ k.transpose(-1, -2)) * scale
    s = s - s.amax(dim=-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    return p @ v


def split_heads(t: Tensor, n_parts: int, heads: int) -> List[Tensor]:
    """'B L (x h d) -> x B h L d'  (:201,290,292,481)."""
    B, L, tot = t.shape
    d = tot // (n_parts * heads)
    t = t.reshape(B, L, n_parts, heads, d).permute(2, 0, 3, 1, 4)
    return [t[i] for i in range(n_parts)]


def merge_heads(t: Tensor) -> Tensor:
    """'B h L d -> B L (h d)'  (:204,298,301,485)."""
    B, h, L, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(B, L, h * d)


def to_tokens(x: Tensor) -> Tensor:
    """'N C H W -> N (H W) C'  (:548,591,621): token n = y*W + x."""
    B, C, H, W = x.shape
    return x.flatten(2).transpose(1, 2)


def to_map(t: Tensor, H: int, W: int) -> Tensor:
    """'N (H W) C -> N C H W'  (:579,647)."""
    B, N, C = t.shape
    return t.transpose(1, 2).reshape(B, C, H, W)


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
def _p(sd, prefix, name):
    return sd[prefix + name]


def mlp(sd, pre: str, t: Tensor) -> Tensor:
    """nn.Sequential(Linear, Identity, GELU, Linear)  (:526-530)."""
    hdn = gelu_erf(linear(t, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"]))
    return linear(hdn, sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"])


def pos_embed(sd, pre: str, x: Tensor) -> Tensor:
    """x + depthwise conv3x3(x), padding 1, bias  (:510 used at :546,589,619)."""
    C = x.shape[1]
    return x + F.conv2d(x, sd[pre + "pos_embed.weight"], sd[pre + "pos_embed.bias"], stride=1, padding=1, groups=C)


def norm1(sd, pre, t):
    return layer_norm(t, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)


def norm2(sd, pre, t):
    return layer_norm(t, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)


def block_c(sd, pre: str, x: Tensor, c: Tensor, heads: int, head_dim: int) -> Tuple[Tensor, Tensor]:
    """'C' block: LeMeBlock.forward_with_c (:584-613) + CrossAttention.forward (:477-486)."""
    xt = to_tokens(pos_embed(sd, pre, x))
    xn, cn = norm1(sd, pre, xt), norm1(sd, pre, c)
    q = linear(cn, sd[pre + "attn.q.weight"], sd[pre + "attn.q.bias"])
    kv = linear(xn, sd[pre + "attn.kv.weight"], sd[pre + "attn.kv.bias"])
    (qh,) = split_heads(q, 1, heads)
    kh, vh = split_heads(kv, 2, heads)
    a = merge_heads(sdpa(qh, kh, vh, head_dim ** -0.5))              # SDPA default scale
    c = c + linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])   # :600
    c = c + mlp(sd, pre, norm2(sd, pre, c))                                       # :601
    return x, c                                                                    # x unchanged :587,610


def block_d(sd, pre: str, x: Tensor, c: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """'D' block: LeMeBlock.forward_with_xc (:542-582) + DualCrossAttention.forward (:252-256,288-302)."""
    B, C, H, W = x.shape
    xt = to_tokens(pos_embed(sd, pre, x))
    N, M = xt.shape[1], c.shape[1]
    xn, cn = norm1(sd, pre, xt), norm1(sd, pre, c)
    scale = C ** -0.5                               # :235 — FULL channel dim
    scale_x = math.log(M, N) * scale                # :255
    scale_c = math.log(N, N) * scale                # :256
    q1, k1, v1 = split_heads(linear(xn, sd[pre + "attn.qkv1.weight"], sd[pre + "attn.qkv1.bias"]), 3, heads)
    q2, k2, v2 = split_heads(linear(cn, sd[pre + "attn.qkv2.weight"], sd[pre + "attn.qkv2.bias"]), 3, heads)
    dx = linear(merge_heads(sdpa(q1, k2, v2, scale_x)), sd[pre + "attn.proj_x.weight"], sd[pre + "attn.proj_x.bias"])
    dc = linear(merge_heads(sdpa(q2, k1, v1, scale_c)), sd[pre + "attn.proj_c.weight"], sd[pre + "attn.proj_c.bias"])
    xt = xt + dx                                    # :561
    xt = xt + mlp(sd, pre, norm2(sd, pre, xt))      # :562
    c = c + dc                                      # :563
    c = c + mlp(sd, pre, norm2(sd, pre, c))         # :564
    return to_map(xt, H, W), c


def self_attention(sd, pre: str, t: Tensor, heads: int, head_dim: int) -> Tensor:
    """StandardAttention.forward (:199-205)."""
    q, k, v = split_heads(linear(t, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]), 3, heads)
    a = merge_heads(sdpa(q, k, v, head_dim ** -0.5))
    return linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def block_s(sd, pre: str, x: Tensor, c: Tensor, heads: int, head_dim: int, update_c: bool) -> Tuple[Tensor, Tensor]:
    """'S' block: LeMeBlock.forward_with_x (:615-650).  The mmseg/mmdet/CD copies omit the two
    meta-token lines (semantic_segmentation/.../lemevit.py:630-636): ``update_c=False``."""
    B, C, H, W = x.shape
    xt = to_tokens(pos_embed(sd, pre, x))
    xt = xt + self_attention(sd, pre, norm1(sd, pre, xt), heads, head_dim)   # :632
    xt = xt + mlp(sd, pre, norm2(sd, pre, xt))                               # :633
    if update_c:
        c = c + self_attention(sd, pre, norm1(sd, pre, c), heads, head_dim)  # :634
        c = c + mlp(sd, pre, norm2(sd, pre, c))                              # :635
    return to_map(xt, H, W), c


def downsample(sd, i: int, x: Tensor, attn_type: Sequence[str]) -> Tensor:
    """downsample_layers[i]  (:698-704 stem; :711-717)."""
    p = f"downsample_layers.{i}."
    if i == 0:
        x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1)
        x = batch_norm_eval(x, sd[p + "1.weight"], sd[p + "1.bias"], sd[p + "1.running_mean"], sd[p + "1.running_var"])
        x = gelu_erf(x)
        x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=2, padding=1)
        return batch_norm_eval(x, sd[p + "4.weight"], sd[p + "4.bias"], sd[p + "4.running_mean"], sd[p + "4.running_var"])
    if attn_type[i - 1] == "C":
        return x                                                          # nn.Identity :711-712
    x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1)
    return batch_norm_eval(x, sd[p + "1.weight"], sd[p + "1.bias"], sd[p + "1.running_mean"], sd[p + "1.running_var"])


def meta_downsample(sd, i: int, c: Tensor) -> Tensor:
    """meta_token_downsample[i]: Linear -> LN(1e-5) -> GELU -> Linear -> LN(1e-5)  (:729-745)."""
    p = f"meta_token_downsample.{i}."
    c = linear(c, sd[p + "0.weight"], sd[p + "0.bias"])
    c = layer_norm(c, sd[p + "1.weight"], sd[p + "1.bias"], 1e-5)
    c = gelu_erf(c)
    c = linear(c, sd[p + "3.weight"], sd[p + "3.bias"])
    return layer_norm(c, sd[p + "4.weight"], sd[p + "4.bias"], 1e-5)


# --------------------------------------------------------------------------------------------
# whole model
# --------------------------------------------------------------------------------------------
def run_stages(sd, cfg: OracleConfig, x: Tensor, backbone: bool, taps: Optional[dict] = None, c: Optional[Tensor] = None):
    """Stage loop of forward_features(x, c) (:809-814).  Returns (x, c, [x after each stage]).
    ``c`` None: the model's own meta tokens, as LeMeViT.forward supplies them (:833)."""
    B = x.shape[0]
    if c is None:
        c = sd["meta_tokens"].unsqueeze(0).expand(B, -1, -1)              # :833 meta_tokens.repeat(B,1,1)
    outs = []
    for i, kind in enumerate(cfg.attn_type):
        x = downsample(sd, i, x, cfg.attn_type)
        c = meta_downsample(sd, i, c)
        heads = cfg.embed_dim[i] // cfg.head_dim                          # :749
        for j in range(cfg.depth[i]):
            pre = f"stages.{i}.{j}."
            if kind == "C":
                x, c = block_c(sd, pre, x, c, heads, cfg.head_dim)
            elif kind == "D":
                x, c = block_d(sd, pre, x, c, heads)
            elif kind == "S":
                x, c = block_s(sd, pre, x, c, heads, cfg.head_dim, update_c=not backbone)
            else:
                raise NotImplementedError("Attention type does not exit")  # reference message :660
            if taps is not None:
                taps[pre + "x"] = x
                taps[pre + "c"] = c
        outs.append(x)
    return x, c, outs


def forward_features_cls(sd, cfg: OracleConfig, x: Tensor, c: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """LeMeViT.forward_features(x, c) of the classification model (:809-829): pre-head features [B, C_last]."""
    x, c, _ = run_stages(sd, cfg, x, backbone=False, taps=taps, c=c)
    x = batch_norm_eval(x, sd["norm.weight"], sd["norm.bias"], sd["norm.running_mean"], sd["norm.running_var"])  # :815
    c = layer_norm(c, sd["norm_c.weight"], sd["norm_c.bias"], 1e-5)                                              # :818
    f = x.flatten(2).mean(-1) + c.transpose(-2, -1).mean(-1)                                                     # :825-827
    if taps is not None:
        taps["features"] = f
    return f


def forward_cls(sd, cfg: OracleConfig, x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """LeMeViT.forward (:831-836): forward_features on the model's own meta tokens, then the head."""
    f = forward_features_cls(sd, cfg, x, None, taps)
    return linear(f, sd["head.weight"], sd["head.bias"])                                                         # :835


def forward_backbone(sd, cfg: OracleConfig, x: Tensor) -> List[Tensor]:
    """mmseg copy: forward_features returns x after stages 1..4 as NCHW maps
    (semantic_segmentation/mmseg/models/backbones/lemevit.py:800-827)."""
    _, _, outs = run_stages(sd, cfg, x, backbone=True)
    return outs[1:]


def cast_state_dict(sd, dtype=torch.float32):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# algorithmic work (2 x MAC of every conv / linear / QK^T / AV contraction) — SURVEY.md §8(d)
# --------------------------------------------------------------------------------------------
def algorithmic_flops_per_image(cfg: OracleConfig, H: int, W: int, backbone: bool = False) -> float:
    M = cfg.queries_len
    mac = 0.0
    C0 = cfg.embed_dim[0]
    h, w = (H + 1) // 2, (W + 1) // 2
    mac += h * w * (C0 // 2) * cfg.in_chans * 9
    h, w = (h + 1) // 2, (w + 1) // 2
    mac += h * w * C0 * (C0 // 2) * 9
    prev = C0
    for i, kind in enumerate(cfg.attn_type):
        C = cfg.embed_dim[i]
        if i > 0 and cfg.attn_type[i - 1] != "C":
            h, w = (h + 1) // 2, (w + 1) // 2
            mac += h * w * C * prev * 9
        N = h * w
        mac += M * (prev * 4 * prev + 4 * prev * C)   # meta_token_downsample runs in every stage, also in the backbone copies
        r = cfg.mlp_ratios[i]
        for _ in range(cfg.depth[i]):
            mac += 9 * C * N                                     # depthwise pos-embed
            if kind == "C":
                mac += M * C * C + N * C * 2 * C + M * C * C     # q, kv, proj
                mac += 2 * M * N * C                             # QK^T + AV
                mac += 2 * r * C * C * M
            elif kind == "D":
                mac += 3 * N * C * C + 3 * M * C * C + N * C * C + M * C * C
                mac += 4 * M * N * C
                mac += 2 * r * C * C * (N + M)
            else:
                mac += 4 * N * C * C + 2 * N * N * C + 2 * r * C * C * N
                if not backbone:
                    mac += 4 * M * C * C + 2 * M * M * C + 2 * r * C * C * M
        prev = C
    if not backbone:
        mac += cfg.embed_dim[-1] * cfg.num_classes
    return 2.0 * mac
