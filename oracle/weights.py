"""Seeded synthetic weights with the reference's state_dict schema — TEST INFRASTRUCTURE.

The reference has no released checkpoint (README.md:79), so parity runs use random weights:
Linear weights ~ trunc_normal(std=.02) as in the reference init (models/lemevit.py:789-796) but
scaled up a little so that attention is not uniform, PLUS randomised biases, LayerNorm/BatchNorm
affine and BatchNorm running statistics — with the defaults (gamma=1, beta=0, mean=0, var=1) every
norm fold would be an identity and folding bugs would be invisible.

Everything is drawn from one ``torch.Generator`` on CPU in a fixed key order, so the same
(variant, seed) yields bit-identical tensors in the build container and on the GPU box (same
torch wheel).  ``fingerprint`` lets a golden fixture assert that.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from .lemevit_oracle import OracleConfig


def state_dict_spec(cfg: OracleConfig, backbone: bool = False) -> "OrderedDict[str, tuple]":
    """Key -> (shape, kind) in the reference's registration order (models/lemevit.py:696-786)."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    E = list(cfg.embed_dim)

    def lin(p, o, i):
        spec[p + ".weight"] = ((o, i), "lin_w")
        spec[p + ".bias"] = ((o,), "bias")

    def ln(p, d):
        spec[p + ".weight"] = ((d,), "gamma")
        spec[p + ".bias"] = ((d,), "beta")

    def bn(p, d):
        spec[p + ".weight"] = ((d,), "gamma")
        spec[p + ".bias"] = ((d,), "beta")
        spec[p + ".running_mean"] = ((d,), "mean")
        spec[p + ".running_var"] = ((d,), "var")
        spec[p + ".num_batches_tracked"] = ((), "count")

    def conv(p, o, i, k=3, groups=1):
        spec[p + ".weight"] = ((o, i // groups, k, k), "conv_w")
        spec[p + ".bias"] = ((o,), "bias")

    spec["meta_tokens"] = ((cfg.queries_len, E[0]), "tokens")
    conv("downsample_layers.0.0", E[0] // 2, cfg.in_chans)
    bn("downsample_layers.0.1", E[0] // 2)
    conv("downsample_layers.0.3", E[0], E[0] // 2)
    bn("downsample_layers.0.4", E[0])
    for i in range(1, len(E)):
        if cfg.attn_type[i - 1] == "C":
            continue
        conv(f"downsample_layers.{i}.0", E[i], E[i - 1])
        bn(f"downsample_layers.{i}.1", E[i])
    for i in range(len(E)):
        prev = E[0] if i == 0 else E[i - 1]
        p = f"meta_token_downsample.{i}"
        lin(p + ".0", 4 * prev, prev)
        ln(p + ".1", 4 * prev)
        lin(p + ".3", E[i], 4 * prev)
        ln(p + ".4", E[i])
    for i, kind in enumerate(cfg.attn_type):
        C = E[i]
        hid = int(cfg.mlp_ratios[i] * C)
        for j in range(cfg.depth[i]):
            p = f"stages.{i}.{j}"
            conv(p + ".pos_embed", C, C, groups=C)
            ln(p + ".norm1", C)
            if kind == "C":
                lin(p + ".attn.q", C, C)
                lin(p + ".attn.kv", 2 * C, C)
                lin(p + ".attn.proj", C, C)
            elif kind == "D":
                lin(p + ".attn.qkv1", 3 * C, C)
                lin(p + ".attn.qkv2", 3 * C, C)
                lin(p + ".attn.proj_x", C, C)
                lin(p + ".attn.proj_c", C, C)
            else:
                lin(p + ".attn.qkv", 3 * C, C)
                lin(p + ".attn.proj", C, C)
            ln(p + ".norm2", C)
            lin(p + ".mlp.0", hid, C)
            lin(p + ".mlp.3", C, hid)
    bn("norm", E[-1])
    ln("norm_c", E[-1])
    if not backbone and cfg.num_classes > 0:
        lin("head", cfg.num_classes, E[-1])
    return spec


def make_state_dict(cfg: OracleConfig, seed: int = 0, backbone: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for key, (shape, kind) in state_dict_spec(cfg, backbone).items():
        if kind == "lin_w":
            # gain/sqrt(fan_in): residual branches stay O(1) through 32 blocks; q/k/v projections get a
            # larger gain so that the softmaxes are far from uniform.
            gain = 1.5 if (".attn.q" in key or ".attn.kv" in key) else 0.6
            t = torch.randn(shape, generator=g).clamp_(-2, 2) * (gain / shape[1] ** 0.5)
        elif kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            # depthwise pos-embed convs sit on the residual stream (x + dw(x)) in all 32 blocks: keep them small
            t = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5 * (0.25 if "pos_embed" in key else 1.0)
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "gamma":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind == "beta":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "mean":
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "tokens":
            t = torch.randn(shape, generator=g)
        elif kind == "count":
            t = torch.tensor(0, dtype=torch.long)
        else:  # pragma: no cover
            raise KeyError(kind)
        sd[key] = t
    return sd


def make_input(batch: int, H: int, W: int, seed: int = 0, in_chans: int = 3) -> torch.Tensor:
    """Standard-normal synthetic images, as benchmark.py:463 (torch.randn)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(77 + seed)
    return torch.randn((batch, in_chans, H, W), generator=g)


def fingerprint(sd: Dict[str, torch.Tensor]) -> float:
    """One float64 number that changes if any weight does."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        if v.is_floating_point():
            acc += float(v.double().abs().sum()) * (1.0 + (i % 7))
    return acc
