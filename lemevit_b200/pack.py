"""Weight packing: reference state_dict -> the folded bf16/f32 tensor list the native plan consumes.

All folds are exact algebra done once in float64 on the host, then cast ONCE to the storage type
(SURVEY.md Appendix A "Algebraic folds"):

  1. ``x + dwconv(x)``            -> depthwise taps with +1 on the centre tap        (models/lemevit.py:546,589,619)
  2. conv -> BatchNorm(eval)      -> conv with scaled weights / shifted bias         (:699-703, :715-716)
  3. LayerNorm affine -> Linear   -> Linear(W·diag(gamma), b + W·beta) on the normalised input; norm1 feeds
     q/kv/qkv1/qkv2/qkv, norm2 feeds mlp.0; x and c share norm1/norm2 so one folded set serves both (:560-564,600-601,632-635)
  4. tail: mean(BN(x)) = BN_affine(mean(x))                                         (:815,825)
  5. ``meta_token_downsample[0](meta_tokens)`` is batch-invariant (c = meta_tokens.repeat, :833) -> precomputed

ORDER of the packed list (checked entry by entry by ``lmv_plan_create`` in csrc/api.cu):

  stem1_w bf16[C0/2, Kp0]  stem1_b f32[C0/2]  stem2_w bf16[C0, 9*C0/2]  stem2_b f32[C0]  c0_init bf16[M, C0]
  for each stage i (Cp = C(i-1), C0 for i = 0):
      if i > 0:  [ds_w bf16[Ci, 9*Cp], ds_b f32[Ci]]   (absent when stage i-1 is 'C': nn.Identity, :711-712)
      md_w0 bf16[4Cp,Cp] md_b0 f32 md_g1 f32 md_be1 f32 md_w3 bf16[Ci,4Cp] md_b3 f32 md_g4 f32 md_be4 f32
      for each block:  dw_w f32[9,C] dw_b f32[C]
          'C': q_w q_b q_cs kv_w kv_b kv_cs kv_kT proj_w proj_b
          'D': qkv1_w qkv1_b qkv1_cs qkv1_qT qkv1_kT qkv2_w qkv2_b qkv2_cs proj_x_w proj_x_b proj_c_w proj_c_b
               (*_qT / *_kT: bf16 [C, C] transposes of the query / key rows, for the fused cross-attention blocks)
          'S': qkv_w qkv_b qkv_cs proj_w proj_b          then: mlp0_w mlp0_b mlp0_cs mlp3_w mlp3_b
          (weights bf16 [out,in], biases f32, *_cs = f32[out] column sums of the bf16 LayerNorm-folded weight)
  classification only:  bn_scale bn_shift norm_c_g norm_c_b (f32[CL])  [head_w bf16[ncls, CL], head_b f32[ncls]]

conv weights are stored [Cout, K] with K = (ky*3 + kx)*Cin + ci (matches the im2col kernels), except stem1
which keeps K = ci*9 + ky*3 + kx zero-padded to Kp0 = round_up(9*in_chans, 8).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch

BN_EPS = 1e-5


def _d(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(device="cpu", dtype=torch.float64)


def _fold_conv_bn(sd, conv: str, bn: str):
    w, b = _d(sd[conv + ".weight"]), _d(sd[conv + ".bias"])
    s = _d(sd[bn + ".weight"]) / torch.sqrt(_d(sd[bn + ".running_var"]) + BN_EPS)
    return w * s[:, None, None, None], (b - _d(sd[bn + ".running_mean"])) * s + _d(sd[bn + ".bias"])


def _fold_ln_linear(sd, ln: str, lin: str):
    w, b = _d(sd[lin + ".weight"]), _d(sd[lin + ".bias"])
    g, be = _d(sd[ln + ".weight"]), _d(sd[ln + ".bias"])
    return w * g[None, :], b + w @ be


def _gelu(t):
    return 0.5 * t * (1.0 + torch.erf(t / math.sqrt(2.0)))


def _ln(t, g, b, eps):
    mu = t.mean(-1, keepdim=True)
    var = ((t - mu) ** 2).mean(-1, keepdim=True)
    return (t - mu) / torch.sqrt(var + eps) * g + b


def pack_state_dict(sd: Dict[str, torch.Tensor], *, depth: Sequence[int], embed_dim: Sequence[int],
                    attn_type: Sequence[str], in_chans: int, num_classes: int, backbone: bool,
                    device: torch.device) -> List[torch.Tensor]:
    out: List[torch.Tensor] = []

    def h(t):  # bf16 weight
        out.append(t.to(torch.bfloat16).contiguous().to(device))

    def f(t):  # f32 vector
        out.append(t.to(torch.float32).contiguous().to(device))

    E = list(embed_dim)
    C0 = E[0]
    # ---- stem
    w, b = _fold_conv_bn(sd, "downsample_layers.0.0", "downsample_layers.0.1")
    kp0 = (in_chans * 9 + 7) // 8 * 8
    w1 = torch.zeros(C0 // 2, kp0, dtype=torch.float64)
    w1[:, : in_chans * 9] = w.reshape(C0 // 2, in_chans * 9)
    h(w1); f(b)
    w, b = _fold_conv_bn(sd, "downsample_layers.0.3", "downsample_layers.0.4")
    h(w.permute(0, 2, 3, 1).reshape(C0, 9 * (C0 // 2))); f(b)
    # ---- c0 = meta_token_downsample[0](meta_tokens)   (models/lemevit.py:729-735)
    p = "meta_token_downsample.0."
    c0 = _d(sd["meta_tokens"])
    c0 = c0 @ _d(sd[p + "0.weight"]).t() + _d(sd[p + "0.bias"])
    c0 = _gelu(_ln(c0, _d(sd[p + "1.weight"]), _d(sd[p + "1.bias"]), 1e-5))
    c0 = c0 @ _d(sd[p + "3.weight"]).t() + _d(sd[p + "3.bias"])
    c0 = _ln(c0, _d(sd[p + "4.weight"]), _d(sd[p + "4.bias"]), 1e-5)
    h(c0)
    # ---- stages
    for i, kind in enumerate(attn_type):
        C = E[i]
        if i > 0 and attn_type[i - 1] != "C":
            w, b = _fold_conv_bn(sd, f"downsample_layers.{i}.0", f"downsample_layers.{i}.1")
            h(w.permute(0, 2, 3, 1).reshape(C, 9 * E[i - 1])); f(b)
        # meta_token_downsample[i]; [0] is only executed for caller-supplied meta tokens (forward_features(x, c))
        p = f"meta_token_downsample.{i}."
        h(_d(sd[p + "0.weight"])); f(_d(sd[p + "0.bias"]))
        f(_d(sd[p + "1.weight"])); f(_d(sd[p + "1.bias"]))
        h(_d(sd[p + "3.weight"])); f(_d(sd[p + "3.bias"]))
        f(_d(sd[p + "4.weight"])); f(_d(sd[p + "4.bias"]))
        for j in range(depth[i]):
            p = f"stages.{i}.{j}."
            dw = _d(sd[p + "pos_embed.weight"]).reshape(C, 9).t().clone()   # [9, C], tap = ky*3 + kx
            dw[4] += 1.0                                                     # x + dwconv(x)
            f(dw); f(_d(sd[p + "pos_embed.bias"]))
            n1, n2 = p + "norm1", p + "norm2"
            if kind == "C":
                names = [("attn.q", n1), ("attn.kv", n1), ("attn.proj", None)]
            elif kind == "D":
                names = [("attn.qkv1", n1), ("attn.qkv2", n1), ("attn.proj_x", None), ("attn.proj_c", None)]
            elif kind == "S":
                names = [("attn.qkv", n1), ("attn.proj", None)]
            else:
                raise NotImplementedError("Attention type does not exit")
            names += [("mlp.0", n2), ("mlp.3", None)]
            for lin, ln in names:
                if ln is None:
                    w, b = _d(sd[p + lin + ".weight"]), _d(sd[p + lin + ".bias"])
                    h(w); f(b)
                else:
                    w, b = _fold_ln_linear(sd, ln, p + lin)
                    h(w); f(b)
                    # column sums of the weights AS STORED (bf16): LN(y) W^T = r (y W^T - mu colsum), csrc/gemm.cu
                    f(w.to(torch.bfloat16).to(torch.float64).sum(dim=1))
                    # transposed copies of the image-side query / key projections for the fused cross-attention blocks, which absorb
                    # them into per-image operands (csrc/kernels.h): T[(h,m)][j] = sum_i vec[m][i] W[i][j] needs W^T with i contiguous
                    if lin == "attn.qkv1":
                        h(w[:C].t()); h(w[C:2 * C].t())
                    elif lin == "attn.kv":
                        h(w[:C].t())
    # ---- tail
    if not backbone:
        s = _d(sd["norm.weight"]) / torch.sqrt(_d(sd["norm.running_var"]) + BN_EPS)
        f(s); f(_d(sd["norm.bias"]) - _d(sd["norm.running_mean"]) * s)
        f(_d(sd["norm_c.weight"])); f(_d(sd["norm_c.bias"]))
        if num_classes > 0:
            h(_d(sd["head.weight"])); f(_d(sd["head.bias"]))
    return out
