"""fp64 torch emulation of what the native schedule (csrc/api.cu build_schedule) computes from the
PACKED weights — test infrastructure that lets the CPU suite validate pack.py's folds and ordering
against the oracle without a GPU."""
import math

import torch
import torch.nn.functional as F


def _gelu(t):
    return 0.5 * t * (1.0 + torch.erf(t / math.sqrt(2.0)))


def _norm(t, eps):
    mu = t.mean(-1, keepdim=True)
    var = ((t - mu) ** 2).mean(-1, keepdim=True)
    return (t - mu) / torch.sqrt(var + eps)


def _ln_linear(t, w, b, cs, eps=1e-6):
    """LayerNorm folded into the linear exactly as csrc/gemm.cu does it: r (t W^T - mu colsum) + b, with colsum the
    column sums of the STORED (bf16) weight."""
    K = t.shape[-1]
    assert torch.allclose(cs, w.sum(-1), rtol=1e-5, atol=1e-5), "colsum does not describe the packed weight"
    mu = t.sum(-1, keepdim=True) / K
    var = (t * t).sum(-1, keepdim=True) / K - mu * mu
    r = 1.0 / torch.sqrt(var + eps)
    return r * (t @ w.t() - mu * cs) + b


def _attn(q, k, v, heads, scale):
    B, Lq, C = q.shape
    d = C // heads
    qh = q.reshape(B, Lq, heads, d).transpose(1, 2)
    kh = k.reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, C)


def _conv_tokens(x_tok, H, W, w, b):
    """x_tok [B, H*W, Cin]; w packed [Cout, 9*Cin] with K = tap*Cin + ci; conv 3x3 s2 p1."""
    B, N, Cin = x_tok.shape
    x = x_tok.transpose(1, 2).reshape(B, Cin, H, W)
    w4 = w.reshape(w.shape[0], 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x, w4, b, stride=2, padding=1)
    return y.flatten(2).transpose(1, 2), y.shape[2], y.shape[3]


def _posembed(x_tok, H, W, dw_w, dw_b):
    B, N, C = x_tok.shape
    x = x_tok.transpose(1, 2).reshape(B, C, H, W)
    w4 = dw_w.t().reshape(C, 1, 3, 3)           # centre tap already holds the +1
    y = F.conv2d(x, w4, dw_b, stride=1, padding=1, groups=C)
    return y.flatten(2).transpose(1, 2)


def forward(packed, x, *, depth, embed_dim, attn_type, head_dim, queries_len, num_classes, in_chans, backbone,
            c_in=None, features=False):
    it = iter([t.detach().cpu().double() for t in packed])
    nx = lambda: next(it)
    M = queries_len
    B = x.shape[0]
    x = x.double()
    stem1_w, stem1_b, stem2_w, stem2_b, c0 = nx(), nx(), nx(), nx(), nx()
    C0 = embed_dim[0]
    w4 = stem1_w[:, : in_chans * 9].reshape(C0 // 2, in_chans, 3, 3)
    y = _gelu(F.conv2d(x, w4, stem1_b, stride=2, padding=1))
    H1, W1 = y.shape[2], y.shape[3]
    xt, H, W = _conv_tokens(y.flatten(2).transpose(1, 2), H1, W1, stem2_w, stem2_b)
    c = c0.unsqueeze(0).expand(B, -1, -1)   # pack-time constant meta_ds_0(meta_tokens) unless the caller brings its own c
    outs = []
    for i, kind in enumerate(attn_type):
        C = embed_dim[i]
        heads = C // head_dim
        if i > 0 and attn_type[i - 1] != "C":
            ds_w, ds_b = nx(), nx()
            xt, H, W = _conv_tokens(xt, H, W, ds_w, ds_b)
        w0, b0, g1, be1, w3, b3, g4, be4 = [nx() for _ in range(8)]
        if i == 0 and c_in is not None:
            c = c_in.double()
        if i > 0 or c_in is not None:
            c = _gelu(_norm(c @ w0.t() + b0, 1e-5) * g1 + be1)
            c = _norm(c @ w3.t() + b3, 1e-5) * g4 + be4
        N = H * W
        for j in range(depth[i]):
            dw_w, dw_b = nx(), nx()
            if kind == "C":
                wq, bq, csq, wkv, bkv, cskv, wkT, wp, bp = [nx() for _ in range(9)]
                assert torch.equal(wkT, wkv[:C].t())       # transposed key rows for the fused cross-attention kernels
            elif kind == "D":
                wa, ba, csa, wqT, wkT, wb, bb, csb, wpx, bpx, wpc, bpc = [nx() for _ in range(12)]
                assert torch.equal(wqT, wa[:C].t()) and torch.equal(wkT, wa[C:2 * C].t())
            else:
                wqkv, bqkv, csqkv, wp, bp = [nx() for _ in range(5)]
            w1, b1, cs1, w2, b2 = [nx() for _ in range(5)]
            mlp = lambda t: _gelu(_ln_linear(t, w1, b1, cs1)) @ w2.t() + b2
            xp = _posembed(xt, H, W, dw_w, dw_b)
            if kind == "C":
                q = _norm(c, 1e-6) @ wq.t() + bq
                kv = _ln_linear(xp, wkv, bkv, cskv)
                a = _attn(q, kv[..., :C], kv[..., C:], heads, head_dim ** -0.5)
                c = c + a @ wp.t() + bp
                c = c + mlp(c)
            elif kind == "D":
                xt = xp
                qkv1 = _ln_linear(xt, wa, ba, csa)
                qkv2 = _norm(c, 1e-6) @ wb.t() + bb
                s = C ** -0.5
                sx = math.log(M) / math.log(N) * s
                ax = _attn(qkv1[..., :C], qkv2[..., C:2 * C], qkv2[..., 2 * C:], heads, sx)
                ac = _attn(qkv2[..., :C], qkv1[..., C:2 * C], qkv1[..., 2 * C:], heads, s)
                xt = xt + ax @ wpx.t() + bpx
                c = c + ac @ wpc.t() + bpc
                xt = xt + mlp(xt)
                c = c + mlp(c)
            else:
                xt = xp
                toks = [xt] if backbone else [xt, c]
                res = []
                for t in toks:
                    qkv = _ln_linear(t, wqkv, bqkv, csqkv)
                    t = t + _attn(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, head_dim ** -0.5) @ wp.t() + bp
                    res.append(t + mlp(t))
                xt = res[0]
                if not backbone:
                    c = res[1]
        if backbone and i >= 1:
            outs.append(xt.transpose(1, 2).reshape(B, C, H, W))
    if backbone:
        return outs
    bn_s, bn_b, g, be = nx(), nx(), nx(), nx()
    feat = bn_s * xt.mean(1) + bn_b + (_norm(c, 1e-5) * g + be).mean(1)
    if num_classes > 0:
        hw, hb = nx(), nx()
    rest = list(it)
    assert not rest, f"{len(rest)} packed tensors were not consumed"
    if features or num_classes <= 0:
        return feat
    return feat @ hw.t() + hb
