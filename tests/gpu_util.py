"""Helpers for the -m gpu tests: raw-pointer calls into the C ABI on torch-owned device buffers."""
import ctypes as C
import math

import torch

from lemevit_b200 import _native

BF16, F32 = _native.DTYPE_BF16, _native.DTYPE_F32


def lib():
    return _native.load()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)   # data_ptr() of a view includes its storage offset


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ok(rc):
    _native.check(rc)


def bf(t):
    return t.to(torch.bfloat16).contiguous()


def rel_err(a, b):
    """max |a-b| / max |b|  — the metric of the stated tolerances."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def describe_mismatch(got, ref, tol):
    """Structure of an error pattern (helps decode descriptor / swizzle bugs from one GPU run)."""
    err = (got.double() - ref.double()).abs()
    bad = err > tol * ref.abs().max()
    rows = bad.any(dim=1).nonzero().flatten()
    cols = bad.any(dim=0).nonzero().flatten()
    msg = [f"bad elems {int(bad.sum())}/{bad.numel()}  max err {float(err.max()):.4g}  ref max {float(ref.abs().max()):.4g}"]
    if rows.numel():
        msg.append(f"bad rows: first {rows[:12].tolist()} ... count {rows.numel()} (mod 8 hist {torch.bincount(rows % 8, minlength=8).tolist()})")
        msg.append(f"bad cols: first {cols[:12].tolist()} ... count {cols.numel()} (mod 64 //8 hist {torch.bincount((cols % 64) // 8, minlength=8).tolist()})")
        r0 = int(rows[0])
        msg.append(f"row {r0} got {got[r0, :8].float().tolist()} ref {ref[r0, :8].float().tolist()}")
    return "\n".join(msg)


def linear(A, W, bias=None, residual=None, gelu=False, out_dtype=torch.bfloat16, simt=False, force_bn=0, ldc=None, out=None):
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N if ldc is None else ldc), dtype=out_dtype, device=A.device)
    ldc = out.stride(0)
    od = BF16 if out.dtype == torch.bfloat16 else F32
    L = lib()
    if simt:
        ok(L.lmv_linear_simt(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(residual), ptr(out), ldc, M, N, K, int(gelu), od, stream()))
    else:
        ok(L.lmv_linear(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(residual), ptr(out), ldc, M, N, K, int(gelu), od, force_bn, stream()))
    return out


def stats_parts(N, simt=False):
    return int(lib().lmv_linear_stats_parts(N, int(simt)))


def linear_fused(A, W, bias=None, residual=None, gelu=False, ln_stats=None, ln_colsum=None, ln_eps=0.0, stats_out=None,
                 simt=False, out=None):
    """ln_stats: [M, parts, 2] partial sums (or [M, 2]); stats_out: [M, stats_parts(N, simt), 2]."""
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=A.device)
    ok(lib().lmv_linear_fused(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(residual), ptr(out), out.stride(0), M, N, K,
                              int(gelu), BF16, ptr(ln_stats), (ln_stats.numel() // (2 * M)) if ln_stats is not None else 1,
                              ptr(ln_colsum), float(ln_eps), ptr(stats_out), int(simt), stream()))
    return out


def mlp_fused(x, W1, b1, W2, b2, ln_stats=None, colsum1=None, ln_eps=1e-6, resid=None, out=None):
    """out = resid + W2 gelu(W1 LN(x) + b1) + b2 in one kernel (lmv_mlp_fused); ln_stats [R, parts, 2] or None."""
    R, Cc = x.shape
    Hd = W1.shape[0]
    if out is None:
        out = torch.empty_like(x)
    parts = (ln_stats.numel() // (2 * R)) if ln_stats is not None else 1
    ok(lib().lmv_mlp_fused(ptr(x), ptr(resid), ptr(out), ptr(W1), ptr(b1), ptr(colsum1), ptr(W2), ptr(b2), ptr(ln_stats), parts,
                           float(ln_eps), R, Cc, Hd, stream()))
    return out


def ref_linear(A, W, bias=None, residual=None, gelu=False):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias.float()
    if gelu:
        y = 0.5 * y * (1.0 + torch.erf(y / math.sqrt(2.0)))
    if residual is not None:
        y = y + residual.float()
    return y


def attention(q, k, v, scale, impl=0):
    """q [B, Lq, h, 32], k/v [B, Lk, h, 32] (any strides with unit innermost stride) -> [B, Lq, h*32]."""
    B, Lq, h, d = q.shape
    Lk = k.shape[1]
    out = torch.empty((B, Lq, h * d), dtype=torch.bfloat16, device=q.device)
    for t in (q, k, v):
        assert t.stride(3) == 1 and t.stride(2) == d
    ok(lib().lmv_attention(ptr(q), q.stride(0), q.stride(1), ptr(k), k.stride(0), k.stride(1), ptr(v), v.stride(0), v.stride(1),
                           ptr(out), out.stride(0), out.stride(1), B, h, Lq, Lk, float(scale), impl, stream()))
    return out


def stem_conv1(x, w, bias):
    """First stem conv (3 -> C1, 3x3 / s2 / p1) + bias + GELU, direct kernel; x NCHW f32|bf16, w [C1, 3, 3, 3] fp32."""
    B, Cin, H, W = x.shape
    C1 = w.shape[0]
    wp = torch.zeros(C1, 32, device=x.device)
    wp[:, :27] = w.reshape(C1, 27)
    wp = bf(wp)
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    out = torch.empty((B, Ho * Wo, C1), dtype=torch.bfloat16, device=x.device)
    ok(lib().lmv_stem_conv1(ptr(x), BF16 if x.dtype == torch.bfloat16 else F32, ptr(wp), ptr(bias), ptr(out), B, Cin, C1, H, W, stream()))
    return out, wp


def attention_self(qkv, heads, N, scale):
    """qkv [B, T, 3C] packed (q | k | v, heads x 32 each): rows < N and rows >= N attend within their own segment."""
    B, T, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty((B, T, Cc), dtype=torch.bfloat16, device=qkv.device)
    q, k, v = qkv[:, :, :Cc], qkv[:, :, Cc:2 * Cc], qkv[:, :, 2 * Cc:]
    need = int(lib().lmv_attention_self_workspace(B, heads, T))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=qkv.device)
    ok(lib().lmv_attention_self(ptr(q), qkv.stride(0), qkv.stride(1), ptr(k), qkv.stride(0), qkv.stride(1), ptr(v), qkv.stride(0), qkv.stride(1),
                                ptr(out), out.stride(0), out.stride(1), B, heads, T, N, float(scale), ptr(ws), need, stream()))
    return out


def attention_meta(q, k, v, scale):
    """Few queries x many keys (meta-token side of the cross attentions): split-N tcgen05 kernel + merge."""
    B, Lq, h, d = q.shape
    Lk = k.shape[1]
    out = torch.empty((B, Lq, h * d), dtype=torch.bfloat16, device=q.device)
    need = int(lib().lmv_attention_meta_workspace(B, h, Lq, Lk))
    ws = torch.empty(need, dtype=torch.uint8, device=q.device)
    ok(lib().lmv_attention_meta(ptr(q), q.stride(0), q.stride(1), ptr(k), k.stride(0), k.stride(1), ptr(v), v.stride(0), v.stride(1),
                                ptr(out), out.stride(0), out.stride(1), B, h, Lq, Lk, float(scale), ptr(ws), ws.numel(), stream()))
    return out


def ref_attention(q, k, v, scale):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    o = (p @ vf).permute(0, 2, 1, 3)
    return o.reshape(o.shape[0], o.shape[1], -1)


def dca_block(kind, xt, c, W, heads, scale_x, scale_c, flags=0, stats_parts=1):
    """Fused 'C' / 'D' block core (lmv_dca_block).  xt [B, N, C] bf16, c [B, 16, C] bf16, W: dict of bf16 weights / fp32 biases
    (wa, ba, wb, bb, wp1, bp1, wp2, bp2, w1, b1, w2, b2).  Returns (xout, stats2, c_out, workspace)."""
    B, N, Cc = xt.shape
    Hd = W["w1"].shape[0]
    st = torch.stack([xt.float().sum(-1), (xt.float() ** 2).sum(-1)], dim=-1).reshape(B * N, 1, 2)
    if stats_parts > 1:
        st = (st / stats_parts).repeat(1, stats_parts, 1)
    st = st.contiguous()
    need = int(lib().lmv_dca_workspace_bytes(B, N, Cc, heads))
    assert need > 0
    ws = torch.zeros(need, dtype=torch.uint8, device=xt.device)
    c_out = c.clone()
    xout = torch.empty_like(xt) if kind == "D" else None
    stats2 = torch.zeros(B * N, 2, device=xt.device) if kind == "D" else None
    if kind == "D":   # transposed copies of the absorbed image-side projections (pack.py emits them next to qkv1 / kv)
        wxt = torch.stack([W["wa"][:Cc].t(), W["wa"][Cc:2 * Cc].t()]).contiguous()
    else:
        wxt = W["wb"][:Cc].t().contiguous()
    ok(lib().lmv_dca_block(ord(kind), ptr(xt), ptr(st), stats_parts, ptr(xout), ptr(stats2), ptr(c_out), ptr(W["wa"]), ptr(W["ba"]), ptr(W["wb"]),
                           ptr(W["bb"]), ptr(wxt), ptr(W["wp1"]), ptr(W["bp1"]), ptr(W.get("wp2")), ptr(W.get("bp2")), ptr(W["w1"]), ptr(W["b1"]),
                           ptr(W["w2"]), ptr(W["b2"]), B, N, Cc, heads, Hd, float(scale_x), float(scale_c), ptr(ws), ws.numel(), int(flags), stream()))
    return xout, stats2, c_out, ws


def ref_dca_block(kind, xt, c, W, heads, scale_x, scale_c):
    """The reference's op order in fp32 (models/lemevit.py:288-302 / :477-486 + :561-564 / :600-601): separate q/k/v projections of
    the LayerNorm-ed tokens, two scaled_dot_product_attentions, output projections, residuals and the meta-token MLP."""
    f = lambda k: W[k].float()
    B, N, Cc = xt.shape
    ln = lambda t: torch.nn.functional.layer_norm(t, (Cc,), eps=1e-6)
    xn, cn = ln(xt.float()), ln(c.float())
    hd = lambda t: t.reshape(t.shape[0], t.shape[1], heads, Cc // heads)
    if kind == "D":
        qkv1 = xn @ f("wa").t() + f("ba")
        qkv2 = cn @ f("wb").t() + f("bb")
        q1, k1, v1 = qkv1.split(Cc, dim=-1)
        q2, k2, v2 = qkv2.split(Cc, dim=-1)
        ax = ref_attention(hd(q1), hd(k2), hd(v2), scale_x)
        ac = ref_attention(hd(q2), hd(k1), hd(v1), scale_c)
        xout = xt.float() + ax @ f("wp1").t() + f("bp1")
        c1 = c.float() + ac @ f("wp2").t() + f("bp2")
    else:
        q = cn @ f("wa").t() + f("ba")
        kv = xn @ f("wb").t() + f("bb")
        k, v = kv.split(Cc, dim=-1)
        ac = ref_attention(hd(q), hd(k), hd(v), scale_c)
        xout = None
        c1 = c.float() + ac @ f("wp1").t() + f("bp1")
    h = ln(c1) @ f("w1").t() + f("b1")
    h = 0.5 * h * (1.0 + torch.erf(h / math.sqrt(2.0)))
    c2 = c1 + h @ f("w2").t() + f("b2")
    return xout, c2
