"""Single-kernel targets for ncu captures.  usage: python tools/ncu_kernel.py mlp R C Hd | gemm M N K [gelu] [ln]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as G

kind = sys.argv[1]
torch.manual_seed(0)
if kind == "mlp":
    R, C, Hd = (int(v) for v in sys.argv[2:5])
    x = G.bf(torch.randn(R, C, device="cuda"))
    W1, W2 = G.bf(torch.randn(Hd, C, device="cuda") * C ** -0.5), G.bf(torch.randn(C, Hd, device="cuda") * Hd ** -0.5)
    b1, b2 = torch.randn(Hd, device="cuda"), torch.randn(C, device="cuda")
    stats = torch.stack([x.float().sum(-1), (x.float() ** 2).sum(-1)], dim=1).contiguous()
    cs = W1.float().sum(-1).contiguous()
    for _ in range(3):
        G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=cs)
elif kind == "gemm":
    M, N, K = (int(v) for v in sys.argv[2:5])
    A, W = G.bf(torch.randn(M, K, device="cuda")), G.bf(torch.randn(N, K, device="cuda") * K ** -0.5)
    bias = torch.randn(N, device="cuda")
    for _ in range(3):
        G.linear(A, W, bias, gelu="gelu" in sys.argv)
torch.cuda.synchronize()
print("done")
if kind == "attn_self":
    B, h, T, N = (int(v) for v in sys.argv[2:6])
    qkv = G.bf(torch.randn(B, T, 3 * h * 32, device="cuda"))
    for _ in range(3):
        G.attention_self(qkv, h, N, 32 ** -0.5)
    torch.cuda.synchronize()
    print("done attn_self")
if kind == "posln":
    B, H, W, C, M = (int(v) for v in sys.argv[2:7])
    T = H * W + M
    tok = G.bf(torch.randn(B, T, C, device="cuda"))
    w9 = torch.randn(9, C, device="cuda") * 0.1
    db = torch.randn(C, device="cuda") * 0.1
    out = torch.empty_like(tok)
    stats = torch.empty(B * T, 2, device="cuda")
    for _ in range(3):
        G.ok(G.lib().lmv_posembed_layernorm(G.ptr(tok), G.ptr(w9), G.ptr(db), G.ptr(out), None, G.ptr(stats), B, H, W, T, C, 1e-6, G.stream()))
    torch.cuda.synchronize()
    print("done posln")
if kind == "attn_meta":
    B, h, Lq, Lk = (int(v) for v in sys.argv[2:6])
    C = h * 32
    qc = G.bf(torch.randn(B, Lq, 3 * C, device="cuda"))
    kvx = G.bf(torch.randn(B, Lk, 3 * C, device="cuda"))
    q = qc[:, :, :C].unflatten(2, (h, 32))
    k = kvx[:, :, C:2 * C].unflatten(2, (h, 32))
    v = kvx[:, :, 2 * C:].unflatten(2, (h, 32))
    for _ in range(3):
        G.attention_meta(q, k, v, C ** -0.5)
    torch.cuda.synchronize()
    print("done attn_meta")
