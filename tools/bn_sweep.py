"""Tile-width sweep of the GEMM kernel (plain bias epilogue) on the LeMeViT-Base shapes: which BN should pick_bn() choose?"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as G

def run(M, N, K, bns):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for bn in bns:
        try:
            for _ in range(2): G.linear(A, W, bias, force_bn=bn, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): G.linear(A, W, bias, force_bn=bn, out=out)
            e1.record(); torch.cuda.synchronize()
            res.append((bn, round(e0.elapsed_time(e1) * 200, 1)))
        except Exception as ex:
            res.append((bn, "err")); torch.cuda.synchronize()
    print(f"M={M} N={N} K={K}: " + "  ".join(f"BN={b}:{t}us" for b, t in res))

run(802816, 288, 96, [0, 96, 160, 192, 288 // 2 // 32 * 32])
run(802816, 96, 96, [0, 96, 64, 32])
run(802816, 192, 96, [0, 192, 96, 64])
run(200704, 576, 192, [0, 192, 128, 96])
run(54272, 1152, 384, [0, 192, 128, 256, 96])
run(54272, 1536, 384, [0, 256, 192, 128])
run(54272, 384, 1536, [0, 192, 128, 96, 256])
run(54272, 384, 384, [0, 192, 128, 96])
