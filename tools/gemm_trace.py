"""Cycle accounting of the GEMM kernel's three roles (TMA producer, MMA issuer, epilogue warps) on the LeMeViT-Base shapes.

Needs a debug build:  LMV_NVCC_EXTRA=-DLMV_GEMM_TRACE python -m lemevit_b200.build --force   (then rebuild without it).
Counters per warp, summed over its tiles (cycles):
  producer : [0] wait for a free ring slot   [1] issue                         [7] total
  mma      : [0] wait acc_empty  [1] wait smem full  [2] issue                 [7] total
  epilogue : [0] wait acc_full   [1] tcgen05.ld   [2] transpose stores   [3] reads+math+global stores
             [4] tile setup + constant loads      [5] release + stats          [7] total
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G  # noqa: E402


def run(name, M, N, K, ln=False, gelu=False, res=False, stats=False):
    dev = "cuda"
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    lnst = torch.stack([A.float().sum(1), (A.float() ** 2).sum(1)], 1).contiguous() if ln else None
    cs = W.float().sum(1).contiguous() if ln else None
    resid = torch.randn(M, N, device=dev).bfloat16() if res else None
    so = torch.empty(M, G.stats_parts(N), 2, device=dev) if stats else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L = G.lib()
    buf = (C.c_ulonglong * (148 * 10 * 8))()
    for _ in range(3):
        G.linear_fused(A, W, bias, resid, gelu, lnst, cs, 1e-6, so, out=out)
    L.lmv_debug_gemm_trace(buf, len(buf))
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        G.linear_fused(A, W, bias, resid, gelu, lnst, cs, 1e-6, so, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    L.lmv_debug_gemm_trace(buf, len(buf))
    t = torch.tensor(list(buf), dtype=torch.float64).view(148, 10, 8) / reps
    prod, mma, epi = t[:, 0].mean(0), t[:, 1].mean(0), t[:, 2:].mean((0, 1))
    f = lambda v: " ".join(f"{int(x):8d}" for x in v)
    print(f"{name}: M={M} N={N} K={K}  {us:.1f} us")
    print(f"  producer  wait_slot,issue,-,-,-,-,-,total      : {f(prod)}")
    print(f"  mma       wait_acc,wait_full,issue,-,-,-,-,total : {f(mma)}")
    print(f"  epilogue  wait_acc,ldtm,sts,math+stg,setup,release,-,total : {f(epi)}")
    print(f"  epilogue max-warp total {int(t[:, 2:, 7].max())}  min {int(t[:, 2:, 7].min())}")


if __name__ == "__main__":
    if "--short" in sys.argv:
        run("plain 54272x1536x384", 54272, 1536, 384)
        run("qkv  ln", 54272, 1152, 384, ln=True)
        run("fc1  ln+gelu", 54272, 1536, 384, ln=True, gelu=True)
        sys.exit(0)
    run("fc1  ln+gelu", 54272, 1536, 384, ln=True, gelu=True)
    run("qkv  ln", 54272, 1152, 384, ln=True)
    run("proj res+stats", 54272, 384, 384, res=True, stats=True)
    run("fc2  res", 54272, 384, 1536, res=True)
    run("s1 qkv ln", 802816, 288, 96, ln=True)
    run("s1 proj res+stats", 802816, 96, 96, res=True, stats=True)
    run("plain 54272x1536x384", 54272, 1536, 384)
