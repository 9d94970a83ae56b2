"""Cycle account of the CTA-pair MLP kernel's MMA issuer (needs the -DLMV_MLP_TRACE build):
   LMV_NVCC_EXTRA=-DLMV_MLP_TRACE python -m lemevit_b200.build --out=lemevit_b200/liblemevit_b200_trace.so
   LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so python tools/mlp_trace.py [R C Hd]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as G
R, Cc, Hd = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 and sys.argv[1].isdigit() else (54272, 384, 1536)))
lib = G.lib()
for a in sys.argv:
    if a.startswith("--debug="):
        lib.lmv_debug_mlp_flags(int(a.split("=")[1]))
        print("debug flags", a)
x = G.bf(torch.randn(R, Cc, device="cuda"))
W1, W2 = G.bf(torch.randn(Hd, Cc, device="cuda") * Cc ** -0.5), G.bf(torch.randn(Cc, Hd, device="cuda") * Hd ** -0.5)
b1, b2 = torch.randn(Hd, device="cuda") * 0.5, torch.randn(Cc, device="cuda")
stats = torch.stack([x.float().sum(-1), (x.float() ** 2).sum(-1)], dim=1).contiguous()
cs = W1.float().sum(-1).contiguous()
out = torch.empty_like(x)
buf = (C.c_ulonglong * (148 * 8))()
for _ in range(3):
    G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=cs, out=out)
lib.lmv_debug_mlp_trace(buf, len(buf))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=cs, out=out); e1.record(); torch.cuda.synchronize()
print(f"mlp_fused R={R} C={Cc} Hd={Hd}: {e0.elapsed_time(e1) * 1e3:.1f} us")
lib.lmv_debug_mlp_trace(buf, len(buf))
names = ["wait x_full", "wait acc1_empty", "wait w_full (fc1)", "issue fc1", "wait hid_full/acc2_empty", "wait w_full (fc2)", "issue fc2 + loop", "total"]
tot, n = [0] * 8, 0
for cta in range(0, 148, 2):
    v = [buf[cta * 8 + i] for i in range(8)]
    if v[7]:
        n += 1
        tot = [a + b for a, b in zip(tot, v)]
chunks = Hd // 128 * ((R + 255) // 256) / max(n, 1)
print(f"{n} issuers, {chunks:.1f} chunks each:", {nm: round(t / n / 1e3, 1) for nm, t in zip(names, tot)}, "kcycles;",
      {nm: round(t / n / chunks) for nm, t in zip(names, tot)}, "cycles per chunk")

names = ["wait acc1_full", "ld + math", "wait hid_empty", "sts + fence + arrive (+ loop)", "output epilogue", "", "", "total"]
tot, n = [0] * 8, 0
for cta in range(1, 148, 2):
    v = [buf[cta * 8 + i] for i in range(8)]
    if v[7]:
        n += 1
        tot = [a + b for a, b in zip(tot, v)]
print(f"epilogue warp 4 of {n} leaders:", {nm: round(t / n / chunks) for nm, t in zip(names, tot) if nm}, "cycles per chunk")
