"""Per-role cycle account of the fused cross-attention kernel (needs the -DLMV_DCA_TRACE build):
   LMV_NVCC_EXTRA=-DLMV_DCA_TRACE python -m lemevit_b200.build --out=lemevit_b200/liblemevit_b200_trace.so
   LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so python tools/dca_trace.py [B N C heads kind]"""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as G
from tests.test_gpu_kernels import _dca_weights
B, N, Cc, heads = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 3136, 96, 3)))
kind = sys.argv[5] if len(sys.argv) > 5 else "D"
flags = int(sys.argv[6]) if len(sys.argv) > 6 else 0
lib = G.lib()
xt = G.bf(torch.randn(B, N, Cc, device="cuda"))
c = G.bf(torch.randn(B, 16, Cc, device="cuda"))
W = _dca_weights(kind, Cc, 4 * Cc)
sx, sc = math.log(16) / math.log(N) * Cc ** -0.5, Cc ** -0.5
buf = (C.c_ulonglong * (148 * 4 * 8))()
for rep in range(3):
    G.dca_block(kind, xt, c, W, heads, sx, sc, flags=flags)
    lib.lmv_debug_dca_trace(buf, len(buf))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.dca_block(kind, xt, c, W, heads, sx, sc, flags=flags); e1.record(); torch.cuda.synchronize()
print(f"dca_block {kind} B={B} N={N} C={Cc}: {e0.elapsed_time(e1) * 1e3:.1f} us (pre + dca_x + post)")
lib.lmv_debug_dca_trace(buf, len(buf))
names = {1: ["loop", "poll", "issue S", "issue Sc", "issue dx", "issue Z", "", "total"], 2: ["stats+consts", "wait s_full", "softmax", "wait dx_full", "epilogue", "", "", "total"],
         3: ["stats+sync", "wait sc_full", "pass1", "wait z_done", "rescale+pass2", "flush", "", "total"]}
for role in (1, 2, 3):
    tot = [0] * 8
    n = 0
    for cta in range(148):
        v = [buf[(cta * 4 + role) * 8 + i] for i in range(8)]
        if v[7]:
            n += 1
            tot = [a + b for a, b in zip(tot, v)]
    if n:
        print(["MMA", "MMA", "x-group", "c-group"][role], {nm: round(t / n / 1e3, 1) for nm, t in zip(names[role], tot) if nm}, "kcycles per CTA")
