"""Per-phase cycle account of the self-attention kernel's softmax warps (needs the -DLMV_ATTN_TRACE build):
   LMV_NVCC_EXTRA=-DLMV_ATTN_TRACE python -m lemevit_b200.build --out=lemevit_b200/liblemevit_b200_trace.so
   LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so python tools/attn_trace.py [B heads T N]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as G
B, heads, T, N = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 12, 212, 196)))
lib = G.lib()
qkv = G.bf(torch.randn(B, T, 3 * heads * 32, device="cuda"))
buf = (C.c_ulonglong * (148 * 5 * 8))()
for rep in range(3):
    G.attention_self(qkv, heads, N, 32 ** -0.5)
    lib.lmv_debug_attn_trace(buf, len(buf))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.attention_self(qkv, heads, N, 32 ** -0.5); e1.record(); torch.cuda.synchronize()
items = B * heads * ((T + 255) // 256) * ((T + 223) // 224)
print(f"attention_self B={B} h={heads} T={T} N={N}: {e0.elapsed_time(e1) * 1e3:.1f} us, {items / 148:.1f} items per CTA")
lib.lmv_debug_attn_trace(buf, len(buf))
names = ["setup", "wait s_full", "pass1", "group sync", "wait p_empty", "pass2", "epilogue", "total"]
for role in range(5):
    tot, n = [0] * 8, 0
    for cta in range(148):
        v = [buf[(cta * 5 + role) * 8 + i] for i in range(8)]
        if v[7]:
            n += 1
            tot = [a + b for a, b in zip(tot, v)]
    if n:
        nm4 = ["wait p_full", "issue S (incl. wait ld_full)", "issue PV (incl. wait o_empty)", "", "", "", "", "", "total"]
        if role == 4:
            print("MMA issuer", {nm: round(t / n / 1e3, 1) for nm, t in zip(nm4, tot) if nm}, "kcycles per CTA")
            continue
        print(f"group {role >> 1} half {role & 1}", {nm: round(t / n / 1e3, 1) for nm, t in zip(names, tot)}, "kcycles per CTA")

if hasattr(lib, "lmv_debug_attn_events") or True:
    try:
        ev = (C.c_longlong * (6 * 16 * 8))()
        lib.lmv_debug_attn_events(ev, len(ev))
        t0 = min(v for v in ev if v > 0)
        rn = ["sm g0h0", "sm g0h1", "sm g1h0", "sm g1h1", "iss g0", "iss g1"]
        sn = [["start", "s_full", "pass1 done", "p_empty", "pass2 done", "epi done", ""], ["p_full", "kq_full", "S issued", "o_empty", "v_full", "PV issued", "PV complete"]]
        rows = []
        for role in range(6):
            for item in range(10):
                for slot in range(7):
                    v = ev[(role * 16 + item) * 8 + slot]
                    if v > 0:
                        rows.append((v - t0, rn[role], item, sn[role >= 4][slot]))
        rows.sort()
        print("timeline of CTA 0 (cycles):")
        for t, r, i, nm in rows:
            if r in ("sm g0h1", "sm g1h1"):
                continue
            print(f"{t:8d}  {r:8s} item {i:2d}  {nm}")
    except AttributeError:
        pass
