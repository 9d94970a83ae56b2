"""Minimal process for ncu: N forwards of one model/batch through the native library (no graph, no oracle).
usage: python tools/ncu_target.py lemevit_base 256 [n_forwards=2] [chunk=0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemevit_b200 as L

name = sys.argv[1] if len(sys.argv) > 1 else "lemevit_base"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
torch.manual_seed(0)
m = getattr(L, name)(native_chunk=chunk).to("cuda", torch.bfloat16)
m.train(False)
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(n):
        y = m(x)
    torch.cuda.synchronize()
print("launches per forward:", m.native_engine(x.device).launch_count(B, 224, 224))
