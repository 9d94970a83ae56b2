"""Timing probe for BASELINE config 5: LeMeViT-Base mmseg-variant backbone, 512x512, 4 feature maps.  usage: python tools/bench_backbone.py [batch=16]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemevit_b200 as L
from oracle import lemevit_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
cfg = O.VARIANTS["lemevit_base"]
m = L.LeMeViTBackbone(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), head_dim=32, mlp_ratios=[4] * 5, attn_type=list("CDDSS"), queries_len=16).to("cuda", torch.bfloat16)
m.train(False)
x = torch.randn(B, 3, 512, 512, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(2):
        outs = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        outs = m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"config": "lemevit_base backbone 512x512", "batch": B, "ms": ms, "img_s": B / ms * 1e3, "shapes": [list(o.shape) for o in outs]}))
eng = m.native_engine(x.device)
eng.set_profile(True)
with torch.no_grad():
    m(x)
print("\n".join(sorted(eng.profile_report().splitlines(), key=lambda l: -float(l.split("ms=")[1].split()[0]))[:8]))
