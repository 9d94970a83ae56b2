"""Print the error margins of the whole-model golden tests (rel-max-err / cosine per output) on the current build."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G
import test_gpu_model as T
for path in sorted(glob.glob(os.path.join(T.GOLDEN, "seg_*.npz"))):
    name, B, H, W, seed = T._parse(path)
    g = np.load(path)
    cfg, sd, m = T._build(name, seed, backbone=True)
    x = T.Wt.make_input(B, H, W, seed).cuda().to(torch.bfloat16)
    outs = m(x)
    for i, o in enumerate(outs):
        if f"out{i}" in g.files:
            ref = torch.from_numpy(g[f"out{i}"]); got = o.float().cpu()
        else:
            ref = torch.from_numpy(g[f"out{i}_sample"]); got = o[:, ::8, ::4, ::4].float().cpu()
        print(os.path.basename(path), i, "rel_err %.4f cos %.6f" % (G.rel_err(got, ref), G.cosine(got, ref)))
for path in sorted(glob.glob(os.path.join(T.GOLDEN, "cls_*.npz"))):
    name, B, H, W, seed = T._parse(path)
    g = np.load(path)
    cfg, sd, m = T._build(name, seed)
    x = T.Wt.make_input(B, H, W, seed).cuda().to(torch.bfloat16)
    y = m(x).float().cpu()
    ref = torch.from_numpy(g[[k for k in g.files if k.startswith("logits") or k == "out"][0]]) if any(k.startswith("logits") or k == "out" for k in g.files) else None
    if ref is not None: print(os.path.basename(path), "rel_err %.4f" % G.rel_err(y, ref))
    else: print(os.path.basename(path), g.files)
