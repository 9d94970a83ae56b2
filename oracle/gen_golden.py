"""Generate tests/golden/*.npz from the UNMODIFIED reference — run in the build container only.

    python -m oracle.gen_golden            # needs /root/reference

The reference has no golden vectors of its own, so these fixtures are the pin:
outputs of the reference ``LeMeViT`` (models/lemevit.py, and the mmseg backbone copy) in fp32 on
CPU for seeded weights (oracle/weights.py) and seeded inputs.  Each file also stores the weight
fingerprint so that a consumer can prove it regenerated the very same weights.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import lemevit_oracle as O
from . import shims
from . import weights as Wt

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CLS_CASES = [
    # (variant, batch, H, W, seed)
    ("lemevit_micro", 2, 64, 64, 0),
    ("lemevit_micro", 1, 96, 64, 1),      # non-square, N != power of two: exercises log_N(M) scale
    ("lemevit_tiny", 2, 224, 224, 0),
    ("lemevit_small", 2, 224, 224, 0),
    ("lemevit_base", 2, 224, 224, 0),
]
SEG_CASES = [
    ("lemevit_micro", 2, 64, 64, 0),
    ("lemevit_base", 1, 256, 256, 0),
    ("lemevit_base", 1, 512, 512, 0),     # BASELINE configs[4] resolution: N = 16384 / 4096 / 1024 / 256 tokens
]
# forward_features(x, c) of the classification model (models/lemevit.py:809-829), with the model's own meta tokens and with
# caller-supplied ones
FEAT_CASES = [
    ("lemevit_micro", 2, 64, 64, 2),
    ("lemevit_tiny", 2, 224, 224, 2),
]


def custom_meta_tokens(cfg: O.OracleConfig, B: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(4242 + seed)
    return torch.randn((B, cfg.queries_len, cfg.embed_dim[0]), generator=g)


def _build_ref_cls(ref, cfg: O.OracleConfig):
    return ref.LeMeViT(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), head_dim=cfg.head_dim,
                       mlp_ratios=list(cfg.mlp_ratios), attn_type=list(cfg.attn_type),
                       queries_len=cfg.queries_len, num_classes=cfg.num_classes, in_chans=cfg.in_chans).eval()


def main():
    ref = shims.load_reference_cls()
    seg = shims.load_reference_mmseg()
    if ref is None or seg is None:
        sys.exit("reference not found under " + shims.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    for name, B, H, W, seed in CLS_CASES:
        cfg = O.VARIANTS[name]
        sd = Wt.make_state_dict(cfg, seed)
        model = _build_ref_cls(ref, cfg)
        model.load_state_dict(sd)
        x = Wt.make_input(B, H, W, seed)
        # block-level taps through forward hooks on every LeMeBlock
        taps = {}
        hooks = []
        for i, stage in enumerate(model.stages):
            for j, blk in enumerate(stage):
                hooks.append(blk.register_forward_hook(
                    lambda m, inp, out, key=f"stages.{i}.{j}.": taps.update({key + "x": out[0], key + "c": out[1]})))
        logits = model(x)
        for h in hooks:
            h.remove()
        payload = {
            "logits": logits.numpy(),
            "fingerprint": np.float64(Wt.fingerprint(sd)),
            "input_sum": np.float64(x.double().sum().item()),
        }
        if name == "lemevit_micro":
            for k, v in taps.items():
                payload["tap/" + k] = v.numpy()
        else:
            for k, v in taps.items():   # keep big variants small: per-tap statistics only
                payload["tapstat/" + k] = np.array([v.double().mean().item(), v.double().abs().mean().item(),
                                                    v.double().abs().max().item()])
        path = os.path.join(OUT, f"cls_{name}_b{B}_{H}x{W}_s{seed}.npz")
        np.savez_compressed(path, **payload)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB", "max|logit|", float(logits.abs().max()))

    for name, B, H, W, seed in FEAT_CASES:
        cfg = O.VARIANTS[name]
        sd = Wt.make_state_dict(cfg, seed)
        model = _build_ref_cls(ref, cfg)
        model.load_state_dict(sd)
        x = Wt.make_input(B, H, W, seed)
        c_own = model.meta_tokens.repeat(B, 1, 1)                 # what LeMeViT.forward passes (:833)
        c_custom = custom_meta_tokens(cfg, B, seed)
        path = os.path.join(OUT, f"feat_{name}_b{B}_{H}x{W}_s{seed}.npz")
        np.savez_compressed(path, fingerprint=np.float64(Wt.fingerprint(sd)),
                            features_own=model.forward_features(x, c_own).numpy(),
                            features_custom=model.forward_features(x, c_custom).numpy())
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")

    for name, B, H, W, seed in SEG_CASES:
        cfg = O.VARIANTS[name]
        sd = Wt.make_state_dict(cfg, seed)             # classification checkpoint → backbone: head.* unexpected
        model = seg.LeMeViT(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), head_dim=cfg.head_dim,
                            mlp_ratios=list(cfg.mlp_ratios), attn_type=list(cfg.attn_type),
                            queries_len=cfg.queries_len)
        model.train(False)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and all(k.startswith("head.") for k in unexpected), (missing, unexpected)
        x = Wt.make_input(B, H, W, seed)
        outs = model(x)
        payload = {"fingerprint": np.float64(Wt.fingerprint(sd))}
        for i, o in enumerate(outs):
            assert o.is_contiguous()
            if name == "lemevit_micro":
                payload[f"out{i}"] = o.numpy()
            else:
                payload[f"out{i}_shape"] = np.array(o.shape)
                payload[f"out{i}_stat"] = np.array([o.double().mean().item(), o.double().abs().mean().item(),
                                                    o.double().abs().max().item()])
                payload[f"out{i}_sample"] = o[:, ::8, ::4, ::4].contiguous().numpy()
        path = os.path.join(OUT, f"seg_{name}_b{B}_{H}x{W}_s{seed}.npz")
        np.savez_compressed(path, **payload)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB", [tuple(o.shape) for o in outs])


if __name__ == "__main__":
    main()
