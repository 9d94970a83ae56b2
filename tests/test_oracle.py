"""The oracle (oracle/lemevit_oracle.py) against the reference: committed golden fixtures produced by
the unmodified reference (oracle/gen_golden.py) and, when /root/reference is mounted, the live reference."""
import glob
import os
import re

import numpy as np
import pytest
import torch

from oracle import lemevit_oracle as O
from oracle import shims
from oracle import weights as Wt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
torch.set_grad_enabled(False)


def _parse(path):
    m = re.match(r"(cls|seg)_(lemevit_\w+?)_b(\d+)_(\d+)x(\d+)_s(\d+)\.npz", os.path.basename(path))
    kind, name, B, H, W, seed = m.groups()
    return kind, name, int(B), int(H), int(W), int(seed)


CLS = sorted(glob.glob(os.path.join(GOLDEN, "cls_*.npz")))
SEG = sorted(glob.glob(os.path.join(GOLDEN, "seg_*.npz")))


def test_fixtures_exist():
    assert len(CLS) >= 5 and len(SEG) >= 2


@pytest.mark.parametrize("path", CLS, ids=os.path.basename)
def test_cls_matches_reference_golden(path):
    _, name, B, H, W, seed = _parse(path)
    g = np.load(path)
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, seed)
    assert Wt.fingerprint(sd) == pytest.approx(float(g["fingerprint"]), rel=1e-12), "weights not reproduced"
    x = Wt.make_input(B, H, W, seed)
    assert float(x.double().sum()) == pytest.approx(float(g["input_sum"]), rel=1e-9, abs=1e-6)
    taps = {}
    y = O.forward_cls(sd, cfg, x, taps=taps).numpy()
    ref = g["logits"]
    # fp32 vs fp32 with a different op order: tolerance 2e-4 of the logit range
    assert np.abs(y - ref).max() <= 2e-4 * np.abs(ref).max()
    for key in g.files:
        if key.startswith("tap/"):
            t = taps[key[4:]].numpy()
            assert np.abs(t - g[key]).max() <= 2e-4 * max(1.0, np.abs(g[key]).max()), key
        elif key.startswith("tapstat/"):
            t = taps[key[8:]].double()
            stat = np.array([t.mean().item(), t.abs().mean().item(), t.abs().max().item()])
            np.testing.assert_allclose(stat, g[key], rtol=2e-4, atol=1e-5, err_msg=key)


@pytest.mark.parametrize("path", SEG, ids=os.path.basename)
def test_backbone_matches_reference_golden(path):
    _, name, B, H, W, seed = _parse(path)
    g = np.load(path)
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, seed)
    assert Wt.fingerprint(sd) == pytest.approx(float(g["fingerprint"]), rel=1e-12)
    outs = O.forward_backbone(sd, cfg, Wt.make_input(B, H, W, seed))
    assert len(outs) == 4
    for i, o in enumerate(outs):
        if f"out{i}" in g.files:
            assert np.abs(o.numpy() - g[f"out{i}"]).max() <= 2e-4 * np.abs(g[f"out{i}"]).max()
        else:
            assert tuple(o.shape) == tuple(g[f"out{i}_shape"])
            s = o[:, ::8, ::4, ::4].numpy()
            assert np.abs(s - g[f"out{i}_sample"]).max() <= 2e-4 * np.abs(g[f"out{i}_sample"]).max()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "feat_*.npz"))), ids=os.path.basename)
def test_oracle_forward_features_against_reference_golden(path):
    """forward_features(x, c) of the classification model (models/lemevit.py:809-829): own and caller-supplied meta tokens."""
    from oracle.gen_golden import custom_meta_tokens
    name, B, H, W, seed = re.match(r"feat_(lemevit_\w+?)_b(\d+)_(\d+)x(\d+)_s(\d+)\.npz", os.path.basename(path)).groups()
    B, H, W, seed = int(B), int(H), int(W), int(seed)
    g = np.load(path)
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, seed)
    assert Wt.fingerprint(sd) == pytest.approx(float(g["fingerprint"]), rel=1e-12)
    x = Wt.make_input(B, H, W, seed)
    own = O.forward_features_cls(sd, cfg, x)
    custom = O.forward_features_cls(sd, cfg, x, custom_meta_tokens(cfg, B, seed))
    assert np.abs(own.numpy() - g["features_own"]).max() <= 2e-4 * np.abs(g["features_own"]).max()
    assert np.abs(custom.numpy() - g["features_custom"]).max() <= 2e-4 * np.abs(g["features_custom"]).max()
    assert np.abs(g["features_custom"] - g["features_own"]).max() > 1e-2      # the meta tokens do reach the features


def test_flop_accounting_matches_survey():
    # 3.892 / 7.912 / 23.481 GFLOP per image at 224^2, 138.468 for base@512 backbone
    f = lambda n, s, b=False: O.algorithmic_flops_per_image(O.VARIANTS[n], s, s, backbone=b) / 1e9
    assert f("lemevit_tiny", 224) == pytest.approx(3.892, abs=2e-3)
    assert f("lemevit_small", 224) == pytest.approx(7.912, abs=2e-3)
    assert f("lemevit_base", 224) == pytest.approx(23.481, abs=2e-3)
    assert f("lemevit_base", 512, True) == pytest.approx(138.468, abs=2e-3)


def test_fp64_oracle_agrees_with_fp32():
    cfg = O.VARIANTS["lemevit_micro"]
    sd = Wt.make_state_dict(cfg, 3)
    x = Wt.make_input(1, 64, 64, 3)
    y32 = O.forward_cls(sd, cfg, x)
    y64 = O.forward_cls(O.cast_state_dict(sd, torch.float64), cfg, x.double())
    assert (y32.double() - y64).abs().max() < 1e-4


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("name", ["lemevit_micro", "lemevit_tiny"])
def test_oracle_against_live_reference(name):
    ref = shims.load_reference_cls()
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, 5)
    m = ref.LeMeViT(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), head_dim=cfg.head_dim,
                    mlp_ratios=list(cfg.mlp_ratios), attn_type=list(cfg.attn_type),
                    queries_len=cfg.queries_len).eval()
    m.load_state_dict(sd)
    x = Wt.make_input(2, 96, 96, 5)
    assert (m(x) - O.forward_cls(sd, cfg, x)).abs().max() < 1e-4


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_state_dict_spec_is_the_reference_schema():
    ref = shims.load_reference_cls()
    for name in ("lemevit_tiny", "lemevit_small", "lemevit_base"):
        sd_ref = getattr(ref, name)().state_dict()
        spec = Wt.state_dict_spec(O.VARIANTS[name])
        assert list(spec.keys()) == list(sd_ref.keys())
        for k, (shape, _) in spec.items():
            assert tuple(sd_ref[k].shape) == tuple(shape), k
