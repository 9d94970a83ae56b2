"""pack.py (weight folding + ordering) against the oracle, on CPU: the packed tensors are kept in float64
(no bf16 rounding) and pushed through tests/packed_emulator.py, which mirrors the native schedule."""
import ctypes as C

import pytest
import torch

import lemevit_b200.pack as pack_mod
from lemevit_b200 import _native
from oracle import lemevit_oracle as O
from oracle import weights as Wt
from tests import packed_emulator as E

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name,H,W", [("lemevit_micro", 64, 64), ("lemevit_micro", 96, 64), ("lemevit_tiny", 64, 64)])
@pytest.mark.parametrize("backbone", [False, True])
def test_packed_forward_matches_oracle(name, H, W, backbone, monkeypatch):
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, 2)
    x = Wt.make_input(2, H, W, 2)
    kw = dict(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), attn_type=list(cfg.attn_type), in_chans=cfg.in_chans,
              num_classes=cfg.num_classes, backbone=backbone)
    # keep float64 through packing so that the comparison isolates the fold algebra from bf16 rounding
    real_to = torch.Tensor.to

    def keep64(self, *a, **k):
        if a and a[0] in (torch.bfloat16, torch.float32):
            return self
        return real_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", keep64)
    packed = pack_mod.pack_state_dict(sd, device="cpu", **kw)
    monkeypatch.undo()
    y = E.forward(packed, x, head_dim=cfg.head_dim, queries_len=cfg.queries_len, **kw)
    sd64 = O.cast_state_dict(sd, torch.float64)
    if backbone:
        ref = O.forward_backbone(sd64, cfg, x.double())
        assert len(y) == 4
        for a, b in zip(y, ref):
            assert a.shape == b.shape
            assert (a - b).abs().max() < 1e-9 * max(1.0, b.abs().max())
    else:
        ref = O.forward_cls(sd64, cfg, x.double())
        assert (y - ref).abs().max() < 1e-9 * max(1.0, ref.abs().max())
        # forward_features(x, c) (models/lemevit.py:809-829): pre-head features, with the model's own and with caller-supplied
        # meta tokens (the latter run meta_token_downsample[0] at run time instead of the pack-time constant)
        f = E.forward(packed, x, head_dim=cfg.head_dim, queries_len=cfg.queries_len, features=True, **kw)
        ref_f = O.forward_features_cls(sd64, cfg, x.double())
        assert (f - ref_f).abs().max() < 1e-9 * max(1.0, ref_f.abs().max())
        c = torch.randn(x.shape[0], cfg.queries_len, cfg.embed_dim[0], dtype=torch.float64, generator=torch.Generator().manual_seed(5))
        f = E.forward(packed, x, head_dim=cfg.head_dim, queries_len=cfg.queries_len, features=True, c_in=c, **kw)
        ref_f = O.forward_features_cls(sd64, cfg, x.double(), c)
        assert (f - ref_f).abs().max() < 1e-9 * max(1.0, ref_f.abs().max())


@pytest.mark.parametrize("name", ["lemevit_micro", "lemevit_tiny", "lemevit_small", "lemevit_base"])
@pytest.mark.parametrize("backbone", [False, True])
def test_packed_count_matches_native_walker(name, backbone):
    cfg = O.VARIANTS[name]
    sd = Wt.make_state_dict(cfg, 0)
    packed = pack_mod.pack_state_dict(sd, depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), attn_type=list(cfg.attn_type),
                                      in_chans=3, num_classes=1000, backbone=backbone, device="cpu")
    lib = _native.load()
    ncfg = _native.make_config(cfg.depth, cfg.embed_dim, [int(r * d) for r, d in zip(cfg.mlp_ratios, cfg.embed_dim)],
                               list(cfg.attn_type), 32, 16, 1000, 3, backbone)
    assert lib.lmv_packed_tensor_count(C.byref(ncfg)) == len(packed)
    # the plan validates numel / dtype / alignment entry by entry — host pointers are fine for that check
    arr = (_native.Tensor * len(packed))()
    for i, t in enumerate(packed):
        arr[i].data, arr[i].numel = t.data_ptr(), t.numel()
        arr[i].dtype = _native.DTYPE_BF16 if t.dtype == torch.bfloat16 else _native.DTYPE_F32
    plan = C.c_void_p()
    assert lib.lmv_plan_create(C.byref(ncfg), arr, len(packed), C.byref(plan)) == 0, lib.lmv_last_error()
    assert lib.lmv_workspace_bytes(plan, 4, 224, 224) > 0
    lib.lmv_plan_destroy(plan)
    # wrong count / wrong size are rejected loudly
    assert lib.lmv_plan_create(C.byref(ncfg), arr, len(packed) - 1, C.byref(plan)) == _native.LMV_ERR_INVALID
    arr[3].numel += 1
    assert lib.lmv_plan_create(C.byref(ncfg), arr, len(packed), C.byref(plan)) == _native.LMV_ERR_INVALID
    assert b"entry 3" in lib.lmv_last_error()
