"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, the drop-in module keeps the
reference's surface, and the multi-process plumbing of bench.py (gloo, world_size 2) shards and reduces the way the N-GPU run does."""
import os
import re
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    from lemevit_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "lemevit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lmv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _native.load()                      # builds if needed; binds every name in SIGNATURES (AttributeError otherwise)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/lemevit_b200.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in lemevit_b200/_native.py"
    assert set(_native.SIGNATURES) <= declared, "ctypes binds symbols the header does not declare"
    assert lib.lmv_version() >= 100


def test_status_codes_and_error_text_without_a_gpu():
    from lemevit_b200 import _native
    lib = _native.load()
    cfg = _native.make_config([1], [64], [256], ["Q"], 32, 16, 10, 3, False)      # unknown attention type
    assert lib.lmv_packed_tensor_count(cfg) == _native.LMV_ERR_UNSUPPORTED
    assert b"Attention type does not exit" in lib.lmv_last_error()                # the reference's message (models/lemevit.py:660)
    cfg = _native.make_config([1], [50], [200], ["S"], 32, 16, 10, 3, False)      # dim % num_heads (reference AssertionError :168)
    assert lib.lmv_packed_tensor_count(cfg) == _native.LMV_ERR_INVALID
    assert lib.lmv_posembed_layernorm(None, None, None, None, None, None, 1, 1, 1, 1, 8, 1e-6, None) == _native.LMV_ERR_INVALID
    assert lib.lmv_mlp_fused(None, None, None, None, None, None, None, None, None, 1, 1e-6, 1, 96, 384, None) == _native.LMV_ERR_INVALID


def test_module_keeps_the_reference_surface_and_refuses_cpu_tensors():
    import lemevit_b200 as L
    m = L.lemevit_tiny(num_classes=10, drop_path_rate=0.1)
    assert m.num_classes == 10 and m.embed_dim == [64, 64, 128, 192, 320] and tuple(m.meta_tokens.shape) == (16, 64)
    assert m.default_cfg["input_size"] == (3, 224, 224) and m.get_classifier() is m.head
    assert m.no_weight_decay() == {"pos_embed", "cls_token"}
    assert sum(p.numel() for p in m.parameters()) > 8e6
    m.train(False)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 224, 224))                                            # no CPU fallback, by design
    with pytest.raises(NotImplementedError):
        L.LeMeViT(attn_type=["C", "X"])


class _FakeRegistry:
    def __init__(self):
        self.modules = {}

    def register_module(self, name=None, force=False, module=None):
        self.modules[name] = module
        return module


def test_backbone_registers_with_openmmlab_style_registries(monkeypatch):
    import types
    import lemevit_b200 as L
    regs = {}
    for pkg in ("mmseg", "mmdet"):
        root, models, builder = types.ModuleType(pkg), types.ModuleType(pkg + ".models"), types.ModuleType(pkg + ".models.builder")
        regs[pkg] = builder.BACKBONES = _FakeRegistry()
        monkeypatch.setitem(sys.modules, pkg, root)
        monkeypatch.setitem(sys.modules, pkg + ".models", models)
        monkeypatch.setitem(sys.modules, pkg + ".models.builder", builder)
    assert L.register_backbones() == ["mmseg", "mmdet"]
    assert regs["mmseg"].modules["LeMeViT"] is L.LeMeViTBackbone
    kw = dict(depth=[1, 1, 1, 1, 1], embed_dim=[32, 32, 64, 96, 128], head_dim=32, queries_len=16)
    bb = regs["mmseg"].modules["LeMeViT"](frozen_stages=-1, **kw)
    assert not any(k.startswith("head.") for k in bb.state_dict())               # backbone copies have no classifier
    assert bb.train(True) is None                                                 # the reference's train() returns None (:874-882)
    assert all(m.training for m in bb.modules() if isinstance(m, torch.nn.BatchNorm2d))    # mmseg: freeze_bn = False (:875)
    # mmdet copy (object_detection/mmdet/models/backbones/lemevit.py:827-842): frozen_stages is a list of stage indices, train()
    # strips requires_grad from those stages and keeps every BatchNorm2d / LayerNorm in eval mode
    det_cls = regs["mmdet"].modules["LeMeViT"]
    assert issubclass(det_cls, L.LeMeViTBackbone) and det_cls.__name__ == "LeMeViT"
    det = det_cls(frozen_stages=[0, 1], **kw)
    assert det.train(True) is None and det.training
    assert not any(m.training for m in det.modules() if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.LayerNorm)))
    assert not any(p.requires_grad for i in (0, 1) for p in det.stages[i].parameters())
    assert all(p.requires_grad for p in det.stages[2].parameters())


def test_change_detection_entrypoints(tmp_path):
    """change_detection/models/lemevit.py:874-963 + networks.py:365-368: lemevit_small(pretrained=path) builds the backbone
    and loads the checkpoint inside the constructor ('model' / 'state_dict' / 'state_dict_ema' / bare, 'backbone.' prefix stripped)."""
    from lemevit_b200 import change_detection as CD
    import lemevit_b200 as L
    src = CD.lemevit_tiny()
    assert isinstance(src, L.LeMeViTBackbone) and src.frozen_stages == [-1]
    sd = {"backbone." + k: torch.randn_like(v) if v.is_floating_point() else v for k, v in src.state_dict().items()}
    path = str(tmp_path / "ckpt.pth")
    torch.save({"state_dict": sd}, path)
    m = CD.lemevit_tiny(pretrained=path)
    got = m.state_dict()
    assert all(torch.equal(got[k[9:]], v) for k, v in sd.items())
    assert set(CD.__all__) == {"lemevit_tiny", "lemevit_small", "lemevit_base"}


def test_training_mode_is_refused_not_silently_different():
    """The native path is the eval-mode forward: in train mode the reference uses BatchNorm batch statistics and DropPath, so
    forward() must refuse instead of returning eval numbers.  Checked before any device work."""
    import lemevit_b200 as L
    m = L.LeMeViT(depth=[1, 1, 1, 1, 1], embed_dim=[32, 32, 64, 96, 128], head_dim=32, queries_len=16)
    assert m.training
    with pytest.raises(RuntimeError, match="inference-only"):
        m._check_inference()
    m.eval()
    with torch.enable_grad(), pytest.warns(UserWarning, match="outside autograd"):
        L.LeMeViT._warned_grad = False
        m._check_inference()
    with torch.no_grad():
        m._check_inference()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator(device="cpu").manual_seed(bench.shard_seed(rank))
    x = torch.randn(4, 3, 8, 8, generator=g)                                      # this rank's shard of the synthetic batch
    ms = bench.max_over_ranks(10.0 * (rank + 1), world)                           # rank 1 is the slow one
    dist.barrier()
    out[rank] = (float(x.sum()), ms, bench.whole_job_rate(256, world, ms))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_over_ranks_timing_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] != r1[0]                              # different images per rank (no replicated work)
    assert r0[1] == r1[1] == 20.0                      # every rank reports the slowest rank's time
    assert r0[2] == r1[2] == pytest.approx(2 * 256 / 20.0 * 1e3)   # whole-job img/s = all ranks' images / max time


def test_uint8_input_layout_codes():
    """Host logic of the 8-bit input edge: the LMV_DTYPE_* code follows the memory layout of the (NCHW-shaped) tensor."""
    import torch
    from lemevit_b200 import _native
    from lemevit_b200.engine import Engine
    u8 = torch.zeros(2, 3, 8, 10, dtype=torch.uint8)
    assert Engine._x_code(u8) == _native.DTYPE_U8
    assert Engine._x_code(u8.contiguous(memory_format=torch.channels_last)) == _native.DTYPE_U8_NHWC
    assert Engine._x_code(torch.zeros(2, 8, 10, 3, dtype=torch.uint8).permute(0, 3, 1, 2)) == _native.DTYPE_U8_NHWC
    assert Engine._x_code(torch.zeros(2, 3, 8, 10, dtype=torch.bfloat16)) == _native.DTYPE_BF16
    assert Engine._x_code(torch.zeros(2, 3, 8, 10)) == _native.DTYPE_F32
