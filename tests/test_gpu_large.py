"""Large-resolution parity: the mmdet backbone runs at 1024 x 1024, where stages 0-1 hold N = 65 536 image
tokens per image — 512 tiles per image in the fused cross-attention kernel (86 softmax segments merged per meta-token row) and a
256-wide map in the positional-embedding kernel.  Micro variant (the oracle finishes in seconds), backbone mode, against the oracle."""
import pytest
import torch

import lemevit_b200 as L
from oracle import lemevit_oracle as O
from oracle import weights as Wt
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.mark.parametrize("H,W", [(1024, 1024), (608, 800)])
def test_backbone_micro_at_detection_resolution(H, W):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = O.VARIANTS["lemevit_micro"]
    sd = Wt.make_state_dict(cfg, 9)
    m = L.LeMeViTBackbone(depth=list(cfg.depth), embed_dim=list(cfg.embed_dim), head_dim=cfg.head_dim, mlp_ratios=list(cfg.mlp_ratios),
                          attn_type=list(cfg.attn_type), queries_len=cfg.queries_len, frozen_stages=[0, 1, 2, 3, 4], norm_eval=True)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing
    m = m.to("cuda", torch.bfloat16)
    m.train(True)                                 # mmdet semantics: frozen stages + norms in eval mode -> the eval forward
    x = Wt.make_input(1, H, W, 9)
    ref = O.forward_backbone(sd, cfg, x)
    outs = m(x.cuda().to(torch.bfloat16))
    assert [tuple(o.shape) for o in outs] == [tuple(r.shape) for r in ref]
    for i, (o, r) in enumerate(zip(outs, ref)):
        got = o.float().cpu()
        rms = float(((got - r).double().pow(2).mean() / r.double().pow(2).mean()).sqrt())
        print(f"{H}x{W} out{i} {tuple(r.shape)}: rel-max-err {G.rel_err(got, r):.4f} rel-rms-err {rms:.4f} cosine {G.cosine(got, r):.6f}")
        assert rms <= 2e-2 and G.cosine(got, r) > 0.9995 and G.rel_err(got, r) <= 6e-2
