"""Per-kernel parity on the GPU (B200): each native kernel, called through the C ABI, against a plain
PyTorch fp32 restatement of the same op on the same bf16 inputs.  Tolerances:
per-kernel max-abs-err / max|ref| <= 1e-2 and cosine >= 0.9999 — bf16 outputs carry 2^-9 relative rounding."""
import math

import pytest
import torch

from tests import gpu_util as G

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.manual_seed(0)
    yield
    torch.cuda.synchronize()


def _rand(*shape, scale=1.0):
    return G.bf(torch.randn(*shape, device="cuda") * scale)


# ---- GEMM --------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    # (M, N, K) — every (N, K) family the three variants use, plus ragged edges
    (128, 32, 64), (128, 128, 64), (256, 256, 128), (300, 96, 96), (1000, 288, 96), (777, 192, 192),
    (512, 384, 96), (640, 576, 192), (200, 768, 192), (424, 1152, 384), (424, 1536, 384), (424, 384, 1536),
    (98, 2048, 512), (98, 512, 2048), (4096, 48, 32), (3136, 96, 432), (784, 192, 864), (196, 384, 1728),
    (49, 512, 3456), (32, 1000, 512), (7, 51, 320), (130, 320, 1280), (130, 960, 320), (1, 64, 8),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_linear_tcgen05_matches_fp32(M, N, K):
    A, W = _rand(M, K), _rand(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    ref = G.ref_linear(A, W, bias)
    out = G.linear(A, W, bias, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert G.rel_err(out, ref) < 2e-3, G.describe_mismatch(out, ref, 2e-3)     # fp32 out: only accumulation-order error
    out16 = G.linear(A, W, bias)
    assert G.rel_err(out16, ref) < TOL and G.cosine(out16, ref) > 0.9999


@pytest.mark.parametrize("force_bn", [32, 64, 96, 128, 160, 192, 224, 256])
def test_linear_every_tile_width(force_bn):
    M, N, K = 384, 512, 256
    A, W = _rand(M, K), _rand(N, K, scale=K ** -0.5)
    ref = G.ref_linear(A, W)
    out = G.linear(A, W, out_dtype=torch.float32, force_bn=force_bn)
    assert G.rel_err(out, ref) < 2e-3, G.describe_mismatch(out, ref, 2e-3)


def test_linear_epilogue_gelu_residual_inplace():
    M, N, K = 1000, 384, 1536
    A, W = _rand(M, K), _rand(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = _rand(M, N)
    ref = G.ref_linear(A, W, bias, residual=res)
    out = res.clone()
    G.linear(A, W, bias, residual=out, out=out)                      # residual stream updated in place
    assert G.rel_err(out, ref) < TOL
    ref_g = G.ref_linear(A, W, bias, gelu=True)
    out_g = G.linear(A, W, bias, gelu=True)
    assert G.rel_err(out_g, ref_g) < TOL and G.cosine(out_g, ref_g) > 0.9999


@pytest.mark.parametrize("M,N,K,gelu", [(1000, 288, 96, False), (777, 384, 96, True), (424, 1152, 384, False), (424, 1536, 384, True),
                                        (300, 768, 192, True), (130, 960, 320, False), (98, 2048, 512, True), (64, 51, 64, False)])
@pytest.mark.parametrize("simt", [False, True])
def test_linear_layernorm_fold(M, N, K, gelu, simt):
    """LayerNorm (no affine) -> Linear (+GELU), the norm folded into the GEMM epilogue from per-row (sum, sum^2)."""
    y = G.bf(torch.randn(M, K, device="cuda") * 1.5 + torch.randn(M, 1, device="cuda") * 2.0)   # rows with large means
    W = _rand(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    stats = torch.stack([y.float().sum(-1), (y.float() ** 2).sum(-1)], dim=1).contiguous()
    if M % 2 == 0:   # the same statistics split into 3 partials per row (what a multi-tile producer GEMM emits)
        stats = torch.stack([stats * 0.5, stats * 0.25, stats * 0.25], dim=1).contiguous()
    colsum = W.float().sum(-1).contiguous()
    ref = G.ref_linear(torch.nn.functional.layer_norm(y.float(), (K,), eps=1e-6), W, bias, gelu=gelu)
    out = G.linear_fused(y, W, bias, gelu=gelu, ln_stats=stats, ln_colsum=colsum, ln_eps=1e-6, simt=simt)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float(), ref, TOL)


@pytest.mark.parametrize("M,N,K,gelu", [(1000, 288, 96, False), (424, 1536, 384, True), (300, 576, 192, False)])
@pytest.mark.parametrize("ratio", [30.0, 100.0, 300.0])
def test_linear_layernorm_fold_large_row_offsets(M, N, K, gelu, ratio):
    """Trained ViTs carry massive-activation channels / large per-row offsets: rows whose |mean| is `ratio` x their std.  The fold
    computes var = E[y^2] - mu^2 and r (acc - mu colsum) in fp32 from the bf16 rows (gemm.cu), both cancellation-prone; checked
    against fp64 LayerNorm of the same bf16-rounded rows (the information the reference's bf16 path sees, too)."""
    std = 0.7
    sign = torch.where(torch.rand(M, 1, device="cuda") < 0.5, -1.0, 1.0)
    y = G.bf(torch.randn(M, K, device="cuda") * std + sign * ratio * std)
    yd = y.double()
    got_ratio = float((yd.mean(-1).abs() / yd.std(-1)).median())
    assert got_ratio > 0.8 * ratio
    W = _rand(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    stats = torch.stack([y.float().sum(-1), (y.float() ** 2).sum(-1)], dim=1).contiguous()
    colsum = W.float().sum(-1).contiguous()
    ln = torch.nn.functional.layer_norm(yd, (K,), eps=1e-6)
    ref = ln @ W.double().t() + bias.double()
    if gelu:
        ref = 0.5 * ref * (1.0 + torch.erf(ref / math.sqrt(2.0)))
    out = G.linear_fused(y, W, bias, gelu=gelu, ln_stats=stats, ln_colsum=colsum, ln_eps=1e-6)
    err = G.rel_err(out, ref)
    print(f"mean/std {got_ratio:.0f}: rel err {err:.5f}")
    assert err < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float(), ref.float(), TOL)


def _dca_weights(kind, C, Hd, gain=1.5):
    W = {}
    lin = lambda o, i, g=1.0: _rand(o, i, scale=g * i ** -0.5)
    if kind == "D":
        W["wa"], W["wb"], W["wp1"], W["wp2"] = lin(3 * C, C, gain), lin(3 * C, C, gain), lin(C, C, 0.6), lin(C, C, 0.6)
        W["ba"], W["bb"] = torch.randn(3 * C, device="cuda") * 0.2, torch.randn(3 * C, device="cuda") * 0.2
        W["bp2"] = torch.randn(C, device="cuda") * 0.1
    else:
        W["wa"], W["wb"], W["wp1"] = lin(C, C, gain), lin(2 * C, C, gain), lin(C, C, 0.6)
        W["ba"], W["bb"] = torch.randn(C, device="cuda") * 0.2, torch.randn(2 * C, device="cuda") * 0.2
    W["bp1"] = torch.randn(C, device="cuda") * 0.1
    W["w1"], W["w2"] = lin(Hd, C, 0.8), lin(C, Hd, 0.8)
    W["b1"], W["b2"] = torch.randn(Hd, device="cuda") * 0.2, torch.randn(C, device="cuda") * 0.1
    return W


DCA_SHAPES = [(3, 300, 96, 3), (2, 3136, 96, 3), (2, 784, 192, 6), (5, 784, 192, 6), (2, 200, 64, 2), (2, 1000, 128, 4), (3, 128, 32, 1),
              (2, 4096, 160, 5), (40, 3136, 96, 3), (160, 784, 192, 6)]


@pytest.mark.parametrize("B,N,C,heads", DCA_SHAPES)
@pytest.mark.parametrize("kind", ["D", "C"])
def test_dca_block_fused(kind, B, N, C, heads):
    """Fused cross-attention block core (meta_pre -> dca_x -> meta_post, absorbed projections) against the reference's op order in
    fp32: DualCrossAttention (models/lemevit.py:252-302) with both residual updates and the meta-token MLP (:561-564), and the
    CrossAttention variant (:477-486, :600-601).  Partial last tiles (N % 128 != 0), duplicated-row (R <= 64) and full-row
    layouts, several segments per image, several segments per CTA."""
    torch.manual_seed(B * 7 + N + C)
    xt = G.bf(torch.randn(B, N, C, device="cuda") * 1.3 + torch.randn(B, N, 1, device="cuda") * 1.5)
    c = G.bf(torch.randn(B, 16, C, device="cuda"))
    W = _dca_weights(kind, C, 4 * C)
    scale_c = C ** -0.5 if kind == "D" else 32 ** -0.5
    scale_x = math.log(16) / math.log(N) * C ** -0.5
    ref_x, ref_c = G.ref_dca_block(kind, xt, c, W, heads, scale_x, scale_c)
    xout, stats2, c_out, _ = G.dca_block(kind, xt, c, W, heads, scale_x, scale_c, stats_parts=2 if B % 2 else 1)
    torch.cuda.synchronize()
    ec = G.rel_err(c_out, ref_c)
    assert ec < TOL and G.cosine(c_out, ref_c) > 0.9999, f"c: rel err {ec}\n" + G.describe_mismatch(c_out.float().reshape(-1, C), ref_c.reshape(-1, C), TOL)
    if kind == "D":
        ex = G.rel_err(xout, ref_x)
        assert ex < TOL and G.cosine(xout, ref_x) > 0.9999, f"x: rel err {ex}\n" + G.describe_mismatch(xout.float().reshape(-1, C), ref_x.reshape(-1, C), TOL)
        st = torch.stack([xout.float().sum(-1), (xout.float() ** 2).sum(-1)], dim=-1).reshape(-1, 2)
        assert torch.allclose(stats2, st, rtol=2e-3, atol=2e-2)
        # the x-branch update must be visible above the rounding floor (otherwise the check above is vacuous)
        assert (ref_x - xt.float()).abs().max() > 0.05 * ref_x.abs().max()
    assert (ref_c - c.float()).abs().max() > 0.05 * ref_c.abs().max()
    # deterministic, and bit-identical with the per-64-channel issue of the c-branch accumulation
    x2, _, c2, _ = G.dca_block(kind, xt, c, W, heads, scale_x, scale_c, stats_parts=2 if B % 2 else 1)
    assert torch.equal(c2, c_out) and (kind == "C" or torch.equal(x2, xout))
    x3, _, c3, _ = G.dca_block(kind, xt, c, W, heads, scale_x, scale_c, flags=1, stats_parts=2 if B % 2 else 1)
    assert G.rel_err(c3, ref_c) < TOL
    # the pipelined schedule (C <= 96, 'C' blocks) and the one-tile-at-a-time schedule compute the same bits
    x4, _, c4, _ = G.dca_block(kind, xt, c, W, heads, scale_x, scale_c, flags=2, stats_parts=2 if B % 2 else 1)
    assert torch.equal(c4, c_out) and (kind == "C" or torch.equal(x4, xout))


def test_dca_block_batch_invariance_and_peaked_softmax():
    """An image's result does not depend on the batch around it (fixed segments, fixed merge order), and a c-branch softmax that is
    nearly one-hot with a late maximum (forces the lazy rescale of the TMEM accumulator) stays exact."""
    B, N, C, heads = 9, 3136, 96, 3
    torch.manual_seed(5)
    xt = G.bf(torch.randn(B, N, C, device="cuda"))
    xt[:, 3000:3010] *= 6.0                       # a few late tokens with large norms -> large scores late in the image
    c = G.bf(torch.randn(B, 16, C, device="cuda"))
    W = _dca_weights("D", C, 4 * C, gain=4.0)
    sc, sx = C ** -0.5, math.log(16) / math.log(N) * C ** -0.5
    ref_x, ref_c = G.ref_dca_block("D", xt, c, W, heads, sx, sc)
    xout, _, c_out, _ = G.dca_block("D", xt, c, W, heads, sx, sc)
    # logits this large (weights with 4x the usual gain) amplify the bf16 rounding of the query / key operands: 2e-2 here
    assert G.rel_err(c_out, ref_c) < 2e-2 and G.rel_err(xout, ref_x) < 2e-2
    x1, _, c1, _ = G.dca_block("D", xt[4:5].contiguous(), c[4:5].contiguous(), W, heads, sx, sc)
    assert torch.equal(x1[0], xout[4]) and torch.equal(c1[0], c_out[4])


@pytest.mark.parametrize("kind,B,N,C,heads", [("D", 5, 300, 96, 3), ("D", 3, 784, 192, 6), ("C", 7, 1000, 96, 3), ("D", 2, 200, 64, 2), ("D", 1, 256, 160, 5)])
def test_dca_block_images_per_cta(kind, B, N, C, heads, monkeypatch):
    """The meta-token kernels run 1, 2 or 4 images per CTA (one weight fragment feeds the MMAs of all of them); the per-image
    arithmetic is the same sequence, so the results are bit-identical, odd batches (last CTA runs its image twice) included."""
    torch.manual_seed(11 * B + C)
    xt = G.bf(torch.randn(B, N, C, device="cuda") * 1.3)
    c = G.bf(torch.randn(B, 16, C, device="cuda"))
    W = _dca_weights(kind, C, 4 * C)
    sc, sx = C ** -0.5, math.log(16) / math.log(N) * C ** -0.5
    ref_x, ref_c = G.ref_dca_block(kind, xt, c, W, heads, sx, sc)
    outs = []
    for im in (1, 2, 4):
        monkeypatch.setenv("LMV_META_IM", str(im))
        xout, _, c_out, _ = G.dca_block(kind, xt, c, W, heads, sx, sc)
        torch.cuda.synchronize()
        assert G.rel_err(c_out, ref_c) < TOL
        outs.append((xout, c_out))
    for xo, co in outs[1:]:
        assert torch.equal(co, outs[0][1]) and (kind == "C" or torch.equal(xo, outs[0][0]))


MLP_SHAPES = [(1000, 96, 384), (777, 192, 768), (300, 64, 256), (260, 160, 640), (130, 128, 512), (129, 192, 1280),
              (40000, 96, 384), (25000, 192, 768),
              # wide variant (256 < C <= 384: 64-column hidden chunks, fc2 in two halves, 3-D W1' boxes) — the stage-3 'S' blocks
              (1000, 384, 1536), (130, 384, 512), (54272, 384, 1536), (333, 384, 128)]


@pytest.mark.parametrize("R,C,Hd", MLP_SHAPES)
@pytest.mark.parametrize("ln", [True, False])
def test_mlp_fused(R, C, Hd, ln):
    """x + W2 gelu(W1 LN(x) + b1) + b2 in one kernel: LayerNorm folded from row statistics, hidden kept on chip."""
    x = G.bf(torch.randn(R, C, device="cuda") * 1.5 + (torch.randn(R, 1, device="cuda") * 2.0 if ln else 0.0))
    W1, W2 = _rand(Hd, C, scale=C ** -0.5), _rand(C, Hd, scale=Hd ** -0.5)
    b1, b2 = torch.randn(Hd, device="cuda") * 0.5, torch.randn(C, device="cuda")
    xin = torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-6) if ln else x.float()
    h = xin @ W1.float().t() + b1
    h = 0.5 * h * (1.0 + torch.erf(h / math.sqrt(2.0)))
    ref = x.float() + h @ W2.float().t() + b2
    stats = colsum = None
    if ln:
        stats = torch.stack([x.float().sum(-1), (x.float() ** 2).sum(-1)], dim=1).contiguous()
        if R % 2 == 0:
            stats = torch.stack([stats * 0.5, stats * 0.25, stats * 0.25], dim=1).contiguous()
        colsum = W1.float().sum(-1).contiguous()
    out = G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=colsum)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float(), ref, TOL)
    # in place on x (the way the forward uses it) gives the identical result, twice (deterministic)
    x2 = x.clone()
    G.mlp_fused(x2, W1, b1, W2, b2, ln_stats=stats, colsum1=colsum, out=x2)
    assert torch.equal(x2, out)


@pytest.mark.parametrize("R,C,Hd", [(1000, 384, 1536), (130, 384, 512), (54272, 384, 1536), (257, 384, 128),
                                    (1000, 320, 1280), (33280, 320, 1280), (129, 320, 256)])
def test_mlp_fused_cta_pair_kernel(R, C, Hd, monkeypatch):
    """The cta_group::2 variant of the wide MLP (two CTAs of a cluster share M = 256 MMAs, each loading half of every weight box):
    checked against the fp32 reference and (C = 384) against the single-CTA wide kernel.  C = 320 (Small stage 3, Tiny stage 4) only
    exists as a pair kernel."""
    x = G.bf(torch.randn(R, C, device="cuda") * 1.5 + torch.randn(R, 1, device="cuda") * 2.0)
    W1, W2 = _rand(Hd, C, scale=C ** -0.5), _rand(C, Hd, scale=Hd ** -0.5)
    b1, b2 = torch.randn(Hd, device="cuda") * 0.5, torch.randn(C, device="cuda")
    stats = torch.stack([x.float().sum(-1), (x.float() ** 2).sum(-1)], dim=1).contiguous()
    colsum = W1.float().sum(-1).contiguous()
    xin = torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-6)
    h = xin @ W1.float().t() + b1
    h = 0.5 * h * (1.0 + torch.erf(h / math.sqrt(2.0)))
    ref = x.float() + h @ W2.float().t() + b2
    monkeypatch.setenv("LMV_MLP_PAIR", "1")
    pair = G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=colsum)
    assert G.rel_err(pair, ref) < TOL and G.cosine(pair, ref) > 0.9999, G.describe_mismatch(pair.float(), ref, TOL)
    if C == 384:
        monkeypatch.setenv("LMV_MLP_PAIR", "0")
        single = G.mlp_fused(x, W1, b1, W2, b2, ln_stats=stats, colsum1=colsum)
        monkeypatch.setenv("LMV_MLP_PAIR", "1")
        assert G.rel_err(pair.float(), single.float()) < 2e-3
    x2 = x.clone()
    G.mlp_fused(x2, W1, b1, W2, b2, ln_stats=stats, colsum1=colsum, out=x2)
    assert torch.equal(x2, pair)


def test_mlp_fused_rejects_unsupported_shapes():
    x = _rand(64, 512)
    W1, W2 = _rand(2048, 512), _rand(512, 2048)
    b1, b2 = torch.zeros(2048, device="cuda"), torch.zeros(512, device="cuda")
    with pytest.raises(RuntimeError):
        G.mlp_fused(x, W1, b1, W2, b2)


@pytest.mark.parametrize("M,N,K", [(1000, 96, 96), (777, 192, 192), (424, 384, 384), (130, 320, 320), (98, 512, 512), (50, 32, 64)])
@pytest.mark.parametrize("simt", [False, True])
def test_linear_residual_row_statistics(M, N, K, simt):
    """x += A W^T + b in place, and stats_out += (sum_n x, sum_n x^2) per row (input of the next LayerNorm fold)."""
    A, W = _rand(M, K), _rand(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    x = _rand(M, N)
    ref = G.ref_linear(A, W, bias, residual=x)
    parts = torch.full((M, G.stats_parts(N, simt), 2), float("nan"), device="cuda")
    out = G.linear_fused(A, W, bias, residual=x, out=x, stats_out=parts, simt=simt)
    assert out.data_ptr() == x.data_ptr() and G.rel_err(out, ref) < TOL
    stats = parts.sum(1)     # the statistics describe the STORED bf16 rows exactly (up to fp32 summation order)
    o = out.float()
    assert torch.allclose(stats[:, 0], o.sum(-1), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[:, 1], (o * o).sum(-1), rtol=1e-4, atol=1e-3)
    # and the run is deterministic
    x2 = x.clone()
    parts2 = torch.empty_like(parts)
    G.linear_fused(A, W, bias, residual=None, out=x2, stats_out=parts2, simt=simt)
    G.linear_fused(A, W, bias, residual=None, out=x, stats_out=parts, simt=simt)
    assert torch.equal(x, x2) and torch.equal(parts, parts2)


def test_gelu_epilogue_close_to_exact_erf():
    """The tcgen05 epilogue uses a tanh-form GELU fitted to the exact-erf GELU (max abs deviation 2.5e-5 + MUFU.TANH
    2^-11); the SIMT cross-check kernel uses erff.  fp32 outputs isolate the activation error."""
    M, N, K = 256, 256, 64
    A = G.bf(torch.linspace(-6, 6, M * K, device="cuda").reshape(M, K))
    W = G.bf(torch.eye(N, K, device="cuda"))
    out = G.linear(A, W, gelu=True, out_dtype=torch.float32)
    ref = G.ref_linear(A, W, gelu=True)
    assert float((out - ref).abs().max()) < 2e-3


def test_linear_many_tiles_persistent_loop():
    # > 148 * 2 tiles: exercises the smem-ring and TMEM double-buffer phase wrap-around
    M, N, K = 128 * 40, 1024, 320
    A, W = _rand(M, K), _rand(N, K, scale=K ** -0.5)
    ref = G.ref_linear(A, W)
    out = G.linear(A, W, out_dtype=torch.float32)
    assert G.rel_err(out, ref) < 2e-3, G.describe_mismatch(out, ref, 2e-3)


def test_linear_strided_operands_and_simt_crosscheck():
    M, N, K = 333, 160, 96
    big = _rand(M, 3 * K)
    A = big[:, K:2 * K]                                              # lda = 3K
    W = _rand(N, K, scale=K ** -0.5)
    ref = G.ref_linear(A, W)
    assert G.rel_err(G.linear(A, W, out_dtype=torch.float32), ref) < 2e-3
    assert G.rel_err(G.linear(A, W, out_dtype=torch.float32, simt=True), ref) < 2e-3


def test_linear_rejects_bad_arguments():
    A, W = _rand(16, 12), _rand(8, 12)                               # K % 8 != 0
    with pytest.raises(RuntimeError, match="multiples of 8"):
        G.linear(A, W)


# ---- positional conv + LayerNorm / LayerNorm --------------------------------------------------------
@pytest.mark.parametrize("B,H,W,C,M", [(2, 14, 14, 384, 16), (3, 7, 5, 64, 0), (2, 28, 28, 192, 0), (1, 9, 11, 512, 16),
                                       (3, 56, 56, 96, 0), (2, 13, 9, 320, 16), (2, 8, 8, 32, 0), (5, 28, 28, 128, 0), (2, 6, 6, 34, 0)])
def test_posembed_layernorm(B, H, W, C, M):
    T = H * W + M
    tok = _rand(B, T, C)
    dw = torch.randn(C, 1, 3, 3, device="cuda") * 0.2
    db = torch.randn(C, device="cuda") * 0.1
    dw_packed = dw.reshape(C, 9).t().contiguous().clone()
    dw_packed[4] += 1.0
    resid = torch.empty_like(tok)
    norm = torch.empty_like(tok)
    L = G.lib()
    stats = torch.full((B * T, 2), -1.0, device="cuda") if C % 8 == 0 else None
    G.ok(L.lmv_posembed_layernorm(G.ptr(tok), G.ptr(dw_packed), G.ptr(db), G.ptr(resid), G.ptr(norm), G.ptr(stats), B, H, W, T, C, 1e-6, G.stream()))
    x = tok[:, :H * W].float().transpose(1, 2).reshape(B, C, H, W)
    xt = (x + torch.nn.functional.conv2d(x, dw, db, padding=1, groups=C)).flatten(2).transpose(1, 2)
    full = torch.cat([xt, tok[:, H * W:].float()], dim=1)
    ref_n = torch.nn.functional.layer_norm(full, (C,), eps=1e-6)
    assert G.rel_err(resid, full) < TOL
    assert G.rel_err(norm, ref_n) < TOL and G.cosine(norm, ref_n) > 0.9999
    if stats is not None:   # statistics describe the STORED (bf16) rows exactly
        r = resid.float().reshape(B * T, C)
        assert torch.allclose(stats[:, 0], r.sum(-1), rtol=1e-4, atol=1e-3)
        assert torch.allclose(stats[:, 1], (r * r).sum(-1), rtol=1e-4, atol=1e-3)
    if stats is not None:   # resid + stats only: the TMA-tiled kernel the block schedule uses
        resid2 = torch.empty_like(tok)
        stats2 = torch.full((B * T, 2), -1.0, device="cuda")
        G.ok(L.lmv_posembed_layernorm(G.ptr(tok), G.ptr(dw_packed), G.ptr(db), G.ptr(resid2), None, G.ptr(stats2), B, H, W, T, C, 1e-6, G.stream()))
        assert G.rel_err(resid2, full) < TOL
        r2 = resid2.float().reshape(B * T, C)
        assert torch.allclose(stats2[:, 0], r2.sum(-1), rtol=1e-4, atol=1e-3)
        assert torch.allclose(stats2[:, 1], (r2 * r2).sum(-1), rtol=1e-4, atol=1e-3)
        assert torch.equal(resid2, resid), "tiled and per-token kernels must round identically"
    # LayerNorm-only mode (no conv), norm output only
    norm2 = torch.empty_like(tok)
    G.ok(L.lmv_posembed_layernorm(G.ptr(tok), None, None, None, G.ptr(norm2), None, B, H, W, T, C, 1e-6, G.stream()))
    assert G.rel_err(norm2, torch.nn.functional.layer_norm(tok.float(), (C,), eps=1e-6)) < TOL


@pytest.mark.parametrize("R,C,gelu,affine", [(64, 2048, True, True), (33, 96, False, False), (512, 384, False, True)])
def test_layernorm(R, C, gelu, affine):
    x = _rand(R, C)
    g = 1 + 0.2 * torch.randn(C, device="cuda") if affine else None
    b = 0.1 * torch.randn(C, device="cuda") if affine else None
    out = torch.empty_like(x)
    G.ok(G.lib().lmv_layernorm(G.ptr(x), G.ptr(out), G.ptr(g), G.ptr(b), R, C, 1e-5, int(gelu), 0, 0, 0, G.stream()))
    ref = torch.nn.functional.layer_norm(x.float(), (C,), g, b, eps=1e-5)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert G.rel_err(out, ref) < TOL


def test_layernorm_row_remap_into_unified_buffer():
    B, M, N, C = 3, 16, 49, 64
    T = N + M
    x = _rand(B * M, C)
    dst = torch.zeros(B, T, C, dtype=torch.bfloat16, device="cuda")
    G.ok(G.lib().lmv_layernorm(G.ptr(x), G.ptr(dst), None, None, B * M, C, 1e-5, 0, M, T, N, G.stream()))
    ref = torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-5).reshape(B, M, C)
    assert G.rel_err(dst[:, N:], ref) < TOL
    assert float(dst[:, :N].abs().max()) == 0.0


# ---- im2col + conv on the GEMM -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_conv_via_im2col_gemm(dtype):
    B, Cin, H, W, Co = 2, 3, 64, 48, 48
    x = torch.randn(B, Cin, H, W, device="cuda").to(dtype)
    w = torch.randn(Co, Cin, 3, 3, device="cuda") * 0.2
    b = torch.randn(Co, device="cuda") * 0.1
    Ho, Wo, Kp = H // 2, W // 2, 32
    patches = torch.empty(B * Ho * Wo, Kp, dtype=torch.bfloat16, device="cuda")
    G.ok(G.lib().lmv_stem_im2col(G.ptr(x), G.F32 if dtype == torch.float32 else G.BF16, G.ptr(patches), B, Cin, H, W, G.stream()))
    wp = torch.zeros(Co, Kp, device="cuda")
    wp[:, :27] = w.reshape(Co, 27)
    out = G.linear(patches, G.bf(wp), b, gelu=True)
    ref = torch.nn.functional.gelu(torch.nn.functional.conv2d(x.float(), G.bf(w).float(), b, stride=2, padding=1))
    ref = ref.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Co)
    assert G.rel_err(out, ref) < TOL


@pytest.mark.parametrize("tc", ["1", "0"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C1", [(2, 224, 224, 48), (3, 64, 96, 32), (1, 33, 47, 48), (37, 224, 224, 48)])
def test_stem_conv1_direct(dtype, B, H, W, C1, tc, monkeypatch):
    """First stem convolution + folded BN + GELU without an im2col detour: the tcgen05 kernel (implicit GEMM, K = 27 -> 32; several
    tiles per persistent CTA at B = 37) and the CUDA-core kernel, against conv2d + exact GELU in fp32."""
    monkeypatch.setenv("LMV_STEM_TC", tc)
    x = torch.randn(B, 3, H, W, device="cuda").to(dtype)
    w = torch.randn(C1, 3, 3, 3, device="cuda") * 0.3
    bias = torch.randn(C1, device="cuda") * 0.2
    out, wp = G.stem_conv1(x, w, bias)
    wr = wp[:, :27].float().reshape(C1, 3, 3, 3)
    xr = G.bf(x).float() if tc == "1" else x.float()      # the tensor-core operand is bf16
    ref = torch.nn.functional.gelu(torch.nn.functional.conv2d(xr, wr, bias, stride=2, padding=1))
    ref = ref.permute(0, 2, 3, 1).reshape(B, -1, C1)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float().reshape(-1, C1), ref.reshape(-1, C1), TOL)


@pytest.mark.parametrize("B,H,W,C,Co,extra", [(2, 28, 28, 48, 96, 0), (2, 14, 14, 192, 384, 16), (1, 7, 9, 64, 128, 0)])
def test_conv3x3s2_via_im2col_gemm(B, H, W, C, Co, extra):
    T = H * W + extra
    tok = _rand(B, T, C)
    w = torch.randn(Co, C, 3, 3, device="cuda") * (9 * C) ** -0.5
    b = torch.randn(Co, device="cuda") * 0.1
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    patches = torch.empty(B * Ho * Wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    G.ok(G.lib().lmv_im2col_3x3s2(G.ptr(tok), G.ptr(patches), B, H, W, T, C, G.stream()))
    wp = G.bf(w.permute(0, 2, 3, 1).reshape(Co, 9 * C))
    out = G.linear(patches, wp, b)
    x = tok[:, :H * W].float().transpose(1, 2).reshape(B, C, H, W)
    ref = torch.nn.functional.conv2d(x, G.bf(w).float(), b, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    assert G.rel_err(out, ref) < TOL


CONV_CASES = [
    # (B, H, W, Cin, Cout, extra rows per input image, output rows per image (0 = dense))
    (3, 112, 112, 48, 96, 0, 0),      # stem conv 2 (Base): 56-wide output rows, 2 per tile, K blocks padded 48 -> 64
    (2, 56, 56, 96, 192, 0, 0),       # stage 2 downsample: 96 = 64 + 32 channels per tap
    (3, 28, 28, 192, 384, 0, 212),    # stage 3 downsample into a unified [B, 196 + 16, C] buffer
    (5, 14, 14, 384, 512, 16, 65),    # stage 4 downsample: two whole 7x7 images per tile (odd batch), unified input and output
    (2, 7, 9, 64, 128, 0, 0),         # odd sizes: right / bottom padding, maps smaller than a tile
    (1, 33, 45, 32, 64, 3, 0),        # rows that do not divide into equal tiles
    (2, 64, 256, 24, 32, 0, 0),       # 128-wide output rows (512 x 512 input of the stem), one per tile; Cin < 64
    (7, 2, 2, 8, 32, 0, 0),           # 1 x 1 output maps, many images per tile
]


@pytest.mark.parametrize("B,H,W,C,Co,extra,out_rows", CONV_CASES)
def test_conv3x3s2_implicit_gemm(B, H, W, C, Co, extra, out_rows):
    """conv3x3 / stride 2 / pad 1 as an implicit GEMM (strided TMA boxes gather the patch rows; reference: nn.Conv2d(.., 3, 2, 1) +
    folded BatchNorm, models/lemevit.py:702-703,715-716) against torch.conv2d, and bit-identical to the im2col + GEMM route."""
    T = H * W + extra
    tok = _rand(B, T, C)
    w = torch.randn(Co, C, 3, 3, device="cuda") * (9 * C) ** -0.5
    b = torch.randn(Co, device="cuda") * 0.1
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    wp = G.bf(w.permute(0, 2, 3, 1).reshape(Co, 9 * C))
    rows = out_rows if out_rows else Ho * Wo
    out = torch.full((B, rows, Co), 7.0, dtype=torch.bfloat16, device="cuda")
    G.ok(G.lib().lmv_conv3x3s2(G.ptr(tok), G.ptr(wp), G.ptr(b), G.ptr(out), B, H, W, T, C, Co, out_rows, G.stream()))
    x = tok[:, :H * W].float().transpose(1, 2).reshape(B, C, H, W)
    ref = torch.nn.functional.conv2d(x, G.bf(w).float(), b, stride=2, padding=1).permute(0, 2, 3, 1).reshape(B, Ho * Wo, Co)
    got = out[:, :Ho * Wo]
    assert G.rel_err(got, ref) < TOL, G.describe_mismatch(got.float().reshape(-1, Co), ref.reshape(-1, Co), TOL)
    assert (out[:, Ho * Wo:] == 7.0).all(), "rows behind the image tokens of a unified buffer must stay untouched"
    patches = torch.empty(B * Ho * Wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    G.ok(G.lib().lmv_im2col_3x3s2(G.ptr(tok), G.ptr(patches), B, H, W, T, C, G.stream()))
    via = G.linear(patches, wp, b).reshape(B, Ho * Wo, Co)
    if C % 16 == 0:   # the K = 16 groups of the MMA instructions then hold the same products in both routes
        assert torch.equal(via, got), "the implicit GEMM accumulates the same products in the same order"
    else:
        assert G.rel_err(got, via.float()) < 2e-3


# ---- attention ----------------------------------------------------------------------------------------
ATTN_CASES = [
    # (B, heads, Lq, Lk)   — S blocks (196/49, and 16x16 meta), D x-branch (N x 16), D/C c-branch (16 x N)
    (2, 6, 196, 196), (2, 10, 49, 49), (3, 12, 16, 16), (2, 3, 784, 16), (2, 3, 16, 784), (1, 2, 16, 3136),
    (1, 4, 1024, 1024), (2, 2, 50, 77),
]


@pytest.mark.parametrize("B,h,Lq,Lk", ATTN_CASES)
@pytest.mark.parametrize("impl", [1, 0])
def test_attention(B, h, Lq, Lk, impl):
    C = h * 32
    qkv_q = _rand(B, Lq, 3 * C)
    qkv_k = _rand(B, Lk, 3 * C)
    q = qkv_q[:, :, :C].unflatten(2, (h, 32))
    k = qkv_k[:, :, C:2 * C].unflatten(2, (h, 32))
    v = qkv_k[:, :, 2 * C:].unflatten(2, (h, 32))
    scale = 0.3
    out = G.attention(q, k, v, scale, impl=impl)
    ref = G.ref_attention(q, k, v, scale)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999


@pytest.mark.parametrize("B,h,Lq,Lk", [(2, 3, 16, 3136), (2, 6, 16, 784), (1, 2, 16, 500), (3, 4, 16, 300), (1, 3, 16, 128),
                                       (2, 8, 16, 1000), (1, 1, 8, 129), (2, 3, 16, 16384)])
def test_attention_meta_queries_over_image_tokens(B, h, Lq, Lk):
    """CrossAttention / DualCrossAttention meta-token branch: 16 queries per head, softmax over all N image tokens."""
    C = h * 32
    qc = _rand(B, Lq, 3 * C)
    kvx = _rand(B, Lk, 3 * C)
    q = qc[:, :, :C].unflatten(2, (h, 32))
    k = kvx[:, :, C:2 * C].unflatten(2, (h, 32))
    v = kvx[:, :, 2 * C:].unflatten(2, (h, 32))
    scale = C ** -0.5 * 3.0        # peaky enough that the per-tile maxima differ
    out = G.attention_meta(q, k, v, scale)
    ref = G.ref_attention(q, k, v, scale)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float().flatten(0, 1), ref.flatten(0, 1), TOL)
    assert torch.equal(out, G.attention_meta(q, k, v, scale)), "split-softmax merge must be deterministic"
    simt = G.attention(q, k, v, scale, impl=1)
    assert G.rel_err(out, simt.float()) < TOL


# ---- tail / export ------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,h,T,N", [(3, 12, 212, 196), (2, 16, 65, 49), (2, 6, 196, 196), (1, 3, 224, 208), (5, 2, 130, 128), (2, 4, 16, 16),
                                     (300, 12, 212, 196), (2, 12, 1024, 1024), (3, 16, 256, 256), (1, 2, 500, 500), (2, 3, 2048, 2048),
                                     (2, 3, 300, 300), (1, 3, 1900, 1900), (1, 4, 4096, 4096)])
def test_attention_self_two_segments(B, h, T, N):
    """Image tokens (rows < N) and meta tokens (rows >= N) of a packed qkv buffer attend within their own segment."""
    Cc = h * 32
    qkv = _rand(B, T, 3 * Cc)
    scale = 32 ** -0.5
    out = G.attention_self(qkv, h, N, scale)
    q, k, v = (qkv[:, :, i * Cc:(i + 1) * Cc].reshape(B, T, h, 32) for i in range(3))
    ref = torch.empty(B, T, Cc, device="cuda")
    ref[:, :N] = G.ref_attention(q[:, :N], k[:, :N], v[:, :N], scale)
    if T > N:
        ref[:, N:] = G.ref_attention(q[:, N:], k[:, N:], v[:, N:], scale)
    assert G.rel_err(out, ref) < TOL and G.cosine(out, ref) > 0.9999, G.describe_mismatch(out.float().reshape(B * T, Cc), ref.reshape(B * T, Cc), TOL)
    assert torch.equal(out, G.attention_self(qkv, h, N, scale))     # deterministic


def test_tail_and_nchw_export():
    B, N, M, C = 3, 49, 16, 320
    T = N + M
    tok = _rand(B, T, C)
    bs, bb = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    g, be = 1 + 0.2 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    feat = torch.empty(B, C, dtype=torch.bfloat16, device="cuda")
    L = G.lib()
    G.ok(L.lmv_tail(G.ptr(tok), T * C, N, G.ptr(tok[:, N:]), T * C, M, C, G.ptr(bs), G.ptr(bb), G.ptr(g), G.ptr(be), 1e-5, G.ptr(feat), B, G.stream()))
    x, c = tok[:, :N].float(), tok[:, N:].float()
    ref = bs * x.mean(1) + bb + torch.nn.functional.layer_norm(c, (C,), g, be, eps=1e-5).mean(1)
    assert G.rel_err(feat, ref) < TOL
    for dt, code in ((torch.float32, G.F32), (torch.bfloat16, G.BF16)):
        out = torch.empty(B, C, 7, 7, dtype=dt, device="cuda")
        G.ok(L.lmv_tokens_to_nchw(G.ptr(tok), G.ptr(out), B, 7, 7, T, C, code, G.stream()))
        assert torch.equal(out.float(), x.transpose(1, 2).reshape(B, C, 7, 7))       # pure data movement: bit exact
