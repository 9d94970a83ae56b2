// Internal launchers of the non-GEMM kernels (tokens.cu, attention.cu).
#pragma once
#include "common.h"

namespace lmv {

struct PosLnArgs {
  const bf16* tokens;   // [B, T, C]
  const float* dw_w;    // [9][C] (centre tap already +1) or null => no conv (LayerNorm only)
  const float* dw_b;    // [C]
  bf16* resid_out;      // nullable, [B, T, C]
  bf16* norm_out;       // nullable, [B, T, C]
  int B, H, W, T, C;
  float eps;
  float* stats_out = nullptr;   // nullable, [B*T][parts][2]: (sum_c v, sum_c v^2) of the bf16-rounded resid_out row (LayerNorm fold input)
  int max_parts = 1;            // the tiled kernel may split the channels over up to this many CTAs, one statistics partial each
};
int posembed_ln_run(const PosLnArgs& a, cudaStream_t s);

// TMA-tiled x' = x + dwconv(x) (+ statistics) — posembed.cu; used whenever resid_out is wanted without norm_out
struct PosEmbedOp {
  CUtensorMap tm;
  PosLnArgs a;
  int TW, TH, tiles_x, tiles_y, cbox, ncb, smem, sub_bytes, threads, col_threads;
  int parts;   // channel slices == statistics partials per row (stats_out is [B*T][parts][2])
};
bool posembed_tile_supported(const PosLnArgs& a);
int posembed_tile_prepare(const PosLnArgs& a, PosEmbedOp* op);
int posembed_tile_run(const PosEmbedOp& op, cudaStream_t s);

struct LnArgs {
  const bf16* in;       // [R, C]
  bf16* out;
  const float* gamma;   // nullable
  const float* beta;    // nullable
  int R, C;
  float eps;
  int act_gelu;
  int grp_rows, grp_stride, grp_off;  // output row remap (grp_rows == 0: identity)
};
int layernorm_run(const LnArgs& a, cudaStream_t s);

// stats[r] = (sum_c x[r,c], sum_c x[r,c]^2) of dense bf16 rows — statistics producer of the SIMT cross-check path
int row_stats_run(const bf16* x, float* stats, int R, int C, cudaStream_t s);

struct AttnArgs {
  const bf16 *q, *k, *v;
  bf16* out;
  long long q_bs, k_bs, v_bs, o_bs;  // per-image strides (elements)
  int q_rs, k_rs, v_rs, o_rs;        // per-row strides (elements)
  int B, heads, Lq, Lk;
  float scale;
};
int attention_simt_run(const AttnArgs& a, cudaStream_t s);
int attention_tc_run(const AttnArgs& a, cudaStream_t s);      // tcgen05 self/cross attention (attention.cu)
bool attention_tc_supported(const AttnArgs& a);
// persistent self-attention; T <= 224 rows per image in two independent segments (T > 224: plain self-attention, split over key blocks): rows [0, N) among themselves (image tokens)
// and rows [N, T) among themselves (the meta tokens of a unified [B, N+M, 3C] qkv buffer); a.Lq / a.Lk are ignored (attention_self.cu)
bool attention_self_supported(const AttnArgs& a, int T, int N);
size_t attention_self_workspace(int B, int heads, int T);   // split-KV partials (T > 224), 0 otherwise
int attention_self_run(const AttnArgs& a, int T, int N, void* workspace, size_t workspace_bytes, cudaStream_t s);
// few queries (meta tokens) x many keys (image tokens): split-N tcgen05 kernel + partial merge (attention_meta.cu)
bool attention_meta_supported(const AttnArgs& a);
size_t attention_meta_workspace(const AttnArgs& a);
int attention_meta_run(const AttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s);

struct StemArgs {
  const void* x;  // NCHW, f32 or bf16; stem_conv1 only: 8-bit pixels, LMV_DTYPE_U8 [B,3,H,W] / LMV_DTYPE_U8_NHWC [B,H,W,3]
  int x_dtype;
  bf16* out;      // patches [B*Ho*Wo, Kp], Kp = round_up(9*Cin, 8), k = ci*9 + ky*3 + kx
  int B, Cin, H, W;
  float mean[3] = {0.485f * 255.f, 0.456f * 255.f, 0.406f * 255.f};   // 8-bit input: bf16((u8 - mean[c]) / std[c])
  float std[3] = {0.229f * 255.f, 0.224f * 255.f, 0.225f * 255.f};
  int tensor_core = 1;   // stem_conv1: tcgen05 implicit-GEMM kernel (0: the CUDA-core kernel, A/B switch "stem_tc")
};
int stem_im2col_run(const StemArgs& a, cudaStream_t s);
// direct first stem convolution + folded BN + GELU: a.out receives token-major [B, Ho*Wo, C1] (no im2col detour)
bool stem_conv1_supported(int Cin, int C1);
int stem_conv1_run(const StemArgs& a, const bf16* w, const float* bias, int C1, cudaStream_t s);

struct Im2colArgs {
  const bf16* in;  // [B, T, C], first H*W rows of each image
  bf16* out;       // [B*Ho*Wo, 9*C]
  int B, H, W, T, C;
};
int im2col_run(const Im2colArgs& a, cudaStream_t s);

struct TailArgs {
  const bf16* x; long long x_bs; int N;
  const bf16* c; long long c_bs; int M;
  int C;
  const float *bn_scale, *bn_shift, *ln_gamma, *ln_beta;
  float eps;
  bf16* feat;  // [B, C]
  int B;
};
int tail_run(const TailArgs& a, cudaStream_t s);

struct ToNchwArgs {
  const bf16* tokens;  // [B, T, C]
  void* out;           // [B, C, H, W]
  int B, H, W, T, C, out_dtype;
};
int tokens_to_nchw_run(const ToNchwArgs& a, cudaStream_t s);

// c[b, m, :] = c0[m, :]   (meta_tokens.repeat(B,1,1), models/lemevit.py:833, after the folded meta_ds_0)
int broadcast_rows_run(const bf16* src, bf16* dst, int rows, int C, int B, long long dst_bs, cudaStream_t s);

// dst[r, :] = src[(r / grp_rows) * grp_stride + grp_off + r % grp_rows, :]   (gather meta-token rows
// out of a unified [B, N+M, C] token buffer)
int gather_rows_run(const bf16* src, bf16* dst, int rows, int C, int grp_rows, int grp_stride, int grp_off,
                    cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// fused MLP branch (mlp_fused.cu): out = resid + W2 gelu(W1' LN(x) + b1') + b2, hidden activation kept on chip
// ------------------------------------------------------------------------------------------------
struct MlpArgs {
  const bf16* x = nullptr;       // [R, C] raw rows (LayerNorm folded through ln_stats / cs1), dense
  const bf16* resid = nullptr;   // [R, C] residual (null: x)
  bf16* out = nullptr;           // [R, C] (may alias x / resid)
  const bf16* W1 = nullptr;      // [Hd, C]  (LayerNorm affine folded in)
  const float* b1 = nullptr;     // [Hd]
  const float* cs1 = nullptr;    // [Hd] column sums of W1 (with ln_stats) or null
  const bf16* W2 = nullptr;      // [C, Hd]
  const float* b2 = nullptr;     // [C]
  const float* ln_stats = nullptr;   // [R][ln_parts][2] partial (sum, sum^2) of the x rows, or null (input already normalised)
  int ln_parts = 1;
  float ln_eps = 1e-6f;
  int R = 0, C = 0, Hd = 0;
};
struct MlpParams {
  int R, C, Hd, tiles, kb1, chunks, nparts2, n2, nx, nh, n1slots, n2slots, slot2_bytes, const_bytes, res_smem;
  int hc;          // hidden columns per chunk: 128 (C <= 256) or 64 (256 < C <= 384: acc2 takes 384 of the 512 TMEM columns)
  int w1_kpb;      // K-blocks per W1' TMA box (> 1: one 3-D box per slot, so 8 KB K-blocks do not pay the per-box cost)
  int w1_boxes;    // W1' boxes per hidden chunk = kb1 / w1_kpb
  int pair;        // 1: CTA-pair kernel (cta_group::2, 256 rows per cluster; `tiles` counts pair tiles)
  const float *b1, *cs1, *b2, *ln_stats;
  int ln_parts;
  float ln_eps, ln_inv_k;
  const bf16* resid;
  bf16* out;
};
struct MlpOp {
  CUtensorMap tmX, tmW1, tmW2;
  MlpParams p;
  int grid = 0, smem_bytes = 0;
};
bool mlp_fused_supported(int C, int Hd);
int mlp_fused_prepare(const MlpArgs& a, MlpOp* op);
int mlp_fused_run(const MlpOp& op, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Fused cross-attention blocks (dca_fused.cu + meta_branch.cu): CrossAttention 'C' blocks (models/lemevit.py:477-486,584-613)
// and DualCrossAttention 'D' blocks (:252-302,542-582) with the image-side projections ABSORBED into per-image operands:
//
//   x-branch   dx[n]   = proj_x( concat_h softmax_m( s_x q1_h[n] . k2_h[m] ) v2_h )         q1 = Wq xn[n] + bq
//                      = sum_(h,m) P[n,(h,m)] Vt[(h,m)] + b_px        S[n,(h,m)] = xn[n] . Kt[(h,m)] + kappa[(h,m)]
//              with  Kt[(h,m)] = s_x Wq_h^T k2_h[m]   kappa = s_x bq_h . k2_h[m]   Vt[(h,m)] = Wpx[:, h] v2_h[m]     (all [R = heads*M, C])
//   c-branch   attn_c[(h,m)] = softmax_n( s_c q2_h[m] . k1_h[n] ) v1_h[n]                     k1 = Wk xn + bk, v1 = Wv xn + bv
//                            = Wv_h Zbar[(h,m)] + bv_h,  Zbar = sum_n softmax_n(Sc) xn[n]     Sc[(h,m),n] = Qt[(h,m)] . xn[n] + beta[(h,m)]
//              with  Qt[(h,m)] = s_c Wk_h^T q2_h[m]   beta = s_c bk_h . q2_h[m]
//
// so the N x 3C projection qkv1 (and q1/k1/v1 themselves) never exists: the image-token kernel only contracts the token tile
// with the [R, C] operands of its image.  xn = LayerNorm_noaffine(xt) is applied AFTER the contraction from the per-row
// statistics (xn = r (xt - mu)): S = r (xt.Kt - mu sum(Kt)) + kappa, Zbar = sum p r (xt - mu).
// ------------------------------------------------------------------------------------------------
constexpr int kDcaM = 16;        // meta tokens per image (queries_len of every published variant); other values use the unfused schedule
constexpr int kDcaTile = 128;    // image tokens per tile

struct DcaGeom {
  int B, N, C, heads, R;         // R = heads * kDcaM <= 128
  int tiles, seg_tiles, segs;    // 128-token tiles per image, cut into `segs` segments of seg_tiles tiles (one softmax partial each)
  int dup, ncopy;                // R <= 64: the (h,m) rows are duplicated at lanes 64.., each copy takes half of a tile's tokens
  int parts;                     // segs * ncopy partials per image
};
bool dca_supported(int N, int C, int heads, int M);
DcaGeom dca_geometry(int B, int N, int C, int heads);

// device scratch of ONE fused block (carved from the caller's workspace; byte offsets from dca_workspace_layout)
struct DcaWs {
  bf16* kt;          // [B][R][C]   x-branch keys in token space (scaled by s_x log2 e)                   ('D' only)
  bf16* qt;          // [B][R][C]   c-branch queries in token space (scaled by s_c log2 e)
  bf16* vt;          // [B][C][R]   x-branch values pushed through proj_x, transposed                      ('D' only)
  float* cst;        // [B][4][R]   (sum_c Kt, kappa, sum_c Qt, beta): sums of the bf16-ROUNDED rows, kappa/beta in the log2 domain
  float4* part_ml;   // [B][parts][R]      (running max, sum p, sum p' mu, -) of a segment, log2 domain
  float* part_z;     // [B][parts][R][C]   sum_n p'_n xt[n]  (p' = bf16(p r_n))
};
size_t dca_workspace_bytes(const DcaGeom& g);
DcaWs dca_workspace_carve(const DcaGeom& g, void* base);

struct MetaPreArgs {
  const bf16* c;             // [B, M, C] meta tokens (block input)
  const bf16* Wc;            // projection applied to LN1(c): 'D' qkv2 [3C, C] (q2 | k2 | v2 rows), 'C' q [C, C]; LN1 affine folded
  const float* bc;
  int nc;                    // rows of Wc (3C or C)
  int q_off, k_off, v_off;   // first row of q2 / k2 / v2 inside that projection (k_off, v_off < 0: absent, 'C' blocks)
  const bf16* WxqT;          // image-side query projection (LN1-folded) TRANSPOSED: WxqT[j][i] = Wq[i][j], [C, C] + bias [C]: 'D' only
  const float* bxq;
  const bf16* WxkT;          // image-side key projection (LN1-folded) transposed [C, C] + bias [C]
  const float* bxk;
  const bf16* Wpx;           // proj_x [C, C] ('D' only)
  float scale_x, scale_c;    // softmax scales of the two branches (natural-log domain)
  int B, C, heads;
  float eps;
  DcaWs ws;
};
int meta_pre_run(const MetaPreArgs& a, cudaStream_t s);

struct MetaPostArgs {
  bf16* c;                   // [B, M, C] meta tokens, updated in place: c += proj(attn_c); c += mlp(LN2(c))
  const bf16* Wxv;           // image-side value rows [C, C] (LN1-folded) + bias
  const float* bxv;
  const bf16* Wp;            // proj_c ('D') / proj ('C')  [C, C]
  const float* bp;
  const bf16* W1;            // mlp.0 with LN2 folded [Hd, C]
  const float* b1;
  const bf16* W2;            // mlp.3 [C, Hd]
  const float* b2;
  int B, C, heads, Hd, parts;
  float eps;
  DcaWs ws;
};
int meta_post_run(const MetaPostArgs& a, cudaStream_t s);
// [post of block j] -> [pre of block j + 1] of the same stage in ONE launch (either may be null)
int meta_chain_run(const MetaPostArgs* post, const MetaPreArgs* pre, cudaStream_t s);

// meta_token_downsample[i] (models/lemevit.py:729-745) in one launch: Linear(Cp, 4Cp) -> LayerNorm -> GELU -> Linear(4Cp, C) -> LayerNorm
struct MetaDsArgs {
  const bf16* in;            // rows of image b: in + b * in_bs, [16, Cp] dense
  long long in_bs;
  bf16* out;                 // rows of image b: out + b * out_bs, [16, C] dense
  long long out_bs;
  const bf16* W0;            // [4Cp, Cp]
  const float *b0, *g1, *be1;
  const bf16* W3;            // [C, 4Cp]
  const float *b3, *g4, *be4;
  int B, Cp, C;
  float eps;
};
bool meta_downsample_supported(int Cp, int C);
int meta_downsample_run(const MetaDsArgs& a, cudaStream_t s);

struct DcaXArgs {
  const bf16* xt;            // [B, N, C] image tokens after x + dwconv(x) (raw; LayerNorm through stats1)
  const float* stats1;       // [B*N][parts1][2] partial (sum, sum^2) of the xt rows
  int parts1;
  float eps;
  int do_x;                  // 1: 'D' block (x-branch + c-branch); 0: 'C' block (c-branch only, x untouched)
  const float* bpx;          // proj_x bias [C]                                  (do_x)
  bf16* xout;                // [B, N, C]: xt + dx   (may alias xt)              (do_x)
  float* stats2;             // [B*N][2]: (sum, sum^2) of the stored xout rows   (do_x)
  DcaGeom g;
  DcaWs ws;
  int zsplit;                // test hook: issue the Z accumulation per 64-channel block instead of one MN-major operand spanning C
  int force_serial;          // test hook: one tile at a time even where the pipelined schedule fits
};
struct DcaXParams {
  DcaGeom g;
  int do_x, kb, kbr, nx, rq, zsplit, parts1, pipe;
  float eps, inv_c;
  const float* stats1;
  const float* cst;
  const float* bpx;
  const bf16* xres;
  bf16* xout;
  float* stats2;
  float4* part_ml;
  float* part_z;
  int smem_x, smem_kt, smem_qt, smem_vt, smem_p, smem_pc;   // byte offsets of the shared-memory regions
};
struct DcaXOp {
  CUtensorMap tmX, tmKt, tmQt, tmVt;
  DcaXParams p;
  int grid = 0, smem_bytes = 0;
};
int dca_x_prepare(const DcaXArgs& a, DcaXOp* op);
int dca_x_run(const DcaXOp& op, cudaStream_t s);

}  // namespace lmv
