/* liblemevit_b200 — C ABI of the Blackwell-native LeMeViT backbone forward pass.
 *
 * The reference (ViTAE-Transformer/LeMeViT) has no C/FFI boundary of its own: its hot path is the
 * Python `LeMeViT.forward` / `forward_features` (models/lemevit.py:809-836) and the mmseg/mmdet
 * backbone `forward` (semantic_segmentation/mmseg/models/backbones/lemevit.py:800-827), reached
 * through `timm.create_model` / `BACKBONES.register_module()`.  This header is the boundary a
 * binding for that path targets: plain pointers and sizes, no torch types.  `lemevit_b200/_native.py`
 * is the ctypes binding; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all buffers;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - every function returns LMV_OK (0) or a negative LMV_ERR_* code; `lmv_last_error()` returns a
 *     thread-local message for the last failure;
 *   - the library is re-entrant per (plan, stream); a plan must not be used from two threads at once;
 *   - there is NO CPU fallback: every entry point launches sm_100a kernels or fails.
 */
#ifndef LEMEVIT_B200_H_
#define LEMEVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMV_OK 0
#define LMV_ERR_INVALID (-1)     /* bad argument / shape / alignment                                  */
#define LMV_ERR_UNSUPPORTED (-2) /* configuration outside what the kernels implement                  */
#define LMV_ERR_CUDA (-3)        /* a CUDA runtime / driver call failed                               */
#define LMV_ERR_OOM (-4)         /* CUDA out of memory (message contains "CUDA out of memory")        */

#define LMV_DTYPE_BF16 0
#define LMV_DTYPE_F32 1
/* input images only (x_dtype of the lmv_forward_* calls and lmv_stem_conv1): raw 8-bit pixels, normalised inside the first stem
 * convolution as bf16((float(u8) - mean[c]) / std[c]) — bit-identical to `x.float().sub_(mean).div_(std).to(bfloat16)`, the GPU side
 * of the timm prefetcher the reference trains and validates with (main.py:399-428) — so the fp32 / bf16 image never exists in HBM. */
#define LMV_DTYPE_U8 2       /* [B, 3, H, W] planar (timm fast_collate)          */
#define LMV_DTYPE_U8_NHWC 3  /* [B, H, W, 3] interleaved (a decoded image as is) */

#define LMV_MAX_STAGES 8

/* Mirrors the structural kwargs of LeMeViT.__init__ (models/lemevit.py:664-687). */
typedef struct lmv_config {
  int num_stages;                  /* len(attn_type), 5 for every published variant                   */
  int depth[LMV_MAX_STAGES];       /* blocks per stage                                                */
  int embed_dim[LMV_MAX_STAGES];   /* channels per stage                                              */
  int mlp_hidden[LMV_MAX_STAGES];  /* int(mlp_ratio * embed_dim)                                      */
  char attn_type[LMV_MAX_STAGES];  /* 'C', 'D' or 'S' (models/lemevit.py:652-660)                     */
  int head_dim;                    /* 32                                                              */
  int queries_len;                 /* meta tokens M, 16                                               */
  int num_classes;                 /* 0 => no head                                                    */
  int in_chans;                    /* 3                                                               */
  int backbone;                    /* 0: classification model; 1: mmseg/mmdet/CD backbone copies
                                      (S-stage does not update meta tokens, returns 4 feature maps)   */
} lmv_config;

typedef struct lmv_tensor {
  const void* data; /* device pointer, 16-byte aligned */
  int64_t numel;
  int dtype;        /* LMV_DTYPE_* */
} lmv_tensor;

typedef struct lmv_plan lmv_plan; /* opaque: packed-weight table + cached launch schedules */

const char* lmv_last_error(void);
int lmv_version(void);

/* ---- whole-model path --------------------------------------------------------------------------
 * Packed weights: the folded/bf16 tensors produced by lemevit_b200/pack.py, in the order documented
 * there (and checked here tensor by tensor: count, numel, dtype).  The plan only stores the
 * pointers; the caller keeps the buffers alive. */
int lmv_packed_tensor_count(const lmv_config* cfg);
int lmv_plan_create(const lmv_config* cfg, const lmv_tensor* packed, int n_packed, lmv_plan** out);
void lmv_plan_destroy(lmv_plan* plan);
/* images processed per pass through the network (0 = whole batch at once); smaller chunks keep the
 * producer->consumer activations L2-resident. */
int lmv_plan_set_chunk(lmv_plan* plan, int images_per_chunk);
/* schedule options (A/B switches; every schedule is rebuilt afterwards).  Known names:
 *   "fused_mlp" (default 1): run `x + mlp(norm2(x))` as ONE kernel (lmv_mlp_fused) where the shape allows it;
 *   "fused_mlp_wide" (default 1; LMV_FUSED_MLP_WIDE in the environment overrides the default): also the C = 384 MLP of the stage-3
 *       'S' blocks — on the cta_group::2 CTA-pair kernel (two CTAs of a cluster share M = 256 MMAs and each loads half of every
 *       weight box), or with LMV_MLP_PAIR=0 on the single-CTA wide kernel;
 *   "direct_stem" (default 1): first stem convolution as a direct kernel (lmv_stem_conv1) instead of im2col + GEMM;
 *   "fused_self_attn" (default 1): image + meta token self-attention of an 'S' block in ONE persistent kernel (lmv_attention_self);
 *   "fused_dca" (default 1): 'C' / 'D' blocks through the fused cross-attention kernels (lmv_dca_block) where the shape allows it;
 *   "dca_pipe" (default 0): the pipelined schedule of the fused cross-attention kernel where tensor memory allows it (A/B switch;
 *   measured slower than the one-tile-at-a-time schedule on B200);
 *   "implicit_conv" (default 1): strided convolutions as implicit GEMMs (0: im2col kernel + GEMM);
 *   "stem_tc" (default 1): the direct first stem convolution on the tcgen05 kernel (0: the CUDA-core kernel). */
int lmv_plan_set_option(lmv_plan* plan, const char* name, int value);
/* test hook (block-level parity against the reference's forward hooks): after block `block` of stage `stage` every forward
 * copies the block's outputs to x_tokens_out [B, N, C] bf16 (token-major) and c_out [B, queries_len, C] bf16 (either may be
 * null); stage < 0 turns the tap off.  The buffers must hold the batch of the forwards that follow. */
int lmv_plan_set_tap(lmv_plan* plan, int stage, int block, void* x_tokens_out, void* c_out);
/* per-channel mean / std of the 8-bit input path, in pixel units (0..255); default: ImageNet mean / std x 255, what the
 * reference's default_cfg (`_cfg()`, models/lemevit.py:866) resolves to in timm's loader.  Needs 3 input channels. */
int lmv_plan_set_input_norm(lmv_plan* plan, const float* mean3, const float* std3);
/* bring-up switch: 1 routes every GEMM / attention through the plain SIMT cross-check kernels. */
int lmv_plan_set_debug_simt(lmv_plan* plan, int enable);
size_t lmv_workspace_bytes(const lmv_plan* plan, int batch, int H, int W);
/* number of kernel launches one forward of this shape issues (for bench.py's gpu_launches). */
int lmv_launch_count(lmv_plan* plan, int batch, int H, int W);

/* per-kernel-class device timing: when enabled, every forward records a CUDA event between consecutive
 * launches on the caller's stream; lmv_plan_get_profile waits for them and returns the totals accumulated since
 * profiling was enabled (returns the number of entries written, or a negative error). */
typedef struct lmv_profile_entry {
  const char* name;      /* kernel class, e.g. "gemm_tcgen05" */
  long long launches;
  double device_ms;      /* sum of CUDA-event durations */
  double flops;          /* sum of algorithmic FLOPs (2 x MAC) of those launches */
  double bytes;          /* sum of algorithmic bytes (compulsory reads + writes) */
} lmv_profile_entry;
int lmv_plan_set_profile(lmv_plan* plan, int enable);
int lmv_plan_get_profile(lmv_plan* plan, lmv_profile_entry* out, int max_entries);
/* human-readable per-launch-shape table of the same data (tuning aid); returns the untruncated length */
int lmv_plan_profile_report(lmv_plan* plan, char* buf, int buf_bytes);

/* replaces LeMeViT.forward (models/lemevit.py:831-836): x[B,in_chans,H,W] NCHW (bf16 or f32), or raw 8-bit pixels
 * (LMV_DTYPE_U8 / LMV_DTYPE_U8_NHWC, normalised in the stem: lmv_plan_set_input_norm) -> logits[B,num_classes] (bf16 or f32,
 * per logits_dtype). */
int lmv_forward_cls(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, void* workspace,
                    size_t workspace_bytes, void* logits, int logits_dtype, void* stream);
/* replaces LeMeViT.forward_features(x, c) (models/lemevit.py:809-829) and, with `logits`, the head on top of it (:831-836):
 * meta_tokens: nullable; [B, queries_len, embed_dim[0]] bf16 — the caller's own meta tokens, run through
 *   meta_token_downsample[0] on the device (null: the model's `meta_tokens` parameter, as LeMeViT.forward does at :833);
 * features: nullable; [B, embed_dim[-1]] bf16 = mean_HW(BN(x)) + mean_M(LN(c)) (:815-827);
 * logits: nullable (must be null when num_classes == 0); [B, num_classes] in logits_dtype. */
int lmv_forward_cls_features(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, const void* meta_tokens,
                             void* workspace, size_t workspace_bytes, void* features, void* logits, int logits_dtype, void* stream);
/* replaces the backbone copies' forward (semantic_segmentation/.../lemevit.py:822-827):
 * -> outs[k] = x after stage k+1 as contiguous NCHW [B, embed_dim[k+1], H/s, W/s], k = 0..3. */
int lmv_forward_features(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, void* workspace,
                         size_t workspace_bytes, void* const* outs, int n_outs, int out_dtype, void* stream);

/* ---- per-kernel entry points (unit tests, ncu) --------------------------------------------------*/
/* nn.Linear (+GELU) (+residual): out[M,N] = act(A[M,K] W[N,K]^T + bias) + residual.  tcgen05 path. */
int lmv_linear(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, void* out,
               int ldc, int M, int N, int K, int act_gelu, int out_dtype, int force_tile_n, void* stream);
/* lmv_linear with the two fused LayerNorm hooks of the block schedule:
 *   ln_stats  [M][ln_parts][2] fp32: partial (sum_k a, sum_k a^2) pairs of every A row (added up by the kernel),
 *             ln_colsum [N] fp32 (sum_k W[n,k]), ln_eps:
 *             out = act(LayerNorm_noaffine(A) W^T + bias) computed as r (A W^T - mu colsum) + bias in the epilogue
 *             (norm1 -> q/kv/qkv*, norm2 -> mlp.0, models/lemevit.py:560-564,600-601,632-635; gamma/beta are folded
 *             into W/bias at pack time);
 *   stats_out [rows][parts][2] fp32 with parts = lmv_linear_stats_parts(N, use_simt): partial (sum_n out, sum_n out^2)
 *             of every stored output row — plain stores, deterministic; needs N % 32 == 0 and a dense bf16 output.
 * use_simt != 0 runs the same contract on the SIMT cross-check kernels. */
int lmv_linear_fused(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, void* out,
                     int ldc, int M, int N, int K, int act_gelu, int out_dtype, const float* ln_stats, int ln_parts,
                     const float* ln_colsum, float ln_eps, float* stats_out, int use_simt, void* stream);
int lmv_linear_stats_parts(int N, int use_simt);
/* Fused MLP branch of a LeMeBlock — `x + mlp(norm2(x))`, models/lemevit.py:526-530 with :562,564,601,633,635 — in one
 * kernel: out[R,C] = resid + W2 gelu(W1 LN(x) + b1) + b2; the [R,Hd] hidden activation never leaves the SM.
 * x: [R,C] bf16 dense raw rows; ln_stats [R][ln_parts][2] + colsum1 [Hd] fold the LayerNorm exactly as in
 * lmv_linear_fused (both null: x is used as is); resid nullable (= x); out may alias x.
 * Requires C % 32 == 0, 32 <= C <= 384, Hd % 128 == 0 (LMV_ERR_UNSUPPORTED otherwise). */
int lmv_mlp_fused(const void* x, const void* resid, void* out, const void* W1, const float* b1, const float* colsum1,
                  const void* W2, const float* b2, const float* ln_stats, int ln_parts, float ln_eps, int R, int C, int Hd,
                  void* stream);
/* same contract as lmv_linear on the SIMT cross-check kernel */
int lmv_linear_simt(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                    void* out, int ldc, int M, int N, int K, int act_gelu, int out_dtype, void* stream);
/* x + dwconv3x3(x) followed by LayerNorm without affine (models/lemevit.py:546,589,619 + :513).
 * tokens: [B, T, C] bf16, the first N = H*W rows of each image are the H x W map, rows N..T-1 (meta
 * tokens) get LayerNorm only.  resid_out (nullable) receives x + dw(x); norm_out (nullable) the normalised rows;
 * stats_out (nullable, [B*T][2] fp32, needs C % 8 == 0) the (sum, sum of squares) of every stored resid_out row —
 * the input (ln_parts = 1) of the LayerNorm fold of lmv_linear_fused. */
int lmv_posembed_layernorm(const void* tokens, const float* dw_weight /*[9][C], centre tap +1*/,
                           const float* dw_bias, void* resid_out, void* norm_out, float* stats_out, int B, int H, int W,
                           int T, int C, float eps, void* stream);
/* LayerNorm over rows [R, C] bf16; gamma/beta nullable (no affine); optional GELU afterwards;
 * out row r -> (r / grp_rows) * grp_stride + grp_off + r % grp_rows when grp_rows > 0. */
int lmv_layernorm(const void* in, void* out, const float* gamma, const float* beta, int R, int C, float eps,
                  int act_gelu, int grp_rows, int grp_stride, int grp_off, void* stream);
/* softmax(scale * Q K^T) V per (image, head), head_dim 32.  Element (b, row, h, d) of Q lives at
 * q + b*q_bs + row*q_rs + h*32 + d (strides in elements), likewise K, V, out.
 * impl: 0 = best available (tcgen05 where implemented), 1 = SIMT cross-check kernel. */
int lmv_attention(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                  long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int Lq, int Lk,
                  float scale, int impl, void* stream);
/* Self-attention of an 'S' block (StandardAttention, models/lemevit.py:199-205) over T rows per image, head_dim 32, in ONE
 * persistent tcgen05 kernel: rows [0, N) (image tokens) attend keys [0, N); rows [N, T) (the meta tokens of a unified
 * [B, N+M, 3C] qkv buffer, forward_with_x :634) attend keys [N, T).  N == T: plain self-attention.  T > 224 (e.g. the 1024 tokens
 * of stage 3 at 512x512): N must equal T; the keys are split into blocks of <= 224 whose partial
 * (max, sum, output) go through `workspace` (lmv_attention_self_workspace bytes, 0 for T <= 224) and a merge kernel.  Same pointer /
 * stride conventions as lmv_attention. */
size_t lmv_attention_self_workspace(int B, int heads, int T);
int lmv_attention_self(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                       long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int T, int N, float scale,
                       void* workspace, size_t workspace_bytes, void* stream);
/* The meta-token side of CrossAttention / DualCrossAttention (models/lemevit.py:484, :300-302): Lq = M (16) queries
 * per head over Lk = N image tokens, heads * Lq <= 128, heads * 32 <= 256.  Persistent tcgen05 kernel (one CTA per SM
 * streams 128-token K/V tiles, block-diagonal Q so that all heads share one accumulation, running softmax state in
 * registers) + deterministic merge of the per-segment softmax partials; the bits of an image's output do not depend on
 * B or on its position in the batch.  workspace: lmv_attention_meta_workspace(...) bytes of device scratch, 16-byte aligned. */
size_t lmv_attention_meta_workspace(int B, int heads, int Lq, int Lk);
int lmv_attention_meta(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                       long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int Lq, int Lk,
                       float scale, void* workspace, size_t workspace_bytes, void* stream);
/* Fused cross-attention block core: everything of LeMeBlock.forward_with_c ('C', models/lemevit.py:584-613 with CrossAttention
 * :477-486) or forward_with_xc ('D', :542-582 with DualCrossAttention :252-302) between the positional embedding and the image-token
 * MLP, in three launches (meta_pre -> dca_x -> meta_post; csrc/kernels.h describes the absorbed-projection formulation):
 *   xt [B, N, C] bf16 = x + dwconv(x) (raw) with stats1 [B*N][parts1][2] = partial (sum, sum^2) of its rows (LayerNorm norm1 is
 *   applied from the statistics; its affine is folded into the weights);
 *   'D': xout [B, N, C] <- xt + proj_x(softmax(s_x q1 k2^T) v2) (may alias xt), stats2 [B*N][2] <- (sum, sum^2) of the stored rows;
 *   c [B, 16, C] bf16 updated IN PLACE with the block's whole meta-token update: c += proj(softmax(s_c q k^T) v); c += mlp(norm2(c)).
 * Weights as packed by lemevit_b200/pack.py (bf16 [out, in], LayerNorm affine folded, fp32 biases):
 *   'D': wa = qkv1 [3C, C], wb = qkv2 [3C, C], wp1 = proj_x, wp2 = proj_c;   'C': wa = q [C, C], wb = kv [2C, C], wp1 = proj (wp2 unused);
 *   w1 = mlp.0 [Hd, C] (norm2 folded), w2 = mlp.3 [C, Hd];
 *   wxt = transposed copies of the image-side projections that get absorbed: 'D' [2][C][C] = (wa[0:C]^T, wa[C:2C]^T), 'C' [C][C] = wb[0:C]^T.
 * scale_x / scale_c: softmax scales of the two branches (:235,255-256; 'C': scale_c = head_dim^-0.5, scale_x unused).
 * Requires 16 meta tokens, C = heads * 32 <= 192.  workspace: lmv_dca_workspace_bytes(...) bytes, 256-byte aligned.
 * flags (test hooks) bit 0: issue the c-branch accumulation per 64-channel block; bit 1: one tile at a time even where the pipelined
 * schedule (C <= 96, or 'C' blocks) fits. */
size_t lmv_dca_workspace_bytes(int B, int N, int C, int heads);
int lmv_dca_block(int kind, const void* xt, const float* stats1, int parts1, void* xout, float* stats2, void* c, const void* wa,
                  const float* ba, const void* wb, const float* bb, const void* wxt, const void* wp1, const float* bp1, const void* wp2,
                  const float* bp2, const void* w1, const float* b1, const void* w2, const float* b2, int B, int N, int C, int heads, int Hd,
                  float scale_x, float scale_c, void* workspace, size_t workspace_bytes, int flags, void* stream);
/* patch gather for the first stem conv 3x3/s2/p1 (models/lemevit.py:699): x NCHW (f32|bf16) ->
 * out[B*Ho*Wo, Kp] bf16 with k = ci*9 + ky*3 + kx, zero padded to Kp = round_up(9*Cin, 8); the conv
 * itself (+ folded BN + GELU, :700-701) is then lmv_linear on the tcgen05 GEMM. */
int lmv_stem_im2col(const void* x, int x_dtype, void* out, int B, int Cin, int H, int W, void* stream);
/* First stem convolution without the im2col detour: x NCHW [B, 3, H, W] (f32 | bf16; LMV_DTYPE_U8 / LMV_DTYPE_U8_NHWC pixels are
 * normalised with the ImageNet mean / std, see lmv_plan_set_input_norm for the plan-level knob) -> conv3x3 / stride 2 / pad 1 with the
 * BatchNorm-folded weights w bf16 [C1][Kp = 32] (k = ci*9 + ky*3 + kx, lemevit_b200/pack.py) + bias -> GELU -> out token-major
 * [B, ceil(H/2)*ceil(W/2), C1] bf16 (models/lemevit.py:699-701).  Cin must be 3, C1 32 or 48. */
int lmv_stem_conv1(const void* x, int x_dtype, const void* w, const float* bias, void* out, int B, int Cin, int C1, int H, int W,
                   void* stream);
/* conv 3x3 / stride 2 / pad 1 (+ bias) on token-major bf16 activations in[B, T, Cin] (first H*W rows of every image, row = y*W + x)
 * as an implicit GEMM on the tcgen05 kernel (strided TMA boxes gather the patch rows, nothing is materialised): the second stem
 * convolution and the stage downsample convolutions with their BatchNorm folded (models/lemevit.py:702-703,715-716).
 * w bf16 [Cout][9*Cin] with k = (ky*3 + kx)*Cin + ci, out bf16 token-major; out_rows_per_image: 0 = dense [B*Ho*Wo, Cout], else the
 * rows of image b start at row b * out_rows_per_image (a unified [B, N + M, C] buffer).  Needs ceil(W/2) <= 128 and Cin % 8 == 0. */
int lmv_conv3x3s2(const void* in, const void* w, const float* bias, void* out, int B, int H, int W, int T, int Cin, int Cout,
                  int out_rows_per_image, void* stream);
/* im2col for conv 3x3/s2/p1 on NHWC bf16 tokens [B, T, C] (first H*W rows): out[B*Ho*Wo, 9*C] */
int lmv_im2col_3x3s2(const void* in, void* out, int B, int H, int W, int T, int C, void* stream);
/* classification tail (models/lemevit.py:815-827): feat[b] = bn_scale*mean_n(x[b]) + bn_shift + mean_m(LN(c[b])) */
int lmv_tail(const void* x, long long x_bs, int N, const void* c, long long c_bs, int M, int C, const float* bn_scale,
             const float* bn_shift, const float* ln_gamma, const float* ln_beta, float eps, void* feat, int B,
             void* stream);
/* token-major [B, T, C] bf16 (first H*W rows) -> NCHW [B, C, H, W] (bf16 or f32) */
int lmv_tokens_to_nchw(const void* tokens, void* out, int B, int H, int W, int T, int C, int out_dtype,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEMEVIT_B200_H_ */
