// C ABI (include/lemevit_b200.h): plan = packed-weight table + cached launch schedules; the whole
// forward of LeMeViT (reference models/lemevit.py:809-836 and the backbone copies' forward,
// semantic_segmentation/mmseg/models/backbones/lemevit.py:800-827) is a list of kernel launches on
// the caller's stream, built once per (batch, H, W, workspace) and replayed afterwards — which also
// makes it CUDA-graph capturable by the host.
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace lmv {

// ------------------------------------------------------------------------------------------------
// error reporting / driver entry point
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(LMV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.  A fresh thread (nn.DataParallel
  // runs every replica in its own worker thread) only gets the runtime's primary context bound by its first runtime call.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    LMV_CUDA_OK(cudaFree(nullptr));
    ctx_bound = true;
  }
  cuuint64_t gdims[5], gstrides[5];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstrides[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu] stride %llu box [%u,%u]",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return fail(LMV_ERR_CUDA, buf);
  }
  return LMV_OK;
}

bool pdl_enabled() {
  static bool v = [] { const char* e = getenv("LMV_PDL"); return !(e && e[0] == '0'); }();
  return v;
}

int device_sm_count() {
  static std::atomic<int> cache[PerDeviceOnce::kMaxDevices];   // zero-initialised: 0 = unknown
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool cached = dev >= 0 && dev < PerDeviceOnce::kMaxDevices;
  if (cached && (v = cache[dev].load(std::memory_order_relaxed)) > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  if (cached) cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct BlockW {
  const float *dw_w, *dw_b;
  // 'C': a = q, b = kv, p1 = proj.  'D': a = qkv1, b = qkv2, p1 = proj_x, p2 = proj_c.  'S': a = qkv, p1 = proj.
  const bf16 *wa, *wb, *wp1, *wp2;
  const float *ba, *bb, *bp1, *bp2;
  const bf16 *w1, *w2;
  const float *b1, *b2;
  const float *csa, *csb, *cs1;   // column sums of the LayerNorm-folded weights wa, wb, w1 (gemm.cu LN fold)
  const bf16 *wxqT, *wxkT;        // transposed image-side query ('D') / key ('C', 'D') projections, [C, C] each (fused cross-attention blocks)
};
struct StageW {
  const bf16* ds_w = nullptr;
  const float* ds_b = nullptr;
  const bf16 *md_w0 = nullptr, *md_w3 = nullptr;
  const float *md_b0 = nullptr, *md_g1 = nullptr, *md_be1 = nullptr, *md_b3 = nullptr, *md_g4 = nullptr, *md_be4 = nullptr;
  std::vector<BlockW> blocks;
};

typedef std::function<int(cudaStream_t)> Launch;

enum OpClass { OP_GEMM = 0, OP_MLP, OP_ATTN_SELF, OP_ATTN_TC, OP_ATTN_META, OP_ATTN_SIMT, OP_POSLN, OP_LN, OP_STEM, OP_IM2COL, OP_MISC, OP_DCA, OP_META, OP_MLP_PAIR, OP_NUM_CLASSES };
static const char* const kOpClassNames[OP_NUM_CLASSES] = {"gemm_tcgen05", "mlp_fused_tcgen05", "attention_self_tcgen05", "attention_tcgen05", "attention_meta_tcgen05", "attention_simt",
                                                          "posembed_layernorm", "layernorm", "stem_conv_direct", "im2col", "misc", "dca_fused_tcgen05", "meta_branch", "mlp_pair_tcgen05"};
struct OpRec {
  Launch fn;
  std::string desc;
  int cls;
  double flops;   // algorithmic FLOPs (2 x MAC) of this launch
  double bytes;   // algorithmic bytes (compulsory reads + writes) of this launch
};

struct IoSlots {
  const void* x = nullptr;
  const void* c_in = nullptr;   // caller-supplied meta tokens [B, M, C0] bf16 (forward_features(x, c), models/lemevit.py:809)
  void* feat = nullptr;         // pre-head features [B, C_last] bf16 (nullable)
  void* logits = nullptr;
  void* outs[LMV_MAX_STAGES] = {nullptr};
};

struct Schedule {
  std::vector<OpRec> ops;
  std::unique_ptr<IoSlots> io{new IoSlots()};
  void push(Launch fn, int cls = OP_MISC, double flops = 0, double bytes = 0, std::string desc = std::string()) {
    if (desc.empty()) desc = kOpClassNames[cls];
    ops.push_back(OpRec{std::move(fn), std::move(desc), cls, flops, bytes});
  }
};

struct ProfDetail { double ms = 0, flops = 0, bytes = 0; long long n = 0; };
struct ProfAcc {
  std::map<std::string, ProfDetail> detail;
  double ms[OP_NUM_CLASSES] = {0}, flops[OP_NUM_CLASSES] = {0}, bytes[OP_NUM_CLASSES] = {0};
  long long launches[OP_NUM_CLASSES] = {0};
};

}  // namespace lmv

struct lmv_plan {
  lmv_config cfg;
  const lmv::bf16 *stem1_w, *stem2_w, *c0_init, *head_w;
  const float *stem1_b, *stem2_b, *bn_scale, *bn_shift, *lnc_g, *lnc_b, *head_b;
  std::vector<lmv::StageW> stages;
  int chunk = 0;
  int debug_simt = 0;
  int fused_mlp = 1;
  // also fuse the C = 384 MLP of the 'S' blocks (LMV_FUSED_MLP_WIDE in the environment overrides the default; with LMV_MLP_PAIR=1 the
  // cta_group::2 CTA-pair kernel runs it — DESIGN.md §4)
  int fused_mlp_wide = [] { const char* e = getenv("LMV_FUSED_MLP_WIDE"); return e ? atoi(e) : 1; }();
  int fused_self_attn = 1;
  int fused_dca = 1;
  int dca_pipe = 0;      // pipelined schedule of the fused cross-attention kernel: measured 5-10 % SLOWER than one tile at a time
  int direct_stem = 1;
  int implicit_conv = 1;   // strided convolutions as implicit GEMMs (0: im2col kernel + GEMM)
  int stem_tc = 1;         // first stem convolution on the tcgen05 kernel (0: the CUDA-core direct kernel)
  // 8-bit input path (LMV_DTYPE_U8 / LMV_DTYPE_U8_NHWC): per-channel mean / std in pixel units, lmv_plan_set_input_norm
  float in_mean[3] = {0.485f * 255.f, 0.456f * 255.f, 0.406f * 255.f};
  float in_std[3] = {0.229f * 255.f, 0.224f * 255.f, 0.225f * 255.f};
  int profile = 0;
  int tap_stage = -1, tap_block = -1;       // test hook: copy (x, c) after this block to tap_x / tap_c
  void *tap_x = nullptr, *tap_c = nullptr;
  std::vector<cudaEvent_t> events;          // profile mode: one event between consecutive launches
  std::vector<const lmv::OpRec*> pending;   // ops whose events have not been harvested yet
  lmv::ProfAcc acc;
  std::map<std::tuple<int, int, int, const void*, int, int, int, int>, std::unique_ptr<lmv::Schedule>> cache;
};

namespace lmv {

static inline int kp0(const lmv_config& c) { return ((c.in_chans * 9 + 7) / 8) * 8; }

// walks the packed-tensor order; when `t` is null only counts
struct PackWalker {
  const lmv_tensor* t;
  int n, i = 0;
  std::string err;
  const void* take(int64_t numel, int dtype, const char* what) {
    if (!t) { ++i; return nullptr; }
    if (i >= n) { if (err.empty()) err = std::string("packed weights: missing tensor ") + what; ++i; return nullptr; }
    const lmv_tensor& e = t[i];
    if (err.empty()) {
      char buf[200];
      if (e.numel != numel || e.dtype != dtype) {
        snprintf(buf, sizeof(buf), "packed weights: entry %d (%s) expects numel %lld dtype %d, got numel %lld dtype %d", i,
                 what, (long long)numel, dtype, (long long)e.numel, e.dtype);
        err = buf;
      } else if (!e.data || (reinterpret_cast<uintptr_t>(e.data) & 15)) {
        snprintf(buf, sizeof(buf), "packed weights: entry %d (%s) is null or not 16-byte aligned", i, what);
        err = buf;
      }
    }
    ++i;
    return e.data;
  }
  const bf16* h(int64_t numel, const char* what) { return static_cast<const bf16*>(take(numel, LMV_DTYPE_BF16, what)); }
  const float* f(int64_t numel, const char* what) { return static_cast<const float*>(take(numel, LMV_DTYPE_F32, what)); }
};

static int validate_config(const lmv_config& c) {
  LMV_REQUIRE(c.num_stages >= 1 && c.num_stages <= LMV_MAX_STAGES, "config: num_stages out of range");
  LMV_REQUIRE(c.head_dim == 32, "config: only head_dim == 32 is implemented (all published variants)");
  LMV_REQUIRE(c.queries_len > 0 && c.queries_len % 8 == 0 && c.queries_len <= 128, "config: queries_len must be a multiple of 8, <= 128");
  LMV_REQUIRE(c.in_chans >= 1, "config: in_chans");
  for (int i = 0; i < c.num_stages; ++i) {
    // reference error conventions: AssertionError on dim % num_heads (models/lemevit.py:168,233,437)
    LMV_REQUIRE(c.embed_dim[i] > 0 && c.embed_dim[i] % c.head_dim == 0, "config: dim not divisible by num_heads");
    LMV_REQUIRE(c.embed_dim[i] <= 512, "config: embed_dim > 512 not implemented");
    LMV_REQUIRE(c.mlp_hidden[i] > 0 && c.mlp_hidden[i] % 8 == 0, "config: mlp hidden must be a multiple of 8");
    LMV_REQUIRE(c.depth[i] >= 0, "config: depth");
    if (!(c.attn_type[i] == 'C' || c.attn_type[i] == 'D' || c.attn_type[i] == 'S'))
      return fail(LMV_ERR_UNSUPPORTED, "Attention type does not exit");  // reference message models/lemevit.py:660
  }
  LMV_REQUIRE(c.embed_dim[0] % 16 == 0, "config: embed_dim[0] must be a multiple of 16");
  return LMV_OK;
}

static void walk(const lmv_config& c, PackWalker& w, lmv_plan* plan) {
  const int C0 = c.embed_dim[0], M = c.queries_len;
  const bf16* p;
  const float* q;
  p = w.h((int64_t)(C0 / 2) * kp0(c), "stem1_w"); if (plan) plan->stem1_w = p;
  q = w.f(C0 / 2, "stem1_b"); if (plan) plan->stem1_b = q;
  p = w.h((int64_t)C0 * 9 * (C0 / 2), "stem2_w"); if (plan) plan->stem2_w = p;
  q = w.f(C0, "stem2_b"); if (plan) plan->stem2_b = q;
  p = w.h((int64_t)M * C0, "c0_init"); if (plan) plan->c0_init = p;
  if (plan) plan->stages.resize(c.num_stages);
  for (int i = 0; i < c.num_stages; ++i) {
    StageW tmp;
    StageW& s = plan ? plan->stages[i] : tmp;
    const int C = c.embed_dim[i], Hd = c.mlp_hidden[i];
    {
      // meta_token_downsample[i] (models/lemevit.py:729-745); [0] maps C0 -> C0 and is only run when the caller supplies
      // its own meta tokens (otherwise its output on `meta_tokens` is the pack-time constant c0_init)
      const int Cp = c.embed_dim[i > 0 ? i - 1 : 0];
      if (i > 0 && c.attn_type[i - 1] != 'C') {
        s.ds_w = w.h((int64_t)C * 9 * Cp, "ds_w");
        s.ds_b = w.f(C, "ds_b");
      }
      s.md_w0 = w.h((int64_t)4 * Cp * Cp, "md_w0");
      s.md_b0 = w.f(4 * Cp, "md_b0");
      s.md_g1 = w.f(4 * Cp, "md_g1");
      s.md_be1 = w.f(4 * Cp, "md_be1");
      s.md_w3 = w.h((int64_t)C * 4 * Cp, "md_w3");
      s.md_b3 = w.f(C, "md_b3");
      s.md_g4 = w.f(C, "md_g4");
      s.md_be4 = w.f(C, "md_be4");
    }
    s.blocks.resize(c.depth[i]);
    for (int j = 0; j < c.depth[i]; ++j) {
      BlockW& b = s.blocks[j];
      memset(&b, 0, sizeof(b));
      b.dw_w = w.f(9 * C, "dw_w");
      b.dw_b = w.f(C, "dw_b");
      switch (c.attn_type[i]) {
        case 'C':
          b.wa = w.h((int64_t)C * C, "q_w"); b.ba = w.f(C, "q_b"); b.csa = w.f(C, "q_colsum");
          b.wb = w.h((int64_t)2 * C * C, "kv_w"); b.bb = w.f(2 * C, "kv_b"); b.csb = w.f(2 * C, "kv_colsum");
          b.wxkT = w.h((int64_t)C * C, "kv_kT");
          b.wp1 = w.h((int64_t)C * C, "proj_w"); b.bp1 = w.f(C, "proj_b");
          break;
        case 'D':
          b.wa = w.h((int64_t)3 * C * C, "qkv1_w"); b.ba = w.f(3 * C, "qkv1_b"); b.csa = w.f(3 * C, "qkv1_colsum");
          b.wxqT = w.h((int64_t)C * C, "qkv1_qT"); b.wxkT = w.h((int64_t)C * C, "qkv1_kT");
          b.wb = w.h((int64_t)3 * C * C, "qkv2_w"); b.bb = w.f(3 * C, "qkv2_b"); b.csb = w.f(3 * C, "qkv2_colsum");
          b.wp1 = w.h((int64_t)C * C, "proj_x_w"); b.bp1 = w.f(C, "proj_x_b");
          b.wp2 = w.h((int64_t)C * C, "proj_c_w"); b.bp2 = w.f(C, "proj_c_b");
          break;
        default:
          b.wa = w.h((int64_t)3 * C * C, "qkv_w"); b.ba = w.f(3 * C, "qkv_b"); b.csa = w.f(3 * C, "qkv_colsum");
          b.wp1 = w.h((int64_t)C * C, "proj_w"); b.bp1 = w.f(C, "proj_b");
      }
      b.w1 = w.h((int64_t)Hd * C, "mlp0_w"); b.b1 = w.f(Hd, "mlp0_b"); b.cs1 = w.f(Hd, "mlp0_colsum");
      b.w2 = w.h((int64_t)C * Hd, "mlp3_w"); b.b2 = w.f(C, "mlp3_b");
    }
  }
  if (!c.backbone) {
    const int CL = c.embed_dim[c.num_stages - 1];
    q = w.f(CL, "bn_scale"); if (plan) plan->bn_scale = q;
    q = w.f(CL, "bn_shift"); if (plan) plan->bn_shift = q;
    q = w.f(CL, "norm_c_g"); if (plan) plan->lnc_g = q;
    q = w.f(CL, "norm_c_b"); if (plan) plan->lnc_b = q;
    if (c.num_classes > 0) {
      p = w.h((int64_t)c.num_classes * CL, "head_w"); if (plan) plan->head_w = p;
      q = w.f(c.num_classes, "head_b"); if (plan) plan->head_b = q;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// geometry + workspace
// ------------------------------------------------------------------------------------------------
struct Geo {
  int H[LMV_MAX_STAGES], W[LMV_MAX_STAGES], N[LMV_MAX_STAGES], T[LMV_MAX_STAGES];
  bool unified[LMV_MAX_STAGES], c_alive[LMV_MAX_STAGES];
  int H1, W1;  // after the first stem conv
};

static int geometry(const lmv_config& c, int H, int W, Geo* g) {
  LMV_REQUIRE(H >= 4 && W >= 4, "input smaller than 4x4");
  g->H1 = (H + 1) / 2; g->W1 = (W + 1) / 2;
  int h = (g->H1 + 1) / 2, w = (g->W1 + 1) / 2;
  // meta tokens matter at stage i iff some stage >= i consumes them
  // (backbone copies: 'S' blocks leave c untouched, so c is dead after the last C/D stage)
  bool alive = !c.backbone;
  for (int i = c.num_stages - 1; i >= 0; --i) {
    if (c.attn_type[i] != 'S') alive = true;
    g->c_alive[i] = alive;
  }
  for (int i = 0; i < c.num_stages; ++i) {
    if (i > 0 && c.attn_type[i - 1] != 'C') { h = (h + 1) / 2; w = (w + 1) / 2; }
    g->H[i] = h; g->W[i] = w; g->N[i] = h * w;
    g->unified[i] = (c.attn_type[i] == 'S') && !c.backbone;
    g->T[i] = g->N[i] + (g->unified[i] ? c.queries_len : 0);
    if (i > 0 && c.attn_type[i - 1] == 'C' && g->unified[i])
      return fail(LMV_ERR_UNSUPPORTED, "an 'S' stage directly after a 'C' stage is not implemented");
  }
  return LMV_OK;
}

struct WsLayout {
  size_t patches, stem1, x0, x1, xn, qkv, hid, ca, cb, cn, cqkv, chid, ctmp, feat, stats1, stats2, cpart, cpart_bytes, dca, total;
};

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

static void ws_layout(const lmv_config& c, const Geo& g, int B, WsLayout* L) {
  const int C0 = c.embed_dim[0], M = c.queries_len;
  size_t patches = (size_t)B * g.H1 * g.W1 * kp0(c);
  patches = std::max(patches, (size_t)B * g.N[0] * 9 * (C0 / 2));
  size_t x = 0, qkv = 0, hid = 0, cmax = 0, chid = 0;
  for (int i = 0; i < c.num_stages; ++i) {
    const size_t rows = (size_t)B * g.T[i];
    x = std::max(x, rows * c.embed_dim[i]);
    qkv = std::max(qkv, rows * 3 * c.embed_dim[i]);
    hid = std::max(hid, rows * c.mlp_hidden[i]);
    cmax = std::max(cmax, (size_t)c.embed_dim[i]);
    chid = std::max(chid, (size_t)c.mlp_hidden[i]);
    if (i > 0) {
      chid = std::max(chid, (size_t)4 * c.embed_dim[i - 1]);
      if (c.attn_type[i - 1] != 'C') patches = std::max(patches, (size_t)B * g.N[i] * 9 * c.embed_dim[i - 1]);
    }
  }
  size_t off = 0;
  auto take = [&](size_t elems) { size_t o = off; off += align_up(elems * 2); return o; };
  L->patches = take(patches);
  L->stem1 = take((size_t)B * g.H1 * g.W1 * (C0 / 2));
  L->x0 = take(x); L->x1 = take(x); L->xn = take(x);
  L->qkv = take(qkv); L->hid = take(hid);
  const size_t crow = (size_t)B * M;
  L->ca = take(crow * cmax); L->cb = take(crow * cmax); L->cn = take(crow * cmax); L->ctmp = take(crow * cmax);
  L->cqkv = take(crow * 3 * cmax); L->chid = take(crow * chid);
  L->feat = take((size_t)B * cmax);
  size_t rows_max = 0;
  for (int i = 0; i < c.num_stages; ++i) rows_max = std::max(rows_max, (size_t)B * g.T[i]);
  L->stats1 = take(rows_max * 16);   // [rows][parts <= 4][2] fp32 (take() counts 2-byte elements)
  L->stats2 = take(rows_max * 16);   // [rows][parts <= 4][2] fp32
  // split-softmax partials of the meta-token attention (attention_meta.cu): a few partial rows per image and (head, query)
  size_t cpart = 0;
  for (int i = 0; i < c.num_stages; ++i)
    if (c.attn_type[i] != 'S')
    {
      AttnArgs ma{};
      ma.B = B; ma.heads = c.embed_dim[i] / c.head_dim; ma.Lq = M; ma.Lk = g.N[i];
      cpart = std::max(cpart, attention_meta_workspace(ma));
    }
  for (int i = 0; i < c.num_stages; ++i)
    if (c.attn_type[i] == 'S') cpart = std::max(cpart, attention_self_workspace(B, c.embed_dim[i] / c.head_dim, g.T[i]));
  L->cpart_bytes = cpart;
  L->cpart = take((cpart + 1) / 2);
  // scratch of the fused cross-attention blocks (per-image operands + c-branch softmax partials)
  size_t dca = 0;
  for (int i = 0; i < c.num_stages; ++i)
    if (c.attn_type[i] != 'S' && dca_supported(g.N[i], c.embed_dim[i], c.embed_dim[i] / c.head_dim, M))
      dca = std::max(dca, dca_workspace_bytes(dca_geometry(B, g.N[i], c.embed_dim[i], c.embed_dim[i] / c.head_dim)));
  L->dca = take((dca + 1) / 2);
  L->total = off;
}

// ------------------------------------------------------------------------------------------------
// schedule builder
// ------------------------------------------------------------------------------------------------
struct Builder {
  lmv_plan* plan;
  Schedule* sc;
  int rc = LMV_OK;
  bool simt;
  void* cpart = nullptr;      // split-softmax partials of the meta-token attention
  size_t cpart_bytes = 0;
  // meta-token update of the last fused cross-attention block, not launched yet: it is chained with the operand build (meta_pre) of
  // the next block of the stage into ONE launch (meta_chain_run); flushed on its own at the end of a stage
  bool has_post = false;
  MetaPostArgs pending_post;
  double pending_post_flops = 0;

  void flush_meta(const MetaPreArgs* pre = nullptr, double pre_flops = 0) {
    if (rc || (!has_post && !pre)) return;
    const MetaPostArgs post = pending_post;
    const bool chain = has_post && pre && post.c == pre->c && post.C == pre->C && post.B == pre->B;
    char d[160];
    if (has_post && !chain) {
      snprintf(d, sizeof(d), "meta_post B=%d C=%d", post.B, post.C);
      sc->push([post](cudaStream_t s) { return meta_chain_run(&post, nullptr, s); }, OP_META, pending_post_flops, 0.0, d);
    }
    if (pre) {
      const MetaPreArgs p2 = *pre;
      if (chain) {
        snprintf(d, sizeof(d), "meta_post+pre B=%d C=%d", post.B, post.C);
        sc->push([post, p2](cudaStream_t s) { return meta_chain_run(&post, &p2, s); }, OP_META, pending_post_flops + pre_flops, 0.0, d);
      } else {
        snprintf(d, sizeof(d), "meta_pre B=%d C=%d", p2.B, p2.C);
        sc->push([p2](cudaStream_t s) { return meta_chain_run(nullptr, &p2, s); }, OP_META, pre_flops, 0.0, d);
      }
    }
    has_post = false;
  }

  void gemm(GemmArgs a) {
    if (rc) return;
    const double fl = 2.0 * a.M * a.N * a.K;
    const double by = 2.0 * ((double)a.M * a.K + (double)a.N * a.K + (double)a.M * a.N * (a.residual ? 2 : 1));
    if (simt) {
      sc->push([a](cudaStream_t s) { return gemm_simt_run(a, s); }, OP_MISC, fl, by);
      return;
    }
    GemmOp op;
    rc = gemm_prepare(a, &op);
    if (rc) return;
    char d[160];
    snprintf(d, sizeof(d), "gemm M=%d N=%d K=%d BN=%d%s%s%s%s%s", a.M, a.N, a.K, op.p.BN, a.ln_stats ? " ln" : "", a.act ? " gelu" : "",
             a.residual ? " res" : "", a.stats_out ? " stats" : "", a.conv_C > 0 ? " conv3x3s2" : "");
    sc->push([op](cudaStream_t s) { return gemm_run(op, s); }, OP_GEMM, fl, by, d);
  }
  void linear(const bf16* A, int lda, const bf16* Wt, const float* bias, int M, int N, int K, bf16* out, int ldc,
              int gelu = 0, const bf16* resid = nullptr, int grp_rows = 0, int grp_stride = 0) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.W = Wt; a.ldw = K; a.M = M; a.N = N; a.K = K;
    a.bias = bias; a.act = gelu; a.residual = resid; a.out = out; a.ldc = ldc;
    a.grp_rows = grp_rows; a.grp_stride = grp_stride;
    gemm(a);
  }
  // conv3x3 / stride 2 / pad 1 (+ folded BatchNorm) on token-major activations in [B, T, Cin] (first H * W rows per image) as an
  // implicit GEMM: the A tiles are gathered by strided TMA boxes straight from the activation (gemm.cu), no patch matrix.
  // Falls back to im2col + GEMM (patches: scratch for B * Ho * Wo * 9 * Cin bf16) where the implicit form does not apply.
  void conv3x3s2(const bf16* in, int B, int H, int W, int T, int Cin, const bf16* Wt, const float* bias, int Cout, bf16* out,
                 int grp_rows, int grp_stride, bf16* patches) {
    if (rc) return;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    if (!simt && plan->implicit_conv && gemm_conv_supported(H, W, Cin)) {
      GemmArgs a;
      a.A = in; a.lda = 0; a.W = Wt; a.ldw = 9 * Cin; a.M = B * Ho * Wo; a.N = Cout; a.K = 9 * Cin;
      a.bias = bias; a.out = out; a.ldc = Cout; a.grp_rows = grp_rows; a.grp_stride = grp_stride;
      a.conv_B = B; a.conv_H = H; a.conv_W = W; a.conv_T = T; a.conv_C = Cin;
      gemm(a);
      return;
    }
    Im2colArgs ia{in, patches, B, H, W, T, Cin};
    sc->push([ia](cudaStream_t s) { return im2col_run(ia, s); }, OP_IM2COL, 0.0, 2.0 * B * Cin * ((double)H * W + 9.0 * Ho * Wo));
    linear(patches, 9 * Cin, Wt, bias, B * Ho * Wo, Cout, 9 * Cin, out, Cout, 0, nullptr, grp_rows, grp_stride);
  }
  // LayerNorm(eps 1e-6, affine folded into Wt/bias at pack time) -> Linear, the norm folded into the epilogue:
  // A holds the raw rows, ln_stats their (sum, sum^2), colsum the column sums of Wt.
  void ln_linear(const bf16* A, const float* ln_stats, int ln_parts, const bf16* Wt, const float* bias, const float* colsum,
                 int M, int N, int K, bf16* out, int gelu = 0) {
    GemmArgs a;
    a.A = A; a.lda = K; a.W = Wt; a.ldw = K; a.M = M; a.N = N; a.K = K;
    a.bias = bias; a.act = gelu; a.out = out; a.ldc = N;
    a.ln_stats = ln_stats; a.ln_parts = ln_parts; a.ln_colsum = colsum; a.ln_eps = 1e-6f;
    gemm(a);
  }
  // Linear + residual (in place on x) that also emits the LayerNorm statistics of the new x rows as
  // `*parts` partial (sum, sum^2) pairs per row (deterministic plain stores; the consumer adds them up)
  void linear_res_stats(const bf16* A, const bf16* Wt, const float* bias, int M, int N, int K, bf16* x, float* stats, int* parts) {
    if (rc) return;
    GemmArgs a;
    a.A = A; a.lda = K; a.W = Wt; a.ldw = K; a.M = M; a.N = N; a.K = K;
    a.bias = bias; a.residual = x; a.out = x; a.ldc = N;
    if (simt) {
      gemm(a);
      sc->push([x, stats, M, N](cudaStream_t s) { return row_stats_run(x, stats, M, N, s); }, OP_MISC, 0.0, 2.0 * M * N, "row_stats");
      *parts = 1;
      return;
    }
    *parts = gemm_stats_parts(N);
    if (*parts > 4) { rc = fail(LMV_ERR_UNSUPPORTED, "linear_res_stats: more than 4 statistics partials per row"); return; }
    a.stats_out = stats;
    gemm(a);
  }
  // x <- x + mlp(norm2(x)) (models/lemevit.py:562,564,601,633,635): one fused kernel when the shape allows it,
  // otherwise fc1 (+LN fold, bias, GELU) and fc2 (+bias, residual) on the GEMM with the hidden activation in `hid`
  void mlp(bf16* x, const float* stats, int parts, const BlockW& bw, int R, int C, int Hd, bf16* hid) {
    if (rc) return;
    if (!simt && plan->fused_mlp && mlp_fused_supported(C, Hd) && (C <= 256 || plan->fused_mlp_wide)) {
      MlpArgs a;
      a.x = x; a.out = x; a.W1 = bw.w1; a.b1 = bw.b1; a.cs1 = bw.cs1; a.W2 = bw.w2; a.b2 = bw.b2;
      a.ln_stats = stats; a.ln_parts = parts; a.ln_eps = 1e-6f; a.R = R; a.C = C; a.Hd = Hd;
      MlpOp op;
      rc = mlp_fused_prepare(a, &op);
      if (rc) return;
      char d[120];
      snprintf(d, sizeof(d), "mlp_fused R=%d C=%d Hd=%d", R, C, Hd);
      sc->push([op](cudaStream_t s) { return mlp_fused_run(op, s); }, op.p.pair ? OP_MLP_PAIR : OP_MLP, 4.0 * R * C * (double)Hd, 4.0 * R * C, d);
      return;
    }
    ln_linear(x, stats, parts, bw.w1, bw.b1, bw.cs1, R, Hd, C, hid, 1);
    linear(hid, Hd, bw.w2, bw.b2, R, C, Hd, x, C, 0, x);
  }
  // StandardAttention of an 'S' block on a [B, T, 3C] qkv buffer: image tokens (rows < N) and meta tokens (rows >= N) attend
  // within their own segment (models/lemevit.py:632-635 run the same attention module on x and on c)
  void self_attn(const bf16* qkv, bf16* out, int B, int heads, int T, int N, int C, float scale) {
    if (rc) return;
    AttnArgs a;
    a.q = qkv; a.k = qkv + C; a.v = qkv + 2 * C; a.out = out;
    a.q_bs = a.k_bs = a.v_bs = (long long)T * 3 * C; a.o_bs = (long long)T * C;
    a.q_rs = a.k_rs = a.v_rs = 3 * C; a.o_rs = C;
    a.B = B; a.heads = heads; a.Lq = T; a.Lk = T; a.scale = scale;
    if (!simt && plan->fused_self_attn && attention_self_supported(a, T, N) && attention_self_workspace(B, heads, T) <= cpart_bytes) {
      void* wsp = cpart;
      const size_t wsb = cpart_bytes;
      const double fl = 4.0 * B * heads * 32.0 * ((double)N * N + (double)(T - N) * (T - N));
      const double by = 2.0 * B * T * 4.0 * C;
      sc->push([a, T, N, wsp, wsb](cudaStream_t s) { return attention_self_run(a, T, N, wsp, wsb, s); }, OP_ATTN_SELF, fl, by,
               "attn_self B=" + std::to_string(B) + " h=" + std::to_string(heads) + " T=" + std::to_string(T) + " N=" + std::to_string(N));
      return;
    }
    attn(qkv, (long long)T * 3 * C, 3 * C, qkv + C, qkv + 2 * C, (long long)T * 3 * C, 3 * C, out, (long long)T * C, C, B, heads, N, N, scale);
    if (T > N) {
      const size_t ro = (size_t)N * 3 * C;
      attn(qkv + ro, (long long)T * 3 * C, 3 * C, qkv + ro + C, qkv + ro + 2 * C, (long long)T * 3 * C, 3 * C, out + (size_t)N * C,
           (long long)T * C, C, B, heads, T - N, T - N, scale);
    }
  }
  // returns the number of statistics partials per row written to `stats` ([B*T][parts][2], parts <= 4)
  int posln(const bf16* tok, const float* dw_w, const float* dw_b, bf16* resid, bf16* norm, int B, int H, int W, int T,
            int C, float* stats = nullptr) {
    if (rc) return 1;
    PosLnArgs a{tok, dw_w, dw_b, resid, norm, B, H, W, T, C, 1e-6f, stats};
    a.max_parts = 4;
    const double rows = (double)B * T;
    char d[160];
    snprintf(d, sizeof(d), "posln rows=%d C=%d conv=%d resid=%d norm=%d", B * T, C, dw_w ? 1 : 0, resid ? 1 : 0, norm ? 1 : 0);
    const double fl = dw_w ? 18.0 * B * H * W * C : 0.0, by = 2.0 * rows * C * (1 + (resid ? 1 : 0) + (norm ? 1 : 0));
    if (posembed_tile_supported(a)) {
      PosEmbedOp op;
      rc = posembed_tile_prepare(a, &op);
      if (rc) return 1;
      sc->push([op](cudaStream_t s) { return posembed_tile_run(op, s); }, OP_POSLN, fl, by, d);
      return op.parts;
    }
    a.max_parts = 1;
    sc->push([a](cudaStream_t s) { return posembed_ln_run(a, s); }, OP_POSLN, fl, by, d);
    return 1;
  }
  void ln(const bf16* in, bf16* out, const float* g, const float* b, int R, int C, float eps, int gelu = 0,
          int grp_rows = 0, int grp_stride = 0, int grp_off = 0) {
    if (rc) return;
    LnArgs a{in, out, g, b, R, C, eps, gelu, grp_rows, grp_stride, grp_off};
    sc->push([a](cudaStream_t s) { return layernorm_run(a, s); }, OP_LN, 0.0, 4.0 * R * C);
  }
  // Fused 'C' / 'D' block core (kernels.h): meta_pre -> dca_x -> meta_post.  xt = x + dw(x) with its LayerNorm statistics in stats1;
  // 'D': x <- xt + proj_x(attention) written to xout (may alias xt) with the LN2 statistics in stats2; c updated in place (full
  // meta-token update of the block including its MLP).
  void dca_block(char kind, const bf16* xt, const float* stats1, int parts1, bf16* xout, float* stats2, bf16* cc, const BlockW& bw,
                 int B, int N, int C, int heads, int Hd, float scale_x, float scale_c, void* dca_ws, int zsplit = 0) {
    if (rc) return;
    const DcaGeom g = dca_geometry(B, N, C, heads);
    const DcaWs w = dca_workspace_carve(g, dca_ws);
    const bool D = kind == 'D';
    MetaPreArgs pre{};
    pre.c = cc; pre.B = B; pre.C = C; pre.heads = heads; pre.eps = 1e-6f; pre.ws = w;
    pre.scale_x = scale_x; pre.scale_c = scale_c;
    MetaPostArgs post{};
    post.c = cc; post.B = B; post.C = C; post.heads = heads; post.Hd = Hd; post.parts = g.parts; post.eps = 1e-6f; post.ws = w;
    post.W1 = bw.w1; post.b1 = bw.b1; post.W2 = bw.w2; post.b2 = bw.b2;
    if (D) {
      pre.Wc = bw.wb; pre.bc = bw.bb; pre.nc = 3 * C; pre.q_off = 0; pre.k_off = C; pre.v_off = 2 * C;
      pre.WxqT = bw.wxqT; pre.bxq = bw.ba; pre.WxkT = bw.wxkT; pre.bxk = bw.ba + C; pre.Wpx = bw.wp1;
      post.Wxv = bw.wa + (size_t)2 * C * C; post.bxv = bw.ba + 2 * C; post.Wp = bw.wp2; post.bp = bw.bp2;
    } else {
      pre.Wc = bw.wa; pre.bc = bw.ba; pre.nc = C; pre.q_off = 0; pre.k_off = -1; pre.v_off = -1;
      pre.WxkT = bw.wxkT; pre.bxk = bw.bb;
      post.Wxv = bw.wb + (size_t)C * C; post.bxv = bw.bb + C; post.Wp = bw.wp1; post.bp = bw.bp1;
    }
    DcaXArgs xa{};
    xa.xt = xt; xa.stats1 = stats1; xa.parts1 = parts1; xa.eps = 1e-6f; xa.do_x = D ? 1 : 0;
    xa.bpx = D ? bw.bp1 : nullptr; xa.xout = D ? xout : nullptr; xa.stats2 = D ? stats2 : nullptr;
    xa.g = g; xa.ws = w; xa.zsplit = zsplit & 1; xa.force_serial = (zsplit >> 1) & 1;
    DcaXOp op;
    rc = dca_x_prepare(xa, &op);
    if (rc) return;
    const double rows = (double)B * 16, R = g.R;
    char d[160];
    flush_meta(&pre, 2.0 * rows * C * (pre.nc + (D ? 3.0 : 1.0) * C));
    // algorithmic FLOPs of what the kernel replaces (SURVEY.md section 8d accounting): the image-side projections + both attentions
    const double fl = 2.0 * B * N * ((D ? 4.0 : 2.0) * C * C + (D ? 2.0 : 1.0) * 2.0 * 16.0 * C);
    snprintf(d, sizeof(d), "dca_x B=%d N=%d C=%d %c", B, N, C, kind);
    sc->push([op](cudaStream_t s) { return dca_x_run(op, s); }, OP_DCA, fl, 2.0 * B * N * C * (D ? 2.0 : 1.0), d);
    pending_post = post;
    pending_post_flops = 2.0 * rows * C * (2.0 * C + 2.0 * Hd) + 2.0 * B * R * C * g.parts;
    has_post = true;
  }
  void attn(const bf16* q, long long q_bs, int q_rs, const bf16* k, const bf16* v, long long kv_bs, int kv_rs, bf16* out,
            long long o_bs, int o_rs, int B, int heads, int Lq, int Lk, float scale) {
    if (rc) return;
    AttnArgs a;
    a.q = q; a.k = k; a.v = v; a.out = out;
    a.q_bs = q_bs; a.k_bs = kv_bs; a.v_bs = kv_bs; a.o_bs = o_bs;
    a.q_rs = q_rs; a.k_rs = kv_rs; a.v_rs = kv_rs; a.o_rs = o_rs;
    a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
    const double fl = 4.0 * B * heads * (double)Lq * Lk * 32;
    const double by = 2.0 * B * heads * 32.0 * (2.0 * Lq + 2.0 * Lk);
    if (!simt && Lk >= 8 * Lq && Lk > 224 && attention_meta_supported(a) && cpart && attention_meta_workspace(a) <= cpart_bytes) {
      void* wsp = cpart;
      const size_t wsb = cpart_bytes;
      sc->push([a, wsp, wsb](cudaStream_t s) { return attention_meta_run(a, wsp, wsb, s); }, OP_ATTN_META, fl, by,
               "attn_meta B=" + std::to_string(B) + " h=" + std::to_string(heads) + " Lq=" + std::to_string(Lq) + " Lk=" + std::to_string(Lk));
      return;
    }
    const bool tc = !simt && attention_tc_supported(a);
    if (!tc && !simt) {
      // no silent drop to the SIMT cross-check kernel (100x slower): an attention shape outside the tcgen05 kernels is an error
      rc = fail(LMV_ERR_UNSUPPORTED, "attention: no tcgen05 kernel covers B=" + std::to_string(B) + " heads=" + std::to_string(heads) + " Lq=" +
                                         std::to_string(Lq) + " Lk=" + std::to_string(Lk) + " (lmv_plan_set_debug_simt(1) runs it on the SIMT cross-check kernel)");
      return;
    }
    sc->push([a, tc](cudaStream_t s) { return tc ? attention_tc_run(a, s) : attention_simt_run(a, s); },
             tc ? OP_ATTN_TC : OP_ATTN_SIMT, fl, by,
             std::string(tc ? "attn_tc" : "attn_simt") + " B=" + std::to_string(B) + " h=" + std::to_string(heads) + " Lq=" +
                 std::to_string(Lq) + " Lk=" + std::to_string(Lk));
  }
};

// schedule flags (part of the cache key)
constexpr int kFlagCustomC = 1;   // meta tokens come from io->c_in and go through meta_token_downsample[0] at run time
constexpr int kFlagFeat = 2;      // copy the pre-head features to io->feat
constexpr int kFlagLogits = 4;    // run the classifier head into io->logits

static int build_schedule(lmv_plan* plan, int B, int H, int W, uint8_t* ws, int x_dtype, int out_dtype, int flags, Schedule* sc) {
  const lmv_config& c = plan->cfg;
  Geo g;
  int rc = geometry(c, H, W, &g);
  if (rc) return rc;
  WsLayout L;
  ws_layout(c, g, B, &L);
  auto P = [&](size_t off) { return reinterpret_cast<bf16*>(ws + off); };
  bf16 *patches = P(L.patches), *stem1 = P(L.stem1), *xn = P(L.xn), *qkv = P(L.qkv), *hid = P(L.hid);
  bf16* xbuf[2] = {P(L.x0), P(L.x1)};
  bf16* cbuf[2] = {P(L.ca), P(L.cb)};
  bf16 *cn = P(L.cn), *ctmp = P(L.ctmp), *cqkv = P(L.cqkv), *chid = P(L.chid), *feat = P(L.feat);
  float *stats1 = reinterpret_cast<float*>(ws + L.stats1), *stats2 = reinterpret_cast<float*>(ws + L.stats2);
  IoSlots* io = sc->io.get();
  Builder b{plan, sc, LMV_OK, plan->debug_simt != 0};
  b.cpart = ws + L.cpart;
  b.cpart_bytes = L.cpart_bytes;
  const int M = c.queries_len, C0 = c.embed_dim[0], S = c.num_stages;

  // ---- stem (models/lemevit.py:698-704): conv3x3/s2 + BN + GELU + conv3x3/s2 + BN, both on the GEMM
  {
    const bool x_u8 = x_dtype == LMV_DTYPE_U8 || x_dtype == LMV_DTYPE_U8_NHWC;
    const double x_bytes = (double)B * c.in_chans * H * W * (x_dtype == LMV_DTYPE_F32 ? 4 : x_u8 ? 1 : 2);
    if (x_u8 && (b.simt || !plan->direct_stem || !stem_conv1_supported(c.in_chans, C0 / 2)))
      return fail(LMV_ERR_UNSUPPORTED, "8-bit input needs the direct stem kernel (3 input channels, embed_dim[0] / 2 in {32, 48}, direct_stem = 1)");
    if (!b.simt && plan->direct_stem && stem_conv1_supported(c.in_chans, C0 / 2)) {
      StemArgs sa{nullptr, x_dtype, stem1, B, c.in_chans, H, W};
      for (int i = 0; i < 3; ++i) { sa.mean[i] = plan->in_mean[i]; sa.std[i] = plan->in_std[i]; }
      sa.tensor_core = plan->stem_tc;
      const bf16* w1 = plan->stem1_w;
      const float* b1 = plan->stem1_b;
      const int C1 = C0 / 2;
      sc->push([sa, io, w1, b1, C1](cudaStream_t s) { StemArgs a = sa; a.x = io->x; return stem_conv1_run(a, w1, b1, C1, s); }, OP_STEM,
               2.0 * B * g.H1 * g.W1 * C1 * 27.0, x_bytes + 2.0 * B * g.H1 * g.W1 * C1, "stem_conv1_direct");
    } else {
      StemArgs sa{nullptr, x_dtype, patches, B, c.in_chans, H, W};
      sc->push([sa, io](cudaStream_t s) { StemArgs a = sa; a.x = io->x; return stem_im2col_run(a, s); }, OP_IM2COL, 0.0,
               x_bytes + 2.0 * B * g.H1 * g.W1 * kp0(c));
      b.linear(patches, kp0(c), plan->stem1_w, plan->stem1_b, B * g.H1 * g.W1, C0 / 2, kp0(c), stem1, C0 / 2, /*gelu=*/1);
    }
  }
  int cur = 0, ccur = 0;
  {
    const bool uni = g.unified[0];
    b.conv3x3s2(stem1, B, g.H1, g.W1, g.H1 * g.W1, C0 / 2, plan->stem2_w, plan->stem2_b, C0, xbuf[cur], uni ? g.N[0] : 0, uni ? g.T[0] : 0,
                patches);
  }

  for (int i = 0; i < S && !b.rc; ++i) {
    const StageW& sw = plan->stages[i];
    const int C = c.embed_dim[i], Hd = c.mlp_hidden[i], N = g.N[i], T = g.T[i], heads = C / c.head_dim;
    const bool uni = g.unified[i];
    const char kind = c.attn_type[i];
    // ---- x downsample (models/lemevit.py:711-717)
    const bf16* c_prev_unified = nullptr;  // where the previous stage left c if it was unified
    if (i > 0) {
      const int Cp = c.embed_dim[i - 1];
      if (g.unified[i - 1]) c_prev_unified = xbuf[cur];
      if (c.attn_type[i - 1] != 'C') {
        b.conv3x3s2(xbuf[cur], B, g.H[i - 1], g.W[i - 1], g.T[i - 1], Cp, sw.ds_w, sw.ds_b, C, xbuf[cur ^ 1], uni ? N : 0, uni ? T : 0,
                    patches);
        cur ^= 1;
      }
    }
    // ---- meta-token path into this stage (models/lemevit.py:729-745, :833)
    if (g.c_alive[i]) {
      if (i == 0 && !(flags & kFlagCustomC)) {
        // meta_ds_0(meta_tokens) is batch-invariant: folded at pack time into c0_init
        bf16* dst = uni ? xbuf[cur] + (size_t)N * C : cbuf[ccur];
        const long long bs = uni ? (long long)T * C : (long long)M * C;
        const bf16* src = plan->c0_init;
        sc->push([src, dst, M, C, B, bs](cudaStream_t s) { return broadcast_rows_run(src, dst, M, C, B, bs, s); });
      } else {
        const int Cp = c.embed_dim[i > 0 ? i - 1 : 0];
        const bf16* cprev = cbuf[ccur];
        if (i == 0) {
          // forward_features(x, c) with the caller's own meta tokens (models/lemevit.py:809-813): stage them next to the
          // other meta-token buffers (the GEMM's tensor map is encoded once per schedule, the caller's pointer changes)
          bf16* dst = cbuf[ccur];
          const size_t bytes = (size_t)B * M * Cp * sizeof(bf16);
          sc->push([io, dst, bytes](cudaStream_t s) {
            if (!io->c_in) return fail(LMV_ERR_INVALID, "forward: this schedule expects caller-supplied meta tokens");
            LMV_CUDA_OK(cudaMemcpyAsync(dst, io->c_in, bytes, cudaMemcpyDeviceToDevice, s));
            return LMV_OK;
          });
        }
        if (!b.simt && M == kDcaM && meta_downsample_supported(Cp, C)) {
          // one kernel per stage transition (meta_branch.cu) instead of gather + linear + LN/GELU + linear + LN
          MetaDsArgs da{};
          if (c_prev_unified) { da.in = c_prev_unified + (size_t)g.N[i - 1] * Cp; da.in_bs = (long long)g.T[i - 1] * Cp; }
          else { da.in = cbuf[ccur]; da.in_bs = (long long)M * Cp; }
          if (uni) { da.out = xbuf[cur] + (size_t)N * C; da.out_bs = (long long)T * C; }
          else { da.out = cbuf[ccur ^ 1]; da.out_bs = (long long)M * C; ccur ^= 1; }
          da.W0 = sw.md_w0; da.b0 = sw.md_b0; da.g1 = sw.md_g1; da.be1 = sw.md_be1; da.W3 = sw.md_w3; da.b3 = sw.md_b3; da.g4 = sw.md_g4; da.be4 = sw.md_be4;
          da.B = B; da.Cp = Cp; da.C = C; da.eps = 1e-5f;
          char d[120];
          snprintf(d, sizeof(d), "meta_downsample B=%d %d->%d", B, Cp, C);
          sc->push([da](cudaStream_t s) { return meta_downsample_run(da, s); }, OP_META, 2.0 * B * M * 4.0 * Cp * (Cp + C), 0.0, d);
        } else {
        if (c_prev_unified) {
          const bf16* src = c_prev_unified;
          bf16* dst = cbuf[ccur];
          const int Np = g.N[i - 1], Tp = g.T[i - 1];
          sc->push([src, dst, B, M, Cp, Np, Tp](cudaStream_t s) { return gather_rows_run(src, dst, B * M, Cp, M, Tp, Np, s); });
        }
        b.linear(cprev, Cp, sw.md_w0, sw.md_b0, B * M, 4 * Cp, Cp, chid, 4 * Cp);
        b.ln(chid, chid, sw.md_g1, sw.md_be1, B * M, 4 * Cp, 1e-5f, /*gelu=*/1);
        b.linear(chid, 4 * Cp, sw.md_w3, sw.md_b3, B * M, C, 4 * Cp, ctmp, C);
        if (uni) b.ln(ctmp, xbuf[cur], sw.md_g4, sw.md_be4, B * M, C, 1e-5f, 0, M, T, N);
        else { b.ln(ctmp, cbuf[ccur ^ 1], sw.md_g4, sw.md_be4, B * M, C, 1e-5f); ccur ^= 1; }
        }
      }
    }
    bf16* cc = cbuf[ccur];
    // fused cross-attention blocks (dca_fused.cu + meta_branch.cu) wherever the shape allows; the unfused schedule otherwise
    const bool fuse_dca = !b.simt && plan->fused_dca && kind != 'S' && !uni && dca_supported(N, C, heads, M);
    void* dca_ws = ws + L.dca;
    // ---- blocks
    for (int j = 0; j < c.depth[i] && !b.rc; ++j) {
      const BlockW& bw = sw.blocks[j];
      if (kind == 'C') {
        // forward_with_c (models/lemevit.py:584-613) + CrossAttention (:477-486); x is returned unchanged
        // xn <- x + dw(x) (raw; norm1 is folded into the kv GEMM through stats1)
        const int sp = b.posln(xbuf[cur], bw.dw_w, bw.dw_b, xn, nullptr, B, g.H[i], g.W[i], T, C, stats1);
        if (fuse_dca) {
          b.dca_block('C', xn, stats1, sp, nullptr, nullptr, cc, bw, B, N, C, heads, Hd, 0.f, 1.0f / sqrtf((float)c.head_dim), dca_ws,
                      plan->dca_pipe ? 0 : 2);
        } else {
        b.ln(cc, cn, nullptr, nullptr, B * M, C, 1e-6f);
        b.linear(cn, C, bw.wa, bw.ba, B * M, C, C, cqkv, C);
        b.ln_linear(xn, stats1, sp, bw.wb, bw.bb, bw.csb, B * N, 2 * C, C, qkv);
        b.attn(cqkv, (long long)M * C, C, qkv, qkv + C, (long long)N * 2 * C, 2 * C, cn, (long long)M * C, C, B, heads, M, N,
               1.0f / sqrtf((float)c.head_dim));
        b.linear(cn, C, bw.wp1, bw.bp1, B * M, C, C, cc, C, 0, cc);
        b.ln(cc, cn, nullptr, nullptr, B * M, C, 1e-6f);
        b.linear(cn, C, bw.w1, bw.b1, B * M, Hd, C, chid, Hd, 1);
        b.linear(chid, Hd, bw.w2, bw.b2, B * M, C, Hd, cc, C, 0, cc);
        }
      } else if (kind == 'D') {
        // forward_with_xc (models/lemevit.py:542-582) + DualCrossAttention (:252-256,288-302)
        const int sp = b.posln(xbuf[cur], bw.dw_w, bw.dw_b, xbuf[cur ^ 1], nullptr, B, g.H[i], g.W[i], T, C, stats1);
        cur ^= 1;
        bf16* x = xbuf[cur];
        const double scale = 1.0 / std::sqrt((double)C);                       // :235 full channel dim
        const double scale_x = std::log((double)M) / std::log((double)N) * scale;  // :255 math.log(M, N)
        if (fuse_dca) {
          b.dca_block('D', x, stats1, sp, x, stats2, cc, bw, B, N, C, heads, Hd, (float)scale_x, (float)scale, dca_ws, plan->dca_pipe ? 0 : 2);
          b.mlp(x, stats2, 1, bw, B * N, C, Hd, hid);                                // norm2 folded, hidden stays on chip
        } else {
        b.ln(cc, cn, nullptr, nullptr, B * M, C, 1e-6f);
        b.ln_linear(x, stats1, sp, bw.wa, bw.ba, bw.csa, B * N, 3 * C, C, qkv);
        b.linear(cn, C, bw.wb, bw.bb, B * M, 3 * C, C, cqkv, 3 * C);
        b.attn(qkv, (long long)N * 3 * C, 3 * C, cqkv + C, cqkv + 2 * C, (long long)M * 3 * C, 3 * C, xn, (long long)N * C, C,
               B, heads, N, M, (float)scale_x);
        b.attn(cqkv, (long long)M * 3 * C, 3 * C, qkv + C, qkv + 2 * C, (long long)N * 3 * C, 3 * C, cn, (long long)M * C, C,
               B, heads, M, N, (float)scale);
        int parts2 = 1;
        b.linear_res_stats(xn, bw.wp1, bw.bp1, B * N, C, C, x, stats2, &parts2);     // x += proj_x(attn), stats2 = LN2 statistics
        b.linear(cn, C, bw.wp2, bw.bp2, B * M, C, C, cc, C, 0, cc);
        b.mlp(x, stats2, parts2, bw, B * N, C, Hd, hid);                             // norm2 folded, hidden stays on chip
        b.ln(cc, cn, nullptr, nullptr, B * M, C, 1e-6f);
        b.linear(cn, C, bw.w1, bw.b1, B * M, Hd, C, chid, Hd, 1);
        b.linear(chid, Hd, bw.w2, bw.b2, B * M, C, Hd, cc, C, 0, cc);
        }
      } else {
        // forward_with_x (models/lemevit.py:615-650) + StandardAttention (:199-205); image and meta tokens
        // share norm1/attn/norm2/mlp, so they travel in one [B, N+M, C] buffer (classification model only)
        const int sp = b.posln(xbuf[cur], bw.dw_w, bw.dw_b, xbuf[cur ^ 1], nullptr, B, g.H[i], g.W[i], T, C, stats1);
        cur ^= 1;
        bf16* x = xbuf[cur];
        b.ln_linear(x, stats1, sp, bw.wa, bw.ba, bw.csa, B * T, 3 * C, C, qkv);
        b.self_attn(qkv, xn, B, heads, T, N, C, 1.0f / sqrtf((float)c.head_dim));
        int parts2 = 1;
        b.linear_res_stats(xn, bw.wp1, bw.bp1, B * T, C, C, x, stats2, &parts2);
        b.mlp(x, stats2, parts2, bw, B * T, C, Hd, hid);
      }
      if (j + 1 == c.depth[i] || (i == plan->tap_stage && j == plan->tap_block)) b.flush_meta();   // end of the stage / tap reads c
      if (i == plan->tap_stage && j == plan->tap_block) {
        // test hook (lmv_plan_set_tap): dense copies of the block's outputs, x as tokens [B, N, C], c as [B, M, C]
        const bf16* xs = xbuf[cur];
        const bf16* cs = uni ? xbuf[cur] + (size_t)N * C : cc;
        const size_t xpitch = (size_t)T * C * 2, cpitch = uni ? xpitch : (size_t)M * C * 2;
        void *tx = plan->tap_x, *tc = plan->tap_c;
        sc->push([=](cudaStream_t s) {
          if (tx) LMV_CUDA_OK(cudaMemcpy2DAsync(tx, (size_t)N * C * 2, xs, xpitch, (size_t)N * C * 2, B, cudaMemcpyDeviceToDevice, s));
          if (tc) LMV_CUDA_OK(cudaMemcpy2DAsync(tc, (size_t)M * C * 2, cs, cpitch, (size_t)M * C * 2, B, cudaMemcpyDeviceToDevice, s));
          return LMV_OK;
        });
      }
    }
    // ---- backbone outputs: x after stages 1..S-1 as NCHW (semantic_segmentation/.../lemevit.py:800-820)
    if (c.backbone && i >= 1) {
      ToNchwArgs ta{xbuf[cur], nullptr, B, g.H[i], g.W[i], T, C, out_dtype};
      const int slot = i - 1;
      sc->push([ta, io, slot](cudaStream_t s) { ToNchwArgs a = ta; a.out = io->outs[slot]; return tokens_to_nchw_run(a, s); },
               OP_MISC, 0.0, (2.0 + (out_dtype == LMV_DTYPE_F32 ? 4 : 2)) * B * N * C);
    }
  }
  if (b.rc) return b.rc;
  if (!c.backbone) {
    // ---- tail + head (models/lemevit.py:815-836)
    const int i = S - 1, C = c.embed_dim[i];
    TailArgs ta;
    ta.x = xbuf[cur]; ta.x_bs = (long long)g.T[i] * C; ta.N = g.N[i];
    if (g.unified[i]) { ta.c = xbuf[cur] + (size_t)g.N[i] * C; ta.c_bs = (long long)g.T[i] * C; }
    else { ta.c = cbuf[ccur]; ta.c_bs = (long long)M * C; }
    ta.M = M; ta.C = C;
    ta.bn_scale = plan->bn_scale; ta.bn_shift = plan->bn_shift; ta.ln_gamma = plan->lnc_g; ta.ln_beta = plan->lnc_b;
    ta.eps = 1e-5f; ta.feat = feat; ta.B = B;
    sc->push([ta](cudaStream_t s) { return tail_run(ta, s); });
    if (flags & kFlagFeat) {
      const size_t bytes = (size_t)B * C * sizeof(bf16);
      sc->push([io, feat, bytes](cudaStream_t s) {
        if (!io->feat) return fail(LMV_ERR_INVALID, "forward: this schedule expects a feature output buffer");
        LMV_CUDA_OK(cudaMemcpyAsync(io->feat, feat, bytes, cudaMemcpyDeviceToDevice, s));
        return LMV_OK;
      });
    }
    if (flags & kFlagLogits) {
      if (c.num_classes <= 0) return fail(LMV_ERR_INVALID, "forward: logits requested from a model without a classifier (num_classes == 0)");
      GemmArgs ga;
      ga.A = feat; ga.lda = C; ga.W = plan->head_w; ga.ldw = C; ga.M = B; ga.N = c.num_classes; ga.K = C;
      ga.bias = plan->head_b; ga.ldc = c.num_classes; ga.out_fp32 = (out_dtype == LMV_DTYPE_F32);
      const bool simt = plan->debug_simt != 0;
      // the logits pointer changes per call/chunk: the tensor maps only cover A and W, so patch `out`
      const double hfl = 2.0 * B * c.num_classes * C;
      if (simt) {
        sc->push([ga, io](cudaStream_t s) { GemmArgs a = ga; a.out = io->logits; return gemm_simt_run(a, s); }, OP_MISC, hfl, 0);
      } else {
        ga.out = feat;  // placeholder for validation; replaced at launch
        ga.out_patched = 1;   // (so the epilogue must store through the pointer, not through a TMA map of the placeholder)
        GemmOp op;
        rc = gemm_prepare(ga, &op);
        if (rc) return rc;
        sc->push([op, io](cudaStream_t s) { GemmOp o = op; o.p.out = io->logits; return gemm_run(o, s); }, OP_GEMM, hfl, 0);
      }
    }
  }
  return LMV_OK;
}

static int harvest_profile(lmv_plan* plan);

static int get_schedule(lmv_plan* plan, int B, int H, int W, void* ws, size_t ws_bytes, int x_dtype, int out_dtype, int flags,
                        Schedule** out) {
  Geo g;
  int rc = geometry(plan->cfg, H, W, &g);
  if (rc) return rc;
  WsLayout L;
  ws_layout(plan->cfg, g, B, &L);
  if (ws_bytes < L.total) return fail(LMV_ERR_INVALID, "workspace too small: need " + std::to_string(L.total) + " bytes");
  LMV_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  auto key = std::make_tuple(B, H, W, (const void*)ws, x_dtype, out_dtype, plan->debug_simt, flags);
  auto it = plan->cache.find(key);
  if (it == plan->cache.end()) {
    if (plan->cache.size() >= 16) {
      rc = harvest_profile(plan);
      if (rc) return rc;
      plan->cache.clear();
    }
    std::unique_ptr<Schedule> sc(new Schedule());
    rc = build_schedule(plan, B, H, W, static_cast<uint8_t*>(ws), x_dtype, out_dtype, flags, sc.get());
    if (rc) return rc;
    it = plan->cache.emplace(key, std::move(sc)).first;
  }
  *out = it->second.get();
  return LMV_OK;
}

// profile mode: fold the event timings of the previous profiled pass into the per-class accumulators
static int harvest_profile(lmv_plan* plan) {
  if (plan->pending.empty()) return LMV_OK;
  LMV_CUDA_OK(cudaEventSynchronize(plan->events[plan->pending.size()]));
  for (size_t k = 0; k < plan->pending.size(); ++k) {
    float ms = 0.f;
    LMV_CUDA_OK(cudaEventElapsedTime(&ms, plan->events[k], plan->events[k + 1]));
    const OpRec* op = plan->pending[k];
    plan->acc.ms[op->cls] += ms;
    plan->acc.flops[op->cls] += op->flops;
    plan->acc.bytes[op->cls] += op->bytes;
    plan->acc.launches[op->cls] += 1;
    ProfDetail& d = plan->acc.detail[op->desc];
    d.ms += ms; d.flops += op->flops; d.bytes += op->bytes; d.n += 1;
  }
  plan->pending.clear();
  return LMV_OK;
}

static int run_forward(lmv_plan* plan, const void* x, int x_dtype, int B, int H, int W, void* ws, size_t ws_bytes,
                       void* logits, void* const* outs, int n_outs, int out_dtype, cudaStream_t stream,
                       const void* c_in = nullptr, void* feat = nullptr) {
  LMV_REQUIRE(plan && x && B > 0, "forward: null plan/input or empty batch");
  LMV_REQUIRE(x_dtype == LMV_DTYPE_BF16 || x_dtype == LMV_DTYPE_F32 || x_dtype == LMV_DTYPE_U8 || x_dtype == LMV_DTYPE_U8_NHWC, "forward: x dtype");
  LMV_REQUIRE(out_dtype == LMV_DTYPE_BF16 || out_dtype == LMV_DTYPE_F32, "forward: output dtype");
  const lmv_config& c = plan->cfg;
  const int chunk = (plan->chunk > 0 && plan->chunk < B) ? plan->chunk : B;
  const size_t xe = x_dtype == LMV_DTYPE_F32 ? 4 : (x_dtype == LMV_DTYPE_BF16 ? 2 : 1), oe = out_dtype == LMV_DTYPE_F32 ? 4 : 2;
  Geo g;
  int rc = geometry(c, H, W, &g);
  if (rc) return rc;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bc = std::min(chunk, B - b0);
    Schedule* sc = nullptr;
    const int flags = (c_in ? kFlagCustomC : 0) | (feat ? kFlagFeat : 0) | (logits ? kFlagLogits : 0);
    rc = get_schedule(plan, bc, H, W, ws, ws_bytes, x_dtype, out_dtype, flags, &sc);
    if (rc) return rc;
    const int CL = c.embed_dim[c.num_stages - 1];
    sc->io->c_in = c_in ? static_cast<const uint8_t*>(c_in) + (size_t)b0 * c.queries_len * c.embed_dim[0] * sizeof(bf16) : nullptr;
    sc->io->feat = feat ? static_cast<uint8_t*>(feat) + (size_t)b0 * CL * sizeof(bf16) : nullptr;
    sc->io->x = static_cast<const uint8_t*>(x) + (size_t)b0 * c.in_chans * H * W * xe;
    if (logits) sc->io->logits = static_cast<uint8_t*>(logits) + (size_t)b0 * c.num_classes * oe;
    for (int k = 0; k < n_outs; ++k)
      sc->io->outs[k] = static_cast<uint8_t*>(outs[k]) + (size_t)b0 * c.embed_dim[k + 1] * g.N[k + 1] * oe;
    if (!plan->profile) {
      for (auto& op : sc->ops) {
        rc = op.fn(stream);
        if (rc) return rc;
      }
    } else {
      rc = harvest_profile(plan);
      if (rc) return rc;
      while (plan->events.size() < sc->ops.size() + 1) {
        cudaEvent_t e;
        LMV_CUDA_OK(cudaEventCreate(&e));
        plan->events.push_back(e);
      }
      LMV_CUDA_OK(cudaEventRecord(plan->events[0], stream));
      for (size_t k = 0; k < sc->ops.size(); ++k) {
        rc = sc->ops[k].fn(stream);
        if (rc) return rc;
        LMV_CUDA_OK(cudaEventRecord(plan->events[k + 1], stream));
        plan->pending.push_back(&sc->ops[k]);
      }
    }
  }
  return LMV_OK;
}

}  // namespace lmv

// ================================================================================================
// extern "C"
// ================================================================================================
using namespace lmv;

extern "C" {

const char* lmv_last_error(void) { return g_err.c_str(); }
int lmv_version(void) { return 100; }

int lmv_packed_tensor_count(const lmv_config* cfg) {
  if (!cfg) return fail(LMV_ERR_INVALID, "null config");
  int rc = validate_config(*cfg);
  if (rc) return rc;
  PackWalker w{nullptr, 0};
  walk(*cfg, w, nullptr);
  return w.i;
}

int lmv_plan_create(const lmv_config* cfg, const lmv_tensor* packed, int n_packed, lmv_plan** out) {
  if (!cfg || !packed || !out) return fail(LMV_ERR_INVALID, "plan_create: null argument");
  int rc = validate_config(*cfg);
  if (rc) return rc;
  std::unique_ptr<lmv_plan> plan(new lmv_plan());
  plan->cfg = *cfg;
  plan->head_w = nullptr; plan->head_b = nullptr;
  PackWalker w{packed, n_packed};
  walk(*cfg, w, plan.get());
  if (!w.err.empty()) return fail(LMV_ERR_INVALID, w.err);
  if (w.i != n_packed) return fail(LMV_ERR_INVALID, "packed weights: expected " + std::to_string(w.i) + " tensors, got " + std::to_string(n_packed));
  *out = plan.release();
  return LMV_OK;
}

void lmv_plan_destroy(lmv_plan* plan) {
  if (!plan) return;
  for (cudaEvent_t e : plan->events) cudaEventDestroy(e);
  delete plan;
}

int lmv_plan_set_profile(lmv_plan* plan, int enable) {
  if (!plan) return fail(LMV_ERR_INVALID, "null plan");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  plan->profile = enable ? 1 : 0;
  if (enable) plan->acc = ProfAcc();
  return LMV_OK;
}

int lmv_plan_profile_report(lmv_plan* plan, char* buf, int buf_bytes) {
  if (!plan || !buf || buf_bytes <= 0) return fail(LMV_ERR_INVALID, "profile_report: bad argument");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  std::string out;
  char line[320];
  for (auto& kv : plan->acc.detail) {
    const ProfDetail& d = kv.second;
    snprintf(line, sizeof(line), "%-52s n=%-5lld ms=%9.3f avg_us=%9.2f TFLOPs=%8.1f GBs=%8.1f\n", kv.first.c_str(), d.n, d.ms,
             d.ms / d.n * 1e3, d.flops / d.ms * 1e-9, d.bytes / d.ms * 1e-6);
    out += line;
  }
  snprintf(buf, buf_bytes, "%s", out.c_str());
  return (int)out.size();
}

int lmv_plan_get_profile(lmv_plan* plan, lmv_profile_entry* out, int max_entries) {
  if (!plan || !out) return fail(LMV_ERR_INVALID, "get_profile: null argument");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  int n = 0;
  for (int c = 0; c < OP_NUM_CLASSES && n < max_entries; ++c) {
    if (!plan->acc.launches[c]) continue;
    out[n].name = kOpClassNames[c];
    out[n].launches = plan->acc.launches[c];
    out[n].device_ms = plan->acc.ms[c];
    out[n].flops = plan->acc.flops[c];
    out[n].bytes = plan->acc.bytes[c];
    ++n;
  }
  return n;
}

int lmv_plan_set_chunk(lmv_plan* plan, int n) {
  if (!plan || n < 0) return fail(LMV_ERR_INVALID, "set_chunk: bad argument");
  plan->chunk = n;
  return LMV_OK;
}
int lmv_plan_set_option(lmv_plan* plan, const char* name, int value) {
  if (!plan || !name) return fail(LMV_ERR_INVALID, "set_option: null argument");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  const std::string n(name);
  if (n == "fused_mlp") plan->fused_mlp = value ? 1 : 0;
  else if (n == "fused_mlp_wide") plan->fused_mlp_wide = value ? 1 : 0;
  else if (n == "fused_self_attn") plan->fused_self_attn = value ? 1 : 0;
  else if (n == "direct_stem") plan->direct_stem = value ? 1 : 0;
  else if (n == "implicit_conv") plan->implicit_conv = value ? 1 : 0;
  else if (n == "stem_tc") plan->stem_tc = value ? 1 : 0;
  else if (n == "fused_dca") plan->fused_dca = value ? 1 : 0;
  else if (n == "dca_pipe") plan->dca_pipe = value ? 1 : 0;
  else return fail(LMV_ERR_INVALID, "set_option: unknown option " + n);
  plan->cache.clear();   // schedules are rebuilt with the new setting
  return LMV_OK;
}
int lmv_plan_set_tap(lmv_plan* plan, int stage, int block, void* x_tokens_out, void* c_out) {
  if (!plan) return fail(LMV_ERR_INVALID, "null plan");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  plan->tap_stage = stage; plan->tap_block = block; plan->tap_x = x_tokens_out; plan->tap_c = c_out;
  plan->cache.clear();
  return LMV_OK;
}
int lmv_plan_set_input_norm(lmv_plan* plan, const float* mean3, const float* std3) {
  if (!plan || !mean3 || !std3) return fail(LMV_ERR_INVALID, "set_input_norm: null argument");
  if (plan->cfg.in_chans != 3) return fail(LMV_ERR_UNSUPPORTED, "set_input_norm: the 8-bit input path needs 3 input channels");
  for (int i = 0; i < 3; ++i)
    if (!(std3[i] > 0.f) || !std::isfinite(mean3[i]) || !std::isfinite(std3[i])) return fail(LMV_ERR_INVALID, "set_input_norm: std must be positive and finite");
  int rc = harvest_profile(plan);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) { plan->in_mean[i] = mean3[i]; plan->in_std[i] = std3[i]; }
  plan->cache.clear();   // the stem launch carries the constants by value
  return LMV_OK;
}
int lmv_plan_set_debug_simt(lmv_plan* plan, int enable) {
  if (!plan) return fail(LMV_ERR_INVALID, "null plan");
  plan->debug_simt = enable ? 1 : 0;
  return LMV_OK;
}

size_t lmv_workspace_bytes(const lmv_plan* plan, int batch, int H, int W) {
  if (!plan || batch <= 0) return 0;
  Geo g;
  if (geometry(plan->cfg, H, W, &g)) return 0;
  const int chunk = (plan->chunk > 0 && plan->chunk < batch) ? plan->chunk : batch;
  WsLayout L;
  ws_layout(plan->cfg, g, chunk, &L);
  return L.total;
}

int lmv_launch_count(lmv_plan* plan, int batch, int H, int W) {
  if (!plan || batch <= 0) return fail(LMV_ERR_INVALID, "launch_count: bad argument");
  const int chunk = (plan->chunk > 0 && plan->chunk < batch) ? plan->chunk : batch;
  int total = 0;
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int bc = std::min(chunk, batch - b0);
    for (auto& kv : plan->cache)
      if (std::get<0>(kv.first) == bc && std::get<1>(kv.first) == H && std::get<2>(kv.first) == W) {
        total += (int)kv.second->ops.size();
        goto next;
      }
    return fail(LMV_ERR_INVALID, "launch_count: run a forward of this shape first");
  next:;
  }
  return total;
}

int lmv_forward_cls(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, void* workspace,
                    size_t workspace_bytes, void* logits, int logits_dtype, void* stream) {
  if (!plan || plan->cfg.backbone) return fail(LMV_ERR_INVALID, "forward_cls: plan was created for the backbone variant");
  if (!logits) return fail(LMV_ERR_INVALID, "forward_cls: null logits");
  return run_forward(plan, x, x_dtype, batch, H, W, workspace, workspace_bytes, logits, nullptr, 0, logits_dtype,
                     static_cast<cudaStream_t>(stream));
}

int lmv_forward_cls_features(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, const void* meta_tokens,
                             void* workspace, size_t workspace_bytes, void* features, void* logits, int logits_dtype, void* stream) {
  if (!plan || plan->cfg.backbone) return fail(LMV_ERR_INVALID, "forward_cls_features: plan was created for the backbone variant");
  if (!features && !logits) return fail(LMV_ERR_INVALID, "forward_cls_features: neither features nor logits requested");
  if (logits && plan->cfg.num_classes <= 0) return fail(LMV_ERR_INVALID, "forward_cls_features: the model has no classifier (num_classes == 0)");
  return run_forward(plan, x, x_dtype, batch, H, W, workspace, workspace_bytes, logits, nullptr, 0, logits ? logits_dtype : LMV_DTYPE_BF16,
                     static_cast<cudaStream_t>(stream), meta_tokens, features);
}

int lmv_forward_features(lmv_plan* plan, const void* x, int x_dtype, int batch, int H, int W, void* workspace,
                         size_t workspace_bytes, void* const* outs, int n_outs, int out_dtype, void* stream) {
  if (!plan || !plan->cfg.backbone) return fail(LMV_ERR_INVALID, "forward_features: plan was created for the classification variant");
  if (!outs || n_outs != plan->cfg.num_stages - 1) return fail(LMV_ERR_INVALID, "forward_features: expects num_stages-1 output maps");
  for (int k = 0; k < n_outs; ++k)
    if (!outs[k]) return fail(LMV_ERR_INVALID, "forward_features: null output");
  return run_forward(plan, x, x_dtype, batch, H, W, workspace, workspace_bytes, nullptr, outs, n_outs, out_dtype,
                     static_cast<cudaStream_t>(stream));
}

// ---- per-kernel entry points -------------------------------------------------------------------
static GemmArgs make_gemm_args(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                               void* out, int ldc, int M, int N, int K, int act_gelu, int out_dtype) {
  GemmArgs a;
  a.A = static_cast<const bf16*>(A); a.lda = lda; a.W = static_cast<const bf16*>(W); a.ldw = ldw;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.act = act_gelu ? 1 : 0;
  a.residual = static_cast<const bf16*>(residual); a.out = out; a.ldc = ldc; a.out_fp32 = out_dtype == LMV_DTYPE_F32;
  return a;
}

int lmv_linear(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, void* out,
               int ldc, int M, int N, int K, int act_gelu, int out_dtype, int force_tile_n, void* stream) {
  GemmArgs a = make_gemm_args(A, lda, W, ldw, bias, residual, out, ldc, M, N, K, act_gelu, out_dtype);
  a.force_bn = force_tile_n;
  GemmOp op;
  int rc = gemm_prepare(a, &op);
  if (rc) return rc;
  return gemm_run(op, static_cast<cudaStream_t>(stream));
}

int lmv_linear_fused(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, void* out,
                     int ldc, int M, int N, int K, int act_gelu, int out_dtype, const float* ln_stats, int ln_parts,
                     const float* ln_colsum, float ln_eps, float* stats_out, int use_simt, void* stream) {
  GemmArgs a = make_gemm_args(A, lda, W, ldw, bias, residual, out, ldc, M, N, K, act_gelu, out_dtype);
  a.ln_stats = ln_stats; a.ln_parts = ln_parts; a.ln_colsum = ln_colsum; a.ln_eps = ln_eps; a.stats_out = stats_out;
  if (use_simt) {
    if (!a.A || !a.W || !a.out || M <= 0 || N <= 0 || K <= 0) return fail(LMV_ERR_INVALID, "linear_fused: bad argument");
    if ((ln_stats == nullptr) != (ln_colsum == nullptr)) return fail(LMV_ERR_INVALID, "linear_fused: ln_stats and ln_colsum go together");
    int rc = gemm_simt_run(a, static_cast<cudaStream_t>(stream));
    if (rc || !stats_out) return rc;
    if (out_dtype != LMV_DTYPE_BF16 || ldc != N) return fail(LMV_ERR_INVALID, "linear_fused(simt): stats_out needs a dense bf16 output");
    return row_stats_run(static_cast<const bf16*>(out), stats_out, M, N, static_cast<cudaStream_t>(stream));
  }
  GemmOp op;
  int rc = gemm_prepare(a, &op);
  if (rc) return rc;
  return gemm_run(op, static_cast<cudaStream_t>(stream));
}

int lmv_mlp_fused(const void* x, const void* resid, void* out, const void* W1, const float* b1, const float* colsum1,
                  const void* W2, const float* b2, const float* ln_stats, int ln_parts, float ln_eps, int R, int C, int Hd,
                  void* stream) {
  MlpArgs a;
  a.x = static_cast<const bf16*>(x); a.resid = static_cast<const bf16*>(resid); a.out = static_cast<bf16*>(out);
  a.W1 = static_cast<const bf16*>(W1); a.b1 = b1; a.cs1 = colsum1; a.W2 = static_cast<const bf16*>(W2); a.b2 = b2;
  a.ln_stats = ln_stats; a.ln_parts = ln_parts; a.ln_eps = ln_eps; a.R = R; a.C = C; a.Hd = Hd;
  MlpOp op;
  int rc = mlp_fused_prepare(a, &op);
  if (rc) return rc;
  return mlp_fused_run(op, static_cast<cudaStream_t>(stream));
}

int lmv_linear_stats_parts(int N, int use_simt) { return use_simt ? 1 : gemm_stats_parts(N); }

int lmv_linear_simt(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                    void* out, int ldc, int M, int N, int K, int act_gelu, int out_dtype, void* stream) {
  GemmArgs a = make_gemm_args(A, lda, W, ldw, bias, residual, out, ldc, M, N, K, act_gelu, out_dtype);
  if (!a.A || !a.W || !a.out || M <= 0 || N <= 0 || K <= 0) return fail(LMV_ERR_INVALID, "linear_simt: bad argument");
  return gemm_simt_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_posembed_layernorm(const void* tokens, const float* dw_weight, const float* dw_bias, void* resid_out,
                           void* norm_out, float* stats_out, int B, int H, int W, int T, int C, float eps, void* stream) {
  if (!tokens || (!resid_out && !norm_out && !stats_out)) return fail(LMV_ERR_INVALID, "posembed_layernorm: null pointer");
  PosLnArgs a{static_cast<const bf16*>(tokens), dw_weight, dw_bias, static_cast<bf16*>(resid_out),
              static_cast<bf16*>(norm_out), B, H, W, T, C, eps, stats_out};
  return posembed_ln_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_layernorm(const void* in, void* out, const float* gamma, const float* beta, int R, int C, float eps,
                  int act_gelu, int grp_rows, int grp_stride, int grp_off, void* stream) {
  if (!in || !out || ((gamma == nullptr) != (beta == nullptr))) return fail(LMV_ERR_INVALID, "layernorm: bad pointers");
  LnArgs a{static_cast<const bf16*>(in), static_cast<bf16*>(out), gamma, beta, R, C, eps, act_gelu, grp_rows, grp_stride, grp_off};
  return layernorm_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_attention(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                  long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int Lq, int Lk,
                  float scale, int impl, void* stream) {
  if (!q || !k || !v || !out) return fail(LMV_ERR_INVALID, "attention: null pointer");
  AttnArgs a;
  a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
  a.out = static_cast<bf16*>(out);
  a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.o_bs = o_bs;
  a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs; a.o_rs = o_rs;
  a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
  if (impl == 0 && attention_tc_supported(a)) return attention_tc_run(a, static_cast<cudaStream_t>(stream));
  if (impl == 2) return attention_tc_supported(a) ? attention_tc_run(a, static_cast<cudaStream_t>(stream))
                                                  : fail(LMV_ERR_UNSUPPORTED, "attention: shape not supported by the tcgen05 kernel");
  return attention_simt_run(a, static_cast<cudaStream_t>(stream));
}

static AttnArgs make_attn_args(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                               long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int Lq, int Lk,
                               float scale) {
  AttnArgs a;
  a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
  a.out = static_cast<bf16*>(out);
  a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.o_bs = o_bs;
  a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs; a.o_rs = o_rs;
  a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
  return a;
}

int lmv_attention_self(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                       long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int T, int N, float scale,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !out) return fail(LMV_ERR_INVALID, "attention_self: null pointer");
  AttnArgs a = make_attn_args(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, out, o_bs, o_rs, B, heads, T, T, scale);
  return attention_self_run(a, T, N, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t lmv_attention_self_workspace(int B, int heads, int T) { return attention_self_workspace(B, heads, T); }

size_t lmv_attention_meta_workspace(int B, int heads, int Lq, int Lk) {
  AttnArgs a{};
  a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk;
  return attention_meta_workspace(a);
}

int lmv_attention_meta(const void* q, long long q_bs, int q_rs, const void* k, long long k_bs, int k_rs, const void* v,
                       long long v_bs, int v_rs, void* out, long long o_bs, int o_rs, int B, int heads, int Lq, int Lk,
                       float scale, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !out) return fail(LMV_ERR_INVALID, "attention_meta: null pointer");
  AttnArgs a = make_attn_args(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, out, o_bs, o_rs, B, heads, Lq, Lk, scale);
  return attention_meta_run(a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t lmv_dca_workspace_bytes(int B, int N, int C, int heads) {
  if (B <= 0 || !dca_supported(N, C, heads, kDcaM)) return 0;
  return dca_workspace_bytes(dca_geometry(B, N, C, heads));
}

int lmv_dca_block(int kind, const void* xt, const float* stats1, int parts1, void* xout, float* stats2, void* c, const void* wa,
                  const float* ba, const void* wb, const float* bb, const void* wxt, const void* wp1, const float* bp1, const void* wp2,
                  const float* bp2, const void* w1, const float* b1, const void* w2, const float* b2, int B, int N, int C, int heads, int Hd,
                  float scale_x, float scale_c, void* workspace, size_t workspace_bytes, int flags, void* stream) {
  if (kind != 'C' && kind != 'D') return fail(LMV_ERR_INVALID, "dca_block: kind must be 'C' or 'D'");
  if (!xt || !stats1 || !c || !wa || !ba || !wb || !bb || !wxt || !wp1 || !bp1 || !w1 || !b1 || !w2 || !b2 || !workspace)
    return fail(LMV_ERR_INVALID, "dca_block: null pointer");
  if (kind == 'D' && (!xout || !stats2 || !wp2 || !bp2)) return fail(LMV_ERR_INVALID, "dca_block: null pointer ('D' block)");
  if (!dca_supported(N, C, heads, kDcaM)) return fail(LMV_ERR_UNSUPPORTED, "dca_block: needs C = heads * 32 <= 192");
  if (workspace_bytes < lmv_dca_workspace_bytes(B, N, C, heads) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(LMV_ERR_INVALID, "dca_block: workspace too small or not 256-byte aligned");
  BlockW bw;
  memset(&bw, 0, sizeof(bw));
  bw.wa = static_cast<const bf16*>(wa); bw.ba = ba; bw.wb = static_cast<const bf16*>(wb); bw.bb = bb;
  bw.wp1 = static_cast<const bf16*>(wp1); bw.bp1 = bp1; bw.wp2 = static_cast<const bf16*>(wp2); bw.bp2 = bp2;
  bw.w1 = static_cast<const bf16*>(w1); bw.b1 = b1; bw.w2 = static_cast<const bf16*>(w2); bw.b2 = b2;
  if (kind == 'D') { bw.wxqT = static_cast<const bf16*>(wxt); bw.wxkT = bw.wxqT + (size_t)C * C; }
  else bw.wxkT = static_cast<const bf16*>(wxt);
  Schedule sc;
  Builder b{nullptr, &sc, LMV_OK, false};
  b.dca_block((char)kind, static_cast<const bf16*>(xt), stats1, parts1, static_cast<bf16*>(xout), stats2, static_cast<bf16*>(c), bw, B, N, C,
              heads, Hd, scale_x, scale_c, workspace, flags & 3);
  b.flush_meta();
  if (b.rc) return b.rc;
  for (auto& op : sc.ops) {
    int rc = op.fn(static_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  return LMV_OK;
}

int lmv_stem_im2col(const void* x, int x_dtype, void* out, int B, int Cin, int H, int W, void* stream) {
  if (!x || !out) return fail(LMV_ERR_INVALID, "stem_im2col: null pointer");
  StemArgs a{x, x_dtype, static_cast<bf16*>(out), B, Cin, H, W};
  return stem_im2col_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_stem_conv1(const void* x, int x_dtype, const void* w, const float* bias, void* out, int B, int Cin, int C1, int H, int W,
                   void* stream) {
  if (!x || !w || !bias || !out) return fail(LMV_ERR_INVALID, "stem_conv1: null pointer");
  StemArgs a{x, x_dtype, static_cast<bf16*>(out), B, Cin, H, W};
  if (const char* e = getenv("LMV_STEM_TC")) a.tensor_core = e[0] != '0';   // unit tests run both kernels
  return stem_conv1_run(a, static_cast<const bf16*>(w), bias, C1, static_cast<cudaStream_t>(stream));
}

int lmv_conv3x3s2(const void* in, const void* w, const float* bias, void* out, int B, int H, int W, int T, int Cin, int Cout,
                  int out_rows_per_image, void* stream) {
  if (!in || !w || !out) return fail(LMV_ERR_INVALID, "conv3x3s2: null pointer");
  if (!gemm_conv_supported(H, W, Cin)) return fail(LMV_ERR_UNSUPPORTED, "conv3x3s2: needs an output width <= 128 and Cin % 8 == 0");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  GemmArgs a;
  a.A = static_cast<const bf16*>(in); a.W = static_cast<const bf16*>(w); a.ldw = 9 * Cin; a.M = B * Ho * Wo; a.N = Cout; a.K = 9 * Cin;
  a.bias = bias; a.out = out; a.ldc = Cout;
  if (out_rows_per_image > 0) { a.grp_rows = Ho * Wo; a.grp_stride = out_rows_per_image; }
  a.conv_B = B; a.conv_H = H; a.conv_W = W; a.conv_T = T; a.conv_C = Cin;
  GemmOp op;
  int rc = gemm_prepare(a, &op);
  if (rc) return rc;
  return gemm_run(op, static_cast<cudaStream_t>(stream));
}

int lmv_im2col_3x3s2(const void* in, void* out, int B, int H, int W, int T, int C, void* stream) {
  if (!in || !out) return fail(LMV_ERR_INVALID, "im2col: null pointer");
  Im2colArgs a{static_cast<const bf16*>(in), static_cast<bf16*>(out), B, H, W, T, C};
  return im2col_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_tail(const void* x, long long x_bs, int N, const void* c, long long c_bs, int M, int C, const float* bn_scale,
             const float* bn_shift, const float* ln_gamma, const float* ln_beta, float eps, void* feat, int B,
             void* stream) {
  if (!x || !c || !feat || !bn_scale || !bn_shift || !ln_gamma || !ln_beta) return fail(LMV_ERR_INVALID, "tail: null pointer");
  TailArgs a;
  a.x = static_cast<const bf16*>(x); a.x_bs = x_bs; a.N = N; a.c = static_cast<const bf16*>(c); a.c_bs = c_bs; a.M = M;
  a.C = C; a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.ln_gamma = ln_gamma; a.ln_beta = ln_beta; a.eps = eps;
  a.feat = static_cast<bf16*>(feat); a.B = B;
  return tail_run(a, static_cast<cudaStream_t>(stream));
}

int lmv_tokens_to_nchw(const void* tokens, void* out, int B, int H, int W, int T, int C, int out_dtype, void* stream) {
  if (!tokens || !out) return fail(LMV_ERR_INVALID, "tokens_to_nchw: null pointer");
  ToNchwArgs a{static_cast<const bf16*>(tokens), out, B, H, W, T, C, out_dtype};
  return tokens_to_nchw_run(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
