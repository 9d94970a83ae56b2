// Meta-token side of the LeMeViT cross attentions on tensor cores: a handful of queries (the M = 16 meta tokens)
// attend over ALL N image tokens, softmax over the image tokens:
//   CrossAttention        (models/lemevit.py:477-486, stage 0):  c <- softmax(q(c) k(x)^T / sqrt(32)) v(x)
//   DualCrossAttention    (models/lemevit.py:300-302, stages 1-2): dc = softmax(q2 k1^T * C^-0.5) v1
// FlashAttention-style kernels waste >75 % of their tiles on 16 query rows; here the roles are arranged so that
// the tcgen05 M dimension is filled by (head, query) pairs and the N/K dimensions by image tokens.
//
// sm_100a design, split-N: one CTA per (image, 128-token tile):
//   * Q is expanded in shared memory to a block-diagonal operand Qbd[(h, j), C] (row (h, j) holds q_j restricted to the
//     32 channels of head h), so ONE accumulation over the full channel dim gives every head's scores:
//         S[(h, j), n] = sum_c Qbd[(h, j), c] K[n, c]                tcgen05.mma M=128, N=128, K=C   (A, B K-major, SW64)
//   * K and V tiles [128 tokens x C] arrive by TMA straight out of the packed kv / qkv activation (64B swizzle, 32-channel
//     boxes; rows past the end of the image are zero-filled by TMA and masked in the softmax);
//   * softmax over the tile's tokens is thread-local (thread r owns TMEM lane r = row (h, j)); P (bf16) goes to a 128B
//     swizzled K-major smem tile;
//   * O[(h, j), c] = sum_n P[(h, j), n] V[n, c]                      tcgen05.mma M=128, N=32 per head chunk, K=128 tokens
//     (B = V as loaded: MN-major);
//   * per tile the kernel emits the split-softmax partial (m, l, O[32]) of every (h, j); a small second kernel merges
//     the partials of all tiles of an image (fixed order, deterministic) and writes merged-heads bf16 output.
// The same partial structure is what a fully fused DualCrossAttention block kernel emits for its meta-token branch.
#include <mutex>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kD = 32;
constexpr int kTile = 128;      // image tokens per CTA
constexpr int kThreads = 128;

struct MetaParams {
  const bf16* q;
  long long q_bs;
  int q_rs;
  float* part_o;     // [B][tiles][R][32]
  float2* part_ml;   // [B][tiles][R]
  int heads, Lq, Lk, R, C, tiles, nchunk, tmem_cols;
  float scale_log2e;
};

struct Ctrl {
  uint64_t bar_load, bar_s, bar_o;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t make_mnmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;   // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
attention_meta_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const MetaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const int chunk_bytes = kTile * kD * 2;                  // [128 rows x 32 ch] bf16 = 8 KB
  uint8_t* sQ = smem + 1024;
  uint8_t* sK = sQ + (size_t)p.nchunk * chunk_bytes;
  uint8_t* sV = sK + (size_t)p.nchunk * chunk_bytes;
  uint8_t* sP = sV + (size_t)p.nchunk * chunk_bytes;        // 2 tiles of [128 x 64] bf16, 128B swizzle
  const int warp = threadIdx.x >> 5;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int n0 = tile * kTile;
  const int valid = min(kTile, p.Lk - n0);

  pdl_launch_dependents();
  pdl_wait();   // the TMA loads below are issued right away
  if (threadIdx.x == 0) {
    mbar_init(&ctrl->bar_load, 1);
    mbar_init(&ctrl->bar_s, 1);
    mbar_init(&ctrl->bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_expect_tx(&ctrl->bar_load, (uint32_t)(2 * p.nchunk * chunk_bytes));
    for (int c = 0; c < p.nchunk; ++c) {
      tma_load_3d(sK + (size_t)c * chunk_bytes, &tmK, &ctrl->bar_load, c * kD, n0, b);
      tma_load_3d(sV + (size_t)c * chunk_bytes, &tmV, &ctrl->bar_load, c * kD, n0, b);
    }
  }
  if (warp == 0) {
    tmem_alloc(&ctrl->tmem_base, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  // ---- block-diagonal Q operand: zero, then row (h, j) <- q[j, 32h .. 32h+31] into chunk h (64B-swizzled K-major) ----
  {
    const int n16 = p.nchunk * chunk_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kThreads) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int r = threadIdx.x;                 // row (h, j) == TMEM lane
  const int h = r / p.Lq, j = r - h * p.Lq;
  if (r < p.R) {
    const uint4* src = reinterpret_cast<const uint4*>(p.q + (long long)b * p.q_bs + (long long)j * p.q_rs + h * kD);
    uint8_t* dst = sQ + (size_t)h * chunk_bytes + (size_t)r * 64;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + ((ch ^ ((r >> 1) & 3)) << 4)) = __ldg(src + ch);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;
  const uint32_t tmemO = tmem + kTile;       // S: columns [0, 128), O: [128, 128 + C)

  if (threadIdx.x == 0) {
    mbar_wait(&ctrl->bar_load, 0, 30);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, kTile);
    for (int c = 0; c < p.nchunk; ++c) {
      const uint64_t dq = make_kmajor_desc<64>(smem_u32(sQ + (size_t)c * chunk_bytes));
      const uint64_t dk = make_kmajor_desc<64>(smem_u32(sK + (size_t)c * chunk_bytes));
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tmem, dq + 2ull * k, dk + 2ull * k, idesc, (uint32_t)((c | k) != 0));
    }
    umma_commit(&ctrl->bar_s);
  }
  mbar_wait(&ctrl->bar_s, 0, 31);
  tc_fence_after();

  // ---- softmax over this tile's tokens (row r) ----
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  float mx = -INFINITY, sum = 0.f;
  if (warp * 32 < p.R) {                       // warp-uniform: tcgen05.ld is .sync.aligned
    for (int c0 = 0; c0 < kTile; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(t_row + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    const float mxs = mx * p.scale_log2e;
    for (int c0 = 0; c0 < kTile; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(t_row + (uint32_t)c0, v);
      tmem_ld_wait();
      uint8_t* tile_p = sP + (size_t)(c0 >> 6) * (kTile * 128) + (size_t)r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int col = c0 + g * 8 + i;
          e[i] = (col < valid) ? exp2f(fmaf(__uint_as_float(v[g * 8 + i]), p.scale_log2e, -mxs)) : 0.f;
        }
        uint4 u;
        u.x = pack_bf16x2(e[0], e[1]); u.y = pack_bf16x2(e[2], e[3]);
        u.z = pack_bf16x2(e[4], e[5]); u.w = pack_bf16x2(e[6], e[7]);
        // the row sum must describe exactly the bf16 probabilities the tensor core multiplies
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        sum += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
        const int ch = ((c0 & 63) >> 3) + g;
        *reinterpret_cast<uint4*>(tile_p + ((ch ^ (r & 7)) << 4)) = u;
      }
    }
    if (r < p.R) p.part_ml[((long long)b * p.tiles + tile) * p.R + r] = make_float2(mxs, sum);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // O[:, 32c .. 32c+31] = P V_c : A = P tiles (K-major, 128B swizzle), B = V chunk as loaded (MN-major, 64B swizzle)
    const uint32_t idesc = make_idesc_bf16(128, kD) | (1u << 16);   // b_major = MN
    const uint32_t pbase = smem_u32(sP);
    for (int c = 0; c < p.nchunk; ++c) {
      const uint32_t vbase = smem_u32(sV + (size_t)c * chunk_bytes);
#pragma unroll
      for (int s = 0; s < kTile / 16; ++s) {
        const uint64_t da = make_kmajor_desc<128>(pbase + (uint32_t)(s >> 2) * (kTile * 128)) + 2ull * (s & 3);
        const uint64_t db = make_mnmajor_sw64_desc(vbase + (uint32_t)s * (16 * kD * 2));
        umma_bf16_ss(tmemO + (uint32_t)(c * kD), da, db, idesc, (uint32_t)(s != 0));
      }
    }
    umma_commit(&ctrl->bar_o);
  }
  mbar_wait(&ctrl->bar_o, 0, 32);
  tc_fence_after();
  if (warp * 32 < p.R) {
    // row (h, j) only needs its own head's 32 channels; h differs per lane, so load lane-uniform chunks and select
    const int h_lo = (warp * 32) / p.Lq, h_hi = min(p.heads - 1, (warp * 32 + 31) / p.Lq);
    float* dst = p.part_o + (((long long)b * p.tiles + tile) * p.R + r) * kD;
    for (int hh = h_lo; hh <= h_hi; ++hh) {
      uint32_t v[32];
      tmem_ld_x32(t_row + (uint32_t)(kTile + hh * kD), v);
      tmem_ld_wait();
      if (r < p.R && hh == h) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                          __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// merge the per-tile partials of one (image, head, query) row: one warp per row, lane = channel
__global__ void __launch_bounds__(256)
attention_meta_merge_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, bf16* __restrict__ out,
                            long long o_bs, int o_rs, int B, int tiles, int R, int Lq) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= (long long)B * R) return;
  const int b = (int)(idx / R), r = (int)(idx % R);
  const int h = r / Lq, j = r - h * Lq;
  float m = -INFINITY;
  for (int t = 0; t < tiles; ++t) m = fmaxf(m, part_ml[((long long)b * tiles + t) * R + r].x);
  float l = 0.f, o = 0.f;
  for (int t = 0; t < tiles; ++t) {
    const long long pr = ((long long)b * tiles + t) * R + r;
    const float2 ml = part_ml[pr];
    const float w = exp2f(ml.x - m);
    l = fmaf(ml.y, w, l);
    o = fmaf(part_o[pr * kD + lane], w, o);
  }
  out[(long long)b * o_bs + (long long)j * o_rs + h * kD + lane] = __float2bfloat16(o / l);
}

std::once_flag g_once;
cudaError_t g_attr = cudaSuccess;

int smem_bytes_for(int nchunk) { return 2048 + 3 * nchunk * kTile * kD * 2 + 2 * kTile * 128; }

}  // namespace

bool attention_meta_supported(const AttnArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int C = a.heads * kD;
  return a.Lq >= 1 && a.heads * a.Lq <= 128 && C <= 256 && a.Lk >= 1 && a.B >= 1 && al16(a.q) && al16(a.k) && al16(a.v) &&
         a.q_rs % 8 == 0 && a.q_bs % 8 == 0 && a.k_rs % 8 == 0 && a.v_rs % 8 == 0 && a.k_bs % 8 == 0 && a.v_bs % 8 == 0 &&
         a.k_rs >= C && a.v_rs >= C && smem_bytes_for(a.heads) <= 227 * 1024;
}

size_t attention_meta_workspace(const AttnArgs& a) {
  const size_t tiles = (a.Lk + kTile - 1) / kTile, R = (size_t)a.heads * a.Lq;
  return (size_t)a.B * tiles * R * (kD * sizeof(float) + sizeof(float2));
}

int attention_meta_run(const AttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!attention_meta_supported(a)) return fail(LMV_ERR_UNSUPPORTED, "attention_meta: unsupported shape / alignment");
  LMV_REQUIRE(workspace && workspace_bytes >= attention_meta_workspace(a) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
              "attention_meta: partial-softmax workspace missing or too small");
  std::call_once(g_once, [] {
    g_attr = cudaFuncSetAttribute(attention_meta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  LMV_CUDA_OK(g_attr);
  MetaParams p;
  p.q = a.q; p.q_bs = a.q_bs; p.q_rs = a.q_rs;
  p.heads = a.heads; p.Lq = a.Lq; p.Lk = a.Lk; p.R = a.heads * a.Lq; p.C = a.heads * kD;
  p.tiles = (a.Lk + kTile - 1) / kTile;
  p.nchunk = a.heads;
  const int need_cols = kTile + p.C;
  p.tmem_cols = need_cols <= 256 ? 256 : 512;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  const size_t rows = (size_t)a.B * p.tiles * p.R;
  p.part_o = static_cast<float*>(workspace);
  p.part_ml = reinterpret_cast<float2*>(p.part_o + rows * kD);
  CUtensorMap tk, tv;
  auto enc = [&](CUtensorMap* m, const bf16* base, long long bs, int rs) {
    uint64_t dims[3] = {(uint64_t)p.C, (uint64_t)a.Lk, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
    uint32_t box[3] = {kD, kTile, 1};
    return encode_tmap_bf16(m, base, 3, dims, strides, box, 64);
  };
  int rc;
  if ((rc = enc(&tk, a.k, a.k_bs, a.k_rs))) return rc;
  if ((rc = enc(&tv, a.v, a.v_bs, a.v_rs))) return rc;
  dim3 grid(p.tiles, a.B);
  LMV_CUDA_OK(launch_kernel(attention_meta_kernel, dim3(grid), dim3(kThreads), (size_t)(smem_bytes_for(p.nchunk)), s, tk, tv, p));
  LMV_CUDA_OK(cudaGetLastError());
  const long long mrows = (long long)a.B * p.R;
  LMV_CUDA_OK(launch_kernel(attention_meta_merge_kernel, dim3((unsigned)((mrows + 7) / 8)), dim3(256), (size_t)(0), s, p.part_o, p.part_ml, a.out, a.o_bs, a.o_rs, a.B, p.tiles, p.R, a.Lq));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
