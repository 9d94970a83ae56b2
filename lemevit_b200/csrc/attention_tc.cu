// Tensor-core attention for head_dim 32: softmax(scale * Q K^T) V per (image, head, 128-query tile).
// Reference: F.scaled_dot_product_attention call sites models/lemevit.py:203 (StandardAttention, stages 3-4 and the
// 16 meta tokens), :297 (DualCrossAttention x-branch: N image queries against the 16 meta tokens).
//
// sm_100a design (single KV block, Lk <= 224 — covers 196 / 49 / 16 keys):
//   TMA (64B swizzle) loads Q[128x32], K[Lkp x32], V[Lkp x32] of one head straight out of the packed qkv activation;
//   S = Q K^T   : tcgen05.mma M=128, N=Lkp, 2 K-steps, A/B K-major                 -> TMEM columns [0, Lkp)
//   softmax     : thread r owns row r (TMEM lane r): two passes of tcgen05.ld, exp2 with the scale folded in,
//                 P written as bf16 into 128B-swizzled K-major smem tiles (the A operand of the next MMA)
//   O = P V     : tcgen05.mma M=128, N=32, Lkp/16 K-steps, A K-major (P), B MN-major (V as loaded) -> TMEM columns [round32(Lkp), +32)
//   epilogue    : tcgen05.ld O, 1/rowsum, bf16, 64-byte row stores into the merged-heads layout.
// CTAs are small (29 KB smem / 64 TMEM columns for the 16 meta-token keys, <= ~100 KB / 256 columns for 196 keys) so
// 7-8 (resp. two) are resident per SM and overlap each other's phases.
#include <mutex>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kD = 32;
constexpr int kQTile = 128;
constexpr int kMaxKeys = 224;
constexpr int kThreads = 128;

struct AttnTcParams {
  bf16* out;
  long long o_bs;
  int o_rs;
  int Lq, Lk, Lkp;   // Lkp = Lk rounded up to 16
  int o_col, tmem_cols;   // O accumulator at the first 32-column boundary past S; allocation = next power of two (>= 32):
                          // 64 columns for the 16 meta-token keys, so eight CTAs fit the 512 TMEM columns of an SM
  float scale_log2e;
};

struct Ctrl {
  uint64_t bar_load, bar_s, bar_o;
  uint32_t tmem_base;
};

// smem descriptor for an MN-major operand stored as rows of 64 bytes (32 bf16 along MN), 64B swizzle, rows = K index:
// 8-row (8 k) atoms of 512 bytes follow each other -> stride byte offset 512; a single 32-wide MN group -> LBO unused.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t make_mnmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;   // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(kThreads)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  uint8_t* sQ = smem + 1024;
  uint8_t* sK = sQ + kQTile * kD * 2;
  uint8_t* sV = sK + p.Lkp * kD * 2;
  uint8_t* sP = sV + p.Lkp * kD * 2;          // ceil(Lkp/64) tiles of [128 x 64] bf16, 128B swizzle
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(&ctrl->bar_load, 1);
    mbar_init(&ctrl->bar_s, 1);
    mbar_init(&ctrl->bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) {
    tmem_alloc(&ctrl->tmem_base, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctrl->tmem_base;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&ctrl->bar_load, (uint32_t)(kQTile * kD * 2 + 2 * p.Lkp * kD * 2));
    tma_load_3d(sQ, &tmQ, &ctrl->bar_load, h * kD, q0, b);
    tma_load_3d(sK, &tmK, &ctrl->bar_load, h * kD, 0, b);
    tma_load_3d(sV, &tmV, &ctrl->bar_load, h * kD, 0, b);
    mbar_wait_lean(&ctrl->bar_load, 0);
    tc_fence_after();
    // S = Q K^T : both operands K-major, 64-byte rows (head_dim 32) -> 64B swizzle; 2 K-steps of 16
    const uint32_t idesc = make_idesc_bf16(kQTile, p.Lkp);
    const uint64_t dq = make_kmajor_desc<64>(smem_u32(sQ));
    const uint64_t dk = make_kmajor_desc<64>(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tmem, dq + 2ull * k, dk + 2ull * k, idesc, (uint32_t)(k != 0));
    umma_commit(&ctrl->bar_s);
  }
  mbar_wait_lean(&ctrl->bar_s, 0);
  tc_fence_after();

  // ---- softmax over the keys of row `r` ----
  const int r = threadIdx.x;                      // tile row == TMEM lane
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  float mx = -INFINITY;
  for (int c0 = 0; c0 < p.Lkp; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(t_row + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < p.Lk) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
  const float mxs = mx * p.scale_log2e;
  float sum = 0.f;
  for (int c0 = 0; c0 < p.Lkp; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(t_row + (uint32_t)c0, v);
    tmem_ld_wait();
    float e[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      e[j] = (c0 + j < p.Lk) ? ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs)) : 0.f;
      sum += e[j];
    }
    // P[r, c0 .. c0+15] -> tile (c0 / 64), 16-byte chunks (c0 % 64) / 8 and +1, XOR-swizzled with (r % 8)
    uint8_t* tile = sP + (size_t)(c0 >> 6) * (kQTile * 128) + (size_t)r * 128;
    const int ch = (c0 & 63) >> 3;
    uint4 u0, u1;
    u0.x = pack_bf16x2(e[0], e[1]);   u0.y = pack_bf16x2(e[2], e[3]);
    u0.z = pack_bf16x2(e[4], e[5]);   u0.w = pack_bf16x2(e[6], e[7]);
    u1.x = pack_bf16x2(e[8], e[9]);   u1.y = pack_bf16x2(e[10], e[11]);
    u1.z = pack_bf16x2(e[12], e[13]); u1.w = pack_bf16x2(e[14], e[15]);
    *reinterpret_cast<uint4*>(tile + (((ch) ^ (r & 7)) << 4)) = u0;
    *reinterpret_cast<uint4*>(tile + (((ch + 1) ^ (r & 7)) << 4)) = u1;
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // O = P V : A = P tiles (K-major, 128B swizzle), B = V as loaded ([key][32] rows = MN-major, 64B swizzle)
    const uint32_t idesc = make_idesc_bf16(kQTile, kD) | (1u << 16);   // b_major = MN
    const uint32_t pbase = smem_u32(sP), vbase = smem_u32(sV);
    const int ksteps = p.Lkp >> 4;
    for (int s = 0; s < ksteps; ++s) {
      const uint64_t da = make_kmajor_desc<128>(pbase + (uint32_t)(s >> 2) * (kQTile * 128)) + 2ull * (s & 3);
      const uint64_t db = make_mnmajor_sw64_desc(vbase + (uint32_t)s * (16 * kD * 2));
      umma_bf16_ss(tmem + (uint32_t)p.o_col, da, db, idesc, (uint32_t)(s != 0));
    }
    umma_commit(&ctrl->bar_o);
  }
  mbar_wait_lean(&ctrl->bar_o, 0);
  tc_fence_after();
  {
    uint32_t v[32];
    tmem_ld_x32(t_row + (uint32_t)p.o_col, v);
    tmem_ld_wait();
    const int qrow = q0 + r;
    if (qrow < p.Lq) {
      const float inv = 1.f / sum;
      bf16* op = p.out + (long long)b * p.o_bs + (long long)qrow * p.o_rs + h * kD;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
        reinterpret_cast<uint4*>(op)[j] = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

PerDeviceOnce g_attr_once;

int smem_bytes_for(int Lkp) { return 2048 + kQTile * kD * 2 + 2 * Lkp * kD * 2 + ((Lkp + 63) / 64) * kQTile * 128; }

}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.Lk >= 1 && a.Lk <= kMaxKeys && a.Lq >= 1 && al16(a.q) && al16(a.k) && al16(a.v) && al16(a.out) &&
         a.q_rs % 8 == 0 && a.k_rs % 8 == 0 && a.v_rs % 8 == 0 && a.o_rs % 8 == 0 && a.q_bs % 8 == 0 && a.k_bs % 8 == 0 &&
         a.v_bs % 8 == 0 && a.o_bs % 8 == 0 && a.q_rs >= a.heads * kD && a.k_rs >= a.heads * kD && a.v_rs >= a.heads * kD;
}

int attention_tc_run(const AttnArgs& a, cudaStream_t s) {
  if (!attention_tc_supported(a)) return fail(LMV_ERR_UNSUPPORTED, "attention_tc: unsupported shape / alignment");
  LMV_CUDA_OK(g_attr_once.run([] { return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(kMaxKeys)); }));
  const int Lkp = (a.Lk + 15) & ~15;
  CUtensorMap tq, tk, tv;
  auto enc = [&](CUtensorMap* m, const bf16* base, long long bs, int rs, int rows, int box_rows) {
    uint64_t dims[3] = {(uint64_t)a.heads * kD, (uint64_t)rows, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
    uint32_t box[3] = {kD, (uint32_t)box_rows, 1};
    return encode_tmap_bf16(m, base, 3, dims, strides, box, 64);
  };
  int rc;
  if ((rc = enc(&tq, a.q, a.q_bs, a.q_rs, a.Lq, kQTile))) return rc;
  if ((rc = enc(&tk, a.k, a.k_bs, a.k_rs, a.Lk, Lkp))) return rc;
  if ((rc = enc(&tv, a.v, a.v_bs, a.v_rs, a.Lk, Lkp))) return rc;
  AttnTcParams p;
  p.out = a.out; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.Lq = a.Lq; p.Lk = a.Lk; p.Lkp = Lkp;
  p.o_col = (Lkp + 31) & ~31;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.o_col + kD) p.tmem_cols <<= 1;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  dim3 grid((a.Lq + kQTile - 1) / kQTile, a.heads, a.B);
  LMV_CUDA_OK(launch_kernel(attention_tc_kernel, dim3(grid), dim3(kThreads), (size_t)(smem_bytes_for(Lkp)), s, tq, tk, tv, p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
