// out[M,N] = epilogue(A[M,K] · W[N,K]^T)  — the workhorse of the LeMeViT forward: every nn.Linear
// (reference models/lemevit.py:175,178,241-246,444-448,526-529,730-733,786) and, through im2col,
// every 3x3/s2 convolution (:702,715) runs through this kernel.
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised:
//   warp 0      TMA producer   — cp.async.bulk.tensor 128B-swizzled A (128x64) and W (BNx64) tiles
//   warp 1      MMA issuer     — one thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM
//   warps 2..5  epilogue       — tcgen05.ld accumulators, + bias, GELU, + residual, bf16/fp32 store
// smem ring of `num_stages` (A,W) tiles with full/empty mbarriers; TMEM holds two 256-column
// accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
#include <algorithm>
#include <mutex>

#include "common.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kAccStride = 256;  // TMEM columns per accumulator stage
constexpr int kSmemLimit = 227 * 1024;

struct SmemCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, long long orow,
                                               int col0, bool vec_ok) {
  const int ncols = min(32, p.N - col0);
  if (row >= p.M || ncols <= 0) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const bool full = (ncols == 32) && vec_ok;
  if (p.bias) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
  const long long off = orow * (long long)p.ldc + col0;
  if (p.residual) {
    if (full) {
      const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u = r4[j];  // plain load: residual may alias out (in-place residual stream)
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[8 * j + 0] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
        v[8 * j + 4] += c.x; v[8 * j + 5] += c.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] += __bfloat162float(p.residual[off + j]);
    }
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + off;
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = v[j];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(p.out) + off;
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        reinterpret_cast<uint4*>(o)[j] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) o[j] = __float2bfloat16(v[j]);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  SmemCtrl* ctrl = reinterpret_cast<SmemCtrl*>(smem);
  uint8_t* tiles = smem + 1024;
  const int a_bytes = BM * BK * 2;
  const int stage_bytes = a_bytes + p.BN * BK * 2;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&ctrl->full[i], 1);
      mbar_init(&ctrl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->acc_full[i], 1);
      mbar_init(&ctrl->acc_empty[i], 4);  // one elected lane per epilogue warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&ctrl->empty[stage], phase ^ 1u, 1);
          uint8_t* sa = tiles + (size_t)stage * stage_bytes;
          mbar_expect_tx(&ctrl->full[stage], (uint32_t)stage_bytes);
          tma_load_2d(sa, &tmA, &ctrl->full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + a_bytes, &tmB, &ctrl->full[stage], kb * BK, n_blk * p.BN);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, p.BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&ctrl->acc_empty[as], aphase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStride);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&ctrl->full[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint64_t da = make_kmajor_desc<128>(sa);
          const uint64_t db = make_kmajor_desc<128>(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle span: start-address field += 2
            umma_bf16_ss(d_tmem, da + 2ull * k, db + 2ull * k, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&ctrl->empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&ctrl->acc_full[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool vec_ok = (p.ldc % 8 == 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      mbar_wait(&ctrl->acc_full[as], aphase, 4);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      long long orow = row;
      if (p.grp_rows > 0) orow = (long long)(row / p.grp_rows) * p.grp_stride + (row % p.grp_rows);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccStride);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        if (n_blk * p.BN + c0 >= p.N) break;  // uniform
        uint32_t r[32];
        tmem_ld_x32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        epilogue_chunk(p, r, row, orow, n_blk * p.BN + c0, vec_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->acc_empty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Bring-up / cross-check kernel: one thread per output element, fp32 accumulate.  Tests only.
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw,
                                 GemmParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.N) return;
  const int row = (int)(idx / p.N), col = (int)(idx % p.N);
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k)
    acc += __bfloat162float(A[(long long)row * lda + k]) * __bfloat162float(W[(long long)col * ldw + k]);
  if (p.bias) acc += p.bias[col];
  if (p.act == 1) acc = gelu_erf(acc);
  long long orow = row;
  if (p.grp_rows > 0) orow = (long long)(row / p.grp_rows) * p.grp_stride + (row % p.grp_rows);
  const long long off = orow * p.ldc + col;
  if (p.residual) acc += __bfloat162float(p.residual[off]);
  if (p.out_fp32) reinterpret_cast<float*>(p.out)[off] = acc;
  else reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16(acc);
}

int pick_bn(int N) {
  if (N % 32 == 0) {
    for (int bn = 256; bn >= 32; bn -= 32)
      if (N % bn == 0) return bn;
  }
  return std::min(256, ((N + 31) / 32) * 32);
}

std::once_flag g_attr_once;
cudaError_t g_attr_err = cudaSuccess;

}  // namespace

int gemm_prepare(const GemmArgs& a, GemmOp* op) {
  LMV_REQUIRE(a.A && a.W && a.out, "gemm: null pointer");
  LMV_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  LMV_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: K, lda, ldw must be multiples of 8 (16-byte TMA strides)");
  LMV_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0, "gemm: operands must be 16-byte aligned");
  LMV_REQUIRE((reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "gemm: output must be 16-byte aligned");
  LMV_REQUIRE(a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  GemmParams& p = op->p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.BN = a.force_bn > 0 ? a.force_bn : pick_bn(a.N);
  LMV_REQUIRE(p.BN % 32 == 0 && p.BN >= 32 && p.BN <= 256, "gemm: tile N must be a multiple of 32 in [32,256]");
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + p.BN - 1) / p.BN;
  p.k_blocks = (a.K + BK - 1) / BK;
  const int stage_bytes = BM * BK * 2 + p.BN * BK * 2;
  p.num_stages = std::min(kMaxStages, (kSmemLimit - 2048) / stage_bytes);
  p.bias = a.bias; p.residual = a.residual; p.out = a.out;
  p.ldc = a.ldc; p.out_fp32 = a.out_fp32; p.act = a.act;
  p.grp_rows = a.grp_rows; p.grp_stride = a.grp_stride;
  op->smem_bytes = 2048 + p.num_stages * stage_bytes;
  op->grid = std::min(p.tiles_m * p.tiles_n, device_sm_count());
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t strides[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = encode_tmap_bf16(&op->tmA, a.A, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t strides[1] = {(uint64_t)a.ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)p.BN};
    int rc = encode_tmap_bf16(&op->tmB, a.W, 2, dims, strides, box, 128);
    if (rc) return rc;
  }
  return LMV_OK;
}

int gemm_run(const GemmOp& op, cudaStream_t stream) {
  std::call_once(g_attr_once, [] {
    g_attr_err = cudaFuncSetAttribute(gemm_bf16_tn_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  });
  LMV_CUDA_OK(g_attr_err);
  gemm_bf16_tn_tcgen05<<<op.grid, kThreads, op.smem_bytes, stream>>>(op.tmA, op.tmB, op.p);
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int gemm_simt_run(const GemmArgs& a, cudaStream_t stream) {
  GemmParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.ldc = a.ldc; p.out_fp32 = a.out_fp32; p.act = a.act;
  p.grp_rows = a.grp_rows; p.grp_stride = a.grp_stride;
  const long long total = (long long)a.M * a.N;
  gemm_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a.A, a.lda, a.W, a.ldw, p);
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
