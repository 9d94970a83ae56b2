// x' = x + dwconv3x3(x) (+ per-row LayerNorm statistics) on token-major [B, T, C] bf16 activations.
// Reference: LeMeBlock.pos_embed (models/lemevit.py:510) applied at :546,589,619, followed by norm1 (:513) whose
// normalisation is folded into the consuming GEMM (gemm.cu) from the (sum, sum^2) emitted here.
//
// sm_100a design (memory-bound op, roofline = read x + write x'): one CTA per (TH x TW) output tile of one image.
//   * one elected thread issues a 4-D TMA load {C, TW+2, TH+2, 1} of the input tile + halo into shared memory; the
//     tensor map spans a single image plane {C, W, H, B}, so the zero padding of the convolution is TMA's out-of-bounds
//     fill (negative / overflowing coordinates) — no border branches in the kernel;
//   * work item = (token, 8-channel vector): 9 x LDS.128 of activations (consecutive lanes -> consecutive 16 B, conflict
//     free), depthwise taps in registers, fp32 accumulate, one coalesced 16-byte store;
//   * statistics are reduced in a fixed order through shared memory (deterministic, no atomics).
// Meta-token rows (t >= H*W of a unified [B, N+M, C] buffer) are passed through by one extra CTA per image.
#include <algorithm>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kThreads = 256;

struct PosTileParams {
  const bf16* tokens;
  const float* dw_w;   // [9][C]
  const float* dw_b;   // [C]
  bf16* out;           // [B, T, C]
  float* stats;        // [B*T][2] or null
  int H, W, T, C;
  int TW, TH, tiles_x, tiles_y;
  int cbox, ncb;       // channel box of one TMA load, number of boxes per CTA (CS = cbox * ncb)
  int CS, parts;       // channels per CTA (blockIdx.z selects the slice) and slices per row; C = CS * parts
  int sub_bytes;       // shared-memory bytes of one channel box of the input tile (128-byte multiple: TMA destination)
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

__global__ void __launch_bounds__(kThreads)
posembed_tile_kernel(const __grid_constant__ CUtensorMap tm, const PosTileParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  const int C = p.C, V = p.CS >> 3, HW = p.H * p.W;   // V: 16-byte channel vectors of this CTA's slice
  const int b = blockIdx.y, slice = blockIdx.z, c_off = slice * p.CS;
  const int ntiles = p.tiles_x * p.tiles_y;
  pdl_launch_dependents();
  pdl_wait();

  if ((int)blockIdx.x >= ntiles) {
    // ---- meta-token rows of a unified buffer: copy + statistics, one warp per row ----
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;          // blockDim need not be a multiple of 32: use the full warps only
    if (warp >= nwarps || slice != 0) return;   // slice 0 copies the whole row; the other partials of the row are zero
    for (int t = HW + warp; t < p.T; t += nwarps) {
      const long long row = (long long)b * p.T + t;
      float s1 = 0.f, s2 = 0.f;
      for (int v = lane; v < (C >> 3); v += 32) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.tokens + row * C) + v);
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
        reinterpret_cast<uint4*>(p.out + row * C)[v] = u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane < p.parts && p.stats)
        *reinterpret_cast<float2*>(p.stats + 2 * (row * p.parts + lane)) = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
    }
    return;
  }

  const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x % p.tiles_x;
  const int x0 = tile_x * p.TW, y0 = tile_y * p.TH;
  const int IW = p.TW + 2, IH = p.TH + 2;
  const int sub_bytes = p.sub_bytes;                             // one channel box of the input tile
  uint8_t* s_tile = smem + 128;
  float2* s_part = reinterpret_cast<float2*>(s_tile + (size_t)p.ncb * sub_bytes);   // [TH*TW][V] partial statistics
  (void)IH;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_expect_tx(bar, (uint32_t)(p.ncb * IH * IW * p.cbox * 2));
    for (int cb = 0; cb < p.ncb; ++cb) tma_load_4d(s_tile + (size_t)cb * sub_bytes, &tm, bar, c_off + cb * p.cbox, x0 - 1, y0 - 1, b);
  }
  // blockDim is a multiple of V, so every thread keeps ONE channel vector for all its items: its 9 x 8 depthwise taps
  // and 8 biases live in registers (shared-memory bandwidth is spent on activations only)
  const int v = threadIdx.x % V;
  float w[9][8], bias[8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.dw_w + tap * C + c_off + v * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.dw_w + tap * C + c_off + v * 8) + 1);
    w[tap][0] = w0.x; w[tap][1] = w0.y; w[tap][2] = w0.z; w[tap][3] = w0.w;
    w[tap][4] = w1.x; w[tap][5] = w1.y; w[tap][6] = w1.z; w[tap][7] = w1.w;
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c_off + v * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c_off + v * 8) + 1);
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w; bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  __syncthreads();          // barrier init visible to every waiter
  mbar_wait(bar, 0, 20);

  const int vpb = p.cbox >> 3;   // 16-byte vectors per channel box
  const int items = p.TW * p.TH * V;
  // blockDim is a multiple of V: a thread's token index advances by blockDim / V per trip, so (tx, ty) are stepped
  // instead of divided out of the item index
  const int tok_step = blockDim.x / V;
  int tx = (threadIdx.x / V) % p.TW, ty = (threadIdx.x / V) / p.TW;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int x = x0 + tx, y = y0 + ty;
    float2 part = make_float2(0.f, 0.f);
    if (x < p.W && y < p.H) {
      const int cb = v / vpb, vv = v - cb * vpb;
      const uint8_t* base = s_tile + (size_t)cb * sub_bytes + (size_t)vv * 16;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bias[j];
      uint4 u[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          u[ky * 3 + kx] = *reinterpret_cast<const uint4*>(base + (size_t)((ty + ky) * IW + tx + kx) * (p.cbox * 2));
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        float f[8];
        unpack8(u[tap], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], w[tap][j], acc[j]);
      }
      uint4 pk;
      pk.x = pack_bf16x2(acc[0], acc[1]); pk.y = pack_bf16x2(acc[2], acc[3]);
      pk.z = pack_bf16x2(acc[4], acc[5]); pk.w = pack_bf16x2(acc[6], acc[7]);
      const long long row = (long long)b * p.T + (long long)y * p.W + x;
      reinterpret_cast<uint4*>(p.out + row * C + c_off)[v] = pk;
      // statistics of the STORED (bf16-rounded) values: exactly what the consuming GEMM reads
      unpack8(pk, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) { part.x += acc[j]; part.y = fmaf(acc[j], acc[j], part.y); }
    }
    if (p.stats) s_part[i] = part;
    tx += tok_step;
    while (tx >= p.TW) { tx -= p.TW; ++ty; }
  }
  if (p.stats) {
    __syncthreads();
    for (int tok = threadIdx.x; tok < p.TW * p.TH; tok += blockDim.x) {
      const int tx = tok % p.TW, ty = tok / p.TW;
      const int x = x0 + tx, y = y0 + ty;
      if (x >= p.W || y >= p.H) continue;
      float s1 = 0.f, s2 = 0.f;
      for (int vv = 0; vv < V; ++vv) {
        const float2 q = s_part[tok * V + vv];
        s1 += q.x; s2 += q.y;
      }
      const long long row = (long long)b * p.T + (long long)y * p.W + x;
      *reinterpret_cast<float2*>(p.stats + 2 * (row * p.parts + slice)) = make_float2(s1, s2);
    }
  }
}

}  // namespace

bool posembed_tile_supported(const PosLnArgs& a) {
  return a.dw_w && a.dw_b && a.resid_out && !a.norm_out && a.C % 8 == 0 && a.C <= 512 && a.H >= 1 && a.W >= 1 &&
         (reinterpret_cast<uintptr_t>(a.tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.resid_out) & 15) == 0 &&
         a.tokens != a.resid_out;
}

int posembed_tile_prepare(const PosLnArgs& a, PosEmbedOp* op) {
  if (!posembed_tile_supported(a)) return fail(LMV_ERR_UNSUPPORTED, "posembed_tile: unsupported arguments");
  LMV_REQUIRE(a.T >= a.H * a.W, "posembed: T < H*W");
  op->a = a;
  const int C = a.C;
  // Channel slices: a CTA that carries all C channels of a wide stage can only afford a 1-2 row tile (halo re-read 2-3.4x
  // through L2 / TMA); slicing the channels over up to max_parts CTAs keeps the tile 8 rows tall.  Each slice emits its
  // own statistics partial, which the consuming GEMM sums (ln_parts <= 4).
  int parts = 1;
  for (int n = 1; n <= std::min(a.max_parts, 4); ++n)
    if (C % (8 * n) == 0) {
      parts = n;
      if (C / n <= 96) break;
    }
  op->parts = parts;
  const int CS = C / parts;
  op->ncb = CS > 256 ? 2 : 1;
  op->cbox = CS / op->ncb;
  LMV_REQUIRE(op->cbox % 8 == 0, "posembed: channel box must be a multiple of 8");
  // tile: <= 16 wide, tall enough to amortise the halo, input tile + halo within ~40 KB so several CTAs share an SM
  int TW = a.W <= 16 ? a.W : ((a.W % 14 == 0) ? 14 : 16);
  int TH = (40 * 1024) / ((TW + 2) * CS * 2) - 2;
  TH = std::max(1, std::min(std::min(TH, 8), a.H));
  op->TW = TW; op->TH = TH;
  op->tiles_x = (a.W + TW - 1) / TW;
  op->tiles_y = (a.H + TH - 1) / TH;
  op->sub_bytes = (((TH + 2) * (TW + 2) * op->cbox * 2 + 127) / 128) * 128;
  op->threads = (CS / 8) * (kThreads / (CS / 8));     // multiple of V = CS/8, <= 256
  op->smem = 128 + 128 + op->ncb * op->sub_bytes + TH * TW * (CS / 8) * 8;
  LMV_REQUIRE(op->smem <= 200 * 1024, "posembed: tile does not fit in shared memory");
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)a.W * C * 2, (uint64_t)a.T * C * 2};
  uint32_t box[4] = {(uint32_t)op->cbox, (uint32_t)(TW + 2), (uint32_t)(TH + 2), 1};
  return encode_tmap_bf16(&op->tm, a.tokens, 4, dims, strides, box, 0);
}

int posembed_tile_run(const PosEmbedOp& op, cudaStream_t s) {
  const PosLnArgs& a = op.a;
  if (a.B == 0) return LMV_OK;
  static PerDeviceOnce attr_once;
  LMV_CUDA_OK(attr_once.run([] { return cudaFuncSetAttribute(posembed_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
  PosTileParams p;
  p.tokens = a.tokens; p.dw_w = a.dw_w; p.dw_b = a.dw_b; p.out = a.resid_out; p.stats = a.stats_out;
  p.H = a.H; p.W = a.W; p.T = a.T; p.C = a.C;
  p.TW = op.TW; p.TH = op.TH; p.tiles_x = op.tiles_x; p.tiles_y = op.tiles_y; p.cbox = op.cbox; p.ncb = op.ncb;
  p.sub_bytes = op.sub_bytes; p.parts = op.parts; p.CS = a.C / op.parts;
  dim3 grid(op.tiles_x * op.tiles_y + (a.T > a.H * a.W ? 1 : 0), a.B, op.parts);
  LMV_CUDA_OK(launch_kernel(posembed_tile_kernel, dim3(grid), dim3(op.threads), (size_t)(op.smem), s, op.tm, p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
