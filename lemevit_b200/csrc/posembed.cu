// x' = x + dwconv3x3(x) (+ per-row LayerNorm statistics) on token-major [B, T, C] bf16 activations.
// Reference: LeMeBlock.pos_embed (models/lemevit.py:510) applied at :546,589,619, followed by norm1 (:513) whose
// normalisation is folded into the consuming GEMM (gemm.cu) from the (sum, sum^2) emitted here.
//
// sm_100a design (memory-bound op, roofline = read x + write x'): persistent CTAs over (TH x TW) output tiles of the images.
//   * one elected thread issues a 4-D TMA load {C, TW+2, TH+2, 1} of the input tile + halo into shared memory; the
//     tensor map spans a single image plane {C, W, H, B}, so the zero padding of the convolution is TMA's out-of-bounds
//     fill (negative / overflowing coordinates) — no border branches in the kernel;
//   * thread = (tile column, 8-channel vector) streaming the input rows of its column once (3 x LDS.128 per row feed up to three
//     output rows through rotating accumulators), depthwise taps in registers, packed fp32 FMAs, one 16-byte store per output;
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
  float* stats;        // [B*T][parts][2] or null
  int B, H, W, T, C;
  int TW, TH, tiles_x, ntiles, n_items;
  int col_threads;     // threads that own (tile column, channel vector) pairs: a multiple of V = CS / 8
  int V, tx_step;      // 16-byte channel vectors of a slice; columns a thread advances by when the tile has more pairs than threads
  int cbox, ncb, vpb;  // channel box of one TMA load, boxes per CTA (CS = cbox * ncb), 16-byte vectors per box
  int CS, parts;       // channels per CTA (blockIdx.z selects the slice) and slices per row; C = CS * parts
  int sub_bytes;       // shared-memory bytes of one channel box of the input tile (128-byte multiple: TMA destination)
  int in_bytes;        // one input tile: ncb * sub_bytes
  int row_pitch, col_pitch;   // bytes between tile rows / tile columns inside a channel box
  int part_elems;      // TH * TW * V statistics partials per buffer
  uint32_t tx_bytes;   // bytes one tile load delivers
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

// Persistent CTAs (grid.x strides over the (image, tile) items of one channel slice) with a double-buffered input tile: the TMA
// load of item i+1 is in flight while item i is computed, and the depthwise taps are fetched once per CTA instead of once per tile.
// Thread = (column tx of the tile, 8-channel vector v); it streams the input rows of its column ONCE: every input row (3 x LDS.128,
// unpacked once) feeds the (up to) three output rows it touches through three rotating accumulators.  The row loop is peeled so
// that no tap is computed for an output row outside the tile (first / last two input rows) and all addressing is three running
// pointers: per input row and thread 3 LDS + 24 unpack + <= 36 packed FMAs, per output row 4 packs + one 16-byte store + ~16
// packed-fp32 instructions for the row statistics.  (The first persistent version spent 2/3 of its issue slots on predicates and
// 64-bit index arithmetic: profiles/r02_ncu_posembed_c384.txt.)  The FMA order per output element is bias, then taps 0..8: the
// per-token kernel (tokens.cu) rounds identically.
__global__ void __launch_bounds__(kThreads)
posembed_tile_kernel(const __grid_constant__ CUtensorMap tm, const PosTileParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);           // full[2]
  const int C = p.C, V = p.V, HW = p.H * p.W;
  const int slice = blockIdx.z, c_off = slice * p.CS;
  pdl_launch_dependents();

  if (blockIdx.y == 1) {
    // ---- meta-token rows of a unified buffer: copy + statistics, one warp per row ----
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;          // blockDim need not be a multiple of 32: use the full warps only
    if (warp >= nwarps || slice != 0) return;   // slice 0 copies the whole row; the other partials of the row are zero
    for (int b = blockIdx.x; b < p.B; b += gridDim.x)
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

  uint8_t* s_in = smem + 128;                                                // [2][in_bytes]
  float2* s_part = reinterpret_cast<float2*>(s_in + 2 * (size_t)p.in_bytes);  // [2][TH*TW][V] partial statistics
  const int ntiles = p.ntiles, n_items = p.n_items;
  auto issue = [&](int item, int buf) {
    const int b = item / ntiles, t = item - b * ntiles;
    const int tile_y = t / p.tiles_x, tile_x = t - tile_y * p.tiles_x;
    mbar_expect_tx(&bar[buf], p.tx_bytes);
    for (int cb = 0; cb < p.ncb; ++cb)
      tma_load_4d(s_in + (size_t)buf * p.in_bytes + (size_t)cb * p.sub_bytes, &tm, &bar[buf], c_off + cb * p.cbox, tile_x * p.TW - 1, tile_y * p.TH - 1, b);
  };
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  // thread = (tile column, channel vector): its 9 x 8 depthwise taps and 8 biases live in registers as packed pairs (weights are
  // constants of the plan, not outputs of the previous kernel: fetched before the grid dependency resolves).  When the tile has more
  // (column, vector) pairs than the CTA has threads, a thread keeps its vector and walks columns tx0, tx0 + tx_step, ..
  const bool active = (int)threadIdx.x < p.col_threads;
  const int tx0 = active ? (int)threadIdx.x / V : 0, v = active ? (int)threadIdx.x - tx0 * V : 0;
  float2 w[9][4], bias[4];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.dw_w + tap * C + c_off + v * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.dw_w + tap * C + c_off + v * 8) + 1);
    w[tap][0] = make_float2(w0.x, w0.y); w[tap][1] = make_float2(w0.z, w0.w);
    w[tap][2] = make_float2(w1.x, w1.y); w[tap][3] = make_float2(w1.z, w1.w);
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c_off + v * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c_off + v * 8) + 1);
    bias[0] = make_float2(b0.x, b0.y); bias[1] = make_float2(b0.z, b0.w); bias[2] = make_float2(b1.x, b1.y); bias[3] = make_float2(b1.z, b1.w);
  }
  const int cb = v / p.vpb, vv = v - cb * p.vpb;
  const uint32_t thr_off = (uint32_t)(cb * p.sub_bytes + vv * 16);      // this thread's vector inside an input tile, column 0
  const uint32_t row_pitch = p.row_pitch, col_pitch = p.col_pitch;
  const long long out_row_step = (long long)p.W * C;                    // elements between vertically adjacent tokens
  const int part_row_step = p.TW * V;
  const int n_tok = p.TW * p.TH;
  const int tok_ty0 = (int)threadIdx.x / p.TW, tok_tx0 = (int)threadIdx.x - tok_ty0 * p.TW;   // first token of the statistics pass
  __syncthreads();          // barrier init visible to every waiter
  pdl_wait();
  if (threadIdx.x == 0 && (int)blockIdx.x < n_items) issue(blockIdx.x, 0);

  int it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int buf = it & 1;
    // prefetch the next item into the other buffer: its previous reader (item it - 1) finished before the trailing __syncthreads
    if (threadIdx.x == 0 && item + (int)gridDim.x < n_items) issue(item + gridDim.x, buf ^ 1);
    const int b = item / ntiles, t = item - b * ntiles;
    const int tile_y = t / p.tiles_x, tile_x = t - tile_y * p.tiles_x;
    const int x0 = tile_x * p.TW, y0 = tile_y * p.TH;
    const int nrows = min(p.TH, p.H - y0);                 // output rows of this tile (>= 1)
    const long long tok0 = (long long)b * p.T + (long long)y0 * p.W + x0;
    mbar_wait(&bar[buf], (uint32_t)(it >> 1) & 1u, 20);
    float2* part = s_part + (size_t)buf * p.part_elems;
    if (active)
      for (int tx = tx0; tx < p.TW && x0 + tx < p.W; tx += p.tx_step) {
        const uint8_t* sp = s_in + (size_t)buf * p.in_bytes + thr_off + (uint32_t)tx * col_pitch;   // input row ir, column tx (kx = 0)
        bf16* op = p.out + (tok0 + tx) * C + c_off + v * 8;                                          // output row ty
        float2* pp = part + tx * V + v;
        float2 in[3][4];
        auto load = [&]() {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint4 u = *reinterpret_cast<const uint4*>(sp + (uint32_t)kx * col_pitch);
            in[kx][0] = unpack2(u.x); in[kx][1] = unpack2(u.y); in[kx][2] = unpack2(u.z); in[kx][3] = unpack2(u.w);
          }
          sp += row_pitch;
        };
        auto top = [&](float2 (&acc)[4]) {       // ky = 0: this input row is the row above the output row
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = ffma2(in[0][j], w[0][j], bias[j]);
            acc[j] = ffma2(in[1][j], w[1][j], acc[j]);
            acc[j] = ffma2(in[2][j], w[2][j], acc[j]);
          }
        };
        auto mid = [&](float2 (&acc)[4]) {       // ky = 1
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = ffma2(in[0][j], w[3][j], acc[j]);
            acc[j] = ffma2(in[1][j], w[4][j], acc[j]);
            acc[j] = ffma2(in[2][j], w[5][j], acc[j]);
          }
        };
        auto bot = [&](float2 (&acc)[4]) {       // ky = 2: completes the output row, which is stored with its statistics
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = ffma2(in[0][j], w[6][j], acc[j]);
            acc[j] = ffma2(in[1][j], w[7][j], acc[j]);
            acc[j] = ffma2(in[2][j], w[8][j], acc[j]);
          }
          uint4 pk;
          pk.x = pack_bf16x2(acc[0].x, acc[0].y); pk.y = pack_bf16x2(acc[1].x, acc[1].y);
          pk.z = pack_bf16x2(acc[2].x, acc[2].y); pk.w = pack_bf16x2(acc[3].x, acc[3].y);
          *reinterpret_cast<uint4*>(op) = pk;
          op += out_row_step;
          if (p.stats) {
            // statistics of the STORED (bf16-rounded) values: exactly what the consuming GEMM reads
            const float2 f0 = unpack2(pk.x), f1 = unpack2(pk.y), f2 = unpack2(pk.z), f3 = unpack2(pk.w);
            const float2 s = fadd2(fadd2(f0, f1), fadd2(f2, f3));
            float2 q = fmul2(f0, f0);
            q = ffma2(f1, f1, q); q = ffma2(f2, f2, q); q = ffma2(f3, f3, q);
            *pp = make_float2(s.x + s.y, q.x + q.y);
            pp += part_row_step;
          }
        };
        float2 a0[4], a1[4], a2[4];
        // input rows nrows, nrows + 1 (M: accumulator whose output row got its ky = 0 taps last, O: the one before)
        auto tail = [&](float2 (&M)[4], float2 (&O)[4]) {
          load();
          if (nrows >= 2) bot(O);
          mid(M);
          load();
          bot(M);
        };
        load(); top(a0);
        if (nrows == 1) {
          tail(a0, a1);
        } else {
          load(); top(a1); mid(a0);
          int ir = 2;
          while (true) {
            if (ir >= nrows) { tail(a1, a0); break; }
            load(); top(a2); mid(a1); bot(a0); ++ir;
            if (ir >= nrows) { tail(a2, a1); break; }
            load(); top(a0); mid(a2); bot(a1); ++ir;
            if (ir >= nrows) { tail(a0, a2); break; }
            load(); top(a1); mid(a0); bot(a2); ++ir;
          }
        }
      }
    __syncthreads();
    if (p.stats) {
      int ty = tok_ty0, txx = tok_tx0;
      for (int tok = threadIdx.x; tok < n_tok; tok += blockDim.x) {
        if (tok != (int)threadIdx.x) { ty = tok / p.TW; txx = tok - ty * p.TW; }
        if (x0 + txx >= p.W || ty >= nrows) continue;
        float s1 = 0.f, s2 = 0.f;
        const float2* pq = part + tok * V;
        for (int q = 0; q < V; ++q) { s1 += pq[q].x; s2 += pq[q].y; }
        const long long row = tok0 + (long long)ty * p.W + txx;
        *reinterpret_cast<float2*>(p.stats + 2 * (row * p.parts + slice)) = make_float2(s1, s2);
      }
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
  TH = (a.H + (a.H + TH - 1) / TH - 1) / ((a.H + TH - 1) / TH);      // equal-height tiles (14 rows -> 7 + 7, not 8 + 6)
  op->TW = TW; op->TH = TH;
  op->tiles_x = (a.W + TW - 1) / TW;
  op->tiles_y = (a.H + TH - 1) / TH;
  op->sub_bytes = (((TH + 2) * (TW + 2) * op->cbox * 2 + 127) / 128) * 128;
  LMV_REQUIRE(CS / 8 <= kThreads, "posembed: channel slice wider than the CTA");
  // one thread per (tile column, 8-channel vector) when that fits the CTA, else a multiple of V threads that walk the columns
  op->col_threads = std::min(TW, kThreads / (CS / 8)) * (CS / 8);
  op->threads = ((op->col_threads + 31) / 32) * 32;
  op->smem = 128 + 128 + 2 * op->ncb * op->sub_bytes + 2 * TH * TW * (CS / 8) * 8;
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
  p.TW = op.TW; p.TH = op.TH; p.tiles_x = op.tiles_x; p.cbox = op.cbox; p.ncb = op.ncb;
  p.sub_bytes = op.sub_bytes; p.parts = op.parts; p.CS = a.C / op.parts; p.B = a.B; p.col_threads = op.col_threads;
  p.ntiles = op.tiles_x * op.tiles_y; p.n_items = p.ntiles * a.B;
  p.V = p.CS / 8; p.tx_step = op.col_threads / p.V; p.vpb = op.cbox / 8;
  p.in_bytes = op.ncb * op.sub_bytes;
  p.col_pitch = op.cbox * 2; p.row_pitch = (op.TW + 2) * p.col_pitch;
  p.part_elems = op.TH * op.TW * p.V;
  p.tx_bytes = (uint32_t)(op.ncb * (op.TH + 2) * (op.TW + 2) * op.cbox * 2);
  // persistent CTAs: as many as can be resident (shared memory / registers allow 2-3 per SM), striding over the (image, tile) items;
  // blockIdx.y == 1: the CTAs that pass the meta-token rows of a unified buffer through
  const int n_items = op.tiles_x * op.tiles_y * a.B;
  const int per_sm = std::max(1, std::min(3, (200 * 1024) / std::max(op.smem, 1)));
  // CTAs per resident slot, measured per shape at Base b256: unsliced (C = 96) and two slices (C = 192) are fastest with one
  // wave of long-lived CTAs, four slices (C = 384, 512: 7-row tiles, 1-2 items per CTA anyway) with three.
  int waves = op.parts >= 4 ? 3 : 1;
  if (const char* e = getenv("LMV_POS_WAVES")) waves = std::max(1, atoi(e));   // A/B knob
  dim3 grid(std::min(n_items, std::max(1, waves * per_sm * device_sm_count() / op.parts)), a.T > a.H * a.W ? 2 : 1, op.parts);
  LMV_CUDA_OK(launch_kernel(posembed_tile_kernel, dim3(grid), dim3(op.threads), (size_t)(op.smem), s, op.tm, p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
