// out[M,N] = epilogue(A[M,K] · W[N,K]^T)  — the workhorse of the LeMeViT forward: every nn.Linear
// (reference models/lemevit.py:175,178,241-246,444-448,526-529,730-733,786) and every 3x3/s2 convolution (:702,715)
// runs through this kernel.
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised:
//   warp 0      TMA producer   — cp.async.bulk.tensor 128B-swizzled A (128x64) and W (BNx64) tiles; for the convolutions
//                                (GemmArgs::conv_*) the A box is a strided 4-D gather straight from the activation
//                                (implicit GEMM: one box per (tap, 64-channel slice), padding by out-of-bounds fill)
//   warp 1      MMA issuer     — one thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM
//   warps 2..9  epilogue       — two warps per TMEM lane quarter (they split the 32-column chunks).  Three forms:
//                                * lane = row (row_epi): tcgen05.ld -> [LayerNorm fold] + bias -> [GELU] -> bf16 row into a
//                                  64B-swizzled staging box -> TMA store;
//                                * lane = row with in-place residual (row_res): the residual box is TMA-loaded one chunk
//                                  ahead into the other staging box, added in place, stored back by TMA, + row statistics;
//                                * transpose (everything else: convolutions' short tiles, fp32 / strided outputs, the head):
//                                  per-warp smem transpose -> coalesced (+ residual) stores or TMA boxes.
// smem ring of `num_stages` (A,W) tiles with full/empty mbarriers (W resident beside an A-only ring when there is one N tile
// and it fits: w_res); TMEM holds two 256-column accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Fusions carried by the epilogue (north star: "LN fused into the following projection, bias+GELU fused into the
// MLP GEMM"): LayerNorm (norm1 -> q/kv/qkv*, norm2 -> mlp.0; models/lemevit.py:560-564,600-601,632-635) is folded
// algebraically: LN(y) W'^T = r (y W'^T - mu colsum(W')), with the row statistics (sum y, sum y^2) produced by the
// kernel that wrote y (posembed kernel, or the proj GEMM's own epilogue through `stats_out`).
#include <algorithm>
#include <mutex>

#include "common.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStageF32 = 32 * 32 * 4;   // per-warp transpose buffer: 32 rows x 32 fp32
constexpr int kStageOut = 2 * 32 * 64;   // per-warp output staging of the TMA-store epilogue: two 32-row x 64-byte (32 bf16) boxes
constexpr int kAccStride = 256;  // TMEM columns per accumulator stage
constexpr int kSmemLimit = 227 * 1024;

struct SmemCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t w_full;      // resident W (p.w_res): all its K blocks have landed
  uint64_t res_full[kEpiWarps][2];   // lane = row residual epilogue (p.row_res): the warp's residual box has landed in its staging buffer
  uint32_t tmem_base;
};

// Generic (slow-path) chunk epilogue: one thread owns one output row; handles ragged N, fp32 output, odd ldc.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, long long orow,
                                               int col0, bool vec_ok, float ln_r, float ln_nrm) {
  const int ncols = min(32, p.N - col0);
  if (row >= p.M || ncols <= 0) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const bool full = (ncols == 32) && vec_ok;
  if (p.ln_stats) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] = fmaf(ln_r, v[j], ln_nrm * __ldg(p.ln_colsum + col0 + j));
  }
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
    for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
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

// Debug build (-DLMV_GEMM_TRACE): per-warp cycle accounting of the three roles, read back with lmv_debug_gemm_trace().
#ifdef LMV_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[148 * 10 * 8];
#define TR_INIT unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64(); const long long tr_t0 = tr_t;
#define TR(i) { const long long now_ = clock64(); tr[i] += (unsigned long long)(now_ - tr_t); tr_t = now_; }
#define TR_FLUSH { tr[7] = (unsigned long long)(clock64() - tr_t0); if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_gemm_trace[((size_t)blockIdx.x * 10 + warp) * 8 + i_] += tr[i_]; }
#else
#define TR_INIT
#define TR(i)
#define TR_FLUSH
#endif

// EPI >= 0: the fast-path epilogue features are fixed at compile time (bit set below: less code, fewer live registers,
// no speculated residual unpack); EPI = -1 decides them at run time from GemmParams (any other combination).
constexpr int kEpiLn = 1, kEpiGelu = 2, kEpiRes = 4, kEpiStats = 8;

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  SmemCtrl* ctrl = reinterpret_cast<SmemCtrl*>(smem);
  // [ctrl 1 KB][8 x 4 KB transpose buffers][8 x 4 KB output staging (TMA-store epilogue)][resident W: k_blocks x BN x 128 B (w_res)][stage ring]
  // (row_epi: the lane = row epilogue needs no transpose buffers)
  const int stg_bytes = p.row_epi ? 0 : kEpiWarps * kStageF32;
  uint8_t* wres = smem + 1024 + stg_bytes + (p.tma_out ? kEpiWarps * kStageOut : 0);
  const int a_bytes = BM * BK * 2, w_bytes = p.BN * BK * 2;
  uint8_t* tiles = wres + (p.w_res ? (size_t)p.k_blocks * w_bytes : 0);
  const int stage_bytes = p.stage_bytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: the role branches stay uniform
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&ctrl->full[i], 1);
      mbar_init(&ctrl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->acc_full[i], 1);
      mbar_init(&ctrl->acc_empty[i], kEpiWarps);  // one elected lane per epilogue warp
    }
    mbar_init(&ctrl->w_full, 1);
    for (int i = 0; i < kEpiWarps; ++i) {
      mbar_init(&ctrl->res_full[i][0], 1);
      mbar_init(&ctrl->res_full[i][1], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmC);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  const uint32_t tmem_base = ctrl->tmem_base;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TR_INIT
      if (p.w_res && (int)blockIdx.x < num_tiles) {
        // one N tile and a W small enough to stay: every K block of it is loaded once per CTA (the strided convolution of the
        // stem otherwise streams as many W bytes as A bytes per tile through the L2 -> SM path that bounds it)
        mbar_expect_tx(&ctrl->w_full, (uint32_t)(p.k_blocks * w_bytes));
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (p.conv) tma_load_3d(wres + (size_t)kb * w_bytes, &tmB, &ctrl->w_full, (kb % p.conv_cpt) * BK, kb / p.conv_cpt, 0);
          else tma_load_2d(wres + (size_t)kb * w_bytes, &tmB, &ctrl->w_full, kb * BK, 0);
        }
      }
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
        // pull the residual tile into L2 while the mainloop of this tile runs: the epilogue reads it ~2 tiles later
        if (p.prefetch_res) tma_prefetch_l2_2d(&tmR, n_blk * p.BN, m_blk * BM);
        // implicit convolution: the tile's first image / output row; K block kb = (tap, 64-channel slice)
        int cb0 = 0, cy0 = 0;
        if (p.conv) {
          if (p.conv_bb > 1) { cb0 = m_blk * p.conv_bb; }
          else { cb0 = m_blk / p.conv_tpi; cy0 = (m_blk - cb0 * p.conv_tpi) * p.conv_bh; }
        }
        int tap = 0, cs = 0;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          TR(1)
          mbar_wait(&ctrl->empty[stage], phase ^ 1u, 1);
          TR(0)
          uint8_t* sa = tiles + (size_t)stage * stage_bytes;
          mbar_expect_tx(&ctrl->full[stage], p.tx_bytes);
          if (p.conv) {
            // A: input pixels (2 oy + ky - 1, 2 ox + kx - 1) of the tile's output pixels: a stride-2 box whose out-of-range
            // coordinates (the padding ring, channels beyond C in the last slice of a tap) arrive as zeros; W: the same slice of
            // tap `tap` from the [N][9][C] view of the packed weights (zero beyond C as well, so the padded k contribute nothing)
            const int ky = tap / 3, kx = tap - 3 * ky;
            tma_load_4d(sa, &tmA, &ctrl->full[stage], cs * BK, kx - 1, 2 * cy0 + ky - 1, cb0);
            if (!p.w_res) tma_load_3d(sa + a_bytes, &tmB, &ctrl->full[stage], cs * BK, tap, n_blk * p.BN);
            if (++cs == p.conv_cpt) { cs = 0; ++tap; }
          } else {
            tma_load_2d(sa, &tmA, &ctrl->full[stage], kb * BK, m_blk * BM);
            if (!p.w_res) tma_load_2d(sa + a_bytes, &tmB, &ctrl->full[stage], kb * BK, n_blk * p.BN);
          }
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
      TR_FLUSH
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // the whole warp runs the (uniform) control flow and one elected lane issues: descriptors and loop state stay in uniform
    // registers (umma.cuh: under `if (lane == 0)` every tcgen05.mma drags a vote / elect / branch sequence behind it)
    {
      const uint32_t idesc = make_idesc_bf16(BM, p.BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      TR_INIT
      if (p.w_res && (int)blockIdx.x < num_tiles) mbar_wait(&ctrl->w_full, 0u, 5);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        TR(2)
        mbar_wait(&ctrl->acc_empty[as], aphase ^ 1u, 2);
        TR(0)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStride);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          TR(2)
          mbar_wait(&ctrl->full[stage], phase, 3);
          TR(1)
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint64_t da = make_kmajor_desc<128>(sa);
          const uint64_t db = make_kmajor_desc<128>(p.w_res ? smem_u32(wres + (size_t)kb * w_bytes) : sa + a_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle span: start-address field += 2
            umma_bf16_ss_warp(d_tmem, da + 2ull * k, db + 2ull * k, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit_warp(&ctrl->empty[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_warp(&ctrl->acc_full[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
      TR_FLUSH
    }
  } else {
    // ---------------- epilogue (warps 2..9) ----------------
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // the two warps of a quarter take alternate 32-column chunks
    const bool f_ln = EPI < 0 ? (p.ln_stats != nullptr) : ((EPI & kEpiLn) != 0);
    const bool f_gelu = EPI < 0 ? (p.act == 1) : ((EPI & kEpiGelu) != 0);
    const bool f_res = EPI < 0 ? (p.residual != nullptr) : ((EPI & kEpiRes) != 0);
    const bool f_stats = EPI < 0 ? (p.stats_out != nullptr) : ((EPI & kEpiStats) != 0);
    constexpr bool kPackedMath = EPI >= 0;
    // the residual + statistics variant has no registers to spare for a second accumulator chunk in flight
    constexpr bool kPipelineLd = !(EPI >= 0 && (EPI & kEpiRes) != 0 && (EPI & kEpiStats) != 0);
    // Lane = row epilogue (no transpose): variants without residual / statistics whose output leaves through TMA tile stores
    // (p.row_epi).  The transpose path moves every accumulator chunk through shared memory twice (fp32 write + read, 2 x 98 KB per
    // 128 x 192 tile) next to the 240 KB of operand reads + 240 KB of TMA fills the tensor pipe needs for the same tile (K = 384).
    // Here a lane keeps its row: LayerNorm fold with the lane's own (r, -r mu), per-column constants by warp-uniform loads, and the
    // only epilogue traffic left in shared memory is the bf16 box the TMA store reads (49 KB per tile).  Same-box A/B at Base b256:
    // qkv 74.6 -> 67.8 us, stage-4 qkv 40.3 -> 38.9, stage-4 fc1 57.4 -> 55.3 (LMV_GEMM_ROW_EPI=0 switches it off).
    constexpr bool kRowEpi = EPI >= 0 && (EPI & (kEpiRes | kEpiStats)) == 0;
    const bool vec_ok = (p.ldc % 8 == 0);
    const bool fast_ok = vec_ok && !p.out_fp32;
    float* stg = reinterpret_cast<float*>(smem + 1024 + (size_t)(warp - 2) * kStageF32);
    // TMA-store epilogue: the bf16 results of a chunk are staged as a 32-row x 64-byte box (64B swizzle: the 16-byte slot of
    // (row, j) is j ^ ((row >> 1) & 3), so the eight rows x four slots a warp writes per pass cover 512 contiguous bytes once) and
    // leave through one cp.async.bulk.tensor store per chunk.  The warp's global stores otherwise queue in the LSU in front of its
    // next bias / colsum / residual loads (profiles/r02_gemm_role_trace.txt: 44 % of the epilogue time sat in the load-issue phase).
    const bool tma_out = p.tma_out != 0;
    uint8_t* ostg = smem + 1024 + stg_bytes + (size_t)(warp - 2) * kStageOut;
    int obuf = 0;
    // ... and with an in-place residual (x += A W^T + b, proj / fc2; + row statistics): p.row_res.  The residual box of chunk n + 1
    // is TMA-loaded into the warp's other staging buffer while chunk n is processed (the box is the output box: same tensor map),
    // the lane adds its row in place and the buffer leaves again as the TMA store — no global loads on the epilogue's critical path,
    // no shuffles for the statistics (a lane owns its row).
    constexpr bool kRowRes = EPI >= 0 && (EPI & kEpiRes) != 0 && (EPI & (kEpiLn | kEpiGelu)) == 0;
    const bool row_res = kRowRes && p.row_res != 0;
    uint32_t rn = 0;                 // chunks this warp has processed (row_res): staging buffer rn & 1, barrier phase (rn >> 1) & 1
    float rs1 = 0.f, rs2 = 0.f;      // row statistics of the lane's row over the warp's chunks of a tile (row_res)
    auto res_issue = [&](int tile, int c0n, uint32_t n) {   // lane 0: residual box of (tile, chunk at column c0n) -> buffer n & 1
      const int mb = tile / p.tiles_n, nb = tile % p.tiles_n;
      uint64_t* bar = &ctrl->res_full[warp - 2][n & 1];
      mbar_expect_tx(bar, (uint32_t)(kStageOut / 2));
      tma_load_2d(ostg + (n & 1) * (kStageOut / 2), &tmC, bar, nb * p.BN + c0n, mb * BM + q * 32);
    };
    if (row_res && lane == 0 && (int)blockIdx.x < num_tiles) res_issue(blockIdx.x, half * 32, 0u);
    const int lr = lane >> 2, lc = (lane & 3) * 8;   // transposed domain: this lane owns rows it*8 + lr, columns lc .. lc+7
    int as = 0;
    uint32_t aphase = 0;
    // LayerNorm statistics of row (q*32 + lane) of the NEXT tile are loaded one tile ahead (coalesced, <= 4 partial
    // pairs per row) so their latency never sits on the epilogue's critical path
    float2 nst[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
    auto load_stats = [&](int tile) {
      const int row = (tile / p.tiles_n) * BM + q * 32 + lane;   // (LayerNorm fold: never an implicit convolution)
      if (row < p.M) {
        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + (long long)row * p.ln_parts;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < p.ln_parts) nst[k] = __ldg(st + k);
      }
    };
    if (f_ln && (int)blockIdx.x < num_tiles) load_stats(blockIdx.x);
    TR_INIT
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      // dense GEMM row of the tile's first row and the number of rows it really has (implicit convolutions: short tiles)
      int row_base = m_blk * BM, row_valid = min(BM, p.M - row_base);
      if (p.conv) {
        if (p.conv_bb > 1) {
          row_base = m_blk * p.tile_rows;
          row_valid = min(p.tile_rows, p.M - row_base);
        } else {
          const int cb = m_blk / p.conv_tpi, cy0 = (m_blk - cb * p.conv_tpi) * p.conv_bh;
          row_base = cb * p.conv_HoWo + cy0 * p.conv_Wo;
          row_valid = min(p.conv_bh, p.conv_Ho - cy0) * p.conv_Wo;
        }
      }
      long long orow_t[4];
      bool rok_t[4];
      float lnr[4], lnn[4];   // LayerNorm fold: v = r * acc + (-r * mu) * colsum[n] + bias[n]
      float own_r = 1.f, own_n = 0.f;
      if (f_ln) {
        const float s1 = (nst[0].x + nst[1].x) + (nst[2].x + nst[3].x), s2 = (nst[0].y + nst[1].y) + (nst[2].y + nst[3].y);
        const float mu = s1 * p.ln_inv_k;
        const float var = fmaxf(fmaf(s2, p.ln_inv_k, -mu * mu), 0.f);
        own_r = rsqrtf(var + p.ln_eps);
        own_n = -own_r * mu;
        if (t + (int)gridDim.x < num_tiles) load_stats(t + gridDim.x);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = row_base + q * 32 + it * 8 + lr;
        rok_t[it] = q * 32 + it * 8 + lr < row_valid;
        orow_t[it] = (p.grp_rows > 0) ? (long long)(rr / p.grp_rows) * p.grp_stride + (rr % p.grp_rows) : (long long)rr;
        lnr[it] = f_ln ? __shfl_sync(0xffffffffu, own_r, it * 8 + lr) : 1.f;
        lnn[it] = f_ln ? __shfl_sync(0xffffffffu, own_n, it * 8 + lr) : 0.f;
      }
      float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
      bool waited = false, loaded = false, released = false;
      uint32_t r[32];   // accumulator chunk in flight: the next chunk's tcgen05.ld is issued as soon as this one sits in smem
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccStride);
      for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
        const int col0 = n_blk * p.BN + c0;
        if (col0 >= p.N) break;  // uniform
        const bool fast = fast_ok && col0 + 32 <= p.N;
        const bool row_path = kRowEpi && p.row_epi != 0 && fast;
        // ---- issue everything that does not depend on the accumulator first (latency hidden behind TMEM/smem) ----
        float bs[8], cs[8];
        uint4 res[4];
        if (fast && !row_path && !row_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { bs[j] = 0.f; cs[j] = 0.f; }
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + lc));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + lc) + 1);
            bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
          }
          if (f_ln) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + lc));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + lc) + 1);
            cs[0] = s0.x; cs[1] = s0.y; cs[2] = s0.z; cs[3] = s0.w; cs[4] = s1.x; cs[5] = s1.y; cs[6] = s1.z; cs[7] = s1.w;
          }
          if (f_res) {
#pragma unroll
            for (int it = 0; it < 4; ++it)
              res[it] = rok_t[it] ? *reinterpret_cast<const uint4*>(p.residual + orow_t[it] * (long long)p.ldc + col0 + lc)
                                  : make_uint4(0u, 0u, 0u, 0u);   // plain load: residual may alias out
          }
        }
        TR(4)   // tile setup + chunk prologue (bias / colsum / residual loads issued)
        if (!waited) {
          mbar_wait(&ctrl->acc_full[as], aphase, 4);
          tc_fence_after();
          waited = true;
        }
        TR(0)
        if (!loaded) tmem_ld_x32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        TR(1)
        if (!fast) {
          const int row = row_base + q * 32 + lane;
          long long orow = row;
          if (p.grp_rows > 0) orow = (long long)(row / p.grp_rows) * p.grp_stride + (row % p.grp_rows);
          epilogue_chunk(p, r, q * 32 + lane < row_valid ? row : p.M, orow, col0, vec_ok, own_r, own_n);
          loaded = false;
          continue;
        }
        if (row_res) {
          // ---- lane = row with in-place residual (every chunk is a fast chunk: checked by the host) ----
          const uint32_t b = rn & 1u;
          {
            // prefetch the next chunk's residual box into the other buffer: its last reader was the TMA store of chunk rn - 1
            const int c1 = c0 + 64;
            const bool more = c1 < p.BN;
            const int tn = t + (int)gridDim.x;
            if (lane == 0 && (more || tn < num_tiles)) {
              tma_store_wait_read<0>();
              if (more) res_issue(t, c1, rn + 1u);
              else res_issue(tn, half * 32, rn + 1u);
            }
          }
          float2 v[16];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (p.bias) {
              b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g8 * 8));
              b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g8 * 8) + 1);
            }
            v[g8 * 4 + 0] = fadd2(make_float2(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), make_float2(b0.x, b0.y));
            v[g8 * 4 + 1] = fadd2(make_float2(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), make_float2(b0.z, b0.w));
            v[g8 * 4 + 2] = fadd2(make_float2(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), make_float2(b1.x, b1.y));
            v[g8 * 4 + 3] = fadd2(make_float2(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), make_float2(b1.z, b1.w));
          }
          {
            const int c1 = c0 + 64;
            const bool more = c1 < p.BN;
            loaded = more;
            if (more) {
              tmem_ld_x32(taddr + (uint32_t)c1, r);   // lands while this chunk is finished
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&ctrl->acc_empty[as]);
              released = true;
            }
          }
          mbar_wait(&ctrl->res_full[warp - 2][b], (rn >> 1) & 1u, 6);      // the residual box of this chunk is in the buffer
          uint8_t* srow = ostg + b * (kStageOut / 2) + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4* slot = reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4));
            const uint4 rv = *slot;
            const float2 o0 = fadd2(v[4 * j + 0], unpack_bf16x2(rv.x)), o1 = fadd2(v[4 * j + 1], unpack_bf16x2(rv.y));
            const float2 o2 = fadd2(v[4 * j + 2], unpack_bf16x2(rv.z)), o3 = fadd2(v[4 * j + 3], unpack_bf16x2(rv.w));
            uint4 wv;
            wv.x = pack_bf16x2(o0.x, o0.y); wv.y = pack_bf16x2(o1.x, o1.y); wv.z = pack_bf16x2(o2.x, o2.y); wv.w = pack_bf16x2(o3.x, o3.y);
            if (f_stats) {   // statistics of the STORED (bf16-rounded) values
              const float2 f0 = unpack_bf16x2(wv.x), f1 = unpack_bf16x2(wv.y), f2 = unpack_bf16x2(wv.z), f3 = unpack_bf16x2(wv.w);
              rs1 += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
              rs2 = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, rs2))));
              rs2 = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, fmaf(f3.x, f3.x, fmaf(f3.y, f3.y, rs2))));
            }
            *slot = wv;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, ostg + b * (kStageOut / 2), col0, row_base + q * 32);   // rows >= M are clipped
            tma_store_commit();
          }
          ++rn;
          TR(3)
          continue;
        }
        if (row_path) {
          // ---- lane = row: math in place on the lane's 32 columns, 8 at a time (constants: warp-uniform 16-byte loads) ----
          uint32_t pk[16];
          const float2 r2 = make_float2(own_r, own_r), n2 = make_float2(own_n, own_n);
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0, s0 = b0, s1 = b0;
            if (p.bias) {
              b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g8 * 8));
              b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g8 * 8) + 1);
            }
            if (f_ln) {
              s0 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + g8 * 8));
              s1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + g8 * 8) + 1);
            }
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
            const float2 cc[4] = {make_float2(s0.x, s0.y), make_float2(s0.z, s0.w), make_float2(s1.x, s1.y), make_float2(s1.z, s1.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 v = make_float2(__uint_as_float(r[g8 * 8 + 2 * j]), __uint_as_float(r[g8 * 8 + 2 * j + 1]));
              v = f_ln ? ffma2(r2, v, ffma2(n2, cc[j], bb[j])) : fadd2(v, bb[j]);
              if (f_gelu) v = gelu_fast2(v);
              pk[g8 * 4 + j] = pack_bf16x2(v.x, v.y);
            }
          }
          {
            const int c1 = c0 + 64;
            const bool more = c1 < p.BN && n_blk * p.BN + c1 < p.N;
            loaded = more;
            if (more) {
              tmem_ld_x32(taddr + (uint32_t)c1, r);   // lands while this chunk is staged and stored
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&ctrl->acc_empty[as]);
              released = true;
            }
          }
          if (lane == 0) tma_store_wait_read<1>();      // the box staged two chunks ago has been read out
          __syncwarp();
          {
            uint8_t* orow_s = ostg + obuf * (kStageOut / 2) + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(orow_s + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, ostg + obuf * (kStageOut / 2), col0, row_base + q * 32);   // rows >= M are clipped
            tma_store_commit();
          }
          obuf ^= 1;
          TR(3)
          continue;
        }
        // ---- transpose the raw fp32 accumulators through the per-warp staging buffer ----
        // lane = row writes 8 x 16 B with the chunk index XOR (row & 7) (conflict-free both ways)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        {
          const int c1 = c0 + 64;
          const bool more = c1 < p.BN && n_blk * p.BN + c1 < p.N;
          loaded = kPipelineLd && more;
          if (loaded) {
            tmem_ld_x32(taddr + (uint32_t)c1, r);   // lands while this chunk goes through math and stores
          } else if (!more) {
            // the accumulator stage is fully read (every lane passed its tcgen05.wait::ld before the __syncwarp): hand it back now
            tc_fence_before();
            if (lane == 0) mbar_arrive(&ctrl->acc_empty[as]);
            released = true;
          }
        }
        TR(2)   // transpose stores
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 8 + lr;
          const int cj = (lane & 3) * 2;
          const float4 a = *reinterpret_cast<const float4*>(stg + rl * 32 + (((cj) ^ (rl & 7)) << 2));
          const float4 b = *reinterpret_cast<const float4*>(stg + rl * 32 + (((cj + 1) ^ (rl & 7)) << 2));
          float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          if (kPackedMath) {
            // compile-time variants: LayerNorm fold, GELU and residual add on the packed fp32 pipe (fma.rn.f32x2 / add.rn.f32x2
            // round like the scalar instructions, so both paths give the same bits)
            const float2 r2 = make_float2(lnr[it], lnr[it]), n2 = make_float2(lnn[it], lnn[it]);
            const uint32_t rw[4] = {f_res ? res[it].x : 0u, f_res ? res[it].y : 0u, f_res ? res[it].z : 0u, f_res ? res[it].w : 0u};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 v = make_float2(o[2 * j], o[2 * j + 1]);
              const float2 c2 = make_float2(cs[2 * j], cs[2 * j + 1]), b2 = make_float2(bs[2 * j], bs[2 * j + 1]);
              v = f_ln ? ffma2(r2, v, ffma2(n2, c2, b2)) : fadd2(v, b2);
              if (f_gelu) v = gelu_fast2(v);
              if (f_res) v = fadd2(v, unpack_bf16x2(rw[j]));
              o[2 * j] = v.x; o[2 * j + 1] = v.y;
            }
          } else {
            if (f_ln) {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = fmaf(lnr[it], o[j], fmaf(lnn[it], cs[j], bs[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] += bs[j];
            }
            if (f_gelu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = gelu_fast(o[j]);
            }
          }
          if (f_res && !kPackedMath) {
            const float2 r0 = unpack_bf16x2(res[it].x), r1 = unpack_bf16x2(res[it].y), r2 = unpack_bf16x2(res[it].z),
                         r3 = unpack_bf16x2(res[it].w);
            o[0] += r0.x; o[1] += r0.y; o[2] += r1.x; o[3] += r1.y; o[4] += r2.x; o[5] += r2.y; o[6] += r3.x; o[7] += r3.y;
          }
          uint4 w;
          w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
          w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
          if (f_stats) {   // statistics of the STORED (bf16-rounded) values
            const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z), f3 = unpack_bf16x2(w.w);
            st1[it] += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
            st2[it] = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, st2[it]))));
            st2[it] = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, fmaf(f3.x, f3.x, fmaf(f3.y, f3.y, st2[it]))));
          }
          if (tma_out) {
            if (it == 0) {   // the box staged two chunks ago must have been read out before its buffer is overwritten
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
            }
            *reinterpret_cast<uint4*>(ostg + obuf * (kStageOut / 2) + rl * 64 + (((lane & 3) ^ ((rl >> 1) & 3)) << 4)) = w;
          } else if (rok_t[it]) {
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + orow_t[it] * (long long)p.ldc + col0 + lc) = w;
          }
        }
        if (tma_out) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, ostg + obuf * (kStageOut / 2), col0, row_base + q * 32);   // rows >= M are clipped
            tma_store_commit();
          }
          obuf ^= 1;
        }
        TR(3)   // transposed reads + math + global stores
      }
      if (!waited) {   // this warp had no chunk in this tile (BN == 32): still consume the phase
        mbar_wait(&ctrl->acc_full[as], aphase, 4);
        tc_fence_after();
      }
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acc_empty[as]);
      }
      if (f_stats && row_res) {
        // the lane owns its row: one plain store per (row, n tile, half), no shuffles
        const int parts = 2 * p.tiles_n, part = 2 * n_blk + half;
        if (q * 32 + lane < row_valid)
          *reinterpret_cast<float2*>(p.stats_out + ((long long)(row_base + q * 32 + lane) * parts + part) * 2) = make_float2(rs1, rs2);
        rs1 = 0.f; rs2 = 0.f;
      } else if (f_stats) {
        // deterministic partials (no atomics): slot (n_blk, half) of every row this warp covers
        const int parts = 2 * p.tiles_n, part = 2 * n_blk + half;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          float a = st1[it], b = st2[it];
          a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
          if ((lane & 3) == 0 && rok_t[it])
            *reinterpret_cast<float2*>(p.stats_out + (orow_t[it] * parts + part) * 2) = make_float2(a, b);
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
      TR(5)   // accumulator release + statistics partials
    }
    if (tma_out && lane == 0) tma_store_wait<0>();   // every tile store has been performed before the CTA (and its staging memory) goes away
    TR_FLUSH
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Bring-up / cross-check kernel: one thread per output element, fp32 accumulate.  Tests only.
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw,
                                 GemmParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.N) return;
  const int row = (int)(idx / p.N), col = (int)(idx % p.N);
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k)
    acc += __bfloat162float(A[(long long)row * lda + k]) * __bfloat162float(W[(long long)col * ldw + k]);
  if (p.ln_stats) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < p.ln_parts; ++k) {
      s1 += p.ln_stats[((long long)row * p.ln_parts + k) * 2];
      s2 += p.ln_stats[((long long)row * p.ln_parts + k) * 2 + 1];
    }
    const float mu = s1 * p.ln_inv_k;
    const float var = fmaxf(s2 * p.ln_inv_k - mu * mu, 0.f);
    const float r = rsqrtf(var + p.ln_eps);
    acc = r * (acc - mu * p.ln_colsum[col]);
  }
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

PerDeviceOnce g_attr_once;

}  // namespace

int gemm_stats_parts(int N, int force_bn) {
  const int bn = force_bn > 0 ? force_bn : pick_bn(N);
  return 2 * ((N + bn - 1) / bn);
}

bool gemm_conv_supported(int H, int W, int C) {
  const int Wo = (W + 1) / 2;
  return H >= 1 && W >= 1 && Wo <= 128 && C >= 8 && C % 8 == 0;
}

int gemm_prepare(const GemmArgs& a, GemmOp* op) {
  const bool conv = a.conv_C > 0;
  LMV_REQUIRE(a.A && a.W && a.out, "gemm: null pointer");
  LMV_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  LMV_REQUIRE(a.K % 8 == 0 && (conv || a.lda % 8 == 0) && a.ldw % 8 == 0, "gemm: K, lda, ldw must be multiples of 8 (16-byte TMA strides)");
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
  p.conv = 0; p.conv_cpt = 0; p.conv_bb = 0; p.conv_bh = 0; p.conv_tpi = 0; p.conv_Ho = 0; p.conv_Wo = 0; p.conv_HoWo = 0;
  p.tile_rows = BM;
  const int stage_bytes = BM * BK * 2 + p.BN * BK * 2;
  p.tx_bytes = (uint32_t)stage_bytes;
  if (conv) {
    const int Ho = (a.conv_H + 1) / 2, Wo = (a.conv_W + 1) / 2, HoWo = Ho * Wo;
    LMV_REQUIRE(gemm_conv_supported(a.conv_H, a.conv_W, a.conv_C), "gemm(conv): needs an output width <= 128 and C % 8 == 0");
    LMV_REQUIRE(a.M == a.conv_B * HoWo && a.K == 9 * a.conv_C && a.ldw == a.K && a.conv_T >= a.conv_H * a.conv_W, "gemm(conv): inconsistent shape");
    LMV_REQUIRE(!a.ln_stats && !a.stats_out && !a.residual, "gemm(conv): bias / activation epilogue only");
    p.conv = 1;
    p.conv_cpt = (a.conv_C + BK - 1) / BK;
    p.k_blocks = 9 * p.conv_cpt;
    p.conv_Ho = Ho; p.conv_Wo = Wo; p.conv_HoWo = HoWo;
    if (HoWo <= BM / 2) {           // several whole images per tile
      p.conv_bb = BM / HoWo; p.conv_bh = Ho; p.conv_tpi = 1;
      p.tile_rows = p.conv_bb * HoWo;
      p.tiles_m = (a.conv_B + p.conv_bb - 1) / p.conv_bb;
    } else {                        // equal-height groups of output rows of one image
      const int max_bh = std::max(1, BM / Wo);
      p.conv_tpi = (Ho + max_bh - 1) / max_bh;
      p.conv_bh = (Ho + p.conv_tpi - 1) / p.conv_tpi;
      p.conv_bb = 1;
      p.tile_rows = p.conv_bh * Wo;
      p.tiles_m = a.conv_B * p.conv_tpi;
    }
    p.tx_bytes = (uint32_t)(p.tile_rows * BK * 2 + p.BN * BK * 2);
  }
  // Shared-memory plan.  Fixed: 2 KB (control + alignment slack) + 8 x 4 KB transpose buffers.  Optional, each only where it does
  // not cost the ring a stage it needs (fewer than min(plain, 4)):
  //   w_res       one N tile and all K blocks of W resident beside a ring of A-only slots (>= 4 of them);
  //   tma_out     8 x 4 KB output staging of the TMA-store epilogue (dense bf16 output, full 128-row tiles).
  // LMV_GEMM_WRES / LMV_GEMM_TMA_OUT = 0 switch one off (A/B runs).  (Measured and dropped: the per-column epilogue constants staged
  // in shared memory one tile ahead behind a named barrier — qkv 70.8 -> 73.9 us.)
  auto env_on = [](const char* name) { const char* e = getenv(name); return !(e && e[0] == '0'); };
  //   row_epi     lane = row epilogue (bias / LayerNorm fold / GELU variants with a dense bf16 output): no transpose buffers at all,
  //               always with tma_out.  LMV_GEMM_ROW_EPI=0 switches it off.
  const bool tma_ok = !conv && !a.out_patched && !a.out_fp32 && a.grp_rows == 0 && a.ldc % 8 == 0 && a.N >= 32 && env_on("LMV_GEMM_TMA_OUT");
  const bool row_epi = tma_ok && !a.residual && !a.stats_out && (a.act == 0 || a.act == 1) && env_on("LMV_GEMM_ROW_EPI");
  // row_res: the same without transposes for `x += A W^T + b` in place (+ row statistics): residual boxes through TMA
  const bool row_res = tma_ok && a.residual && a.residual == a.out && !a.ln_stats && a.act == 0 && a.N % p.BN == 0 && p.BN >= 64 &&
                       env_on("LMV_GEMM_ROW_RES");
  p.row_res = row_res ? 1 : 0;
  p.row_epi = (row_epi || row_res) ? 1 : 0;      // shared-memory layout without transpose buffers
  const int fixed = 2048 + (p.row_epi ? 0 : kEpiWarps * kStageF32);
  const int a_bytes = BM * BK * 2, w_bytes = p.BN * BK * 2;
  const int stages_plain = std::min(kMaxStages, (kSmemLimit - fixed) / stage_bytes);
  const int want = std::min(stages_plain, 4);
  int extra = 0;
  p.w_res = 0; p.stage_bytes = stage_bytes;
  if (p.tiles_n == 1 && env_on("LMV_GEMM_WRES") && (kSmemLimit - fixed - p.k_blocks * w_bytes) / a_bytes >= 4 && p.tiles_m > device_sm_count()) {
    p.w_res = 1;
    p.stage_bytes = a_bytes;
    p.tx_bytes -= (uint32_t)w_bytes;
    extra += p.k_blocks * w_bytes;
  }
  auto stages_with = [&](int more) { return std::min(kMaxStages, (kSmemLimit - fixed - extra - more) / p.stage_bytes); };
  bool tma_out = p.row_epi || (tma_ok && stages_with(kEpiWarps * kStageOut) >= (p.w_res ? 4 : want) && stages_with(kEpiWarps * kStageOut) >= 2);
  p.tma_out = tma_out ? 1 : 0;
  if (tma_out) extra += kEpiWarps * kStageOut;
  p.num_stages = stages_with(0);
  p.bias = a.bias; p.residual = a.residual; p.out = a.out;
  p.ldc = a.ldc; p.out_fp32 = a.out_fp32; p.act = a.act;
  p.grp_rows = a.grp_rows; p.grp_stride = a.grp_stride;
  LMV_REQUIRE((a.ln_stats == nullptr) == (a.ln_colsum == nullptr), "gemm: ln_stats and ln_colsum go together");
  LMV_REQUIRE(a.ln_colsum == nullptr || (reinterpret_cast<uintptr_t>(a.ln_colsum) & 15) == 0, "gemm: ln_colsum must be 16-byte aligned");
  p.ln_stats = a.ln_stats; p.ln_colsum = a.ln_colsum; p.ln_eps = a.ln_eps; p.ln_inv_k = 1.0f / (float)a.K;
  p.ln_parts = a.ln_parts > 0 ? a.ln_parts : 1;
  LMV_REQUIRE(p.ln_parts <= 4, "gemm: at most 4 LayerNorm statistics partials per row");
  p.stats_out = a.stats_out;
  LMV_REQUIRE(a.stats_out == nullptr || (a.N % 32 == 0 && a.ldc % 8 == 0 && !a.out_fp32),
              "gemm: stats_out needs N % 32 == 0, ldc % 8 == 0 and a bf16 output");
  op->smem_bytes = fixed + extra + p.num_stages * p.stage_bytes;
  op->grid = std::min(p.tiles_m * p.tiles_n, device_sm_count());
  if (conv) {
    // activation [B][T rows of C] seen as {C, W, H, B}; the box walks W and H with stride 2 from (kx - 1, 2 oy0 + ky - 1)
    const int C = a.conv_C;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)a.conv_W, (uint64_t)a.conv_H, (uint64_t)a.conv_B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)a.conv_W * C * 2, (uint64_t)a.conv_T * C * 2};
    uint32_t box[4] = {BK, (uint32_t)(2 * p.conv_Wo), (uint32_t)(2 * p.conv_bh), (uint32_t)p.conv_bb};
    uint32_t estr[4] = {1, 2, 2, 1};
    int rc = encode_tmap_bf16(&op->tmA, a.A, 4, dims, strides, box, 128, estr);
    if (rc) return rc;
    uint64_t wdims[3] = {(uint64_t)C, 9, (uint64_t)a.N};
    uint64_t wstrides[2] = {(uint64_t)C * 2, (uint64_t)9 * C * 2};
    uint32_t wbox[3] = {BK, 1, (uint32_t)p.BN};
    rc = encode_tmap_bf16(&op->tmB, a.W, 3, wdims, wstrides, wbox, 128);
    if (rc) return rc;
  } else {
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
  }
  op->tmC = op->tmA;
  if (p.tma_out) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t strides[1] = {(uint64_t)a.ldc * 2};
    uint32_t box[2] = {32, 32};
    int rc = encode_tmap_bf16(&op->tmC, a.out, 2, dims, strides, box, 64);
    if (rc) return rc;
  }
  op->tmR = op->tmA;
  p.prefetch_res = 0;
  if (!conv && a.residual && a.grp_rows == 0 && a.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t strides[1] = {(uint64_t)a.ldc * 2};
    uint32_t box[2] = {(uint32_t)std::min(p.BN, ((a.N + 7) / 8) * 8), BM};
    if (encode_tmap_bf16(&op->tmR, a.residual, 2, dims, strides, box, 0) == LMV_OK) p.prefetch_res = 1;
    else op->tmR = op->tmA;
  }
  return LMV_OK;
}

using GemmKernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
// the epilogue combinations the LeMeViT schedule uses get their own instantiation; everything else runs the run-time one
static const struct { int flags; GemmKernel fn; } kGemmVariants[] = {
    {0, gemm_bf16_tn_tcgen05<0>},                                   // bias only (convolutions as GEMM, c-path q/kv)
    {kEpiLn, gemm_bf16_tn_tcgen05<kEpiLn>},                         // LN-folded qkv
    {kEpiLn | kEpiGelu, gemm_bf16_tn_tcgen05<kEpiLn | kEpiGelu>},   // LN-folded fc1 + GELU
    {kEpiGelu, gemm_bf16_tn_tcgen05<kEpiGelu>},                     // fc1 + GELU on already-normalised input
    {kEpiRes, gemm_bf16_tn_tcgen05<kEpiRes>},                       // fc2 + residual
    {kEpiRes | kEpiStats, gemm_bf16_tn_tcgen05<kEpiRes | kEpiStats>},   // proj + residual + statistics for the next LN
    {-1, gemm_bf16_tn_tcgen05<-1>},
};

int gemm_run(const GemmOp& op, cudaStream_t stream) {
  LMV_CUDA_OK(g_attr_once.run([] {
    cudaError_t err = cudaSuccess;
    for (const auto& v : kGemmVariants) {
      const cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (e != cudaSuccess) err = e;
    }
    return err;
  }));
  const GemmParams& gp = op.p;
  const int flags = (gp.ln_stats ? kEpiLn : 0) | (gp.act == 1 ? kEpiGelu : 0) | (gp.residual ? kEpiRes : 0) | (gp.stats_out ? kEpiStats : 0);
  GemmKernel fn = gemm_bf16_tn_tcgen05<-1>;
  if (gp.act == 0 || gp.act == 1)
    for (const auto& v : kGemmVariants)
      if (v.flags == flags) { fn = v.fn; break; }
  LMV_CUDA_OK(launch_kernel(fn, dim3(op.grid), dim3(kThreads), (size_t)(op.smem_bytes), stream, op.tmA, op.tmB, op.tmR, op.tmC, op.p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int gemm_simt_run(const GemmArgs& a, cudaStream_t stream) {
  GemmParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.ldc = a.ldc; p.out_fp32 = a.out_fp32; p.act = a.act;
  p.grp_rows = a.grp_rows; p.grp_stride = a.grp_stride;
  p.ln_stats = a.ln_stats; p.ln_colsum = a.ln_colsum; p.ln_eps = a.ln_eps; p.ln_inv_k = 1.0f / (float)a.K;
  p.ln_parts = a.ln_parts > 0 ? a.ln_parts : 1;
  p.stats_out = nullptr;   // the SIMT path gets its row statistics from row_stats_run (kernels.h)
  const long long total = (long long)a.M * a.N;
  LMV_CUDA_OK(launch_kernel(gemm_simt_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), stream, a.A, a.lda, a.W, a.ldw, p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv

#ifdef LMV_GEMM_TRACE
// copies out and clears the [148 CTAs][10 warps][8 counters] cycle table of the traced GEMM launches (debug builds only)
extern "C" int lmv_debug_gemm_trace(unsigned long long* host, int n) {
  static unsigned long long zero[148 * 10 * 8];
  if (n > 148 * 10 * 8) n = 148 * 10 * 8;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(host, lmv::g_gemm_trace, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(lmv::g_gemm_trace, zero, sizeof(zero)) != cudaSuccess;
}
#endif
