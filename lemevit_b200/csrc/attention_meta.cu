// Meta-token side of the LeMeViT cross attentions on tensor cores: a handful of queries (the M = 16 meta tokens)
// attend over ALL N image tokens, softmax over the image tokens:
//   CrossAttention        (models/lemevit.py:477-486, stage 0):  c <- softmax(q(c) k(x)^T / sqrt(32)) v(x)
//   DualCrossAttention    (models/lemevit.py:300-302, stages 1-2): dc = softmax(q2 k1^T * C^-0.5) v1
// FlashAttention-style kernels waste >75 % of their tiles on 16 query rows; here the roles are arranged so that
// the tcgen05 M dimension is filled by (head, query) pairs and the N/K dimensions by image tokens.  The op is
// HBM-bound (K and V are read once, 2 * N * C bf16 per image), so the kernel is a persistent stream over token tiles.
//
// sm_100a design: one persistent CTA per SM walks a contiguous range of (image, 128-token tile) items:
//   * warp 4 (one thread) is the control thread: TMA loads of the K / V tiles [128 tokens x C] straight out of the packed
//     kv / qkv activation (64B swizzle, 32-channel boxes, rows past the end of the image zero-filled) into a 1-2 deep
//     ring, and both MMAs of every tile;
//   * Q is expanded in shared memory to a block-diagonal operand Qbd[(h, j), C] (row (h, j) holds q_j restricted to the
//     32 channels of head h), so ONE accumulation over the full channel dim gives every head's scores:
//         S[(h, j), n] = sum_c Qbd[(h, j), c] K[n, c]                tcgen05.mma M=128, N=128, K=C   (A, B K-major, SW64)
//     S is double-buffered in TMEM so the scores of tile i+1 are computed while tile i is in its softmax;
//   * warps 0-3 do the softmax, thread r owning TMEM lane r = row (h, j).  With heads * Lq <= 64 the rows are duplicated
//     at lanes 64.. and each copy takes half of the tile's tokens, so all four warps work;  P (bf16) goes to a 128B
//     swizzled K-major smem tile;
//   * O[(h, j), c] = sum_n P[(h, j), n] V[n, c]                      tcgen05.mma M=128, N=32 per head chunk, K=128 tokens
//     (B = V as loaded: MN-major); each thread folds its row of O into a running (m, l, O[32]) in registers;
//   * an image's tiles are cut into fixed segments (a function of N only); at the end of a segment the running state is
//     written as one split-softmax partial and a small second kernel merges the segments x copies partials of an image
//     in a fixed order, so every output bit is independent of batch size, batch position and SM count.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kD = 32;
constexpr int kTile = 128;        // image tokens per item
constexpr int kThreads = 160;     // 4 softmax warps + 1 control warp
constexpr int kChunkBytes = kTile * kD * 2;   // [128 rows x 32 ch] bf16 = 8 KB
constexpr int kTmemS = 0, kTmemO = 2 * kTile; // S[2]: columns [0, 256), O: [256, 256 + C)

struct MetaParams {
  const bf16* q;
  long long q_bs;
  int q_rs;
  float* part_o;     // [B][parts][R][32]
  float2* part_ml;   // [B][parts][R]
  int heads, Lq, Lk, R, C, tiles, nchunk;
  int dup;           // rows duplicated at lanes 64..: copy 0 takes tokens [0, 64) of a tile, copy 1 tokens [64, 128)
  int seg_tiles, segs;         // an image's tiles are cut into `segs` segments of seg_tiles tiles: one partial per segment,
                               // so the summation order (and every output bit) is independent of batch size and position
  int total_tiles;             // B * tiles: CTAs take contiguous, tile-balanced runs of whole segments
  int kst, vst;      // K / V ring depth
  float scale_log2e;
};

struct Ctrl {
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full[2], s_empty[2], p_full, o_full, o_empty, q_ready;
  uint32_t tmem_base;
};

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

__global__ void __launch_bounds__(kThreads, 1)
attention_meta_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const MetaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const int kv_bytes = p.nchunk * kChunkBytes;
  uint8_t* sQ = smem + 1024;
  uint8_t* sK = sQ + kv_bytes;
  uint8_t* sV = sK + (size_t)p.kst * kv_bytes;
  uint8_t* sP = sV + (size_t)p.vst * kv_bytes;               // 2 tiles of [128 x 64] bf16, 128B swizzle
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto seg_first_tile = [&](int sidx) {   // flattened tile index at which flattened segment sidx starts
    const int b = sidx / p.segs, sg = sidx - b * p.segs;
    return b * p.tiles + min(sg * p.seg_tiles, p.tiles);
  };
  // CTA c starts at the first segment boundary at or after tile c * total / grid: whole segments, tile-balanced
  auto cta_first_seg = [&](int c) {
    const long long tau = (long long)c * p.total_tiles / (long long)gridDim.x;
    const int b = (int)(tau / p.tiles), within = (int)(tau - (long long)b * p.tiles);
    return b * p.segs + (within + p.seg_tiles - 1) / p.seg_tiles;
  };
  const int s_begin = cta_first_seg(blockIdx.x), s_end = cta_first_seg(blockIdx.x + 1);
  const int t_begin = seg_first_tile(s_begin), t_end = seg_first_tile(s_end);
  const int n = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->k_full[i], 1); mbar_init(&ctrl->k_empty[i], 1);
      mbar_init(&ctrl->v_full[i], 1); mbar_init(&ctrl->v_empty[i], 1);
      mbar_init(&ctrl->s_full[i], 1); mbar_init(&ctrl->s_empty[i], 4);
    }
    mbar_init(&ctrl->p_full, 4);
    mbar_init(&ctrl->o_full, 1);
    mbar_init(&ctrl->o_empty, 4);
    mbar_init(&ctrl->q_ready, 4);
    fence_mbar_init();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 4) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctrl->tmem_base;
  if (n <= 0) {   // (only when the grid was rounded up)
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
    return;
  }

  if (warp == 4) {
    // ---------------- control thread: TMA loads + both MMAs of every tile ----------------
    if (lane == 0) {
      // (image, tile) cursors advance by one tile at a time: no divisions in the loops; ring depths are 1 or 2
      struct Cur { int b, tile; };
      const int b0 = t_begin / p.tiles;
      const Cur c0 = {b0, t_begin - b0 * p.tiles};
      auto step = [&](Cur& c) { if (++c.tile == p.tiles) { c.tile = 0; ++c.b; } };
      const int ksh = p.kst - 1, vsh = p.vst - 1;   // slot = i & sh, use count = i >> sh
      Cur ck = c0, cv = c0, cs = c0, ci = c0;       // next K load, next V load, next S issue, tile i
      auto load_kv = [&](const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, Cur& c) {
        mbar_expect_tx(bar, (uint32_t)kv_bytes);
#pragma unroll 1
        for (int ch = 0; ch < p.nchunk; ++ch) tma_load_3d(dst + (size_t)ch * kChunkBytes, tm, bar, ch * kD, c.tile * kTile, c.b);
        step(c);
      };
      auto load_k = [&](int i) { load_kv(&tmK, sK + (size_t)(i & ksh) * kv_bytes, &ctrl->k_full[i & ksh], ck); };
      auto load_v = [&](int i) { load_kv(&tmV, sV + (size_t)(i & vsh) * kv_bytes, &ctrl->v_full[i & vsh], cv); };
      const uint32_t idesc_s = make_idesc_bf16(128, kTile);
      const uint32_t idesc_o = make_idesc_bf16(128, kD) | (1u << 16);   // b_major = MN
      int k_next = 0;                       // next K tile to request
      for (; k_next < min(p.kst, n); ++k_next) load_k(k_next);
      for (int i = 0; i < min(p.vst, n); ++i) load_v(i);
      mbar_wait_lean(&ctrl->q_ready, 0);   // Qbd of the first image written, P tile zeroed
      int s_next = 0;                       // next tile whose scores have not been issued
      for (int i = 0; i < n; ++i) {
        // refill the K ring slots whose scores were issued in an earlier trip (their MMAs retired long ago)
#pragma unroll 1
        while (k_next < n && k_next - p.kst < s_next) {
          const int js = k_next - p.kst;
          mbar_wait_lean(&ctrl->k_empty[js & ksh], (uint32_t)((js >> ksh) & 1));
          load_k(k_next++);
        }
        // scores: S(i) if still missing, and S(i+1) while tile i is in its softmax — unless it belongs to a new image,
        // whose Qbd is only staged before p_full(i) (single issue site: the loop body exists once in the instruction stream)
#pragma unroll 1
        while (s_next < n && (s_next == i || (s_next == i + 1 && cs.b == ci.b))) {
          const int js = s_next++, kb = js & ksh;
          step(cs);
          while (k_next <= js) {   // single-slot ring only: K(js) can be requested once S(js - 1), issued just now, has retired
            const int jp = k_next - p.kst;
            if (jp >= 0) mbar_wait_lean(&ctrl->k_empty[jp & ksh], (uint32_t)((jp >> ksh) & 1));
            load_k(k_next++);
          }
          mbar_wait_lean(&ctrl->k_full[kb], (uint32_t)((js >> ksh) & 1));
          if (js >= 2) mbar_wait_lean(&ctrl->s_empty[js & 1], (uint32_t)(((js >> 1) - 1) & 1));
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(kTmemS + (js & 1) * kTile);
#pragma unroll 1
          for (int c = 0; c < p.nchunk; ++c) {
            const uint64_t dq = make_kmajor_desc<64>(smem_u32(sQ + (size_t)c * kChunkBytes));
            const uint64_t dk = make_kmajor_desc<64>(smem_u32(sK + (size_t)kb * kv_bytes + (size_t)c * kChunkBytes));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(d, dq + 2ull * k, dk + 2ull * k, idesc_s, (uint32_t)((c | k) != 0));
          }
          umma_commit(&ctrl->s_full[js & 1]);
          umma_commit(&ctrl->k_empty[kb]);
        }
        mbar_wait_lean(&ctrl->p_full, (uint32_t)(i & 1));
        mbar_wait_lean(&ctrl->v_full[i & vsh], (uint32_t)((i >> vsh) & 1));
        if (i > 0) mbar_wait_lean(&ctrl->o_empty, (uint32_t)((i - 1) & 1));
        tc_fence_after();
        // O[:, 32c .. 32c+31] = P V_c : A = P tiles (K-major, 128B swizzle), B = V chunk as loaded (MN-major, 64B swizzle)
        const uint32_t pbase = smem_u32(sP);
        for (int c = 0; c < p.nchunk; ++c) {
          const uint32_t vbase = smem_u32(sV + (size_t)(i & vsh) * kv_bytes + (size_t)c * kChunkBytes);
#pragma unroll
          for (int s = 0; s < kTile / 16; ++s) {
            const uint64_t da = make_kmajor_desc<128>(pbase + (uint32_t)(s >> 2) * (kTile * 128)) + 2ull * (s & 3);
            const uint64_t db = make_mnmajor_sw64_desc(vbase + (uint32_t)s * (16 * kD * 2));
            umma_bf16_ss(tmem + (uint32_t)(kTmemO + c * kD), da, db, idesc_o, (uint32_t)(s != 0));
          }
        }
        umma_commit(&ctrl->o_full);
        umma_commit(&ctrl->v_empty[i & vsh]);
        if (i + p.vst < n) {
          mbar_wait_lean(&ctrl->v_empty[i & vsh], (uint32_t)((i >> vsh) & 1));
          load_v(i + p.vst);
        }
        step(ci);
      }
    }
  } else {
    // ---------------- softmax warps: thread r owns TMEM lane r ----------------
    const int r = threadIdx.x;
    const int rr = p.dup ? (r & 63) : r;          // logical row (h, j)
    const int copy = p.dup ? (r >> 6) : 0;
    const int ncopy = p.dup ? 2 : 1;
    const bool rvalid = rr < p.R;
    const int h = rvalid ? rr / p.Lq : 0, j = rr - h * p.Lq;
    const int c_lo = p.dup ? copy * 64 : 0, c_n = p.dup ? 64 : 128;   // token columns of a tile this thread owns
    // zero Qbd and the P tile once: the block-diagonal / copy structure never changes, only the written parts do
    {
      const int n16 = (kv_bytes) / 16;
      for (int i = threadIdx.x; i < n16; i += 128) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0u, 0u, 0u, 0u);
      for (int i = threadIdx.x; i < 2 * kTile * 128 / 16; i += 128) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    auto write_q = [&](int b) {
      if (rvalid) {
        const uint4* src = reinterpret_cast<const uint4*>(p.q + (long long)b * p.q_bs + (long long)j * p.q_rs + h * kD);
        uint8_t* dst = sQ + (size_t)h * kChunkBytes + (size_t)r * 64;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + ((ch ^ ((r >> 1) & 3)) << 4)) = __ldg(src + ch);
      }
      fence_proxy_async_smem();
    };
    // cursor of tile i: image, tile in the image, segment id, tiles left in the segment (advanced once per trip, no divisions)
    int cb = t_begin / p.tiles, ctile = t_begin - cb * p.tiles;
    int seg_cur = cb * p.segs + ctile / p.seg_tiles;        // segment the running state belongs to
    int seg_i = seg_cur, seg_pos = ctile % p.seg_tiles;     // segment of tile i and its position inside it
    write_q(cb);
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctrl->q_ready);

    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
    const int h_lo = ((warp * 32) & (p.dup ? 63 : 127)) / p.Lq;
    const int h_hi = min(p.heads - 1, (((warp * 32) & (p.dup ? 63 : 127)) + 31) / p.Lq);
    float m_run = -INFINITY, l_run = 0.f, o_run[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o_run[i] = 0.f;
    float m_prev = -INFINITY, l_prev = 0.f;   // (max, sum) of the tile whose O is still in flight

    // fold O of tile `it` (already complete in TMEM) into the running state
    auto fold = [&](int it) {
      mbar_wait_lean(&ctrl->o_full, (uint32_t)(it & 1));
      tc_fence_after();
      const float m_new = fmaxf(m_run, m_prev);
      const float a = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      const float w = (m_prev == -INFINITY) ? 0.f : exp2f(m_prev - m_new);
      for (int hh = h_lo; hh <= h_hi; ++hh) {   // h differs per lane: load lane-uniform chunks and select
        uint32_t v[32];
        tmem_ld_x32(t_row + (uint32_t)(kTmemO + hh * kD), v);
        tmem_ld_wait();
        if (hh == h) {
#pragma unroll
          for (int i = 0; i < kD; ++i) o_run[i] = fmaf(o_run[i], a, w * __uint_as_float(v[i]));
        }
      }
      l_run = fmaf(l_run, a, w * l_prev);
      m_run = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->o_empty);
    };
    auto flush = [&](int sidx) {
      if (rvalid) {
        const long long pr = ((long long)sidx * ncopy + copy) * p.R + rr;
        p.part_ml[pr] = make_float2(m_run, l_run);
        float4* dst = reinterpret_cast<float4*>(p.part_o + pr * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(o_run[4 * i], o_run[4 * i + 1], o_run[4 * i + 2], o_run[4 * i + 3]);
      }
      m_run = -INFINITY; l_run = 0.f;
#pragma unroll
      for (int i = 0; i < kD; ++i) o_run[i] = 0.f;
    };

    // One extra trip (i == n) folds and flushes the last tile, so fold / flush / the 32-column block bodies exist once in
    // the instruction stream: the five warps of this kernel cannot hide instruction-cache misses.
    for (int i = 0; i <= n; ++i) {
      const bool last = (i == n);
      const int valid = min(kTile, p.Lk - ctile * kTile);
      const uint32_t s_row = t_row + (uint32_t)(kTmemS + (i & 1) * kTile + c_lo);
      float mx = -INFINITY;
      // one 32-column block of scores; columns past the end of the image become -inf (-> probability 0)
      auto load_block = [&](uint32_t (&v)[32], int c0) {
        tmem_ld_x32(s_row + (uint32_t)c0, v);
        tmem_ld_wait();
        const int nv = valid - (c_lo + c0);
        if (nv < 32) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k >= nv) v[k] = 0xff800000u;
        }
      };
      if (!last) {
        mbar_wait_lean(&ctrl->s_full[i & 1], (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        // S(i) is complete, so every MMA that reads the current Qbd has retired: stage the next image's queries now
        if (i + 1 < n && ctile + 1 == p.tiles) write_q(cb + 1);
        // ---- pass 1: tile maximum of this thread's columns ----
#pragma unroll 1
        for (int c0 = 0; c0 < c_n; c0 += 32) {
          uint32_t v[32];
          load_block(v, c0);
#pragma unroll
          for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(v[k]));
        }
      }
      // ---- previous tile: its O is needed before P (single buffer) can be overwritten ----
      if (i > 0) {
        fold(i - 1);
        if (last || seg_i != seg_cur) { flush(seg_cur); seg_cur = seg_i; }
      }
      if (last) break;
      // ---- pass 2: probabilities -> P tile (bf16), row sum of exactly the values the tensor core multiplies ----
      const float mxs = mx * p.scale_log2e;
      const float mxs_safe = (mx == -INFINITY) ? 0.f : mxs;   // no real token in this thread's columns: every p is exp2(-inf) = 0
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < c_n; c0 += 32) {
        uint32_t v[32];
        load_block(v, c0);
        const int col0 = c_lo + c0;
        uint8_t* tile_p = sP + (size_t)(col0 >> 6) * (kTile * 128) + (size_t)r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) e[k] = ex2_approx(fmaf(__uint_as_float(v[g * 8 + k]), p.scale_log2e, -mxs_safe));
          uint4 u;
          u.x = pack_bf16x2(e[0], e[1]); u.y = pack_bf16x2(e[2], e[3]);
          u.z = pack_bf16x2(e[4], e[5]); u.w = pack_bf16x2(e[6], e[7]);
          const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
          sum += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
          const int ch = ((col0 & 63) >> 3) + g;
          if (rvalid) *reinterpret_cast<uint4*>(tile_p + ((ch ^ (r & 7)) << 4)) = u;
        }
      }
      m_prev = mxs; l_prev = sum;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctrl->s_empty[i & 1]);
        mbar_arrive(&ctrl->p_full);
      }
      // advance the cursor to tile i + 1
      if (++ctile == p.tiles) { ctile = 0; ++cb; seg_pos = 0; ++seg_i; }
      else if (++seg_pos == p.seg_tiles) { seg_pos = 0; ++seg_i; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// merge the partials of one (image, head, query) row: one warp per row, lane = channel
__global__ void __launch_bounds__(256)
attention_meta_merge_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, bf16* __restrict__ out,
                            long long o_bs, int o_rs, int B, int parts, int R, int Lq) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= (long long)B * R) return;
  const int b = (int)(idx / R), r = (int)(idx % R);
  const int h = r / Lq, j = r - h * Lq;
  const long long base = (long long)b * parts;   // segments x copies partial rows per image, merged in a fixed order
  float m = -INFINITY;
  for (int t = 0; t < parts; ++t) m = fmaxf(m, part_ml[(base + t) * R + r].x);
  float l = 0.f, o = 0.f;
  for (int t = 0; t < parts; ++t) {
    const long long pr = (base + t) * R + r;
    const float2 ml = part_ml[pr];
    const float w = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - m);
    l = fmaf(ml.y, w, l);
    o = fmaf(part_o[pr * kD + lane], w, o);
  }
  out[(long long)b * o_bs + (long long)j * o_rs + h * kD + lane] = __float2bfloat16(o / l);
}

PerDeviceOnce g_attr_once;

struct MetaShape {
  int nchunk, kst, vst, dup, tiles, seg_tiles, segs, grid, ncopy, smem;
};

MetaShape shape_for(const AttnArgs& a) {
  MetaShape s;
  s.nchunk = a.heads;
  s.dup = (a.heads * a.Lq <= 64) ? 1 : 0;
  s.ncopy = s.dup ? 2 : 1;
  // Q + K ring + V ring + P (32 KB) + control/alignment (2 KB) within 227 KB
  const int kv = s.nchunk * kChunkBytes, budget = 227 * 1024 - 2048 - 2 * kTile * 128 - kv;
  s.kst = (budget >= 3 * kv) ? 2 : 1;
  s.vst = (budget >= 4 * kv) ? 2 : 1;
  s.smem = 2048 + kv * (1 + s.kst + s.vst) + 2 * kTile * 128;
  s.tiles = (a.Lk + kTile - 1) / kTile;
  static const int seg_env = [] { const char* e = getenv("LMV_META_SEG"); return e ? atoi(e) : 0; }();
  s.seg_tiles = seg_env > 0 ? seg_env : (s.tiles >= 12 ? 6 : (s.tiles >= 6 ? 3 : 1));   // a function of Lk only (never of B): see MetaParams::seg_tiles
  s.segs = (s.tiles + s.seg_tiles - 1) / s.seg_tiles;
  s.grid = std::min(a.B * s.segs, device_sm_count());
  return s;
}

}  // namespace

bool attention_meta_supported(const AttnArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int C = a.heads * kD;
  return a.Lq >= 1 && a.heads * a.Lq <= 128 && C <= 256 && a.Lk >= 1 && a.B >= 1 && al16(a.q) && al16(a.k) && al16(a.v) &&
         a.q_rs % 8 == 0 && a.q_bs % 8 == 0 && a.k_rs % 8 == 0 && a.v_rs % 8 == 0 && a.k_bs % 8 == 0 && a.v_bs % 8 == 0 &&
         a.k_rs >= C && a.v_rs >= C && (long long)a.B * ((a.Lk + kTile - 1) / kTile) < (1ll << 28) &&
         shape_for(a).smem <= 227 * 1024;
}

size_t attention_meta_workspace(const AttnArgs& a) {
  const MetaShape s = shape_for(a);
  return (size_t)a.B * s.segs * s.ncopy * ((size_t)a.heads * a.Lq) * (kD * sizeof(float) + sizeof(float2));
}

int attention_meta_run(const AttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!attention_meta_supported(a)) return fail(LMV_ERR_UNSUPPORTED, "attention_meta: unsupported shape / alignment");
  LMV_REQUIRE(workspace && workspace_bytes >= attention_meta_workspace(a) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
              "attention_meta: partial-softmax workspace missing or too small");
  LMV_CUDA_OK(g_attr_once.run([] { return cudaFuncSetAttribute(attention_meta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  const MetaShape sh = shape_for(a);
  MetaParams p;
  p.q = a.q; p.q_bs = a.q_bs; p.q_rs = a.q_rs;
  p.heads = a.heads; p.Lq = a.Lq; p.Lk = a.Lk; p.R = a.heads * a.Lq; p.C = a.heads * kD;
  p.tiles = sh.tiles; p.nchunk = sh.nchunk; p.dup = sh.dup;
  p.seg_tiles = sh.seg_tiles; p.segs = sh.segs; p.total_tiles = a.B * sh.tiles; p.kst = sh.kst; p.vst = sh.vst;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  const size_t rows = (size_t)a.B * sh.segs * sh.ncopy * p.R;
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
  LMV_CUDA_OK(launch_kernel(attention_meta_kernel, dim3(sh.grid), dim3(kThreads), (size_t)(sh.smem), s, tk, tv, p));
  LMV_CUDA_OK(cudaGetLastError());
  const long long mrows = (long long)a.B * p.R;
  LMV_CUDA_OK(launch_kernel(attention_meta_merge_kernel, dim3((unsigned)((mrows + 7) / 8)), dim3(256), (size_t)(0), s, p.part_o, p.part_ml, a.out, a.o_bs, a.o_rs, a.B, sh.segs * sh.ncopy, p.R, a.Lq));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
