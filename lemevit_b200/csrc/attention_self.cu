// Self-attention of the late LeMeViT stages on tensor cores, persistent and warp-specialised:
//   StandardAttention (models/lemevit.py:199-205; F.scaled_dot_product_attention :203) on the N image tokens of every image
//   AND, in the same launch, on the M = 16 meta tokens that share the block's weights (forward_with_x :632-635): the
//   [B, N+M, 3C] qkv buffer holds both, rows [0, N) attend keys [0, N), rows [N, N+M) attend keys [N, N+M).
//
// sm_100a design — one CTA per SM, looping over work items (image, head, pair of 128-row query tiles):
//   warp 0        TMA producer: K, V [Lkp x 32] and the two Q tiles [128 x 32] of the item (64B swizzle) into a 2-stage ring —
//                 K and V are fetched ONCE for both query tiles
//   warps 1, 3    one issuer per softmax group:  S_g = Q_g K^T  (M=128, N=Lkp<=224, K=32)   -> TMEM columns [224 g, 224 g + Lkp)
//                                                 O_g = P_g V    (M=128, N=32,  K=Lkp)       -> TMEM columns [448 + 32 g, +32)
//                 S_g of item i+1 is issued BEFORE P_g V of item i, so a softmax group always finds its next scores ready
//   warps 4..11   softmax group 0 (query tile 0): two warps per TMEM lane quarter, each owning half of the key columns of its 32
//   warps 12..19  softmax group 1 (query tile 1)   rows: partial row max -> exchange through smem -> exp2 with the scale folded
//                 in, P as bf16 into 128B-swizzled K-major smem tiles (A operand of P V), partial row sums exchanged; the output
//                 epilogue of item i (O / rowsum -> bf16 -> global) runs after the softmax of item i+1 was started, so nobody
//                 ever waits for the P V MMA
// The two groups drift apart naturally: while one is in its MUFU-bound exp phase the other loads TMEM or stores results.
#include <mutex>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kD = 32;
constexpr bool kPolyExp = true;   // every fourth pair of exponentials of the exp pass on the FMA pipe (exp2_poly2)
constexpr int kQTile = 128;
constexpr int kMaxKeys = 224;
constexpr int kSCols = 224;              // TMEM columns reserved per S accumulator
constexpr int kOCol = 2 * kSCols;        // 448
constexpr int kFirstSoftmaxWarp = 4;
constexpr int kGroupWarps = 8;               // two warps per TMEM lane quarter
constexpr int kThreads = 32 * (kFirstSoftmaxWarp + 2 * kGroupWarps);
constexpr int kQBytes = kQTile * kD * 2;           // 8 KB
constexpr int kKVBytes = kMaxKeys * kD * 2;        // 14 KB
constexpr int kStageBytes = 2 * kQBytes + 2 * kKVBytes;   // 44 KB
constexpr int kPTileBytes = kQTile * 128;          // [128 x 64] bf16
constexpr int kPBytes = 4 * kPTileBytes;           // 64 KB per group (Lkp <= 256)
constexpr int kXchgBytes = 2 * 2 * kQTile * 4 * 3;   // per group and column half: row max + two parities of row sums
constexpr int kSmemBytes = 2048 + ((kXchgBytes + 1023) / 1024) * 1024 + 2 * kStageBytes + 2 * kPBytes;

struct SelfParams {
  bf16* out;
  long long o_bs;
  int o_rs;
  int B, heads, T, N, Lkp, pairs, qtiles;
  int nkv, KB;               // key blocks per (image, head) and keys per block (nkv > 1: split-KV, partials merged by a second kernel)
  float* part_o;             // [B*heads][qtiles*128][nkv][32] unnormalised outputs of every key block (nkv > 1)
  float2* part_ml;           // [B*heads][qtiles*128][nkv] (row max * scale * log2e, row sum)
  long long items;
  float scale_log2e;
};

struct Ctrl {
  uint64_t kq_full[2], kq_empty[2];   // Q tiles + K of a stage: free again as soon as the item's two S MMAs are done
  uint64_t v_full[2], v_empty[2];     // V of a stage: read by the item's P V MMAs, one softmax later
  uint64_t s_full[2], p_full[2], p_empty[2], o_full[2], o_empty[2];

  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// 2^x for x <= 0 on the FMA / ALU pipes, two values at a time (the exp pass is MUFU-bound: every fourth pair goes here instead of
// through ex2.approx).  Cody-Waite: n = round(x) via the 1.5 * 2^23 trick, 2^(x - n) by a cubic on [-0.5, 0.5] (max relative error
// 2.2e-4, an eighth of the bf16 rounding P gets anyway), 2^n added straight into the exponent field.  Arguments below -126
// (masked keys are -inf) give ~1e-38 instead of 0: 38 orders of magnitude below the row maximum's 1.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  const float kMagic = 12582912.f;
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = fadd2(x, make_float2(kMagic, kMagic));
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
  const float2 r = ffma2(n, make_float2(-1.f, -1.f), x);
  float2 q = ffma2(r, make_float2(0.05286743491888046f, 0.05286743491888046f), make_float2(0.2421518862247467f, 0.2421518862247467f));
  q = ffma2(q, r, make_float2(0.6935867667198181f, 0.6935867667198181f));
  q = ffma2(q, r, make_float2(0.9999627470970154f, 0.9999627470970154f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)), __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
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

// Debug build (-DLMV_ATTN_TRACE): cycle account of one warp per (group, column half), read back with lmv_debug_attn_trace().
// slots: 0 item setup, 1 wait s_full, 2 pass 1, 3 group sync, 4 wait p_empty, 5 pass 2, 6 output epilogue (incl. wait o_full), 7 total
#ifdef LMV_ATTN_TRACE
__device__ unsigned long long g_attn_trace[148 * 5 * 8];
#define TR_INIT unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64(); const long long tr_t0 = tr_t;
#define TR(i) { const long long now_ = clock64(); tr[i] += (unsigned long long)(now_ - tr_t); tr_t = now_; }
#define TR_FLUSH(role) { tr[7] = (unsigned long long)(clock64() - tr_t0); if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_attn_trace[((size_t)blockIdx.x * 5 + (role)) * 8 + i_] += tr[i_]; }
#else
#define TR_INIT
#define TR(i)
#define TR_FLUSH(role)
#endif
// -DLMV_ATTN_TRACE=3: absolute timestamps of CTA 0's first 16 items: [role 0..5][item][slot 0..7] (roles: softmax g0 h0, g0 h1, g1 h0, g1 h1
// (lane quarter 0), issuer g0, issuer g1)
#if defined(LMV_ATTN_TRACE) && LMV_ATTN_TRACE == 3
__device__ long long g_attn_events[6 * 16 * 8];
#define EV(role, item, slot) { if (blockIdx.x == 0 && lane == 0 && (item) < 16) g_attn_events[((role) * 16 + (int)(item)) * 8 + (slot)] = clock64(); }
#else
#define EV(role, item, slot)
#endif

__global__ void __launch_bounds__(kThreads, 1)
attention_self_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const SelfParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  float* sXchg = reinterpret_cast<float*>(smem + 1024);   // [group][half][max | sum parity 0 | sum parity 1][128 rows]
  uint8_t* sStage = smem + 1024 + ((kXchgBytes + 1023) / 1024) * 1024;   // 2 x {Q0, Q1, K, V}
  uint8_t* sP = sStage + 2 * kStageBytes;              // 2 groups x 4 tiles
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const long long n_my = ((long long)blockIdx.x < p.items) ? (p.items - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->kq_full[i], 1);
      mbar_init(&ctrl->kq_empty[i], p.qtiles > 1 ? 2 : 1);   // one commit per issuer warp
      mbar_init(&ctrl->v_full[i], 1);
      mbar_init(&ctrl->v_empty[i], p.qtiles > 1 ? 2 : 1);
      mbar_init(&ctrl->s_full[i], 1);
      mbar_init(&ctrl->p_full[i], kGroupWarps);
      mbar_init(&ctrl->p_empty[i], 1);
      mbar_init(&ctrl->o_full[i], 1);
      mbar_init(&ctrl->o_empty[i], kGroupWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 2) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctrl->tmem_base;
  const bool two = p.qtiles > 1;                 // does query tile 1 of a pair ever exist
  const uint32_t kv_bytes = (uint32_t)(p.Lkp * kD * 2);

  // item -> (b, h, pair of query tiles, key block); the key block is the fastest index: neighbouring CTAs share Q in L2
  // (32-bit arithmetic: the host checks items < 2^31; 64-bit divisions cost the softmax warps ~1.5k cycles per item)
  auto decode = [&](long long it, int& b, int& h, int& pr, int& kvb) {
    uint32_t item = blockIdx.x + (uint32_t)it * gridDim.x;
    kvb = 0; pr = 0;
    if (p.nkv > 1) { const uint32_t t = item / (uint32_t)p.nkv; kvb = (int)(item - t * (uint32_t)p.nkv); item = t; }
    if (p.pairs > 1) { const uint32_t t = item / (uint32_t)p.pairs; pr = (int)(item - t * (uint32_t)p.pairs); item = t; }
    const uint32_t bq = item / (uint32_t)p.heads;
    h = (int)(item - bq * (uint32_t)p.heads);
    b = (int)bq;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    // Q/K of item i+1 are requested before V of item i: their slot was released an item earlier (after S(i-1)), so the S MMA of the
    // next item never waits for memory; V(i) is only needed by P V(i), after the softmax of item i.
    if (lane == 0) {
      auto load_qk = [&](long long it) {
        const int s = (int)(it & 1);
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        int b, h, pr, kvb;
        decode(it, b, h, pr, kvb);
        const bool has1 = (2 * pr + 1) < p.qtiles;
        mbar_wait_lean(&ctrl->kq_empty[s], ph ^ 1u);
        uint8_t* st = sStage + (size_t)s * kStageBytes;
        mbar_expect_tx(&ctrl->kq_full[s], (uint32_t)(kQBytes * (has1 ? 2 : 1)) + kv_bytes);
        tma_load_3d(st, &tmQ, &ctrl->kq_full[s], h * kD, (2 * pr) * kQTile, b);
        if (has1) tma_load_3d(st + kQBytes, &tmQ, &ctrl->kq_full[s], h * kD, (2 * pr + 1) * kQTile, b);
        tma_load_3d(st + 2 * kQBytes, &tmK, &ctrl->kq_full[s], h * kD, kvb * p.KB, b);   // rows past T are zero-filled
      };
      auto load_v = [&](long long it) {
        const int s = (int)(it & 1);
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        int b, h, pr, kvb;
        decode(it, b, h, pr, kvb);
        mbar_wait_lean(&ctrl->v_empty[s], ph ^ 1u);
        mbar_expect_tx(&ctrl->v_full[s], kv_bytes);
        tma_load_3d(sStage + (size_t)s * kStageBytes + 2 * kQBytes + kKVBytes, &tmV, &ctrl->v_full[s], h * kD, kvb * p.KB, b);
      };
      if (n_my > 0) load_qk(0);
      for (long long it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) load_qk(it + 1);
        load_v(it);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ---------------- MMA issuers: warp 1 for softmax group 0, warp 3 for group 1 ----------------
    // (whole warp in uniform control flow, one elected lane issues: umma.cuh).  Both S_g(i+1) and P_g V(i) hang on ONE event, "group g
    // has written P_g(i)", so each issuer blocks on p_full[g] (hardware-suspended try_wait: ~100 cycles from the arrive to the first
    // MMA; an event loop polling four barriers per group needed ~900) and issues S_g(i+1) first — the group's warps are waiting for
    // it — then the 14 small P V MMAs, whose ISSUE (~900 cycles next to four busy softmax warps on the same scheduler) is the slow
    // part: with one issuer for both groups it sat in front of the other group's S.
    //   S_g(i+1)  needs P_g(i) written (=> S_g(i) fully consumed) and the Q/K slot of item i+1
    //   PV_g(i)   needs P_g(i) written, V(i), and the output epilogue of item i-1 done with O_g
    const int g = warp >> 1;
    if (g == 0 || two) {
      const uint32_t idesc_s = make_idesc_bf16(kQTile, p.Lkp);
      const uint32_t idesc_o = make_idesc_bf16(kQTile, kD) | (1u << 16);   // b_major = MN (V as loaded)
      const int ksteps = p.Lkp >> 4;
      const uint32_t d_s = tmem + (uint32_t)(g * kSCols), d_o = tmem + kOCol + (uint32_t)(g * kD);
      auto issue_s = [&](long long it) {   // S_g(it) = Q_g K^T; the Q/K slot frees when both groups' S MMAs have finished
        const uint8_t* st = sStage + (size_t)(it & 1) * kStageBytes;
        const uint64_t dq = make_kmajor_desc<64>(smem_u32(st + (size_t)g * kQBytes));
        const uint64_t dk = make_kmajor_desc<64>(smem_u32(st + 2 * kQBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ss_warp(d_s, dq + 2ull * k, dk + 2ull * k, idesc_s, (uint32_t)(k != 0));
        umma_commit_warp(&ctrl->s_full[g]);
        umma_commit_warp(&ctrl->kq_empty[it & 1]);
      };
      auto wait_warp = [&](uint64_t* bar, uint32_t parity) { mbar_wait_lean(bar, parity); __syncwarp(); };
      TR_INIT
      if (n_my > 0) {
        wait_warp(&ctrl->kq_full[0], 0u);
        tc_fence_after();
        issue_s(0);
      }
      for (long long i = 0; i < n_my; ++i) {
        TR(2)
        wait_warp(&ctrl->p_full[g], (uint32_t)i & 1u);
        TR(0)
        EV(4 + g, i, 0)
        if (i + 1 < n_my) {
          wait_warp(&ctrl->kq_full[(i + 1) & 1], (uint32_t)((i + 1) >> 1) & 1u);
          EV(4 + g, i, 1)
          tc_fence_after();
          issue_s(i + 1);
        }
        TR(1)
        EV(4 + g, i, 2)
        if (i > 0) wait_warp(&ctrl->o_empty[g], (uint32_t)(i - 1) & 1u);
        EV(4 + g, i, 3)
        wait_warp(&ctrl->v_full[i & 1], (uint32_t)(i >> 1) & 1u);
        EV(4 + g, i, 4)
        tc_fence_after();
        // descriptors advance by plain adds on the 16-byte-unit address field (all operands sit below 256 KB)
        uint64_t da = make_kmajor_desc<128>(smem_u32(sP + (size_t)g * kPBytes));
        uint64_t db = make_mnmajor_sw64_desc(smem_u32(sStage + (size_t)(i & 1) * kStageBytes + 2 * kQBytes + kKVBytes));
        for (int s = 0; s < ksteps; ++s) {
          umma_bf16_ss_warp(d_o, da, db, idesc_o, (uint32_t)(s != 0));
          da += ((s & 3) == 3) ? (uint64_t)((kPTileBytes >> 4) - 6) : 2ull;   // next 16 keys: +32 B inside a 64-key tile, else the next tile
          db += (uint64_t)((16 * kD * 2) >> 4);
        }
        umma_commit_warp(&ctrl->o_full[g]);
        umma_commit_warp(&ctrl->p_empty[g]);
        umma_commit_warp(&ctrl->v_empty[i & 1]);   // the V slot frees when both groups' P V MMAs of item i have finished
        EV(4 + g, i, 5)
#if defined(LMV_ATTN_TRACE) && LMV_ATTN_TRACE == 3
        wait_warp(&ctrl->o_full[g], (uint32_t)i & 1u);   // probe: when did the P V MMAs complete (the issuer has nothing else to do until the next p_full)
        EV(4 + g, i, 6)
#endif
      }
      TR(2)
      if (g == 0) { TR_FLUSH(4) }
    }
  } else if (warp >= kFirstSoftmaxWarp) {
    // ---------------- softmax groups ----------------
    const int g = (warp - kFirstSoftmaxWarp) / kGroupWarps;     // group == query tile of the pair
    const int hcol = ((warp - kFirstSoftmaxWarp) >> 2) & 1;      // which half of the key columns (and of the 32 output columns)
    const int q = warp & 3;                            // TMEM lane quarter
    const int r = q * 32 + lane;                       // row inside the tile == TMEM lane
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = t_row + (uint32_t)(g * kSCols), t_o = t_row + kOCol + (uint32_t)(g * kD);
    uint8_t* pg = sP + (size_t)g * kPBytes + (size_t)r * 128;
    const int nchunk = (p.Lkp + 31) >> 5;
    const int nsplit = (nchunk + 1) >> 1;
    const int cbeg = hcol ? nsplit : 0, cend = hcol ? nchunk : nsplit;   // this warp's 32-column chunks
    float* x_max = sXchg + ((g * 2 + hcol) * 3 + 0) * kQTile;       // own slots; the partner's are at (hcol ^ 1)
    float* x_sum = sXchg + ((g * 2 + hcol) * 3 + 1) * kQTile;
    const float* y_max = sXchg + ((g * 2 + (hcol ^ 1)) * 3 + 0) * kQTile;
    const float* y_sum = sXchg + ((g * 2 + (hcol ^ 1)) * 3 + 1) * kQTile;
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kGroupWarps * 32) : "memory"); };
    // deferred output epilogue state of the previous item
    float prev_ps = 0.f;          // this warp's partial row sum
    float prev_mxs = 0.f;         // row max * scale * log2e (split-KV partials)
    uint32_t prev_par = 0;        // which x_sum / y_sum slot holds the previous item's partial sums
    bf16* prev_out = nullptr;     // final output row (single key block)
    long long prev_prow = -1;     // partial row index (split-KV)
    bool prev_any = false;
    uint32_t n_done = 0;          // items this group has processed (phase counter of its barriers)
    auto output_epilogue = [&]() {
      mbar_wait_lean(&ctrl->o_full[g], (n_done - 1) & 1u);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_x16(t_o + (uint32_t)(hcol * 16), v);     // this warp's 16 of the 32 output columns
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->o_empty[g]);
      if (prev_out) {
        const float inv = 1.f / (prev_ps + y_sum[prev_par * kQTile + r]);
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
        bf16* dst = prev_out + hcol * 16;
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);   // (one 256-bit store instead was measured 4 % slower here)
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
      } else if (prev_prow >= 0) {
        // split-KV: unnormalised partial of this key block; the merge kernel rescales and sums the blocks
        // lane = row stores of 64 contiguous bytes: two 256-bit stores where the workspace allows it (the LSU pays per cache line)
        float* dst = p.part_o + prev_prow * kD + hcol * 16;
        if ((reinterpret_cast<uintptr_t>(p.part_o) & 31) == 0) {
          st_global_256(dst, reinterpret_cast<const uint32_t (&)[8]>(v[0]));
          st_global_256(dst + 8, reinterpret_cast<const uint32_t (&)[8]>(v[8]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
        if (hcol == 0) p.part_ml[prev_prow] = make_float2(prev_mxs, prev_ps + y_sum[prev_par * kQTile + r]);
      }
    };
    const float2 sc2 = make_float2(p.scale_log2e, p.scale_log2e);
    // per-(query pair, key block) state: recomputed only when the pair or the key block changes (never, in the single-block case)
    int st_pr = -1, st_kvb = -1, tile = 0, row = 0, kbeg = 0, kend = 0;
    bool rok = false, warp_active = false;
    uint32_t full_mask = 0, part_mask = 0, vis = 0;
    TR_INIT

    for (long long it = 0; it < n_my; ++it) {
      if (g >= p.qtiles) break;                        // a single query tile per (image, head): nothing for group 1
      int b, h, pr, kvb;
      decode(it, b, h, pr, kvb);
      if (pr != st_pr || kvb != st_kvb) {
        st_pr = pr; st_kvb = kvb;
        tile = 2 * pr + g;
        row = tile * kQTile + r;
        rok = row < p.T;
        // key range of this row inside the key block.  One block (T <= 224): image tokens attend image tokens, meta tokens attend
        // meta tokens (rows that do not exist behave like image rows).  Split-KV: every row sees the block's keys below T.
        if (p.nkv == 1) {
          kbeg = (rok && row >= p.N) ? p.N : 0;
          kend = (rok && row >= p.N) ? p.T : p.N;
        } else {
          kbeg = 0;
          kend = max(0, min(p.KB, p.T - kvb * p.KB));
        }
        warp_active = tile * kQTile + q * 32 < p.T;   // warp-uniform: any valid row in this warp
        // Visibility of the 32-key chunks, per lane (= row): fully visible, fully hidden, or partial.
        //   bit c of full_mask: every lane of this warp sees all 32 keys of chunk c (no masking at all)
        //   bit c of part_mask: some lane sees only part of chunk c -> per-element masking of the loaded registers (at most the chunk
        //                       that holds the image/meta boundary and the last chunk of the key block)
        //   otherwise every lane sees the chunk entirely or not at all (the warp that holds image AND meta rows): bit c of vis says
        //   which, and a hidden chunk costs nothing — pass 1 skips its maximum, pass 2 runs it with (scale, -max) = (0, -inf), so P = 0
        full_mask = 0; part_mask = 0; vis = 0;
        for (int c = cbeg; c < cend; ++c) {
          const int c0 = c * 32;
          const bool l_full = kbeg <= c0 && c0 + 32 <= kend, l_none = kend <= c0 || kbeg >= c0 + 32;
          if (l_full) vis |= 1u << c;
          if (__all_sync(0xffffffffu, l_full)) full_mask |= 1u << c;
          // (columns at or beyond Lkp were never written by the S MMA: stale TMEM bits, possibly NaN, must go through the select)
          else if (__any_sync(0xffffffffu, !(l_full || l_none)) || c0 + 32 > p.Lkp) part_mask |= 1u << c;
        }
      }
      TR(0)
      if (q == 0) { EV(g * 2 + hcol, it, 0) }
      mbar_wait_lean(&ctrl->s_full[g], n_done & 1u);
      tc_fence_after();
      TR(1)
      if (q == 0) { EV(g * 2 + hcol, it, 1) }
      float psum = 0.f, mxs_item = 0.f;
      float pm = -INFINITY;
      auto mask_chunk = [&](uint32_t (&v)[32], int c0) {
        // keys outside this row's segment get a score of -inf (exp2 -> exactly 0); applied to the loaded registers of the few
        // chunks that are not fully visible, so that each pass has ONE math body (instruction-cache footprint matters here)
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < kbeg || c0 + j >= kend) v[j] = 0xff800000u;
      };
      if (warp_active) {
        // ---- pass 1: partial row max over this warp's chunks (four independent chains) ----
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
        for (int c = cbeg; c < cend; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_s + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          if ((part_mask >> c) & 1u) mask_chunk(v, c * 32);
          const bool hide = !((full_mask | part_mask) >> c & 1u) && !((vis >> c) & 1u);   // this lane sees nothing of an unmasked chunk
          float c0m = m0, c1m = m1, c2m = m2, c3m = m3;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {   // FMNMX3: two new values per instruction
            c0m = fmax3(c0m, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            c1m = fmax3(c1m, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            c2m = fmax3(c2m, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            c3m = fmax3(c3m, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
          }
          if (!hide) { m0 = c0m; m1 = c1m; m2 = c2m; m3 = c3m; }
        }
        pm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      x_max[r] = pm;
      TR(2)
      if (q == 0) { EV(g * 2 + hcol, it, 2) }
      group_sync();                                     // partial maxima (and the previous item's partial sums) are visible
      TR(3)
      // the P V MMA of the previous item must have finished reading this group's P tiles
      mbar_wait_lean(&ctrl->p_empty[g], (n_done & 1u) ^ 1u);
      // (Measured: forcing the two groups to take turns at the exp pass does not help — one group alone is latency-bound in it, ~3.5k
      // cycles per item either way — so the groups are left to run in lockstep, where the four warps of a scheduler keep MUFU busy
      // (~80% during pass 2).  Walking the chunks in 16-column halves with the next tcgen05.ld in flight changes nothing either.)
      TR(4)
      if (q == 0) { EV(g * 2 + hcol, it, 3) }
      if (warp_active) {
        const float mxs = fmaxf(pm, y_max[r]) * p.scale_log2e;
        mxs_item = mxs;
        const float2 nm2 = make_float2(-mxs, -mxs);
        // ---- pass 2: p = exp2(s * scale - max), partial row sum (two packed chains), bf16 P tile ----
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = cbeg; c < cend; ++c) {
          const int c0 = c * 32;
          uint32_t v[32];
          tmem_ld_x32(t_s + (uint32_t)c0, v);
          tmem_ld_wait();
          if ((part_mask >> c) & 1u) mask_chunk(v, c0);
          const bool hide = !((full_mask | part_mask) >> c & 1u) && !((vis >> c) & 1u);
          const float2 scc = hide ? make_float2(0.f, 0.f) : sc2, nmc = hide ? make_float2(-INFINITY, -INFINITY) : nm2;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {   // eight scores: six exponentials on MUFU, one pair on the FMA pipe
            float2 a0 = ffma2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), scc, nmc);
            float2 a1 = ffma2(make_float2(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3])), scc, nmc);
            float2 a2 = ffma2(make_float2(__uint_as_float(v[2 * j + 4]), __uint_as_float(v[2 * j + 5])), scc, nmc);
            float2 a3 = ffma2(make_float2(__uint_as_float(v[2 * j + 6]), __uint_as_float(v[2 * j + 7])), scc, nmc);
            a0.x = ex2_approx(a0.x); a0.y = ex2_approx(a0.y);
            a1.x = ex2_approx(a1.x); a1.y = ex2_approx(a1.y);
            a2.x = ex2_approx(a2.x); a2.y = ex2_approx(a2.y);
            a3 = kPolyExp ? exp2_poly2(a3) : make_float2(ex2_approx(a3.x), ex2_approx(a3.y));
            s0 = fadd2(s0, fadd2(a0, a2));
            s1 = fadd2(s1, fadd2(a1, a3));
            pk[j] = pack_bf16x2(a0.x, a0.y);
            pk[j + 1] = pack_bf16x2(a1.x, a1.y);
            pk[j + 2] = pack_bf16x2(a2.x, a2.y);
            pk[j + 3] = pack_bf16x2(a3.x, a3.y);
          }
          // P[r, c0 .. c0+31] -> tile (c0 / 64), 16-byte chunks ((c0 % 64) / 8) .. +3, XOR-swizzled with (r % 8)
          uint8_t* tile_p = pg + (size_t)(c0 >> 6) * kPTileBytes;
          const int ch = (c0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(tile_p + (((ch + j) ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        psum = (s0.x + s0.y) + (s1.x + s1.y);
      } else {
        // rows of this warp do not exist: P = 0 keeps the (unused) accumulator rows finite
        for (int c = cbeg; c < cend; ++c) {
          uint8_t* tile_p = pg + (size_t)((c * 32) >> 6) * kPTileBytes;
          const int ch = ((c * 32) & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(tile_p + (((ch + j) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      const uint32_t par = n_done & 1u;
      x_sum[par * kQTile + r] = psum;                  // read by the partner warp in its deferred output epilogue
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->p_full[g]);
      TR(5)
      if (q == 0) { EV(g * 2 + hcol, it, 4) }
      // ---- deferred output epilogue of the previous item, while the tensor pipe works on this one ----
      if (prev_any) output_epilogue();
      TR(6)
      if (q == 0) { EV(g * 2 + hcol, it, 5) }
      ++n_done;
      prev_any = true;
      prev_ps = psum;
      prev_mxs = mxs_item;
      prev_par = par;
      prev_out = nullptr;
      prev_prow = -1;
      if (rok && warp_active) {
        if (p.nkv == 1) prev_out = p.out + (long long)b * p.o_bs + (long long)row * p.o_rs + h * kD;
        else prev_prow = (((long long)b * p.heads + h) * (p.qtiles * kQTile) + row) * p.nkv + kvb;
      }
    }
    if (prev_any) {
      group_sync();                                     // the partner's partial sums of the last item
      output_epilogue();
    }
    if (q == 0) { TR_FLUSH(g * 2 + hcol) }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// split-KV merge: one warp per (image, head, query row), lane = channel; fixed block order (deterministic)
__global__ void __launch_bounds__(256)
attention_self_merge_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, bf16* __restrict__ out, long long o_bs, int o_rs,
                            int B, int heads, int T, int Tpad, int nkv) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= (long long)B * heads * T) return;
  const int row = (int)(idx % T);
  const long long bh = idx / T;
  const int h = (int)(bh % heads), b = (int)(bh / heads);
  const long long base = (bh * Tpad + row) * nkv;
  float m = -INFINITY;
  for (int k = 0; k < nkv; ++k) m = fmaxf(m, part_ml[base + k].x);
  float l = 0.f, o = 0.f;
  for (int k = 0; k < nkv; ++k) {
    const float2 ml = part_ml[base + k];
    const float w = exp2f(ml.x - m);
    l = fmaf(ml.y, w, l);
    o = fmaf(part_o[(base + k) * kD + lane], w, o);
  }
  out[(long long)b * o_bs + (long long)row * o_rs + h * kD + lane] = __float2bfloat16(o / l);
}

PerDeviceOnce g_attr_once;

}  // namespace

// T <= 224: one key block (image + meta segments allowed).  Larger T (e.g. 1024 tokens at 512x512): split-KV over
// ceil(T / 224) key blocks, plain self-attention only (N == T) and an even number of 128-row query tiles.
static int self_nkv(int T) { return (T + kMaxKeys - 1) / kMaxKeys; }
static int self_kb(int T) { const int n = self_nkv(T); return (((T + n - 1) / n) + 15) & ~15; }

bool attention_self_supported(const AttnArgs& a, int T, int N) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int qtiles = (T + kQTile - 1) / kQTile;
  const bool shape_ok = T >= 1 && N >= 1 && N <= T && (T <= kMaxKeys || (N == T && self_kb(T) <= kMaxKeys));   // (an odd number of query tiles leaves group 1 idle in the last pair)
  return shape_ok && a.B >= 1 && a.heads >= 1 && al16(a.q) && al16(a.k) && al16(a.v) && al16(a.out) &&
         a.q_rs % 8 == 0 && a.k_rs % 8 == 0 && a.v_rs % 8 == 0 && a.o_rs % 8 == 0 && a.q_bs % 8 == 0 && a.k_bs % 8 == 0 &&
         a.v_bs % 8 == 0 && a.o_bs % 8 == 0 && a.q_rs >= a.heads * kD && a.k_rs >= a.heads * kD && a.v_rs >= a.heads * kD;
}

size_t attention_self_workspace(int B, int heads, int T) {
  const int nkv = self_nkv(T);
  if (nkv <= 1) return 0;
  const size_t rows = (size_t)B * heads * (((T + kQTile - 1) / kQTile) * kQTile) * nkv;
  return rows * (kD * sizeof(float) + sizeof(float2));
}

// a.Lq / a.Lk are ignored: T rows per image, the first N attend among themselves, the remaining T - N among themselves
int attention_self_run(const AttnArgs& a, int T, int N, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!attention_self_supported(a, T, N)) return fail(LMV_ERR_UNSUPPORTED, "attention_self: unsupported shape / alignment");
  LMV_REQUIRE(attention_self_workspace(a.B, a.heads, T) == 0 ||
                  (workspace && workspace_bytes >= attention_self_workspace(a.B, a.heads, T) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0),
              "attention_self: split-KV workspace missing or too small");
  LMV_CUDA_OK(g_attr_once.run([] { return cudaFuncSetAttribute(attention_self_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); }));
  SelfParams p;
  p.out = a.out; p.o_bs = a.o_bs; p.o_rs = a.o_rs;
  p.B = a.B; p.heads = a.heads; p.T = T; p.N = N;
  p.nkv = self_nkv(T);
  p.KB = p.nkv == 1 ? ((T + 15) & ~15) : self_kb(T);
  p.Lkp = p.KB;
  p.qtiles = (T + kQTile - 1) / kQTile;
  p.pairs = (p.qtiles + 1) / 2;
  p.items = (long long)a.B * a.heads * p.pairs * p.nkv;
  LMV_REQUIRE(p.items < (1ll << 31), "attention_self: too many work items");
  p.part_o = nullptr; p.part_ml = nullptr;
  if (p.nkv > 1) {
    const size_t rows = (size_t)a.B * a.heads * (p.qtiles * kQTile) * p.nkv;
    p.part_o = static_cast<float*>(workspace);
    p.part_ml = reinterpret_cast<float2*>(p.part_o + rows * kD);
  }
  p.scale_log2e = a.scale * 1.4426950408889634f;
  CUtensorMap tq, tk, tv;
  auto enc = [&](CUtensorMap* m, const bf16* base, long long bs, int rs, int box_rows) {
    uint64_t dims[3] = {(uint64_t)a.heads * kD, (uint64_t)T, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
    uint32_t box[3] = {kD, (uint32_t)box_rows, 1};
    return encode_tmap_bf16(m, base, 3, dims, strides, box, 64);
  };
  int rc;
  if ((rc = enc(&tq, a.q, a.q_bs, a.q_rs, kQTile))) return rc;
  if ((rc = enc(&tk, a.k, a.k_bs, a.k_rs, p.Lkp))) return rc;
  if ((rc = enc(&tv, a.v, a.v_bs, a.v_rs, p.Lkp))) return rc;
  const int grid = (int)std::min<long long>(p.items, device_sm_count());
  LMV_CUDA_OK(launch_kernel(attention_self_kernel, dim3(grid), dim3(kThreads), (size_t)(kSmemBytes), s, tq, tk, tv, p));
  LMV_CUDA_OK(cudaGetLastError());
  if (p.nkv > 1) {
    const long long mrows = (long long)a.B * a.heads * T;
    LMV_CUDA_OK(launch_kernel(attention_self_merge_kernel, dim3((unsigned)((mrows + 7) / 8)), dim3(256), (size_t)0, s, p.part_o, p.part_ml, a.out, a.o_bs,
                              a.o_rs, a.B, a.heads, T, p.qtiles * kQTile, p.nkv));
    LMV_CUDA_OK(cudaGetLastError());
  }
  return LMV_OK;
}

}  // namespace lmv

#if defined(LMV_ATTN_TRACE) && LMV_ATTN_TRACE == 3
extern "C" int lmv_debug_attn_events(long long* host, int n) {
  if (n > 6 * 16 * 8) n = 6 * 16 * 8;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  return cudaMemcpyFromSymbol(host, lmv::g_attn_events, sizeof(long long) * n) != cudaSuccess;
}
#endif
#ifdef LMV_ATTN_TRACE
// copies out and clears the [148 CTAs][4 warps][8 counters] cycle table of the traced launches (debug builds only)
extern "C" int lmv_debug_attn_trace(unsigned long long* host, int n) {
  static unsigned long long zero[148 * 5 * 8];
  if (n > 148 * 5 * 8) n = 148 * 5 * 8;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(host, lmv::g_attn_trace, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(lmv::g_attn_trace, zero, sizeof(zero)) != cudaSuccess;
}
#endif
