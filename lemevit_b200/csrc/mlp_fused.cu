// out = x + W2 · gelu(W1' · LN(x) + b1') + b2      — the whole MLP branch of a LeMeBlock in ONE kernel.
// Reference: `self.mlp = nn.Sequential(Linear, Identity, GELU, Linear)` (models/lemevit.py:526-530) applied as
// `x = x + self.drop_path(self.mlp(self.norm2(x)))` (:562,564 D blocks; :601 C blocks; :633,635 S blocks).
// MLP is 2/3 of the linear FLOPs of the network; unfused, the [rows, 4C] hidden activation is written to and read
// back from HBM (8C bytes/token against 4C for x in + x out).  Here it only ever exists as TMEM accumulators and
// bf16 shared-memory tiles.
//
// sm_100a design — persistent CTA (one per SM) over 128-row tiles, hidden dimension in chunks of 128, warp-specialised:
//   warp 0      TMA producer A: the X tile (128 x C as K-blocks of 64, 128B swizzle) + ring 1 of W1' boxes [128 hidden x 64 k]
//   warp 1      TMA producer B: ring 2 of W2 boxes [n2 out x 64 hidden]   (its own ring: W1' prefetch never queues behind
//               W2 boxes that wait for the epilogue)
//   warp 2      one thread issues tcgen05.mma:  fc1(g): acc1[g&1][128 x 128] = X · W1'[g]^T      (K = C)
//                                               fc2(g): acc2[128 x C]      += H_g · W2[:, g]^T    (K = 128)
//               issue order fc1(g+1), fc2(g): the tensor pipe works on the next hidden chunk while the epilogue warps turn
//               chunk g into bf16
//   warps 4..19 epilogue (4 warps per TMEM lane quarter, each owning 32 of the chunk's 128 columns): tcgen05.ld acc1 -> LayerNorm fold
//               (r·acc + (−r·mu)·colsum + b1, constants in smem) -> GELU (packed fp32 math + MUFU.TANH) -> bf16 ->
//               128B-swizzled K-major smem tiles H_g (the A operand of fc2).  After the last chunk of a tile they run the
//               output epilogue: acc2 + b2 + residual -> bf16 -> global.
// TMEM: acc2 in columns [0, C <= 256), the two acc1 buffers in columns [256, 512).
// Wide variant (C = 384, the 'S' blocks of stage 3; HC = 64): acc2 [0, 384), two 64-column acc1 buffers [384, 512); fc2 is
// issued as two N = C/2 halves; one X buffer (96 KB); the epilogue constants come through L1 (__ldg broadcasts) instead of
// shared memory, and ONE ring of four 24 KB slots carries both weights in the order the tensor pipe consumes them —
// W1'(j+1) as two 3-D boxes of three K-blocks, then the two [C/2 x 64] halves of W2(j) — so every byte of the 96 KB left
// beside X and the hidden buffers is lookahead.  Per 128-row tile the weights (2.4 MB) stream from L2; unfused, the same
// block writes and re-reads a [rows, 4C] hidden activation that no longer fits the L2.
// LayerNorm (norm2) is folded exactly as in gemm.cu: the producer of x emitted per-row (sum, sum^2) partials.
#include <algorithm>
#include <mutex>

#include "common.h"
#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int BM = 128;          // rows per tile
constexpr int BK = 64;           // K-block (one 128B swizzle span of bf16)
constexpr int kEpiWarps = 16;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);
constexpr int kMaxSlots = 8;
constexpr int kXBlockBytes = BM * BK * 2;      // 16 KB
constexpr int kSmemLimit = 227 * 1024;

struct Ctrl {
  uint64_t x_full[2], x_empty[2];
  uint64_t w1_full[kMaxSlots], w1_empty[kMaxSlots];
  uint64_t w2_full[kMaxSlots], w2_empty[kMaxSlots];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t hid_full[2], hid_empty[2];
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_base;
};
static_assert(sizeof(Ctrl) <= 1024, "control block");



template <int HC>   // hidden columns per chunk
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_tcgen05(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  constexpr bool WIDE = HC == 64;
  constexpr int kHidBytes = BM * HC * 2;         // hidden buffer: HC / 64 K-blocks of [128 rows x 128 B]
  constexpr int kW1KbBytes = HC * BK * 2;        // one K-block of a W1' box
  constexpr int kAcc1Col = WIDE ? 384 : 256;     // first TMEM column of the two fc1 accumulators
  constexpr int kW1Kpb = WIDE ? 3 : 1;           // K-blocks per W1' box
  constexpr int kW1BoxBytes = kW1Kpb * kW1KbBytes;   // 16 KB (narrow) / 24 KB (wide: == one W2 half, the unified ring's slot)
  const int x_bytes = p.kb1 * kXBlockBytes;
  float4* sCB = reinterpret_cast<float4*>(smem + 1024);               // [Hd/2] (cs0, cs1, b0, b1) of two hidden columns   (narrow only)
  float* sB2 = reinterpret_cast<float*>(smem + 1024 + (size_t)p.Hd * 8);   // [C]                                            (narrow only)
  uint8_t* sX = smem + 1024 + p.const_bytes;                          // nx buffers of kb1 K-blocks
  uint8_t* sH = sX + (size_t)p.nx * x_bytes;                          // nh hidden buffers
  uint8_t* sW1 = sH + (size_t)p.nh * kHidBytes;                       // ring 1
  uint8_t* sW2 = sW1 + (size_t)p.n1slots * kW1BoxBytes;               // ring 2 (narrow only)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int J = p.chunks;
  const int my_tiles = ((int)blockIdx.x < p.tiles) ? (p.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->x_full[i], 1);
      mbar_init(&ctrl->x_empty[i], p.res_smem ? 1 + kEpiWarps : 1);   // + the output epilogue when it reads the residual from the X tile
      mbar_init(&ctrl->acc1_full[i], 1);
      mbar_init(&ctrl->acc1_empty[i], kEpiWarps);
      mbar_init(&ctrl->hid_full[i], kEpiWarps);
      mbar_init(&ctrl->hid_empty[i], 1);
    }
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&ctrl->w1_full[i], 1);
      mbar_init(&ctrl->w1_empty[i], 1);
      mbar_init(&ctrl->w2_full[i], 1);
      mbar_init(&ctrl->w2_empty[i], 1);
    }
    mbar_init(&ctrl->acc2_full, 1);
    mbar_init(&ctrl->acc2_empty, kEpiWarps);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  pdl_wait();   // barrier init / TMEM allocation above overlapped the previous kernel's tail
  // epilogue constants -> shared memory (read on every chunk by every epilogue warp)
  if constexpr (!WIDE) {
    for (int i = threadIdx.x; i < p.Hd / 2; i += kThreads) {
      const float2 b = __ldg(reinterpret_cast<const float2*>(p.b1) + i);
      const float2 c = p.cs1 ? __ldg(reinterpret_cast<const float2*>(p.cs1) + i) : make_float2(0.f, 0.f);
      sCB[i] = make_float4(c.x, c.y, b.x, b.y);
    }
    for (int i = threadIdx.x; i < p.C; i += kThreads) sB2[i] = __ldg(p.b2 + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer A: X tiles + W1' ring ----------------
    if (lane == 0) {
      int slot = 0, xb = 0;
      uint32_t wphase = 0, xphase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int t = blockIdx.x + it * gridDim.x;
        mbar_wait(&ctrl->x_empty[xb], xphase ^ 1u, 40);
        mbar_expect_tx(&ctrl->x_full[xb], (uint32_t)x_bytes);
        for (int kb = 0; kb < p.kb1; ++kb)
          tma_load_2d(sX + (size_t)xb * x_bytes + (size_t)kb * kXBlockBytes, &tmX, &ctrl->x_full[xb], kb * BK, t * BM);
        if (++xb == p.nx) { xb = 0; xphase ^= 1u; }
        if constexpr (!WIDE) {
          for (int j = 0; j < J; ++j)
            for (int kb = 0; kb < p.kb1; ++kb) {
              mbar_wait(&ctrl->w1_empty[slot], wphase ^ 1u, 41);
              mbar_expect_tx(&ctrl->w1_full[slot], (uint32_t)kW1BoxBytes);
              tma_load_2d(sW1 + (size_t)slot * kW1BoxBytes, &tmW1, &ctrl->w1_full[slot], kb * BK, j * HC);
              if (++slot == p.n1slots) { slot = 0; wphase ^= 1u; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- TMA producer B: W2 ring ----------------
    if (lane == 0) {
      int slot = 0;
      uint32_t wphase = 0;
      if constexpr (!WIDE) {
        for (int it = 0; it < my_tiles; ++it)
          for (int j = 0; j < J; ++j)
            for (int kb = 0; kb < HC / BK; ++kb)
              for (int q = 0; q < p.nparts2; ++q) {
                mbar_wait(&ctrl->w2_empty[slot], wphase ^ 1u, 42);
                mbar_expect_tx(&ctrl->w2_full[slot], (uint32_t)(p.n2 * BK * 2));
                tma_load_2d(sW2 + (size_t)slot * p.slot2_bytes, &tmW2, &ctrl->w2_full[slot], j * HC + kb * BK, q * p.n2);
                if (++slot == p.n2slots) { slot = 0; wphase ^= 1u; }
              }
      } else {
        // unified ring (the w1_* barriers), in the issuer's consumption order: W1'(0), W1'(1), W2(0), W1'(2), W2(1), ..., W2(J-1)
        auto put_w1 = [&](int j) {
          for (int bx = 0; bx < p.w1_boxes; ++bx) {
            mbar_wait(&ctrl->w1_empty[slot], wphase ^ 1u, 41);
            mbar_expect_tx(&ctrl->w1_full[slot], (uint32_t)kW1BoxBytes);
            tma_load_3d(sW1 + (size_t)slot * kW1BoxBytes, &tmW1, &ctrl->w1_full[slot], 0, j * HC, bx * kW1Kpb);   // (k in block, hidden row, K-block)
            if (++slot == p.n1slots) { slot = 0; wphase ^= 1u; }
          }
        };
        auto put_w2 = [&](int j) {
          for (int q = 0; q < p.nparts2; ++q) {
            mbar_wait(&ctrl->w1_empty[slot], wphase ^ 1u, 42);
            mbar_expect_tx(&ctrl->w1_full[slot], (uint32_t)(p.n2 * BK * 2));
            tma_load_2d(sW1 + (size_t)slot * kW1BoxBytes, &tmW2, &ctrl->w1_full[slot], j * HC, q * p.n2);
            if (++slot == p.n1slots) { slot = 0; wphase ^= 1u; }
          }
        };
        for (int it = 0; it < my_tiles; ++it) {
          put_w1(0);
          for (int j = 1; j < J; ++j) { put_w1(j); put_w2(j - 1); }
          put_w2(J - 1);
        }
      }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer: whole warp in uniform control flow, one elected lane issues (umma.cuh) ----------------
    {
      const uint32_t idesc1 = make_idesc_bf16(BM, HC);
      const uint32_t idesc2 = make_idesc_bf16(BM, p.n2);
      const int G = my_tiles * J;
      int s1 = 0, s2 = 0, xb = 0;
      uint32_t ph1 = 0, ph2 = 0, xphase = 0;
      uint32_t xaddr = 0;
      int i1 = 0, i2 = 0;     // next fc1 / fc2 chunk (global over this CTA's tiles)
      auto fc1 = [&](int g) {
        const int j = g % J;
        if (j == 0) {   // first chunk of a tile: its X must have landed
          mbar_wait(&ctrl->x_full[xb], xphase, 50);
          xaddr = smem_u32(sX + (size_t)xb * x_bytes);
        }
        const uint32_t buf = (uint32_t)g & 1u, use = (uint32_t)g >> 1;
        mbar_wait(&ctrl->acc1_empty[buf], (use & 1u) ^ 1u, 51);
        tc_fence_after();
        const uint32_t d = tmem_base + kAcc1Col + buf * HC;
        for (int bx = 0; bx < p.w1_boxes; ++bx) {
          mbar_wait(&ctrl->w1_full[s1], ph1, 52);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kW1Kpb; ++kk) {
            const int kb = bx * kW1Kpb + kk;
            const uint64_t da = make_kmajor_desc<128>(xaddr + (uint32_t)kb * kXBlockBytes);
            const uint64_t db = make_kmajor_desc<128>(smem_u32(sW1 + (size_t)s1 * kW1BoxBytes + (size_t)kk * kW1KbBytes));
            const int ks = WIDE ? BK / 16 : min(BK / 16, (p.C - kb * BK) / 16);
            for (int k = 0; k < ks; ++k) umma_bf16_ss_warp(d, da + 2ull * k, db + 2ull * k, idesc1, (uint32_t)((kb | k) != 0));
          }
          umma_commit_warp(&ctrl->w1_empty[s1]);
          if (++s1 == p.n1slots) { s1 = 0; ph1 ^= 1u; }
        }
        umma_commit_warp(&ctrl->acc1_full[buf]);
        if (j == J - 1) {   // X tile no longer needed once the last fc1 of the tile retires
          umma_commit_warp(&ctrl->x_empty[xb]);
          if (++xb == p.nx) { xb = 0; xphase ^= 1u; }
        }
      };
      auto fc2 = [&](int g) {
        const int j = g % J;
        const uint32_t tile_it = (uint32_t)(g / J);
        const uint32_t hb = (uint32_t)g & 1u, use = (uint32_t)g >> 1;
        mbar_wait(&ctrl->hid_full[hb], use & 1u, 53);
        if (j == 0) mbar_wait(&ctrl->acc2_empty, (tile_it & 1u) ^ 1u, 54);
        tc_fence_after();
        for (int kb = 0; kb < HC / BK; ++kb) {
          const uint64_t da = make_kmajor_desc<128>(smem_u32(sH + (size_t)hb * kHidBytes + (size_t)kb * kXBlockBytes));
          for (int q = 0; q < p.nparts2; ++q) {
            const uint32_t d = tmem_base + (uint32_t)(q * p.n2);
            if constexpr (!WIDE) {
              mbar_wait(&ctrl->w2_full[s2], ph2, 55);
              tc_fence_after();
              const uint64_t db = make_kmajor_desc<128>(smem_u32(sW2 + (size_t)s2 * p.slot2_bytes));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_warp(d, da + 2ull * k, db + 2ull * k, idesc2, (uint32_t)((j | kb | k) != 0));
              umma_commit_warp(&ctrl->w2_empty[s2]);
              if (++s2 == p.n2slots) { s2 = 0; ph2 ^= 1u; }
            } else {   // unified ring: the next slot holds this half of W2(j)
              mbar_wait(&ctrl->w1_full[s1], ph1, 55);
              tc_fence_after();
              const uint64_t db = make_kmajor_desc<128>(smem_u32(sW1 + (size_t)s1 * kW1BoxBytes));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_warp(d, da + 2ull * k, db + 2ull * k, idesc2, (uint32_t)((j | kb | k) != 0));
              umma_commit_warp(&ctrl->w1_empty[s1]);
              if (++s1 == p.n1slots) { s1 = 0; ph1 ^= 1u; }
            }
          }
        }
        umma_commit_warp(&ctrl->hid_empty[hb]);
        if (j == J - 1) umma_commit_warp(&ctrl->acc2_full);
      };
      while (i2 < G) {
        // fc1 runs one chunk ahead of fc2 (two accumulators); with a single X buffer never run ahead into the next tile
        while (i1 < G && i1 < i2 + 2 && (p.nx == 2 || i1 / J == i2 / J)) fc1(i1++);
        fc2(i2++);
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ---------------- epilogue warps ----------------
    const int q = warp & 3;                        // TMEM lane quarter
    const int e = (warp - kFirstEpiWarp) >> 2;     // 0..3
    // this warp's HC / 4 of the chunk's hidden columns: columns e * HCW .. of the hidden tile (K-blocks of 64 columns, 16-byte chunks of 8)
    constexpr int HCW = HC / 4;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int rloc = q * 32 + lane;                // row inside the tile == TMEM lane
    // LayerNorm statistics of this thread's row are fetched one tile ahead (<= 4 partial pairs)
    float2 nst[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
    auto load_stats = [&](int tile) {
      const int r = tile * BM + rloc;
      if (r < p.R) {
        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + (long long)r * p.ln_parts;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < p.ln_parts) nst[k] = __ldg(st + k);
      }
    };
    if (p.ln_stats && my_tiles > 0) load_stats(blockIdx.x);
    // ---- output epilogue of tile iteration `oit` (all 16 warps): acc2 + b2 + residual -> out.  It runs AFTER the first hidden
    // chunk of the next tile, so the wait for the last fc2 of the tile is hidden behind useful work.
    auto output_epilogue = [&](int oit) {
      const int row = (blockIdx.x + oit * gridDim.x) * BM + rloc;
      const bool rok = row < p.R;
        bool waited = false;
        const int xb_o = oit & 1;      // res_smem implies two X buffers: tile `oit` sits in buffer oit & 1
        if (p.res_smem) mbar_wait(&ctrl->x_full[xb_o], (uint32_t)(oit >> 1) & 1u, 63);   // (completed long ago: acquire of the TMA writes)
        for (int c0 = e * 32; c0 < p.C; c0 += 128) {
          uint4 res[4];
          if (p.res_smem) {
            // the residual IS the X tile that fed fc1 (resid == x): read it back from shared memory (128B-swizzled K-blocks)
            // instead of a second, latency-exposed trip to global memory
            const uint8_t* xr = sX + (size_t)xb_o * x_bytes + (size_t)(c0 >> 6) * kXBlockBytes + (size_t)rloc * 128;
            const int ch0 = (c0 & 63) >> 3;
  #pragma unroll
            for (int i = 0; i < 4; ++i) res[i] = *reinterpret_cast<const uint4*>(xr + (((ch0 + i) ^ (rloc & 7)) << 4));
          } else if (rok) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.resid + (long long)row * p.C + c0);
  #pragma unroll
            for (int i = 0; i < 4; ++i) res[i] = rp[i];   // plain loads: resid may alias out
          }
          if (!waited) {
            mbar_wait(&ctrl->acc2_full, (uint32_t)oit & 1u, 62);
            tc_fence_after();
            waited = true;
          }
          uint32_t v[32];
          tmem_ld_x32(lane_addr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (rok) {
            uint4 o[4];
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b0 = WIDE ? __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + 8 * i)) : *reinterpret_cast<const float4*>(sB2 + c0 + 8 * i);
              const float4 b1 = WIDE ? __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + 8 * i + 4)) : *reinterpret_cast<const float4*>(sB2 + c0 + 8 * i + 4);
              const float2 r0 = unpack_bf16x2(res[i].x), r1 = unpack_bf16x2(res[i].y), r2_ = unpack_bf16x2(res[i].z),
                           r3 = unpack_bf16x2(res[i].w);
              o[i].x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) + b0.x + r0.x, __uint_as_float(v[8 * i + 1]) + b0.y + r0.y);
              o[i].y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) + b0.z + r1.x, __uint_as_float(v[8 * i + 3]) + b0.w + r1.y);
              o[i].z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) + b1.x + r2_.x, __uint_as_float(v[8 * i + 5]) + b1.y + r2_.y);
              o[i].w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) + b1.z + r3.x, __uint_as_float(v[8 * i + 7]) + b1.w + r3.y);
            }
            // lane = row stores: the LSU pays per cache line touched, so one full 32-byte sector per lane and instruction where the
            // rows are 32-byte aligned
            bf16* op = p.out + (long long)row * p.C + c0;
            if ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) {
              const uint32_t w0[8] = {o[0].x, o[0].y, o[0].z, o[0].w, o[1].x, o[1].y, o[1].z, o[1].w};
              const uint32_t w1[8] = {o[2].x, o[2].y, o[2].z, o[2].w, o[3].x, o[3].y, o[3].z, o[3].w};
              st_global_256(op, w0);
              st_global_256(op + 16, w1);
            } else {
  #pragma unroll
              for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(op)[i] = o[i];
            }
          }
        }
        if (!waited) {   // this warp owns no output columns (C < 128): still consume the phase
          mbar_wait(&ctrl->acc2_full, (uint32_t)oit & 1u, 62);
          tc_fence_after();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctrl->acc2_empty);
          if (p.res_smem) mbar_arrive(&ctrl->x_empty[xb_o]);   // the X tile may be refilled (tile oit + 2)
        }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int t = blockIdx.x + it * gridDim.x;
      const int row = t * BM + rloc;
      const bool rok = row < p.R;
      float own_r = 1.f, own_n = 0.f;
      if (p.ln_stats) {
        const float s1 = (nst[0].x + nst[1].x) + (nst[2].x + nst[3].x), s2 = (nst[0].y + nst[1].y) + (nst[2].y + nst[3].y);
        const float mu = s1 * p.ln_inv_k;
        const float var = fmaxf(fmaf(s2, p.ln_inv_k, -mu * mu), 0.f);
        own_r = rsqrtf(var + p.ln_eps);
        own_n = -own_r * mu;
#pragma unroll
        for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
        if (it + 1 < my_tiles) load_stats(t + gridDim.x);
      }
      const float2 r2 = make_float2(own_r, own_r), n2v = make_float2(own_n, own_n);
      for (int j = 0; j < J; ++j) {
        const uint32_t g = (uint32_t)(it * J + j);
        const uint32_t buf = g & 1u, use1 = g >> 1;
        const uint32_t hb = g & 1u, useh = g >> 1;
        // wide variant: the constants of this warp's 16 hidden columns come through L1 (uniform addresses: one broadcast per load),
        // requested before the accumulator wait
        float4 gc[WIDE ? HCW / 4 : 1], gb[WIDE ? HCW / 4 : 1];
        if constexpr (WIDE) {
#pragma unroll
          for (int i = 0; i < HCW / 4; ++i) {
            gb[i] = __ldg(reinterpret_cast<const float4*>(p.b1 + j * HC + e * HCW) + i);
            gc[i] = p.cs1 ? __ldg(reinterpret_cast<const float4*>(p.cs1 + j * HC + e * HCW) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        mbar_wait(&ctrl->acc1_full[buf], use1 & 1u, 60);
        tc_fence_after();
        uint32_t v[HCW];
        if constexpr (HCW == 32) tmem_ld_x32(lane_addr + kAcc1Col + buf * HC + (uint32_t)(e * HCW), v);
        else tmem_ld_x16(lane_addr + kAcc1Col + buf * HC + (uint32_t)(e * HCW), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acc1_empty[buf]);   // accumulator drained: the next fc1 may overwrite it
        const float4* cb = sCB + ((j * HC + e * HCW) >> 1);
        uint32_t pk[HCW / 2];
#pragma unroll
        for (int i = 0; i < HCW / 2; ++i) {
          float4 c;   // (cs, cs, b1, b1) of hidden columns 2i, 2i+1 of this warp's columns
          if constexpr (WIDE) {
            const float4 c4 = gc[i >> 1], b4 = gb[i >> 1];
            c = (i & 1) ? make_float4(c4.z, c4.w, b4.z, b4.w) : make_float4(c4.x, c4.y, b4.x, b4.y);
          } else {
            c = cb[i];   // broadcast read from shared memory
          }
          float2 a = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          a = ffma2(r2, a, ffma2(n2v, make_float2(c.x, c.y), make_float2(c.z, c.w)));
          a = gelu_fast2(a);
          pk[i] = pack_bf16x2(a.x, a.y);
        }
        mbar_wait(&ctrl->hid_empty[hb], (useh & 1u) ^ 1u, 61);   // the fc2 that last read this buffer has retired
        uint8_t* hrow = sH + (size_t)hb * kHidBytes + (size_t)((e * HCW) >> 6) * kXBlockBytes + (size_t)rloc * 128;
#pragma unroll
        for (int i = 0; i < HCW / 8; ++i) {
          const int ch = (((e * HCW) & 63) >> 3) + i;
          *reinterpret_cast<uint4*>(hrow + ((ch ^ (rloc & 7)) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->hid_full[hb]);
        if (p.nx == 2 && j == 0 && it > 0) output_epilogue(it - 1);   // deferred output epilogue of the previous tile
      }
      if (p.nx != 2) output_epilogue(it);
    }
    if (p.nx == 2 && my_tiles > 0) output_epilogue(my_tiles - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------
// Pair variant of the wide MLP (C = 384): two CTAs of a cluster (one TPC) work on 256 rows with cta_group::2 MMAs (M = 256).
// Each CTA keeps ITS 128 rows of X, of the hidden tile and of both accumulators, but only HALF of every weight box (the B operand
// of a pair MMA is split along N: rank r holds rows [r N/2, (r+1) N/2)), so the L2 -> SM weight traffic per row halves and the
// ring holds a whole 128-column hidden chunk per slot — the single-CTA wide variant is bound by the weight bytes it can keep in
// flight (~2k cycles of L2 latency under load against the 48 B/clk it needs).
//   TMEM             acc2 [0, 384), ONE fc1 accumulator [384, 512): hidden chunks of 128 columns.  The epilogue drains acc1 into
//                    registers right away, so fc1(g+1) only waits for that drain, and the pipe runs fc2(g-1) meanwhile:
//                    issue order fc1(0), { fc1(g+1), fc2(g) }.  Chunks of 64 with two accumulators were measured slower: every
//                    chunk costs ~2.7k cycles of hand-shakes across the pair regardless of its width.  Handing the hidden tile
//                    to fc2 in two 64-column halves was slower too (the single hidden buffer has to wait for fc2(g-1) anyway).
//   leader (rank 0)  warp 2 issues every MMA; its mbarriers collect the TMA bytes of both CTAs (x_full, w_full), the arrivals of
//                    its own epilogue warps and ONE forwarded arrival per event from the peer (acc1_empty, hid_full, acc2_empty)
//   both CTAs        producers load their halves (completing on the leader's barriers), epilogue warps work on their own rows
//                    and only ever arrive on their OWN CTA's barriers; tcgen05.commit multicasts to the barrier at the same
//                    offset in both CTAs (w_empty, x_empty, acc1_full, hid_empty, acc2_full)
//   peer (rank 1)    warps 2 and 3 forward "all 16 epilogue warps have arrived" to the leader's barrier: a cluster-scope
//                    release arrive costs the issuing warp ~700 cycles — twice per chunk in every epilogue warp when they
//                    signalled the leader directly (measured: the epilogue, not the tensor pipe, set the pace)
// Ring slot (48 KB per CTA): W1'(j) = one 3-D box [6 K-blocks][64 hidden rows][64]; W2(j) = four boxes [96 out rows][64 hidden]
// ([hidden K-block][output half]).
// ------------------------------------------------------------------------------------------------
constexpr int kPairHC = 128;                    // hidden columns per chunk
constexpr int kPairSlots = 2;
constexpr int kPairHidBytes = BM * kPairHC * 2;       // 32 KB (two K-blocks), single buffer
constexpr int kPairW1KbBytes = (kPairHC / 2) * BK * 2;   // 8 KB: this CTA's 64 hidden rows of one K-block
constexpr int kPairAcc1Col = 384;
constexpr int kPairConstBytes = 2 * kPairHC * 8;   // two buffers of (colsum, b1) per hidden column of a chunk
// shapes that depend on the channel count C (384: stage 3 of Base, stage 4 of Small; 320: stage 3 of Small, stage 4 of Tiny)
template <int C>
struct PairShape {
  static constexpr int kKb = C / BK;                        // K-blocks of X
  static constexpr int kXBytes = kKb * kXBlockBytes;        // 96 / 80 KB
  static constexpr int kN2 = C / 2;                         // fc2 in two halves of N = C / 2
  static constexpr int kW2PartBytes = (kN2 / 2) * BK * 2;   // this CTA's out rows of one half, one hidden K-block (12 / 10 KB)
  static constexpr int kSlotBytes = kKb * kPairW1KbBytes;   // 48 / 40 KB
  // [alignment pad <= 1023][ctrl 1 KB][X][H][ring][constants page 2 KB — only addressable when the pad is 0, see the epilogue]
  static constexpr int kSmemBytes = 1024 + kXBytes + kPairHidBytes + kPairSlots * kSlotBytes + kPairConstBytes;
  static_assert(C % BK == 0 && C <= kPairAcc1Col && kN2 % 16 == 0 && kN2 <= 256, "pair MLP channel count");
  static_assert(kSmemBytes <= kSmemLimit && kSmemBytes - kPairConstBytes + 1023 <= kSmemLimit, "pair MLP shared memory");
  static_assert(4 * kW2PartBytes == kSlotBytes && kW2PartBytes % 1024 == 0, "ring slot layout");
};
static bool pair_channels_ok(int C) { return C == 384 || C == 320; }

// Debug build (-DLMV_MLP_TRACE): cycle account of the pair kernel's MMA issuer (leader CTAs) and of one epilogue warp, read back
// with lmv_debug_mlp_trace().
// issuer slots:   0 wait x_full, 1 wait acc1_empty, 2 wait w_full (fc1), 3 issue fc1, 4 wait hid_full (+ acc2_empty), 5 wait w_full (fc2), 6 issue fc2, 7 total
// epilogue slots: 0 wait acc1_full, 1 ld + math, 2 wait hid_empty, 3 stores + arrive (+ loop), 4 output epilogue, 7 total
#ifdef LMV_MLP_TRACE
__device__ unsigned long long g_mlp_trace[148 * 8];
__device__ int g_mlp_debug;   // timing experiments (lmv_debug_mlp_flags; results become wrong): 1 skip the fc1 MMAs, 2 skip the fc2 MMAs
#define DBG(bit) (dbg_flags & (bit))
#define DBG_INIT const int dbg_flags = g_mlp_debug;
#define TR_INIT unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64(); const long long tr_t0 = tr_t;
#define TR(i) { const long long now_ = clock64(); tr[i] += (unsigned long long)(now_ - tr_t); tr_t = now_; }
#define TR_FLUSH { tr[7] = (unsigned long long)(clock64() - tr_t0); if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_mlp_trace[(size_t)blockIdx.x * 8 + i_] += tr[i_]; }
#else
#define TR_INIT
#define TR(i)
#define TR_FLUSH
#define DBG(bit) 0
#define DBG_INIT
#endif

struct CtrlPair {
  uint64_t x_full, x_empty;
  uint64_t w_full[kPairSlots], w_empty[kPairSlots];
  uint64_t acc1_full, acc1_empty;
  uint64_t hid_full, hid_empty;
  uint64_t acc2_full, acc2_empty;
  uint32_t tmem_base;
};

template <int kPairC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_pair_tcgen05(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  using Shape = PairShape<kPairC>;
  constexpr int kPairKb = Shape::kKb, kPairXBytes = Shape::kXBytes, kPairN2 = Shape::kN2, kPairW2PartBytes = Shape::kW2PartBytes,
                kPairSlotBytes = Shape::kSlotBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  CtrlPair* ctrl = reinterpret_cast<CtrlPair*>(smem);
  uint8_t* sX = smem + 1024;
  uint8_t* sH = sX + kPairXBytes;
  uint8_t* sW = sH + kPairHidBytes;
  // The last 2 KB of the allocation hold the epilogue constants of the current and the next hidden chunk — but only when the
  // dynamic shared memory happens to start 1024-aligned (pad == 0: it does on every driver seen so far); otherwise those bytes are
  // the alignment pad and the epilogue broadcasts its constants with shuffles instead.
  float2* sConst = reinterpret_cast<float2*>(sW + kPairSlots * kPairSlotBytes);
  const bool const_page = pad == 0;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int J = p.chunks;
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int my_tiles = (pair < p.tiles) ? (p.tiles - 1 - pair) / npairs + 1 : 0;   // p.tiles counts 256-row pair tiles
  const int G = my_tiles * J;                                                      // hidden chunks of this pair
  auto row0_of = [&](int it) { return ((pair + it * npairs) * 2 + (int)rank) * BM; };
  auto leader = [&](uint64_t* bar) { return mapa_u32(smem_u32(bar), 0u); };   // the leader's copy of a barrier

  if (threadIdx.x == 0) {
    const uint32_t nepi = kEpiWarps + (rank == 0 ? 1 : 0);   // own epilogue warps (+ the peer's forwarder on the leader)
    mbar_init(&ctrl->x_full, 1);
    mbar_init(&ctrl->x_empty, 1);
    for (int i = 0; i < kPairSlots; ++i) {
      mbar_init(&ctrl->w_full[i], 1);
      mbar_init(&ctrl->w_empty[i], 1);
    }
    mbar_init(&ctrl->acc1_full, 1);
    mbar_init(&ctrl->acc1_empty, nepi);
    mbar_init(&ctrl->hid_full, nepi);
    mbar_init(&ctrl->hid_empty, 1);
    mbar_init(&ctrl->acc2_full, 1);
    mbar_init(&ctrl->acc2_empty, nepi);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) {
    tmem_alloc_2sm(&ctrl->tmem_base, 512);
    tmem_relinquish_2sm();
  }
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer A: this CTA's 128 rows of X ----------------
    if (lane == 0) {
      const uint32_t bar = leader(&ctrl->x_full);
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&ctrl->x_empty, ((uint32_t)it & 1u) ^ 1u, 40);
        if (rank == 0) mbar_expect_tx(&ctrl->x_full, 2u * kPairXBytes);
        for (int kb = 0; kb < kPairKb; ++kb) tma_load_2d_2sm(sX + (size_t)kb * kXBlockBytes, &tmX, bar, kb * BK, row0_of(it));
      }
    }
  } else if (warp == 1) {
    // ---------------- TMA producer B: this CTA's halves of the weight boxes, in the issuer's consumption order ----------------
    if (lane == 0) {
      int slot = 0;
      uint32_t wphase = 0;
      auto next = [&]() { if (++slot == kPairSlots) { slot = 0; wphase ^= 1u; } };
      auto put_w1 = [&](int j) {
        mbar_wait(&ctrl->w_empty[slot], wphase ^ 1u, 41);
        if (rank == 0) mbar_expect_tx(&ctrl->w_full[slot], 2u * kPairSlotBytes);
        tma_load_3d_2sm(sW + (size_t)slot * kPairSlotBytes, &tmW1, leader(&ctrl->w_full[slot]), 0, j * kPairHC + (int)rank * (kPairHC / 2), 0);
        next();
      };
      auto put_w2 = [&](int j) {
        mbar_wait(&ctrl->w_empty[slot], wphase ^ 1u, 42);
        if (rank == 0) mbar_expect_tx(&ctrl->w_full[slot], 2u * kPairSlotBytes);
        const uint32_t bar = leader(&ctrl->w_full[slot]);
        for (int kb = 0; kb < kPairHC / BK; ++kb)
          for (int q = 0; q < 2; ++q)
            tma_load_2d_2sm(sW + (size_t)slot * kPairSlotBytes + (size_t)(kb * 2 + q) * kPairW2PartBytes, &tmW2, bar, j * kPairHC + kb * BK,
                            q * kPairN2 + (int)rank * (kPairN2 / 2));
        next();
      };
      for (int it = 0; it < my_tiles; ++it) {   // W1'(0), W1'(1), W2(0), W1'(2), W2(1), ..., W2(J-1)
        put_w1(0);
        for (int j = 1; j < J; ++j) { put_w1(j); put_w2(j - 1); }
        put_w2(J - 1);
      }
    }
  } else if (warp == 2) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader CTA only): whole warp in uniform control flow, one elected lane issues ----------------
      const uint32_t idesc1 = make_idesc_bf16(2 * BM, kPairHC);
      const uint32_t idesc2 = make_idesc_bf16(2 * BM, kPairN2);
      int s = 0;
      uint32_t ph = 0;
      const uint32_t xaddr = smem_u32(sX), haddr = smem_u32(sH);
      auto wait_warp = [&](uint64_t* bar, uint32_t parity) { mbar_wait_cluster(bar, parity); __syncwarp(); };
      TR_INIT
      DBG_INIT
      auto fc1 = [&](int g) {
        const int j = g % J;
        TR(6)
        if (j == 0) wait_warp(&ctrl->x_full, (uint32_t)(g / J) & 1u);   // both CTAs' X tiles have landed
        TR(0)
        wait_warp(&ctrl->acc1_empty, ((uint32_t)g & 1u) ^ 1u);          // both CTAs' epilogues have drained acc1(g-1)
        TR(1)
        wait_warp(&ctrl->w_full[s], ph);
        TR(2)
        tc_fence_after();
        const uint32_t d = tmem_base + kPairAcc1Col;
        const uint32_t wbase = smem_u32(sW + (size_t)s * kPairSlotBytes);
#pragma unroll
        for (int kb = 0; kb < kPairKb; ++kb) {
          const uint64_t da = make_kmajor_desc<128>(xaddr + (uint32_t)kb * kXBlockBytes);
          const uint64_t db = make_kmajor_desc<128>(wbase + (uint32_t)kb * kPairW1KbBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            if (!DBG(1)) umma_bf16_ss_warp_2sm(d, da + 2ull * k, db + 2ull * k, idesc1, (uint32_t)((kb | k) != 0));
        }
        umma_commit_warp_2sm(&ctrl->w_empty[s]);
        if (++s == kPairSlots) { s = 0; ph ^= 1u; }
        umma_commit_warp_2sm(&ctrl->acc1_full);
        if (j == J - 1) umma_commit_warp_2sm(&ctrl->x_empty);   // the X tiles are free once the last fc1 of the tile retires
      };
      auto fc2 = [&](int g) {
        const int j = g % J;
        TR(3)
        wait_warp(&ctrl->hid_full, (uint32_t)g & 1u);
        if (j == 0) wait_warp(&ctrl->acc2_empty, ((uint32_t)(g / J) & 1u) ^ 1u);
        TR(4)
        wait_warp(&ctrl->w_full[s], ph);
        TR(5)
        tc_fence_after();
        const uint32_t wbase = smem_u32(sW + (size_t)s * kPairSlotBytes);
#pragma unroll
        for (int kb = 0; kb < kPairHC / BK; ++kb) {
          const uint64_t da = make_kmajor_desc<128>(haddr + (uint32_t)kb * kXBlockBytes);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint64_t db = make_kmajor_desc<128>(wbase + (uint32_t)(kb * 2 + q) * kPairW2PartBytes);
            const uint32_t d = tmem_base + (uint32_t)(q * kPairN2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (!DBG(2)) umma_bf16_ss_warp_2sm(d, da + 2ull * k, db + 2ull * k, idesc2, (uint32_t)((j | kb | k) != 0));
          }
        }
        umma_commit_warp_2sm(&ctrl->w_empty[s]);
        if (++s == kPairSlots) { s = 0; ph ^= 1u; }
        umma_commit_warp_2sm(&ctrl->hid_empty);
        if (j == J - 1) umma_commit_warp_2sm(&ctrl->acc2_full);
      };
      // fc1 runs one chunk ahead of fc2 inside a tile (the ring's order); a single X buffer: never run ahead into the next tile
      for (int g = 0; g < G; ++g) {
        if (g % J == 0) fc1(g);
        if ((g + 1) % J != 0) fc1(g + 1);
        fc2(g);
      }
      TR(6)
      TR_FLUSH
    } else {
      // peer CTA: forward hid_full (all 16 local epilogue warps have written their part of H_g) to the leader
      const uint32_t dst = leader(&ctrl->hid_full);
      for (int g = 0; g < G; ++g) {
        mbar_wait(&ctrl->hid_full, (uint32_t)g & 1u, 70);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(dst);
      }
    }
  } else if (warp == 3) {
    if (rank == 1) {
      // peer CTA: forward acc1_empty (per chunk) and acc2_empty (per tile) to the leader, in the order the epilogue warps arrive:
      // the output epilogue of a tile is deferred behind the first chunk of the next one
      const uint32_t dst1 = leader(&ctrl->acc1_empty), dst2 = leader(&ctrl->acc2_empty);
      auto forward = [&](uint64_t* bar, uint32_t parity, uint32_t dst) {
        mbar_wait(bar, parity, 71);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(dst);
      };
      for (int it = 0; it < my_tiles; ++it)
        for (int j = 0; j < J; ++j) {
          forward(&ctrl->acc1_empty, (uint32_t)(it * J + j) & 1u, dst1);
          if (j == 0 && it > 0) forward(&ctrl->acc2_empty, (uint32_t)(it - 1) & 1u, dst2);
        }
      if (my_tiles > 0) forward(&ctrl->acc2_empty, (uint32_t)(my_tiles - 1) & 1u, dst2);
    }
  } else if (warp >= kFirstEpiWarp) {
    // ---------------- epilogue warps (both CTAs, on their own 128 rows) ----------------
    const int q = warp & 3;                        // TMEM lane quarter
    const int e = (warp - kFirstEpiWarp) >> 2;     // 0..3: this warp's 32 of the chunk's 128 hidden columns
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int rloc = q * 32 + lane;                // row inside this CTA's tile == TMEM lane
    float2 nst[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
    auto load_stats = [&](int it) {
      const int r = row0_of(it) + rloc;
      if (r < p.R) {
        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + (long long)r * p.ln_parts;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < p.ln_parts) nst[k] = __ldg(st + k);
      }
    };
    if (p.ln_stats && my_tiles > 0) load_stats(0);
    // Epilogue constants of this warp's 32 hidden columns of a chunk, one register pair per lane: (colsum(W1'), b1) of column
    // (lane); one coalesced load per chunk, requested a whole chunk ahead.  Every warp writes them to the constants page (buffer
    // g & 1; the four warps that share the columns write identical bytes, so nobody waits for anybody) and reads them back as
    // broadcast 16-byte loads; without the page they are broadcast with shuffles, whose throughput (one warp per clock and SM)
    // then limits the chunk.  Buffer reuse is safe: a warp starts chunk g+2 only after acc1_full(g+2), i.e. after fc1(g+2) was
    // issued, i.e. after every warp drained acc1(g+1) — which it does after finishing the math of chunk g.
    // (warp-uniform 16-byte loads through L1, even prefetched, were measured 20% slower than the shuffles.)
    auto load_consts = [&](int j) {
      const int col = j * kPairHC + e * 32 + lane;
      return make_float2(p.cs1 ? __ldg(p.cs1 + col) : 0.f, __ldg(p.b1 + col));
    };
    float2 cnext = my_tiles > 0 ? load_consts(0) : make_float2(0.f, 0.f);
    TR_INIT
    // ---- output epilogue of tile iteration `oit`: acc2 + b2 + residual -> out.  It runs AFTER the first hidden chunk of the next tile
    // was handed to fc2, so the wait for the last fc2 of the tile and the output stores hide behind fc1 of the next tile.
    auto output_epilogue = [&](int oit) {
      const int row = row0_of(oit) + rloc;
      const bool rok = row < p.R;
      TR(3)
      // this warp's output columns: c0 = 32 e + 128 k.  b2 sits one column per lane (broadcast by shuffles), the residual of chunk
      // k+1 is requested before chunk k is processed: the three dependent trips to L2 of a straightforward loop were the largest
      // single item of this kernel's per-tile time, and while the epilogue warps are here nobody drains the next fc1 accumulator.
      constexpr int kChunks = (kPairC + 127) / 128;
      float b2r[kChunks];
#pragma unroll
      for (int k = 0; k < kChunks; ++k) {
        const int c0 = e * 32 + k * 128;
        b2r[k] = c0 < kPairC ? __ldg(p.b2 + c0 + lane) : 0.f;
      }
      // lane = row: every lane touches its own cache lines, and the LSU pays per line — 256-bit accesses (a full sector per lane)
      // where the rows are 32-byte aligned, 128-bit ones otherwise
      const bool wide_ls = ((reinterpret_cast<uintptr_t>(p.resid) | reinterpret_cast<uintptr_t>(p.out)) & 31) == 0;
      auto load_res = [&](int c0, uint32_t (&res)[16]) {
        if (rok && c0 < kPairC) {
          const bf16* rp = p.resid + (long long)row * kPairC + c0;   // plain loads: resid may alias out
          if (wide_ls) {
            ld_global_256(rp, reinterpret_cast<uint32_t (&)[8]>(res[0]));
            ld_global_256(rp + 16, reinterpret_cast<uint32_t (&)[8]>(res[8]));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 u = reinterpret_cast<const uint4*>(rp)[i];
              res[4 * i] = u.x; res[4 * i + 1] = u.y; res[4 * i + 2] = u.z; res[4 * i + 3] = u.w;
            }
          }
        }
      };
      uint32_t res[2][16];
      load_res(e * 32, res[0]);
      mbar_wait(&ctrl->acc2_full, (uint32_t)oit & 1u, 62);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kChunks; ++k) {
        const int c0 = e * 32 + k * 128;
        if (c0 >= kPairC) break;   // (warp-uniform)
        if (k + 1 < kChunks) load_res(c0 + 128, res[(k + 1) & 1]);
        uint32_t v[32];
        tmem_ld_x32(lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 rr = unpack_bf16x2(res[k & 1][i]);
          const float b0 = __shfl_sync(0xffffffffu, b2r[k], 2 * i), b1 = __shfl_sync(0xffffffffu, b2r[k], 2 * i + 1);
          o[i] = pack_bf16x2(__uint_as_float(v[2 * i]) + b0 + rr.x, __uint_as_float(v[2 * i + 1]) + b1 + rr.y);
        }
        if (rok) {
          bf16* op = p.out + (long long)row * kPairC + c0;
          if (wide_ls) {
            st_global_256(op, reinterpret_cast<const uint32_t (&)[8]>(o[0]));
            st_global_256(op + 16, reinterpret_cast<const uint32_t (&)[8]>(o[8]));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(op)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->acc2_empty);
      TR(4)
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int row = row0_of(it) + rloc;
      const bool rok = row < p.R;
      float own_r = 1.f, own_n = 0.f;
      if (p.ln_stats) {
        const float s1 = (nst[0].x + nst[1].x) + (nst[2].x + nst[3].x), s2 = (nst[0].y + nst[1].y) + (nst[2].y + nst[3].y);
        const float mu = s1 * p.ln_inv_k;
        const float var = fmaxf(fmaf(s2, p.ln_inv_k, -mu * mu), 0.f);
        own_r = rsqrtf(var + p.ln_eps);
        own_n = -own_r * mu;
#pragma unroll
        for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
        if (it + 1 < my_tiles) load_stats(it + 1);
      }
      const float2 r2 = make_float2(own_r, own_r), n2v = make_float2(own_n, own_n);
      for (int j = 0; j < J; ++j) {
        const uint32_t g = (uint32_t)(it * J + j);
        const float2 ccur = cnext;
        cnext = load_consts(j + 1 < J ? j + 1 : 0);
        TR(3)
        mbar_wait(&ctrl->acc1_full, g & 1u, 60);
        TR(0)
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_x32(lane_addr + kPairAcc1Col + (uint32_t)(e * 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acc1_empty);   // accumulator drained (own CTA's barrier; the peer's is forwarded)
        uint32_t pk[16];
        if (const_page) {
          float2* cpage = sConst + (g & 1u) * kPairHC + e * 32;
          cpage[lane] = ccur;
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 c = *reinterpret_cast<const float4*>(cpage + 2 * i);   // (cs, b1) of columns 2i, 2i+1 — broadcast read
            float2 a = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            a = ffma2(r2, a, ffma2(n2v, make_float2(c.x, c.z), make_float2(c.y, c.w)));
            a = gelu_fast2(a);
            pk[i] = pack_bf16x2(a.x, a.y);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 cs = make_float2(__shfl_sync(0xffffffffu, ccur.x, 2 * i), __shfl_sync(0xffffffffu, ccur.x, 2 * i + 1));
            const float2 bb = make_float2(__shfl_sync(0xffffffffu, ccur.y, 2 * i), __shfl_sync(0xffffffffu, ccur.y, 2 * i + 1));
            float2 a = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            a = ffma2(r2, a, ffma2(n2v, cs, bb));
            a = gelu_fast2(a);
            pk[i] = pack_bf16x2(a.x, a.y);
          }
        }
        TR(1)
        mbar_wait(&ctrl->hid_empty, (g & 1u) ^ 1u, 61);   // fc2(g-1), the last reader of the hidden tile, has retired
        TR(2)
        uint8_t* hrow = sH + (size_t)(e >> 1) * kXBlockBytes + (size_t)rloc * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = (e & 1) * 4 + i;
          *reinterpret_cast<uint4*>(hrow + ((ch ^ (rloc & 7)) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->hid_full);
        if (j == 0 && it > 0) output_epilogue(it - 1);   // deferred output epilogue of the previous tile
      }
    }
    if (my_tiles > 0) output_epilogue(my_tiles - 1);
#ifdef LMV_MLP_TRACE
    if (warp == kFirstEpiWarp && rank == 0) { tr[7] = (unsigned long long)(clock64() - tr_t0); if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_mlp_trace[(size_t)(blockIdx.x + 1) * 8 + i_] += tr[i_]; }
#endif
  }
  tc_fence_before();
  cluster_sync_all();   // the peer may still be signalling this CTA's barriers / reading its shared memory
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

PerDeviceOnce g_attr_once;

}  // namespace

static int mlp_hc(int C) { return C <= 256 ? 128 : 64; }

bool mlp_fused_supported(int C, int Hd) {
  // acc2 [128 x C] + two acc1 buffers must fit the 512 TMEM columns; the X tile, two hidden buffers and both weight rings
  // must fit 227 KB of shared memory (C = 384: 96 KB X tile, 14 KB constants at Hd = 1536)
  const bool narrow = C >= 32 && C <= 192 && C % 32 == 0;
  const bool wide_ok = C == 384 && Hd <= 1536;                   // single-CTA wide kernel (constants page sized for Hd <= 1536)
  const char* pair_env = getenv("LMV_MLP_PAIR");
  const bool pair_ok = pair_channels_ok(C) && !(pair_env && atoi(pair_env) == 0);   // CTA-pair kernel
  return (narrow || wide_ok || pair_ok) && Hd % 128 == 0 && Hd >= 128 && Hd <= 4096;
}

int mlp_fused_prepare(const MlpArgs& a, MlpOp* op) {
  LMV_REQUIRE(a.x && a.W1 && a.W2 && a.b1 && a.b2 && a.out, "mlp_fused: null pointer");
  LMV_REQUIRE(a.R > 0, "mlp_fused: empty problem");
  if (!mlp_fused_supported(a.C, a.Hd))
    return fail(LMV_ERR_UNSUPPORTED, "mlp_fused: needs C % 32 == 0 with 32 <= C <= 192 (or C == 384 with hidden <= 1536), hidden % 128 == 0, hidden <= 4096");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  LMV_REQUIRE(al16(a.x) && al16(a.W1) && al16(a.W2) && al16(a.b1) && al16(a.b2) && al16(a.out) && al16(a.resid ? a.resid : a.x),
              "mlp_fused: pointers must be 16-byte aligned");
  LMV_REQUIRE((a.ln_stats == nullptr) == (a.cs1 == nullptr), "mlp_fused: ln_stats and colsum go together");
  LMV_REQUIRE(a.cs1 == nullptr || al16(a.cs1), "mlp_fused: colsum must be 16-byte aligned");
  LMV_REQUIRE(a.ln_parts <= 4, "mlp_fused: at most 4 LayerNorm statistics partials per row");
  MlpParams& p = op->p;
  p.R = a.R; p.C = a.C; p.Hd = a.Hd;
  p.tiles = (a.R + BM - 1) / BM;
  const int HC = mlp_hc(a.C);
  const int kHidBytes = BM * HC * 2, kW1KbBytes = HC * BK * 2;
  p.hc = HC;
  // the wide shape runs on the CTA-pair kernel unless LMV_MLP_PAIR=0 (read per call: schedules are built once, tests switch it)
  const char* pair_env = getenv("LMV_MLP_PAIR");
  p.pair = (HC == 64 && pair_channels_ok(a.C) && !(pair_env && atoi(pair_env) == 0)) ? 1 : 0;
  if (p.pair) {
    // CTA-pair kernel: 256 rows per cluster, each CTA loads half of every weight box
    p.tiles = (a.R + 2 * BM - 1) / (2 * BM);
    const int kb = a.C / BK, n2 = a.C / 2;
    p.kb1 = kb; p.chunks = a.Hd / kPairHC; p.nparts2 = 2; p.n2 = n2; p.slot2_bytes = 0; p.const_bytes = 0;
    p.nx = 1; p.nh = 1; p.n1slots = kPairSlots; p.n2slots = 0; p.w1_kpb = kb; p.w1_boxes = 1; p.res_smem = 0;
    op->smem_bytes = a.C == 384 ? PairShape<384>::kSmemBytes : PairShape<320>::kSmemBytes;
    p.b1 = a.b1; p.cs1 = a.cs1; p.b2 = a.b2;
    p.ln_stats = a.ln_stats; p.ln_parts = a.ln_parts > 0 ? a.ln_parts : 1; p.ln_eps = a.ln_eps; p.ln_inv_k = 1.0f / (float)a.C;
    p.resid = a.resid ? a.resid : a.x;
    p.out = a.out;
    op->grid = 2 * std::min(p.tiles, device_sm_count() / 2);
    int rc;
    {
      uint64_t dims[2] = {(uint64_t)a.C, (uint64_t)a.R};
      uint64_t strides[1] = {(uint64_t)a.C * 2};
      uint32_t box[2] = {BK, BM};
      if ((rc = encode_tmap_bf16(&op->tmX, a.x, 2, dims, strides, box, 128))) return rc;
    }
    {   // (k inside a K-block, hidden row, K-block): this CTA's 64 hidden rows of all six K-blocks in one box
      uint64_t dims[3] = {(uint64_t)BK, (uint64_t)a.Hd, (uint64_t)kb};
      uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)BK * 2};
      uint32_t box[3] = {BK, (uint32_t)(kPairHC / 2), (uint32_t)kb};
      if ((rc = encode_tmap_bf16(&op->tmW1, a.W1, 3, dims, strides, box, 128))) return rc;
    }
    {
      uint64_t dims[2] = {(uint64_t)a.Hd, (uint64_t)a.C};
      uint64_t strides[1] = {(uint64_t)a.Hd * 2};
      uint32_t box[2] = {BK, (uint32_t)(n2 / 2)};
      if ((rc = encode_tmap_bf16(&op->tmW2, a.W2, 2, dims, strides, box, 128))) return rc;
    }
    return LMV_OK;
  }
  p.kb1 = (a.C + BK - 1) / BK;
  p.chunks = a.Hd / HC;
  p.nparts2 = a.C <= 256 ? 1 : 2;
  p.n2 = a.C / p.nparts2;
  p.slot2_bytes = p.n2 * BK * 2;
  p.const_bytes = HC == 128 ? ((a.Hd * 8 + a.C * 4 + 1023) / 1024) * 1024 : 0;   // wide: constants come through L1
  const int x_bytes = p.kb1 * kXBlockBytes;
  const int budget = kSmemLimit - 2048 - p.const_bytes;     // ctrl + alignment slack
  p.nh = 2;
  if (HC == 128) {
    p.w1_kpb = 1;
    p.w1_boxes = p.kb1;
    // X double-buffered when that still leaves one chunk of lookahead in each ring (kb1 W1' boxes, 2 W2 boxes)
    const int min_rings = (p.kb1 + 1) * kW1KbBytes + 3 * p.slot2_bytes;
    p.nx = 2;
    if (budget - p.nx * x_bytes - p.nh * kHidBytes < min_rings) p.nx = 1;
    const int rest = budget - p.nx * x_bytes - p.nh * kHidBytes;
    p.n2slots = std::min(4, std::max(2, (rest - (p.kb1 + 1) * kW1KbBytes) / p.slot2_bytes));
    p.n1slots = std::min(kMaxSlots, (rest - p.n2slots * p.slot2_bytes) / kW1KbBytes);
  } else {
    // wide variant: one X buffer and ONE ring of 24 KB slots for both weights (W1' boxes of 3 K-blocks, W2 halves [C/2 x 64])
    LMV_REQUIRE(p.kb1 % 3 == 0 && p.n2 * BK * 2 == 3 * kW1KbBytes, "mlp_fused: wide variant needs C == 384");
    p.nx = 1;
    p.w1_kpb = 3;
    p.w1_boxes = p.kb1 / 3;
    p.n2slots = 0;
    p.n1slots = std::min(kMaxSlots, (budget + 1024 - p.nx * x_bytes - p.nh * kHidBytes) / (3 * kW1KbBytes));   // (no constants page: only the alignment slack is reserved)
    LMV_REQUIRE(p.n1slots >= 3, "mlp_fused: shared memory budget (weight ring)");
  }
  LMV_REQUIRE(p.n1slots >= 2 && (HC == 64 || p.n2slots >= 2), "mlp_fused: shared memory budget (weight rings)");
  op->smem_bytes = (HC == 128 ? 2048 : 1024 + 1023) + p.const_bytes + p.nx * x_bytes + p.nh * kHidBytes + p.n1slots * p.w1_kpb * kW1KbBytes + p.n2slots * p.slot2_bytes;
  LMV_REQUIRE(op->smem_bytes <= kSmemLimit, "mlp_fused: shared memory budget");
  p.b1 = a.b1; p.cs1 = a.cs1; p.b2 = a.b2;
  p.ln_stats = a.ln_stats; p.ln_parts = a.ln_parts > 0 ? a.ln_parts : 1; p.ln_eps = a.ln_eps; p.ln_inv_k = 1.0f / (float)a.C;
  p.resid = a.resid ? a.resid : a.x;
  p.out = a.out;
  p.res_smem = (p.nx == 2 && p.resid == a.x) ? 1 : 0;
  op->grid = std::min(p.tiles, device_sm_count());
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)a.C, (uint64_t)a.R};
    uint64_t strides[1] = {(uint64_t)a.C * 2};
    uint32_t box[2] = {BK, BM};
    if ((rc = encode_tmap_bf16(&op->tmX, a.x, 2, dims, strides, box, 128))) return rc;
  }
  {
    if (HC == 128) {
      uint64_t dims[2] = {(uint64_t)a.C, (uint64_t)a.Hd};
      uint64_t strides[1] = {(uint64_t)a.C * 2};
      uint32_t box[2] = {BK, (uint32_t)HC};
      if ((rc = encode_tmap_bf16(&op->tmW1, a.W1, 2, dims, strides, box, 128))) return rc;
    } else {   // (k inside a K-block, hidden row, K-block): one box = w1_kpb stacked [HC x 64] K-blocks
      uint64_t dims[3] = {(uint64_t)BK, (uint64_t)a.Hd, (uint64_t)p.kb1};
      uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)BK * 2};
      uint32_t box[3] = {BK, (uint32_t)HC, (uint32_t)p.w1_kpb};
      if ((rc = encode_tmap_bf16(&op->tmW1, a.W1, 3, dims, strides, box, 128))) return rc;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)a.Hd, (uint64_t)a.C};
    uint64_t strides[1] = {(uint64_t)a.Hd * 2};
    uint32_t box[2] = {BK, (uint32_t)p.n2};
    if ((rc = encode_tmap_bf16(&op->tmW2, a.W2, 2, dims, strides, box, 128))) return rc;
  }
  return LMV_OK;
}

int mlp_fused_run(const MlpOp& op, cudaStream_t stream) {
  LMV_CUDA_OK(g_attr_once.run([] {
    const cudaError_t e0 = cudaFuncSetAttribute(mlp_fused_tcgen05<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    const cudaError_t e1 = cudaFuncSetAttribute(mlp_fused_tcgen05<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    const cudaError_t e2 = cudaFuncSetAttribute(mlp_pair_tcgen05<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    const cudaError_t e3 = cudaFuncSetAttribute(mlp_pair_tcgen05<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    return e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
  }));
  auto fn = op.p.pair ? (op.p.C == 384 ? mlp_pair_tcgen05<384> : mlp_pair_tcgen05<320>)
                      : (op.p.hc == 128 ? mlp_fused_tcgen05<128> : mlp_fused_tcgen05<64>);
  LMV_CUDA_OK(launch_kernel(fn, dim3(op.grid), dim3(kThreads), (size_t)(op.smem_bytes), stream, op.tmX, op.tmW1, op.tmW2, op.p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv

#ifdef LMV_MLP_TRACE
extern "C" int lmv_debug_mlp_flags(int flags) { return cudaMemcpyToSymbol(lmv::g_mlp_debug, &flags, sizeof(int)) != cudaSuccess; }
extern "C" int lmv_debug_mlp_trace(unsigned long long* host, int n) {
  static unsigned long long zero[148 * 8];
  if (n > 148 * 8) n = 148 * 8;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(host, lmv::g_mlp_trace, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(lmv::g_mlp_trace, zero, sizeof(zero)) != cudaSuccess;
}
#endif
