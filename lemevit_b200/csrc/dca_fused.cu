// Image-token side of the fused cross-attention blocks — the DualCrossAttention of a 'D' block (models/lemevit.py:252-302 inside
// forward_with_xc :542-582: qkv1 projection, both scaled QK^T, both softmaxes, both AV products and proj_x + residual) and the
// image side of the CrossAttention of a 'C' block (:477-486) — as ONE persistent tcgen05 kernel over 128-token tiles.
//
// The image-side projections are absorbed into per-image [R = heads*16, C] operands built by meta_pre (kernels.h, meta_branch.cu),
// so per tile of 128 image tokens X (= x + dwconv(x), raw bf16; LayerNorm applied from the row statistics after the contraction):
//
//   S  [128 tok, R]   = X Kt^T        tcgen05.mma M=128 N=R   K=C     x-branch scores of every head against the 16 meta keys
//   Sc [R, 128 tok]   = Qt X^T        tcgen05.mma M=128 N=128 K=C     c-branch scores of the 16 meta queries of every head
//   P  = softmax over the 16 keys of each head (thread = token)       -> bf16 smem tile (A operand)
//   P' = exp2(Sc - m) r_n             (thread = (head, query) row)    -> bf16 smem tile (A operand); running (m, l, sum p' mu)
//   dx [128 tok, C]   = P Vt          tcgen05.mma M=128 N=C K=R       = proj_x(attention output) - b_px
//   Z  [R, C]        += P' X          tcgen05.mma M=128 N=C K=128     X is the B operand AS LOADED (MN-major), accumulated in TMEM
//                                                                      across the tiles of a segment with lazy rescaling
//   x_out = X + dx + b_px  (+ per-row (sum, sum^2) of the stored rows: the LayerNorm statistics the fused MLP kernel consumes)
//
// Neither qkv1 [N, 3C] nor the attention output ever exists; x is read once and written once per block, the c-branch leaves the
// kernel as one (m, l, t, Z[C]) partial per (segment, head, query) that meta_post merges.
//
// Roles (320 threads):  warp 0 TMA producer (X tiles, per-image operands) | warp 1 MMA issuer | warps 2-5 x-group (softmax over
// the meta keys + output epilogue, thread = token = TMEM lane) | warps 6-9 c-group (softmax over the image tokens, thread =
// (head, query) row = TMEM lane; rows duplicated at lanes 64.. when R <= 64 so that every thread has half a tile of columns).
// TMEM columns: [0, C) S then dx | [C, C+128) Sc | [C+128, 2C+128) Z.
#include <algorithm>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kThreads = 320;
constexpr int kXGroup0 = 2, kCGroup0 = 6;     // first warp of the x-group / c-group
constexpr int kXBlockBytes = 128 * 128;        // [128 rows x 64 ch] bf16, 128B swizzle
constexpr int kPBlockBytes = 128 * 64;         // [128 rows x 32 (h,m)] bf16, 64B swizzle
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHeaderBytes = 7 * 1024;    // barriers, per-image constants, proj_x bias, per-tile (r, mu) of the tokens
constexpr float kRescaleThreshold = 8.f;       // log2 domain: the running maximum may lag the true one by a factor <= 256

struct Ctrl {
  uint64_t x_full[3], x_empty[3];
  uint64_t op_full, op_empty;
  uint64_t s_full, p_full[2], dx_full, dx_free;
  uint64_t sc_full[2], pc_full[2], z_done;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major operand in the 128B-swizzled [rows x 64 ch] blocks a TMA box {64, rows} writes: 8 x 128 B atoms along K (rows),
// SBO = 1024 B between 8-row groups, LBO = distance between 64-element chunks along MN (= one block)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// Debug build (-DLMV_DCA_TRACE): cycle accounting per role (lane 0 of one warp per role), read back with lmv_debug_dca_trace().
// slots: producer {0 wait op_empty, 1 wait x_empty, 7 total} | MMA {0 issue, 1 wait/poll, 7 total}
//        x-group {0 stats + consts, 1 wait s_full, 2 softmax, 3 wait dx_full, 4 epilogue, 7 total}
//        c-group {0 stats staging + sync, 1 wait sc_full, 2 pass 1, 3 wait z_done, 4 rescale + pass 2, 5 flush, 7 total}
#ifdef LMV_DCA_TRACE
__device__ unsigned long long g_dca_trace[148 * 4 * 8];
#define TR_INIT unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64(); const long long tr_t0 = tr_t;
#define TR(i) { const long long now_ = clock64(); tr[i] += (unsigned long long)(now_ - tr_t); tr_t = now_; }
#define TR_FLUSH(role) { tr[7] = (unsigned long long)(clock64() - tr_t0); if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_dca_trace[((size_t)blockIdx.x * 4 + (role)) * 8 + i_] += tr[i_]; }
#else
#define TR_INIT
#define TR(i)
#define TR_FLUSH(role)
#endif

// position of a role in this CTA's (segment, tile) sequence
struct Cursor {
  int s, t, it;     // flattened segment index, tile inside the segment, tiles done so far
  int b, sg;        // image and segment inside the image (s = b * segs + sg)
  int t0, nt;       // first tile of the segment inside the image, tiles of the segment
  int xb, xu;       // X-tile ring: buffer of this tile and how often that buffer was used before
};

// PIPE: the tensor-memory budget allows a second score buffer for the c-branch and separate S / dx regions (C <= 96, or no x-branch):
// the MMA warp becomes an event loop that issues whichever of S / Sc / dx / Z of the next tiles has its inputs ready, the x-group
// runs the softmax of tile i + 1 before the output epilogue of tile i, and the c-group never waits for its scores.
template <bool PIPE>
__global__ void __launch_bounds__(kThreads, 1)
dca_x_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmKt, const __grid_constant__ CUtensorMap tmQt,
             const __grid_constant__ CUtensorMap tmVt, const DcaXParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  float* sKc = reinterpret_cast<float*>(smem + 256);            // [2][R]   (sum Kt, kappa) of the current image          (x-group)
  float* sBias = sKc + 2 * 128;                                  // [C]      proj_x bias
  float4* sStat = reinterpret_cast<float4*>(sBias + 256);        // [2][128] (r_n, -r_n mu_n, mu_n, -) of the tile's tokens, two buffers (c-group)
  uint8_t* sX = smem + p.smem_x;
  uint8_t* sKt = smem + p.smem_kt;
  uint8_t* sQt = smem + p.smem_qt;
  uint8_t* sVt = smem + p.smem_vt;
  uint8_t* sP = smem + p.smem_p;
  uint8_t* sPc = smem + p.smem_pc;
  const DcaGeom& g = p.g;
  const int C = g.C, R = g.R;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: role branches stay uniform
  const int lane = threadIdx.x & 31;
  const int x_bytes = p.kb * kXBlockBytes;
  const int kt_block = R * 128, qt_block = p.rq * 128, vt_block = C * 64;
  const int p_bytes = p.kbr * kPBlockBytes;

  // this CTA's contiguous run of whole segments (flattened segment index = image * segs + segment)
  const int total_segs = g.B * g.segs;
  const int s_begin = (int)((long long)blockIdx.x * total_segs / gridDim.x);
  const int s_end = (int)((long long)(blockIdx.x + 1) * total_segs / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&ctrl->x_full[i], 1);
      mbar_init(&ctrl->x_empty[i], p.do_x ? 5 : 1);      // the last MMA that reads the X tile (+ the x-group's residual read)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->p_full[i], 4); mbar_init(&ctrl->sc_full[i], 1); mbar_init(&ctrl->pc_full[i], 4);
    }
    mbar_init(&ctrl->op_full, 1); mbar_init(&ctrl->op_empty, 1);
    mbar_init(&ctrl->s_full, 1); mbar_init(&ctrl->dx_full, 1); mbar_init(&ctrl->dx_free, 4);
    mbar_init(&ctrl->z_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmQt);
    if (p.do_x) { tma_prefetch_desc(&tmKt); tma_prefetch_desc(&tmVt); }
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  // the P' tile keeps zeros wherever nobody writes (the other copy's half of a duplicated row, rows past R)
  for (int i = threadIdx.x; i < 2 * kXBlockBytes / 16; i += kThreads) reinterpret_cast<uint4*>(sPc)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctrl->tmem_base;
  // tensor-memory columns.  serial: [0, C) S then dx | [C, C+128) Sc | [C+128, 2C+128) Z
  //                         PIPE:   [0, 64) S | [64, 64+C) dx | two Sc buffers of 128 | Z      (without x-branch: Sc at 0 / 128, Z at 256)
  const uint32_t colS = 0;
  const uint32_t colDx = (PIPE && p.do_x) ? 64u : 0u;
  const uint32_t colSc = PIPE ? (p.do_x ? 64u + (uint32_t)C : 0u) : (uint32_t)C;
  const uint32_t colZ = colSc + (PIPE ? 256u : 128u);

  // cursors advance incrementally: no divisions on the per-tile paths (the MMA-issuing thread is a single instruction stream)
  auto cur_valid = [&](const Cursor& c) { return c.s < s_end; };
  auto cur_advance = [&](Cursor& c) {
    ++c.it;
    if (++c.xb == p.nx) { c.xb = 0; ++c.xu; }
    if (++c.t == c.nt) {
      c.t = 0; ++c.s;
      if (++c.sg == g.segs) { c.sg = 0; ++c.b; }
      c.t0 = c.sg * g.seg_tiles;
      c.nt = min(g.seg_tiles, g.tiles - c.t0);
    }
  };
  auto cur_last_of_image = [&](const Cursor& c) {   // last tile this CTA processes of the tile's image
    return c.t + 1 == c.nt && (c.s + 1 == s_end || c.sg + 1 == g.segs);
  };
  Cursor c_begin;
  c_begin.s = s_begin; c_begin.t = 0; c_begin.it = 0;
  c_begin.b = s_begin / g.segs; c_begin.sg = s_begin - c_begin.b * g.segs;
  c_begin.t0 = c_begin.sg * g.seg_tiles; c_begin.nt = min(g.seg_tiles, g.tiles - c_begin.t0);
  c_begin.xb = 0; c_begin.xu = 0;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int run = 0, cur_b = -1;
      for (Cursor c = c_begin; cur_valid(c); cur_advance(c)) {
        const int b = c.b;
        if (b != cur_b) {
          // per-image operands: every MMA that read the previous image's operands has retired
          mbar_wait(&ctrl->op_empty, ((uint32_t)run & 1u) ^ 1u, 10);
          const uint32_t bytes = (uint32_t)(p.kb * (g.dup ? 2 : 1) * R * 128 + (p.do_x ? p.kb * kt_block + p.kbr * vt_block : 0));
          mbar_expect_tx(&ctrl->op_full, bytes);
          for (int kb = 0; kb < p.kb; ++kb) {
            tma_load_3d(sQt + (size_t)kb * qt_block, &tmQt, &ctrl->op_full, kb * 64, 0, b);
            if (g.dup) tma_load_3d(sQt + (size_t)kb * qt_block + 64 * 128, &tmQt, &ctrl->op_full, kb * 64, 0, b);
            if (p.do_x) tma_load_3d(sKt + (size_t)kb * kt_block, &tmKt, &ctrl->op_full, kb * 64, 0, b);
          }
          if (p.do_x)
            for (int j = 0; j < p.kbr; ++j) tma_load_3d(sVt + (size_t)j * vt_block, &tmVt, &ctrl->op_full, j * 32, 0, b);
          cur_b = b;
          ++run;
        }
        const int buf = c.xb;
        const uint32_t use = (uint32_t)c.xu;
        mbar_wait(&ctrl->x_empty[buf], (use & 1u) ^ 1u, 11);
        mbar_expect_tx(&ctrl->x_full[buf], (uint32_t)x_bytes);
        for (int kb = 0; kb < p.kb; ++kb)
          tma_load_3d(sX + (size_t)buf * x_bytes + (size_t)kb * kXBlockBytes, &tmX, &ctrl->x_full[buf], kb * 64, (c.t0 + c.t) * kDcaTile, b);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues ----------------
    {
      const uint32_t idesc_s = make_idesc_bf16(128, R);
      const uint32_t idesc_sc = make_idesc_bf16(128, kDcaTile);
      const uint32_t idesc_dx = make_idesc_bf16(128, C);
      const uint32_t idesc_z = make_idesc_bf16(128, C) | (1u << 16);          // B (the X tile) is MN-major
      // operand descriptors are affine in the smem address: build the bases once, add block / K-step offsets per instruction
      const uint64_t dX0 = make_kmajor_desc<128>(smem_u32(sX)), dKt = make_kmajor_desc<128>(smem_u32(sKt));
      const uint64_t dQt = make_kmajor_desc<128>(smem_u32(sQt)), dVt = make_kmajor_desc<64>(smem_u32(sVt));
      const uint64_t dP0 = make_kmajor_desc<64>(smem_u32(sP)), dPc = make_kmajor_desc<128>(smem_u32(sPc));
      const uint64_t dXmn0 = make_mnmajor_sw128_desc(smem_u32(sX), kXBlockBytes);
      const uint64_t x_step = (uint64_t)(x_bytes >> 4), kt_step = (uint64_t)(kt_block >> 4), qt_step = (uint64_t)(qt_block >> 4);
      const uint64_t vt_step = (uint64_t)(vt_block >> 4), p_step = (uint64_t)(p_bytes >> 4);
      constexpr uint64_t kXB = kXBlockBytes >> 4, kPB = kPBlockBytes >> 4;
      const int ks_last = (C - (p.kb - 1) * 64) / 16;          // K-steps of the last 64-channel block (4 in the others)
      const int ksr_last = (R - (p.kbr - 1) * 32) / 16;        // K-steps of the last 32-row block of P / Vt (2 in the others)
      auto issue_s = [&](const Cursor& c) {
        const uint64_t da0 = dX0 + (uint64_t)c.xb * x_step;
        for (int kb = 0; kb < p.kb; ++kb) {
          const uint64_t da = da0 + (uint64_t)kb * kXB, db = dKt + (uint64_t)kb * kt_step;
          const int ks = (kb + 1 == p.kb) ? ks_last : 4;
#pragma unroll 4
          for (int k = 0; k < ks; ++k) umma_bf16_ss_warp(tmem + colS, da + 2ull * k, db + 2ull * k, idesc_s, (uint32_t)((kb | k) != 0));
        }
        umma_commit_warp(&ctrl->s_full);
      };
      auto issue_sc = [&](const Cursor& c) {
        const uint64_t db0 = dX0 + (uint64_t)c.xb * x_step;
        const int sb = PIPE ? (c.it & 1) : 0;
        const uint32_t d = tmem + colSc + (uint32_t)sb * 128u;
        for (int kb = 0; kb < p.kb; ++kb) {
          const uint64_t da = dQt + (uint64_t)kb * qt_step, db = db0 + (uint64_t)kb * kXB;
          const int ks = (kb + 1 == p.kb) ? ks_last : 4;
#pragma unroll 4
          for (int k = 0; k < ks; ++k) umma_bf16_ss_warp(d, da + 2ull * k, db + 2ull * k, idesc_sc, (uint32_t)((kb | k) != 0));
        }
        umma_commit_warp(&ctrl->sc_full[sb]);
      };
      auto issue_dx = [&](int it) {
        const uint64_t da0 = dP0 + (uint64_t)(PIPE ? (it & 1) : 0) * p_step;
        for (int j = 0; j < p.kbr; ++j) {
          const uint64_t da = da0 + (uint64_t)j * kPB, db = dVt + (uint64_t)j * vt_step;
          const int ks = (j + 1 == p.kbr) ? ksr_last : 2;
          for (int k = 0; k < ks; ++k) umma_bf16_ss_warp(tmem + colDx, da + 2ull * k, db + 2ull * k, idesc_dx, (uint32_t)((j | k) != 0));
        }
        umma_commit_warp(&ctrl->dx_full);
      };
      auto issue_z = [&](const Cursor& c) {
        const int t = c.t;
        const uint64_t xo = (uint64_t)c.xb * x_step;
        if (!p.zsplit) {
          const uint64_t db0 = dXmn0 + xo;
#pragma unroll
          for (int k = 0; k < kDcaTile / 16; ++k)      // A: 64-token K-blocks of P' (4 K-steps each); B: 16 token rows = 2048 B per K-step
            umma_bf16_ss_warp(tmem + colZ, dPc + (uint64_t)(k >> 2) * kXB + 2ull * (k & 3), db0 + (uint64_t)k * (2048 >> 4), idesc_z, (uint32_t)((t | k) != 0));
        } else {
          for (int kb = 0; kb < p.kb; ++kb) {
            const int nn = min(64, C - kb * 64);
            const uint32_t idz = make_idesc_bf16(128, nn) | (1u << 16);
            const uint64_t db0 = dXmn0 + xo + (uint64_t)kb * kXB;
            for (int k = 0; k < kDcaTile / 16; ++k)
              umma_bf16_ss_warp(tmem + colZ + (uint32_t)kb * 64u, dPc + (uint64_t)(k >> 2) * kXB + 2ull * (k & 3), db0 + (uint64_t)k * (2048 >> 4), idz,
                           (uint32_t)((t | k) != 0));
          }
        }
        umma_commit_warp(&ctrl->z_done);
        umma_commit_warp(&ctrl->x_empty[c.xb]);       // the Z accumulation is the last reader of the X tile
      };
      if (!PIPE) {
        // ---- one tile at a time: S, Sc, (softmax) dx, (softmax) Z
        int run = 0, cur_b = -1;
        for (Cursor c = c_begin; cur_valid(c); cur_advance(c)) {
          const int b = c.b;
          if (b != cur_b) {
            mbar_wait(&ctrl->op_full, (uint32_t)run & 1u, 20);
            cur_b = b;
            ++run;
          }
          const uint32_t par = (uint32_t)c.it & 1u;
          mbar_wait(&ctrl->x_full[c.xb], (uint32_t)c.xu & 1u, 21);
          if (p.do_x) {
            mbar_wait(&ctrl->dx_free, par ^ 1u, 22);     // the x-group has drained dx of the previous tile (same columns as S)
            tc_fence_after();
            issue_s(c);
          }
          // Sc: the c-group finished reading the previous tile's scores before it signalled pc_full, awaited below
          tc_fence_after();
          issue_sc(c);
          if (p.do_x) {
            mbar_wait(&ctrl->p_full[0], par, 23);
            tc_fence_after();
            issue_dx(c.it);
          }
          mbar_wait(&ctrl->pc_full[0], par, 24);
          tc_fence_after();
          issue_z(c);
          if (cur_last_of_image(c)) umma_commit_warp(&ctrl->op_empty);
        }
      } else {
        // ---- event loop over four cursors; every barrier test is of the current or the immediately preceding phase
        Cursor cS = c_begin, cSc = c_begin, cDx = c_begin, cZ = c_begin;
        int run = 0, cur_b = -1;
        long long t_idle = clock64();
        TR_INIT
        // the tile's image operands are resident: when the image changes, everything of the previous image must have been issued
        auto image_ready = [&](const Cursor& c) {
          const int b = c.b;
          if (b == cur_b) return true;
          if (cZ.it != c.it || (p.do_x && (cDx.it != c.it || cS.it != cSc.it))) return false;
          if (cur_b >= 0) umma_commit_warp(&ctrl->op_empty);
          mbar_wait(&ctrl->op_full, (uint32_t)run & 1u, 20);
          cur_b = b;
          ++run;
          return true;
        };
        while (cur_valid(cZ) || (p.do_x && cur_valid(cDx))) {
          bool progress = false;
          // Z(i) releases the X tile: S(i) and Sc(i), its other readers, must have been issued before
          if (cur_valid(cZ) && cZ.it < cSc.it && (!p.do_x || cZ.it < cS.it) &&
              mbar_test_wait_warp(&ctrl->pc_full[cZ.it & 1], (uint32_t)(cZ.it >> 1) & 1u)) {
            tc_fence_after();
            TR(0)
            issue_z(cZ);
            TR(5)
            cur_advance(cZ);
            progress = true;
          }
          if (p.do_x && cur_valid(cDx) && cDx.it < cS.it && mbar_test_wait_warp(&ctrl->p_full[cDx.it & 1], (uint32_t)(cDx.it >> 1) & 1u) &&
              (cDx.it == 0 || mbar_test_wait_warp(&ctrl->dx_free, (uint32_t)(cDx.it - 1) & 1u))) {
            tc_fence_after();
            TR(0)
            issue_dx(cDx.it);
            TR(4)
            cur_advance(cDx);
            progress = true;
          }
          if (p.do_x && cur_valid(cS) && cS.it <= cSc.it && image_ready(cS) &&
              mbar_test_wait_warp(&ctrl->x_full[cS.xb], (uint32_t)cS.xu & 1u) &&
              (cS.it == 0 || mbar_test_wait_warp(&ctrl->p_full[(cS.it - 1) & 1], (uint32_t)((cS.it - 1) >> 1) & 1u))) {
            tc_fence_after();
            TR(0)
            issue_s(cS);
            TR(2)
            cur_advance(cS);
            progress = true;
          }
          if (cur_valid(cSc) && (!p.do_x || cSc.it < cS.it) && image_ready(cSc) &&
              mbar_test_wait_warp(&ctrl->x_full[cSc.xb], (uint32_t)cSc.xu & 1u) &&
              (cSc.it < 2 || mbar_test_wait_warp(&ctrl->pc_full[cSc.it & 1], (uint32_t)((cSc.it >> 1) - 1) & 1u))) {
            tc_fence_after();
            TR(0)
            issue_sc(cSc);
            TR(3)
            cur_advance(cSc);
            progress = true;
          }
          if (progress) { t_idle = clock64(); TR(0) }
          else { TR(1) if (clock64() - t_idle > (1ll << 31)) __trap(); }     // protocol bug: fail the launch, never hang
        }
        TR_FLUSH(1)
      }
    }
  } else if (warp < kCGroup0) {
    // ---------------- x-group: softmax over the 16 meta keys of every head, then the output epilogue ----------------
    if (p.do_x) {
      const int q = warp & 3;                       // TMEM lane quarter of this warp
      const int row = q * 32 + lane;                // token inside the tile == TMEM lane
      const int gt = threadIdx.x - kXGroup0 * 32;   // 0..127 inside the group
      const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
      for (int i = gt; i < C; i += 128) sBias[i] = __ldg(p.bpx + i);
      int cur_b = -1;
      TR_INIT
      // LayerNorm statistics of this thread's token in the NEXT softmax tile: loaded one tile ahead (global latency off the chain)
      float2 nst[4];
      auto prefetch_stats = [&](const Cursor& c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
        if (!cur_valid(c)) return;
        const int tok = (c.t0 + c.t) * kDcaTile + row;
        if (tok < g.N) {
          const float2* st = reinterpret_cast<const float2*>(p.stats1) + ((long long)c.b * g.N + tok) * p.parts1;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < p.parts1) nst[k] = __ldg(st + k);
        }
      };
      prefetch_stats(c_begin);
      // ---- scores -> probabilities of tile c (16 columns = one head; two heads per tensor-memory load) -> P tile (A operand of dx)
      auto softmax_x = [&](const Cursor& c) {
        const int b = c.b;
        if (b != cur_b) {
          group_sync(1);                            // nobody still reads the previous image's constants
          const float* cst = p.cst + (size_t)b * 4 * R;
          for (int i = gt; i < 2 * R; i += 128) sKc[i] = __ldg(cst + i);
          group_sync(1);
          cur_b = b;
        }
        float ln_r, ln_n;                           // score = ln_r * acc + ln_n * sum(Kt) + kappa
        {
          const float s1 = (nst[0].x + nst[1].x) + (nst[2].x + nst[3].x), s2 = (nst[0].y + nst[1].y) + (nst[2].y + nst[3].y);
          const float mu = s1 * p.inv_c;
          ln_r = rsqrtf(fmaxf(fmaf(s2, p.inv_c, -mu * mu), 0.f) + p.eps);
          ln_n = -ln_r * mu;
          Cursor n = c;
          cur_advance(n);
          prefetch_stats(n);
        }
        uint8_t* pbuf = sP + (size_t)(PIPE ? (c.it & 1) : 0) * p_bytes;
        TR(0)
        mbar_wait(&ctrl->s_full, (uint32_t)c.it & 1u, 30);
        tc_fence_after();
        TR(1)
        auto one_head = [&](int h, const uint32_t* v) {
          float sc[16];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sc[j] = fmaf(ln_r, __uint_as_float(v[j]), fmaf(ln_n, sKc[h * kDcaM + j], sKc[R + h * kDcaM + j]));
            mx = fmaxf(mx, sc[j]);
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { sc[j] = ex2_approx(sc[j] - mx); sum += sc[j]; }
          const float inv = 1.f / sum;
          uint4 u0, u1;
          u0.x = pack_bf16x2(sc[0] * inv, sc[1] * inv); u0.y = pack_bf16x2(sc[2] * inv, sc[3] * inv);
          u0.z = pack_bf16x2(sc[4] * inv, sc[5] * inv); u0.w = pack_bf16x2(sc[6] * inv, sc[7] * inv);
          u1.x = pack_bf16x2(sc[8] * inv, sc[9] * inv); u1.y = pack_bf16x2(sc[10] * inv, sc[11] * inv);
          u1.z = pack_bf16x2(sc[12] * inv, sc[13] * inv); u1.w = pack_bf16x2(sc[14] * inv, sc[15] * inv);
          // P tile: 32-column K-blocks [128 rows x 64 B], 64B swizzle: 16-byte chunk index XOR ((row >> 1) & 3)
          uint8_t* prow = pbuf + (size_t)(h >> 1) * kPBlockBytes + (size_t)row * 64;
          const int ch = (h & 1) * 2, sw = (row >> 1) & 3;
          *reinterpret_cast<uint4*>(prow + (((ch) ^ sw) << 4)) = u0;
          *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ sw) << 4)) = u1;
        };
        for (int h = 0; h < g.heads; h += 2) {
          if (h + 1 < g.heads) {
            uint32_t v[32];
            tmem_ld_x32(lane_addr + colS + (uint32_t)(h * kDcaM), v);
            tmem_ld_wait();
            one_head(h, v);
            one_head(h + 1, v + 16);
          } else {
            uint32_t v[16];
            tmem_ld_x16(lane_addr + colS + (uint32_t)(h * kDcaM), v);
            tmem_ld_wait();
            one_head(h, v);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->p_full[PIPE ? (c.it & 1) : 0]);
        TR(2)
      };
      // ---- output epilogue of tile c: x_out = X + dx + b_px, LayerNorm statistics of the stored rows.  The residual is the X tile
      // itself, still in shared memory (128B-swizzled K-blocks): no second trip to global memory
      auto epilogue_x = [&](const Cursor& c) {
        const int tok = (c.t0 + c.t) * kDcaTile + row;
        const bool rok = tok < g.N;
        const long long grow = (long long)c.b * g.N + tok;
        const int xbuf = c.xb;
        const uint8_t* xrow = sX + (size_t)xbuf * x_bytes + (size_t)row * 128;
        TR(4)
        mbar_wait(&ctrl->x_full[xbuf], (uint32_t)c.xu & 1u, 32);     // (completed long ago: acquire of the TMA writes)
        mbar_wait(&ctrl->dx_full, (uint32_t)c.it & 1u, 31);
        tc_fence_after();
        TR(3)
        float st1 = 0.f, st2 = 0.f;
        for (int c0 = 0; c0 < C; c0 += 32) {
         {
          uint32_t v[32];
          tmem_ld_x32(lane_addr + colDx + (uint32_t)c0, v);
          uint4 res[4];
          {
            const uint8_t* xb = xrow + (size_t)(c0 >> 6) * kXBlockBytes;
            const int ch0 = (c0 & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) res[i] = *reinterpret_cast<const uint4*>(xb + (((ch0 + i) ^ (row & 7)) << 4));
          }
          tmem_ld_wait();
          if (rok) {
            uint4 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0 + 8 * i);
              const float4 b1 = *reinterpret_cast<const float4*>(sBias + c0 + 8 * i + 4);
              const float2 r0 = unpack_bf16x2(res[i].x), r1 = unpack_bf16x2(res[i].y), r2 = unpack_bf16x2(res[i].z), r3 = unpack_bf16x2(res[i].w);
              o[i].x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) + b0.x + r0.x, __uint_as_float(v[8 * i + 1]) + b0.y + r0.y);
              o[i].y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) + b0.z + r1.x, __uint_as_float(v[8 * i + 3]) + b0.w + r1.y);
              o[i].z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) + b1.x + r2.x, __uint_as_float(v[8 * i + 5]) + b1.y + r2.y);
              o[i].w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) + b1.z + r3.x, __uint_as_float(v[8 * i + 7]) + b1.w + r3.y);
              const float2 f0 = unpack_bf16x2(o[i].x), f1 = unpack_bf16x2(o[i].y), f2 = unpack_bf16x2(o[i].z), f3 = unpack_bf16x2(o[i].w);
              st1 += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
              st2 = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, st2))));
              st2 = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, fmaf(f3.x, f3.x, fmaf(f3.y, f3.y, st2))));
            }
            // lane = token stores: the LSU pays per cache line touched — one full 32-byte sector per lane and instruction where the
            // rows are 32-byte aligned (C % 16 == 0 always holds here)
            bf16* op = p.xout + grow * C + c0;
            if ((reinterpret_cast<uintptr_t>(p.xout) & 31) == 0 && (C & 15) == 0) {
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
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctrl->dx_free);
          mbar_arrive(&ctrl->x_empty[xbuf]);        // the X tile (residual source) may be refilled
        }
        if (rok) *reinterpret_cast<float2*>(p.stats2 + grow * 2) = make_float2(st1, st2);
        TR(4)
      };
      if (!PIPE) {
        for (Cursor c = c_begin; cur_valid(c); cur_advance(c)) {
          softmax_x(c);
          epilogue_x(c);
        }
      } else {
        // softmax of tile i + 1 before the epilogue of tile i: the dx MMA of tile i runs behind the softmax of tile i + 1
        Cursor cs = c_begin, ce = c_begin;
        if (cur_valid(cs)) { softmax_x(cs); cur_advance(cs); }
        while (cur_valid(ce)) {
          if (cur_valid(cs)) { softmax_x(cs); cur_advance(cs); }
          epilogue_x(ce);
          cur_advance(ce);
        }
      }
      if (warp == kXGroup0) { TR_FLUSH(2) }
    }
  } else {
    // ---------------- c-group: softmax over the image tokens, thread = (head, query) row ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;                        // TMEM lane
    const int gt = threadIdx.x - kCGroup0 * 32;         // 0..127 inside the group (a permutation of the lanes: q != (warp - 6))
    const int rr = g.dup ? (r & 63) : r;                // logical (h, m) row
    const int copy = g.dup ? (r >> 6) : 0;
    const bool rvalid = rr < R;
    const int c_lo = g.dup ? copy * 64 : 0, c_n = g.dup ? 64 : 128;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f, t_run = 0.f;
    float sumq = 0.f, beta = 0.f;
    int cur_b = -1;
    TR_INIT
    // LayerNorm statistics of token `gt` of the NEXT tile: loaded one tile ahead
    float2 nst[4];
    auto prefetch_stats = [&](const Cursor& c) {
#pragma unroll
      for (int k = 0; k < 4; ++k) nst[k] = make_float2(0.f, 0.f);
      if (!cur_valid(c)) return;
      const int tok = (c.t0 + c.t) * kDcaTile + gt;
      if (tok < g.N) {
        const float2* st = reinterpret_cast<const float2*>(p.stats1) + ((long long)c.b * g.N + tok) * p.parts1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < p.parts1) nst[k] = __ldg(st + k);
      }
    };
    prefetch_stats(c_begin);
    for (Cursor c = c_begin; cur_valid(c); cur_advance(c)) {
      const int b = c.b, t = c.t, it = c.it;
      const int t_abs = c.t0 + t;
      if (b != cur_b) {
        if (rvalid) {
          sumq = __ldg(p.cst + (size_t)b * 4 * R + 2 * R + rr);
          beta = __ldg(p.cst + (size_t)b * 4 * R + 3 * R + rr);
        }
        cur_b = b;
      }
      const uint32_t par = (uint32_t)it & 1u;
      const int sb = PIPE ? (it & 1) : 0;                                   // score buffer / barrier slot of this tile
      const uint32_t spar = PIPE ? ((uint32_t)(it >> 1) & 1u) : par;        // its phase
      // (r_n, -r_n mu_n, mu_n) of the tile's 128 tokens -> shared (two buffers: a fast thread may already stage tile it + 1)
      float4* stat = sStat + (it & 1) * 128;
      {
        const float s1 = (nst[0].x + nst[1].x) + (nst[2].x + nst[3].x), s2 = (nst[0].y + nst[1].y) + (nst[2].y + nst[3].y);
        const float mu = s1 * p.inv_c;
        const float rn = (t_abs * kDcaTile + gt < g.N) ? rsqrtf(fmaxf(fmaf(s2, p.inv_c, -mu * mu), 0.f) + p.eps) : 0.f;
        stat[gt] = make_float4(rn, -rn * mu, mu, 0.f);
        Cursor n = c;
        cur_advance(n);
        prefetch_stats(n);
      }
      group_sync(2);
      const int valid = min(kDcaTile, g.N - t_abs * kDcaTile);
      TR(0)
      mbar_wait(&ctrl->sc_full[sb], spar, 40);
      tc_fence_after();
      TR(1)
      const uint32_t s_row = lane_addr + colSc + (uint32_t)sb * 128u + (uint32_t)c_lo;
      // ---- pass 1: tile maximum of this thread's columns (LayerNorm-corrected scores, log2 domain; columns past the end of the
      // image are skipped)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < c_n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(s_row + (uint32_t)c0, v);
        tmem_ld_wait();
        const int nv = valid - (c_lo + c0);
        if (nv >= 32) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float2 rm = *reinterpret_cast<const float2*>(stat + c_lo + c0 + k);
            mx = fmaxf(mx, fmaf(rm.x, __uint_as_float(v[k]), fmaf(rm.y, sumq, beta)));
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float2 rm = *reinterpret_cast<const float2*>(stat + c_lo + c0 + k);
            const float val = fmaf(rm.x, __uint_as_float(v[k]), fmaf(rm.y, sumq, beta));
            mx = fmaxf(mx, (k < nv) ? val : -INFINITY);
          }
        }
      }
      // the previous tile's Z accumulation has retired: P' may be overwritten, Z may be rescaled
      TR(2)
      mbar_wait(&ctrl->z_done, par ^ 1u, 41);
      tc_fence_after();
      TR(3)
      // ---- running maximum with lazy rescaling of the TMEM accumulator (first tile of a segment: Z is overwritten)
      float scale_old = 1.f;
      bool need = false;
      if (t == 0) {
        m_run = mx; l_run = 0.f; t_run = 0.f;
      } else if (mx > m_run + kRescaleThreshold || (m_run == -INFINITY && mx > -INFINITY)) {
        scale_old = (m_run == -INFINITY) ? 0.f : exp2f(m_run - mx);
        m_run = mx;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(lane_addr + colZ + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * scale_old);
          tmem_st_x32(lane_addr + colZ + (uint32_t)c0, v);
        }
        tmem_st_wait();
        l_run *= scale_old;
        t_run *= scale_old;
      }
      // ---- pass 2: p = exp2(score - m), P' = bf16(p r_n) -> shared (A operand of the Z accumulation); the shift is folded into
      // the per-row constant: score - m = r acc + nrm sumq + (beta - m)
      const float beta_m = beta - ((m_run == -INFINITY) ? 0.f : m_run);
#pragma unroll 1
      for (int c0 = 0; c0 < c_n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(s_row + (uint32_t)c0, v);
        tmem_ld_wait();
        const int col0 = c_lo + c0;
        const int nv = valid - col0;
        uint8_t* tile_p = sPc + (size_t)(col0 >> 6) * kXBlockBytes + (size_t)r * 128;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 rm = stat[col0 + gq * 8 + k];
            float pk = ex2_approx(fmaf(rm.x, __uint_as_float(v[gq * 8 + k]), fmaf(rm.y, sumq, beta_m)));
            if (gq * 8 + k >= nv) pk = 0.f;                     // (compiles to a select; nv >= 32 on every full tile)
            l_run += pk;
            e[k] = pk * rm.x;
            t_run = fmaf(e[k], rm.z, t_run);
          }
          uint4 u;
          u.x = pack_bf16x2(e[0], e[1]); u.y = pack_bf16x2(e[2], e[3]);
          u.z = pack_bf16x2(e[4], e[5]); u.w = pack_bf16x2(e[6], e[7]);
          const int ch = ((col0 & 63) >> 3) + gq;
          if (rvalid) *reinterpret_cast<uint4*>(tile_p + ((ch ^ (r & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->pc_full[sb]);
      TR(4)
      // ---- end of the segment: one split-softmax partial per (segment, copy, row)
      if (t + 1 == c.nt) {
        mbar_wait(&ctrl->z_done, par, 42);
        tc_fence_after();
        const long long pr = ((long long)c.s * g.ncopy + copy) * R + rr;
        if (rvalid) p.part_ml[pr] = make_float4(m_run, l_run, t_run, 0.f);
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(lane_addr + colZ + (uint32_t)c0, v);
          tmem_ld_wait();
          if (rvalid) {
            // lane = row stores of 128 contiguous bytes: 256-bit where the partial buffer allows it (the LSU pays per cache line)
            float* dst = p.part_z + pr * C + c0;
            if ((reinterpret_cast<uintptr_t>(p.part_z) & 31) == 0 && (C & 7) == 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i) st_global_256(dst + 8 * i, reinterpret_cast<const uint32_t (&)[8]>(v[8 * i]));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
          }
        }
        tc_fence_before();
        TR(5)
      }
    }
    if (warp == kCGroup0) { TR_FLUSH(3) }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

PerDeviceOnce g_attr_once;

}  // namespace

int dca_x_prepare(const DcaXArgs& a, DcaXOp* op) {
  const DcaGeom& g = a.g;
  LMV_REQUIRE(a.xt && a.stats1 && a.ws.qt && a.ws.cst && a.ws.part_ml && a.ws.part_z, "dca_x: null pointer");
  LMV_REQUIRE(!a.do_x || (a.bpx && a.xout && a.stats2 && a.ws.kt && a.ws.vt), "dca_x: null pointer (x-branch)");
  if (!dca_supported(g.N, g.C, g.heads, kDcaM)) return fail(LMV_ERR_UNSUPPORTED, "dca_x: needs C = heads * 32 <= 192 and 16 meta tokens");
  LMV_REQUIRE(a.parts1 >= 1 && a.parts1 <= 4, "dca_x: 1..4 statistics partials per row");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  LMV_REQUIRE(al16(a.xt) && al16(a.xout ? a.xout : a.xt) && al16(a.ws.kt) && al16(a.ws.qt) && al16(a.ws.vt) && al16(a.ws.part_z),
              "dca_x: pointers must be 16-byte aligned");
  DcaXParams& p = op->p;
  p.g = g;
  p.do_x = a.do_x ? 1 : 0;
  p.kb = (g.C + 63) / 64;
  p.kbr = (g.R + 31) / 32;
  p.rq = g.dup ? 64 + g.R : g.R;
  p.zsplit = a.zsplit;
  p.parts1 = a.parts1;
  p.eps = a.eps; p.inv_c = 1.0f / (float)g.C;
  p.stats1 = a.stats1; p.cst = a.ws.cst; p.bpx = a.bpx; p.xres = a.xt; p.xout = a.xout; p.stats2 = a.stats2;
  p.part_ml = a.ws.part_ml; p.part_z = a.ws.part_z;
  // shared-memory plan (every region 1024-byte aligned: 128B-swizzled tiles)
  auto up = [](int v) { return (v + 1023) & ~1023; };
  const int x_bytes = p.kb * kXBlockBytes;
  const int qt_bytes = up(p.kb * p.rq * 128), kt_bytes = p.do_x ? up(p.kb * g.R * 128) : 0, vt_bytes = p.do_x ? up(p.kbr * g.C * 64) : 0;
  // pipelined schedule where tensor memory has room for a second score buffer and separate S / dx regions (kernel header)
  p.pipe = (!a.force_serial && (p.do_x ? (64 + 2 * g.C + 256 <= 512 && g.R <= 64) : true)) ? 1 : 0;
  const int p_bytes = p.do_x ? (p.pipe ? 2 : 1) * p.kbr * kPBlockBytes : 0, pc_bytes = 2 * kXBlockBytes;
  const int fixed = kHeaderBytes + qt_bytes + kt_bytes + vt_bytes + p_bytes + pc_bytes + 4096 /* Qt overrun of the M = 128 operand read */;
  p.nx = std::min(3, (kSmemLimit - 1024 - fixed) / x_bytes);     // X-tile ring: 3 deep where it fits (the x-group holds a tile until its epilogue)
  LMV_REQUIRE(p.nx >= 1, "dca_x: shared memory budget (X tile)");
  if (p.nx < 2) p.pipe = 0;        // (only reachable with the single P buffer already counted: the pipelined schedule needs both)
  int off = kHeaderBytes;
  p.smem_qt = off; off += qt_bytes;
  p.smem_kt = off; off += kt_bytes;
  p.smem_vt = off; off += vt_bytes;
  p.smem_x = off; off += p.nx * x_bytes;        // directly behind the operands: the Qt over-read (rows R..127) stays inside the allocation
  p.smem_p = off; off += p_bytes;
  p.smem_pc = off; off += pc_bytes;
  op->smem_bytes = off + 1024;
  LMV_REQUIRE(op->smem_bytes <= kSmemLimit, "dca_x: shared memory budget");
  LMV_REQUIRE(qt_bytes + kt_bytes + vt_bytes + p.nx * x_bytes >= p.kb * 128 * 128, "dca_x: operand over-read window");
  op->grid = std::min(g.B * g.segs, device_sm_count());
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)g.C, (uint64_t)g.N, (uint64_t)g.B};
    uint64_t strides[2] = {(uint64_t)g.C * 2, (uint64_t)g.N * g.C * 2};
    uint32_t box[3] = {64, (uint32_t)kDcaTile, 1};
    if ((rc = encode_tmap_bf16(&op->tmX, a.xt, 3, dims, strides, box, 128))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)g.C, (uint64_t)g.R, (uint64_t)g.B};
    uint64_t strides[2] = {(uint64_t)g.C * 2, (uint64_t)g.R * g.C * 2};
    uint32_t box[3] = {64, (uint32_t)g.R, 1};
    if ((rc = encode_tmap_bf16(&op->tmQt, a.ws.qt, 3, dims, strides, box, 128))) return rc;
    op->tmKt = op->tmQt;
    if (a.do_x && (rc = encode_tmap_bf16(&op->tmKt, a.ws.kt, 3, dims, strides, box, 128))) return rc;
  }
  op->tmVt = op->tmQt;
  if (a.do_x) {
    uint64_t dims[3] = {(uint64_t)g.R, (uint64_t)g.C, (uint64_t)g.B};
    uint64_t strides[2] = {(uint64_t)g.R * 2, (uint64_t)g.R * g.C * 2};
    uint32_t box[3] = {32, (uint32_t)g.C, 1};
    if ((rc = encode_tmap_bf16(&op->tmVt, a.ws.vt, 3, dims, strides, box, 64))) return rc;
  }
  return LMV_OK;
}

int dca_x_run(const DcaXOp& op, cudaStream_t s) {
  LMV_CUDA_OK(g_attr_once.run([] {
    const cudaError_t e = cudaFuncSetAttribute(dca_x_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    return e != cudaSuccess ? e : cudaFuncSetAttribute(dca_x_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  }));
  if (op.p.pipe)
    LMV_CUDA_OK(launch_kernel(dca_x_kernel<true>, dim3(op.grid), dim3(kThreads), (size_t)op.smem_bytes, s, op.tmX, op.tmKt, op.tmQt, op.tmVt, op.p));
  else
    LMV_CUDA_OK(launch_kernel(dca_x_kernel<false>, dim3(op.grid), dim3(kThreads), (size_t)op.smem_bytes, s, op.tmX, op.tmKt, op.tmQt, op.tmVt, op.p));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv

#ifdef LMV_DCA_TRACE
// copies out and clears the [148 CTAs][4 roles][8 counters] cycle table of the traced dca_x launches (debug builds only)
extern "C" int lmv_debug_dca_trace(unsigned long long* host, int n) {
  static unsigned long long zero[148 * 4 * 8];
  if (n > 148 * 4 * 8) n = 148 * 4 * 8;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(host, lmv::g_dca_trace, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(lmv::g_dca_trace, zero, sizeof(zero)) != cudaSuccess;
}
#endif
