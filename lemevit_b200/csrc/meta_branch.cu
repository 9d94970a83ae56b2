// Meta-token side of the fused cross-attention blocks: everything that happens to the M = 16 meta tokens of an image in a
// 'C' block (CrossAttention, models/lemevit.py:477-486 inside forward_with_c :584-613) or a 'D' block (DualCrossAttention
// :252-302 inside forward_with_xc :542-582), in TWO kernels per block instead of ~8 launch-latency-bound ones:
//
//   meta_pre   cn = LN1(c);  [q2 | k2 | v2] = Wc cn + bc;  builds the per-image operands of the image-token kernel
//              (dca_fused.cu) with the image-side projections absorbed (kernels.h): Kt, Qt, Vt^T and their constants;
//   meta_post  merges the c-branch softmax partials of the image-token kernel into Zbar = sum_n softmax_n(Sc) xn[n], then
//              attn_c = Wv_h Zbar + bv;  c += proj_c(attn_c);  c += mlp(LN2(c))          (:563-564 / :600-601).
//
// One CTA per IM images (IM = 2 for real batches, see pick_im): the kernels are bound by the latency of streaming the block's
// weights from L2 (and, for the wide downsample layers, by L2 bandwidth), so every weight fragment a lane loads feeds the MMAs of
// IM images.  Rows of image i of the CTA are rows 16 i .. 16 i + 15 of every shared-memory tile; the per-image arithmetic is
// unchanged (same MMA sequence per image: results do not depend on IM or on the batch position).  16 rows are one eighth of a tcgen05 tile (M = 128) but exactly one warp-level tensor-core tile, so every
// contraction here is `mma.sync.m16n8k16` (bf16 in, fp32 accumulate): the 16 x K activation sits in shared memory as bf16 (the A
// fragments of a warp are two 16-byte loads per 32 k), the weights are never staged — each lane reads its B fragment straight from
// L2 with one 16-byte load per (8 columns x 32 k), by pairing the k indices of the two fragments in the order they lie in memory
// (a consistent permutation of k on both operands leaves the dot product unchanged).  Per image: ~1.2 MB of weights from L2 and
// ~19 MFLOP at C = 192, 0.3 MB / 5 MFLOP at C = 96.
#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int M = kDcaM;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
// row pitch (bf16 elements) of a 16 x K activation tile in shared memory: pitch bytes = 64 (mod 128), so that the eight lanes of
// a quarter warp (two rows x four 16-byte chunks) hit 128 distinct bytes when they load their A fragments
__host__ __device__ inline int pitch(int K) { return K + ((K % 64 == 0) ? 32 : 0); }
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// out[m][n] = sum_k A[m][k] W[n][k] for the 16 rows of an image.  A: bf16 in shared memory, row pitch lda (pitch(K), 16-byte
// aligned rows); W: bf16 in global memory, K contiguous, row pitch ldw; K % 32 == 0, N % 8 == 0.  Each warp takes groups of four
// 8-column tiles; epi(n, m, v0, v1) receives out[m][n], out[m][n + 1] (fp32) for the lane's two rows m = g and g + 8.
// a_group_stride: the A tile used for the 32-column group starting at column n is A0 + (n / 32) * a_group_stride (one A tile per
// head in the value contraction of meta_post; 0 everywhere else).
// NT: 8-column tiles per warp trip (4, 2 or 1): chosen by rows16_mma so that all 16 warps have work even when N is small
// (N = 192 has only 24 tiles), at the price of re-loading the A fragments more often.
template <int NT, int IM, typename Epi>
__device__ __forceinline__ void rows16_mma_nt(const bf16* __restrict__ A0, int lda, int a_img_stride, const bf16* __restrict__ W, int ldw, int N, int K,
                                               Epi& epi, int a_group_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int ntiles = N >> 3;
  for (int tile0 = warp * NT; tile0 < ntiles; tile0 += kWarps * NT) {
    const int nt = min(NT, ntiles - tile0);
    const bf16* A = A0 + (size_t)(tile0 >> 2) * a_group_stride + (size_t)g * lda + tq * 8;
    float acc[IM][NT][4];
#pragma unroll
    for (int i = 0; i < IM; ++i)
#pragma unroll
      for (int t = 0; t < NT; ++t) { acc[i][t][0] = 0.f; acc[i][t][1] = 0.f; acc[i][t][2] = 0.f; acc[i][t][3] = 0.f; }
    const bf16* wrow[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) wrow[t] = W + (size_t)((tile0 + min(t, nt - 1)) * 8 + g) * ldw + tq * 8;
#pragma unroll(NT == 4 ? 2 : 4)
    for (int kc = 0; kc < K; kc += 32) {
      uint4 b[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) b[t] = __ldg(reinterpret_cast<const uint4*>(wrow[t] + kc));
#pragma unroll
      for (int i = 0; i < IM; ++i) {
        const uint4 alo = *reinterpret_cast<const uint4*>(A + (size_t)i * a_img_stride + kc);
        const uint4 ahi = *reinterpret_cast<const uint4*>(A + (size_t)i * a_img_stride + (size_t)8 * lda + kc);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mma_bf16_16816(acc[i][t], alo.x, ahi.x, alo.y, ahi.y, b[t].x, b[t].y);
          mma_bf16_16816(acc[i][t], alo.z, ahi.z, alo.w, ahi.w, b[t].z, b[t].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < IM; ++i)
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (t < nt) {
          const int n = (tile0 + t) * 8 + tq * 2;
          epi(n, i * M + g, acc[i][t][0], acc[i][t][1]);
          epi(n, i * M + g + 8, acc[i][t][2], acc[i][t][3]);
        }
  }
}

// a_img_stride: distance (elements) between the A tiles of consecutive images of the CTA; epi's row index m runs over IM * 16.
template <int IM, typename Epi>
__device__ __forceinline__ void rows16_mma(const bf16* __restrict__ A0, int lda, int a_img_stride, const bf16* __restrict__ W, int ldw, int N, int K,
                                            Epi epi, int a_group_stride = 0) {
  const int ntiles = N >> 3;
  if (IM < 4 && ntiles >= 4 * kWarps) rows16_mma_nt<(IM < 4 ? 4 : 2), IM>(A0, lda, a_img_stride, W, ldw, N, K, epi, a_group_stride);
  else if (ntiles >= 2 * kWarps) rows16_mma_nt<2, IM>(A0, lda, a_img_stride, W, ldw, N, K, epi, a_group_stride);
  else rows16_mma_nt<1, IM>(A0, lda, a_img_stride, W, ldw, N, K, epi, a_group_stride);
}

// Per-head contractions with K = 32 (the absorbed projections of meta_pre): out_h[m][n] = sum_d A[m][32h + d] W[n][32h + d] for every
// head h and column n < N.  All (head, 8-column tile) pairs form ONE flat work list over the 16 warps, so the single 16-byte weight
// load each pair needs is in flight for several pairs at once (a rows16_mma call per head would serialise heads x L2 latency).
template <int IM, typename Epi>
__device__ __forceinline__ void heads16_mma_k32(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, int heads, int N, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int ntiles = N >> 3, total = heads * ntiles;
#pragma unroll 4
  for (int w = warp; w < total; w += kWarps) {
    const int h = w / ntiles, tile = w - h * ntiles;
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(tile * 8 + g) * ldw + 32 * h + tq * 8));
    const int n = tile * 8 + tq * 2;
#pragma unroll
    for (int i = 0; i < IM; ++i) {
      const uint4 alo = *reinterpret_cast<const uint4*>(A + (size_t)(i * M + g) * lda + 32 * h + tq * 8);
      const uint4 ahi = *reinterpret_cast<const uint4*>(A + (size_t)(i * M + g + 8) * lda + 32 * h + tq * 8);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(acc, alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
      mma_bf16_16816(acc, alo.z, ahi.z, alo.w, ahi.w, b.z, b.w);
      epi(h, n, i * M + g, acc[0], acc[1]);
      epi(h, n, i * M + g + 8, acc[2], acc[3]);
    }
  }
}

// LayerNorm without affine of the `rows` fp32 rows -> bf16 rows (A operand of the next contraction); one warp per row
__device__ __forceinline__ void rows16_layernorm(const float* in, int ld, bf16* out, int ldo, int C, float eps, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < rows; m += kWarps) {
    float s1 = 0.f;
    for (int k = lane; k < C; k += 32) s1 += in[m * ld + k];
    const float mu = warp_sum(s1) / (float)C;
    float s2 = 0.f;
    for (int k = lane; k < C; k += 32) { const float d = in[m * ld + k] - mu; s2 = fmaf(d, d, s2); }
    const float r = rsqrtf(warp_sum(s2) / (float)C + eps);
    for (int k = lane; k < C; k += 32) out[m * ldo + k] = __float2bfloat16((in[m * ld + k] - mu) * r);
  }
}

// ------------------------------------------------------------------------------------------------
// meta_pre
// ------------------------------------------------------------------------------------------------
// s_c: [IM * M][C] fp32 meta tokens at the front of shared memory (loaded here unless `have_c`: the post body of the same launch
// left the block's output there); `scratch`: the rest of the dynamic shared memory.  bi[i]: batch index of the CTA's image i
// (an odd batch makes the last CTA run its only image twice: identical values are stored twice).
template <int IM>
__device__ __forceinline__ void meta_pre_body(const MetaPreArgs& a, float* s_c, uint8_t* scratch, const int (&bi)[IM], bool have_c) {
  const int C = a.C, R = a.heads * M, nc = a.nc;
  const int lda = pitch(C), ldp = pitch(nc);
  bf16* s_n = reinterpret_cast<bf16*>(scratch);                    // [IM * M][C + pad]  LN1(c)
  bf16* s_p = s_n + IM * M * lda;                                  // [IM * M][nc + pad] projection of LN1(c): q2 | k2 | v2
  bf16* s_t = s_p + IM * M * ldp;                                  // [IM][R][C] staging of Kt / Qt, [IM][C][R] staging of Vt^T
  if (!have_c) {
#pragma unroll
    for (int i = 0; i < IM; ++i) {
      const bf16* cb = a.c + (size_t)bi[i] * M * C;
      for (int j = threadIdx.x; j < M * C; j += kThreads) s_c[i * M * C + j] = __bfloat162float(cb[j]);
    }
  }
  __syncthreads();
  rows16_layernorm(s_c, C, s_n, lda, C, a.eps, IM * M);
  __syncthreads();
  rows16_mma<IM>(s_n, lda, M * lda, a.Wc, C, nc, C, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.bc + n));
    *reinterpret_cast<uint32_t*>(s_p + m * ldp + n) = pack_bf16x2(v0 + bias.x, v1 + bias.y);
  });
  __syncthreads();

  // T[(h,m)][j] = scale * sum_d vec[m][32h + d] Wx[32h + d][j], through the transposed copy WxT[j][i] (K contiguous);
  // then the row sums of the bf16-rounded T and the bias constant scale * sum_d bx[32h + d] vec[m][32h + d]
  auto absorb = [&](int off, const bf16* WxT, const float* bx, float scale, bf16* out_g, int sum_off) {
    heads16_mma_k32<IM>(s_p + off, ldp, WxT, C, a.heads, C, [&](int h, int n, int m, float v0, float v1) {
      const int i = m >> 4, mm = m & 15;
      *reinterpret_cast<uint32_t*>(s_t + ((size_t)i * R + h * M + mm) * C + n) = pack_bf16x2(v0 * scale, v1 * scale);
    });
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int ri = warp; ri < IM * R; ri += kWarps) {
      const int i = ri / R, r = ri - i * R;
      float s = 0.f;
      for (int k = lane; k < C; k += 32) s += __bfloat162float(s_t[(size_t)ri * C + k]);
      s = warp_sum(s);
      const int h = r / M, m = r - h * M;
      const float kc = __bfloat162float(s_p[(i * M + m) * ldp + off + 32 * h + lane]) * __ldg(bx + 32 * h + lane);
      const float k2 = warp_sum(kc) * scale;
      if (lane == 0) {
        float* cst = a.ws.cst + (size_t)bi[i] * 4 * R + sum_off;
        cst[r] = s;
        cst[R + r] = k2;
      }
    }
#pragma unroll
    for (int i = 0; i < IM; ++i) {
      const uint4* src = reinterpret_cast<const uint4*>(s_t + (size_t)i * R * C);
      uint4* dst = reinterpret_cast<uint4*>(out_g + (size_t)bi[i] * R * C);
      for (int j = threadIdx.x; j < R * C / 8; j += kThreads) dst[j] = src[j];
    }
    __syncthreads();
  };
  if (a.WxqT)   // 'D': x-branch keys in token space
    absorb(a.k_off, a.WxqT, a.bxq, a.scale_x * kLog2e, a.ws.kt, 0);
  absorb(a.q_off, a.WxkT, a.bxk, a.scale_c * kLog2e, a.ws.qt, 2 * R);
  if (a.Wpx) {
    // Vt^T[j][(h,m)] = sum_d Wpx[j][32h + d] v2[m][32h + d]
    heads16_mma_k32<IM>(s_p + a.v_off, ldp, a.Wpx, C, a.heads, C, [&](int h, int n, int m, float v0, float v1) {
      const int i = m >> 4, mm = m & 15;
      bf16* t = s_t + (size_t)i * C * R;
      t[(size_t)n * R + h * M + mm] = __float2bfloat16(v0);
      t[(size_t)(n + 1) * R + h * M + mm] = __float2bfloat16(v1);
    });
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IM; ++i) {
      const uint4* src = reinterpret_cast<const uint4*>(s_t + (size_t)i * C * R);
      uint4* dst = reinterpret_cast<uint4*>(a.ws.vt + (size_t)bi[i] * C * R);
      for (int j = threadIdx.x; j < R * C / 8; j += kThreads) dst[j] = src[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// meta_post
// ------------------------------------------------------------------------------------------------
// Leaves the block's output meta tokens (the bf16-rounded values that were stored) in s_c for a following pre body.
template <int IM>
__device__ __forceinline__ void meta_post_body(const MetaPostArgs& a, float* s_c, uint8_t* scratch, const int (&bi)[IM]) {
  const int C = a.C, R = a.heads * M, Hd = a.Hd, P = a.parts;
  const int lda = pitch(C), ldh = pitch(Hd);
  float* s_w = reinterpret_cast<float*>(scratch);              // [IM][P + 1][R]  merge weights, 1 / l
  bf16* s_a = reinterpret_cast<bf16*>(s_w + IM * (P + 1) * R); // [IM * M][C + pad]   attn_c, later LN2(c)
  bf16* s_z = s_a + IM * M * lda;                              // [IM][R][C + pad]   Zbar (bf16); reused as the MLP hidden activation [IM * M][Hd + pad]
#pragma unroll
  for (int i = 0; i < IM; ++i) {
    const bf16* cb = a.c + (size_t)bi[i] * M * C;
    for (int j = threadIdx.x; j < M * C; j += kThreads) s_c[i * M * C + j] = __bfloat162float(cb[j]);
  }
  // ---- merge the segment partials (fixed order: deterministic, independent of batch size and position)
  for (int ri = threadIdx.x; ri < IM * R; ri += kThreads) {
    const int i = ri / R, r = ri - i * R;
    const float4* ml = a.ws.part_ml + (size_t)bi[i] * P * R;
    float* w_i = s_w + i * (P + 1) * R;
    float mx = -INFINITY;
    for (int p = 0; p < P; ++p) mx = fmaxf(mx, ml[p * R + r].x);
    float l = 0.f;
    for (int p = 0; p < P; ++p) {
      const float4 v = ml[p * R + r];
      const float w = (v.x == -INFINITY) ? 0.f : exp2f(v.x - mx);
      w_i[p * R + r] = w;
      l = fmaf(v.y, w, l);
    }
    w_i[P * R + r] = 1.f / l;
  }
  __syncthreads();
  const int per_img = R * C / 4;
  for (int ji = threadIdx.x; ji < IM * per_img; ji += kThreads) {
    const int i = ji / per_img, j = ji - i * per_img;
    const int r = (j * 4) / C, col = j * 4 - r * C;
    const float4* ml = a.ws.part_ml + (size_t)bi[i] * P * R;
    const float* pz = a.ws.part_z + (size_t)bi[i] * P * R * C;
    const float* w_i = s_w + i * (P + 1) * R;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p0 = 0; p0 < P; p0 += 4) {      // four partials in flight
      float4 z[4];
      float w[4], t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = min(p0 + q, P - 1);
        w[q] = (p0 + q < P) ? w_i[p * R + r] : 0.f;
        z[q] = __ldg(reinterpret_cast<const float4*>(pz + (size_t)p * R * C) + j);
        t[q] = ml[p * R + r].z;     // sum_n p'_n mu_n: Zbar = sum p' (xt - mu)
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (w[q] != 0.f) {          // an empty partial (no valid token in that copy's columns) may hold anything
          acc.x = fmaf(w[q], z[q].x - t[q], acc.x); acc.y = fmaf(w[q], z[q].y - t[q], acc.y);
          acc.z = fmaf(w[q], z[q].z - t[q], acc.z); acc.w = fmaf(w[q], z[q].w - t[q], acc.w);
        }
    }
    const float il = w_i[P * R + r];
    uint2 pk;
    pk.x = pack_bf16x2(acc.x * il, acc.y * il);
    pk.y = pack_bf16x2(acc.z * il, acc.w * il);
    *reinterpret_cast<uint2*>(s_z + ((size_t)i * R + r) * lda + col) = pk;
  }
  __syncthreads();
  // ---- attn_c[m][32h + d] = Wv[32h + d][:] . Zbar[(h,m)][:] + bv
  rows16_mma<IM>(s_z, lda, R * lda, a.Wxv, C, C, C, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.bxv + n));
    *reinterpret_cast<uint32_t*>(s_a + m * lda + n) = pack_bf16x2(v0 + bias.x, v1 + bias.y);
  }, M * lda);
  __syncthreads();
  // ---- c += proj(attn_c)
  rows16_mma<IM>(s_a, lda, M * lda, a.Wp, C, C, C, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.bp + n));
    s_c[m * C + n] += v0 + bias.x;
    s_c[m * C + n + 1] += v1 + bias.y;
  });
  __syncthreads();
  // ---- c += mlp(LN2(c))   (LN2 affine folded into W1 / b1)
  rows16_layernorm(s_c, C, s_a, lda, C, a.eps, IM * M);
  __syncthreads();
  bf16* s_h = s_z;
  rows16_mma<IM>(s_a, lda, M * lda, a.W1, C, Hd, C, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.b1 + n));
    *reinterpret_cast<uint32_t*>(s_h + (size_t)m * ldh + n) = pack_bf16x2(gelu_erf(v0 + bias.x), gelu_erf(v1 + bias.y));
  });
  __syncthreads();
  rows16_mma<IM>(s_h, ldh, M * ldh, a.W2, Hd, C, Hd, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.b2 + n));
    const uint32_t pk = pack_bf16x2(s_c[m * C + n] + v0 + bias.x, s_c[m * C + n + 1] + v1 + bias.y);
    *reinterpret_cast<uint32_t*>(a.c + ((size_t)bi[m >> 4] * M + (m & 15)) * C + n) = pk;
    const float2 f = unpack_bf16x2(pk);
    s_c[m * C + n] = f.x;
    s_c[m * C + n + 1] = f.y;
  });
}

// One launch = [post of block j] -> [pre of block j + 1] of the same stage (either may be absent): one meta-token kernel per block.
template <int IM>
__global__ void __launch_bounds__(kThreads)
meta_chain_kernel(MetaPostArgs post, MetaPreArgs pre, int do_post, int do_pre, int C, int B) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_c = reinterpret_cast<float*>(smem_raw);
  uint8_t* scratch = smem_raw + (size_t)IM * M * C * 4;
  int bi[IM];
#pragma unroll
  for (int i = 0; i < IM; ++i) bi[i] = min((int)blockIdx.x * IM + i, B - 1);
  pdl_launch_dependents();
  pdl_wait();
  if (do_post) {
    meta_post_body<IM>(post, s_c, scratch, bi);
    __syncthreads();
  }
  if (do_pre) meta_pre_body<IM>(pre, s_c, scratch, bi, do_post != 0);
}

// meta_token_downsample[i] (models/lemevit.py:729-745): Linear(Cp, 4Cp) -> LayerNorm(eps 1e-5) -> GELU -> Linear(4Cp, C) -> LayerNorm,
// one CTA per IM images instead of five launches.  in / out rows of image b start at in + b * in_bs / out + b * out_bs (the meta
// tokens may live behind the image tokens of a unified [B, N + M, C] buffer).  The reference's bf16 pipeline stores each Linear
// output before the LayerNorm, so those rows are kept as bf16 in shared memory (the hidden rows are normalised in place).
template <int IM>
__global__ void __launch_bounds__(kThreads)
meta_downsample_kernel(MetaDsArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int Cp = a.Cp, C = a.C, H4 = 4 * a.Cp;
  const int ldi = pitch(Cp), ldh = pitch(H4);
  bf16* s_i = reinterpret_cast<bf16*>(smem_raw);                   // [IM * M][Cp + pad]  input rows
  bf16* s_h = s_i + IM * M * ldi;                                  // [IM * M][4Cp + pad] hidden rows
  bf16* s_o = s_h + IM * M * ldh;                                  // [IM * M][C] output rows before the last LayerNorm
  int bi[IM];
#pragma unroll
  for (int i = 0; i < IM; ++i) bi[i] = min((int)blockIdx.x * IM + i, a.B - 1);
  pdl_launch_dependents();
  pdl_wait();
#pragma unroll
  for (int i = 0; i < IM; ++i) {
    const bf16* in = a.in + (size_t)bi[i] * a.in_bs;
    for (int j = threadIdx.x; j < M * Cp / 8; j += kThreads) {
      const int m = (j * 8) / Cp, k = j * 8 - m * Cp;
      *reinterpret_cast<uint4*>(s_i + (i * M + m) * ldi + k) = *reinterpret_cast<const uint4*>(in + (size_t)m * Cp + k);
    }
  }
  __syncthreads();
  rows16_mma<IM>(s_i, ldi, M * ldi, a.W0, Cp, H4, Cp, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.b0 + n));
    *reinterpret_cast<uint32_t*>(s_h + (size_t)m * ldh + n) = pack_bf16x2(v0 + bias.x, v1 + bias.y);
  });
  __syncthreads();
  // LayerNorm(width, affine, eps) [-> GELU] of bf16 rows -> bf16 rows (may be in place: a lane re-reads only what it writes); one warp per row
  auto ln_rows = [&](const bf16* src, int lds, int width, const float* gamma, const float* beta, bool gelu, auto out_row) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = warp; m < IM * M; m += kWarps) {
      const bf16* row = src + (size_t)m * lds;
      float s1 = 0.f;
      for (int k = lane; k < width; k += 32) s1 += __bfloat162float(row[k]);
      const float mu = warp_sum(s1) / (float)width;
      float s2 = 0.f;
      for (int k = lane; k < width; k += 32) { const float d = __bfloat162float(row[k]) - mu; s2 = fmaf(d, d, s2); }
      const float r = rsqrtf(warp_sum(s2) / (float)width + a.eps);
      bf16* out = out_row(m);
      for (int k = lane; k < width; k += 32) {
        float v = (__bfloat162float(row[k]) - mu) * r * __ldg(gamma + k) + __ldg(beta + k);
        if (gelu) v = gelu_erf(v);
        out[k] = __float2bfloat16(v);
      }
    }
  };
  ln_rows(s_h, ldh, H4, a.g1, a.be1, true, [&](int m) { return s_h + (size_t)m * ldh; });
  __syncthreads();
  rows16_mma<IM>(s_h, ldh, M * ldh, a.W3, H4, C, H4, [&](int n, int m, float v0, float v1) {
    const float2 bias = __ldg(reinterpret_cast<const float2*>(a.b3 + n));
    *reinterpret_cast<uint32_t*>(s_o + (size_t)m * C + n) = pack_bf16x2(v0 + bias.x, v1 + bias.y);
  });
  __syncthreads();
  ln_rows(s_o, C, C, a.g4, a.be4, false, [&](int m) { return a.out + (size_t)bi[m >> 4] * a.out_bs + (size_t)(m & 15) * C; });
}

// Images per CTA: 2 once one image per CTA would need more than one wave of CTAs (one per SM: 128 registers x 512 threads), 1
// below that (the kernels are latency chains: a CTA with two images takes ~1.8x as long, so halving the CTA count only pays when
// it saves a wave; measured at Base b256 on one lane: 1.33 -> 1.14 ms per forward, 384 -> 512 downsample 123 -> 85 us).  LMV_META_IM overrides (1, 2, 4; read per call so that tests can switch it); a value whose
// shared memory does not fit falls back to the next smaller one.
int pick_im(int B) {
  const char* e = getenv("LMV_META_IM");
  const int forced = e ? atoi(e) : 0;
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  return B > device_sm_count() ? 2 : 1;
}

PerDeviceOnce g_chain_once[3], g_ds_once[3];
constexpr int kMetaSmemMax = 200 * 1024;

}  // namespace

bool dca_supported(int N, int C, int heads, int Mq) {
  return Mq == kDcaM && heads * 32 == C && C % 32 == 0 && C >= 32 && C <= 192 && heads * kDcaM <= 128 && N >= 1;
}

DcaGeom dca_geometry(int B, int N, int C, int heads) {
  DcaGeom g;
  g.B = B; g.N = N; g.C = C; g.heads = heads; g.R = heads * kDcaM;
  g.tiles = (N + kDcaTile - 1) / kDcaTile;
  const int nseg = (g.tiles + 5) / 6;                 // segments of <= 6 tiles: a function of N only (batch-invariant bits)
  g.seg_tiles = (g.tiles + nseg - 1) / nseg;
  g.segs = (g.tiles + g.seg_tiles - 1) / g.seg_tiles;
  g.dup = g.R <= 64 ? 1 : 0;
  g.ncopy = g.dup ? 2 : 1;
  g.parts = g.segs * g.ncopy;
  return g;
}

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

size_t dca_workspace_bytes(const DcaGeom& g) {
  const size_t rc = (size_t)g.B * g.R * g.C;
  return 3 * al256(rc * 2) + al256((size_t)g.B * 4 * g.R * 4) + al256((size_t)g.B * g.parts * g.R * 16) + al256((size_t)g.B * g.parts * g.R * g.C * 4);
}

DcaWs dca_workspace_carve(const DcaGeom& g, void* base) {
  uint8_t* p = static_cast<uint8_t*>(base);
  const size_t rc = (size_t)g.B * g.R * g.C;
  DcaWs w;
  w.kt = reinterpret_cast<bf16*>(p); p += al256(rc * 2);
  w.qt = reinterpret_cast<bf16*>(p); p += al256(rc * 2);
  w.vt = reinterpret_cast<bf16*>(p); p += al256(rc * 2);
  w.cst = reinterpret_cast<float*>(p); p += al256((size_t)g.B * 4 * g.R * 4);
  w.part_ml = reinterpret_cast<float4*>(p); p += al256((size_t)g.B * g.parts * g.R * 16);
  w.part_z = reinterpret_cast<float*>(p);
  return w;
}

static size_t pre_scratch_bytes(const MetaPreArgs& a, int im) {
  const int R = a.heads * M;
  return im * ((size_t)M * pitch(a.C) * 2 + (size_t)M * pitch(a.nc) * 2 + (size_t)R * a.C * 2);
}
static size_t post_scratch_bytes(const MetaPostArgs& a, int im) {
  const int R = a.heads * M;
  const size_t zbytes = std::max((size_t)R * pitch(a.C), (size_t)M * pitch(a.Hd)) * 2;
  return im * ((size_t)(a.parts + 1) * R * 4 + (size_t)M * pitch(a.C) * 2 + zbytes);
}
static int check_pre(const MetaPreArgs& a) {
  LMV_REQUIRE(a.c && a.Wc && a.bc && a.WxkT && a.bxk && a.ws.qt && a.ws.cst, "meta_pre: null pointer");
  LMV_REQUIRE((a.WxqT == nullptr) == (a.Wpx == nullptr), "meta_pre: WxqT and Wpx go together (x-branch)");
  LMV_REQUIRE(a.C % 32 == 0 && a.heads * 32 == a.C && a.nc % 8 == 0, "meta_pre: C must be heads * 32");
  return LMV_OK;
}
static int check_post(const MetaPostArgs& a) {
  LMV_REQUIRE(a.c && a.Wxv && a.bxv && a.Wp && a.bp && a.W1 && a.b1 && a.W2 && a.b2 && a.ws.part_ml && a.ws.part_z, "meta_post: null pointer");
  LMV_REQUIRE(a.C % 32 == 0 && a.heads * 32 == a.C && a.Hd % 32 == 0 && a.parts >= 1, "meta_post: shape");
  return LMV_OK;
}

// [post of block j] -> [pre of block j + 1] in one launch; either pointer may be null.  Both must belong to the same stage (same
// B, C and meta-token buffer): the pre body continues on the tokens the post body leaves in shared memory.
int meta_chain_run(const MetaPostArgs* post, const MetaPreArgs* pre, cudaStream_t s) {
  LMV_REQUIRE(post || pre, "meta_chain: nothing to run");
  int rc;
  if (post && (rc = check_post(*post))) return rc;
  if (pre && (rc = check_pre(*pre))) return rc;
  if (post && pre) LMV_REQUIRE(post->C == pre->C && post->B == pre->B && post->c == pre->c, "meta_chain: post and pre of different stages");
  const int C = post ? post->C : pre->C, B = post ? post->B : pre->B;
  auto smem_for = [&](int im) {
    return (size_t)im * M * C * 4 + std::max(post ? post_scratch_bytes(*post, im) : 0, pre ? pre_scratch_bytes(*pre, im) : 0);
  };
  int im = pick_im(B);
  while (im > 1 && smem_for(im) > (size_t)kMetaSmemMax) im >>= 1;
  const size_t smem = smem_for(im);
  LMV_REQUIRE(smem <= (size_t)kMetaSmemMax, "meta_chain: shared memory budget");
  const MetaPostArgs pa = post ? *post : MetaPostArgs{};
  const MetaPreArgs ra = pre ? *pre : MetaPreArgs{};
  const dim3 grid((B + im - 1) / im);
  auto go = [&](auto kern, int slot) -> int {
    LMV_CUDA_OK(g_chain_once[slot].run([kern] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetaSmemMax); }));
    LMV_CUDA_OK(launch_kernel(kern, grid, dim3(kThreads), smem, s, pa, ra, post ? 1 : 0, pre ? 1 : 0, C, B));
    LMV_CUDA_OK(cudaGetLastError());
    return LMV_OK;
  };
  return im == 4 ? go(meta_chain_kernel<4>, 2) : im == 2 ? go(meta_chain_kernel<2>, 1) : go(meta_chain_kernel<1>, 0);
}

int meta_pre_run(const MetaPreArgs& a, cudaStream_t s) { return meta_chain_run(nullptr, &a, s); }
int meta_post_run(const MetaPostArgs& a, cudaStream_t s) { return meta_chain_run(&a, nullptr, s); }

bool meta_downsample_supported(int Cp, int C) { return Cp % 32 == 0 && C % 8 == 0 && Cp >= 32 && Cp <= 512 && C <= 2048; }

int meta_downsample_run(const MetaDsArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.in && a.out && a.W0 && a.b0 && a.g1 && a.be1 && a.W3 && a.b3 && a.g4 && a.be4, "meta_downsample: null pointer");
  if (!meta_downsample_supported(a.Cp, a.C)) return fail(LMV_ERR_UNSUPPORTED, "meta_downsample: needs Cp % 32 == 0, C % 8 == 0");
  LMV_REQUIRE((reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && a.in_bs % 8 == 0, "meta_downsample: input rows must be 16-byte aligned");
  const int H4 = 4 * a.Cp;
  auto smem_for = [&](int im) { return (size_t)im * M * (pitch(a.Cp) + pitch(H4) + a.C) * 2; };
  int im = pick_im(a.B);
  while (im > 1 && smem_for(im) > (size_t)kMetaSmemMax) im >>= 1;
  const size_t smem = smem_for(im);
  LMV_REQUIRE(smem <= (size_t)kMetaSmemMax, "meta_downsample: shared memory budget");
  const dim3 grid((a.B + im - 1) / im);
  auto go = [&](auto kern, int slot) -> int {
    LMV_CUDA_OK(g_ds_once[slot].run([kern] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetaSmemMax); }));
    LMV_CUDA_OK(launch_kernel(kern, grid, dim3(kThreads), smem, s, a));
    LMV_CUDA_OK(cudaGetLastError());
    return LMV_OK;
  };
  return im == 4 ? go(meta_downsample_kernel<4>, 2) : im == 2 ? go(meta_downsample_kernel<2>, 1) : go(meta_downsample_kernel<1>, 0);
}

}  // namespace lmv
