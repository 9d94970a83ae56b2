// Meta-token side of the fused cross-attention blocks: everything that happens to the M = 16 meta tokens of an image in a
// 'C' block (CrossAttention, models/lemevit.py:477-486 inside forward_with_c :584-613) or a 'D' block (DualCrossAttention
// :252-302 inside forward_with_xc :542-582), in TWO kernels per block instead of ~8 launch-latency-bound ones:
//
//   meta_pre   cn = LN1(c);  [q2 | k2 | v2] = Wc cn + bc;  builds the per-image operands of the image-token kernel
//              (dca_fused.cu) with the image-side projections absorbed (kernels.h): Kt, Qt, Vt^T and their constants;
//   meta_post  merges the c-branch softmax partials of the image-token kernel into Zbar = sum_n softmax_n(Sc) xn[n], then
//              attn_c = Wv_h Zbar + bv;  c += proj_c(attn_c);  c += mlp(LN2(c))          (:563-564 / :600-601).
//
// 16 rows per image are far below a tensor-core tile, so these are CUDA-core kernels: one CTA per image, the 16 rows held in
// shared memory as fp32, one thread per output column (16 accumulators), weights streamed from L2 with 16-byte loads.  The
// work is ~6 (pre) + ~10 (post) C^2 MACs per row — 1-6 MFLOP per image.
#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int M = kDcaM;
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[m][n] = sum_k in[m][k] W[n][k]  for the 16 rows in shared memory; one thread per output column n, fn(n, acc[16]) consumes
// the result.  W rows are read with 16-byte loads (K % 8 == 0), `in` rows with broadcast float4 loads (ld = row pitch in floats).
template <typename Fn>
__device__ __forceinline__ void rows16_linear(const float* __restrict__ in, int ld, int K, const bf16* __restrict__ W, int ldw,
                                               int n_begin, int n_end, Fn fn) {
  for (int n = n_begin + (int)threadIdx.x; n < n_end; n += kThreads) {
    float acc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = 0.f;
    const uint4* wrow = reinterpret_cast<const uint4*>(W + (size_t)n * ldw);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint4 u = __ldg(wrow + k8);
      const float2 w0 = unpack_bf16x2(u.x), w1 = unpack_bf16x2(u.y), w2 = unpack_bf16x2(u.z), w3 = unpack_bf16x2(u.w);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float4 a = *reinterpret_cast<const float4*>(in + m * ld + k8 * 8);
        const float4 b = *reinterpret_cast<const float4*>(in + m * ld + k8 * 8 + 4);
        acc[m] = fmaf(a.x, w0.x, fmaf(a.y, w0.y, fmaf(a.z, w1.x, fmaf(a.w, w1.y, acc[m]))));
        acc[m] = fmaf(b.x, w2.x, fmaf(b.y, w2.y, fmaf(b.z, w3.x, fmaf(b.w, w3.y, acc[m]))));
      }
    }
    fn(n, acc);
  }
}

// LayerNorm without affine over the 16 rows (one warp per two rows): out[m][:] = (in[m][:] - mu) rsqrt(var + eps)
__device__ __forceinline__ void rows16_layernorm(const float* in, float* out, int ld, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += kThreads / 32) {
    float s1 = 0.f;
    for (int k = lane; k < C; k += 32) s1 += in[m * ld + k];
    const float mu = warp_sum(s1) / (float)C;
    float s2 = 0.f;
    for (int k = lane; k < C; k += 32) { const float d = in[m * ld + k] - mu; s2 = fmaf(d, d, s2); }
    const float r = rsqrtf(warp_sum(s2) / (float)C + eps);
    for (int k = lane; k < C; k += 32) out[m * ld + k] = (in[m * ld + k] - mu) * r;
  }
}

// ------------------------------------------------------------------------------------------------
// meta_pre
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
meta_pre_kernel(MetaPreArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = a.C, R = a.heads * M, nc = a.nc, b = blockIdx.x;
  float* s_c = reinterpret_cast<float*>(smem_raw);         // [M][C]   raw, then normalised meta tokens
  float* s_p = s_c + M * C;                                 // [M][nc]  projection of LN1(c)
  bf16* s_t = reinterpret_cast<bf16*>(s_p + M * nc);        // [R][C]   staging of Kt / Qt (bf16, as stored)
  pdl_launch_dependents();
  pdl_wait();
  const bf16* cb = a.c + (size_t)b * M * C;
  for (int i = threadIdx.x; i < M * C; i += kThreads) s_c[i] = __bfloat162float(cb[i]);
  __syncthreads();
  rows16_layernorm(s_c, s_c, C, C, a.eps);     // in place: every element is read and written by the same lane
  __syncthreads();
  rows16_linear(s_c, C, C, a.Wc, C, 0, nc, [&](int n, const float (&acc)[M]) {
    const float bias = __ldg(a.bc + n);
#pragma unroll
    for (int m = 0; m < M; ++m) s_p[m * nc + n] = acc[m] + bias;
  });
  __syncthreads();

  // T[(h,m)][j] = scale * sum_d vec[m][32h + d] Wx[32h + d][j]   (Wx: image-side projection rows, [C, C] row-major, j contiguous)
  // + its row sums (of the bf16-rounded values) and the bias constant scale * sum_d bx[32h + d] vec[m][32h + d]
  auto absorb = [&](const float* vec /* s_p + offset */, const bf16* Wx, const float* bx, float scale, bf16* out_g, float* sum_g,
                    float* cst_g) {
    for (int j2 = threadIdx.x; j2 < C / 2; j2 += kThreads) {
      for (int h = 0; h < a.heads; ++h) {
        float2 acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = make_float2(0.f, 0.f);
        for (int d = 0; d < 32; ++d) {
          const float2 w = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(Wx + (size_t)(32 * h + d) * C) + j2));
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const float v = vec[m * nc + 32 * h + d];
            acc[m].x = fmaf(v, w.x, acc[m].x);
            acc[m].y = fmaf(v, w.y, acc[m].y);
          }
        }
#pragma unroll
        for (int m = 0; m < M; ++m)
          *reinterpret_cast<uint32_t*>(s_t + (size_t)(h * M + m) * C + 2 * j2) = pack_bf16x2(acc[m].x * scale, acc[m].y * scale);
      }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < R; r += kThreads / 32) {
      float s = 0.f;
      for (int k = lane; k < C; k += 32) s += __bfloat162float(s_t[(size_t)r * C + k]);
      s = warp_sum(s);
      const int h = r / M, m = r - h * M;
      const float kc = vec[m * nc + 32 * h + lane] * __ldg(bx + 32 * h + lane);
      const float k2 = warp_sum(kc) * scale;
      if (lane == 0) { sum_g[r] = s; cst_g[r] = k2; }
    }
    const uint4* src = reinterpret_cast<const uint4*>(s_t);
    uint4* dst = reinterpret_cast<uint4*>(out_g);
    for (int i = threadIdx.x; i < R * C / 8; i += kThreads) dst[i] = src[i];
    __syncthreads();
  };
  float* cst = a.ws.cst + (size_t)b * 4 * R;
  if (a.Wxq)   // 'D': x-branch keys in token space
    absorb(s_p + a.k_off, a.Wxq, a.bxq, a.scale_x * kLog2e, a.ws.kt + (size_t)b * R * C, cst, cst + R);
  absorb(s_p + a.q_off, a.Wxk, a.bxk, a.scale_c * kLog2e, a.ws.qt + (size_t)b * R * C, cst + 2 * R, cst + 3 * R);
  if (a.Wpx) {
    // Vt^T[j][(h,m)] = sum_d Wpx[j][32h + d] v2[m][32h + d]: one thread per output channel j, its row of Vt^T is contiguous
    const float* v2 = s_p + a.v_off;
    bf16* vt = a.ws.vt + (size_t)b * C * R;
    for (int j = threadIdx.x; j < C; j += kThreads) {
      for (int h = 0; h < a.heads; ++h) {
        float w[32];
        const uint4* wr = reinterpret_cast<const uint4*>(a.Wpx + (size_t)j * C + 32 * h);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = __ldg(wr + q);
          const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
          w[8 * q + 0] = f0.x; w[8 * q + 1] = f0.y; w[8 * q + 2] = f1.x; w[8 * q + 3] = f1.y;
          w[8 * q + 4] = f2.x; w[8 * q + 5] = f2.y; w[8 * q + 6] = f3.x; w[8 * q + 7] = f3.y;
        }
        uint32_t pk[M / 2];
#pragma unroll
        for (int m = 0; m < M; m += 2) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            s0 = fmaf(w[d], v2[m * nc + 32 * h + d], s0);
            s1 = fmaf(w[d], v2[(m + 1) * nc + 32 * h + d], s1);
          }
          pk[m / 2] = pack_bf16x2(s0, s1);
        }
        uint4* dst = reinterpret_cast<uint4*>(vt + (size_t)j * R + h * M);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// meta_post
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
meta_post_kernel(MetaPostArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = a.C, R = a.heads * M, Hd = a.Hd, b = blockIdx.x, P = a.parts;
  float* s_c = reinterpret_cast<float*>(smem_raw);   // [M][C]  meta tokens (residual stream, fp32)
  float* s_a = s_c + M * C;                           // [M][C]  attn_c, later LN2(c)
  float* s_w = s_a + M * C;                           // [P][R]  merge weights, [R] 1 / l
  float* s_z = s_w + (P + 1) * R;                     // [R][C]  Zbar; reused as the MLP hidden activation [M][Hd]
  pdl_launch_dependents();
  pdl_wait();
  bf16* cb = a.c + (size_t)b * M * C;
  for (int i = threadIdx.x; i < M * C; i += kThreads) s_c[i] = __bfloat162float(cb[i]);
  // ---- merge the segment partials (fixed order: deterministic, independent of batch size and position)
  const float4* ml = a.ws.part_ml + (size_t)b * P * R;
  for (int r = threadIdx.x; r < R; r += kThreads) {
    float mx = -INFINITY;
    for (int p = 0; p < P; ++p) mx = fmaxf(mx, ml[p * R + r].x);
    float l = 0.f;
    for (int p = 0; p < P; ++p) {
      const float4 v = ml[p * R + r];
      const float w = (v.x == -INFINITY) ? 0.f : exp2f(v.x - mx);
      s_w[p * R + r] = w;
      l = fmaf(v.y, w, l);
    }
    s_w[P * R + r] = 1.f / l;
  }
  __syncthreads();
  const float* pz = a.ws.part_z + (size_t)b * P * R * C;
  for (int i = threadIdx.x; i < R * C / 4; i += kThreads) {
    const int r = (i * 4) / C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < P; ++p) {
      const float w = s_w[p * R + r];
      if (w == 0.f) continue;      // an empty partial (no valid token in that copy's columns) may hold anything
      const float4 z = __ldg(reinterpret_cast<const float4*>(pz + (size_t)p * R * C) + i);
      const float t = ml[p * R + r].z;    // sum_n p'_n mu_n: Zbar = sum p' (xt - mu)
      acc.x = fmaf(w, z.x - t, acc.x); acc.y = fmaf(w, z.y - t, acc.y);
      acc.z = fmaf(w, z.z - t, acc.z); acc.w = fmaf(w, z.w - t, acc.w);
    }
    const float il = s_w[P * R + r];
    reinterpret_cast<float4*>(s_z)[i] = make_float4(acc.x * il, acc.y * il, acc.z * il, acc.w * il);
  }
  __syncthreads();
  // ---- attn_c[m][32h + d] = Wv[32h + d][:] . Zbar[(h,m)][:] + bv   (a warp's 32 output channels share the head: broadcast reads)
  for (int n = threadIdx.x; n < C; n += kThreads) {
    const int h = n >> 5;
    float acc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = 0.f;
    const uint4* wrow = reinterpret_cast<const uint4*>(a.Wxv + (size_t)n * C);
    const float* z = s_z + (size_t)h * M * C;
    for (int k8 = 0; k8 < C / 8; ++k8) {
      const uint4 u = __ldg(wrow + k8);
      const float2 w0 = unpack_bf16x2(u.x), w1 = unpack_bf16x2(u.y), w2 = unpack_bf16x2(u.z), w3 = unpack_bf16x2(u.w);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float4 p = *reinterpret_cast<const float4*>(z + m * C + k8 * 8);
        const float4 q = *reinterpret_cast<const float4*>(z + m * C + k8 * 8 + 4);
        acc[m] = fmaf(p.x, w0.x, fmaf(p.y, w0.y, fmaf(p.z, w1.x, fmaf(p.w, w1.y, acc[m]))));
        acc[m] = fmaf(q.x, w2.x, fmaf(q.y, w2.y, fmaf(q.z, w3.x, fmaf(q.w, w3.y, acc[m]))));
      }
    }
    const float bias = __ldg(a.bxv + n);
#pragma unroll
    for (int m = 0; m < M; ++m) s_a[m * C + n] = acc[m] + bias;
  }
  __syncthreads();
  // ---- c += proj(attn_c)
  rows16_linear(s_a, C, C, a.Wp, C, 0, C, [&](int n, const float (&acc)[M]) {
    const float bias = __ldg(a.bp + n);
#pragma unroll
    for (int m = 0; m < M; ++m) s_c[m * C + n] += acc[m] + bias;
  });
  __syncthreads();
  // ---- c += mlp(LN2(c))   (LN2 affine folded into W1 / b1)
  rows16_layernorm(s_c, s_a, C, C, a.eps);
  __syncthreads();
  float* s_h = s_z;
  rows16_linear(s_a, C, C, a.W1, C, 0, Hd, [&](int n, const float (&acc)[M]) {
    const float bias = __ldg(a.b1 + n);
#pragma unroll
    for (int m = 0; m < M; ++m) s_h[m * Hd + n] = gelu_erf(acc[m] + bias);
  });
  __syncthreads();
  rows16_linear(s_h, Hd, Hd, a.W2, Hd, 0, C, [&](int n, const float (&acc)[M]) {
    const float bias = __ldg(a.b2 + n);
#pragma unroll
    for (int m = 0; m < M; ++m) cb[m * C + n] = __float2bfloat16(s_c[m * C + n] + acc[m] + bias);
  });
}

PerDeviceOnce g_pre_once, g_post_once;
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

int meta_pre_run(const MetaPreArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.c && a.Wc && a.bc && a.Wxk && a.bxk && a.ws.qt && a.ws.cst, "meta_pre: null pointer");
  LMV_REQUIRE((a.Wxq == nullptr) == (a.Wpx == nullptr), "meta_pre: Wxq and Wpx go together (x-branch)");
  LMV_REQUIRE(a.C % 32 == 0 && a.heads * 32 == a.C && a.nc % 8 == 0, "meta_pre: C must be heads * 32");
  const int R = a.heads * M;
  const size_t smem = (size_t)M * a.C * 4 + (size_t)M * a.nc * 4 + (size_t)R * a.C * 2;
  LMV_REQUIRE(smem <= (size_t)kMetaSmemMax, "meta_pre: shared memory budget");
  LMV_CUDA_OK(g_pre_once.run([] { return cudaFuncSetAttribute(meta_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetaSmemMax); }));
  LMV_CUDA_OK(launch_kernel(meta_pre_kernel, dim3(a.B), dim3(kThreads), smem, s, a));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int meta_post_run(const MetaPostArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.c && a.Wxv && a.bxv && a.Wp && a.bp && a.W1 && a.b1 && a.W2 && a.b2 && a.ws.part_ml && a.ws.part_z, "meta_post: null pointer");
  LMV_REQUIRE(a.C % 32 == 0 && a.heads * 32 == a.C && a.Hd % 8 == 0 && a.parts >= 1, "meta_post: shape");
  const int R = a.heads * M;
  const size_t zbytes = std::max((size_t)R * a.C, (size_t)M * a.Hd) * 4;
  const size_t smem = (size_t)2 * M * a.C * 4 + (size_t)(a.parts + 1) * R * 4 + zbytes;
  LMV_REQUIRE(smem <= (size_t)kMetaSmemMax, "meta_post: shared memory budget");
  LMV_CUDA_OK(g_post_once.run([] { return cudaFuncSetAttribute(meta_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetaSmemMax); }));
  LMV_CUDA_OK(launch_kernel(meta_post_kernel, dim3(a.B), dim3(kThreads), smem, s, a));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
