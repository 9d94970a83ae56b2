// Memory-bound kernels of the LeMeViT forward: positional depthwise conv + LayerNorm, LayerNorm,
// im2col for the strided 3x3 convolutions (which then run on the tcgen05 GEMM), the classification
// tail and the NCHW export of the backbone feature maps.  All activations are token-major
// [B, T, C] bf16 (NHWC); the reference's NCHW<->token rearranges (models/lemevit.py:548,579) vanish.
#include <algorithm>
#include <type_traits>

#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// x + dwconv3x3(x)  ->  LayerNorm (no affine; gamma/beta are folded into the consuming linear)
// reference: LeMeBlock pos_embed (models/lemevit.py:510 at :546,589,619) + norm1 (:513)
// one warp per token; lane l owns channel pairs l, l+32, ...  (C <= 512)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPairs = 8;

__global__ void __launch_bounds__(256)
posembed_ln_kernel(PosLnArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= (long long)a.B * a.T) return;
  const int b = (int)(row / a.T), t = (int)(row % a.T);
  const int C = a.C, npairs = C >> 1;
  const bf16* img = a.tokens + (long long)b * a.T * C;
  float2 val[kMaxPairs];
  const bool conv = (a.dw_w != nullptr) && (t < a.H * a.W);
  if (conv) {
    const int y = t / a.W, x = t % a.W;
#pragma unroll
    for (int i = 0; i < kMaxPairs; ++i) {
      const int p = lane + 32 * i;
      val[i] = (p < npairs) ? __ldg(reinterpret_cast<const float2*>(a.dw_b) + p) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= a.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= a.W) continue;
        const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(img + (long long)(yy * a.W + xx) * C);
        const float2* wt = reinterpret_cast<const float2*>(a.dw_w + (ky * 3 + kx) * C);
#pragma unroll
        for (int i = 0; i < kMaxPairs; ++i) {
          const int p = lane + 32 * i;
          if (p < npairs) {
            const float2 v = __bfloat1622float2(src[p]);
            const float2 w = __ldg(wt + p);
            val[i].x = fmaf(v.x, w.x, val[i].x);
            val[i].y = fmaf(v.y, w.y, val[i].y);
          }
        }
      }
    }
  } else {
    const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(img + (long long)t * C);
#pragma unroll
    for (int i = 0; i < kMaxPairs; ++i) {
      const int p = lane + 32 * i;
      val[i] = (p < npairs) ? __bfloat1622float2(src[p]) : make_float2(0.f, 0.f);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPairs; ++i) s += val[i].x + val[i].y;   // out-of-range pairs hold zeros
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPairs; ++i) {
    if (lane + 32 * i < npairs) {
      const float dx = val[i].x - mean, dy = val[i].y - mean;
      ss += dx * dx + dy * dy;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + a.eps);
  const long long off = row * C;
#pragma unroll
  for (int i = 0; i < kMaxPairs; ++i) {
    const int p = lane + 32 * i;
    if (p < npairs) {
      if (a.resid_out) reinterpret_cast<__nv_bfloat162*>(a.resid_out + off)[p] = __floats2bfloat162_rn(val[i].x, val[i].y);
      if (a.norm_out)
        reinterpret_cast<__nv_bfloat162*>(a.norm_out + off)[p] =
            __floats2bfloat162_rn((val[i].x - mean) * rstd, (val[i].y - mean) * rstd);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2 of the same op for C % 8 == 0: G lanes per token, each lane owns 8 consecutive channels (one 16-byte vector)
// per ITER; consecutive lanes -> consecutive 16-byte chunks (fully coalesced), the 9 taps x 8 channels of depthwise
// weights live in registers (ITERS == 1) and are reused over `iters` tokens; a block covers a compact run of
// ~4 image rows so the vertical taps hit L1.  Emits the LayerNorm statistics of the stored (bf16-rounded) row
// so that the consuming GEMM can fold norm1 into its epilogue (gemm.cu).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int G, int ITERS>
__global__ void __launch_bounds__(256)
posembed_ln_v2_kernel(PosLnArgs a, int iters) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int SLOTS = 256 / G;
  const int slot = threadIdx.x / G, gl = threadIdx.x % G;
  const int C = a.C, V = C >> 3, HW = a.H * a.W;
  const long long rows = (long long)a.B * a.T;
  const long long base = (long long)blockIdx.x * (SLOTS * iters);
  const bool has_conv = a.dw_w != nullptr;
  float w[ITERS == 1 ? 9 : 1][8], bias[8];
  if (ITERS == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) bias[j] = 0.f;
    if (has_conv && gl < V) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.dw_w + tap * C + gl * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.dw_w + tap * C + gl * 8) + 1);
        w[tap][0] = w0.x; w[tap][1] = w0.y; w[tap][2] = w0.z; w[tap][3] = w0.w;
        w[tap][4] = w1.x; w[tap][5] = w1.y; w[tap][6] = w1.z; w[tap][7] = w1.w;
      }
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.dw_b + gl * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.dw_b + gl * 8) + 1);
      bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w; bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
  }
  for (int k = 0; k < iters; ++k) {
    const long long row = base + (long long)k * SLOTS + slot;
    const bool active = row < rows;
    const int b = active ? (int)(row / a.T) : 0, t = active ? (int)(row % a.T) : 0;
    const bf16* img = a.tokens + (long long)b * a.T * C;
    const bool conv = has_conv && t < HW;
    const int y = t / a.W, x = t - y * a.W;
    uint4 packed[ITERS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int v = gl + G * it;
      float val[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) val[j] = 0.f;
      if (active && v < V) {
        if (conv) {
          if (ITERS == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] = bias[j];
          } else {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.dw_b + v * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.dw_b + v * 8) + 1);
            val[0] = b0.x; val[1] = b0.y; val[2] = b0.z; val[3] = b0.w; val[4] = b1.x; val[5] = b1.y; val[6] = b1.z; val[7] = b1.w;
          }
          // issue all 9 tap loads before any arithmetic (memory-level parallelism); out-of-image taps read the
          // clamped address and are replaced by zeros
          uint4 tapv[9];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int yy = y + ky - 1, xx = x + kx - 1;
              const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
              const int yc = min(max(yy, 0), a.H - 1), xc = min(max(xx, 0), a.W - 1);
              const uint4 ld = __ldg(reinterpret_cast<const uint4*>(img + (long long)(yc * a.W + xc) * C) + v);
              tapv[ky * 3 + kx] = ok ? ld : make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            float f[8];
            unpack8(tapv[tap], f);
            if (ITERS == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) val[j] = fmaf(f[j], w[tap][j], val[j]);
            } else {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.dw_w + tap * C + v * 8));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.dw_w + tap * C + v * 8) + 1);
              val[0] = fmaf(f[0], w0.x, val[0]); val[1] = fmaf(f[1], w0.y, val[1]);
              val[2] = fmaf(f[2], w0.z, val[2]); val[3] = fmaf(f[3], w0.w, val[3]);
              val[4] = fmaf(f[4], w1.x, val[4]); val[5] = fmaf(f[5], w1.y, val[5]);
              val[6] = fmaf(f[6], w1.z, val[6]); val[7] = fmaf(f[7], w1.w, val[7]);
            }
          }
        } else {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(img + (long long)t * C) + v);
          unpack8(u, val);
        }
      }
      // round to the stored precision first: the statistics must describe exactly what the consumer reads
      uint4 pk;
      pk.x = pack_bf16x2(val[0], val[1]); pk.y = pack_bf16x2(val[2], val[3]);
      pk.z = pack_bf16x2(val[4], val[5]); pk.w = pack_bf16x2(val[6], val[7]);
      packed[it] = pk;
      unpack8(pk, val);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += val[j]; s2 = fmaf(val[j], val[j], s2); }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (!active) continue;
    if (a.stats_out && gl == 0) *reinterpret_cast<float2*>(a.stats_out + 2 * row) = make_float2(s1, s2);
    float mean = 0.f, rstd = 0.f;
    if (a.norm_out) {
      mean = s1 / (float)C;
      // centred second pass over the registers (exactly the reference's biased variance of the rounded row)
      float ss = 0.f;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        if (gl + G * it < V) {
          float f[8];
          unpack8(packed[it], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      rstd = rsqrtf(ss / (float)C + a.eps);
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int v = gl + G * it;
      if (v >= V) continue;
      if (a.resid_out) reinterpret_cast<uint4*>(a.resid_out + row * C)[v] = packed[it];
      if (a.norm_out) {
        float f[8];
        unpack8(packed[it], f);
        uint4 pk;
        pk.x = pack_bf16x2((f[0] - mean) * rstd, (f[1] - mean) * rstd); pk.y = pack_bf16x2((f[2] - mean) * rstd, (f[3] - mean) * rstd);
        pk.z = pack_bf16x2((f[4] - mean) * rstd, (f[5] - mean) * rstd); pk.w = pack_bf16x2((f[6] - mean) * rstd, (f[7] - mean) * rstd);
        reinterpret_cast<uint4*>(a.norm_out + row * C)[v] = pk;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic LayerNorm over rows (any even C): warp per row, three L1-resident passes.
// reference: nn.LayerNorm call sites models/lemevit.py:513,525,731,734,774
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_kernel(LnArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= a.R) return;
  const int npairs = a.C >> 1;
  const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(a.in + row * a.C);
  float s = 0.f;
  for (int p = lane; p < npairs; p += 32) {
    const float2 v = __bfloat1622float2(src[p]);
    s += v.x + v.y;
  }
  const float mean = warp_sum(s) / (float)a.C;
  float ss = 0.f;
  for (int p = lane; p < npairs; p += 32) {
    const float2 v = __bfloat1622float2(src[p]);
    const float dx = v.x - mean, dy = v.y - mean;
    ss += dx * dx + dy * dy;
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)a.C + a.eps);
  long long orow = row;
  if (a.grp_rows > 0) orow = (row / a.grp_rows) * a.grp_stride + a.grp_off + (row % a.grp_rows);
  __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(a.out + orow * a.C);
  for (int p = lane; p < npairs; p += 32) {
    float2 v = __bfloat1622float2(src[p]);
    v.x = (v.x - mean) * rstd;
    v.y = (v.y - mean) * rstd;
    if (a.gamma) {
      const float2 g = __ldg(reinterpret_cast<const float2*>(a.gamma) + p);
      const float2 be = __ldg(reinterpret_cast<const float2*>(a.beta) + p);
      v.x = fmaf(v.x, g.x, be.x);
      v.y = fmaf(v.y, g.y, be.y);
    }
    if (a.act_gelu) {
      v.x = gelu_erf(v.x);
      v.y = gelu_erf(v.y);
    }
    dst[p] = __floats2bfloat162_rn(v.x, v.y);
  }
}

// ------------------------------------------------------------------------------------------------
// stem im2col: x NCHW (f32|bf16) -> patches [B*Ho*Wo, Kp] bf16, k = ci*9 + ky*3 + kx, zero padded to Kp
// reference: first stem conv nn.Conv2d(in_chans, C0/2, 3, 2, 1) (models/lemevit.py:699)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const T* __restrict__ x, bf16* __restrict__ out, int B, int Cin, int H, int W, int Ho, int Wo,
                   int Kp) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per (output pixel, group of 8 k values)
  const int groups = Kp >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Ho * Wo * groups;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = g * 8 + j;
    float val = 0.f;
    if (k < Cin * 9) {
      const int ci = k / 9, r = k % 9, ky = r / 3, kx = r % 3;
      const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = ld_as_float<T>(x + (((long long)b * Cin + ci) * H + iy) * W + ix);
    }
    v[j] = val;
  }
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out + pix * Kp)[g] = u;
}

// ------------------------------------------------------------------------------------------------
// first stem convolution, direct: x NCHW (f32|bf16, 3 channels) -> conv3x3/s2/p1 (+ folded BatchNorm) -> GELU -> token-major
// [B, Ho*Wo, C1] bf16.  reference: nn.Conv2d(in_chans, C0/2, 3, 2, 1) + BatchNorm2d + GELU (models/lemevit.py:699-701).
// K = 27 is far too small for a tensor-core tile and the im2col detour writes 2.7x the output volume; here one thread owns two
// output pixels: their 2 x 27 inputs sit in registers, the folded weights are broadcast from shared memory ([k][C1] fp32), and the C1
// results leave as one contiguous 2*C1-byte run (neighbouring threads -> neighbouring runs).
// ------------------------------------------------------------------------------------------------

struct StemNorm { float mean[3], std[3]; };

// LAYOUT 0: x[B, 3, H, W] of T (float | bf16 | uint8_t); 1: x[B, H, W, 3] uint8_t.  8-bit pixels go through a 3 x 256-entry table of
// bf16((v - mean[c]) / std[c]) built once per CTA with IEEE subtraction / division: the same bits as the torch expression
// x.float().sub_(mean).div_(std).to(bfloat16), without a cvt / sub / div per tap.
template <typename T, int C1, int LAYOUT>
__global__ void __launch_bounds__(128)
stem_conv1_kernel(const T* __restrict__ x, const bf16* __restrict__ w /*[C1][Kp], k = ci*9 + ky*3 + kx*/, const float* __restrict__ bias,
                  bf16* __restrict__ out, int B, int H, int W, int Ho, int Wo, int Kp, StemNorm nrm) {
  constexpr bool kU8 = sizeof(T) == 1;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 sw[27 * (C1 / 4)];     // [k][C1]
  __shared__ float4 sb[C1 / 4];
  __shared__ float lut[kU8 ? 3 * 256 : 1];
  for (int i = threadIdx.x; i < 27 * C1; i += blockDim.x) {
    const int k = i / C1, c = i - k * C1;
    reinterpret_cast<float*>(sw)[i] = __bfloat162float(w[c * Kp + k]);
  }
  for (int i = threadIdx.x; i < C1; i += blockDim.x) reinterpret_cast<float*>(sb)[i] = bias[i];
  if (kU8)
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
      const int c = i >> 8;
      lut[i] = __bfloat162float(__float2bfloat16(__fdiv_rn(__fsub_rn((float)(i & 255), nrm.mean[c]), nrm.std[c])));
    }
  __syncthreads();
  // two output pixels per thread: every broadcast weight vector read from shared memory feeds 8 FMAs instead of 4
  const long long total = (long long)B * Ho * Wo;
  const long long pix0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (pix0 >= total) return;
  const bool two = pix0 + 1 < total;
  float in[2][27];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const long long pix = pix0 + ((q && two) ? 1 : 0);
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
          float v = 0.f;     // the zero padding pads the NORMALISED image
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            if constexpr (kU8) {
              const long long idx = LAYOUT == 1 ? (((long long)b * H + iy) * W + ix) * 3 + ci : (((long long)b * 3 + ci) * H + iy) * W + ix;
              v = lut[ci * 256 + (int)__ldg(reinterpret_cast<const unsigned char*>(x) + idx)];
            } else {
              v = ld_as_float<T>(x + (((long long)b * 3 + ci) * H + iy) * W + ix);
            }
          }
          in[q][ci * 9 + ky * 3 + kx] = v;
        }
  }
  float4 acc[2][C1 / 4];
#pragma unroll
  for (int c = 0; c < C1 / 4; ++c) { acc[0][c] = sb[c]; acc[1][c] = sb[c]; }
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float v0 = in[0][k], v1 = in[1][k];
#pragma unroll
    for (int c = 0; c < C1 / 4; ++c) {
      const float4 wv = sw[k * (C1 / 4) + c];     // same address in every lane: broadcast
      float2 a;
      a = ffma2(make_float2(v0, v0), make_float2(wv.x, wv.y), make_float2(acc[0][c].x, acc[0][c].y)); acc[0][c].x = a.x; acc[0][c].y = a.y;
      a = ffma2(make_float2(v0, v0), make_float2(wv.z, wv.w), make_float2(acc[0][c].z, acc[0][c].w)); acc[0][c].z = a.x; acc[0][c].w = a.y;
      a = ffma2(make_float2(v1, v1), make_float2(wv.x, wv.y), make_float2(acc[1][c].x, acc[1][c].y)); acc[1][c].x = a.x; acc[1][c].y = a.y;
      a = ffma2(make_float2(v1, v1), make_float2(wv.z, wv.w), make_float2(acc[1][c].z, acc[1][c].w)); acc[1][c].z = a.x; acc[1][c].w = a.y;
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q && !two) break;
    // lane = pixel stores (C1 * 2 = 64 or 96 contiguous bytes each): the LSU pays per cache line touched, so one full 32-byte sector
    // per lane and instruction where the output is 32-byte aligned
    bf16* dst = out + (pix0 + q) * C1;
    const bool al32 = (reinterpret_cast<uintptr_t>(out) & 31) == 0;
#pragma unroll
    for (int c = 0; c < C1 / 16; ++c) {
      uint32_t w[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // packed GELU (FFMA2 / FMUL2 + two MUFU.TANH per pair): this kernel is issue-bound, and the activation is a third of its instructions
        const float4 a0 = acc[q][4 * c + 2 * h], a1 = acc[q][4 * c + 2 * h + 1];
        const float2 g0 = gelu_fast2(make_float2(a0.x, a0.y)), g1 = gelu_fast2(make_float2(a0.z, a0.w));
        const float2 g2 = gelu_fast2(make_float2(a1.x, a1.y)), g3 = gelu_fast2(make_float2(a1.z, a1.w));
        w[4 * h] = pack_bf16x2(g0.x, g0.y);
        w[4 * h + 1] = pack_bf16x2(g1.x, g1.y);
        w[4 * h + 2] = pack_bf16x2(g2.x, g2.y);
        w[4 * h + 3] = pack_bf16x2(g3.x, g3.y);
      }
      if (al32) {
        st_global_256(dst + 16 * c, w);
      } else {
        reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// first stem convolution on the tensor cores: the same op as stem_conv1_kernel (conv3x3 / s2 / p1 + folded BN + GELU,
// models/lemevit.py:699-701), as an implicit GEMM with K = 27 -> 32.  The CUDA-core kernel above spends 2/3 of its issue slots on
// the 27 x C1 FMAs of a pixel; here they are two tcgen05.mma (M = 128 pixels, N = C1, K = 16) per tile.
//   warps 0..3  thread = pixel, for the gather AND the epilogue, software-pipelined: the 27 loads of tile i + 1 (NCHW f32 | bf16 |
//               8-bit pixels) are issued, then the epilogue of tile i runs (tcgen05.ld of the thread's TMEM lane -> + bias ->
//               GELU -> bf16 -> 256-bit stores) and hides their latency — the gather is a chain of cold image rows — and only then
//               are they packed into one 64-byte K-major row of the A tile (64B swizzle: 16-byte slot j of row r sits at
//               j ^ ((r >> 1) & 3); 8-bit pixels go through the normalisation table here) and handed to the issuer
//   warp 4      issues the two MMAs of a tile into one of two 64-column TMEM accumulators; the weights [C1][32] bf16 (pack.py's
//               layout is already K-major) sit in shared memory for the whole kernel
// Persistent CTAs of 160 threads, three per SM (128 columns of TMEM each).
// ------------------------------------------------------------------------------------------------
constexpr int kStemTcThreads = 5 * 32;

template <typename T, int C1, int LAYOUT>
__global__ void __launch_bounds__(kStemTcThreads, 3)
stem_conv1_tc_kernel(const T* __restrict__ x, const bf16* __restrict__ w /*[C1][32]*/, const float* __restrict__ bias, bf16* __restrict__ out,
                     int B, int H, int W, int Ho, int Wo, StemNorm nrm) {
  constexpr bool kU8 = sizeof(T) == 1;
  using Raw = typename std::conditional<kU8, uint8_t, float>::type;      // what a tap is held as between its load and its use
  __shared__ __align__(1024) uint8_t sA[2][128 * 64];
  __shared__ __align__(1024) uint8_t sW[C1 * 64];
  __shared__ __align__(16) float sBias[C1];
  __shared__ float lut[kU8 ? 3 * 256 : 1];
  __shared__ uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 4);       // one elected lane per pixel warp
      mbar_init(&a_empty[i], 1);      // tcgen05.commit
      mbar_init(&acc_full[i], 1);     // tcgen05.commit
      mbar_init(&acc_empty[i], 4);    // one elected lane per pixel warp
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_slot, 128);
    tmem_relinquish();
  }
  // constants of the plan (not outputs of the previous kernel): weights, bias, normalisation table
  for (int i = threadIdx.x; i < C1 * 4; i += blockDim.x) {
    const int n = i >> 2, j = i & 3;
    *reinterpret_cast<uint4*>(sW + n * 64 + ((j ^ ((n >> 1) & 3)) << 4)) = __ldg(reinterpret_cast<const uint4*>(w + n * 32) + j);
  }
  for (int i = threadIdx.x; i < C1; i += blockDim.x) sBias[i] = bias[i];
  if (kU8)
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
      const int c = i >> 8;
      lut[i] = __bfloat162float(__float2bfloat16(__fdiv_rn(__fsub_rn((float)(i & 255), nrm.mean[c]), nrm.std[c])));
    }
  fence_proxy_async_smem();     // sW is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  const long long total = (long long)B * Ho * Wo;
  const int num_tiles = (int)((total + 127) / 128);

  if (warp < 4) {
    const int r = threadIdx.x;       // row of the tile = TMEM lane
    const int HoWo = Ho * Wo, npix = (int)total;      // (< 2^31 output pixels: checked by the launcher)
    const bool al32 = (reinterpret_cast<uintptr_t>(out) & 31) == 0;
    // raw taps of a tile's pixel: loads only, nothing here waits for them (out-of-image taps: `zero`, the padding of the
    // NORMALISED image; for 8-bit input a 28th table entry would do, a flag bit per tap is cheaper)
    auto load_raw = [&](int t, Raw (&raw)[27], uint32_t& inside) {
      const int pix = t * 128 + r;
      inside = 0u;
#pragma unroll
      for (int i = 0; i < 27; ++i) raw[i] = Raw(0);
      if (t >= num_tiles || pix >= npix) return;
      const int b = pix / HoWo, rem = pix - b * HoWo, oy = rem / Wo, ox = rem - oy * Wo;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1, k = ci * 9 + ky * 3 + kx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
              inside |= 1u << k;
              if constexpr (kU8) {
                const long long idx = LAYOUT == 1 ? (((long long)b * H + iy) * W + ix) * 3 + ci : (((long long)b * 3 + ci) * H + iy) * W + ix;
                raw[k] = __ldg(reinterpret_cast<const unsigned char*>(x) + idx);
              } else {
                raw[k] = ld_as_float<T>(x + (((long long)b * 3 + ci) * H + iy) * W + ix);
              }
            }
          }
    };
    // pack the taps into the thread's K-major row of A tile `buf` and hand the tile to the issuer
    auto store_row = [&](const Raw (&raw)[27], uint32_t inside, int buf, uint32_t empty_parity) {
      float in[28];
      in[27] = 0.f;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        if constexpr (kU8) in[k] = ((inside >> k) & 1u) ? lut[(k / 9) * 256 + (int)raw[k]] : 0.f;
        else in[k] = raw[k];
      }
      uint32_t pk[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) pk[i] = pack_bf16x2(in[2 * i], in[2 * i + 1]);
      mbar_wait(&a_empty[buf], empty_parity, 11);
      uint8_t* row = sA[buf] + r * 64;
      const int sw = (r >> 1) & 3;
      *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      *reinterpret_cast<uint4*>(row + ((2 ^ sw) << 4)) = make_uint4(pk[8], pk[9], pk[10], pk[11]);
      *reinterpret_cast<uint4*>(row + ((3 ^ sw) << 4)) = make_uint4(pk[12], pk[13], 0u, 0u);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    };
    Raw raw[27];
    uint32_t inside;
    if ((int)blockIdx.x < num_tiles) {
      load_raw(blockIdx.x, raw, inside);
      store_row(raw, inside, 0, 1u);
    }
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const int tn = t + gridDim.x;
      load_raw(tn, raw, inside);                     // in flight during the epilogue below
      // ---- epilogue of tile t
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u, 14);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 64);
      uint32_t acc[C1 / 16][16];
#pragma unroll
      for (int c = 0; c < C1 / 16; ++c) tmem_ld_x16(taddr + (uint32_t)(c * 16), acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      const int pix = t * 128 + r;
      if (pix < npix) {
        bf16* dst = out + (long long)pix * C1;
#pragma unroll
        for (int c = 0; c < C1 / 16; ++c) {
          uint32_t wv[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bq = *reinterpret_cast<const float4*>(sBias + c * 16 + 4 * j);
            const float2 g0 = gelu_fast2(fadd2(make_float2(__uint_as_float(acc[c][4 * j]), __uint_as_float(acc[c][4 * j + 1])), make_float2(bq.x, bq.y)));
            const float2 g1 = gelu_fast2(fadd2(make_float2(__uint_as_float(acc[c][4 * j + 2]), __uint_as_float(acc[c][4 * j + 3])), make_float2(bq.z, bq.w)));
            wv[2 * j] = pack_bf16x2(g0.x, g0.y);
            wv[2 * j + 1] = pack_bf16x2(g1.x, g1.y);
          }
          if (al32) {
            st_global_256(dst + 16 * c, wv);
          } else {
            reinterpret_cast<uint4*>(dst + 16 * c)[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            reinterpret_cast<uint4*>(dst + 16 * c)[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
          }
        }
      }
      // ---- tile t + stride: its taps have arrived by now
      if (tn < num_tiles) store_row(raw, inside, buf ^ 1, ((uint32_t)((it + 1) >> 1) & 1u) ^ 1u);
    }
  } else {
    // ---------------- MMA issuer (warp-uniform control flow, one elected lane issues) ----------------
    const uint32_t idesc = make_idesc_bf16(128, C1);
    const uint64_t db = make_kmajor_desc<64>(smem_u32(sW));
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[buf], ph ^ 1u, 12);
      mbar_wait(&a_full[buf], ph, 13);
      tc_fence_after();
      const uint64_t da = make_kmajor_desc<64>(smem_u32(sA[buf]));
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 64);
      umma_bf16_ss_warp(d_tmem, da, db, idesc, 0u);
      umma_bf16_ss_warp(d_tmem, da + 2ull, db + 2ull, idesc, 1u);     // +32 bytes: k = 16..31 inside the 64-byte swizzle span
      umma_commit_warp(&a_empty[buf]);
      umma_commit_warp(&acc_full[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------
// im2col for conv3x3 stride 2 pad 1 on token-major activations: out[(b,oy,ox), tap*C + ci]
// reference: stem conv 2 and downsample convs (models/lemevit.py:702,715)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_3x3s2_kernel(Im2colArgs a, int Ho, int Wo) {
  // one warp per output pixel: its 9 * C/8 16-byte vectors are contiguous in the patch row, so lane j writes vector j
  // (fully coalesced) and reads vector (j % vecs) of tap (j / vecs); the pixel is decomposed once per warp in 32-bit math
  pdl_launch_dependents();
  pdl_wait();
  const int vecs = a.C >> 3, S = 9 * vecs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npix = a.B * Ho * Wo, HoWo = Ho * Wo;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= npix) return;
  const int b = pix / HoWo, rem = pix - b * HoWo, oy = rem / Wo, ox = rem - oy * Wo;
  const uint4* in_b = reinterpret_cast<const uint4*>(a.in + (long long)b * a.T * a.C);
  uint4* out_p = reinterpret_cast<uint4*>(a.out + (long long)pix * (9LL * a.C));
#pragma unroll 4
  for (int j = lane; j < S; j += 32) {
    const int tap = j / vecs, vec = j - tap * vecs;
    const int ky = tap / 3, kx = tap - 3 * ky;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) u = __ldg(in_b + (long long)(iy * a.W + ix) * vecs + vec);
    out_p[j] = u;
  }
}

// ------------------------------------------------------------------------------------------------
// classification tail: BN(eval) -> spatial mean, LN(c) -> token mean, sum  (models/lemevit.py:815-827)
// mean(BN(x)) == BN_affine(mean(x)); one CTA per image.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tail_kernel(TailArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // mu[M], rstd[M]
  float* mu = sm;
  float* rs = sm + a.M;
  const int b = blockIdx.x;
  const bf16* x = a.x + (long long)b * a.x_bs;
  const bf16* c = a.c + (long long)b * a.c_bs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int m = warp; m < a.M; m += nwarps) {
    const bf16* r = c + (long long)m * a.C;
    float s = 0.f;
    for (int ch = lane; ch < a.C; ch += 32) s += __bfloat162float(r[ch]);
    const float mean = warp_sum(s) / (float)a.C;
    float ss = 0.f;
    for (int ch = lane; ch < a.C; ch += 32) {
      const float d = __bfloat162float(r[ch]) - mean;
      ss += d * d;
    }
    const float var = warp_sum(ss) / (float)a.C;
    if (lane == 0) { mu[m] = mean; rs[m] = rsqrtf(var + a.eps); }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < a.C; ch += blockDim.x) {
    float sx = 0.f;
    for (int n = 0; n < a.N; ++n) sx += __bfloat162float(x[(long long)n * a.C + ch]);
    float sc = 0.f;
    for (int m = 0; m < a.M; ++m) sc += (__bfloat162float(c[(long long)m * a.C + ch]) - mu[m]) * rs[m];
    const float f = a.bn_scale[ch] * (sx / (float)a.N) + a.bn_shift[ch] + a.ln_gamma[ch] * (sc / (float)a.M) + a.ln_beta[ch];
    a.feat[(long long)b * a.C + ch] = __float2bfloat16(f);
  }
}

// ------------------------------------------------------------------------------------------------
// token-major -> NCHW (backbone outputs, semantic_segmentation/.../lemevit.py:800-820)
// ------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256)
tokens_to_nchw_kernel(const bf16* __restrict__ tok, TO* __restrict__ out, int N, int T, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bf16* src = tok + (long long)b * T * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, ch = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && ch < C) ? __bfloat162float(src[(long long)n * C + ch]) : 0.f;
  }
  __syncthreads();
  TO* dst = out + (long long)b * C * N;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i, n = n0 + threadIdx.x;
    if (n < N && ch < C) dst[(long long)ch * N + n] = (TO)tile[threadIdx.x][i];
  }
}

__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int per_image_vecs, int B, long long dst_bs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * per_image_vecs) return;
  const int b = (int)(idx / per_image_vecs), v = (int)(idx % per_image_vecs);
  reinterpret_cast<uint4*>(dst + (long long)b * dst_bs)[v] = __ldg(reinterpret_cast<const uint4*>(src) + v);
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int rows, int vecs, int C, int grp_rows,
                   int grp_stride, int grp_off) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * vecs) return;
  const int r = (int)(idx / vecs), v = (int)(idx % vecs);
  const long long srow = (long long)(r / grp_rows) * grp_stride + grp_off + (r % grp_rows);
  reinterpret_cast<uint4*>(dst + (long long)r * C)[v] = __ldg(reinterpret_cast<const uint4*>(src + srow * C) + v);
}

__global__ void __launch_bounds__(256)
row_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats, int R, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= R) return;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = __bfloat162float(x[row * C + c]);
    s1 += v; s2 = fmaf(v, v, s2);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { stats[2 * row] = s1; stats[2 * row + 1] = s2; }
}

inline unsigned blocks_for(long long total, int per_block) { return (unsigned)((total + per_block - 1) / per_block); }

}  // namespace

template <int G, int ITERS>
static void launch_posln_v2(const PosLnArgs& a, cudaStream_t s) {
  constexpr int SLOTS = 256 / G;
  const long long rows = (long long)a.B * a.T;
  int iters = (4 * a.W + SLOTS - 1) / SLOTS;                       // ~4 image rows per block
  iters = std::max(1, std::min(iters, 16));
  const long long want_blocks = 2LL * device_sm_count();            // keep at least two waves of blocks
  while (iters > 1 && (rows + (long long)SLOTS * iters - 1) / (SLOTS * iters) < want_blocks) --iters;
  const long long per_block = (long long)SLOTS * iters;
  (void)launch_kernel(posembed_ln_v2_kernel<G, ITERS>, dim3((unsigned)((rows + per_block - 1) / per_block)), dim3(256), (size_t)(0), s, a, iters);   // error surfaces through cudaGetLastError in the caller
}

int posembed_ln_run(const PosLnArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.C % 2 == 0 && a.C <= 64 * kMaxPairs, "posembed_layernorm: C must be even and <= 512");
  LMV_REQUIRE(a.T >= (a.dw_w ? a.H * a.W : 0), "posembed_layernorm: T < H*W");
  const long long rows = (long long)a.B * a.T;
  if (rows == 0) return LMV_OK;
  if (posembed_tile_supported(a)) {
    PosEmbedOp op;
    int rc = posembed_tile_prepare(a, &op);
    if (rc) return rc;
    return posembed_tile_run(op, s);
  }
  if (a.C % 8 == 0) {
    const int V = a.C / 8;
    if (V <= 4) launch_posln_v2<4, 1>(a, s);
    else if (V <= 8) launch_posln_v2<8, 1>(a, s);
    else if (V <= 16) launch_posln_v2<16, 1>(a, s);
    else if (V <= 32) launch_posln_v2<32, 1>(a, s);
    else launch_posln_v2<32, 2>(a, s);
  } else {
    LMV_REQUIRE(a.stats_out == nullptr, "posembed_layernorm: statistics output needs C % 8 == 0");
    LMV_CUDA_OK(launch_kernel(posembed_ln_kernel, dim3(blocks_for(rows, 8)), dim3(256), (size_t)(0), s, a));
  }
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int row_stats_run(const bf16* x, float* stats, int R, int C, cudaStream_t s) {
  if (R == 0) return LMV_OK;
  LMV_CUDA_OK(launch_kernel(row_stats_kernel, dim3(blocks_for(R, 8)), dim3(256), (size_t)(0), s, x, stats, R, C));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int layernorm_run(const LnArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.C % 2 == 0, "layernorm: C must be even");
  if (a.R == 0) return LMV_OK;
  LMV_CUDA_OK(launch_kernel(layernorm_kernel, dim3(blocks_for(a.R, 8)), dim3(256), (size_t)(0), s, a));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int stem_im2col_run(const StemArgs& a, cudaStream_t s) {
  const int Ho = (a.H + 1) / 2, Wo = (a.W + 1) / 2;
  const int Kp = ((a.Cin * 9 + 7) / 8) * 8;
  const long long total = (long long)a.B * Ho * Wo * (Kp / 8);
  if (total == 0) return LMV_OK;
  if (a.x_dtype == LMV_DTYPE_F32)
    LMV_CUDA_OK(launch_kernel(stem_im2col_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), s, (const float*)a.x, a.out, a.B, a.Cin, a.H, a.W, Ho, Wo, Kp));
  else
    LMV_CUDA_OK(launch_kernel(stem_im2col_kernel<bf16>, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), s, (const bf16*)a.x, a.out, a.B, a.Cin, a.H, a.W, Ho, Wo, Kp));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

bool stem_conv1_supported(int Cin, int C1) { return Cin == 3 && (C1 == 32 || C1 == 48); }

int stem_conv1_run(const StemArgs& a, const bf16* w, const float* bias, int C1, cudaStream_t s) {
  LMV_REQUIRE(stem_conv1_supported(a.Cin, C1), "stem_conv1: needs 3 input channels and 32 or 48 output channels");
  const int Ho = (a.H + 1) / 2, Wo = (a.W + 1) / 2, Kp = ((a.Cin * 9 + 7) / 8) * 8;
  const long long total = (long long)a.B * Ho * Wo;
  if (total == 0) return LMV_OK;
  StemNorm nrm;
  for (int i = 0; i < 3; ++i) { nrm.mean[i] = a.mean[i]; nrm.std[i] = a.std[i]; }
  const bool use_tc = a.tensor_core && Kp == 32 && total < (1ll << 31) - 256;
  const unsigned grid = use_tc ? (unsigned)std::min<long long>((total + 127) / 128, 3LL * device_sm_count())   /* three resident CTAs per SM */
                               : blocks_for((total + 1) / 2, 128);   // CUDA-core kernel: two output pixels per thread
  auto go = [&](auto kern_tc, auto kern, auto* xp) -> int {
    if (use_tc) LMV_CUDA_OK(launch_kernel(kern_tc, dim3(grid), dim3(kStemTcThreads), (size_t)(0), s, xp, w, bias, a.out, a.B, a.H, a.W, Ho, Wo, nrm));
    else LMV_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(128), (size_t)(0), s, xp, w, bias, a.out, a.B, a.H, a.W, Ho, Wo, Kp, nrm));
    LMV_CUDA_OK(cudaGetLastError());
    return LMV_OK;
  };
#define LMV_STEM_GO(T, L, ptr) (C1 == 32 ? go(stem_conv1_tc_kernel<T, 32, L>, stem_conv1_kernel<T, 32, L>, ptr) \
                                          : go(stem_conv1_tc_kernel<T, 48, L>, stem_conv1_kernel<T, 48, L>, ptr))
  switch (a.x_dtype) {
    case LMV_DTYPE_F32: return LMV_STEM_GO(float, 0, (const float*)a.x);
    case LMV_DTYPE_BF16: return LMV_STEM_GO(bf16, 0, (const bf16*)a.x);
    case LMV_DTYPE_U8: return LMV_STEM_GO(uint8_t, 0, (const uint8_t*)a.x);
    case LMV_DTYPE_U8_NHWC: return LMV_STEM_GO(uint8_t, 1, (const uint8_t*)a.x);
    default: return fail(LMV_ERR_INVALID, "stem_conv1: x dtype");
  }
#undef LMV_STEM_GO
}

int im2col_run(const Im2colArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.C % 8 == 0, "im2col: C must be a multiple of 8");
  const int Ho = (a.H + 1) / 2, Wo = (a.W + 1) / 2;
  const long long npix = (long long)a.B * Ho * Wo;
  if (npix == 0) return LMV_OK;
  LMV_REQUIRE(npix < (1ll << 31), "im2col: more than 2^31 output pixels");
  LMV_CUDA_OK(launch_kernel(im2col_3x3s2_kernel, dim3((unsigned)((npix + 7) / 8)), dim3(256), (size_t)(0), s, a, Ho, Wo));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int tail_run(const TailArgs& a, cudaStream_t s) {
  if (a.B == 0) return LMV_OK;
  LMV_CUDA_OK(launch_kernel(tail_kernel, dim3(a.B), dim3(256), (size_t)(2 * a.M * sizeof(float)), s, a));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int tokens_to_nchw_run(const ToNchwArgs& a, cudaStream_t s) {
  const int N = a.H * a.W;
  if (a.B == 0 || N == 0) return LMV_OK;
  dim3 grid((N + 31) / 32, (a.C + 31) / 32, a.B), block(32, 8);
  if (a.out_dtype == LMV_DTYPE_F32)
    LMV_CUDA_OK(launch_kernel(tokens_to_nchw_kernel<float>, dim3(grid), dim3(block), (size_t)(0), s, a.tokens, (float*)a.out, N, a.T, a.C));
  else
    LMV_CUDA_OK(launch_kernel(tokens_to_nchw_kernel<bf16>, dim3(grid), dim3(block), (size_t)(0), s, a.tokens, (bf16*)a.out, N, a.T, a.C));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int broadcast_rows_run(const bf16* src, bf16* dst, int rows, int C, int B, long long dst_bs, cudaStream_t s) {
  LMV_REQUIRE((rows * C) % 8 == 0 && dst_bs % 8 == 0, "broadcast_rows: sizes must be multiples of 8 elements");
  const int vecs = rows * C / 8;
  const long long total = (long long)B * vecs;
  if (total == 0) return LMV_OK;
  LMV_CUDA_OK(launch_kernel(broadcast_rows_kernel, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), s, src, dst, vecs, B, dst_bs));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

int gather_rows_run(const bf16* src, bf16* dst, int rows, int C, int grp_rows, int grp_stride, int grp_off,
                    cudaStream_t s) {
  LMV_REQUIRE(C % 8 == 0 && grp_rows > 0, "gather_rows: C must be a multiple of 8");
  const long long total = (long long)rows * (C / 8);
  if (total == 0) return LMV_OK;
  LMV_CUDA_OK(launch_kernel(gather_rows_kernel, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), s, src, dst, rows, C / 8, C, grp_rows, grp_stride, grp_off));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
