// SIMT attention: softmax(scale * Q K^T) V per (image, head), head_dim = 32, arbitrary Lq / Lk and
// strides.  This is the cross-check kernel for the tensor-core attention kernels and the first
// correct path for the degenerate-shaped attentions of the C / D blocks (16 meta tokens on one
// side; reference models/lemevit.py:297-301,484) until the fused DCA kernel replaces them.
// One thread owns one query (optionally a 1/nsplit share of the keys), online softmax in fp32.
#include "kernels.h"
#include "umma.cuh"

namespace lmv {

namespace {

constexpr int kD = 32;
constexpr int kKeysPerTile = 64;
constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
attention_simt_kernel(AttnArgs a, int QT, int nsplit) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float sK[kKeysPerTile][kD];
  __shared__ __align__(16) float sV[kKeysPerTile][kD];
  __shared__ float sM[kThreads], sL[kThreads];
  __shared__ float sAcc[kThreads][kD + 1];

  const int b = blockIdx.z, h = blockIdx.y;
  const int ql = threadIdx.x % QT, split = threadIdx.x / QT;
  const int qrow = blockIdx.x * QT + ql;
  const bool valid = qrow < a.Lq;

  float q[kD];
  if (valid) {
    const bf16* qp = a.q + (long long)b * a.q_bs + (long long)qrow * a.q_rs + h * kD;
#pragma unroll
    for (int d = 0; d < kD; ++d) q[d] = __bfloat162float(qp[d]) * a.scale;
  } else {
#pragma unroll
    for (int d = 0; d < kD; ++d) q[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f, acc[kD];
#pragma unroll
  for (int d = 0; d < kD; ++d) acc[d] = 0.f;

  const bf16* kbase = a.k + (long long)b * a.k_bs + h * kD;
  const bf16* vbase = a.v + (long long)b * a.v_bs + h * kD;
  for (int k0 = 0; k0 < a.Lk; k0 += kKeysPerTile) {
    const int tile = min(kKeysPerTile, a.Lk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kKeysPerTile * kD / 2; i += kThreads) {
      const int key = i / (kD / 2), dp = i % (kD / 2);
      float2 kk = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
      if (key < tile) {
        kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kbase + (long long)(k0 + key) * a.k_rs + 2 * dp));
        vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vbase + (long long)(k0 + key) * a.v_rs + 2 * dp));
      }
      sK[key][2 * dp] = kk.x; sK[key][2 * dp + 1] = kk.y;
      sV[key][2 * dp] = vv.x; sV[key][2 * dp + 1] = vv.y;
    }
    __syncthreads();
    if (valid) {
      for (int j = split; j < tile; j += nsplit) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < kD; ++d) s = fmaf(q[d], sK[j][d], s);
        if (s > m) {
          const float corr = expf(m - s);  // exp(-inf) = 0 on the first key
          l *= corr;
#pragma unroll
          for (int d = 0; d < kD; ++d) acc[d] *= corr;
          m = s;
        }
        const float p = expf(s - m);
        l += p;
#pragma unroll
        for (int d = 0; d < kD; ++d) acc[d] = fmaf(p, sV[j][d], acc[d]);
      }
    }
  }
  sM[threadIdx.x] = m;
  sL[threadIdx.x] = l;
#pragma unroll
  for (int d = 0; d < kD; ++d) sAcc[threadIdx.x][d] = acc[d];
  __syncthreads();
  if (split == 0 && valid) {
    float mm = m;
    for (int s2 = 1; s2 < nsplit; ++s2) mm = fmaxf(mm, sM[s2 * QT + ql]);
    float ll = 0.f, o[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) o[d] = 0.f;
    for (int s2 = 0; s2 < nsplit; ++s2) {
      const int t = s2 * QT + ql;
      const float w = (sM[t] == -INFINITY) ? 0.f : expf(sM[t] - mm);
      ll += sL[t] * w;
#pragma unroll
      for (int d = 0; d < kD; ++d) o[d] = fmaf(sAcc[t][d], w, o[d]);
    }
    const float inv = 1.f / ll;
    bf16* op = a.out + (long long)b * a.o_bs + (long long)qrow * a.o_rs + h * kD;
#pragma unroll
    for (int d = 0; d < kD; d += 2)
      *reinterpret_cast<__nv_bfloat162*>(op + d) = __floats2bfloat162_rn(o[d] * inv, o[d + 1] * inv);
  }
}

}  // namespace

int attention_simt_run(const AttnArgs& a, cudaStream_t s) {
  LMV_REQUIRE(a.Lq > 0 && a.Lk > 0 && a.B > 0 && a.heads > 0, "attention: empty problem");
  LMV_REQUIRE(a.q_rs % 2 == 0 && a.k_rs % 2 == 0 && a.v_rs % 2 == 0 && a.o_rs % 2 == 0, "attention: odd row stride");
  int QT = 128;
  if (a.Lq <= 16) QT = 16;
  else if (a.Lq <= 32) QT = 32;
  else if (a.Lq <= 64) QT = 64;
  const int nsplit = kThreads / QT;
  dim3 grid((a.Lq + QT - 1) / QT, a.heads, a.B);
  LMV_CUDA_OK(launch_kernel(attention_simt_kernel, dim3(grid), dim3(kThreads), (size_t)(0), s, a, QT, nsplit));
  LMV_CUDA_OK(cudaGetLastError());
  return LMV_OK;
}

}  // namespace lmv
