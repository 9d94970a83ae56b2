// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of warps reading (their own lane quarter of) TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tools/micro/tmem_ld_bench.cu && ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__global__ void __launch_bounds__(1024, 1) k(long long* out, int iters, float* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    const uint32_t col = (uint32_t)(((i + warp) * 32) & 255);
    if (SHAPE == 32) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
          "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(t + col));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j += 8) acc += __uint_as_float(v[j]);
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 148 * 32 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 12, 16, 20, 24, 32}) {
    k<32><<<1, warps * 32>>>(d, iters, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double bytes = (double)warps * iters * 32 * 32 * 4;
    printf("warps %2d: %8lld cycles for %d x32 loads per warp -> %.1f cycles per load per warp, %.1f B/clk per SM\n", warps, mx, iters, (double)mx / iters, bytes / mx);
  }
  return 0;
}
