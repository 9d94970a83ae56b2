// Micro-benchmark: chip-wide global store bandwidth (pure write stream) vs read and copy, for buffers around and above the L2 size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_write_bench tools/micro/hbm_write_bench.cu && ./hbm_write_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_write(uint4* dst, size_t n, uint32_t v) {
  const uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = x;
}
__global__ void k_read(const uint4* src, size_t n, uint32_t* sink) {
  uint32_t a = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const uint4 x = src[i]; a ^= x.x ^ x.y ^ x.z ^ x.w; }
  if (a == 0x12345678u) sink[0] = a;
}
__global__ void k_copy(const uint4* src, uint4* dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
  const size_t maxb = 2ull << 30;
  uint4 *a, *b; uint32_t* sink;
  cudaMalloc(&a, maxb); cudaMalloc(&b, maxb); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, maxb); cudaMemset(b, 2, maxb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (size_t mb : {42, 84, 167, 334, 1024, 2048}) {
    const size_t bytes = mb << 20, n = bytes / 16;
    float tw = 1e9f, tr = 1e9f, tc = 1e9f, t;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0); k_write<<<148 * 8, 512>>>(a, n, rep); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&t, e0, e1); if (rep && t < tw) tw = t;
      cudaEventRecord(e0); k_read<<<148 * 8, 512>>>(a, n, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&t, e0, e1); if (rep && t < tr) tr = t;
      cudaEventRecord(e0); k_copy<<<148 * 8, 512>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&t, e0, e1); if (rep && t < tc) tc = t;
    }
    printf("%5zu MB: write %7.1f us = %6.2f TB/s | read %7.1f us = %6.2f TB/s | copy %7.1f us = %6.2f TB/s (read + write)\n", mb, tw * 1e3, bytes / tw / 1e9,
           tr * 1e3, bytes / tr / 1e9, tc * 1e3, 2.0 * bytes / tc / 1e9);
  }
  return 0;
}
