// Standalone sm_100a microbenchmarks behind the cost model in DESIGN.md (not part of the library):
//   mma     cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) for N = 32..256, issued back to back by one thread
//   tmemld  cycles per tcgen05.ld.32x32b.x32 with 4/8/16 warps in flight
//   tanh    MUFU.TANH and MUFU.EX2 throughput per SM
//   tma     latency and streaming rate of 8/16 KB TMA boxes from L2 into a ring of S slots
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "../lemevit_b200/csrc/umma.cuh"

using namespace lmv;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_bench(int N, int reps, int same_a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  // A: 4 K-blocks of [128 x 64] (64 KB), B: 4 K-blocks of [256 x 64] (128 KB) — contents irrelevant (zeros)
  for (int i = threadIdx.x; i < (192 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int kb = same_a ? 0 : (r & 3);
      const uint64_t da = make_kmajor_desc<128>(a0 + kb * 16384), db = make_kmajor_desc<128>(b0 + kb * 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tbase, da + 2ull * k, db + 2ull * k, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) tmemld_bench(int reps, long long* out) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const uint32_t addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t v[32];
    tmem_ld_x32(addr + (uint32_t)((r * 32) & 255), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) out[2] = acc;
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) mufu_bench(int reps, int which, float* sink, long long* out) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (which == 0) {
      x0 = tanh_approx(x0); x1 = tanh_approx(x1); x2 = tanh_approx(x2); x3 = tanh_approx(x3);
    } else {
      asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

// ---------------------------------------------------------------------------------------------
// one producer thread streams `boxes` TMA boxes of [rows x 64] bf16 from a 2-D tensor into a ring of `slots` slots;
// a consumer thread releases each slot as soon as it is full.  Measures cycles for the whole stream (rate) and the
// latency of the first box.
__global__ void __launch_bounds__(64, 1) tma_bench(const __grid_constant__ CUtensorMap tm, int rows, int slots, int boxes, int nrow_tiles,
                                                    long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int bytes = rows * 128;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&empty[s], ph ^ 1u, 1);
      mbar_expect_tx(&full[s], (uint32_t)bytes);
      const int tile = (i * 7 + blockIdx.x * 13) % nrow_tiles;
      tma_load_2d(smem + (size_t)s * bytes, &tm, &full[s], 0, tile * rows);
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
    out[2 * blockIdx.x + 0] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    long long first = 0;
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&full[s], ph, 2);
      if (i == 0) first = clock64() - t0;
      mbar_arrive(&empty[s]);
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
    out[2 * blockIdx.x + 1] = clock64() - t0;
    if (blockIdx.x == 0) out[2 * gridDim.x] = first;
  }
}


// ---------------------------------------------------------------------------------------------
// TMA streaming, second form: 3-D tensor {64, rows_total, kblocks}; one request fetches `kbs` K-blocks of [rows x 64]
// (rows*kbs*128 bytes).  mode 0: every CTA streams the SAME box sequence (weight streaming); mode 1: every CTA its own region.
// cluster > 1: CTA r of a cluster loads rows [r*rows/cluster, +rows/cluster) of every box and multicasts it to all CTAs.
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__global__ void __launch_bounds__(64, 1) tma3_bench(const __grid_constant__ CUtensorMap tm, int rows, int kbs, int slots, int boxes, int ntiles,
                                                     int mode, int cluster, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t full[16], empty[16];
  const uint32_t rank = cluster > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (uint32_t)cluster); }
    fence_mbar_init();
  }
  __syncthreads();
  if (cluster > 1) cluster_sync_all();
  const int bytes = rows * kbs * 128;
  const int part_rows = rows / cluster;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&empty[s], ph ^ 1u, 1);
      mbar_expect_tx(&full[s], (uint32_t)bytes);
      const int base = mode == 0 ? 0 : (int)(blockIdx.x / cluster) * 7;
      const int tile = (i + base) % ntiles;
      if (cluster == 1) {
        tma_load_3d(smem + (size_t)s * bytes, &tm, &full[s], 0, tile * rows, 0);
      } else {
        // each K-block's [rows x 64] slab is split by rows between the CTAs of the cluster: one request per K-block
        for (int kb = 0; kb < kbs; ++kb)
          tma_load_3d_mc(smem + (size_t)s * bytes + (size_t)kb * rows * 128 + (size_t)rank * part_rows * 128, &tm, &full[s], 0,
                         tile * rows + rank * part_rows, kb, (uint16_t)((1u << cluster) - 1));
      }
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&full[s], ph, 2);
      if (cluster == 1) mbar_arrive(&empty[s]);
      else for (int c = 0; c < cluster; ++c) mbar_arrive_cluster(&empty[s], c);
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (cluster > 1) cluster_sync_all();
}


// ---------------------------------------------------------------------------------------------
// TMA streaming, third form: P producer threads (one per warp, each owning every P-th slot) and either tensor (2-D contiguous)
// or descriptor-less 1-D bulk copies of `bytes` each.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(160, 1) tma4_bench(const __grid_constant__ CUtensorMap tm, const uint8_t* src, int bytes, int slots, int boxes,
                                                      int P, int use_bulk, long long total_bytes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < P && lane == 0) {
    for (int i = warp; i < boxes; i += P) {
      const int s = i % slots;
      const uint32_t ph = (uint32_t)(i / slots) & 1u;
      mbar_wait(&empty[s], ph ^ 1u, 1);
      mbar_expect_tx(&full[s], (uint32_t)bytes);
      const long long off = ((long long)(i + blockIdx.x * 5) * bytes) % total_bytes;
      if (use_bulk) bulk_load_1d(smem + (size_t)s * bytes, src + off, (uint32_t)bytes, &full[s]);
      else tma_load_2d(smem + (size_t)s * bytes, &tm, &full[s], 0, (int)(off / 128));
    }
  } else if (warp == 4 && lane == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < boxes; ++i) {
      const int s = i % slots;
      mbar_wait(&full[s], (uint32_t)(i / slots) & 1u, 2);
      mbar_arrive(&empty[s]);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}


// ---------------------------------------------------------------------------------------------
// tma5: K requests of `bytes` issued back to back (mode 0: by one thread; mode 1: by K threads of K warps; mode 2: by K lanes of
// one warp) on ONE barrier; reports issue time and completion time.  Repeated `reps` times, averaged.
__global__ void __launch_bounds__(256, 1) tma5_bench(const __grid_constant__ CUtensorMap tm, int rows, int K, int mode, int reps, int nrow_tiles,
                                                      long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); tma_prefetch_desc(&tm); }
  __syncthreads();
  const int bytes = rows * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t_issue = 0, t_done = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, (uint32_t)(bytes * K));
    const int base = (r * 8 + blockIdx.x * 3) % (nrow_tiles - 8);
    if (mode == 0) {
      if (threadIdx.x == 0) {
        for (int k = 0; k < K; ++k) tma_load_2d(smem + (size_t)k * bytes, &tm, &bar, 0, (base + k) * rows);
        t_issue += clock64() - t0;
      }
    } else if (mode == 1) {
      if (warp < K && lane == 0) tma_load_2d(smem + (size_t)warp * bytes, &tm, &bar, 0, (base + warp) * rows);
      if (threadIdx.x == 0) t_issue += clock64() - t0;
    } else {
      if (warp == 0 && lane < K) tma_load_2d(smem + (size_t)lane * bytes, &tm, &bar, 0, (base + lane) * rows);
      if (threadIdx.x == 0) t_issue += clock64() - t0;
    }
    if (threadIdx.x == 0) {
      mbar_wait(&bar, (uint32_t)r & 1u, 3);
      t_done += clock64() - t0;
    }
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t_issue; out[2 * blockIdx.x + 1] = t_done; }
}


// ---------------------------------------------------------------------------------------------
// mma2: MMA stream (as mma_bench) while a second thread streams TMA boxes into a separate smem ring at full rate
// (and optionally 8 more warps hammer shared memory with LDS/STS like an epilogue transpose does).
__global__ void __launch_bounds__(384, 1) mma2_bench(const __grid_constant__ CUtensorMap tm, int N, int reps, int tma_on, int lds_on, int box_rows,
                                                      int nrow_tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t bar, tfull[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    // two bf16 values in (-2, 2): sign + exponent 0x3f/0x3e + random mantissa
    const uint32_t v = (lds_on & 4) ? ((h & 0x807f807fu) | 0x3f003f00u) : 0u;
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&tfull[i], 1); fence_mbar_init(); stop = 0; tma_prefetch_desc(&tm); }
  if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + 96 * 1024;        // 3 x 32 KB ring for the TMA stream
  uint8_t* scratch = smem + 192 * 1024;    // 16 KB for the LDS/STS warps
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);   // A: 2 K-blocks (32 KB), B: 2 K-blocks of 256 rows (64 KB)
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int kb = r & 1;
      const uint64_t da = make_kmajor_desc<128>(a0 + kb * 16384), db = make_kmajor_desc<128>(b0 + kb * 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tbase, da + 2ull * k, db + 2ull * k, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t2 = clock64();
    stop = 1;
    out[3 * blockIdx.x] = t2 - t0;
  } else if (warp == 1 && lane == 0 && tma_on) {
    const int bytes = box_rows * 128;
    long long n = 0;
    uint32_t ph[3] = {0, 0, 0};
    for (int s = 0; s < 3; ++s) {
      mbar_expect_tx(&tfull[s], (uint32_t)bytes);
      tma_load_2d(ring + (size_t)s * 32768, &tm, &tfull[s], 0, ((int)(n + blockIdx.x * 5) % nrow_tiles) * box_rows);
      ++n;
    }
    int s = 0;
    while (!stop) {
      mbar_wait(&tfull[s], ph[s], 2);
      ph[s] ^= 1u;
      mbar_expect_tx(&tfull[s], (uint32_t)bytes);
      tma_load_2d(ring + (size_t)s * 32768, &tm, &tfull[s], 0, ((int)(n + blockIdx.x * 5) % nrow_tiles) * box_rows);
      ++n;
      if (++s == 3) s = 0;
    }
    for (int k = 0; k < 3; ++k) { mbar_wait(&tfull[s], ph[s], 3); if (++s == 3) s = 0; }
    out[3 * blockIdx.x + 1] = n * bytes;
  } else if (warp >= 4 && (lds_on & 2)) {
    // epilogue-like TMEM drain of the OTHER accumulator half while the MMA stream accumulates into columns [0, N)
    const uint32_t addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    uint32_t acc = 0;
    long long n = 0;
    while (!stop) {
      uint32_t v[32];
      tmem_ld_x32(addr + (uint32_t)((n * 32) & 255), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
      ++n;
    }
    if (lane == 0 && warp == 4) out[3 * blockIdx.x + 2] = n * 4096 / 512 * 512;
    if (acc == 0x12345u) out[0] = 0;
  } else if (warp >= 4 && (lds_on & 1)) {
    float4* sc = reinterpret_cast<float4*>(scratch) + (warp - 4) * 128;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[lane + 32 * i] = v;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float4 w = sc[(lane ^ 5) + 32 * i]; v.x += w.x; v.y += w.y; }
      __syncwarp();
      n += 8;
    }
    if (lane == 0 && warp == 4) out[3 * blockIdx.x + 2] = n * 512;
    if (v.x == 123.456f) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}


// ---------------------------------------------------------------------------------------------
// tma6: boxes of `rows` rows x `cols` bf16 columns cut out of a wide row-major matrix (row pitch `pitch` bytes), i.e. many short
// row segments per request — what an attention kernel does when it fetches one head's [keys x 32] slice of a packed qkv tensor.
__global__ void __launch_bounds__(64, 1) tma6_bench(const __grid_constant__ CUtensorMap tm, int box_bytes, int slots, int boxes, int ncol_tiles,
                                                     int nrow_tiles, int rows, int cols, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&empty[s], ph ^ 1u, 1);
      mbar_expect_tx(&full[s], (uint32_t)box_bytes);
      const int ct = (i + blockIdx.x) % ncol_tiles, rt = (i * 3 + blockIdx.x * 7) % nrow_tiles;
      tma_load_2d(smem + (size_t)s * 32768, &tm, &full[s], ct * cols, rt * rows);
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(&full[s], ph, 2);
      mbar_arrive(&empty[s]);
      if (++s == slots) { s = 0; ph ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  long long* out;
  CK(cudaMallocManaged(&out, 4096 * sizeof(long long)));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  // ---- mma
  CK(cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int same_a = 0; same_a < 2; ++same_a)
    for (int N : {32, 64, 96, 128, 192, 256}) {
      const int reps = 256;
      mma_bench<<<sms, 128, 193 * 1024 + 1024>>>(N, reps, same_a, out);
      CK(cudaDeviceSynchronize());
      printf("mma  M=128 N=%3d K=16 same_a=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (ideal %5.1f)\n", N, same_a, out[0] / (4.0 * reps),
             out[1] / (4.0 * reps), N / 2.0);
    }
  // ---- tmem ld
  for (int warps : {4, 8, 16, 32}) {
    const int reps = 512;
    tmemld_bench<<<sms, warps * 32>>>(reps, out);
    CK(cudaDeviceSynchronize());
    printf("tmem ld.x32: %2d warps: %6.1f cyc per ld per warp -> %6.1f B/cyc/SM\n", warps, out[0] / (double)reps, warps * 4096.0 * reps / out[0]);
  }
  // ---- mufu
  float* sink;
  CK(cudaMalloc(&sink, sms * 1024 * sizeof(float)));
  for (int which = 0; which < 2; ++which)
    for (int threads : {128, 256, 512, 1024}) {
      const int reps = 2048;
      mufu_bench<<<sms, threads>>>(reps, which, sink, out);
      CK(cudaDeviceSynchronize());
      printf("mufu %s %4d threads: %6.2f ops/cyc/SM\n", which ? "ex2 " : "tanh", threads, 4.0 * reps * threads / out[0]);
    }
  // ---- tma
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int total_rows = 64 * 1024;   // 64K rows x 64 cols bf16 = 8 MB: L2 resident
    void* buf;
    CK(cudaMalloc(&buf, (size_t)total_rows * 128));
    CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
    CK(cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rows : {64, 128, 256})
      for (int slots : {1, 2, 4, 8}) {
        if (rows * 128 * slots > 190 * 1024) continue;
        CUtensorMap tm;
        cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)rows}, estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int boxes = 512;
        for (int rep = 0; rep < 2; ++rep) {
          tma_bench<<<sms, 64, rows * 128 * slots + 1024>>>(tm, rows, slots, boxes, total_rows / rows, out);
          CK(cudaDeviceSynchronize());
        }
        double avg = 0;
        for (int b = 0; b < sms; ++b) avg += out[2 * b + 1];
        avg /= sms;
        printf("tma  box %3d x 64 (%2d KB) slots %d: first-box latency %5lld cyc, %7.1f cyc/box -> %5.1f B/cyc/SM (all %d SMs streaming)\n", rows,
               rows * 128 / 1024, slots, out[2 * sms], avg / boxes, rows * 128.0 * boxes / avg, sms);
      }
  }
  // ---- tma, 3-D boxes / same-vs-distinct addresses / cluster multicast
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int total_rows = 8192, kblocks = 6;     // W-like tensor [8192 rows x 384] bf16 = 6 MB (L2 resident)
    void* buf;
    CK(cudaMalloc(&buf, (size_t)total_rows * kblocks * 128));
    CK(cudaMemset(buf, 0, (size_t)total_rows * kblocks * 128));
    CK(cudaFuncSetAttribute(tma3_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(tma3_bench, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    struct Cfg { int rows, kbs, slots, cluster; };
    const Cfg cfgs[] = {{64, 1, 4, 1}, {128, 1, 4, 1}, {256, 1, 4, 1}, {64, 6, 3, 1}, {128, 3, 3, 1}, {128, 6, 2, 1}, {64, 3, 4, 1},
                        {128, 1, 4, 2}, {128, 1, 4, 4}, {256, 1, 4, 2}, {256, 1, 4, 4}, {128, 3, 3, 2}, {128, 3, 3, 4}, {256, 1, 4, 8}};
    for (const Cfg& c : cfgs)
      for (int mode = 0; mode < 2; ++mode) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {64, (cuuint64_t)total_rows, (cuuint64_t)kblocks};
        cuuint64_t strides[2] = {(cuuint64_t)kblocks * 128, 128};
        const int box_rows = c.rows / c.cluster, box_kb = c.cluster > 1 ? 1 : c.kbs;
        cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_kb}, estr[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode3 failed %d\n", (int)r); return 1; }
        const int boxes = 384, grid = (sms / c.cluster) * c.cluster;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(64); lc.dynamicSmemBytes = (size_t)c.rows * c.kbs * 128 * c.slots + 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) {
          cudaError_t e = cudaLaunchKernelEx(&lc, tma3_bench, tm, c.rows, c.kbs, c.slots, boxes, total_rows / c.rows, mode, c.cluster, out);
          if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); break; }
          CK(cudaDeviceSynchronize());
        }
        double avg = 0;
        for (int b = 0; b < grid; ++b) avg += out[b];
        avg /= grid;
        printf("tma3 box %3d rows x %d kb (%3d KB) slots %d cluster %d %s: %7.1f cyc/box -> %6.1f B/cyc/SM delivered\n", c.rows, c.kbs,
               c.rows * c.kbs * 128 / 1024, c.slots, c.cluster, mode ? "distinct" : "same    ", avg / boxes, c.rows * c.kbs * 128.0 * boxes / avg);
      }
  }
  // ---- tma6: short row segments out of a wide matrix
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int width = 1152, total_rows = 16384;          // [16384 x 1152] bf16 = 37.7 MB (L2 resident)
    void* buf;
    CK(cudaMalloc(&buf, (size_t)total_rows * width * 2));
    CK(cudaMemset(buf, 0, (size_t)total_rows * width * 2));
    CK(cudaFuncSetAttribute(tma6_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { int rows, cols, swz; };
    const Cfg cfgs[] = {{224, 32, 64}, {128, 32, 64}, {224, 64, 128}, {128, 64, 128}, {64, 64, 128}, {256, 16, 32}};
    for (const Cfg& c : cfgs) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)total_rows};
      cuuint64_t strides[1] = {(cuuint64_t)width * 2};
      cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows}, estr[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       c.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : c.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode7 failed %d\n", (int)r); return 1; }
      const int boxes = 512, slots = 4, box_bytes = c.rows * c.cols * 2;
      for (int rep = 0; rep < 2; ++rep) {
        tma6_bench<<<sms, 64, 4 * 32768 + 1024>>>(tm, box_bytes, slots, boxes, width / c.cols, total_rows / c.rows, c.rows, c.cols, out);
        CK(cudaDeviceSynchronize());
      }
      double avg = 0;
      for (int b = 0; b < sms; ++b) avg += out[b];
      avg /= sms;
      printf("tma6 box %3d rows x %2d cols (%3d B segments, pitch 2304 B): %7.1f cyc/box = %5.2f cyc/row -> %5.1f B/cyc/SM\n", c.rows, c.cols, c.cols * 2,
             avg / boxes, avg / boxes / c.rows, box_bytes * (double)boxes / avg);
    }
  }
  // ---- mma2: MMA rate under concurrent TMA / LDS traffic
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int total_rows = 64 * 1024;
    void* buf;
    CK(cudaMalloc(&buf, (size_t)total_rows * 128));
    CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
    CK(cudaFuncSetAttribute(mma2_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const int box_rows = 256;
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode6 failed %d\n", (int)r); return 1; }
    for (int N : {64, 128, 192, 256})
      for (int cfg : {0, 1, 2, 4, 5, 8, 9, 13}) {
        if (N == 256 && (cfg & 4)) continue;   // the TMEM-drain variant needs columns [256, 512) free
        const int tma_on = cfg & 1, lds_on = cfg >> 1, reps = 2048;
        for (int i = 0; i < 3 * sms; ++i) out[i] = 0;
        for (int rep = 0; rep < 2; ++rep) {
          mma2_bench<<<sms, 384, 209 * 1024 + 1024>>>(tm, N, reps, tma_on, lds_on, box_rows, total_rows / box_rows, out);
          CK(cudaDeviceSynchronize());
        }
        double cyc = 0, tb = 0, lb = 0;
        for (int b = 0; b < sms; ++b) { cyc += out[3 * b]; tb += out[3 * b + 1]; lb += out[3 * b + 2]; }
        cyc /= sms;
        printf("mma2 N=%3d tma=%d mode(1=lds,2=tmemld,4=random data)=%d: %6.1f cyc/mma (ideal %5.1f), concurrent TMA %5.1f B/cyc/SM, LDS+STS %5.1f B/cyc/SM(warp 4 only x8)\n", N, tma_on, lds_on,
               cyc / (4.0 * reps), N / 2.0, tb / sms / cyc, 8.0 * lb / sms / cyc);
      }
  }
  // ---- tma5: issue cost vs completion of K back-to-back requests
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int total_rows = 64 * 1024;
    void* buf;
    CK(cudaMalloc(&buf, (size_t)total_rows * 128));
    CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
    CK(cudaFuncSetAttribute(tma5_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rows : {64, 128, 256})
      for (int mode = 0; mode < 3; ++mode)
        for (int K : {1, 2, 4, 6}) {
          if (rows * 128 * K > 192 * 1024) continue;
          CUtensorMap tm;
          cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
          cuuint64_t strides[1] = {128};
          cuuint32_t box[2] = {64, (cuuint32_t)rows}, estr[2] = {1, 1};
          CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode5 failed %d\n", (int)r); return 1; }
          const int reps = 64;
          for (int rep = 0; rep < 2; ++rep) {
            tma5_bench<<<sms, 256, (size_t)rows * 128 * K + 1024>>>(tm, rows, K, mode, reps, total_rows / rows, out);
            CK(cudaDeviceSynchronize());
          }
          double ti = 0, td = 0;
          for (int b = 0; b < sms; ++b) { ti += out[2 * b]; td += out[2 * b + 1]; }
          printf("tma5 %2d KB x K=%d mode %d (%s): issue %6.0f cyc, all done %6.0f cyc -> %6.1f B/cyc/SM burst\n", rows * 128 / 1024, K, mode,
                 mode == 0 ? "1 thread " : mode == 1 ? "K warps  " : "K lanes  ", ti / sms / reps, td / sms / reps, rows * 128.0 * K / (td / sms / reps));
        }
  }
  // ---- tma4: producers x {tensor 2-D contiguous, 1-D bulk}
  {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const long long total = 8ll << 20;
    uint8_t* buf;
    CK(cudaMalloc(&buf, total));
    CK(cudaMemset(buf, 0, total));
    CK(cudaFuncSetAttribute(tma4_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int use_bulk = 0; use_bulk < 2; ++use_bulk)
      for (int bytes : {8192, 16384, 32768})
        for (int P : {1, 2, 4}) {
          if (!use_bulk && bytes > 32768) continue;
          const int slots = std::min(8, (int)(190 * 1024 / bytes));
          CUtensorMap tm;
          cuuint64_t dims[2] = {64, (cuuint64_t)(total / 128)};
          cuuint64_t strides[1] = {128};
          cuuint32_t box[2] = {64, (cuuint32_t)std::min(256, bytes / 128)}, estr[2] = {1, 1};
          CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode4 failed %d\n", (int)r); return 1; }
          const int boxes = 512;
          for (int rep = 0; rep < 2; ++rep) {
            tma4_bench<<<sms, 160, (size_t)bytes * slots + 1024>>>(tm, buf, bytes, slots, boxes, P, use_bulk, total, out);
            CK(cudaDeviceSynchronize());
          }
          double avg = 0;
          for (int b = 0; b < sms; ++b) avg += out[b];
          avg /= sms;
          printf("tma4 %s %2d KB slots %d producers %d: %7.1f cyc/req -> %6.1f B/cyc/SM\n", use_bulk ? "bulk1d" : "tensor", bytes / 1024, slots, P,
                 avg / boxes, (double)bytes * boxes / avg);
        }
  }
  return 0;
}
