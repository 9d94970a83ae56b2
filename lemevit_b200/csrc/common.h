// Host-side common definitions for liblemevit_b200: status codes, error reporting, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda, so the library
// also loads on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/lemevit_b200.h"

namespace lmv {

typedef __nv_bfloat16 bf16;

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);  // records msg, returns code

#define LMV_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      return ::lmv::fail(_e == cudaErrorMemoryAllocation ? LMV_ERR_OOM : LMV_ERR_CUDA,             \
                         std::string(#expr) + ": " + cudaGetErrorString(_e) +                      \
                             (_e == cudaErrorMemoryAllocation ? " (CUDA out of memory)" : ""));    \
    }                                                                                              \
  } while (0)

#define LMV_REQUIRE(cond, msg)                                               \
  do {                                                                       \
    if (!(cond)) return ::lmv::fail(LMV_ERR_INVALID, std::string(msg));      \
  } while (0)

// Encode a tiled bf16 tensor map.  dims/strides innermost first; strides in BYTES for dims 1..rank-1.
// swizzle_bytes in {0, 32, 64, 128}.  Returns 0 on success.
// elem_strides (optional, innermost first): traversal stride per dimension; dimension i then delivers ceil(box[i] / elem_strides[i])
// elements (the strided gather of the implicit-GEMM convolutions).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides = nullptr);

int device_sm_count();   // SM count of the CURRENT device (cached per device)

// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) are per device: a process that runs the model on a second
// GPU (nn.DataParallel replicas — reference validate.py:260-261 — or cuda:1 after cuda:0) must set them again there.
// PerDeviceOnce runs `fn` the first time it is called with each device current (thread-safe, lock-free afterwards) and
// remembers that device's result.
struct PerDeviceOnce {
  static constexpr int kMaxDevices = 64;
  std::atomic<int> state[kMaxDevices];     // 0 = not run, 1 = ok, 2 = failed
  cudaError_t err[kMaxDevices];
  std::mutex mu;
  PerDeviceOnce() { for (auto& s : state) s.store(0); }
  template <typename F>
  cudaError_t run(F&& fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return fn();   // beyond the cache: just do it every time
    int st = state[dev].load(std::memory_order_acquire);
    if (st == 0) {
      std::lock_guard<std::mutex> lock(mu);
      st = state[dev].load(std::memory_order_relaxed);
      if (st == 0) {
        err[dev] = fn();
        st = err[dev] == cudaSuccess ? 1 : 2;
        state[dev].store(st, std::memory_order_release);
      }
    }
    return st == 1 ? cudaSuccess : err[dev];
  }
};

// Every kernel goes through this launcher: programmatic stream serialization (PDL) lets consecutive kernels of the forward overlap
// the prologue of kernel n+1 with the tail of kernel n (each kernel calls pdl_wait() before touching dependent global memory).
// LMV_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// GEMM:  out[M,N] = epilogue( A[M,K] * W[N,K]^T )     (gemm.cu — tcgen05 / TMEM / TMA)
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
  const bf16* A = nullptr;   // [M,K] row-major, leading dim lda (elements)
  int lda = 0;
  const bf16* W = nullptr;   // [N,K] row-major (nn.Linear layout), leading dim ldw
  int ldw = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;       // [N] fp32 or null
  int act = 0;                       // 0 none, 1 exact-erf GELU
  const bf16* residual = nullptr;    // same row mapping / leading dim as out, or null
  void* out = nullptr;               // bf16 (default) or fp32
  int ldc = 0;
  int out_fp32 = 0;
  // optional row remap of the OUTPUT (and residual): row r -> (r / grp_rows) * grp_stride + r % grp_rows
  int grp_rows = 0, grp_stride = 0;
  int force_bn = 0;                  // test hook: override the tile-N heuristic
  int out_patched = 0;               // `out` is replaced at launch time (GemmOp::p.out): nothing may bake its address into a tensor map
  // Implicit-GEMM convolution 3x3 / stride 2 / pad 1 (conv_C > 0): A is not a matrix but the token-major activation
  // [conv_B, conv_T, conv_C] (first conv_H * conv_W rows of every image, row = y * W + x); the GEMM rows are the output pixels
  // (b, oy, ox), M = conv_B * ceil(H/2) * ceil(W/2), K = 9 * conv_C with k = (ky * 3 + kx) * C + ci — the layout of the
  // materialised patch matrix (im2col_3x3s2_kernel) this replaces, so W [N, 9C] is used as packed.  lda is ignored.
  int conv_B = 0, conv_H = 0, conv_W = 0, conv_T = 0, conv_C = 0;
  // LayerNorm folded into the GEMM (LN(y) W'^T = r (y W'^T - mu colsum(W'))): A holds the RAW rows y,
  // ln_stats[row] = (sum_k y, sum_k y^2) of A row `row`, ln_colsum[n] = sum_k W'[n,k] (of the bf16 weights).
  const float* ln_stats = nullptr;   // [M][2] or null
  const float* ln_colsum = nullptr;  // [N]
  float ln_eps = 0.f;
  int ln_parts = 1;                  // ln_stats holds [M][ln_parts][2]: partial sums, added up by the consumer
  // per OUTPUT row, partial (sum, sum of squares) of the stored bf16 values: stats_out[orow][part] with
  // part = 2 * n_tile + warp_half, gemm_stats_parts(N) parts per row — plain stores, deterministic, no zeroing needed
  float* stats_out = nullptr;        // [rows][gemm_stats_parts(N)][2] or null
};

struct GemmParams {
  int M, N, K;
  int BN, num_stages, tiles_m, tiles_n, k_blocks;
  const float* bias;
  const bf16* residual;
  void* out;
  int ldc, out_fp32, act;
  int grp_rows, grp_stride;
  const float* ln_stats;
  const float* ln_colsum;
  float ln_eps, ln_inv_k;
  int ln_parts;
  float* stats_out;
  int prefetch_res;
  // implicit-GEMM convolution (conv != 0): an M tile is `tile_rows` <= 128 output pixels = conv_bb whole images (maps of <= 64
  // pixels) or conv_bh output rows of one image (conv_tpi such tiles per image); K blocks walk (tap, 64-channel slice)
  int conv, conv_cpt, conv_bb, conv_bh, conv_tpi, conv_Ho, conv_Wo, conv_HoWo;
  int tile_rows;        // GEMM rows per M tile (128 unless conv)
  uint32_t tx_bytes;    // bytes one pipeline stage receives (A box + W box)
  int row_epi;      // lane = row epilogue (no shared-memory transpose; implies tma_out)
  int row_res;      // ... with the in-place residual read through TMA boxes (x += A W^T + b)
  int w_res;        // tiles_n == 1 and all K blocks of W fit beside the A ring: W is loaded once per CTA and stays resident
  int stage_bytes;  // bytes of one ring slot (A tile, + W tile unless w_res)
  int tma_out;   // the fast-path epilogue leaves through TMA tile stores (tmC) instead of per-lane global stores
};

struct GemmOp {
  CUtensorMap tmA, tmB, tmR;   // tmR: residual tile, only used for L2 prefetch
  CUtensorMap tmC;             // output, 32 x 32 boxes (64-byte rows, 64B swizzle): p.tma_out
  GemmParams p;
  int grid = 0;
  int smem_bytes = 0;
};

int gemm_stats_parts(int N, int force_bn = 0);   // partial-sum slots per row that stats_out receives
bool gemm_conv_supported(int H, int W, int C);   // implicit-GEMM conv3x3/s2/p1 on [B, T, C] tokens (GemmArgs::conv_*)
int gemm_prepare(const GemmArgs& a, GemmOp* op);
int gemm_run(const GemmOp& op, cudaStream_t stream);
int gemm_simt_run(const GemmArgs& a, cudaStream_t stream);  // bring-up cross-check kernel (tests only)

}  // namespace lmv
