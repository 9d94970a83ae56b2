// Host-side common definitions for liblemevit_b200: status codes, error reporting, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda, so the library
// also loads on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

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
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes);

int device_sm_count();

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
};

struct GemmParams {
  int M, N, K;
  int BN, num_stages, tiles_m, tiles_n, k_blocks;
  const float* bias;
  const bf16* residual;
  void* out;
  int ldc, out_fp32, act;
  int grp_rows, grp_stride;
};

struct GemmOp {
  CUtensorMap tmA, tmB;
  GemmParams p;
  int grid = 0;
  int smem_bytes = 0;
};

int gemm_prepare(const GemmArgs& a, GemmOp* op);
int gemm_run(const GemmOp& op, cudaStream_t stream);
int gemm_simt_run(const GemmArgs& a, cudaStream_t stream);  // bring-up cross-check kernel (tests only)

}  // namespace lmv
