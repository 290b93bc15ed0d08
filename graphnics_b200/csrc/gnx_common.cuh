// Shared device helpers: error plumbing, warp/block reductions and scans.
#ifndef GNX_COMMON_CUH
#define GNX_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "gnx_index.h"

void gnx_set_error(const char* fmt, ...);

#define GNX_CHECK_ARG(cond)                                                       \
  do {                                                                            \
    if (!(cond)) {                                                                \
      gnx_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);        \
      return GNX_ERR_ARG;                                                         \
    }                                                                             \
  } while (0)

#define GNX_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t err__ = (call);                                                   \
    if (err__ != cudaSuccess) {                                                   \
      gnx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                    cudaGetErrorString(err__));                                   \
      return GNX_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define GNX_LAUNCH_CHECK() GNX_CUDA(cudaGetLastError())

static inline int gnx_blocks(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

constexpr unsigned GNX_FULL = 0xffffffffu;

__device__ __forceinline__ double gnx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(GNX_FULL, v, o);
  return v;
}

// Sum over the block; result valid in thread 0 (and broadcast to all if BCAST).  Deterministic
// (fixed tree).  `sh` must hold >= 32 doubles.
template <bool BCAST>
__device__ __forceinline__ double gnx_block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = gnx_warp_sum(v);
  __syncthreads();  // protect sh reuse
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? sh[lane] : 0.0;
    t = gnx_warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  if (BCAST) {
    __syncthreads();
    v = sh[0];
  } else {
    v = sh[0];  // only meaningful in thread 0 of warp 0
  }
  return v;
}

// Every block reduces the same `n` partials in the same order -> identical bits in every block.
__device__ __forceinline__ double gnx_reduce_partials(const double* __restrict__ part, int n,
                                                      double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  return gnx_block_sum<true>(v, sh);
}

// inclusive scan inside a lane group of width `w` (power of two <= 32)
__device__ __forceinline__ double gnx_group_incl_scan(double v, int w) {
  const int gl = (threadIdx.x & 31) & (w - 1);
  for (int o = 1; o < w; o <<= 1) {
    double t = __shfl_up_sync(GNX_FULL, v, o, w);
    if (gl >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ double gnx_group_sum(double v, int w) {
  for (int o = w >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(GNX_FULL, v, o, w);
  return v;
}

#endif  // GNX_COMMON_CUH
