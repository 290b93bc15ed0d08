// Shared device helpers: error plumbing, warp/block reductions and scans.
#ifndef GNX_COMMON_CUH
#define GNX_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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

// Programmatic dependent launch (griddepcontrol): a kernel launched with gnx_launch_dependent may be
// scheduled while its predecessor on the stream is still draining; it must run gnx_pdl_wait() before it
// touches anything an earlier kernel wrote (a no-op under an ordinary launch).  A predecessor that calls
// gnx_pdl_trigger() lets the dependent grid become resident as soon as all of ITS blocks have started.
__device__ __forceinline__ void gnx_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gnx_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// GNX_PDL=0 (tuning knob): ordinary launches.  Measured on the C2 step: 59.4 -> 53.3 us with dependent
// launches; letting the back-substitution grid become resident DURING the Schur kernel (an early trigger
// there) takes SM slots from the overlapped assembly and loses (61.5 us).
static inline bool gnx_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GNX_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}
template <class... KArgs, class... Args>
static inline cudaError_t gnx_launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                               cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = gnx_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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

// ---------------------------------------------------------------------------------------
// Block-segmented scan / reduce: the block is tiled by segments of `seg` consecutive threads
// (seg a power of two, 1 <= seg <= blockDim.x <= 1024, blockDim.x a multiple of max(seg,32)).
// Fixed combination order -> deterministic.  `sh` holds >= 32 doubles per value; every call
// contains block barriers, so all threads of the block must call it.
// ---------------------------------------------------------------------------------------
// inclusive scan of v inside the segment; *total = segment sum
__device__ __forceinline__ double gnx_seg_incl_scan(double v, int seg, double* sh, double* total) {
  const int lane = threadIdx.x & 31;
  const int w = seg < 32 ? seg : 32;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(GNX_FULL, v, o, w);
    if (o < w && (lane & (w - 1)) >= o) v += t;
  }
  if (seg <= 32) {
    *total = __shfl_sync(GNX_FULL, v, w - 1, w);
    return v;
  }
  const int wid = threadIdx.x >> 5;
  const int wps = seg >> 5;                         // warps per segment (2..32)
  __syncthreads();                                  // sh reuse
  if (lane == 31) sh[wid] = v;
  __syncthreads();
  const int w0 = wid & ~(wps - 1);                  // first warp of my segment
  double t = (lane < wps) ? sh[w0 + lane] : 0.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(GNX_FULL, t, o);
    if (o < wps && lane >= o) t += u;
  }
  *total = __shfl_sync(GNX_FULL, t, wps - 1);
  const int me = wid - w0;
  const double before = __shfl_sync(GNX_FULL, t, me > 0 ? me - 1 : 0);
  return me > 0 ? v + before : v;
}

// segment sums of K values, broadcast to every thread of the segment
template <int K>
__device__ __forceinline__ void gnx_seg_sum(double (&v)[K], int seg, double* sh) {
  const int lane = threadIdx.x & 31;
  const int w = seg < 32 ? seg : 32;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(GNX_FULL, v[k], o, w);
      if (o < w) v[k] += t;
    }
  }
  if (seg <= 32) return;
  const int wid = threadIdx.x >> 5;
  const int wps = seg >> 5;
  const int w0 = wid & ~(wps - 1);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = (lane < wps) ? sh[k * 32 + w0 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(GNX_FULL, t, o);
    v[k] = t;
  }
}

#endif  // GNX_COMMON_CUH
