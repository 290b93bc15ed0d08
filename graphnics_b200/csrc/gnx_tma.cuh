// 1-D bulk copies between shared and global memory by the TMA engine (PTX cp.async.bulk, SASS
// UBLKCP): one elected thread hands a whole contiguous tile to the copy engine instead of the
// block spending its issue slots on LDS/STG (or LDG/STS) loops.  The edge-tile kernels own whole
// edges, so each of their inputs and outputs is ONE contiguous range per array.
//
// Alignment: global and shared addresses must be 16-byte aligned and sizes multiples of 16 bytes.
// The arrays are doubles and tile ranges start at arbitrary (even or odd) element offsets, so a
// range is staged SHIFTED by the parity of its first global index (sh = base & 1: element i of
// the range lives at stage[sh + i]); the 16-byte aligned body goes through the copy engine and
// the possible first / last element through ordinary loads or stores.
#ifndef GNX_TMA_CUH
#define GNX_TMA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t gnx_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// generic-proxy writes to shared memory -> visible to the async proxy (run by every writer, before
// the barrier that precedes the bulk store)
__device__ __forceinline__ void gnx_fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void gnx_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(gnx_smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gnx_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread: all of its committed bulk stores have finished READING shared memory
__device__ __forceinline__ void gnx_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent group (double-buffered staging: the buffer of the store before last is free)
__device__ __forceinline__ void gnx_bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// Block-wide: g[base .. base+len) <- stage[sh .. sh+len), sh = base & 1.  `stage` is 16-byte aligned,
// `g` points to a 16-byte aligned array.  Every writer of `stage` has run gnx_fence_async_smem() and the
// block has synchronised.  Thread 0 issues (and must gnx_bulk_commit() / gnx_bulk_wait_read() before
// `stage` is reused or the block exits); thread 32 (or 0 in a one-warp block) stores the odd ends.
__device__ __forceinline__ void gnx_bulk_stream_out(const double* stage, double* __restrict__ g, int64_t base,
                                                    int32_t len) {
  const int32_t sh = (int32_t)(base & 1);
  const int32_t rest = len - sh;                           // elements from the first 16-byte aligned one
  if (threadIdx.x == 0) {
    const int32_t body = rest > 0 ? (rest & ~1) : 0;
    if (body > 0) gnx_bulk_s2g(g + base + sh, stage + 2 * sh, (uint32_t)body * 8u);
  }
  if (threadIdx.x == (blockDim.x > 32 ? 32 : 0)) {
    if (sh && len > 0) g[base] = stage[1];
    if (rest > 0 && (rest & 1)) g[base + len - 1] = stage[sh + len - 1];
  }
}

// ---- global -> shared, completion on an mbarrier ---------------------------------------------
__device__ __forceinline__ void gnx_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(gnx_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void gnx_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(gnx_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gnx_bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(gnx_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(gnx_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gnx_mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(gnx_smem_u32(bar)), "r"(phase) : "memory");
}

// bytes of the range [base, base+len) that travel through the copy engine
__device__ __forceinline__ uint32_t gnx_bulk_body_bytes(int64_t base, int32_t len) {
  const int32_t rest = len - (int32_t)(base & 1);
  return rest > 0 ? (uint32_t)(rest & ~1) * 8u : 0u;
}

// Thread 0 only, after gnx_mbar_expect_tx(bar, sum of gnx_bulk_body_bytes):
// stage[sh .. sh+len) <- g[base .. base+len) through the copy engine (+ plain loads of the odd ends).
// All threads then gnx_mbar_wait(bar, phase); the plain stores of thread 0 are ordered for the others
// by the __syncthreads() callers add after the wait.
__device__ __forceinline__ void gnx_bulk_stream_in(double* stage, const double* __restrict__ g, int64_t base,
                                                   int32_t len, uint64_t* bar) {
  const int32_t sh = (int32_t)(base & 1);
  const int32_t rest = len - sh;
  const uint32_t bytes = gnx_bulk_body_bytes(base, len);
  if (bytes) gnx_bulk_g2s(stage + 2 * sh, g + base + sh, bytes, bar);
  if (sh && len > 0) stage[1] = g[base];
  if (rest > 0 && (rest & 1)) stage[sh + len - 1] = g[base + len - 1];
}

#endif
