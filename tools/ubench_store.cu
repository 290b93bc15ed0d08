// Micro-benchmark: how fast can an SM push shared-memory tiles to global memory?
//   mode 0: cp.async.bulk shared->global, one op of CHUNK bytes at a time per block (commit, wait_group.read 0)
//   mode 1: the same, two buffers in flight (wait_group.read 1)
//   mode 2: LSU: the block's threads copy the tile with 16-byte ld.shared / st.global
//   mode 3: bulk, tile split into 4 ops issued back to back
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_store tools/ubench_store.cu
// Run:   tools/ubench_store            (prints GB/s for every mode x blocks per SM x tile size)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_s2g(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g), "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int tile_bytes, int tiles_per_block, int nbuf) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int td = tile_bytes / 8;
  for (int i = tid; i < td * nbuf; i += 256) sm[i] = (double)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int t = 0; t < tiles_per_block; ++t) {
    double* g = out + ((size_t)blockIdx.x * tiles_per_block + t) * td;
    double* s = sm + (size_t)(t % nbuf) * td;
    if (MODE == 0) {
      if (tid == 0) { bulk_s2g(g, s, tile_bytes); commit(); wait0(); }
      __syncthreads();
    } else if (MODE == 1) {
      if (tid == 0) { bulk_s2g(g, s, tile_bytes); commit(); wait1(); }
      __syncthreads();
    } else if (MODE == 3) {
      if (tid == 0) {
        for (int c = 0; c < 4; ++c) bulk_s2g(g + c * (td / 4), s + c * (td / 4), tile_bytes / 4);
        commit(); wait0();
      }
      __syncthreads();
    } else {
      const double2* s2 = reinterpret_cast<const double2*>(s);
      double2* g2 = reinterpret_cast<double2*>(g);
      for (int i = tid; i < td / 2; i += 256) g2[i] = s2[i];
      __syncthreads();
    }
  }
  if (MODE != 2 && tid == 0) wait0();
}

int main() {
  const size_t total = (size_t)1 << 30;           // 1 GiB written per launch
  double* out; cudaMalloc(&out, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sms = 148;
  for (int mode = 0; mode < 4; ++mode)
    for (int bps : {1, 2, 4, 6})
      for (int kb : {8, 16, 32}) {
        const int tile = kb * 1024, nbuf = mode == 1 ? 2 : 1;
        const int blocks = sms * bps;
        const int tpb = (int)(total / ((size_t)blocks * tile));
        const size_t smem = (size_t)tile * nbuf;
        if (smem * bps > 220 * 1024) continue;
        auto launch = [&]() {
          switch (mode) {
            case 0: cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<blocks, 256, smem>>>(out, tile, tpb, nbuf); break;
            case 1: cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<blocks, 256, smem>>>(out, tile, tpb, nbuf); break;
            case 2: cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<blocks, 256, smem>>>(out, tile, tpb, nbuf); break;
            default: cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<3><<<blocks, 256, smem>>>(out, tile, tpb, nbuf); break;
          }
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)blocks * tpb * tile;
        printf("mode %d blocks/SM %d tile %2d KB: %7.1f GB/s (%5.1f GB/s per SM)  %s\n", mode, bps, kb, bytes / ms / 1e6, bytes / ms / 1e6 / sms,
               cudaGetLastError() == cudaSuccess ? "" : "ERROR");
      }
  return 0;
}
