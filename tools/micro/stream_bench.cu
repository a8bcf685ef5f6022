// Microbenchmark: how fast can 148 persistent CTAs pull a contiguous per-CTA byte stream from HBM
// into shared memory / registers?  Variants:
//   mode 0: one producer thread, cp.async.bulk of `chunk` bytes into a ring of `slots`, consumer warps
//           wait full / read a little / arrive empty
//   mode 1: same but each ring slot is filled by `split` smaller bulk copies (more copies in flight)
//   mode 2: plain LDG.128 by 256 threads, `unroll` loads in flight per thread, no shared memory
// usage: stream_bench <mode> <chunk_bytes> <slots> <split> <total_MB_per_cta_x148>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

extern __shared__ __align__(128) unsigned char smem[];

__global__ void __launch_bounds__(320, 1) k_bulk(const unsigned char* __restrict__ src, size_t per_cta, int chunk, int slots,
                                                int split, unsigned* sink) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  unsigned char* ring = smem + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < n; ++it) {
        int s = it % slots;
        if (it >= slots) mbar_wait(&empty[s], ((it / slots) - 1) & 1);
        mbar_expect_tx(&full[s], chunk);
        int piece = chunk / split;
        for (int j = 0; j < split; ++j)
          bulk_g2s(ring + (size_t)s * chunk + j * piece, base + (size_t)it * chunk + j * piece, piece, &full[s]);
      }
    }
    return;
  }
  if (warp == 9) return;
  unsigned acc = 0;
  for (int it = 0; it < n; ++it) {
    int s = it % slots;
    mbar_wait(&full[s], (it / slots) & 1);
    acc += *reinterpret_cast<const unsigned*>(ring + (size_t)s * chunk + (threadIdx.x * 16) % chunk);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256, 1) k_ldg(const uint4* __restrict__ src, size_t per_cta, unsigned* sink) {
  const uint4* base = src + (size_t)blockIdx.x * (per_cta / 16);
  const size_t n = per_cta / 16;
  unsigned acc = 0;
  for (size_t i = threadIdx.x; i + (size_t)(U - 1) * 256 < n; i += (size_t)U * 256) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(base + i + (size_t)u * 256);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int chunk = argc > 2 ? atoi(argv[2]) : 32768;
  int slots = argc > 3 ? atoi(argv[3]) : 5;
  int split = argc > 4 ? atoi(argv[4]) : 1;
  size_t per_cta = (size_t)(argc > 5 ? atoi(argv[5]) : 32) << 20;
  int G = 148;
  per_cta = per_cta / chunk * chunk;
  unsigned char* src;
  unsigned* sink;
  cudaMalloc(&src, per_cta * G);
  cudaMemset(src, 1, per_cta * G);
  cudaMalloc(&sink, 16);
  size_t sm = 256 + (size_t)chunk * slots;
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0 || mode == 1) k_bulk<<<G, 320, sm>>>(src, per_cta, chunk, slots, split, sink);
    else if (split == 4) k_ldg<4><<<G, 256>>>((const uint4*)src, per_cta, sink);
    else if (split == 8) k_ldg<8><<<G, 256>>>((const uint4*)src, per_cta, sink);
    else k_ldg<16><<<G, 256>>>((const uint4*)src, per_cta, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("mode %d chunk %6d slots %2d split %2d : %.3f ms  %.1f GB/s  (%s)\n", mode, chunk, slots, split, best,
         per_cta * G / best / 1e6, cudaGetErrorString(e));
  return 0;
}
