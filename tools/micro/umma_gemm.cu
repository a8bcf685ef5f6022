// Stand-alone prototype (NOT part of libcsm_b200.so, not on the product path): C[M,N] = A[M,K] * W[N,K]^T in bf16 with
// fp32 accumulation on the 5th-generation tensor cores -- tcgen05.mma issued by one thread, operands staged in shared
// memory by the TMA engine (128-byte swizzle), the 128x128 accumulator in tensor memory, read back with tcgen05.ld.
// This is the GEMM shape of the prefill projections (csm_prefill.cu calls cuBLAS for them today; DESIGN.md section 7,
// item 5): A = activations [rows, in], W = nn.Linear weight [out, in], both K-major.
//
// STATUS: written and compile-checked here (nvcc 12.9, sm_100a: ptxas accepts every tcgen05 / TMA form below);
// NOT yet run on a B200 -- the first GPU session has to run `umma_gemm 1024 2048 2048` (self-check against a plain
// CUDA reference GEMM, exits non-zero on mismatch) before anything is built on it.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/micro/umma_gemm tools/micro/umma_gemm.cu
// Run:    tools/micro/umma_gemm [M N K]      (multiples of 128 / 128 / 64; wrap in `timeout 60`)
//
// One CTA per 128x128 tile of C, 192 threads:
//   warp 0        TMA producer: per 64-wide k-block one A tile [128 x 64] and one W tile [128 x 64] into a 4-stage ring,
//                 completing on full[stage]; waits on empty[stage] before re-using a slot; also owns the TMEM allocation
//   warp 1        MMA issuer (one elected lane): 4 x tcgen05.mma (K = 16 each) per stage, tcgen05.commit -> empty[stage];
//                 after the last k-block tcgen05.commit -> acc_full
//   warps 2..5    epilogue: warp w reads TMEM lanes 32*(w%4).. (its quarter of the 128 rows), 32 columns at a time,
//                 converts to bf16 and stores its row segment
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef __nv_bfloat16 bf16;

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 128;   // fp32 accumulator 128 lanes x 128 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor): K-major tile whose rows are 128 bytes
// (64 bf16) written by TMA with the 128-byte swizzle; 8-row groups are 1024 bytes apart (stride byte offset); the
// leading byte offset is not used by swizzled K-major layouts; version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);            // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                               // [16,30) leading byte offset >> 4 (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;                     // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                               // [46,48) descriptor version
  d |= (uint64_t)2 << 61;                               // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor (InstrDescriptor): D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {   // lane i of the warp: TMEM lane base+i, 32 columns
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(192, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, bf16* __restrict__ C,
                 int M, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128: 1024-byte tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 0) {   // TMEM allocation (whole warp), address written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);                      // (first pass: a fresh barrier passes parity 1)
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(smem + s * STAGE_BYTES, &map_a, &full[s], kb * BK, m0);
        tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &map_w, &full[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {             // 32 bytes along K inside the 128-byte swizzle atom
          umma_bf16(tmem_acc, umma_desc_k128(a0 + k * UMMA_K * 2), umma_desc_k128(b0 + k * UMMA_K * 2), idesc,
                    (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);                             // slot free once these MMAs have read it
      }
      umma_commit(acc_full);                                // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 only
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row < M) {
        uint4* dst = reinterpret_cast<uint4*>(C + (size_t)row * N + n0 + c);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[8 * v + 2 * e]), __uint_as_float(r[8 * v + 2 * e + 1]));
            w[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          dst[v] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

// plain reference: one thread per output element, fp32 accumulation in k order
__global__ void ref_gemm_kernel(const bf16* A, const bf16* W, bf16* C, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
  C[(size_t)m * N + n] = __float2bfloat16_rn(acc);
}

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 2; } \
  } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, K] bf16 row-major, box [box_rows x 64] with the 128-byte swizzle
static int make_map(EncodeTiled enc, CUtensorMap* map, void* base, int rows, int K, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  int M = 1024, N = 2048, K = 2048;
  if (argc >= 4) { M = atoi(argv[1]); N = atoi(argv[2]); K = atoi(argv[3]); }
  if (M % BM || N % BN || K % BK) { fprintf(stderr, "M, N, K must be multiples of %d, %d, %d\n", BM, BN, BK); return 2; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr) { fprintf(stderr, "cuTensorMapEncodeTiled not available\n"); return 2; }
  EncodeTiled enc = (EncodeTiled)fn;

  std::vector<bf16> hA((size_t)M * K), hW((size_t)N * K);
  uint32_t st = 12345u;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (auto& v : hA) v = __float2bfloat16(rnd());
  for (auto& v : hW) v = __float2bfloat16(rnd() * 0.1f);
  bf16 *dA, *dW, *dC, *dR;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 2)); CK(cudaMalloc(&dR, (size_t)M * N * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 2));
  CUtensorMap map_a, map_w;
  if (make_map(enc, &map_a, dA, M, K, BM) || make_map(enc, &map_w, dW, N, K, BN)) return 2;
  CK(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(N / BN, M / BM);
  umma_gemm_kernel<<<grid, 192, SMEM_BYTES>>>(map_a, map_w, dC, M, N, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  ref_gemm_kernel<<<dim3((N + 127) / 128, M), 128>>>(dA, dW, dR, M, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<bf16> hC((size_t)M * N), hR((size_t)M * N);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hR.data(), dR, hR.size() * 2, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  size_t bad = 0;
  for (size_t i = 0; i < hC.size(); ++i) {
    const double c = __bfloat162float(hC[i]), r = __bfloat162float(hR[i]);
    const double e = fabs(c - r);
    if (!(e <= 0.02 * fabs(r) + 0.02)) ++bad;   // (also catches NaN)
    if (e > maxerr) maxerr = e;
    if (fabs(r) > maxref) maxref = fabs(r);
  }
  printf("umma_gemm M=%d N=%d K=%d: max |err| %.5f (max |ref| %.3f), %zu of %zu outside tolerance\n", M, N, K, maxerr, maxref, bad,
         hC.size());
  if (bad) return 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 20;
  for (int i = 0; i < 3; ++i) umma_gemm_kernel<<<grid, 192, SMEM_BYTES>>>(map_a, map_w, dC, M, N, K);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) umma_gemm_kernel<<<grid, 192, SMEM_BYTES>>>(map_a, map_w, dC, M, N, K);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%.3f ms per GEMM, %.1f TFLOP/s\n", ms / reps, 2.0 * M * N * K / (ms / reps * 1e-3) / 1e12);
  return 0;
}
