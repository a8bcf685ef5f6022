// Dense projections of the context prefill on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
//   C[R, N] = A[R, K] * W[N, K]^T      bf16 operands, fp32 accumulation in tensor memory
//
// A = activation rows (tokens), W = an nn.Linear weight in its natural [out, in] layout -- the q/k/v/o and
// gate/up/down projections of hf LlamaAttention / LlamaMLP (modeling_llama.py:262-264,288,183) that the reference
// reaches through CSMModel.forward (modeling_csm.py:345-354) with S = T context frames, and the projection of the
// audio embedding table at create time (modeling_csm.py:564-565).  Replaces the cuBLAS calls of round 1.
//
// One persistent CTA per SM, 192 threads, tiles of 128 rows x 256 columns:
//   warp 0      TMA producer: per 64-wide k-block one A box [128 x 64] and one W box [256 x 64] (128-byte swizzle) into
//               a 4-stage ring, completing on full[stage]; owns the TMEM allocation (512 columns = two accumulators)
//   warp 1      MMA issuer (one lane): 4 x tcgen05.mma (M 128, N 256, K 16) per stage; tcgen05.commit frees the stage
//               and, after the last k-block of a tile, hands the accumulator to the epilogue
//   warps 2..5  epilogue of tile i while the tensor core works on tile i+1 (the other accumulator): tcgen05.ld, the
//               fused element-wise tail of the projection, bf16 stores.  Fused tails (all rounding points as the
//               reference: the linear output is rounded to bf16 first):
//                 EPI_STORE   plain store
//                 EPI_RESID   h += y                      (hf modeling_llama.py:325,331)
//                 EPI_SWIGLU  bf16(silu(gate)) * up       (hf modeling_llama.py:183; W rows interleaved gate_j, up_j)
//                 EPI_QKV     RoPE on q and k, q -> C, k and v -> the KV cache at their positions
//                             (hf modeling_llama.py:146-168,267-270; cache_utils.py:102-121 as an in-place write)
// Tiles are dealt round-robin (tile = m-block fastest), so the CTAs working at the same time share W tiles through L2.
#include <cuda.h>

#include "csm_common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 512;          // two fp32 accumulators of 128 lanes x 256 columns
constexpr int GEMM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Shared-memory matrix descriptor: K-major tile whose rows are 128 bytes (64 bf16) written by TMA with the 128-byte
// swizzle; 8-row groups are 1024 bytes apart (stride byte offset); version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (the tile is stored [64 k][64 m/n] boxes of 128-byte rows, 128-byte swizzle, as TMA writes a box of a
// [K, MN] row-major matrix): canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 8 k-rows of one
// 64-element MN chunk form a 1024-byte swizzle atom; SBO = 1024 (next 8 k-rows inside a box), LBO = 8192 (next
// 64-element MN chunk = next box).
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor: D = F32, A = B = BF16, N >> 3 at [17,23), M >> 4 at [24,29); bit 15 / 16: A / B is MN-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ float lin(uint32_t acc_bits) { return bfround(__uint_as_float(acc_bits)); }   // nn.Linear output is bf16

// ---- fused tails: one thread = one row of the tile, 32 (or 64) consecutive columns at a time
__device__ __forceinline__ void store32(bf16* dst, const uint32_t (&r)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int v = 0; v < 4; ++v)
    d[v] = make_uint4(pack_bf16(__uint_as_float(r[8 * v]), __uint_as_float(r[8 * v + 1])),
                      pack_bf16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                      pack_bf16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])),
                      pack_bf16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
}
__device__ __forceinline__ void resid32(bf16* dst, const uint32_t (&r)[32]) {   // residual + f(x), both bf16
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint4 h = d[v];
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack_bf16(bf_lo(hw[e]) + lin(r[8 * v + 2 * e]), bf_hi(hw[e]) + lin(r[8 * v + 2 * e + 1]));
    d[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__device__ __forceinline__ void swiglu32(bf16* dst, const uint32_t (&r)[32]) {   // columns (2j, 2j+1) = (gate_j, up_j) -> 16 outputs
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g0 = lin(r[4 * j]), u0 = lin(r[4 * j + 1]), g1 = lin(r[4 * j + 2]), u1 = lin(r[4 * j + 3]);
    const float s0 = bfround(g0 / (1.f + expf(-g0))), s1 = bfround(g1 / (1.f + expf(-g1)));
    o[j] = pack_bf16(s0 * u0, s1 * u1);
  }
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(o[0], o[1], o[2], o[3]);
  d[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

template <int EPI, bool A_MN = false, bool B_MN = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
csm_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128: 1024-byte tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // accumulator ready (MMA -> epilogue)
  uint64_t* tempty = tfull + 2;          // accumulator drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (p.R + BM - 1) / BM, nt = (p.N + BN - 1) / BN, ntiles = mt * nt;
  const int nkb = (p.K + BK - 1) / BK;   // (a ragged last k-block is zero-filled by TMA on both operands)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 0) {   // TMEM allocation (whole warp), address written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile % mt) * BM, n0 = (tile / mt) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);                    // (first pass: a fresh barrier passes parity 1)
          mbar_expect_tx(&full[s], STAGE_BYTES);           // rows / columns outside the tensor are zero-filled and counted
          unsigned char* sa = smem + s * STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &map_a, &full[s], kb * BK, m0);
          } else {                                          // [64 k x 64 m] boxes of the [K, M] matrix
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &map_a, &full[s], m0 + 64 * i, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sa + A_BYTES, &map_w, &full[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sa + A_BYTES + i * 8192, &map_w, &full[s], n0 + 64 * i, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&tempty[buf], aph ^ 1u);                  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 32 bytes along K inside the 128-byte swizzle atom; MN-major: 16 k-rows = two 1024-byte atoms
            const uint64_t da = A_MN ? umma_desc_mn128(a0 + k * 2048) : umma_desc_k128(a0 + k * UMMA_K * 2);
            const uint64_t db = B_MN ? umma_desc_mn128(b0 + k * 2048) : umma_desc_k128(b0 + k * UMMA_K * 2);
            umma_bf16(tacc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);                           // slot free once these MMAs have read it
        }
        umma_commit(&tfull[buf]);                           // accumulator complete
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 only
    const int q = warp & 3;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++acc_it) {
      const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      const int m0 = (tile % mt) * BM, n0 = (tile / mt) * BN;
      mbar_wait(&tfull[buf], aph);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool rok = row < p.R;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
      if (EPI == EPI_QKV) {
        // columns: q heads | k heads | v heads, 64 wide each; this tile holds 4 whole heads of one kind
        const int hd = 64, half = 32;
        const int nq = p.heads * hd, nkv = p.kv * hd;
        const int bl = rok ? row / p.S : 0, sp = rok ? row - bl * p.S : 0;
        const int pos = p.pos0 + sp, b = p.b0 + bl;
        uint32_t cs[16], sn[16];   // cos / sin of this row's position, packed pairs (hf: rounded to bf16 first)
        if (n0 < nq + nkv) {
          const uint4* cp = reinterpret_cast<const uint4*>(p.cos_t + (size_t)pos * half);
          const uint4* sp4 = reinterpret_cast<const uint4*>(p.sin_t + (size_t)pos * half);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 c4 = __ldg(cp + v), s4 = __ldg(sp4 + v);
            cs[4 * v] = c4.x; cs[4 * v + 1] = c4.y; cs[4 * v + 2] = c4.z; cs[4 * v + 3] = c4.w;
            sn[4 * v] = s4.x; sn[4 * v + 1] = s4.y; sn[4 * v + 2] = s4.z; sn[4 * v + 3] = s4.w;
          }
        }
#pragma unroll 1
        for (int hg = 0; hg < BN / 64; ++hg) {
          const int col0 = n0 + hg * 64;
          if (col0 >= p.N) break;                           // (warp-uniform)
          uint32_t a[32], bq[32];
          tmem_ld32(tacc + hg * 64, a);
          tmem_ld32(tacc + hg * 64 + 32, bq);
          tmem_ld_wait();
          if (!rok) continue;
          if (col0 < nq + nkv) {
            const bool isq = col0 < nq;
            const int head = (isq ? col0 : col0 - nq) / hd;
            bf16* dst = isq ? p.C + (size_t)row * p.ldc + col0
                            : p.kc + ((((size_t)p.layer * p.Bmax + b) * p.kv + head) * p.Tcap + pos) * hd;
            uint32_t o1[16], o2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // apply_rotary_pos_emb (hf modeling_llama.py:146-168): every product and the sum round to bf16
              const float x1a = lin(a[2 * i]), x1b = lin(a[2 * i + 1]), x2a = lin(bq[2 * i]), x2b = lin(bq[2 * i + 1]);
              const float ca = bf_lo(cs[i]), cb = bf_hi(cs[i]), sa = bf_lo(sn[i]), sb = bf_hi(sn[i]);
              o1[i] = pack_bf16(bfround(x1a * ca) + bfround(-x2a * sa), bfround(x1b * cb) + bfround(-x2b * sb));
              o2[i] = pack_bf16(bfround(x2a * ca) + bfround(x1a * sa), bfround(x2b * cb) + bfround(x1b * sb));
            }
            uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              d[v] = make_uint4(o1[4 * v], o1[4 * v + 1], o1[4 * v + 2], o1[4 * v + 3]);
              d[4 + v] = make_uint4(o2[4 * v], o2[4 * v + 1], o2[4 * v + 2], o2[4 * v + 3]);
            }
          } else {
            const int head = (col0 - nq - nkv) / hd;
            bf16* dst = p.vc + ((((size_t)p.layer * p.Bmax + b) * p.kv + head) * p.Tcap + pos) * hd;
            store32(dst, a);
            store32(dst + 32, bq);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          if (n0 + c >= p.N) break;                         // (warp-uniform)
          uint32_t r[32];
          tmem_ld32(tacc + c, r);
          tmem_ld_wait();
          if (!rok) continue;
          if (EPI == EPI_STORE) store32(p.C + (size_t)row * p.ldc + n0 + c, r);
          else if (EPI == EPI_RESID) resid32(p.C + (size_t)row * p.ldc + n0 + c, r);
          else swiglu32(p.C + (size_t)row * p.ldc + ((n0 + c) >> 1), r);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled g_encode = nullptr;

template <int EPI, bool A_MN = false, bool B_MN = false>
cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mw, const GemmParams& p, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void*)csm_gemm_kernel<EPI, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES);
  if (e != cudaSuccess) return e;
  csm_gemm_kernel<EPI, A_MN, B_MN><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(ma, mw, p);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

// Tensor map of a row-major bf16 matrix [rows, K] (row pitch `pitch` elements) for boxes of [box_rows x 64] with the
// 128-byte swizzle.  `out` points to 128 bytes (a CUtensorMap).  box_rows: 128 for the A operand, 256 for W.
int csm_tmap_2d(void* out, const void* base, long long rows, int K, long long pitch, int box_rows) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr)
      return 1;
    g_encode = (EncodeTiled)fn;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// Tensor map of an MN-major operand: the matrix is stored [k_rows, mn] row-major (pitch elements between k rows); boxes
// of [64 k x 64 mn] with the 128-byte swizzle.
int csm_tmap_2d_mn(void* out, const void* base, long long k_rows, long long mn, long long pitch) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr)
      return 1;
    g_encode = (EncodeTiled)fn;
  }
  cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)k_rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

int csm_gemm_box_rows_a() { return BM; }
int csm_gemm_box_rows_w() { return BN; }

cudaError_t csm_gemm_launch(const void* map_a, const void* map_w, const GemmParams* p, int sms, cudaStream_t st) {
  if (p->R <= 0) return cudaSuccess;
  // STORE / RESID: any K (the tensor maps zero-fill past the end), N a multiple of 32; the fused tails need whole tiles
  if (p->N % 32 || p->K < 1) return cudaErrorInvalidValue;
  if ((p->epi == EPI_SWIGLU || p->epi == EPI_QKV) && (p->K % BK || p->N % 64)) return cudaErrorInvalidValue;
  const int mt = (p->R + BM - 1) / BM, nt = (p->N + BN - 1) / BN;
  const int grid = mt * nt < sms ? mt * nt : sms;
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(map_a);
  const CUtensorMap& mw = *reinterpret_cast<const CUtensorMap*>(map_w);
  if (p->a_mn || p->b_mn) {   // transposed operands (gradients): plain store / accumulate only
    if (p->epi == EPI_STORE && !p->a_mn && p->b_mn) return launch<EPI_STORE, false, true>(ma, mw, *p, grid, st);
    if (p->epi == EPI_STORE && p->a_mn && p->b_mn) return launch<EPI_STORE, true, true>(ma, mw, *p, grid, st);
    return cudaErrorInvalidValue;
  }
  switch (p->epi) {
    case EPI_STORE: return launch<EPI_STORE>(ma, mw, *p, grid, st);
    case EPI_RESID: return launch<EPI_RESID>(ma, mw, *p, grid, st);
    case EPI_SWIGLU: return launch<EPI_SWIGLU>(ma, mw, *p, grid, st);
    case EPI_QKV: return launch<EPI_QKV>(ma, mw, *p, grid, st);
  }
  return cudaErrorInvalidValue;
}

}  // extern "C"
