// Causal GQA flash attention BACKWARD on tcgen05 / TMEM / TMA, head dim 64 (training step, csm_train.cu).
//
// Adjoint of csm_flash_tc_kernel over the qkv rows [rows, W] = rotated q heads | rotated k heads | v heads:
//   given dO [rows, heads*64], the forward's log-sum-exp and D = rowsum(dO * O), produce dK and dV (bf16, into the k and
//   v columns of the gradient rows) and dQ (fp32 reductions into a buffer the caller rounds afterwards).
//
// One CTA = one block of 128 keys of one kv head; it keeps dK and dV of its keys in TENSOR MEMORY across the 4 query heads
// of the GQA group and every query block at or after the key block (no atomics for dK / dV).  576 threads:
//   warp 0      TMA producer: K and V block once; per iteration (query head, 128-query block) the Q and dO tiles (2 stages)
//   warp 1      MMA issuer.  Per iteration, all five products on tcgen05 (M 128):
//                 S^T  = K Q^T        (K-major x K-major, N 128)      dP^T = V dO^T       (same)
//                 dV  += P^T dO       (P^T K-major from shared memory; dO tile read as an MN-major B operand, N 64)
//                 dK  += dS^T Q       (likewise with the Q tile)
//                 dQ   = dS K         (the dS^T tile read as an MN-major A operand, the K tile as an MN-major B operand)
//               S^T / dP^T of iteration i+1 are issued as soon as the softmax threads have pulled iteration i's into registers
//   warps 2..17 four threads per key row: scores and dP out of TMEM, P^T = 2^(scale' s - lse'), dS'^T = P^T (dP^T - D)
//               (the softmax scale is applied to dQ and dK on the way out) written as bf16 tiles in the 128-byte-swizzle
//               layout; the dQ tile of the previous iteration (two accumulators) is read out of TMEM (thread = query row)
//               and added to the fp32 buffer with 16-byte reductions; at the end the dK / dV rows
#include <cuda.h>
#include <string.h>

#include "csm_tc.cuh"

namespace {

constexpr int BB = 128, HDB = 64;
constexpr int TILE_B = BB * HDB * 2;          // a [128 x 64] bf16 tile: 16 KB
constexpr int FB_THREADS = 576;   // TMA warp, MMA warp, 16 softmax warps
constexpr int FB_SMEM = 2 * TILE_B /* K, V */ + 4 * TILE_B /* Q, dO x 2 stages */ + 4 * TILE_B /* P^T, dS^T: 2 halves each */ +
                        1024 /* alignment */ + 256 /* barriers */ + 2 * 128 * 8 /* (lse, D) of the query block, 2 stages */;

struct FlashBwdParams {
  int S, heads, kv, nseq;
  float scale;
  const unsigned char* valid;   // [nseq * S] or null
  const float* lse;             // [nseq * S, heads]
  const float* delta;           // [nseq * S, heads]
  bf16* dqkv;                   // [nseq * S, W]: the k and v columns are written
  float* dq_acc;                // [nseq * S, heads * 64] fp32, zeroed by the caller
};

__global__ void __launch_bounds__(FB_THREADS, 1)
csm_flash_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                        const FlashBwdParams p) {
  extern __shared__ unsigned char fb_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)fb_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sK = smem;
  unsigned char* sV = sK + TILE_B;
  unsigned char* sQ = sV + TILE_B;              // [2]
  unsigned char* sdO = sQ + 2 * TILE_B;         // [2]
  unsigned char* sPt = sdO + 2 * TILE_B;        // [128 keys][128 queries] as two 64-query halves
  unsigned char* sdSt = sPt + 2 * TILE_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdSt + 2 * TILE_B);
  uint64_t* kv_full = bars;
  uint64_t* qd_full = bars + 1;     // 2
  uint64_t* qd_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* p_empty = bars + 8;
  uint64_t* dq_full = bars + 9;     // 2
  uint64_t* dq_empty = bars + 11;   // 2
  uint64_t* fin_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float2* sLD = reinterpret_cast<float2*>(bars + 32);   // [2][128]: (lse * log2 e, D) of the iteration's queries

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvb = blockIdx.x, kvh = blockIdx.y, bl = blockIdx.z;
  const int rep = p.heads / p.kv, nq = p.heads * HDB, nkv = p.kv * HDB, width = nq + 2 * nkv;
  const int row0 = bl * p.S, k0 = kvb * BB;
  const int nqb = (p.S + BB - 1) / BB;
  const int per_head = nqb - kvb, nit = rep * per_head;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 16);
    mbar_init(p_full, 16);
    mbar_init(p_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 16); }
    mbar_init(fin_full, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ft_fence_before();
  __syncthreads();
  ft_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tSt = tmem, tdPt = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TILE_B);
      ft_tma_load(sK, &map_qkv, kv_full, nq + kvh * HDB, row0 + k0);
      ft_tma_load(sV, &map_qkv, kv_full, nq + nkv + kvh * HDB, row0 + k0);
      for (int it = 0; it < nit; ++it) {
        const int st = it & 1, head = kvh * rep + it / per_head, q0 = (kvb + it % per_head) * BB;
        mbar_wait(&qd_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&qd_full[st], 2 * TILE_B);
        ft_tma_load(sQ + st * TILE_B, &map_qkv, &qd_full[st], head * HDB, row0 + q0);
        ft_tma_load(sdO + st * TILE_B, &map_do, &qd_full[st], head * HDB, row0 + q0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idS = ft_idesc(BB, BB, false), idKV = ft_idesc(BB, HDB, true), idQ = ft_idesc(BB, HDB, true, true);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPt = smem_u32(sPt), adSt = smem_u32(sdSt);
      mbar_wait(kv_full, 0);
      auto issue_sp = [&](int it) {
        const int st = it & 1;
        mbar_wait(&qd_full[st], (it >> 1) & 1);
        mbar_wait(s_empty, (it & 1) ^ 1);
        ft_fence_after();
        const uint32_t aQ = smem_u32(sQ + st * TILE_B), adO = smem_u32(sdO + st * TILE_B);
#pragma unroll
        for (int k = 0; k < HDB / 16; ++k)
          ft_umma(tSt, ft_desc_k128(aK + k * 32), ft_desc_k128(aQ + k * 32), idS, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HDB / 16; ++k)
          ft_umma(tdPt, ft_desc_k128(aV + k * 32), ft_desc_k128(adO + k * 32), idS, k != 0 ? 1u : 0u);
        ft_commit(s_full);
      };
      issue_sp(0);
      for (int it = 0; it < nit; ++it) {
        const int st = it & 1;
        if (it + 1 < nit) issue_sp(it + 1);
        const uint32_t aQ = smem_u32(sQ + st * TILE_B), adO = smem_u32(sdO + st * TILE_B);
        mbar_wait(p_full, it & 1);
        ft_fence_after();
#pragma unroll
        for (int k = 0; k < BB / 16; ++k)      // contraction over the 128 queries, 16 per MMA
          ft_umma(tdV, ft_desc_k128(aPt + (k >> 2) * TILE_B + (k & 3) * 32), ft_desc_mn128(adO + k * 2048), idKV,
                  (it | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BB / 16; ++k)
          ft_umma(tdK, ft_desc_k128(adSt + (k >> 2) * TILE_B + (k & 3) * 32), ft_desc_mn128(aQ + k * 2048), idKV,
                  (it | k) != 0 ? 1u : 0u);
        mbar_wait(&dq_empty[st], ((it >> 1) & 1) ^ 1);   // (two dQ accumulators: the read-out of iteration it-1 is not waited for)
        ft_fence_after();
#pragma unroll
        for (int k = 0; k < BB / 16; ++k)      // contraction over the 128 keys: dS^T tile as an MN-major A operand
          ft_umma(tdQ + st * HDB, ft_desc_mn128(adSt + k * 2048, TILE_B), ft_desc_mn128(aK + k * 2048), idQ, k != 0 ? 1u : 0u);
        ft_commit(&dq_full[st]);
        ft_commit(&qd_empty[st]);
        ft_commit(p_empty);
      }
      ft_commit(fin_full);
    }
  } else {
    // Four threads per key row: warps 2..5 own queries 0..31 of every block, 6..9 queries 32..63, ... (a warp may touch
    // TMEM lanes 32 (warp % 4) .. +31 only; the four warps of a lane group split the columns).  The same split applies to
    // the 64 dims when a thread reads a dQ / dK / dV row (16 dims each).
    const int qd = warp & 3;                      // TMEM lanes 32 qd .. +31
    const int qq = (warp - 2) >> 2;               // which 32 queries of the block / which 16 dims
    const int r = qd * 32 + lane;                 // key row of the block (query row when reading the dQ tile)
    const int key = k0 + r;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const float l2e = 1.4426950408889634f, sl2 = p.scale * l2e;
    const int st_tid = threadIdx.x - 64;
    const bool key_ok = key < p.S && (p.valid == nullptr || p.valid[(size_t)row0 + key] != 0);
    unsigned char* prow = sPt + (qq >> 1) * TILE_B + (r >> 3) * 1024 + (r & 7) * 128;
    unsigned char* drow = sdSt + (qq >> 1) * TILE_B + (r >> 3) * 1024 + (r & 7) * 128;
    int head = kvh * rep, qb = kvb;               // iteration (query head, query block), advanced incrementally
    int head_prev = 0, q0_prev = 0;
    auto flush_dq = [&](int itp) {               // dQ tile of iteration itp: thread = query row, 16 dims
      mbar_wait(&dq_full[itp & 1], (itp >> 1) & 1);
      ft_fence_after();
      uint32_t v[16];
      ft_ld16(tdQ + (itp & 1) * HDB + lane_base + qq * 16, v);
      ft_ld_wait();
      ft_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_empty[itp & 1]);
      const int q = q0_prev + r;
      if (q < p.S) {
        float* dst = p.dq_acc + ((size_t)row0 + q) * nq + head_prev * HDB + qq * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "f"(__uint_as_float(v[i]) * p.scale),
                       "f"(__uint_as_float(v[i + 1]) * p.scale), "f"(__uint_as_float(v[i + 2]) * p.scale),
                       "f"(__uint_as_float(v[i + 3]) * p.scale)
                       : "memory");
      }
    };
    auto fetch_ld = [&](int headn, int qbn) {    // (lse * log2 e, D) of query st_tid of a block; past the end: p = 2^-inf = 0
      float2 ld = make_float2(INFINITY, 0.f);
      if (st_tid < 128) {
        const int q = qbn * BB + st_tid;
        if (q < p.S) ld = make_float2(__ldg(p.lse + ((size_t)row0 + q) * p.heads + headn) * l2e,
                                      __ldg(p.delta + ((size_t)row0 + q) * p.heads + headn));
      }
      return ld;
    };
    float2 ld_next = fetch_ld(head, qb);
    for (int it = 0; it < nit; ++it) {
      const int st = it & 1, q0 = qb * BB;
      const bool diag = qb == kvb;
      int head_n = head, qb_n = qb + 1;           // the next iteration
      if (qb_n == nqb) { qb_n = kvb; ++head_n; }
      if (st_tid < 128) sLD[st * 128 + st_tid] = ld_next;   // fetched one iteration ahead
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (it + 1 < nit) ld_next = fetch_ld(head_n, qb_n);
      mbar_wait(s_full, it & 1);
      ft_fence_after();
      uint32_t s[32], d[32];
      ft_ld32(tSt + lane_base + qq * 32, s);
      ft_ld32(tdPt + lane_base + qq * 32, d);
      ft_ld_wait();
      ft_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);        // S^T / dP^T of the next iteration may be issued
      const float2* ldq = sLD + st * 128 + qq * 32;
      uint32_t pp[16], pd[16];                    // this thread's 32 queries of P^T and dS^T as packed bf16 pairs
      if (!diag && key_ok) {                      // nothing to mask (a warp-uniform branch except on padded rows)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 a = ldq[i], b = ldq[i + 1];
          const float pa = ft_ex2(__uint_as_float(s[i]) * sl2 - a.x), pb = ft_ex2(__uint_as_float(s[i + 1]) * sl2 - b.x);
          pp[i >> 1] = pack_bf16(pa, pb);
          pd[i >> 1] = pack_bf16(pa * (__uint_as_float(d[i]) - a.y), pb * (__uint_as_float(d[i + 1]) - b.y));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 a = ldq[i], b = ldq[i + 1];
          const bool va = key_ok && (!diag || r <= qq * 32 + i), vb = key_ok && (!diag || r <= qq * 32 + i + 1);
          const float pa = va ? ft_ex2(__uint_as_float(s[i]) * sl2 - a.x) : 0.f;
          const float pb = vb ? ft_ex2(__uint_as_float(s[i + 1]) * sl2 - b.x) : 0.f;
          pp[i >> 1] = pack_bf16(pa, pb);
          pd[i >> 1] = pack_bf16(pa * (__uint_as_float(d[i]) - a.y), pb * (__uint_as_float(d[i + 1]) - b.y));
        }
      }
      // (only now: the previous iteration's products have read the P^T / dS^T tiles -- they ran during the arithmetic above)
      mbar_wait(p_empty, (it & 1) ^ 1);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int off = (((qq & 1) * 4 + q4) ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + off) = make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
        *reinterpret_cast<uint4*>(drow + off) = make_uint4(pd[4 * q4], pd[4 * q4 + 1], pd[4 * q4 + 2], pd[4 * q4 + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      ft_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (it > 0) flush_dq(it - 1);               // (its products ran while this iteration's tiles were computed)
      head_prev = head;
      q0_prev = q0;
      head = head_n;
      qb = qb_n;
    }
    flush_dq(nit - 1);
    // dK / dV of this key row: 16 dims each.  dS was kept without the softmax scale: dK = scale * (dS'^T Q)
    mbar_wait(fin_full, 0);
    ft_fence_after();
    {
      uint32_t a[16], b[16];
      ft_ld16(tdK + lane_base + qq * 16, a);
      ft_ld16(tdV + lane_base + qq * 16, b);
      ft_ld_wait();
      if (key < p.S) {
        bf16* dk = p.dqkv + ((size_t)row0 + key) * width + nq + kvh * HDB + qq * 16;
        bf16* dv = dk + nkv;
#pragma unroll
        for (int q8 = 0; q8 < 2; ++q8) {
          reinterpret_cast<uint4*>(dk)[q8] =
              make_uint4(pack_bf16(__uint_as_float(a[8 * q8]) * p.scale, __uint_as_float(a[8 * q8 + 1]) * p.scale),
                         pack_bf16(__uint_as_float(a[8 * q8 + 2]) * p.scale, __uint_as_float(a[8 * q8 + 3]) * p.scale),
                         pack_bf16(__uint_as_float(a[8 * q8 + 4]) * p.scale, __uint_as_float(a[8 * q8 + 5]) * p.scale),
                         pack_bf16(__uint_as_float(a[8 * q8 + 6]) * p.scale, __uint_as_float(a[8 * q8 + 7]) * p.scale));
          reinterpret_cast<uint4*>(dv)[q8] =
              make_uint4(pack_bf16(__uint_as_float(b[8 * q8]), __uint_as_float(b[8 * q8 + 1])),
                         pack_bf16(__uint_as_float(b[8 * q8 + 2]), __uint_as_float(b[8 * q8 + 3])),
                         pack_bf16(__uint_as_float(b[8 * q8 + 4]), __uint_as_float(b[8 * q8 + 5])),
                         pack_bf16(__uint_as_float(b[8 * q8 + 6]), __uint_as_float(b[8 * q8 + 7])));
        }
      }
    }
    ft_fence_before();
  }
  ft_fence_before();
  __syncthreads();
  if (warp == 0) {
    ft_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" {
int csm_tmap_2d(void* out, const void* base, long long rows, int K, long long pitch, int box_rows);

// qkv, dqkv [nseq * S, W]; d_out [nseq * S, heads * 64]; lse, delta [nseq * S, heads]; dq_acc fp32 [nseq * S, heads * 64] (zeroed).
cudaError_t csm_flash_tc_bwd_launch(const bf16* qkv, const bf16* d_out, const float* lse, const float* delta, int S, int nseq,
                                    int heads, int kv, float scale, const unsigned char* valid, bf16* dqkv, float* dq_acc,
                                    cudaStream_t st) {
  const int W = (heads + 2 * kv) * HDB, nq = heads * HDB;
  const long long rows = (long long)nseq * S;
  CUtensorMap mq, mdo;
  if (csm_tmap_2d(&mq, qkv, rows, W, W, BB) || csm_tmap_2d(&mdo, d_out, rows, nq, nq, BB)) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void*)csm_flash_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
  if (e != cudaSuccess) return e;
  FlashBwdParams p;
  memset(&p, 0, sizeof p);
  p.S = S; p.heads = heads; p.kv = kv; p.nseq = nseq; p.scale = scale; p.valid = valid; p.lse = lse; p.delta = delta;
  p.dqkv = dqkv; p.dq_acc = dq_acc;
  dim3 grid((S + BB - 1) / BB, kv, nseq);
  csm_flash_tc_bwd_kernel<<<grid, FB_THREADS, FB_SMEM, st>>>(mq, mdo, p);
  return cudaGetLastError();
}
}  // extern "C"
