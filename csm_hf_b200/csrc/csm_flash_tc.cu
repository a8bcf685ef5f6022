// Causal GQA flash attention forward on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head dim 64.
//
// Replaces the mma.sync forward of the training step (csm_train_kernels.cuh flash_fwd_kernel<64>) for whole sequences of
// qkv rows [rows, W] = rotated q heads | rotated k heads | v heads (hf sdpa_attention_forward, integrations/
// sdpa_attention.py:40-104, as LlamaAttention.forward calls it, modeling_llama.py:251-289).
//
// One CTA = 128 query rows of one head; key blocks of 128.  192 threads:
//   warp 0      TMA producer: the Q tile once, then K and V blocks into a 2-stage ring (boxes of the qkv matrix itself:
//               K as a K-major [128 keys x 64] tile, V as two MN-major [64 keys x 64 dims] boxes)
//   warp 1      MMA issuer (one lane): S = Q K^T (M 128, N 128, K 64) into one of two TMEM accumulators, and, once the
//               softmax warps have written P, O_blk = P V (M 128, N 64, K 128: P K-major from shared memory, V MN-major);
//               S of block j+1 is issued before P V of block j, so the tensor core works during the softmax
//   warps 2..5  softmax, one thread per query row (= TMEM lane): two passes over the S row in TMEM (row max, then
//               exp2 / row sum / bf16 P written to shared memory in the 128-byte-swizzle layout the MMA reads), then
//               O = O * corr + O_blk in registers (fp32); at the end O / l -> bf16 rows and the log-sum-exp
// Masking: causal inside the diagonal block; keys of padded frames (valid[] == 0) everywhere.
#include <cuda.h>
#include <string.h>

#include "csm_tc.cuh"

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2, K_BYTES = BKV * HD * 2, V_BYTES = BKV * HD * 2, P_BYTES = BQ * BKV * 2;
constexpr int FT_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps
constexpr int FT_SMEM = Q_BYTES + 2 * (K_BYTES + V_BYTES) + P_BYTES + 1024 /* alignment */ + 256 /* barriers, masks */ +
                        4 * 128 * 4 /* row maxima of the two halves, double-buffered */;

struct FlashTcParams {
  int S, heads, kv, nseq;
  // where the K / V rows of (sequence bl, kv head h, position t) live in their matrices:
  //   row = kv_row_base + bl * kv_row_seq + h * kv_row_head + t,   column = k_col / v_col + h * kv_col_head
  // training: rows of the qkv matrix itself; prefill: the KV cache [layer][sequence][kv head][Tcap][64]
  long long kv_row_base;
  int kv_row_seq, kv_row_head, k_col, v_col, kv_col_head;
  float scale;
  const unsigned char* valid;   // [nseq * S] or null
  bf16* out;                    // [nseq * S, heads * 64]
  float* lse;                   // [nseq * S, heads] or null
};

__global__ void __launch_bounds__(FT_THREADS, 1)
csm_flash_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const FlashTcParams p) {
  extern __shared__ unsigned char ft_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)ft_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + Q_BYTES;                 // [2][K_BYTES]
  unsigned char* sV = sK + 2 * K_BYTES;             // [2][V_BYTES]
  unsigned char* sP = sV + 2 * V_BYTES;             // two K-halves of [128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // 2
  uint64_t* kv_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;        // 2
  uint64_t* s_empty = bars + 7;       // 2
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* o_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  uint32_t* sVm = reinterpret_cast<uint32_t*>(bars + 14);            // [2][4] visibility masks of a key block (valid != null)
  float* sMax = reinterpret_cast<float*>(bars + 32);                 // [2 stages][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, bl = blockIdx.z;   // heaviest tiles first
  const int kvh = head / (p.heads / p.kv);
  const int nq = p.heads * HD;
  const int row0 = bl * p.S;                     // first row of this sequence in the qkv matrix
  const int q0 = qt * BQ;
  const int nblk = qt + 1;                       // causal: key blocks 0 .. qt

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ft_fence_before();
  __syncthreads();
  ft_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS0 = tmem, tO = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      ft_tma_load(sQ, &map_q, q_full, head * HD, row0 + q0);
      const int krow = (int)(p.kv_row_base + (long long)bl * p.kv_row_seq + (long long)kvh * p.kv_row_head);
      const int kcol = p.k_col + kvh * p.kv_col_head, vcol = p.v_col + kvh * p.kv_col_head;
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], K_BYTES + V_BYTES);
        ft_tma_load(sK + st * K_BYTES, &map_k, &kv_full[st], kcol, krow + j * BKV);
        ft_tma_load(sV + st * V_BYTES, &map_v, &kv_full[st], vcol, krow + j * BKV);
        ft_tma_load(sV + st * V_BYTES + 8192, &map_v, &kv_full[st], vcol, krow + j * BKV + 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idS = ft_idesc(BQ, BKV, false), idO = ft_idesc(BQ, HD, true);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
        ft_fence_after();
        const uint32_t aK = smem_u32(sK + st * K_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ft_umma(tS0 + st * BKV, ft_desc_k128(aQ + k * 32), ft_desc_k128(aK + k * 32), idS, k != 0 ? 1u : 0u);
        ft_commit(&s_full[st]);
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        ft_fence_after();
        const uint32_t aV = smem_u32(sV + st * V_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)     // 16 keys per MMA: P column block (k / 4 = 64-key half), V k-rows 16k..
          ft_umma(tO, ft_desc_k128(aP + (k >> 2) * (BQ * 128) + (k & 3) * 32), ft_desc_mn128(aV + k * 2048), idO, k != 0 ? 1u : 0u);
        ft_commit(o_full);
        ft_commit(&kv_empty[st]);
        ft_commit(p_empty);
      }
    }
  } else {
    // Two threads per query row: warps 2..5 own keys 0..63 of every block and output dims 0..31, warps 6..9 keys 64..127
    // and dims 32..63 (a warp may touch TMEM lanes 32 (warp % 4) .. +31 only; both halves of a row share those lanes).
    // The halves exchange their row maximum through shared memory once per block; the row sums are combined at the end.
    const int qd = warp & 3;                      // TMEM lanes 32 qd .. 32 qd + 31
    const int hf = (warp - 2) >> 2;               // which half of the keys / output dims
    const int r = qd * 32 + lane;                 // row of the tile = query q0 + r
    const int qrow = q0 + r;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, corr_prev = 1.f;
    const int st_tid = threadIdx.x - 64;          // 0..255 among the softmax threads
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      const bool diag = j == qt;
      // visibility of this thread's 64 keys as two 32-bit masks: padding (valid[] == 0; ballots of the first four
      // softmax warps, one key each, exchanged through shared memory) and the causal limit of the row (diagonal block)
      uint32_t vm0 = 0xffffffffu, vm1 = 0xffffffffu;
      if (p.valid != nullptr) {
        if (st_tid < 128) {
          const int key = j * BKV + st_tid;
          const bool ok = key < p.S && p.valid[(size_t)row0 + key] != 0;
          const uint32_t bl_ = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) sVm[st * 4 + (st_tid >> 5)] = bl_;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        vm0 = sVm[st * 4 + 2 * hf];
        vm1 = sVm[st * 4 + 2 * hf + 1];
      }
      if (diag) {
        const int n0 = qrow - (j * BKV + hf * 64) + 1, n1 = n0 - 32;      // keys of each chunk at or before the query
        vm0 &= n0 >= 32 ? 0xffffffffu : (n0 <= 0 ? 0u : ((1u << n0) - 1u));
        vm1 &= n1 >= 32 ? 0xffffffffu : (n1 <= 0 ? 0u : ((1u << n1) - 1u));
      }
      mbar_wait(&s_full[st], (j >> 1) & 1);
      ft_fence_after();
      const uint32_t tS = tS0 + st * BKV + lane_base + hf * 64;
      // this thread's 64 scores: both TMEM loads in flight together, kept in registers for both passes
      uint32_t v0[32], v1[32];
      ft_ld32(tS, v0);
      ft_ld32(tS + 32, v1);
      ft_ld_wait();
      // pass 1: maximum over this thread's visible keys, then over the row (the other half's through shared memory)
      float mx = -INFINITY;
      if (vm0 == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v0[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((vm0 >> i) & 1u) ? __uint_as_float(v0[i]) : -INFINITY);
      }
      if (vm1 == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v1[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((vm1 >> i) & 1u) ? __uint_as_float(v1[i]) : -INFINITY);
      }
      sMax[(st * 2 + hf) * 128 + r] = mx;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      mx = fmaxf(fmaxf(mx, sMax[(st * 2 + (hf ^ 1)) * 128 + r]), m);
      const float corr = (mx == -INFINITY) ? 1.f : ft_ex2((m - mx) * sl2);   // (m = -inf: ex2(-inf) = 0)
      const float ms = (mx == -INFINITY) ? 0.f : mx * sl2;
      m = mx;
      mbar_wait(p_empty, (j & 1) ^ 1);             // P V of the previous block has read the P tile
      // pass 2: p = 2^(scale' s - scale' max), partial row sum, bf16 P -> this half's [128 x 64] tile (128-byte swizzle)
      float sum = 0.f;
      unsigned char* base = sP + hf * (BQ * 128) + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t w = c == 0 ? vm0 : vm1;
        uint32_t pk[16];
        if (w == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(c == 0 ? v0[i] : v1[i]), s1 = __uint_as_float(c == 0 ? v0[i + 1] : v1[i + 1]);
            const float p0 = ft_ex2(s0 * sl2 - ms), p1 = ft_ex2(s1 * sl2 - ms);
            sum += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(c == 0 ? v0[i] : v1[i]), s1 = __uint_as_float(c == 0 ? v0[i + 1] : v1[i + 1]);
            const float p0 = ((w >> i) & 1u) ? ft_ex2(s0 * sl2 - ms) : 0.f;
            const float p1 = ((w >> (i + 1)) & 1u) ? ft_ex2(s1 * sl2 - ms) : 0.f;
            sum += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = c * 4 + q4;             // 16-byte chunk of the 128-byte row
          *reinterpret_cast<uint4*>(base + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      l = l * corr + sum;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P written here is read by the tensor core (async proxy)
      ft_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[st]);                 // this warp's part of S is free for block j + 2
      }
      // O = O * corr + P V of the PREVIOUS block (this thread's 32 output dims): P V of block j runs on the tensor core
      // while these threads do the softmax of block j + 1, its result is folded in one iteration later
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        ft_fence_after();
        uint32_t v[32];
        ft_ld32(tO + lane_base + hf * 32, v);
        ft_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(v[i]);
        ft_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
      corr_prev = corr;
    }
    {
      mbar_wait(o_full, (nblk - 1) & 1);
      ft_fence_after();
      uint32_t v[32];
      ft_ld32(tO + lane_base + hf * 32, v);
      ft_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_prev + __uint_as_float(v[i]);
      ft_fence_before();
    }
    // the row sum = both halves' partial sums (same running maximum on both sides)
    sMax[hf * 128 + r] = l;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    l += sMax[(hf ^ 1) * 128 + r];
    if (qrow < p.S) {
      const float inv = l > 0.f ? 1.f / l : 0.f;    // a row that saw no key (padded frame): zero output
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)row0 + qrow) * nq + head * HD + hf * 32);
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8)
        dst[q8] = make_uint4(pack_bf16(o[8 * q8] * inv, o[8 * q8 + 1] * inv), pack_bf16(o[8 * q8 + 2] * inv, o[8 * q8 + 3] * inv),
                             pack_bf16(o[8 * q8 + 4] * inv, o[8 * q8 + 5] * inv), pack_bf16(o[8 * q8 + 6] * inv, o[8 * q8 + 7] * inv));
      if (p.lse && hf == 0) p.lse[((size_t)row0 + qrow) * p.heads + head] = l > 0.f ? m * p.scale + logf(l) : 0.f;
    }
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
int csm_tmap_2d_mn(void* out, const void* base, long long k_rows, long long mn, long long pitch);

static cudaError_t ft_launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const FlashTcParams& p,
                             cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void*)csm_flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
  if (e != cudaSuccess) return e;
  dim3 grid((p.S + BQ - 1) / BQ, p.heads, p.nseq);
  csm_flash_tc_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(mq, mk, mv, p);
  return cudaGetLastError();
}

// Training: qkv [nseq * S, W] rows (W = (heads + 2 kv) * 64) hold q, k and v; out [nseq * S, heads * 64],
// lse [nseq * S, heads] (may be null).
cudaError_t csm_flash_tc_launch(const bf16* qkv, int S, int nseq, int heads, int kv, float scale, const unsigned char* valid,
                                bf16* out, float* lse, cudaStream_t st) {
  const int W = (heads + 2 * kv) * HD;
  const long long rows = (long long)nseq * S;
  CUtensorMap mqk, mv;
  if (csm_tmap_2d(&mqk, qkv, rows, W, W, BQ) || csm_tmap_2d_mn(&mv, qkv, rows, W, W)) return cudaErrorInvalidValue;
  FlashTcParams p;
  memset(&p, 0, sizeof p);
  p.S = S; p.heads = heads; p.kv = kv; p.nseq = nseq; p.scale = scale; p.valid = valid; p.out = out; p.lse = lse;
  p.kv_row_base = 0; p.kv_row_seq = S; p.kv_row_head = 0; p.k_col = heads * HD; p.v_col = (heads + kv) * HD; p.kv_col_head = HD;
  return ft_launch(mqk, mqk, mv, p, st);
}

// Context prefill into an empty cache (CSMModel.generate_frame with S > 1, modeling_csm.py:484-552 -> hf LlamaAttention):
// q rows [nseq * S, W] of this group of sequences; K / V already rotated and stored in the cache
// [layer][Bmax][kv][Tcap][64] by the QKV GEMM's epilogue; sequences b0 .. b0 + nseq - 1; out [nseq * S, heads * 64].
cudaError_t csm_flash_tc_prefill_launch(const bf16* qkv, int S, int b0, int nseq, int heads, int kv, const bf16* kc,
                                        const bf16* vc, int layer, int layers, int Bmax, int Tcap, float scale,
                                        const unsigned char* valid, bf16* out, cudaStream_t st) {
  const int W = (heads + 2 * kv) * HD;
  const long long crow = (long long)layers * Bmax * kv * Tcap;
  if (crow >= (1ll << 31)) return cudaErrorInvalidValue;
  CUtensorMap mq, mk, mv;
  if (csm_tmap_2d(&mq, qkv, (long long)nseq * S, W, W, BQ) || csm_tmap_2d(&mk, kc, crow, HD, HD, BKV) ||
      csm_tmap_2d_mn(&mv, vc, crow, HD, HD))
    return cudaErrorInvalidValue;
  FlashTcParams p;
  memset(&p, 0, sizeof p);
  p.S = S; p.heads = heads; p.kv = kv; p.nseq = nseq; p.scale = scale; p.valid = valid; p.out = out; p.lse = nullptr;
  p.kv_row_base = ((long long)layer * Bmax + b0) * kv * Tcap; p.kv_row_seq = kv * Tcap; p.kv_row_head = Tcap;
  p.k_col = 0; p.v_col = 0; p.kv_col_head = 0;
  return ft_launch(mq, mk, mv, p, st);
}
}  // extern "C"
