// Device kernels of the training step (CSMModel.forward(labels=...) + backward, reference modeling_csm.py:292-482):
// everything that is not a dense projection (those run on the tcgen05 GEMM of csm_gemm.cu).  Included by csm_train.cu.
//
//   rope_rows_kernel         apply_rotary_pos_emb on the q and k heads of qkv rows, forward and adjoint
//   flash_fwd_kernel<HD>     causal GQA attention over the qkv rows of whole sequences, saving the log-sum-exp
//   attn_delta_kernel        D = rowsum(dO * O)
//   flash_bwd_kernel<HD>     dK, dV (one CTA per 64-key block and kv head) and dQ (fp32 reductions)
//   swiglu_fwd / swiglu_bwd  LlamaMLP's silu(gate) * up on [R, 2I] rows (gate | up)
//   rmsnorm_bwd_kernel       adjoint of LlamaRMSNorm incl. the residual pass-through and the weight gradient
//   ce_rows_kernel           cross entropy per row: loss and (softmax - onehot) / count written over the logits
//   gathers / scatters       decoder inputs (modeling_csm.py:405-443) and the embedding-table gradients
#pragma once
#include "csm_common.cuh"

namespace {

__device__ __forceinline__ void t_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void t_ldsm_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ uint32_t t_movmatrix(uint32_t a) {   // transpose of an 8x8 b16 matrix held in fragment layout
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void t_cp_async16(void* dst_smem, const void* src, bool pred) {
  const int n = pred ? 16 : 0;   // (src-size 0: the 16 bytes are zero-filled)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------- RoPE on the q and k heads of [R, ld] rows
// hf modeling_llama.py:146-168: x*cos + rotate_half(x)*sin, every product and the sum rounded to bf16.  BWD: the adjoint
// (dx1 = dy1*c + dy2*s, dx2 = dy2*c - dy1*s), one rounding.  Row r is position r % S.  cos/sin: [n_pos][hd/2] bf16.
template <bool BWD>
__global__ void rope_rows_kernel(bf16* __restrict__ x, int ld, int rows, int S, int nheads, int hd,
                                 const bf16* __restrict__ cs, const bf16* __restrict__ sn) {
  const int half = hd >> 1, h8 = half >> 3;     // 8 rotation pairs (two 16-byte vectors of x, one each of cos / sin) per thread
  const long long total = (long long)rows * nheads * h8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % h8) * 8, h = (int)((i / h8) % nheads), r = (int)(i / ((long long)h8 * nheads));
    const int pos = r % S;
    const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(cs + (size_t)pos * half + e));
    const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sn + (size_t)pos * half + e));
    bf16* p = x + (size_t)r * ld + h * hd + e;
    const uint4 a4 = *reinterpret_cast<const uint4*>(p), b4 = *reinterpret_cast<const uint4*>(p + half);
    const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w}, sw[4] = {s4.x, s4.y, s4.z, s4.w};
    const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a0 = bf_lo(aw[q]), a1 = bf_hi(aw[q]), b0 = bf_lo(bw[q]), b1 = bf_hi(bw[q]);
      const float c0 = bf_lo(cw[q]), c1 = bf_hi(cw[q]), s0 = bf_lo(sw[q]), s1 = bf_hi(sw[q]);
      if (!BWD) {
        o1[q] = pack_bf16(bfround(a0 * c0) + bfround(-b0 * s0), bfround(a1 * c1) + bfround(-b1 * s1));
        o2[q] = pack_bf16(bfround(b0 * c0) + bfround(a0 * s0), bfround(b1 * c1) + bfround(a1 * s1));
      } else {
        o1[q] = pack_bf16(a0 * c0 + b0 * s0, a1 * c1 + b1 * s1);
        o2[q] = pack_bf16(b0 * c0 - a0 * s0, b1 * c1 - a1 * s1);
      }
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(p + half) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

// ---------------------------------------------------------------- causal GQA flash attention forward over qkv rows
// grid (ceil(S/128), heads, nseq), 256 threads = 8 warps x 16 query rows; K/V blocks of 64 keys double-buffered with
// cp.async.  qkv row = rotated q heads | rotated k heads | v heads.  Query i of a sequence sees keys <= i that are valid
// (valid == null: all).  Saves lse = log(sum exp(scale * s)) per (row, head); a row that sees nothing: out 0, lse 0.
template <int HD>
__global__ void __launch_bounds__(256, HD == 64 ? 2 : 1) flash_fwd_kernel(const bf16* __restrict__ qkv, int S, int heads, int kv, float scale,
                                                        const unsigned char* __restrict__ valid, bf16* __restrict__ out,
                                                        float* __restrict__ lse) {
  constexpr int BQ = 128, BK = 64, LDS = HD + 8, C8 = HD / 8;
  extern __shared__ __align__(16) unsigned char fsm[];
  bf16* sK = reinterpret_cast<bf16*>(fsm);                 // [2][BK*LDS]
  bf16* sV = sK + 2 * BK * LDS;                            // [2][BK*LDS]
  unsigned char* sOk = reinterpret_cast<unsigned char*>(sV + 2 * BK * LDS);   // [2][BK]
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, bl = blockIdx.z;
  const int kvh = head / (heads / kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nq = heads * HD, nkv = kv * HD, width = nq + 2 * nkv;
  const int q0 = qt * BQ + warp * 16;
  const int r_lo = q0 + g, r_hi = q0 + g + 8;
  const bf16* base = qkv + (size_t)bl * S * width;
  uint32_t qa[HD / 16][4];
  {
    const bf16* qlo = base + (size_t)r_lo * width + head * HD;
    const bf16* qhi = base + (size_t)r_hi * width + head * HD;
#pragma unroll
    for (int kt = 0; kt < HD / 16; ++kt) {
      const int col = kt * 16 + 2 * t;
      qa[kt][0] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qlo + col) : 0u;
      qa[kt][1] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qhi + col) : 0u;
      qa[kt][2] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qlo + col + 8) : 0u;
      qa[kt][3] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qhi + col + 8) : 0u;
    }
  }
  float o[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[j][q] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const int last_key = min(qt * BQ + BQ - 1, S - 1);
  const float sl2 = scale * 1.4426950408889634f;
  const int nblk = last_key / BK + 1;
  const bf16* kbase = base + nq + kvh * HD;
  const bf16* vbase = base + nq + nkv + kvh * HD;
  auto load_block = [&](int kb) {
    const int k0 = kb * BK, buf = kb & 1;
    for (int i = threadIdx.x; i < BK * C8; i += 256) {
      const int kr = i / C8, c8 = i % C8;
      const bool ok = k0 + kr <= last_key;
      const size_t off = (size_t)(ok ? k0 + kr : 0) * width + c8 * 8;
      t_cp_async16(&sK[buf * BK * LDS + kr * LDS + c8 * 8], kbase + off, ok);
      t_cp_async16(&sV[buf * BK * LDS + kr * LDS + c8 * 8], vbase + off, ok);
    }
    if (threadIdx.x < BK) {
      const int kk = k0 + (int)threadIdx.x;
      sOk[buf * BK + threadIdx.x] = kk >= S ? 0 : (valid == nullptr ? 1 : valid[(size_t)bl * S + kk]);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_block(0);
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * BK, buf = kb & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (kb + 1 < nblk) load_block(kb + 1);
    if (k0 > q0 + 15) continue;            // whole block is in this warp's future
    const bf16* cK = sK + buf * BK * LDS;
    const bf16* cV = sV + buf * BK * LDS;
    const unsigned char* cOk = sOk + buf * BK;
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sc[j][q] = 0.f;
#pragma unroll
      for (int kp = 0; kp < HD / 32; ++kp) {
        uint32_t kf[4];
        t_ldsm_x4(kf, cK + (8 * j + (lane & 7)) * LDS + 32 * kp + 8 * (lane >> 3));
        mma16816(sc[j], qa[2 * kp], kf[0], kf[1]);
        mma16816(sc[j], qa[2 * kp + 1], kf[2], kf[3]);
      }
    }
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = k0 + 8 * j + 2 * t;
      const bool ok0 = cOk[8 * j + 2 * t] != 0, ok1 = cOk[8 * j + 2 * t + 1] != 0;
      if (key > r_lo || !ok0) sc[j][0] = -INFINITY;
      if (key + 1 > r_lo || !ok1) sc[j][1] = -INFINITY;
      if (key > r_hi || !ok0) sc[j][2] = -INFINITY;
      if (key + 1 > r_hi || !ok1) sc[j][3] = -INFINITY;
      mx_lo = fmaxf(mx_lo, fmaxf(sc[j][0], sc[j][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(sc[j][2], sc[j][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float c_lo = (mx_lo == -INFINITY) ? 1.f : exp2f((m_lo - mx_lo) * sl2);
    const float c_hi = (mx_hi == -INFINITY) ? 1.f : exp2f((m_hi - mx_hi) * sl2);
    m_lo = mx_lo;
    m_hi = mx_hi;
    l_lo *= c_lo;
    l_hi *= c_hi;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) { o[j][0] *= c_lo; o[j][1] *= c_lo; o[j][2] *= c_hi; o[j][3] *= c_hi; }
    const float ms_lo = (mx_lo == -INFINITY) ? 0.f : mx_lo * sl2, ms_hi = (mx_hi == -INFINITY) ? 0.f : mx_hi * sl2;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(sc[j][0] * sl2 - ms_lo), p1 = exp2f(sc[j][1] * sl2 - ms_lo);
      const float p2 = exp2f(sc[j][2] * sl2 - ms_hi), p3 = exp2f(sc[j][3] * sl2 - ms_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      const int kk = j >> 1;
      if ((j & 1) == 0) { pa[kk][0] = pack_bf16(p0, p1); pa[kk][1] = pack_bf16(p2, p3); }
      else { pa[kk][2] = pack_bf16(p0, p1); pa[kk][3] = pack_bf16(p2, p3); }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jd = 0; jd < HD / 8; ++jd) {
        uint32_t b0r, b1r;
        t_ldsm_x2_trans(b0r, b1r, cV + (16 * kk + (lane & 15)) * LDS + 8 * jd);
        mma16816(o[jd], pa[kk], b0r, b1r);
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, i_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  bf16* olo = out + ((size_t)bl * S + r_lo) * nq + head * HD;
  bf16* ohi = out + ((size_t)bl * S + r_hi) * nq + head * HD;
#pragma unroll
  for (int jd = 0; jd < HD / 8; ++jd) {
    const int col = 8 * jd + 2 * t;
    if (r_lo < S) *reinterpret_cast<uint32_t*>(olo + col) = pack_bf16(o[jd][0] * i_lo, o[jd][1] * i_lo);
    if (r_hi < S) *reinterpret_cast<uint32_t*>(ohi + col) = pack_bf16(o[jd][2] * i_hi, o[jd][3] * i_hi);
  }
  if (t == 0) {
    if (r_lo < S) lse[((size_t)bl * S + r_lo) * heads + head] = l_lo > 0.f ? m_lo * scale + logf(l_lo) : 0.f;
    if (r_hi < S) lse[((size_t)bl * S + r_hi) * heads + head] = l_hi > 0.f ? m_hi * scale + logf(l_hi) : 0.f;
  }
}

// D[r, head] = sum_d dO[r, head, d] * O[r, head, d]: one warp per (row, head)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, int rows, int heads, int hd,
                                  float* __restrict__ delta) {
  const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const size_t off = (size_t)w * hd;     // (row, head) pairs are contiguous: row stride = heads * hd
  float s = 0.f;
  for (int i = lane * 2; i < hd; i += 64) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(o + off + i), b = *reinterpret_cast<const uint32_t*>(d_o + off + i);
    s += bf_lo(a) * bf_lo(b) + bf_hi(a) * bf_hi(b);
  }
  s = warp_sum(s);
  if (lane == 0) delta[w] = s;
}

// ---------------------------------------------------------------- flash attention backward
// grid (ceil(S/64), kv heads, nseq), 128 threads: warp w owns keys 16w..16w+15 of the CTA's 64-key block and keeps
// their dK and dV in registers over all the query heads of the kv group and all query blocks at or after the key block.
// Per (query head, 64-query block): S^T = K Q^T and dP^T = V dO^T as [16 keys x 64 queries] accumulators per warp;
// P^T = exp(scale s - lse), dS^T = P^T (dP^T - D) scale; dV += P^T dO and dK += dS^T Q reuse the accumulator registers
// as A fragments; dQ += dS K takes the transposed 8x8 blocks (movmatrix) and is reduced into an fp32 buffer in global
// memory (several key blocks and four warps contribute to a query row).
template <int HD>
__global__ void __launch_bounds__(128, HD == 64 ? 3 : 1) flash_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_out,
                                                        const float* __restrict__ lse, const float* __restrict__ delta,
                                                        int S, int heads, int kv, float scale,
                                                        const unsigned char* __restrict__ valid, bf16* __restrict__ dqkv,
                                                        float* __restrict__ dq_acc) {
  constexpr int BK = 64, LDS = HD + 8, C8 = HD / 8, LDD = BK + 8, TILE = BK * LDS;
  extern __shared__ __align__(16) unsigned char fsm[];
  bf16* sK = reinterpret_cast<bf16*>(fsm);
  bf16* sV = sK + TILE;
  bf16* sQ = sV + TILE;             // [2][TILE]
  bf16* sdO = sQ + 2 * TILE;        // [2][TILE]
  bf16* sdS = sdO + 2 * TILE;       // [BK keys][LDD]: dS^T of the current block, all four warps' key rows
  float* sLse = reinterpret_cast<float*>(sdS + BK * LDD);     // [2][BK]
  float* sD = sLse + 2 * BK;        // [2][BK]
  unsigned char* sOk = reinterpret_cast<unsigned char*>(sD + 2 * BK);
  const int kvb = blockIdx.x, kvh = blockIdx.y, bl = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rep = heads / kv, nq = heads * HD, nkv = kv * HD, width = nq + 2 * nkv;
  const bf16* base = qkv + (size_t)bl * S * width;
  const int k0 = kvb * BK;
  const int nqb = (S + BK - 1) / BK;
  const int per_head = nqb - kvb, nit = rep * per_head;     // iterations: (query head of the group, query block >= key block)
  const float l2e = 1.4426950408889634f;
  auto load_q = [&](int it) {
    const int head = kvh * rep + it / per_head, q0 = (kvb + it % per_head) * BK, buf = it & 1;
    for (int i = threadIdx.x; i < BK * C8; i += 128) {
      const int qr = i / C8, c8 = i % C8;
      const bool ok = q0 + qr < S;
      const size_t row = (size_t)bl * S + (ok ? q0 + qr : 0);
      t_cp_async16(&sQ[buf * TILE + qr * LDS + c8 * 8], qkv + row * width + head * HD + c8 * 8, ok);
      t_cp_async16(&sdO[buf * TILE + qr * LDS + c8 * 8], d_out + row * nq + head * HD + c8 * 8, ok);
    }
    if (threadIdx.x < BK) {
      const int qq = q0 + (int)threadIdx.x;
      const size_t row = (size_t)bl * S + qq;
      sLse[buf * BK + threadIdx.x] = qq < S ? lse[row * heads + head] * l2e : INFINITY;   // (exp2 domain; past the end: p = 0)
      sD[buf * BK + threadIdx.x] = qq < S ? delta[row * heads + head] : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // K and V block -> shared memory (rows past the sequence are zero-filled and flagged invisible)
  for (int i = threadIdx.x; i < BK * C8; i += 128) {
    const int kr = i / C8, c8 = i % C8;
    const bool ok = k0 + kr < S;
    const size_t off = (size_t)(ok ? k0 + kr : 0) * width + c8 * 8;
    t_cp_async16(&sK[kr * LDS + c8 * 8], base + nq + kvh * HD + off, ok);
    t_cp_async16(&sV[kr * LDS + c8 * 8], base + nq + nkv + kvh * HD + off, ok);
  }
  if (threadIdx.x < BK) {
    const int kk = k0 + (int)threadIdx.x;
    sOk[threadIdx.x] = kk >= S ? 0 : (valid == nullptr ? 1 : valid[(size_t)bl * S + kk]);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_q(0);
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // K / V have landed (the first Q / dO block may still be in flight)
  __syncthreads();
  // A fragments of this warp's 16 keys: K and V rows
  uint32_t ka[HD / 16][4], va[HD / 16][4];
#pragma unroll
  for (int kt = 0; kt < HD / 16; ++kt) {
    const bf16* klo = sK + (16 * warp + g) * LDS + kt * 16 + 2 * t;
    const bf16* vlo = sV + (16 * warp + g) * LDS + kt * 16 + 2 * t;
    ka[kt][0] = *reinterpret_cast<const uint32_t*>(klo);
    ka[kt][1] = *reinterpret_cast<const uint32_t*>(klo + 8 * LDS);
    ka[kt][2] = *reinterpret_cast<const uint32_t*>(klo + 8);
    ka[kt][3] = *reinterpret_cast<const uint32_t*>(klo + 8 * LDS + 8);
    va[kt][0] = *reinterpret_cast<const uint32_t*>(vlo);
    va[kt][1] = *reinterpret_cast<const uint32_t*>(vlo + 8 * LDS);
    va[kt][2] = *reinterpret_cast<const uint32_t*>(vlo + 8);
    va[kt][3] = *reinterpret_cast<const uint32_t*>(vlo + 8 * LDS + 8);
  }
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { dk[j][q] = 0.f; dv[j][q] = 0.f; }
  const int key_lo = k0 + 16 * warp + g, key_hi = key_lo + 8;
  const bool okk_lo = sOk[16 * warp + g] != 0, okk_hi = sOk[16 * warp + g + 8] != 0;
  const float sl2 = scale * l2e;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    const int head = kvh * rep + it / per_head, q0 = (kvb + it % per_head) * BK, buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                        // block `it` has landed; everyone is done with the other buffer and with sdQ's flush
    if (it + 1 < nit) load_q(it + 1);       // in flight while this block is multiplied
    const bf16* cQ = sQ + buf * TILE;
    const bf16* cdO = sdO + buf * TILE;
    const float* cLse = sLse + buf * BK;
    const float* cD = sD + buf * BK;
    // The 64 queries of the block in two halves of 32 (register budget: three CTAs per SM at head dim 64).  Per half:
    // S^T and dP^T as [16 keys x 32 queries] accumulators; P^T = exp(scale s - lse), dS^T = P^T (dP^T - D) scale as bf16
    // A fragments (k16 tile = two query n-tiles); dV += P^T dO and dK += dS^T Q (B fragments: k = query rows, n = head
    // dims -> transposed ldmatrix); dS^T goes to shared memory for the dQ product below.
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      float st[4][4], dp[4][4];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int j = 4 * hf + j4;
#pragma unroll
        for (int q = 0; q < 4; ++q) { st[j4][q] = 0.f; dp[j4][q] = 0.f; }
#pragma unroll
        for (int kp = 0; kp < HD / 32; ++kp) {
          uint32_t f[4];
          t_ldsm_x4(f, cQ + (8 * j + (lane & 7)) * LDS + 32 * kp + 8 * (lane >> 3));
          mma16816(st[j4], ka[2 * kp], f[0], f[1]);
          mma16816(st[j4], ka[2 * kp + 1], f[2], f[3]);
          t_ldsm_x4(f, cdO + (8 * j + (lane & 7)) * LDS + 32 * kp + 8 * (lane >> 3));
          mma16816(dp[j4], va[2 * kp], f[0], f[1]);
          mma16816(dp[j4], va[2 * kp + 1], f[2], f[3]);
        }
      }
      uint32_t pa[2][4], dsa[2][4];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int qi = 8 * (4 * hf + j4) + 2 * t, qq = q0 + qi;
        const float l0 = cLse[qi], l1 = cLse[qi + 1], d0 = cD[qi], d1 = cD[qi + 1];
        const bool m00 = okk_lo && key_lo <= qq, m01 = okk_lo && key_lo <= qq + 1;
        const bool m10 = okk_hi && key_hi <= qq, m11 = okk_hi && key_hi <= qq + 1;
        const float p0 = m00 ? exp2f(st[j4][0] * sl2 - l0) : 0.f, p1 = m01 ? exp2f(st[j4][1] * sl2 - l1) : 0.f;
        const float p2 = m10 ? exp2f(st[j4][2] * sl2 - l0) : 0.f, p3 = m11 ? exp2f(st[j4][3] * sl2 - l1) : 0.f;
        const float s0 = p0 * (dp[j4][0] - d0) * scale, s1 = p1 * (dp[j4][1] - d1) * scale;
        const float s2 = p2 * (dp[j4][2] - d0) * scale, s3 = p3 * (dp[j4][3] - d1) * scale;
        const int kk = j4 >> 1, o2 = (j4 & 1) * 2;
        pa[kk][o2] = pack_bf16(p0, p1);
        pa[kk][o2 + 1] = pack_bf16(p2, p3);
        dsa[kk][o2] = pack_bf16(s0, s1);
        dsa[kk][o2 + 1] = pack_bf16(s2, s3);
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int kk = 2 * hf + k2;
#pragma unroll
        for (int jd = 0; jd < HD / 8; ++jd) {
          uint32_t b0, b1;
          t_ldsm_x2_trans(b0, b1, cdO + (16 * kk + (lane & 15)) * LDS + 8 * jd);
          mma16816(dv[jd], pa[k2], b0, b1);
          t_ldsm_x2_trans(b0, b1, cQ + (16 * kk + (lane & 15)) * LDS + 8 * jd);
          mma16816(dk[jd], dsa[k2], b0, b1);
        }
      }
      bf16* wr = sdS + (16 * warp + g) * LDD + 32 * hf + 2 * t;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int kk = j4 >> 1, o2 = (j4 & 1) * 2;
        *reinterpret_cast<uint32_t*>(wr + 8 * j4) = dsa[kk][o2];
        *reinterpret_cast<uint32_t*>(wr + 8 * LDD + 8 * j4) = dsa[kk][o2 + 1];
      }
    }
    // dQ = dS K needs all 64 keys of a query row: after the exchange through shared memory warp w owns queries
    // 16w..16w+15 (A fragments = transposed 8x8 blocks of the [key][query] tile) and adds its finished rows to the fp32
    // buffer in global memory (other key blocks add to the same rows)
    __syncthreads();
    {
      float dq[HD / 8][4];
#pragma unroll
      for (int jd = 0; jd < HD / 8; ++jd)
#pragma unroll
        for (int q = 0; q < 4; ++q) dq[jd][q] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        if (q0 + 16 * warp + 15 < k0 + 16 * kt) continue;   // every query of this warp precedes every key of the tile
        uint32_t a[4];
        {
          const bf16* p = sdS + (16 * kt + (lane & 7) + ((lane >> 4) & 1) * 8) * LDD + 16 * warp + ((lane >> 3) & 1) * 8;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                       : "r"(smem_u32(p)));
        }
#pragma unroll
        for (int jd = 0; jd < HD / 8; ++jd) {
          uint32_t b0, b1;
          t_ldsm_x2_trans(b0, b1, sK + (16 * kt + (lane & 15)) * LDS + 8 * jd);
          mma16816(dq[jd], a, b0, b1);
        }
      }
      const int ql = q0 + 16 * warp + g, qh = ql + 8;
      float* rl = dq_acc + ((size_t)bl * S + ql) * nq + head * HD + 2 * t;
      float* rh = dq_acc + ((size_t)bl * S + qh) * nq + head * HD + 2 * t;
      if (q0 + 16 * warp + 15 >= k0) {
#pragma unroll
        for (int jd = 0; jd < HD / 8; ++jd) {
          if (ql < S) red_add_v2(rl + 8 * jd, dq[jd][0], dq[jd][1]);
          if (qh < S) red_add_v2(rh + 8 * jd, dq[jd][2], dq[jd][3]);
        }
      }
    }
  }
  // dK, dV of this warp's keys -> the k and v columns of the gradient rows
  bf16* dlo = dqkv + ((size_t)bl * S + key_lo) * width + nq + kvh * HD + 2 * t;
  bf16* dhi = dqkv + ((size_t)bl * S + key_hi) * width + nq + kvh * HD + 2 * t;
#pragma unroll
  for (int jd = 0; jd < HD / 8; ++jd) {
    if (key_lo < S) {
      *reinterpret_cast<uint32_t*>(dlo + 8 * jd) = pack_bf16(dk[jd][0], dk[jd][1]);
      *reinterpret_cast<uint32_t*>(dlo + nkv + 8 * jd) = pack_bf16(dv[jd][0], dv[jd][1]);
    }
    if (key_hi < S) {
      *reinterpret_cast<uint32_t*>(dhi + 8 * jd) = pack_bf16(dk[jd][2], dk[jd][3]);
      *reinterpret_cast<uint32_t*>(dhi + nkv + 8 * jd) = pack_bf16(dv[jd][2], dv[jd][3]);
    }
  }
}

// fp32 [rows, cols] -> bf16 into dst rows of pitch ld
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst, int ld) {
  const long long total = (long long)rows * cols / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const int r = (int)(e / cols), c = (int)(e % cols);
    const float2 v = *reinterpret_cast<const float2*>(src + e);
    *reinterpret_cast<uint32_t*>(dst + (size_t)r * ld + c) = pack_bf16(v.x, v.y);
  }
}

// ---------------------------------------------------------------- SwiGLU on [R, 2I] rows = gate | up
// hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16
__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, int rows, int I, bf16* __restrict__ act) {
  const long long total = (long long)rows * I / 8;     // 8 elements (one 16-byte vector of gate, up and the output) per thread
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const int r = (int)(e / I), c = (int)(e % I);
    const uint4 g4 = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + c);
    const uint4 u4 = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + I + c);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g0 = bf_lo(gw[q]), g1 = bf_hi(gw[q]);
      const float s0 = bfround(g0 / (1.f + expf(-g0))), s1 = bfround(g1 / (1.f + expf(-g1)));
      o[q] = pack_bf16(s0 * bf_lo(uw[q]), s1 * bf_hi(uw[q]));
    }
    *reinterpret_cast<uint4*>(act + (size_t)r * I + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// d gate = d act * up * silu'(gate), d up = d act * silu(gate)   (autograd of F.silu(g) * u on bf16 tensors)
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dact, int rows, int I,
                                  bf16* __restrict__ dgu) {
  const long long total = (long long)rows * I / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const int r = (int)(e / I), c = (int)(e % I);
    const uint4 g4 = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + c);
    const uint4 u4 = *reinterpret_cast<const uint4*>(gu + (size_t)r * 2 * I + I + c);
    const uint4 d4 = *reinterpret_cast<const uint4*>(dact + (size_t)r * I + c);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w}, dw[4] = {d4.x, d4.y, d4.z, d4.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float dg[2], du[2];
      const float gg[2] = {bf_lo(gw[q]), bf_hi(gw[q])}, uu[2] = {bf_lo(uw[q]), bf_hi(uw[q])}, dd[2] = {bf_lo(dw[q]), bf_hi(dw[q])};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float sg = 1.f / (1.f + expf(-gg[k]));
        const float silu = bfround(gg[k] * sg);
        du[k] = dd[k] * silu;
        const float dsilu = bfround(dd[k] * uu[k]);            // gradient reaching silu's output (a bf16 tensor)
        dg[k] = dsilu * (sg * (1.f + gg[k] * (1.f - sg)));
      }
      og[q] = pack_bf16(dg[0], dg[1]);
      ou[q] = pack_bf16(du[0], du[1]);
    }
    *reinterpret_cast<uint4*>(dgu + (size_t)r * 2 * I + c) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dgu + (size_t)r * 2 * I + I + c) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// ---------------------------------------------------------------- adjoint of LlamaRMSNorm (+ residual pass-through)
// y = w * bf16(x * rstd).  With xh = x * rstd and g = dy * w:  dx = rstd * (g - xh * mean(g * xh)).
// dh[r] = (dres ? dres[r] : 0) + bf16(dx)   (the residual stream's gradient, updated in place when dres == dh);
// dw += sum_r dy[r] * bf16(xh[r]) accumulated in fp32 (one atomic per column per block of rows).
// One warp per row, 8 rows per block of 256 threads; H <= 2048.
__global__ void __launch_bounds__(256, 2) rmsnorm_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                          const bf16* dy, const bf16* dres, float eps, int H,
                                                          int rows, int rows_per_block, bf16* dh, float* __restrict__ dw_acc) {
  __shared__ float sdw[2048];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < H; i += 256) sdw[i] = 0.f;
  __syncthreads();
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  float dwl[8][8];   // this lane's columns (i*32 + lane)*8 + q, q < 8
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) dwl[i][q] = 0.f;
  for (int r = r_begin + warp; r < r_end; r += 8) {
    const bf16* xr = x + (size_t)r * H;
    const bf16* dyr = dy + (size_t)r * H;
    uint4 xv[8], gv[8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i * 32 + lane) * 8;
      if (idx < H) {
        xv[i] = *reinterpret_cast<const uint4*>(xr + idx);
        gv[i] = *reinterpret_cast<const uint4*>(dyr + idx);
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&xv[i]);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float a = bf_lo(u[q]), b = bf_hi(u[q]); ss += a * a + b * b; }
      }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)H + eps);
    float dot = 0.f;   // sum g * xh
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i * 32 + lane) * 8;
      if (idx < H) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + idx));
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&xv[i]);
        const uint32_t* gg = reinterpret_cast<const uint32_t*>(&gv[i]);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x0 = bf_lo(u[q]) * rstd, x1 = bf_hi(u[q]) * rstd;
          const float d0 = bf_lo(gg[q]), d1 = bf_hi(gg[q]);
          dot += d0 * bf_lo(ww[q]) * x0 + d1 * bf_hi(ww[q]) * x1;
          dwl[i][2 * q] += d0 * bfround(x0);
          dwl[i][2 * q + 1] += d1 * bfround(x1);
        }
      }
    }
    dot = warp_sum(dot) / (float)H;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i * 32 + lane) * 8;
      if (idx < H) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + idx));
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&xv[i]);
        const uint32_t* gg = reinterpret_cast<const uint32_t*>(&gv[i]);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (dres) rv = *reinterpret_cast<const uint4*>(dres + (size_t)r * H + idx);
        const uint32_t* rr = reinterpret_cast<const uint32_t*>(&rv);
        uint4 o;
        uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x0 = bf_lo(u[q]) * rstd, x1 = bf_hi(u[q]) * rstd;
          const float dx0 = bfround(rstd * (bf_lo(gg[q]) * bf_lo(ww[q]) - x0 * dot));
          const float dx1 = bfround(rstd * (bf_hi(gg[q]) * bf_hi(ww[q]) - x1 * dot));
          ou[q] = pack_bf16(bf_lo(rr[q]) + dx0, bf_hi(rr[q]) + dx1);
        }
        *reinterpret_cast<uint4*>(dh + (size_t)r * H + idx) = o;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = (i * 32 + lane) * 8;
    if (idx < H) {
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(&sdw[idx + q], dwl[i][q]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += 256) atomicAdd(&dw_acc[i], sdw[i]);
}

// ---------------------------------------------------------------- cross entropy over rows of logits
// One warp per row.  label < 0 (ignore_index -100): the row contributes nothing and its gradient is zero.
// In place: logits[row] <- (softmax - onehot) / count for columns < V, 0 for the padding columns [V, ld).
// row_loss[row] = lse - logit[label]  (BF16_LOSS: rounded to bf16, as nn.CrossEntropyLoss on bf16 logits returns
// its per-row log-probabilities in bf16 -- modeling_csm.py:464-467; the codebook-0 loss is taken on float32 logits :389).
template <bool BF16_LOSS>
__global__ void ce_rows_kernel(bf16* __restrict__ logits, int ld, int V, int rows, const int* __restrict__ label,
                               const int* __restrict__ count, float* __restrict__ row_loss) {
  const int r = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  bf16* row = logits + (size_t)r * ld;
  const int lab = label[r];
  if (lab < 0 || lab >= V) {
    for (int i = lane; i < ld; i += 32) row[i] = __float2bfloat16_rn(0.f);
    if (lane == 0) row_loss[r] = 0.f;
    return;
  }
  float mx = -INFINITY;
  for (int i = lane; i < V; i += 32) mx = fmaxf(mx, __bfloat162float(row[i]));
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < V; i += 32) sum += expf(__bfloat162float(row[i]) - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const float inv = 1.f / (float)max(*count, 1);
  const float xt = __bfloat162float(row[lab]);
  for (int i = lane; i < ld; i += 32) {
    float gval = 0.f;
    if (i < V) {
      const float p = expf(__bfloat162float(row[i]) - lse);
      gval = (p - (i == lab ? 1.f : 0.f)) * inv;
    }
    row[i] = __float2bfloat16_rn(gval);
  }
  if (lane == 0) row_loss[r] = BF16_LOSS ? bfround(lse - xt) : (lse - xt);
}
// out[0] = sum(row_loss) / count (0 rows: 0), summed in a fixed order by one block
__global__ void mean_loss_kernel(const float* __restrict__ row_loss, int rows, const int* __restrict__ count, int bf16_out,
                                 float* __restrict__ out) {
  __shared__ float part[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < rows; i += 1024) s += row_loss[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int c = *count;
    float v = c > 0 ? part[0] / (float)c : 0.f;
    out[0] = bf16_out ? bfround(v) : v;
  }
}

// ---------------------------------------------------------------- labels, frame list
// codebook-0 labels with the causal shift (modeling_csm.py:376-384): row (b, t) predicts labels[b, t+1, 0]
__global__ void shift_labels_kernel(const long long* __restrict__ labels, int B, int S, int* __restrict__ row_label,
                                    int* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * S) return;
  const int t = r % S;
  int lab = -100;
  if (t < S - 1) lab = (int)labels[(size_t)(r + 1) * 33];
  row_label[r] = lab;
  if (lab >= 0) atomicAdd(count, 1);
}
// frames whose 32 audio labels are all present (modeling_csm.py:397-399)
__global__ void frame_flag_kernel(const long long* __restrict__ labels, int rows, unsigned char* __restrict__ flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  bool ok = true;
  for (int c = 0; c < 32; ++c) ok = ok && labels[(size_t)r * 33 + c] != -100;
  flag[r] = ok ? 1 : 0;
}
// ordered compaction by one block of 1024 threads: frames[i] = row index of the i-th flagged frame; *n = how many
__global__ void frame_list_kernel(const unsigned char* __restrict__ flag, int rows, int* __restrict__ frames, int* __restrict__ n) {
  __shared__ int cnt[1024];
  const int per = (rows + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(rows, b + per);
  int c = 0;
  for (int r = b; r < e; ++r) c += flag[r];
  cnt[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) { const int v = cnt[i]; cnt[i] = run; run += v; }
    *n = run;
  }
  __syncthreads();
  int o = cnt[threadIdx.x];
  for (int r = b; r < e; ++r)
    if (flag[r]) frames[o++] = r;
}
// decoder targets: row c * F + i -> labels[frame i, c + 1]   (modeling_csm.py:459-461)
__global__ void decoder_labels_kernel(const long long* __restrict__ labels, const int* __restrict__ frames, int F,
                                      int* __restrict__ row_label, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * 31) return;
  const int c = i / F, f = i % F;
  const int lab = (int)labels[(size_t)frames[f] * 33 + c + 1];
  row_label[i] = lab;
  if (lab >= 0) atomicAdd(count, 1);
}

// ---------------------------------------------------------------- decoder inputs (modeling_csm.py:405-443)
// row (i, 0) = backbone hidden of the position before frame i (t = 0 wraps to the last position of the sequence, as
// h[b, t-1] does); rows (i, 1 + c) = audio_embeddings[ids[frame, c] + c * V].  One block per row.
__global__ void decoder_gather_kernel(const bf16* __restrict__ hf, const bf16* __restrict__ audio_emb,
                                      const long long* __restrict__ ids, const int* __restrict__ frames, int S, int V, int H,
                                      bf16* __restrict__ dec_in) {
  const int row = blockIdx.x, i = row / 33, p = row % 33;
  const int fr = frames[i], b = fr / S, t = fr % S;
  const bf16* src;
  if (p == 0) src = hf + ((size_t)b * S + (t == 0 ? S - 1 : t - 1)) * H;
  else src = audio_emb + ((size_t)ids[(size_t)fr * 33 + (p - 1)] + (size_t)(p - 1) * V) * H;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dec_in + (size_t)row * H);
  for (int k = threadIdx.x; k < H / 8; k += blockDim.x) d4[k] = s4[k];
}
// adjoint: row (i, 0) is added to the hidden-state gradient (each position feeds at most one frame), rows (i, 1 + c) to
// the fp32 gradient of the audio embedding table
__global__ void decoder_scatter_kernel(const bf16* __restrict__ d_in, const long long* __restrict__ ids,
                                       const int* __restrict__ frames, int S, int V, int H, bf16* __restrict__ dhf,
                                       float* __restrict__ audio_acc) {
  const int row = blockIdx.x, i = row / 33, p = row % 33;
  const int fr = frames[i], b = fr / S, t = fr % S;
  const bf16* src = d_in + (size_t)row * H;
  if (p == 0) {
    bf16* dst = dhf + ((size_t)b * S + (t == 0 ? S - 1 : t - 1)) * H;
    for (int k = threadIdx.x * 2; k < H; k += blockDim.x * 2) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(src + k), d = *reinterpret_cast<const uint32_t*>(dst + k);
      *reinterpret_cast<uint32_t*>(dst + k) = pack_bf16(bf_lo(a) + bf_lo(d), bf_hi(a) + bf_hi(d));
    }
  } else {
    float* dst = audio_acc + ((size_t)ids[(size_t)fr * 33 + (p - 1)] + (size_t)(p - 1) * V) * H;
    for (int k = threadIdx.x * 4; k < H; k += blockDim.x * 4) {
      const uint2 a = *reinterpret_cast<const uint2*>(src + k);
      red_add_v4(dst + k, bf_lo(a.x), bf_hi(a.x), bf_lo(a.y), bf_hi(a.y));
    }
  }
}
// adjoint of the 33-way masked embedding sum (modeling_csm.py:319-334): every active slot of a token adds the token's
// hidden-state gradient to its table row.  mask == null: all 33 slots.  One block per token row.
__global__ void embed_bwd_kernel(const bf16* __restrict__ dh, const long long* __restrict__ ids, const int* __restrict__ mask,
                                 int V, int H, float* __restrict__ audio_acc, float* __restrict__ text_acc) {
  const int r = blockIdx.x;
  const bf16* src = dh + (size_t)r * H;
  for (int j = 0; j < 33; ++j) {
    if (mask != nullptr && mask[(size_t)r * 33 + j] == 0) continue;
    const long long tok = ids[(size_t)r * 33 + j];
    float* dst = j < 32 ? audio_acc + ((size_t)tok + (size_t)j * V) * H : text_acc + (size_t)tok * H;
    for (int k = threadIdx.x * 4; k < H; k += blockDim.x * 4) {
      const uint2 a = *reinterpret_cast<const uint2*>(src + k);
      red_add_v4(dst + k, bf_lo(a.x), bf_hi(a.x), bf_lo(a.y), bf_hi(a.y));
    }
  }
}

// ---------------------------------------------------------------- layout helpers
// dst[c, r] = src[r, c]: src [rows, cols] with pitch lds, dst [cols, rows] with pitch ldd.  64 x 64 tiles, block (32, 8):
// a warp reads 128 contiguous bytes of a source row and writes 128 contiguous bytes of a destination row (pairs of
// elements per thread on both sides; lds, ldd, rows-pitch even; a ragged edge falls back to single elements).
__global__ void __launch_bounds__(256) transpose_kernel(const bf16* __restrict__ src, int rows, int cols, long long lds,
                                                        bf16* __restrict__ dst, long long ldd) {
  __shared__ bf16 tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool even = ((lds | ldd) & 1) == 0;
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    if (r < rows) {
      if (even && c + 1 < cols) {
        *reinterpret_cast<uint32_t*>(&tile[i][2 * tx]) = *reinterpret_cast<const uint32_t*>(src + (size_t)r * lds + c);
      } else {
        if (c < cols) tile[i][2 * tx] = src[(size_t)r * lds + c];
        if (c + 1 < cols) tile[i][2 * tx + 1] = src[(size_t)r * lds + c + 1];
      }
    }
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;     // destination row c, destination columns r, r + 1
    if (c < cols) {
      if (even && r + 1 < rows) {
        __nv_bfloat162 v;
        v.x = tile[2 * tx][i];
        v.y = tile[2 * tx + 1][i];
        *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)c * ldd + r) = v;
      } else {
        if (r < rows) dst[(size_t)c * ldd + r] = tile[2 * tx][i];
        if (r + 1 < rows) dst[(size_t)c * ldd + r + 1] = tile[2 * tx + 1][i];
      }
    }
  }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, long long n, bf16* __restrict__ dst) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2) {
    if (i + 1 < n) *reinterpret_cast<uint32_t*>(dst + i) = pack_bf16(src[i], src[i + 1]);
    else dst[i] = __float2bfloat16_rn(src[i]);
  }
}

}  // namespace
