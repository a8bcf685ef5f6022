// The frame engine: ONE persistent kernel executes a whole audio frame -- backbone decode step,
// codebook-0 head, 32 decoder positions x 4 layers, 31 codebook heads, greedy sampling and the
// embedding gathers between them -- as a table of phases separated by grid-wide barriers.
//
// Replaces, per frame, the ~7000 ATen kernel launches of CSMModel.generate_frame
// (reference modeling_csm.py:484-589 driving hf LlamaModel.forward x32).
//
// Structure of a CTA (one per SM, 148 on B200):
//   warps 0..7  compute: stage activations, tensor-core MMA on weight chunks, fused epilogues
//   warp  8     weight stream: walks the SAME phase table and keeps a ring of shared-memory
//               slots full with this CTA's slice of every weight matrix, using bulk
//               asynchronous copies (TMA engine, SASS UBLKCP) that complete on mbarriers.
//               It never waits on the grid barrier, so HBM keeps streaming the NEXT phases'
//               weights while the compute warps sit in a barrier or an epilogue.
//   warp  9     activation stream for K=8192 (down_proj) phases, whose activations do not
//               fit in shared memory: bulk-copies [B, k-chunk] tiles after the barrier.
//
// Work split of a matrix W[N,K]: rows are divided evenly over the CTAs (granule 1 or 2 rows);
// csm_pack.cu stores each CTA's rows contiguously, k16-tile major, in mma B-fragment order, so
// one bulk copy brings a chunk and one conflict-free LDS.64 per lane feeds an
// mma.sync.m16n8k16 (activations = A operand: batch rows x k; weights = B operand: k x 8 rows).
#include "csm_common.cuh"

namespace {

constexpr int SM_BAR_BYTES = 256;

struct Ctx {
  uint64_t *full, *empty, *afull, *aempty;
  volatile int* sflag;   // [0] last-arriver flag, [1] scratch
  float* red;
  unsigned char* actreg;
  unsigned char* ring;
  int tid, warp, lane, c, G;
};

__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned target) {
  while (ld_acquire_u32(counter) < target) {
  }
}

// ------------------------------------------------------------------ activation staging
// Rows of the phase input -> shared memory [M][K+8] bf16, applying RMSNorm exactly as
// LlamaRMSNorm.forward (hf modeling_llama.py:62-67): fp32 x*rsqrt(mean(x^2)+eps) -> bf16 -> *w -> bf16.
__device__ __forceinline__ void stage_act(const StreamParams& p, const Phase& P, const Ctx& cx, int astride) {
  const int K = P.K, M = p.B;
  const float eps = P.stack ? p.dec.eps : p.bb.eps;
  bf16* dst = reinterpret_cast<bf16*>(cx.actreg);
  for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
    const bf16* src;
    if (P.act_mode == ACT_GATHER) {
      // _embed_audio (modeling_csm.py:247-259): row tok + codebook*V of the shared audio table
      int tok = ldcg_i32(p.fed + m * CSM_NQ + P.cb);
      src = P.act + (size_t)(tok + P.cb * p.V) * K;
    } else {
      src = P.act + (size_t)m * P.act_stride;
    }
    uint4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = (i * 32 + cx.lane) * 8;
      if (idx < K) {
        v[i] = ldcg_u4(src + idx);
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = bf_lo(u[q]), b = bf_hi(u[q]);
          ss += a * a + b * b;
        }
      }
    }
    if (P.act_mode == ACT_NORM) {
      ss = warp_sum(ss);
      const float rstd = rsqrtf(ss / (float)K + eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int idx = (i * 32 + cx.lane) * 8;
        if (idx < K) {
          uint4 wv = __ldg(reinterpret_cast<const uint4*>(P.norm_w + idx));
          const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&wv);
          uint4 o;
          uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float y0 = bfround(bf_lo(u[q]) * rstd), y1 = bfround(bf_hi(u[q]) * rstd);
            ou[q] = pack_bf16(bf_lo(w[q]) * y0, bf_hi(w[q]) * y1);
          }
          *reinterpret_cast<uint4*>(dst + (size_t)m * astride + idx) = o;
          if (P.norm_out != nullptr && (m % cx.G) == cx.c)
            *reinterpret_cast<uint4*>(P.norm_out + (size_t)m * K + idx) = o;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int idx = (i * 32 + cx.lane) * 8;
        if (idx < K) *reinterpret_cast<uint4*>(dst + (size_t)m * astride + idx) = v[i];
      }
    }
  }
}

// ------------------------------------------------------------------ head argmax (greedy sampling)
// sample_topk at topk=1 (modeling_csm.py:179-189) with the canonical lowest-index tie-break.
// Every CTA publishes the best (value, index) of its rows; the last CTA to arrive reduces them.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void head_finish(const StreamParams& p, const Phase& P, const Ctx& cx, const Geom& g,
                                            int rows_pad) {
  const int M = p.B;
  for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = cx.lane; n < g.rows; n += 32) {
      float v = cx.red[(size_t)m * rows_pad + n];
      if (better(v, g.row0 + n, best, bi)) { best = v; bi = g.row0 + n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (cx.lane == 0) p.head_part[(size_t)cx.c * p.Bmax + m] = make_float2(best, __int_as_float(bi));
  }
  compute_sync();
  if (cx.tid == 0) {
    __threadfence();
    unsigned old = atomicAdd(p.head_cnt, 1u);
    cx.sflag[0] = (old == (unsigned)cx.G - 1u);
    cx.sflag[1] = 0;
  }
  compute_sync();
  if (!cx.sflag[0]) return;
  // ---- last CTA: reduce the per-CTA candidates
  __threadfence();
  for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int cc = cx.lane; cc < cx.G; cc += 32) {
      float2 pr = __ldcg(p.head_part + (size_t)cc * p.Bmax + m);
      int oi = __float_as_int(pr.y);
      if (better(pr.x, oi, best, bi)) { best = pr.x; bi = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (cx.lane == 0) {
      p.samples[m * CSM_NQ + P.cb] = bi;
      if (!p.forced) p.fed[m * CSM_NQ + P.cb] = bi;
    }
  }
  if (cx.tid == 0) *p.head_cnt = 0u;
  if (P.cb != CSM_NQ - 1) return;
  // ---- frame complete: publish the 32 ids (modeling_csm.py:657-666), evaluate the stop rule (:662)
  __threadfence();
  compute_sync();
  int nz = 0;
  for (int e = cx.tid; e < M * CSM_NQ; e += CSM_COMPUTE_THREADS) {
    int tok = ldcg_i32(p.samples + e);
    nz |= (tok != 0);
    if (p.out_frames) {
      int m = e / CSM_NQ, q = e % CSM_NQ;
      p.out_frames[(size_t)m * p.out_stride + p.out_off + q] = (long long)tok;
    }
  }
  if (nz) cx.sflag[1] = 1;
  compute_sync();
  if (cx.tid == 0) {
    if (p.stop_on_zeros && !cx.sflag[1]) *p.stop_flag = 1;   // all-zero frame: not kept, generation ends
    else if (p.n_frames) *p.n_frames += 1;
  }
}

// ------------------------------------------------------------------ GEMV / skinny-GEMM phase
__device__ __forceinline__ void gemv_phase(const StreamParams& p, const Phase& P, const Ctx& cx, uint32_t& it,
                                           uint32_t& ait) {
  const int M = p.B, K = P.K;
  const bool stream = (P.act_mode == ACT_STREAM);
  const Geom g = csm_geom(P.N, K, P.gran, cx.G, cx.c, p.slot_bytes, stream ? p.stream_tpc_max : 0x7fffffff);
  const int mt = (p.m_alloc + 15) >> 4;
  int astride;
  if (!stream) {
    astride = K + 8;
    stage_act(p, P, cx, astride);
    compute_sync();
  } else {
    astride = g.tpc * 16 + 8;
  }
  if (g.rows == 0) {
    if (P.epi == EPI_HEAD) head_finish(p, P, cx, g, 8);
    return;
  }
  float acc[2][2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][mi][q] = 0.f;

  const int gq = cx.lane >> 2, tq = cx.lane & 3;
  const int ng = cx.warp % g.ns, kg = cx.warp / g.ns;

  for (int ch = 0; ch < g.nchunks; ++ch) {
    const int T0 = ch * g.tpc;
    const int tiles = min(g.tpc, g.ntiles - T0);
    const uint32_t s = it % (uint32_t)p.n_slots;
    mbar_wait(&cx.full[s], (it / (uint32_t)p.n_slots) & 1u);
    const unsigned char* wslot = cx.ring + (size_t)s * p.slot_bytes;
    const bf16* abase;
    int acol0;
    uint32_t as = 0;
    if (stream) {
      as = ait & 1u;
      mbar_wait(&cx.afull[as], (ait >> 1) & 1u);
      abase = reinterpret_cast<const bf16*>(cx.actreg + (size_t)as * (p.act_region_bytes / 2));
      acol0 = 0;
    } else {
      abase = reinterpret_cast<const bf16*>(cx.actreg);
      acol0 = T0 * 16;
    }
    int tl = (kg - (T0 % g.ks) + g.ks) % g.ks;
    for (; tl < tiles; tl += g.ks) {
      const int col = acol0 + tl * 16 + 2 * tq;
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r0 = mi * 16 + gq, r1 = r0 + 8;
        const bf16* p0 = abase + (size_t)r0 * astride + col;
        const bf16* p1 = abase + (size_t)r1 * astride + col;
        const bool v0 = (mi < mt) && (r0 < M), v1 = (mi < mt) && (r1 < M);
        a[mi][0] = v0 ? *reinterpret_cast<const uint32_t*>(p0) : 0u;
        a[mi][1] = v1 ? *reinterpret_cast<const uint32_t*>(p1) : 0u;
        a[mi][2] = v0 ? *reinterpret_cast<const uint32_t*>(p0 + 8) : 0u;
        a[mi][3] = v1 ? *reinterpret_cast<const uint32_t*>(p1 + 8) : 0u;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = ng + j * g.ns;
        if (i < g.nt) {
          const int row = 8 * i + gq;
          uint2 b = make_uint2(0u, 0u);
          if (row < g.rows) b = *reinterpret_cast<const uint2*>(wslot + ((size_t)tl * g.rows + row) * 32 + tq * 8);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
            if (mi < mt) mma16816(acc[j][mi], a[mi], b.x, b.y);
        }
      }
    }
    __syncwarp();
    if (cx.lane == 0) {
      mbar_arrive(&cx.empty[s]);
      if (stream) mbar_arrive(&cx.aempty[as]);
    }
    ++it;
    if (stream) ++ait;
  }

  // ---- cross-warp (split-K) reduction through shared memory: red[kg][m][n]
  const int rows_pad = g.nt * 8;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = ng + j * g.ns;
    if (i < g.nt) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (mi < mt) {
          const int n = 8 * i + 2 * tq;
          const int m0 = mi * 16 + gq, m1 = m0 + 8;
          if (m0 < M)
            *reinterpret_cast<float2*>(cx.red + ((size_t)kg * p.m_alloc + m0) * rows_pad + n) =
                make_float2(acc[j][mi][0], acc[j][mi][1]);
          if (m1 < M)
            *reinterpret_cast<float2*>(cx.red + ((size_t)kg * p.m_alloc + m1) * rows_pad + n) =
                make_float2(acc[j][mi][2], acc[j][mi][3]);
        }
      }
    }
  }
  compute_sync();

  // ---- fused epilogues
  const int gran = P.gran;
  const int upc = g.rows / gran;
  const StackDims& sd = P.stack ? p.dec : p.bb;
  for (int e = cx.tid; e < M * upc; e += CSM_COMPUTE_THREADS) {
    const int m = e / upc, u = e % upc, n = u * gran;
    float v0 = 0.f, v1 = 0.f;
    for (int kk = 0; kk < g.ks; ++kk) {
      const float* r = cx.red + ((size_t)kk * p.m_alloc + m) * rows_pad + n;
      v0 += r[0];
      if (gran == 2) v1 += r[1];
    }
    v0 = bfround(v0);   // nn.Linear output is bf16
    v1 = bfround(v1);
    const int gn = g.row0 + n;   // packed row index
    switch (P.epi) {
      case EPI_STORE:
        P.out[(size_t)m * P.out_stride + gn] = __float2bfloat16_rn(v0);
        break;
      case EPI_RESID: {   // hf modeling_llama.py:325,331: residual + f(x), both bf16
        bf16* o = P.out + (size_t)m * P.out_stride + gn;
        float r = ldcg_bf16(o);
        *o = __float2bfloat16_rn(r + v0);
        break;
      }
      case EPI_SWIGLU: {  // hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16 ; rows (2j,2j+1)=(gate_j,up_j)
        float sl = bfround(v0 / (1.f + expf(-v0)));
        P.out[(size_t)m * P.out_stride + (gn >> 1)] = __float2bfloat16_rn(sl * v1);
        break;
      }
      case EPI_QKV: {     // rows (2j,2j+1) = RoPE pair (i, i+hd/2) of q/k, or two adjacent v features
        const int half = sd.hd >> 1;
        const int pidx = gn >> 1;
        const int nq = sd.heads * half, nk = sd.kv * half;
        const int pos = P.stack ? P.dec_pos : p.pos;
        const int cap = P.stack ? CSM_DEC_POS : p.Tcap;
        bf16* kc = P.stack ? p.kc_dec : p.kc_bb;
        bf16* vc = P.stack ? p.vc_dec : p.vc_bb;
        if (pidx < nq + nk) {
          const bool isq = pidx < nq;
          const int pp = isq ? pidx : pidx - nq;
          const int head = pp / half, i = pp % half;
          const bf16* ct = (P.stack ? p.cos_dec : p.cos_bb) + (size_t)pos * half + i;
          const bf16* st = (P.stack ? p.sin_dec : p.sin_bb) + (size_t)pos * half + i;
          const float cs = __bfloat162float(*ct), sn = __bfloat162float(*st);
          // apply_rotary_pos_emb (hf modeling_llama.py:146-168): every product and the sum round to bf16
          const float o1 = bfround(bfround(v0 * cs) + bfround(-v1 * sn));
          const float o2 = bfround(bfround(v1 * cs) + bfround(v0 * sn));
          bf16* dstp;
          if (isq) dstp = P.out + (size_t)m * P.out_stride + head * sd.hd + i;
          else dstp = kc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + i;
          dstp[0] = __float2bfloat16_rn(o1);
          dstp[half] = __float2bfloat16_rn(o2);
        } else {
          const int f = (pidx - nq - nk) * 2;
          const int head = f / sd.hd, d = f % sd.hd;
          bf16* dstp = vc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + d;
          *reinterpret_cast<uint32_t*>(dstp) = pack_bf16(v0, v1);
        }
        break;
      }
      case EPI_HEAD: {
        if (P.out) P.out[(size_t)m * P.out_stride + gn] = __float2bfloat16_rn(v0);
        cx.red[(size_t)m * rows_pad + n] = v0;   // kk = 0 plane, own element only
        break;
      }
    }
  }
  if (P.epi == EPI_HEAD) {
    compute_sync();
    head_finish(p, P, cx, g, rows_pad);
  }
}

// ------------------------------------------------------------------ 33-way masked embedding gather-sum
// _embed_tokens + mask multiply + sum (modeling_csm.py:261-282,327-334): fp32 accumulate, one bf16 rounding.
__device__ __forceinline__ void embed_phase(const StreamParams& p, const Ctx& cx) {
  const int H = p.bb.H;
  for (int m = cx.c; m < p.B; m += cx.G) {
    for (int d2 = cx.tid; d2 < H / 2; d2 += CSM_COMPUTE_THREADS) {
      float a0 = 0.f, a1 = 0.f;
      for (int slot = 0; slot <= CSM_NQ; ++slot) {
        int mk = p.mask ? p.mask[m * (CSM_NQ + 1) + slot] : (slot < CSM_NQ ? 1 : 0);
        if (mk == 0) continue;
        long long tok;
        if (p.ids) tok = p.ids[m * (CSM_NQ + 1) + slot];
        else tok = slot < CSM_NQ ? (long long)ldcg_i32(p.fed + m * CSM_NQ + slot) : 0;
        const bf16* row = slot < CSM_NQ ? p.audio_emb + (size_t)(tok + (long long)slot * p.V) * H
                                        : p.text_emb + (size_t)tok * H;
        uint32_t u = __ldg(reinterpret_cast<const unsigned int*>(row) + d2);
        a0 += bf_lo(u) * (float)mk;
        a1 += bf_hi(u) * (float)mk;
      }
      reinterpret_cast<uint32_t*>(p.h_bb + (size_t)m * H)[d2] = pack_bf16(a0, a1);
    }
  }
}

// ------------------------------------------------------------------ backbone decode attention (split-KV, GQA)
// One unit = (sequence, kv-head, 128 cached positions); the 4 (rep) query heads of the group share
// every K/V byte read.  Units write (max, sum, o[64]) partials; the last unit of a (sequence,
// kv-head) merges them -- no extra grid barrier.  Softmax in fp32 (sdpa_attention_forward,
// hf integrations/sdpa_attention.py:40-104; decode step attends to every cached position).
template <int REP>
__device__ __forceinline__ void attn_bb_phase(const StreamParams& p, const Phase& P, const Ctx& cx) {
  constexpr int HD = 64;
  const int Ttot = p.pos + 1;
  const int nsplit = (Ttot + CSM_ATT_SPLIT - 1) / CSM_ATT_SPLIT;
  const int nk = p.bb.kv;
  const int nunits = p.B * nk * nsplit;
  float* sm_o = cx.red;                         // [8 warps][REP][64]
  float* sm_m = cx.red + 8 * REP * HD;          // [8][REP]
  float* sm_l = sm_m + 8 * REP;                 // [8][REP]
  const int grp = cx.lane >> 3, dl = cx.lane & 7;   // 4 positions per load, 8 lanes x 8 dims each
  for (int unit = cx.c; unit < nunits; unit += cx.G) {
    const int sp = unit % nsplit;
    const int kvh = (unit / nsplit) % nk;
    const int b = unit / (nsplit * nk);
    const size_t kvbase = (((size_t)P.layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap * HD;
    const bf16* Kp = p.kc_bb + kvbase;
    const bf16* Vp = p.vc_bb + kvbase;
    // q slice of this lane: REP heads x 8 dims, pre-scaled
    float q[REP][8];
#pragma unroll
    for (int h = 0; h < REP; ++h) {
      uint4 qv = ldcg_u4(p.q_bb + (size_t)b * (p.bb.heads * HD) + (kvh * REP + h) * HD + dl * 8);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&qv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q[h][2 * i] = bf_lo(u[i]) * p.bb.scale;
        q[h][2 * i + 1] = bf_hi(u[i]) * p.bb.scale;
      }
    }
    const int pbase = sp * CSM_ATT_SPLIT + cx.warp * 16;
    float s[REP][4];
    uint4 kv4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pj = pbase + 4 * j + grp;
      if (pj < Ttot) kv4[j] = ldcg_u4(Kp + (size_t)pj * HD + dl * 8);
      else kv4[j] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pj = pbase + 4 * j + grp;
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&kv4[j]);
      float kf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { kf[2 * i] = bf_lo(u[i]); kf[2 * i + 1] = bf_hi(u[i]); }
#pragma unroll
      for (int h = 0; h < REP; ++h) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d += q[h][i] * kf[i];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        s[h][j] = (pj < Ttot) ? d : -INFINITY;
      }
    }
    // V loads issued before the softmax math
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pj = pbase + 4 * j + grp;
      if (pj < Ttot) kv4[j] = ldcg_u4(Vp + (size_t)pj * HD + dl * 8);
      else kv4[j] = make_uint4(0, 0, 0, 0);
    }
    float mx[REP], ls[REP], o[REP][8];
#pragma unroll
    for (int h = 0; h < REP; ++h) {
      float m = fmaxf(fmaxf(s[h][0], s[h][1]), fmaxf(s[h][2], s[h][3]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
      mx[h] = m;
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = (m == -INFINITY) ? 0.f : __expf(s[h][j] - m);
        s[h][j] = pv;
        l += pv;
      }
      l += __shfl_xor_sync(0xffffffffu, l, 8);
      l += __shfl_xor_sync(0xffffffffu, l, 16);
      ls[h] = l;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[h][i] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&kv4[j]);
      float vf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { vf[2 * i] = bf_lo(u[i]); vf[2 * i + 1] = bf_hi(u[i]); }
#pragma unroll
      for (int h = 0; h < REP; ++h)
#pragma unroll
        for (int i = 0; i < 8; ++i) o[h][i] += s[h][j] * vf[i];
    }
#pragma unroll
    for (int h = 0; h < REP; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = o[h][i];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        o[h][i] = v;
      }
    if (grp == 0) {
#pragma unroll
      for (int h = 0; h < REP; ++h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm_o[(cx.warp * REP + h) * HD + dl * 8 + i] = o[h][i];
        if (dl == 0) { sm_m[cx.warp * REP + h] = mx[h]; sm_l[cx.warp * REP + h] = ls[h]; }
      }
    }
    compute_sync();
    // merge the 8 warps: thread (h, d)
    if (cx.tid < REP * HD) {
      const int h = cx.tid / HD, d = cx.tid % HD;
      float Mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) Mx = fmaxf(Mx, sm_m[w * REP + h]);
      float L = 0.f, O = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        float mw = sm_m[w * REP + h];
        float f = (mw == -INFINITY) ? 0.f : __expf(mw - Mx);
        L += f * sm_l[w * REP + h];
        O += f * sm_o[(w * REP + h) * HD + d];
      }
      float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + h) * p.nsplit_max + sp) * (HD + 2);
      part[2 + d] = O;
      if (d == 0) { part[0] = Mx; part[1] = L; }
    }
    compute_sync();
    if (cx.tid == 0) {
      __threadfence();
      unsigned old = atomicAdd(p.attn_cnt + b * nk + kvh, 1u);
      cx.sflag[0] = (old == (unsigned)nsplit - 1u);
    }
    compute_sync();
    if (cx.sflag[0]) {
      __threadfence();
      if (cx.tid < REP * HD) {
        const int h = cx.tid / HD, d = cx.tid % HD;
        const float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + h) * p.nsplit_max) * (HD + 2);
        float Mx = -INFINITY;
        for (int s2 = 0; s2 < nsplit; ++s2) Mx = fmaxf(Mx, ldcg_f32(part + (size_t)s2 * (HD + 2)));
        float L = 0.f, O = 0.f;
        for (int s2 = 0; s2 < nsplit; ++s2) {
          const float* ps = part + (size_t)s2 * (HD + 2);
          float f = __expf(ldcg_f32(ps) - Mx);
          L += f * ldcg_f32(ps + 1);
          O += f * ldcg_f32(ps + 2 + d);
        }
        p.attn_bb[(size_t)b * (p.bb.heads * HD) + (kvh * REP + h) * HD + d] = __float2bfloat16_rn(O / L);
      }
      if (cx.tid == 0) p.attn_cnt[b * nk + kvh] = 0u;
    }
    compute_sync();
  }
}

// ------------------------------------------------------------------ decoder attention (<= 32 positions, hd 128)
__device__ __forceinline__ void attn_dec_phase(const StreamParams& p, const Phase& P, const Ctx& cx) {
  constexpr int HD = 128;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int T = P.dec_pos + 1;
  const int nunits = p.B * nh;
  for (int unit = cx.warp * cx.G + cx.c; unit < nunits; unit += CSM_COMPUTE_WARPS * cx.G) {
    const int b = unit / nh, head = unit % nh, kvh = head / rep;
    const size_t kvbase = (((size_t)P.layer * p.Bmax + b) * nk + kvh) * (size_t)CSM_DEC_POS * HD;
    const bf16* qp = p.q_dec + (size_t)b * (nh * HD) + head * HD;
    float sc = -INFINITY;
    if (cx.lane < T) {
      const bf16* kp = p.kc_dec + kvbase + (size_t)cx.lane * HD;
      float d = 0.f;
#pragma unroll 4
      for (int ci = 0; ci < HD / 8; ++ci) {
        uint4 qv = ldcg_u4(qp + ci * 8);
        uint4 kv = ldcg_u4(kp + ci * 8);
        const uint32_t* qu = reinterpret_cast<const uint32_t*>(&qv);
        const uint32_t* ku = reinterpret_cast<const uint32_t*>(&kv);
#pragma unroll
        for (int i = 0; i < 4; ++i) d += bf_lo(qu[i]) * bf_lo(ku[i]) + bf_hi(qu[i]) * bf_hi(ku[i]);
      }
      sc = d * p.dec.scale;
    }
    const float mx = warp_max(sc);
    const float pe = (cx.lane < T) ? __expf(sc - mx) : 0.f;
    const float l = warp_sum(pe);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float pv = __shfl_sync(0xffffffffu, pe, t);
      uint2 vv = ldcg_u2(p.vc_dec + kvbase + (size_t)t * HD + cx.lane * 4);
      o0 += pv * bf_lo(vv.x); o1 += pv * bf_hi(vv.x);
      o2 += pv * bf_lo(vv.y); o3 += pv * bf_hi(vv.y);
    }
    const float inv = 1.f / l;
    uint2 ov = make_uint2(pack_bf16(o0 * inv, o1 * inv), pack_bf16(o2 * inv, o3 * inv));
    *reinterpret_cast<uint2*>(p.attn_dec + (size_t)b * (nh * HD) + head * HD + cx.lane * 4) = ov;
  }
}

}  // namespace

extern __shared__ __align__(128) unsigned char csm_smem[];

__global__ void __launch_bounds__(CSM_THREADS, 1) csm_stream_kernel(const StreamParams p) {
  if (p.stop_flag != nullptr && *p.stop_flag) return;   // generation already ended (set by an earlier launch)

  Ctx cx;
  cx.full = reinterpret_cast<uint64_t*>(csm_smem);
  cx.empty = cx.full + CSM_MAX_SLOTS;
  cx.afull = cx.empty + CSM_MAX_SLOTS;
  cx.aempty = cx.afull + 2;
  cx.sflag = reinterpret_cast<volatile int*>(cx.aempty + 2);
  cx.red = reinterpret_cast<float*>(csm_smem + SM_BAR_BYTES);
  cx.actreg = csm_smem + SM_BAR_BYTES + p.red_bytes;
  cx.ring = cx.actreg + p.act_region_bytes;
  cx.tid = threadIdx.x;
  cx.warp = threadIdx.x >> 5;
  cx.lane = threadIdx.x & 31;
  cx.c = blockIdx.x;
  cx.G = gridDim.x;

  if (cx.tid == 0) {
    for (int s = 0; s < CSM_MAX_SLOTS; ++s) {
      mbar_init(&cx.full[s], 1);
      mbar_init(&cx.empty[s], CSM_COMPUTE_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&cx.afull[s], 1);
      mbar_init(&cx.aempty[s], CSM_COMPUTE_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (cx.warp == CSM_COMPUTE_WARPS) {
    // ===================== weight stream producer =====================
    if (cx.lane == 0) {
      uint32_t it = 0;
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type != PH_GEMV) continue;
        const Geom g = csm_geom(P.N, P.K, P.gran, cx.G, cx.c, p.slot_bytes,
                                P.act_mode == ACT_STREAM ? p.stream_tpc_max : 0x7fffffff);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2;
        for (int ch = 0; ch < g.nchunks; ++ch) {
          const int tiles = min(g.tpc, g.ntiles - ch * g.tpc);
          const uint32_t bytes = (uint32_t)tiles * g.rows * 32u;
          const uint32_t s = it % (uint32_t)p.n_slots;
          if (it >= (uint32_t)p.n_slots) mbar_wait(&cx.empty[s], ((it / (uint32_t)p.n_slots) - 1u) & 1u);
          mbar_expect_tx(&cx.full[s], bytes);
          bulk_g2s(cx.ring + (size_t)s * p.slot_bytes, src, bytes, &cx.full[s]);
          src += bytes;
          ++it;
        }
      }
    }
    return;
  }
  if (cx.warp == CSM_COMPUTE_WARPS + 1) {
    // ===================== activation stream producer (K=8192 phases) =====================
    if (cx.lane == 0) {
      uint32_t ait = 0;
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type != PH_GEMV || P.act_mode != ACT_STREAM) continue;
        const Geom g = csm_geom(P.N, P.K, P.gran, cx.G, cx.c, p.slot_bytes, p.stream_tpc_max);
        if (g.nchunks == 0) continue;
        if (p.use_barrier && ph > p.phase_begin) grid_wait(p.bar_counter, (unsigned)(ph - p.phase_begin) * cx.G);
        fence_proxy_async();
        const int astride_b = (g.tpc * 16 + 8) * 2;
        for (int ch = 0; ch < g.nchunks; ++ch) {
          const int tiles = min(g.tpc, g.ntiles - ch * g.tpc);
          const uint32_t rowbytes = (uint32_t)tiles * 32u;
          const uint32_t s = ait & 1u;
          if (ait >= 2u) mbar_wait(&cx.aempty[s], ((ait >> 1) - 1u) & 1u);
          mbar_expect_tx(&cx.afull[s], rowbytes * (uint32_t)p.B);
          unsigned char* dst = cx.actreg + (size_t)s * (p.act_region_bytes / 2);
          const unsigned char* src =
              reinterpret_cast<const unsigned char*>(P.act) + (size_t)ch * g.tpc * 32;
          for (int m = 0; m < p.B; ++m)
            bulk_g2s(dst + (size_t)m * astride_b, src + (size_t)m * P.act_stride * 2, rowbytes, &cx.afull[s]);
          ++ait;
        }
      }
    }
    return;
  }

  // ===================== compute warps =====================
  uint32_t it = 0, ait = 0;
  for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
    if (p.use_barrier && ph > p.phase_begin) {
      if (cx.tid == 0) grid_wait(p.bar_counter, (unsigned)(ph - p.phase_begin) * cx.G);
      compute_sync();
    }
    const Phase& P = p.phases[ph];
    if (p.prof != nullptr && cx.c == 0 && cx.tid == 0) p.prof[2 * ph] = clock64();
    switch (P.type) {
      case PH_EMBED: embed_phase(p, cx); break;
      case PH_GEMV: gemv_phase(p, P, cx, it, ait); break;
      case PH_ATTN_BB: {
        const int rep = p.bb.heads / p.bb.kv;
        if (rep == 4) attn_bb_phase<4>(p, P, cx);
        else if (rep == 2) attn_bb_phase<2>(p, P, cx);
        else attn_bb_phase<1>(p, P, cx);
        break;
      }
      case PH_ATTN_DEC: attn_dec_phase(p, P, cx); break;
    }
    if (p.prof != nullptr && cx.c == 0 && cx.tid == 0) p.prof[2 * ph + 1] = clock64();
    if (p.use_barrier && ph + 1 < p.phase_end) {
      compute_sync();
      if (cx.tid == 0) {
        __threadfence();
        atomicAdd(p.bar_counter, 1u);
      }
    }
  }
}

// ------------------------------------------------------------------ host launch
extern "C" cudaError_t csm_launch_stream(const StreamParams* p, int grid, size_t smem, cudaStream_t stream,
                                         int cooperative) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(csm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (cooperative) {
    void* args[] = {(void*)p};
    return cudaLaunchCooperativeKernel((const void*)csm_stream_kernel, dim3(grid), dim3(CSM_THREADS), args, smem, stream);
  }
  csm_stream_kernel<<<grid, CSM_THREADS, smem, stream>>>(*p);
  return cudaGetLastError();
}
