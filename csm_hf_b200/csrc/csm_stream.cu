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
// csm_pack.cu stores each CTA's rows contiguously, k16-tile major, 32 bytes per (tile,row) in
// mma fragment order.  The WEIGHTS are the 16-row A operand of mma.sync.m16n8k16 (two
// conflict-free LDS.64 per lane), the batch rows of the activations the 8-column B operand, so a
// batch of 1..8 sequences costs one MMA per 256 weights and nothing is wasted on padding.
//
// Latency rules this file follows (the path is a chain of ~700 dependent phases per frame):
//   * phase descriptors are prefetched into shared memory one phase ahead;
//   * every cross-CTA read is issued as one batch of independent loads (one L2 round trip);
//   * all 8 compute warps take part in activation staging, also for a single sequence;
//   * greedy sampling needs no extra synchronisation: every CTA publishes its best (logit, id)
//     with the head phase's normal barrier and the consumers reduce the 148 candidates themselves.
#include "csm_common.cuh"

namespace {

struct Ctx {
  uint64_t *full, *empty, *afull, *aempty;
  volatile int* sflag;   // [0] last-arriver flag, [1] scratch
  Phase* desc;           // [2] descriptor slots
  float* scratch;        // 256 floats
  int* tok;              // [32] tokens gathered by this phase
  bf16* rope;            // cos_dec | sin_dec ([32][hd/2] each) | cos_bb[pos] | sin_bb[pos]
  float* red;
  unsigned char* actreg;
  unsigned char* ring;
  int tid, warp, lane, c, G;
  uint32_t slot, slot_par, aslot, aslot_par;   // ring positions of the consumer side
  unsigned long long* prof;                    // debug stamps of this phase (thread 0 of the first / last CTA) or null
};

#define CSM_STAMP(cx, i)                      \
  do {                                        \
    if ((cx).prof) (cx).prof[i] = clock64();  \
  } while (0)

__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned target) {
  while (ld_acquire_u32(counter) < target) {
  }
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// ------------------------------------------------------------------ greedy sample of a finished head phase
// sample_topk at topk=1 (modeling_csm.py:179-189) with the canonical lowest-index tie-break: reduce
// the (best logit, id) candidates every CTA published in the head phase for codebook `cb`.
// Result in cx.tok[m]; CTA 0 also records samples / fed.  Ends with a compute_sync.
__device__ __forceinline__ void reduce_candidates(const StreamParams& p, const Ctx& cx, int cb) {
  const int M = p.B;
  for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    float2 pr[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int cc = cx.lane + 32 * j;
      pr[j] = cc < cx.G ? __ldcg(p.cand + (size_t)cc * p.Bmax + m) : make_float2(-INFINITY, __int_as_float(0x7fffffff));
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int oi = __float_as_int(pr[j].y);
      if (better(pr[j].x, oi, best, bi)) { best = pr[j].x; bi = oi; }
    }
    for (int cc = cx.lane + 160; cc < cx.G; cc += 32) {   // grids larger than 160 CTAs (not B200)
      float2 q = __ldcg(p.cand + (size_t)cc * p.Bmax + m);
      const int oi = __float_as_int(q.y);
      if (better(q.x, oi, best, bi)) { best = q.x; bi = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (cx.lane == 0) {
      int fedtok = bi;
      if (p.forced) fedtok = ldcg_i32(p.fed + m * CSM_NQ + cb);
      cx.tok[m] = fedtok;
      if (cx.c == 0) {
        p.samples[m * CSM_NQ + cb] = bi;
        if (!p.forced) p.fed[m * CSM_NQ + cb] = bi;
      }
    }
  }
  compute_sync();
}

// ------------------------------------------------------------------ decoder attention (<= 32 positions, hd 128)
// One warp per (sequence, query head).  Lane t owns cached position t for the
// scores and output dims 4*lane.. for P.V; every load of a stage is issued before its first use.
__device__ __forceinline__ void attn_dec_unit(const StreamParams& p, int layer, int dec_pos, int b, int head, bf16* dst,
                                              int lane) {
  constexpr int HD = 128;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int T = dec_pos + 1;
  {
    const int kvh = head / rep;
    const size_t kvbase = (((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)CSM_DEC_POS * HD;
    const bf16* qp = p.q_dec + (size_t)b * (nh * HD) + head * HD;
    const bf16* kp = p.kc_dec + kvbase + (size_t)lane * HD;
    uint4 kq[HD / 8];
    // q: lane l holds dims 4l..4l+3 (8 bytes), redistributed by shuffles below
    const uint2 qmine = ldcg_u2(qp + lane * 4);
    if (lane < T) {
#pragma unroll
      for (int ci = 0; ci < HD / 8; ++ci) kq[ci] = ldcg_u4(kp + ci * 8);
    } else {
#pragma unroll
      for (int ci = 0; ci < HD / 8; ++ci) kq[ci] = make_uint4(0, 0, 0, 0);
    }
    // V rows (independent of the scores): 8 bytes per lane per position, first half issued with K
    const bf16* vp = p.vc_dec + kvbase + lane * 4;
    uint2 va[CSM_DEC_POS / 2];
#pragma unroll
    for (int t = 0; t < CSM_DEC_POS / 2; ++t) va[t] = t < T ? ldcg_u2(vp + (size_t)t * HD) : make_uint2(0, 0);
    float d = 0.f;
#pragma unroll
    for (int ci = 0; ci < HD / 8; ++ci) {
      // dims 8ci..8ci+7 of q live in lanes 2ci (first 4) and 2ci+1 (last 4)
      const uint32_t q0 = __shfl_sync(0xffffffffu, qmine.x, 2 * ci), q1 = __shfl_sync(0xffffffffu, qmine.y, 2 * ci);
      const uint32_t q2 = __shfl_sync(0xffffffffu, qmine.x, 2 * ci + 1), q3 = __shfl_sync(0xffffffffu, qmine.y, 2 * ci + 1);
      const uint4 kv = kq[ci];
      d += bf_lo(q0) * bf_lo(kv.x) + bf_hi(q0) * bf_hi(kv.x);
      d += bf_lo(q1) * bf_lo(kv.y) + bf_hi(q1) * bf_hi(kv.y);
      d += bf_lo(q2) * bf_lo(kv.z) + bf_hi(q2) * bf_hi(kv.z);
      d += bf_lo(q3) * bf_lo(kv.w) + bf_hi(q3) * bf_hi(kv.w);
    }
    uint2 vb[CSM_DEC_POS / 2];
#pragma unroll
    for (int t = 0; t < CSM_DEC_POS / 2; ++t)
      vb[t] = (t + CSM_DEC_POS / 2) < T ? ldcg_u2(vp + (size_t)(t + CSM_DEC_POS / 2) * HD) : make_uint2(0, 0);
    const float sc = lane < T ? d * p.dec.scale : -INFINITY;
    const float mx = warp_max(sc);
    const float pe = (lane < T) ? __expf(sc - mx) : 0.f;
    const float l = warp_sum(pe);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
    for (int t = 0; t < CSM_DEC_POS / 2; ++t) {
      const float pv = __shfl_sync(0xffffffffu, pe, t);
      o0 += pv * bf_lo(va[t].x); o1 += pv * bf_hi(va[t].x);
      o2 += pv * bf_lo(va[t].y); o3 += pv * bf_hi(va[t].y);
    }
#pragma unroll
    for (int t = 0; t < CSM_DEC_POS / 2; ++t) {
      const float pv = __shfl_sync(0xffffffffu, pe, t + CSM_DEC_POS / 2);
      o0 += pv * bf_lo(vb[t].x); o1 += pv * bf_hi(vb[t].x);
      o2 += pv * bf_lo(vb[t].y); o3 += pv * bf_hi(vb[t].y);
    }
    const float inv = 1.f / l;
    uint2 ov = make_uint2(pack_bf16(o0 * inv, o1 * inv), pack_bf16(o2 * inv, o3 * inv));
    *reinterpret_cast<uint2*>(dst + lane * 4) = ov;
  }
}


// Separate-phase form (larger batches): units spread over the CTAs, result to global memory.
__device__ __forceinline__ void attn_dec_phase(const StreamParams& p, const Phase& P, const Ctx& cx) {
  const int nh = p.dec.heads;
  const int nunits = p.B * nh;
  for (int unit = cx.warp * cx.G + cx.c; unit < nunits; unit += CSM_COMPUTE_WARPS * cx.G) {
    const int b = unit / nh, head = unit - b * nh;
    attn_dec_unit(p, P.layer, P.dec_pos, b, head, p.attn_dec + (size_t)b * (nh * p.dec.hd) + head * p.dec.hd, cx.lane);
  }
}

// ------------------------------------------------------------------ activation staging
// Rows of the phase input -> shared memory [M][K+8] bf16, applying RMSNorm exactly as
// LlamaRMSNorm.forward (hf modeling_llama.py:62-67): fp32 x*rsqrt(mean(x^2)+eps) -> bf16 -> *w -> bf16.
// A row is spread over K/8 threads (one 16-byte load each), 256*8/K rows per pass, so that even
// a single sequence is loaded by all warps with one round trip to L2.
__device__ __forceinline__ void stage_act(const StreamParams& p, const Phase& P, const Ctx& cx, int astride) {
  const int K = P.K, M = p.B;
  const float eps = P.stack ? p.dec.eps : p.bb.eps;
  bf16* dst = reinterpret_cast<bf16*>(cx.actreg);
  if (P.act_mode == ACT_GATHER) reduce_candidates(p, cx, P.cb);
  if (P.act_mode == ACT_ATTN) {
    // decoder attention of every (sequence, head), computed redundantly by every CTA straight into the
    // activation rows of the o_proj that consumes it (small batches only: saves a phase and a barrier)
    const int nh = p.dec.heads;
    for (int unit = cx.warp; unit < M * nh; unit += CSM_COMPUTE_WARPS) {
      const int b = unit / nh, head = unit - b * nh;
      attn_dec_unit(p, P.layer, P.dec_pos, b, head, dst + (size_t)b * astride + head * p.dec.hd, cx.lane);
    }
    return;
  }
  if (K > 2048) {
    // wide plain rows (the MLP activations when they fit in shared memory): straight 16-byte copies
    const int cpr = K >> 3;
    for (int m = 0; m < M; ++m) {
      const uint4* src = reinterpret_cast<const uint4*>(P.act + (size_t)m * P.act_stride);
      uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)m * astride);
      for (int c0 = cx.tid; c0 < cpr; c0 += 4 * CSM_COMPUTE_THREADS) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = c0 + j * CSM_COMPUTE_THREADS;
          if (cc < cpr) v[j] = ldcg_u4(src + cc);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = c0 + j * CSM_COMPUTE_THREADS;
          if (cc < cpr) d4[cc] = v[j];
        }
      }
    }
    return;
  }
  const int tpr = K >> 3;                      // threads per row (K <= 2048 -> <= 256)
  const int rpp = CSM_COMPUTE_THREADS / tpr;   // rows per pass
  const int wpr = tpr >> 5;                    // warps per row (0 when a row is narrower than a warp)
  const int rl = cx.tid / tpr, col = (cx.tid - rl * tpr) * 8;
  uint4 wv = make_uint4(0, 0, 0, 0);
  if (P.act_mode == ACT_NORM && rl < rpp) wv = __ldg(reinterpret_cast<const uint4*>(P.norm_w + col));
  for (int m0 = 0; m0 < M; m0 += rpp) {
    const int m = m0 + rl;
    const bool on = rl < rpp && m < M;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (on) {
      const bf16* src;
      if (P.act_mode == ACT_GATHER)   // _embed_audio (modeling_csm.py:247-259): row tok + codebook*V of the audio table
        src = P.act + (size_t)(cx.tok[m] + P.cb * p.V) * K;
      else
        src = P.act + (size_t)m * P.act_stride;
      v = ldcg_u4(src + col);
    }
    if (P.act_mode == ACT_NORM) {
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float a = bf_lo(u[q]), b = bf_hi(u[q]);
        ss += a * a + b * b;
      }
      // reduce over the threads of the row: inside the warp, then across the row's warps
      if (wpr == 0) {
        for (int o = tpr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      } else {
        ss = warp_sum(ss);
        if (wpr > 1) {
          if (cx.lane == 0) cx.scratch[cx.warp] = ss;
          compute_sync();
          const int w0 = (cx.warp / wpr) * wpr;
          ss = 0.f;
          for (int w = 0; w < wpr; ++w) ss += cx.scratch[w0 + w];
          compute_sync();
        }
      }
      if (on) {
        const float rstd = rsqrtf(ss / (float)K + eps);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&wv);
        uint4 o;
        uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float y0 = bfround(bf_lo(u[q]) * rstd), y1 = bfround(bf_hi(u[q]) * rstd);
          ou[q] = pack_bf16(bf_lo(w[q]) * y0, bf_hi(w[q]) * y1);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)m * astride + col) = o;
        if (P.norm_out != nullptr && (m % cx.G) == cx.c) *reinterpret_cast<uint4*>(P.norm_out + (size_t)m * K + col) = o;
      }
    } else if (on) {
      *reinterpret_cast<uint4*>(dst + (size_t)m * astride + col) = v;
    }
  }
}

// ------------------------------------------------------------------ tensor-core inner loop
// One ring chunk holds this CTA's rows for `tiles` k16-tiles as [tile][k-half][row][8 bf16] (csm_pack.cu), so
// the 16x16 A fragment of an m-tile is ONE ldmatrix.x4 (four conflict-free 8x8 matrices).  The B fragments
// (8 batch rows x 16 k) of two k-tiles come from the activation rows with one more ldmatrix.x4.  Rows past the
// CTA's last weight row and batch rows past M re-read a valid row: they only feed accumulator rows /
// columns that are never stored.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// acc[j][nb] (+)= W[m-tile] x act[n-tile nb] over this warp's k16-tiles (tl0, tl0+ks, ...) of one chunk, two
// k-tiles per iteration.  SINGLE: one m-tile, the two k-tiles of an iteration go to the two accumulator
// sets (two independent MMA chains); otherwise accumulator set j belongs to m-tile j.
template <int NB, bool SINGLE>
__device__ __forceinline__ void mma_chunk(float (&acc)[2][NB][4], uint32_t wa0, uint32_t wa1, const uint32_t (&ab)[NB],
                                          int tiles, int tl0, int ks, uint32_t tile_bytes) {
  const uint32_t wsec = (uint32_t)ks * tile_bytes;        // second k-tile of the iteration
  const uint32_t wstep = 2u * wsec, astep = 2u * (uint32_t)ks * 32u;
  wa0 += (uint32_t)tl0 * tile_bytes;
  wa1 += (uint32_t)tl0 * tile_bytes;
  uint32_t aoff = (uint32_t)tl0 * 32u;
  for (int tl = tl0; tl < tiles; tl += 2 * ks) {
    const bool two = tl + ks < tiles;
    uint32_t aA[4], aB[4], aC[4], aD[4], b[NB][4];
    ldsm_x4(aA, wa0);
    if (!SINGLE) ldsm_x4(aB, wa1);
    if (two) {
      ldsm_x4(aC, wa0 + wsec);
      if (!SINGLE) ldsm_x4(aD, wa1 + wsec);
    }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) ldsm_x4(b[nb], ab[nb] + aoff);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      mma16816(acc[0][nb], aA, b[nb][0], b[nb][1]);
      if (!SINGLE) mma16816(acc[1][nb], aB, b[nb][0], b[nb][1]);
    }
    if (two) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (SINGLE) {
          mma16816(acc[1][nb], aC, b[nb][2], b[nb][3]);
        } else {
          mma16816(acc[0][nb], aC, b[nb][2], b[nb][3]);
          mma16816(acc[1][nb], aD, b[nb][2], b[nb][3]);
        }
      }
    }
    wa0 += wstep;
    wa1 += wstep;
    aoff += astep;
  }
}

// All chunks of one phase for this warp; partial sums -> red[kg][m][rows_pad].
template <int NB>
__device__ __forceinline__ void gemv_core(const StreamParams& p, const Phase& P, Ctx& cx, const Geom& g, bool stream,
                                          int astride, int rows_pad) {
  const int M = p.B;
  float acc[2][NB][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][nb][q] = 0.f;
  const int gq = cx.lane >> 2, tq = cx.lane & 3;
  const int ng = cx.warp & (g.ns - 1), kg = cx.warp / g.ns;
  const int mt0 = ng, mt1 = (ng + g.ns < g.mtiles) ? ng + g.ns : -1;
  const bool active = mt0 < g.mtiles;
  const bool single = mt1 < 0;
  // per-lane ldmatrix row addresses: lane = 8*mat + r
  const int mat = cx.lane >> 3, r8 = cx.lane & 7;
  const uint32_t tile_bytes = (uint32_t)g.rows * 32u;
  int ra0 = 16 * mt0 + (mat & 1) * 8 + r8, ra1 = 16 * (single ? mt0 : mt1) + (mat & 1) * 8 + r8;
  if (ra0 >= g.rows) ra0 = 0;
  if (ra1 >= g.rows) ra1 = 0;
  const uint32_t offA0 = (uint32_t)((mat >> 1) * g.rows + ra0) * 16u;
  const uint32_t offA1 = (uint32_t)((mat >> 1) * g.rows + ra1) * 16u;
  uint32_t offB[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int n = nb * 8 + r8;   // batch rows past M re-read row M-1: their accumulator columns are never stored
    offB[nb] = (uint32_t)((n < M ? n : M - 1) * astride + (mat >> 1) * g.ks * 16 + (mat & 1) * 8) * 2u;
  }
  const uint32_t ring0 = smem_u32(cx.ring), act0 = smem_u32(cx.actreg);

  for (int ch = 0; ch < g.nchunks; ++ch) {
    const int T0 = ch * g.tpc;
    const int tiles = min(g.tpc, g.ntiles - T0);
    const uint32_t s = cx.slot;
    mbar_wait(&cx.full[s], cx.slot_par);
    uint32_t abase = act0;
    uint32_t as = 0;
    if (stream) {
      as = cx.aslot;
      mbar_wait(&cx.afull[as], cx.aslot_par);
      abase += as * (uint32_t)(p.act_region_bytes / 2);
    } else {
      abase += (uint32_t)T0 * 32u;
    }
    if (active) {
      const int tl0 = (kg - T0) & (g.ks - 1);
      const uint32_t wbase = ring0 + s * (uint32_t)p.slot_bytes;
      uint32_t ab[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) ab[nb] = abase + offB[nb];
      if (single) mma_chunk<NB, true>(acc, wbase + offA0, wbase + offA0, ab, tiles, tl0, g.ks, tile_bytes);
      else mma_chunk<NB, false>(acc, wbase + offA0, wbase + offA1, ab, tiles, tl0, g.ks, tile_bytes);
    }
    __syncwarp();
    if (cx.lane == 0) {
      mbar_arrive(&cx.empty[s]);
      if (stream) mbar_arrive(&cx.aempty[as]);
    }
    if (++cx.slot == (uint32_t)p.n_slots) { cx.slot = 0; cx.slot_par ^= 1u; }
    if (stream) { cx.aslot ^= 1u; if (cx.aslot == 0) cx.aslot_par ^= 1u; }
  }
  if (!active) return;
  if (single) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[0][nb][q] += acc[1][nb][q];
  }
  // D fragment: c0,c1 = (weight row g, batch 2t, 2t+1), c2,c3 = (row g+8, same batch columns)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int mt = j == 0 ? mt0 : mt1;
    if (j == 1 && single) break;
    const int row = 16 * mt + gq;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int n0 = nb * 8 + 2 * tq, n1 = n0 + 1;
      float* r0 = cx.red + ((size_t)kg * p.m_alloc + n0) * rows_pad + row;
      float* r1 = cx.red + ((size_t)kg * p.m_alloc + n1) * rows_pad + row;
      if (n0 < M) { r0[0] = acc[j][nb][0]; r0[8] = acc[j][nb][2]; }
      if (n1 < M) { r1[0] = acc[j][nb][1]; r1[8] = acc[j][nb][3]; }
    }
  }
}

// ------------------------------------------------------------------ GEMV / skinny-GEMM phase
template <int NB>
__device__ __forceinline__ void gemv_phase(const StreamParams& p, const Phase& P, Ctx& cx) {
  const int M = p.B, K = P.K;
  const bool stream = (P.act_mode == ACT_STREAM);
  const Geom g = csm_geom(P, cx.c);
  int astride;
  if (!stream) {
    astride = K + 8;
    stage_act(p, P, cx, astride);
    compute_sync();
    CSM_STAMP(cx, 4);   // activations staged
  } else {
    astride = g.tpc * 16 + 8;
  }
  const int rows_pad = g.mtiles * 16 + 4;
  // epilogue mapping: thread -> (batch row m, granule u), granules padded to a power of two
  const int gran = P.gran;
  const int upc = g.rows / gran;
  int up2 = 1, ush = 0;
  while (up2 < upc) { up2 <<= 1; ++ush; }
  const int u = cx.tid & (up2 - 1);
  const int mstep = up2 >= CSM_COMPUTE_THREADS ? 1 : CSM_COMPUTE_THREADS >> ush;
  const int m_first = up2 >= CSM_COMPUTE_THREADS ? 0 : cx.tid >> ush;
  // residual value of the first element this thread will update: fetched now, used after the MMAs
  float resid0 = 0.f;
  if (P.epi == EPI_RESID && u < upc && m_first < M)
    resid0 = ldcg_bf16(P.out + (size_t)m_first * P.out_stride + g.row0 + u);
  if (g.rows > 0) gemv_core<NB>(p, P, cx, g, stream, astride, rows_pad);
  CSM_STAMP(cx, 5);     // this warp's MMAs done
  compute_sync();
  CSM_STAMP(cx, 6);     // all warps' MMAs done

  // ---- fused epilogues
  const StackDims& sd = P.stack ? p.dec : p.bb;
  const int half = sd.hd >> 1;
  if (u < upc) {
    for (int m = m_first; m < M; m += mstep) {
      const int n = u * gran;
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
          const float r = m == m_first ? resid0 : ldcg_bf16(o);
          *o = __float2bfloat16_rn(r + v0);
          break;
        }
        case EPI_SWIGLU: {  // hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16 ; rows (2j,2j+1)=(gate_j,up_j)
          float sl = bfround(v0 / (1.f + expf(-v0)));
          P.out[(size_t)m * P.out_stride + (gn >> 1)] = __float2bfloat16_rn(sl * v1);
          break;
        }
        case EPI_QKV: {     // rows (2j,2j+1) = RoPE pair (i, i+hd/2) of q/k, or two adjacent v features
          const int pidx = gn >> 1;
          const int nq = sd.heads * half, nk = sd.kv * half;
          const int pos = P.stack ? P.dec_pos : p.pos;
          const int cap = P.stack ? CSM_DEC_POS : p.Tcap;
          bf16* kc = P.stack ? p.kc_dec : p.kc_bb;
          bf16* vc = P.stack ? p.vc_dec : p.vc_bb;
          if (pidx < nq + nk) {
            const bool isq = pidx < nq;
            const int pp = isq ? pidx : pidx - nq;
            const int head = pp / half, i = pp - head * half;
            // rope tables staged in shared memory at kernel start: decoder [32][half] cos|sin, backbone row `pos`
            const bf16* ct = P.stack ? cx.rope + pos * half + i : cx.rope + 2 * CSM_DEC_POS * (p.dec.hd >> 1) + i;
            const bf16* st = P.stack ? ct + CSM_DEC_POS * half : ct + half;
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
            const int head = f / sd.hd, d = f - head * sd.hd;
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
  }
  if (P.epi == EPI_HEAD) {
    // publish this CTA's best (logit, id) per sequence; consumers reduce after the barrier
    compute_sync();
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
      if (cx.lane == 0) p.cand[(size_t)cx.c * p.Bmax + m] = make_float2(best, __int_as_float(bi));
    }
  }
}

// ------------------------------------------------------------------ end of frame
// After the last head: sample codebook 31, publish the 32 ids (modeling_csm.py:657-666) and evaluate
// the stop rule torch.all(new_frame == 0) (:662).  CTA 0 only.
__device__ __forceinline__ void finish_phase(const StreamParams& p, const Ctx& cx) {
  if (cx.c != 0) return;
  const int M = p.B;
  reduce_candidates(p, cx, CSM_NQ - 1);
  __threadfence_block();
  if (cx.tid == 0) cx.sflag[1] = 0;
  compute_sync();
  int nz = 0;
  for (int e = cx.tid; e < M * CSM_NQ; e += CSM_COMPUTE_THREADS) {
    const int tok = p.samples[e];   // written by this CTA (this phase or earlier ones of this launch)
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

// ------------------------------------------------------------------ 33-way masked embedding gather-sum
// _embed_tokens + mask multiply + sum (modeling_csm.py:261-282,327-334): fp32 accumulate in slot order,
// one bf16 rounding.  Unit = (sequence, 256-column chunk), one warp each, spread over the CTAs.
__device__ __forceinline__ void embed_phase(const StreamParams& p, const Ctx& cx) {
  const int H = p.bb.H;
  const int nchunk = (H + 255) / 256;
  const int nunits = p.B * nchunk;
  for (int unit = cx.warp * cx.G + cx.c; unit < nunits; unit += CSM_COMPUTE_WARPS * cx.G) {
    const int m = unit / nchunk, col = (unit - m * nchunk) * 256 + cx.lane * 8;
    // lane l holds (id, mask) of slot l; slot 32 (text) is held by every lane
    long long my_tok, txt_tok;
    int my_mk, txt_mk;
    if (p.ids) {
      my_tok = p.ids[m * (CSM_NQ + 1) + cx.lane];
      txt_tok = p.ids[m * (CSM_NQ + 1) + CSM_NQ];
    } else {
      my_tok = (long long)ldcg_i32(p.fed + m * CSM_NQ + cx.lane);
      txt_tok = 0;
    }
    if (p.mask) {
      my_mk = p.mask[m * (CSM_NQ + 1) + cx.lane];
      txt_mk = p.mask[m * (CSM_NQ + 1) + CSM_NQ];
    } else {
      my_mk = 1;
      txt_mk = 0;
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const bool incol = col < H;
#pragma unroll
    for (int s0 = 0; s0 < 33; s0 += 11) {
      uint4 v[11];
      int mk[11];
#pragma unroll
      for (int j = 0; j < 11; ++j) {
        const int slot = s0 + j;
        long long tok;
        if (slot < CSM_NQ) {
          tok = __shfl_sync(0xffffffffu, my_tok, slot);
          mk[j] = __shfl_sync(0xffffffffu, my_mk, slot);
        } else {
          tok = txt_tok;
          mk[j] = txt_mk;
        }
        const bf16* row = slot < CSM_NQ ? p.audio_emb + (size_t)(tok + (long long)slot * p.V) * H
                                        : p.text_emb + (size_t)tok * H;
        v[j] = make_uint4(0, 0, 0, 0);
        if (mk[j] != 0 && incol) v[j] = __ldg(reinterpret_cast<const uint4*>(row + col));
      }
#pragma unroll
      for (int j = 0; j < 11; ++j) {
        if (mk[j] == 0) continue;
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[j]);
        const float f = (float)mk[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += bf_lo(u[i]) * f;
          acc[2 * i + 1] += bf_hi(u[i]) * f;
        }
      }
    }
    if (incol) {
      uint4 o;
      o.x = pack_bf16(acc[0], acc[1]);
      o.y = pack_bf16(acc[2], acc[3]);
      o.z = pack_bf16(acc[4], acc[5]);
      o.w = pack_bf16(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(p.h_bb + (size_t)m * H + col) = o;
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
    const int pbase = sp * CSM_ATT_SPLIT + cx.warp * 16;
    // K and V of this warp's 16 positions: all eight 16-byte loads issued before anything is used
    uint4 kv4[4], vv4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pj = pbase + 4 * j + grp;
      if (pj < Ttot) {
        kv4[j] = ldcg_u4(Kp + (size_t)pj * HD + dl * 8);
        vv4[j] = ldcg_u4(Vp + (size_t)pj * HD + dl * 8);
      } else {
        kv4[j] = make_uint4(0, 0, 0, 0);
        vv4[j] = make_uint4(0, 0, 0, 0);
      }
    }
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
    float s[REP][4];
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
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&vv4[j]);
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

}  // namespace

extern __shared__ __align__(128) unsigned char csm_smem[];

template <int NB, int REP>
__global__ void __launch_bounds__(CSM_THREADS, 1) csm_stream_kernel(const StreamParams p) {
  if (p.stop_flag != nullptr && *p.stop_flag) return;   // generation already ended (set by an earlier launch)

  Ctx cx;
  cx.full = reinterpret_cast<uint64_t*>(csm_smem);
  cx.empty = cx.full + CSM_MAX_SLOTS;
  cx.afull = cx.empty + CSM_MAX_SLOTS;
  cx.aempty = cx.afull + 2;
  cx.sflag = reinterpret_cast<volatile int*>(cx.aempty + 2);
  cx.desc = reinterpret_cast<Phase*>(csm_smem + 256);
  cx.scratch = reinterpret_cast<float*>(csm_smem + 512);
  cx.tok = reinterpret_cast<int*>(csm_smem + 1536);
  cx.rope = reinterpret_cast<bf16*>(csm_smem + CSM_SM_HDR_BYTES);
  cx.red = reinterpret_cast<float*>(csm_smem + CSM_SM_HDR_BYTES + p.rope_bytes);
  cx.actreg = csm_smem + CSM_SM_HDR_BYTES + p.rope_bytes + p.red_bytes;
  cx.ring = cx.actreg + p.act_region_bytes;
  cx.tid = threadIdx.x;
  cx.warp = threadIdx.x >> 5;
  cx.lane = threadIdx.x & 31;
  cx.c = blockIdx.x;
  cx.G = gridDim.x;
  cx.slot = cx.slot_par = cx.aslot = cx.aslot_par = 0;

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
  if (cx.tid < CSM_COMPUTE_THREADS) {
    // rope tables -> shared memory: decoder cos|sin for its 32 positions, backbone row `pos`
    const int hd2 = p.dec.hd >> 1, hb2 = p.bb.hd >> 1;
    const int nd = CSM_DEC_POS * hd2;
    for (int i = cx.tid; i < nd; i += CSM_COMPUTE_THREADS) {
      cx.rope[i] = p.cos_dec[i];
      cx.rope[nd + i] = p.sin_dec[i];
    }
    if (cx.tid < hb2) {
      cx.rope[2 * nd + cx.tid] = p.cos_bb[(size_t)p.pos * hb2 + cx.tid];
      cx.rope[2 * nd + hb2 + cx.tid] = p.sin_bb[(size_t)p.pos * hb2 + cx.tid];
    }
    // first phase descriptor
    if (cx.tid < 8)
      reinterpret_cast<uint4*>(&cx.desc[p.phase_begin & 1])[cx.tid] =
          __ldg(reinterpret_cast<const uint4*>(p.phases + p.phase_begin) + cx.tid);
  }
  __syncthreads();

  if (cx.warp == CSM_COMPUTE_WARPS) {
    // ===================== weight stream producer =====================
    if (cx.lane == 0) {
      uint32_t s = 0, round = 0;
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase P = p.phases[ph];
        if (P.type != PH_GEMV) continue;
        const Geom g = csm_geom(P, cx.c);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2;
        for (int ch = 0; ch < g.nchunks; ++ch) {
          const int tiles = min(g.tpc, g.ntiles - ch * g.tpc);
          const uint32_t bytes = (uint32_t)tiles * g.rows * 32u;
          if (round > 0) mbar_wait(&cx.empty[s], (round - 1u) & 1u);
          mbar_expect_tx(&cx.full[s], bytes);
          bulk_g2s(cx.ring + (size_t)s * p.slot_bytes, src, bytes, &cx.full[s]);
          src += bytes;
          if (++s == (uint32_t)p.n_slots) { s = 0; ++round; }
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
        const Phase P = p.phases[ph];
        if (P.type != PH_GEMV || P.act_mode != ACT_STREAM) continue;
        const Geom g = csm_geom(P, cx.c);
        if (g.nchunks == 0) continue;
        const bf16* actp = P.act;
        const int act_stride = P.act_stride;
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
          const unsigned char* src = reinterpret_cast<const unsigned char*>(actp) + (size_t)ch * g.tpc * 32;
          for (int m = 0; m < p.B; ++m)
            bulk_g2s(dst + (size_t)m * astride_b, src + (size_t)m * act_stride * 2, rowbytes, &cx.afull[s]);
          ++ait;
        }
      }
    }
    return;
  }

  // ===================== compute warps =====================
  for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
    unsigned long long* prof = nullptr;   // debug stamps of the first and the last CTA: [cta][phase][8]
    if (p.prof != nullptr && cx.tid == 0 && (cx.c == 0 || cx.c == cx.G - 1))
      prof = p.prof + ((size_t)(cx.c == 0 ? 0 : 1) * p.n_phases_total + ph) * 8;
    cx.prof = prof;
    if (p.use_barrier && ph > p.phase_begin) {
      if (cx.tid == 0) grid_wait(p.bar_counter, (unsigned)(ph - p.phase_begin) * cx.G);
      if (prof) prof[0] = clock64();     // barrier observed
      compute_sync();
    }
    // descriptor of this phase is in shared memory; fetch the next one while this phase runs
    const Phase P = cx.desc[ph & 1];
    uint4 nxt = make_uint4(0, 0, 0, 0);
    const bool fetch = cx.warp == CSM_COMPUTE_WARPS - 1 && cx.lane < 8 && ph + 1 < p.phase_end;
    if (fetch) nxt = __ldg(reinterpret_cast<const uint4*>(p.phases + ph + 1) + cx.lane);
    if (prof) prof[1] = clock64();       // phase body starts
    switch (P.type) {
      case PH_EMBED: embed_phase(p, cx); break;
      case PH_GEMV: gemv_phase<NB>(p, P, cx); break;
      case PH_ATTN_BB: attn_bb_phase<REP>(p, P, cx); break;
      case PH_ATTN_DEC: attn_dec_phase(p, P, cx); break;
      case PH_FINISH: finish_phase(p, cx); break;
    }
    if (prof) prof[2] = clock64();       // this thread's share of the body done
    if (fetch) reinterpret_cast<uint4*>(&cx.desc[(ph + 1) & 1])[cx.lane] = nxt;
    if (ph + 1 < p.phase_end) {
      compute_sync();
      if (p.use_barrier && cx.tid == 0) {
        // release: everything this CTA wrote (ordered before by the CTA barrier) becomes visible before the count
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar_counter) : "memory");
      }
    }
    if (prof) prof[3] = clock64();       // arrived at the grid barrier
  }
}

// ------------------------------------------------------------------ host launch
// One instantiation per (batch n-tiles, backbone GQA ratio): the kernel then carries a single GEMV and
// attention variant, which keeps its instruction footprint small.
typedef void (*StreamKernel)(const StreamParams);

static StreamKernel pick_kernel(int nb, int rep) {
#define CSM_PICK(NBV)                                                    \
  switch (rep) {                                                         \
    case 1: return csm_stream_kernel<NBV, 1>;                            \
    case 2: return csm_stream_kernel<NBV, 2>;                            \
    default: return csm_stream_kernel<NBV, 4>;                           \
  }
  if (nb <= 1) { CSM_PICK(1) }
  if (nb <= 2) { CSM_PICK(2) }
  CSM_PICK(4)
#undef CSM_PICK
}

extern "C" cudaError_t csm_launch_stream(const StreamParams* p, int grid, size_t smem, cudaStream_t stream,
                                         int cooperative) {
  const int nb = (p->B + 7) / 8, rep = p->bb.heads / p->bb.kv;
  StreamKernel k = pick_kernel(nb, rep);
  cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  if (cooperative) {
    void* args[] = {(void*)p};
    return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(CSM_THREADS), args, smem, stream);
  }
  k<<<grid, CSM_THREADS, smem, stream>>>(*p);
  return cudaGetLastError();
}
