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
//   warp 10     L2 prefetcher: walks the phase table further ahead still and pulls this CTA's weight
//               slices (and norm weights, and the K/V blocks of its attention units) from HBM into L2
//               with cp.async.bulk.prefetch.L2, so HBM keeps streaming even when the ring is full.
//
// Phases are chained by dataflow: every vector that crosses CTAs is an array of tagged words
// (bf16 | 16-bit tag of the producing phase, csm_common.cuh) that the consumer polls while staging, so
// one hand-over costs one trip through L2 -- no fence, no atomic, no separate barrier round.  Only the
// K=8192 phases fed by the TMA engine at batch > 4 still use a grid barrier (plain bf16 input).
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
  volatile unsigned int* sprog;   // bytes of this CTA's weight stream issued so far (read by the prefetcher)
  int ph;                // phase being executed
  int rep;               // which copy of the tagged vectors this CTA reads (c % repl)
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
__device__ __forceinline__ uint32_t tg(const StreamParams& p, int ph) { return (p.tagbase + (uint32_t)ph) & 0xffffu; }

// ------------------------------------------------------------------ greedy sample of a finished head phase
// sample_topk at topk=1 (modeling_csm.py:179-189) with the canonical lowest-index tie-break: reduce
// the (best logit, id) candidates every CTA published in head phase `head_ph` for codebook `cb`
// (tagged 64-bit words: polling them IS the synchronisation with that phase).
// Result in cx.tok[m]; CTA 0 also records samples / fed.  Ends with a compute_sync.
__device__ __forceinline__ void reduce_candidates(const StreamParams& p, const Ctx& cx, int cb, int head_ph) {
  const int M = p.B;
  const unsigned long long tag = tg(p, head_ph);
  for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
    unsigned long long w[5];
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int cc = cx.lane + 32 * j;
        w[j] = 0;
        if (cc < cx.G) {
          w[j] = ld_tag64(p.cand + ((size_t)cx.rep * cx.G + cc) * p.Bmax + m);
          ok &= ((w[j] >> 32) & 0xffffull) == tag;
        }
      }
    } while (!__all_sync(0xffffffffu, ok));
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (cx.lane + 32 * j < cx.G) {
        const int oi = (int)((w[j] >> 16) & 0xffffull);
        const float ov = tw_val((uint32_t)w[j]);
        if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
      }
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
// One warp per (sequence, query head).  Lane t owns cached position t for the scores and output dims
// 4*lane.. for P.V.  q and the K/V of the position being processed come as tagged words straight from
// the qkv phase (polled here); older positions come from the cache, loaded before the poll starts.
// Returns the normalised output dims 4*lane..4*lane+3.
__device__ __forceinline__ void attn_dec_unit(const StreamParams& p, const uint32_t* qbase, int layer, int dec_pos, int b,
                                              int head, uint32_t qtag, int lane, float (&out)[4]) {
  constexpr int HD = 128;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int T = dec_pos + 1;
  const int kvh = head / rep;
  const int W = (nh + 2 * nk) * HD;
  const size_t kvbase = (((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)CSM_DEC_POS * HD;
  const bf16* kp = p.kc_dec + kvbase + (size_t)lane * HD;
  const bf16* vp = p.vc_dec + kvbase + lane * 4;
  uint4 kq[HD / 8];
  if (lane < dec_pos) {
#pragma unroll
    for (int ci = 0; ci < HD / 8; ++ci) kq[ci] = ldcg_u4(kp + ci * 8);
  } else {
#pragma unroll
    for (int ci = 0; ci < HD / 8; ++ci) kq[ci] = make_uint4(0, 0, 0, 0);
  }
  uint2 va[CSM_DEC_POS / 2], vb[CSM_DEC_POS / 2];
#pragma unroll
  for (int t = 0; t < CSM_DEC_POS / 2; ++t) va[t] = t < dec_pos ? ldcg_u2(vp + (size_t)t * HD) : make_uint2(0, 0);
#pragma unroll
  for (int t = 0; t < CSM_DEC_POS / 2; ++t)
    vb[t] = (t + CSM_DEC_POS / 2) < dec_pos ? ldcg_u2(vp + (size_t)(t + CSM_DEC_POS / 2) * HD) : make_uint2(0, 0);
  // q | k | v of this position: lane l holds dims 4l..4l+3 of each
  const uint32_t* qw = qbase + (size_t)b * W + head * HD + lane * 4;
  const uint32_t* kw = qbase + (size_t)b * W + nh * HD + kvh * HD + lane * 4;
  const uint32_t* vw = kw + nk * HD;
  uint4 q4, k4, v4;
  bool ok;
  do {
    q4 = ld_tag4(qw);
    k4 = ld_tag4(kw);
    v4 = ld_tag4(vw);
    ok = tw_ok4(q4, qtag) & tw_ok4(k4, qtag) & tw_ok4(v4, qtag);
  } while (!__all_sync(0xffffffffu, ok));
  const uint2 qmine = make_uint2(tw_pair(q4.x, q4.y), tw_pair(q4.z, q4.w));
  const uint2 kmine = make_uint2(tw_pair(k4.x, k4.y), tw_pair(k4.z, k4.w));
  const uint2 vmine = make_uint2(tw_pair(v4.x, v4.y), tw_pair(v4.z, v4.w));
  float d = 0.f;
#pragma unroll
  for (int ci = 0; ci < HD / 8; ++ci) {
    // dims 8ci..8ci+7 live in lanes 2ci (first 4) and 2ci+1 (last 4)
    const uint32_t q0 = __shfl_sync(0xffffffffu, qmine.x, 2 * ci), q1 = __shfl_sync(0xffffffffu, qmine.y, 2 * ci);
    const uint32_t q2 = __shfl_sync(0xffffffffu, qmine.x, 2 * ci + 1), q3 = __shfl_sync(0xffffffffu, qmine.y, 2 * ci + 1);
    const uint32_t k0 = __shfl_sync(0xffffffffu, kmine.x, 2 * ci), k1 = __shfl_sync(0xffffffffu, kmine.y, 2 * ci);
    const uint32_t k2 = __shfl_sync(0xffffffffu, kmine.x, 2 * ci + 1), k3 = __shfl_sync(0xffffffffu, kmine.y, 2 * ci + 1);
    uint4 kv = kq[ci];
    if (lane == dec_pos) kv = make_uint4(k0, k1, k2, k3);
    d += bf_lo(q0) * bf_lo(kv.x) + bf_hi(q0) * bf_hi(kv.x);
    d += bf_lo(q1) * bf_lo(kv.y) + bf_hi(q1) * bf_hi(kv.y);
    d += bf_lo(q2) * bf_lo(kv.z) + bf_hi(q2) * bf_hi(kv.z);
    d += bf_lo(q3) * bf_lo(kv.w) + bf_hi(q3) * bf_hi(kv.w);
  }
  const float sc = lane < T ? d * p.dec.scale : -INFINITY;
  const float mx = warp_max(sc);
  const float pe = (lane < T) ? __expf(sc - mx) : 0.f;
  const float l = warp_sum(pe);
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
  for (int t = 0; t < CSM_DEC_POS / 2; ++t) {
    const float pv = __shfl_sync(0xffffffffu, pe, t);
    const uint2 vv = (t == dec_pos) ? vmine : va[t];
    o0 += pv * bf_lo(vv.x); o1 += pv * bf_hi(vv.x);
    o2 += pv * bf_lo(vv.y); o3 += pv * bf_hi(vv.y);
  }
#pragma unroll
  for (int t = 0; t < CSM_DEC_POS / 2; ++t) {
    const float pv = __shfl_sync(0xffffffffu, pe, t + CSM_DEC_POS / 2);
    const uint2 vv = (t + CSM_DEC_POS / 2 == dec_pos) ? vmine : vb[t];
    o0 += pv * bf_lo(vv.x); o1 += pv * bf_hi(vv.x);
    o2 += pv * bf_lo(vv.y); o3 += pv * bf_hi(vv.y);
  }
  const float inv = 1.f / l;
  out[0] = o0 * inv; out[1] = o1 * inv; out[2] = o2 * inv; out[3] = o3 * inv;
}

// Separate-phase form: units spread over the CTAs, result published as tagged words.
__device__ __forceinline__ void attn_dec_phase(const StreamParams& p, const Phase& P, const Ctx& cx) {
  const int nh = p.dec.heads;
  const int nunits = p.B * nh;
  const uint32_t qtag = tg(p, P.src_ph), otag = tg(p, cx.ph);
  const size_t qrs = (size_t)p.Bmax * (nh + 2 * p.dec.kv) * p.dec.hd, ors = (size_t)p.Bmax * nh * p.dec.hd;
  for (int unit = cx.warp * cx.G + cx.c; unit < nunits; unit += CSM_COMPUTE_WARPS * cx.G) {
    const int b = unit / nh, head = unit - b * nh;
    float o[4];
    attn_dec_unit(p, p.q_dec + cx.rep * qrs, P.layer, P.dec_pos, b, head, qtag, cx.lane, o);
    st_tag4_r(p.attn_dec + (size_t)b * (nh * p.dec.hd) + head * p.dec.hd + cx.lane * 4, tw_pack(o[0], otag),
              tw_pack(o[1], otag), tw_pack(o[2], otag), tw_pack(o[3], otag), p.repl, ors);
  }
}

// ------------------------------------------------------------------ activation staging
// Rows of the phase input -> shared memory [M][K+8] bf16.  The input is an array of tagged words written
// by the CTAs of phase P.src_ph.  Work item = 4 consecutive words (one 16-byte load); items are dealt
// round-robin to the 256 compute threads, up to 8 loads in flight per thread, repeated until every word
// carries the producer's tag -- the poll IS the load, one trip through L2 after the last producer's store.
// RMSNorm exactly as LlamaRMSNorm.forward (hf modeling_llama.py:62-67): fp32 x*rsqrt(mean(x^2)+eps) ->
// bf16 -> *w -> bf16.  Pass 1 stores the raw rows and per-warp partial sums of squares; after one CTA
// barrier pass 2 scales the thread's own elements in place (fixed summation order: deterministic).
__device__ __forceinline__ void stage_act(const StreamParams& p, const Phase& P, const Ctx& cx, int astride) {
  const int K = P.K, M = p.B;
  bf16* dst = reinterpret_cast<bf16*>(cx.actreg);
  if (P.act_mode == ACT_GATHER) {
    // _embed_audio (modeling_csm.py:247-259): row tok + codebook*V of the audio table (plain bf16, read-only)
    reduce_candidates(p, cx, P.cb, P.res_ph);
    const int gpr = K >> 3;                      // 16-byte groups per row
    const int total = M * gpr;
    for (int i0 = cx.tid; i0 < total; i0 += 8 * CSM_COMPUTE_THREADS) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (i < total) {
          const int m = i / gpr, g = i - m * gpr;
          v[j] = __ldg(reinterpret_cast<const uint4*>(P.act + (size_t)(cx.tok[m] + P.cb * p.V) * K) + g);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (i < total) {
          const int m = i / gpr, g = i - m * gpr;
          *reinterpret_cast<uint4*>(dst + (size_t)m * astride + g * 8) = v[j];
        }
      }
    }
    return;
  }
  const uint32_t tag = tg(p, P.src_ph);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(P.act) + (size_t)cx.rep * p.Bmax * P.act_stride;
  const bool norm = P.act_mode == ACT_NORM;
  const int gpr = K >> 2;                        // 4-word groups per row (multiple of 32: a warp stays inside a row)
  const int total = M * gpr;
  const int ppr = gpr >> 5;                      // warp-sized pieces per row (<= 16 for K <= 2048)
  // norm weights of this thread's (at most two) column groups, requested before the poll starts
  uint2 nw0 = make_uint2(0, 0), nw1 = make_uint2(0, 0);
  if (norm) {
    nw0 = __ldg(reinterpret_cast<const uint2*>(P.norm_w) + (cx.tid % gpr));
    nw1 = __ldg(reinterpret_cast<const uint2*>(P.norm_w) + ((cx.tid + CSM_COMPUTE_THREADS) % gpr));
  }
  bool first = true;
  for (int i0 = cx.tid; i0 < total; i0 += 8 * CSM_COMPUTE_THREADS) {
    uint4 w[8];
    bool ok;
    int iters = 0;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (i < total) {
          const int m = i / gpr, g = i - m * gpr;
          w[j] = ld_tag4(base + (size_t)m * P.act_stride + g * 4);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j * CSM_COMPUTE_THREADS < total) ok &= tw_ok4(w[j], tag);
      ++iters;
    } while (!__all_sync(0xffffffffu, ok));
    if (first && cx.prof) { cx.prof[8] = clock64(); cx.prof[9] = (unsigned long long)iters; }
    first = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * CSM_COMPUTE_THREADS;   // warp-uniform validity and row (gpr % 32 == 0)
      if (i < total) {
        const int m = i / gpr, g = i - m * gpr;
        *reinterpret_cast<uint2*>(dst + (size_t)m * astride + g * 4) =
            make_uint2(tw_pair(w[j].x, w[j].y), tw_pair(w[j].z, w[j].w));
        if (norm) {
          const float a = tw_val(w[j].x), b = tw_val(w[j].y), c = tw_val(w[j].z), d = tw_val(w[j].w);
          const float ss = warp_sum(a * a + b * b + c * c + d * d);
          if (cx.lane == 0) cx.scratch[m * ppr + (g >> 5)] = ss;
        }
      }
    }
  }
  if (!norm) return;
  compute_sync();
  const float eps = P.stack ? p.dec.eps : p.bb.eps;
  int jj = 0;
  for (int i = cx.tid; i < total; i += CSM_COMPUTE_THREADS, ++jj) {
    const int m = i / gpr, g = i - m * gpr;
    float ss = 0.f;
    for (int q = 0; q < ppr; ++q) ss += cx.scratch[m * ppr + q];
    const float rstd = rsqrtf(ss / (float)K + eps);
    const uint2 nw = (gpr > CSM_COMPUTE_THREADS && (jj & 1)) ? nw1 : nw0;
    uint2* px = reinterpret_cast<uint2*>(dst + (size_t)m * astride + g * 4);
    const uint2 x = *px;
    const float y0 = bfround(bf_lo(x.x) * rstd), y1 = bfround(bf_hi(x.x) * rstd);
    const float y2 = bfround(bf_lo(x.y) * rstd), y3 = bfround(bf_hi(x.y) * rstd);
    const uint2 o = make_uint2(pack_bf16(bf_lo(nw.x) * y0, bf_hi(nw.x) * y1), pack_bf16(bf_lo(nw.y) * y2, bf_hi(nw.y) * y3));
    *px = o;
    if (P.norm_out != nullptr && (m % cx.G) == cx.c) *reinterpret_cast<uint2*>(P.norm_out + (size_t)m * K + g * 4) = o;
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
    if (ch == 0) CSM_STAMP(cx, 7);   // first weight chunk of the phase is in shared memory
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
__device__ __forceinline__ float resid_poll(const uint32_t* p, uint32_t tag) {
  uint32_t w = ld_tag(p);
  while ((w >> 16) != tag) w = ld_tag(p);
  return tw_val(w);
}

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
  const uint32_t otag = tg(p, cx.ph);
  uint32_t* outw = reinterpret_cast<uint32_t*>(P.out);
  const size_t ors = (size_t)p.Bmax * P.out_stride;   // words between the copies of the output vector
  const int R = p.repl;
  // residual word of the first element this thread will update: requested now, used after the MMAs
  // (own element of the previous residual phase, or the stream's first value written by another CTA)
  uint32_t resid0 = 0;
  const uint32_t rtag = tg(p, P.res_ph);
  if (P.epi == EPI_RESID && u < upc && m_first < M)
    resid0 = ld_tag(outw + cx.rep * ors + (size_t)m_first * P.out_stride + g.row0 + u);
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
          st_tag_r(outw + (size_t)m * P.out_stride + gn, tw_pack(v0, otag), R, ors);
          break;
        case EPI_RESID: {   // hf modeling_llama.py:325,331: residual + f(x), both bf16
          uint32_t* o = outw + (size_t)m * P.out_stride + gn;
          float r;
          if (m == m_first && (resid0 >> 16) == rtag) r = tw_val(resid0);
          else r = resid_poll(o + cx.rep * ors, rtag);
          st_tag_r(o, tw_pack(r + v0, otag), R, ors);
          break;
        }
        case EPI_SWIGLU: {  // hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16 ; rows (2j,2j+1)=(gate_j,up_j)
          const float sl = bfround(v0 / (1.f + expf(-v0)));
          if (P.flags & CSM_PF_OUT_PLAIN) P.out[(size_t)m * P.out_stride + (gn >> 1)] = __float2bfloat16_rn(sl * v1);
          else st_tag_r(outw + (size_t)m * P.out_stride + (gn >> 1), tw_pack(sl * v1, otag), R, ors);
          break;
        }
        case EPI_QKV: {     // rows (2j,2j+1) = RoPE pair (i, i+hd/2) of q/k, or two adjacent v features
          const int pidx = gn >> 1;
          const int nq = sd.heads * half, nk = sd.kv * half;
          const int pos = P.stack ? P.dec_pos : p.pos;
          const int cap = P.stack ? CSM_DEC_POS : p.Tcap;
          bf16* kc = P.stack ? p.kc_dec : p.kc_bb;
          bf16* vc = P.stack ? p.vc_dec : p.vc_bb;
          uint32_t* qrow = outw + (size_t)m * P.out_stride;   // tagged q | k | v of this position
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
            uint32_t* qd = qrow + (isq ? 0 : sd.heads * sd.hd) + head * sd.hd + i;
            st_tag_r(qd, tw_pack(o1, otag), R, ors);
            st_tag_r(qd + half, tw_pack(o2, otag), R, ors);
            if (!isq) {   // DynamicCache.update (hf cache_utils.py:102-121) as an in-place write at `pos`
              bf16* dstp = kc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + i;
              dstp[0] = __float2bfloat16_rn(o1);
              dstp[half] = __float2bfloat16_rn(o2);
            }
          } else {
            const int f = (pidx - nq - nk) * 2;
            const int head = f / sd.hd, d = f - head * sd.hd;
            st_tag2_r(qrow + (sd.heads + sd.kv) * sd.hd + f, tw_pack(v0, otag), tw_pack(v1, otag), R, ors);
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
    // publish this CTA's best (logit, id) per sequence as one tagged 64-bit word; consumers poll and reduce
    compute_sync();
    for (int m = cx.warp; m < M; m += CSM_COMPUTE_WARPS) {
      float best = -INFINITY;
      int bi = 0xffff;
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
      if (cx.lane < R)   // one copy per lane
        st_tag64(p.cand + ((size_t)cx.lane * cx.G + cx.c) * p.Bmax + m, ((unsigned long long)otag << 32) |
                                                                          ((unsigned long long)(bi & 0xffff) << 16) |
                                                                          (unsigned long long)float_to_bf16_bits(best));
    }
  }
}

// ------------------------------------------------------------------ end of frame
// After the last head: sample codebook 31, publish the 32 ids (modeling_csm.py:657-666) and evaluate
// the stop rule torch.all(new_frame == 0) (:662).  CTA 0 only.
__device__ __forceinline__ void finish_phase(const StreamParams& p, const Phase& P, const Ctx& cx) {
  if (cx.c != 0) return;
  const int M = p.B;
  reduce_candidates(p, cx, CSM_NQ - 1, P.res_ph);
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
      const uint32_t otag = tg(p, cx.ph);
      uint32_t* o = p.h_bb + (size_t)m * H + col;
      const size_t rs = (size_t)p.Bmax * H;
      st_tag4_r(o, tw_pack(acc[0], otag), tw_pack(acc[1], otag), tw_pack(acc[2], otag), tw_pack(acc[3], otag), p.repl, rs);
      st_tag4_r(o + 4, tw_pack(acc[4], otag), tw_pack(acc[5], otag), tw_pack(acc[6], otag), tw_pack(acc[7], otag), p.repl, rs);
    }
  }
}

// ------------------------------------------------------------------ backbone decode attention (split-KV, GQA)
// One unit = (sequence, kv-head, 128 cached positions); the 4 (rep) query heads of the group share
// every K/V byte read.  Units write (max, sum, o[64]) partials; the last unit of a (sequence,
// kv-head) merges them and publishes the head outputs as tagged words.  q and the K/V of the position
// being processed are polled from the qkv phase's tagged output; older positions come from the cache.
// Softmax in fp32 (sdpa_attention_forward, hf integrations/sdpa_attention.py:40-104; a decode step
// attends to every cached position).
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
  const uint32_t qtag = tg(p, P.src_ph), otag = tg(p, cx.ph);
  const int Wq = (p.bb.heads + 2 * nk) * HD;        // tagged q | k | v row
  const uint32_t* qbase = p.q_bb + (size_t)cx.rep * p.Bmax * Wq;
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
      if (pj < p.pos) {
        kv4[j] = ldcg_u4(Kp + (size_t)pj * HD + dl * 8);
        vv4[j] = ldcg_u4(Vp + (size_t)pj * HD + dl * 8);
      } else {
        kv4[j] = make_uint4(0, 0, 0, 0);
        vv4[j] = make_uint4(0, 0, 0, 0);
      }
    }
    // q slice of this lane: REP heads x 8 dims, pre-scaled (tagged words from the qkv phase)
    float q[REP][8];
    {
      const uint32_t* qw = qbase + (size_t)b * Wq + (kvh * REP) * HD + dl * 8;
      uint4 qa[REP], qb[REP];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int h = 0; h < REP; ++h) {
          qa[h] = ld_tag4(qw + h * HD);
          qb[h] = ld_tag4(qw + h * HD + 4);
          ok &= tw_ok4(qa[h], qtag) & tw_ok4(qb[h], qtag);
        }
      } while (!__all_sync(0xffffffffu, ok));
#pragma unroll
      for (int h = 0; h < REP; ++h) {
        q[h][0] = tw_val(qa[h].x) * p.bb.scale; q[h][1] = tw_val(qa[h].y) * p.bb.scale;
        q[h][2] = tw_val(qa[h].z) * p.bb.scale; q[h][3] = tw_val(qa[h].w) * p.bb.scale;
        q[h][4] = tw_val(qb[h].x) * p.bb.scale; q[h][5] = tw_val(qb[h].y) * p.bb.scale;
        q[h][6] = tw_val(qb[h].z) * p.bb.scale; q[h][7] = tw_val(qb[h].w) * p.bb.scale;
      }
    }
    // the position being processed: its K/V are in flight to the cache, take them from the tagged row
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (pbase + 4 * j + grp == p.pos) {
        const uint32_t* kw = qbase + (size_t)b * Wq + p.bb.heads * HD + kvh * HD + dl * 8;
        const uint32_t* vw = kw + nk * HD;
        uint4 k0, k1, v0, v1;
        do {
          k0 = ld_tag4(kw); k1 = ld_tag4(kw + 4);
          v0 = ld_tag4(vw); v1 = ld_tag4(vw + 4);
        } while (!(tw_ok4(k0, qtag) & tw_ok4(k1, qtag) & tw_ok4(v0, qtag) & tw_ok4(v1, qtag)));
        kv4[j] = make_uint4(tw_pair(k0.x, k0.y), tw_pair(k0.z, k0.w), tw_pair(k1.x, k1.y), tw_pair(k1.z, k1.w));
        vv4[j] = make_uint4(tw_pair(v0.x, v0.y), tw_pair(v0.z, v0.w), tw_pair(v1.x, v1.y), tw_pair(v1.z, v1.w));
      }
    }
    __syncwarp();
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
        st_tag_r(p.attn_bb + (size_t)b * (p.bb.heads * HD) + (kvh * REP + h) * HD + d, tw_pack(O / L, otag), p.repl,
                 (size_t)p.Bmax * p.bb.heads * HD);
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
  cx.sprog = reinterpret_cast<volatile unsigned int*>(cx.aempty + 3);
  cx.desc = reinterpret_cast<Phase*>(csm_smem + 256);
  cx.scratch = reinterpret_cast<float*>(csm_smem + 512);    // 512 floats
  cx.tok = reinterpret_cast<int*>(csm_smem + 2560);
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
  cx.ph = p.phase_begin;
  cx.rep = blockIdx.x % p.repl;
  cx.prof = nullptr;

  if (cx.tid == 0) {
    for (int s = 0; s < CSM_MAX_SLOTS; ++s) {
      mbar_init(&cx.full[s], 1);
      mbar_init(&cx.empty[s], CSM_COMPUTE_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&cx.afull[s], 1);
      mbar_init(&cx.aempty[s], CSM_COMPUTE_WARPS);
    }
    *cx.sprog = 0u;
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
  const int bar_base = p.phases[p.phase_begin].bar_idx;   // grid-barrier events before/at the first phase

  if (cx.warp == CSM_COMPUTE_WARPS) {
    // ===================== weight stream producer =====================
    if (cx.lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      uint32_t s = 0, round = 0, prog = 0;
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
          if (p.evict_first) bulk_g2s_hint(cx.ring + (size_t)s * p.slot_bytes, src, bytes, &cx.full[s], pol);
          else bulk_g2s(cx.ring + (size_t)s * p.slot_bytes, src, bytes, &cx.full[s]);
          src += bytes;
          prog += bytes;
          *cx.sprog = prog;
          if (++s == (uint32_t)p.n_slots) { s = 0; ++round; }
        }
      }
    }
    return;
  }
  if (cx.warp == CSM_COMPUTE_WARPS + 1) {
    // ===================== activation stream producer (K=8192 phases at batch > 4) =====================
    if (cx.lane == 0) {
      uint32_t ait = 0;
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase P = p.phases[ph];
        if (P.type != PH_GEMV || P.act_mode != ACT_STREAM) continue;
        const Geom g = csm_geom(P, cx.c);
        if (g.nchunks == 0) continue;
        const bf16* actp = P.act;
        const int act_stride = P.act_stride;
        if (p.use_barrier && ph > p.phase_begin) grid_wait(p.bar_counter, (unsigned)(P.bar_idx - bar_base) * cx.G);
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
  if (cx.warp == CSM_COMPUTE_WARPS + 2) {
    // ===================== L2 prefetcher =====================
    // Issues HBM->L2 prefetches for this CTA's weight stream up to l2_ahead_bytes beyond what the ring
    // has requested, so that DRAM never idles while the ring is full and the compute warps are inside
    // a latency-bound stretch (staging, epilogue, attention).  Also: the norm weights of upcoming
    // phases (one CTA each) and the K/V blocks of this CTA's first backbone attention units.
    if (cx.lane == 0 && p.l2_ahead_bytes > 0) {
      uint32_t pf = 0;
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase P = p.phases[ph];
        if (P.type == PH_ATTN_BB) {
          const int Ttot = p.pos + 1, nk = p.bb.kv;
          const int nsplit = (Ttot + CSM_ATT_SPLIT - 1) / CSM_ATT_SPLIT;
          const int nunits = p.B * nk * nsplit;
          int done = 0;
          for (int unit = cx.c; unit < nunits && done < 4; unit += cx.G, ++done) {
            const int sp = unit % nsplit, kvh = (unit / nsplit) % nk, b = unit / (nsplit * nk);
            const size_t off = ((((size_t)P.layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap + (size_t)sp * CSM_ATT_SPLIT) * 64;
            const int npos = min(CSM_ATT_SPLIT, p.pos - sp * CSM_ATT_SPLIT);   // cached positions only
            if (npos > 0) {
              bulk_prefetch_l2(p.kc_bb + off, (uint32_t)npos * 128u);
              bulk_prefetch_l2(p.vc_bb + off, (uint32_t)npos * 128u);
            }
          }
          continue;
        }
        if (P.type != PH_GEMV) continue;
        if (P.norm_w != nullptr && (ph % cx.G) == cx.c) bulk_prefetch_l2(P.norm_w, (uint32_t)P.K * 2u);
        const Geom g = csm_geom(P, cx.c);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2;
        const uint32_t total = (uint32_t)g.rows * (uint32_t)P.K * 2u;
        for (uint32_t off = 0; off < total; off += 32768u) {
          const uint32_t n = min(32768u, total - off);
          uint32_t prog = *cx.sprog;
          while ((int)(pf - prog) > p.l2_ahead_bytes) {
            __nanosleep(500);
            prog = *cx.sprog;
          }
          if ((int)(pf + n - prog) > 0) bulk_prefetch_l2(src + off, n);   // skip what the ring has already asked for
          pf += n;
        }
      }
    }
    return;
  }

  // ===================== compute warps =====================
  for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
    unsigned long long* prof = nullptr;   // debug stamps of the first and the last CTA: [cta][phase][8]
    if (p.prof != nullptr && cx.tid == 0 && (cx.c == 0 || cx.c == cx.G - 1))
      prof = p.prof + ((size_t)(cx.c == 0 ? 0 : 1) * p.n_phases_total + ph) * 16;
    cx.prof = prof;
    cx.ph = ph;
    // descriptor of this phase is in shared memory; fetch the next one while this phase runs
    const Phase P = cx.desc[ph & 1];
    if ((P.flags & CSM_PF_BAR_IN) && p.use_barrier && ph > p.phase_begin) {
      if (cx.tid == 0) grid_wait(p.bar_counter, (unsigned)(P.bar_idx - bar_base) * cx.G);
      compute_sync();
    }
    if (prof) prof[0] = clock64();       // (barrier observed)
    uint4 nxt = make_uint4(0, 0, 0, 0);
    const bool fetch = cx.warp == CSM_COMPUTE_WARPS - 1 && cx.lane < 8 && ph + 1 < p.phase_end;
    if (fetch) nxt = __ldg(reinterpret_cast<const uint4*>(p.phases + ph + 1) + cx.lane);
    if (prof) prof[1] = clock64();       // phase body starts
    switch (P.type) {
      case PH_EMBED: embed_phase(p, cx); break;
      case PH_GEMV: gemv_phase<NB>(p, P, cx); break;
      case PH_ATTN_BB: attn_bb_phase<REP>(p, P, cx); break;
      case PH_ATTN_DEC: attn_dec_phase(p, P, cx); break;
      case PH_FINISH: finish_phase(p, P, cx); break;
    }
    if (prof) prof[2] = clock64();       // this thread's share of the body done
    if (fetch) reinterpret_cast<uint4*>(&cx.desc[(ph + 1) & 1])[cx.lane] = nxt;
    if (ph + 1 < p.phase_end) {
      compute_sync();                    // shared-memory reuse between phases; next descriptor visible
      if ((P.flags & CSM_PF_BAR_OUT) && p.use_barrier && cx.tid == 0) {
        // release: everything this CTA wrote (ordered before by the CTA barrier) becomes visible before the count
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar_counter) : "memory");
      }
    }
    if (prof) prof[3] = clock64();       // end of phase
    if (p.prof != nullptr && cx.tid == 0) {   // wall-clock end of this phase for every CTA (skew between CTAs)
      unsigned long long gt;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      p.prof[(size_t)32 * p.n_phases_total + (size_t)cx.c * p.n_phases_total + ph] = gt;
    }
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
