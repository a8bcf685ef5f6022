// The frame engine: ONE persistent kernel executes a whole audio frame -- backbone decode step,
// codebook-0 head, 32 decoder positions x 4 layers, 31 codebook heads, greedy sampling and the
// embedding gathers between them -- as a table of ~780 dependent phases.
//
// Replaces, per frame, the ~7000 ATen kernel launches of CSMModel.generate_frame
// (reference modeling_csm.py:484-589 driving hf LlamaModel.forward x32).
//
// Structure of a CTA (one per SM, 148 on B200):
//   warps 0..7  compute: stage activations, tensor-core MMA on weight chunks, fused epilogues
//   warp  8     weight stream: walks the SAME phase table and keeps a ring of shared-memory
//               slots full with this CTA's slice of every weight matrix, using bulk
//               asynchronous copies (TMA engine, SASS UBLKCP) that complete on mbarriers.
//               It never waits for the compute warps' inputs, so HBM keeps streaming the NEXT
//               phases' weights while the compute warps sit in a poll or an epilogue.
//   warp  9     activation stream for K=8192 (down_proj) phases at batch > 4, whose activations do not
//               fit in shared memory: bulk-copies [B, k-chunk] tiles after a grid barrier.
//   warp 10     L2 prefetcher: walks the phase table further ahead still and pulls this CTA's weight
//               slices (and norm weights, and the K/V blocks of its attention units) from HBM into L2
//               with cp.async.bulk.prefetch.L2, so HBM keeps streaming even when the ring is full.
//
// Phases are chained by dataflow: every vector that crosses CTAs is an array of tagged words
// (bf16 | 16-bit tag of the producing phase, csm_common.cuh) that the consumer polls while staging, so
// one hand-over costs one trip through L2 -- no fence, no atomic, no separate barrier round.
//
// Work split of a matrix W[N,K]: rows are divided evenly over the CTAs (granule 1 or 2 rows);
// csm_pack.cu stores each CTA's rows contiguously, k16-tile major, in ldmatrix order.  The WEIGHTS are
// the 16-row A operand of mma.sync.m16n8k16, the batch rows of the activations the 8-column B operand,
// so a batch of 1..8 sequences costs one MMA per 256 weights and nothing is wasted on padding.
//
// The chain is ~780 phases long, so what bounds a frame at small batch is the number of INSTRUCTIONS on
// the critical path of a phase, not bytes.  Rules this file follows:
//   * the per-phase geometry is computed on the host (GeoC in the descriptor), descriptors are prefetched
//     into shared memory one phase ahead and read in place;
//   * the hot path (staging, MMA loop, epilogue) is kept small and loop-invariants are pinned in registers
//     (the optimiser otherwise re-derives strides inside the MMA loop); rarely executed phases
//     (embedding, backbone attention, finish) are out of line;
//   * every cross-CTA read is issued as one batch of independent loads (one L2 round trip);
//   * greedy sampling needs no extra synchronisation: every CTA publishes its best (logit, id) as a tagged
//     word and the consumers reduce the 148 candidates themselves.
//
// This file is the kernel family of engines for <= 2 sequences (pure dataflow, nothing on the hot path that is not
// needed there), compiled twice: greedy and STOCH (stochastic top-k sampling) -- separate translation units, because
// the small-batch greedy kernel is bound by the instruction count and register pressure of its hot path: code it
// never executes still costs it 10-20 % when compiled in.  Engines for 3..32 sequences run csm_batch.inl (plain bf16
// hand-over, grid barriers, TMA-staged activations).
#include "csm_common.cuh"
#include "csm_sample.cuh"

#if !defined(CSM_BUILD_SMALL) || !defined(CSM_BUILD_STOCH) || !CSM_BUILD_SMALL
#error "include this file from csm_stream_small{,_stoch}.cu (the general kernel family is csm_batch.inl)"
#endif

// build-time experiment knobs (tools/gpu_variants.sh builds several libraries and times them in one GPU call)
#ifndef CSM_MMA_UNROLL
#define CSM_MMA_UNROLL 4
#endif
#ifndef CSM_KV_EARLY
#define CSM_KV_EARLY 1
#endif
#ifndef CSM_NORM_SPLIT
#define CSM_NORM_SPLIT 1
#endif
#ifndef CSM_GUARD
#define CSM_GUARD (!CSM_BUILD_SMALL)
#endif
#ifndef CSM_PROGRESS_HOOKS
#define CSM_PROGRESS_HOOKS (!CSM_BUILD_SMALL)
#endif
#define CSM_STR2(x) #x
#define CSM_STR(x) CSM_STR2(x)

extern __shared__ __align__(128) unsigned char csm_smem[];

namespace {

// ---- shared-memory header (CSM_SM_HDR_BYTES = 4096) ----
//   [0,64) full[8] | [64,128) empty[8] | [128,144) afull[2] | [144,160) aempty[2] | [160,168) sflag[2] |
//   [168,172) weight-stream progress | [256,768) 2 phase descriptors | [768,2816) 512 floats scratch | [2816,2944) tok[32] |
//   [2944,3072) rstd[32]
__device__ __forceinline__ uint64_t* sm_full() { return reinterpret_cast<uint64_t*>(csm_smem); }
__device__ __forceinline__ uint64_t* sm_empty() { return reinterpret_cast<uint64_t*>(csm_smem + 64); }
__device__ __forceinline__ uint64_t* sm_afull() { return reinterpret_cast<uint64_t*>(csm_smem + 128); }
__device__ __forceinline__ uint64_t* sm_aempty() { return reinterpret_cast<uint64_t*>(csm_smem + 144); }
__device__ __forceinline__ volatile int* sm_flag() { return reinterpret_cast<volatile int*>(csm_smem + 160); }
__device__ __forceinline__ volatile unsigned int* sm_prog() { return reinterpret_cast<volatile unsigned int*>(csm_smem + 168); }
__device__ __forceinline__ Phase* sm_desc() { return reinterpret_cast<Phase*>(csm_smem + 256); }
__device__ __forceinline__ float* sm_scratch() { return reinterpret_cast<float*>(csm_smem + 768); }
__device__ __forceinline__ int* sm_tok() { return reinterpret_cast<int*>(csm_smem + 2816); }
__device__ __forceinline__ float* sm_rstd() { return reinterpret_cast<float*>(csm_smem + 2944); }   // 32 floats
// cos_dec | sin_dec ([32][hd/2] each) | cos_bb[pos] | sin_bb[pos]
__device__ __forceinline__ bf16* sm_rope() { return reinterpret_cast<bf16*>(csm_smem + CSM_SM_HDR_BYTES); }
__device__ __forceinline__ float* sm_red(const StreamParams& p) {
  return reinterpret_cast<float*>(csm_smem + CSM_SM_HDR_BYTES + p.rope_bytes);
}
__device__ __forceinline__ unsigned char* sm_act(const StreamParams& p) {
  return csm_smem + CSM_SM_HDR_BYTES + p.rope_bytes + p.red_bytes;
}
__device__ __forceinline__ unsigned char* sm_ring(const StreamParams& p) {
  return csm_smem + CSM_SM_HDR_BYTES + p.rope_bytes + p.red_bytes + p.act_region_bytes;
}

// Per-thread state of a compute warp.  Plain scalars; only ever passed to inlined code, so it stays in
// registers (a context struct whose address escapes into an out-of-line call is demoted to local memory
// and every access becomes a load).
struct Lane {
  int tid, warp, lane, c, G;
  uint32_t slot, slot_par, aslot, aslot_par;   // ring positions of the consumer side
  int ph;                                      // phase being executed
  unsigned long long* prof;                    // debug stamps of this phase (thread 0 of the first / last CTA) or null
};

#define CSM_STAMP(L, i)                     \
  do {                                      \
    if ((L).prof) (L).prof[i] = clock64();  \
  } while (0)
#define CSM_PROGRESS(p, c, tid, slot, v)                                      \
  do {                                                                        \
    if (CSM_PROGRESS_HOOKS && (p).progress != nullptr && (tid) == 0) (p).progress[(c) * 4 + (slot)] = (v); \
  } while (0)

// ---- hang guard ----
// Every wait in this kernel is a spin on memory another CTA (or the TMA engine) will write.  A protocol bug or
// a lost CTA would otherwise wedge the GPU for good; instead a wait that lasts longer than ~2 s records who waited
// for what and raises the abort flag, which makes every wait in the grid give up and every later launch return at
// once; the host reports it as an error (csm_frames_done / csm_generate_frame).  Cost: one counter increment
// per failed poll.
enum WaitId { W_STAGE = 1, W_CAND = 2, W_ATTN_DEC = 3, W_ATTN_BB_Q = 4, W_ATTN_BB_KV = 5, W_RESID = 6, W_GRID = 7,
              W_FULL = 8, W_AFULL = 9, W_EMPTY = 10, W_AEMPTY = 11 };

__device__ __noinline__ bool spin_slow(const StreamParams& p, unsigned n, int ph, int id, unsigned a, unsigned b) {
  if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) return true;
  if (n < (1u << 22)) return false;            // >= 4M failed polls of >= 0.15-0.5 us each: seconds
  if (atomicCAS(p.abort_flag, 0, 1) == 0) {
    p.abort_flag[1] = (int)blockIdx.x; p.abort_flag[2] = ph; p.abort_flag[3] = id; p.abort_flag[4] = (int)a;
    p.abort_flag[5] = (int)b; p.abort_flag[6] = (int)threadIdx.x;
    __threadfence();
  }
  return true;
}
// Back-off between failed polls of the general (many-sequence) kernels: with 8-32 sequences a CTA re-reads 32-128 KB
// per polling round; 148 CTAs doing that back to back saturate L2 and slow down the very producers they wait for
// (observed as multi-second stalls at 24-32 sequences).  Below 16 sequences nobody sleeps.
__device__ __forceinline__ void poll_backoff(const StreamParams& p, unsigned n) {
  if (!CSM_BUILD_SMALL && p.B >= 16) __nanosleep(n < 4u ? 100u : (n < 32u ? 400u : 1500u));
}
// call once per failed poll (one add and one test on the fast path); true = give up
__device__ __forceinline__ bool spin_giveup(const StreamParams& p, unsigned& n, int ph, int id, unsigned a = 0,
                                            unsigned b = 0) {
  if (!CSM_GUARD) return false;
  if ((++n & 0xffffu) != 0) return false;
  return spin_slow(p, n, ph, id, a, b);
}

__device__ __forceinline__ void grid_wait(const StreamParams& p, const unsigned int* counter, unsigned target, int ph) {
  unsigned n = 0;
  while (ld_acquire_u32(counter) < target) {
    if (spin_giveup(p, n, ph, W_GRID, target)) break;
  }
}
__device__ __forceinline__ void mbar_wait_g(const StreamParams& p, uint64_t* bar, uint32_t parity, int ph, int id) {
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (spin_giveup(p, n, ph, id, parity)) break;
  }
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }
__device__ __forceinline__ uint32_t tg(const StreamParams& p, int ph) { return (p.tagbase + (uint32_t)ph) & 0xffffu; }
// Keep a loop-invariant value in its register: stops the optimiser from re-deriving it inside a loop.
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(int& x) { asm volatile("" : "+r"(x)); }

// ------------------------------------------------------------------ greedy sample of a finished head phase
// sample_topk at topk=1 (modeling_csm.py:179-189) with the canonical lowest-index tie-break: reduce
// the (best logit, id) candidates every CTA published in head phase `head_ph` for codebook `cb`
// (tagged 64-bit words: polling them IS the synchronisation with that phase).
// Result in tok[m]; CTA 0 also records samples / fed.  Ends with a compute_sync.
__device__ __forceinline__ void reduce_candidates(const StreamParams& p, int warp, int lane, int c, int G, int cb,
                                                  int head_ph) {
  const int M = p.B;
  const unsigned long long tag = tg(p, head_ph);
  int* tok = sm_tok();
#pragma unroll 1
  for (int m = warp; m < M; m += CSM_COMPUTE_WARPS) {
    unsigned long long w[5];
    bool ok;
    unsigned spin = 0;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int cc = lane + 32 * j;
        w[j] = 0;
        if (cc < G) {
          w[j] = ld_tag64(p.cand + (size_t)cc * p.Bmax + m);
          ok &= ((w[j] >> 32) & 0xffffull) == tag;
        }
      }
      if (!ok) poll_backoff(p, spin);
      if (!ok && spin_giveup(p, spin, head_ph, W_CAND, (unsigned)m, (unsigned)lane)) ok = true;
    } while (!__all_sync(0xffffffffu, ok));
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (lane + 32 * j < G) {
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
    if (lane == 0) {
      int fedtok = bi;
      if (p.forced) fedtok = ldcg_i32(p.fed + m * CSM_NQ + cb);
      tok[m] = fedtok;
      if (c == 0) {
        p.samples[m * CSM_NQ + cb] = bi;
        if (!p.forced) p.fed[m * CSM_NQ + cb] = bi;
      }
    }
  }
  compute_sync();
}

#if CSM_BUILD_STOCH
// ------------------------------------------------------------------ stochastic top-k sample of a finished head phase
// sample_topk(logits, topk, temperature) (modeling_csm.py:179-189) for codebook `cb`: every CTA polls the tagged
// logits the head phase published, keeps them as 16-bit sort keys in shared memory (the activation region, free
// at this point), and one warp per sequence selects the k-th largest and draws by Gumbel-max with hashed noise
// (csm_sample.cuh) -- every CTA draws the same token.  Out of line: only stochastic runs execute it.
__device__ __noinline__ void sample_tokens(const StreamParams& p, int cb, int head_ph) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, c = blockIdx.x;
  const int M = p.B, V = p.V, Vs = p.lgt_stride;
  unsigned short* keys = reinterpret_cast<unsigned short*>(sm_act(p));   // [M][Vs]
  const uint32_t tag = tg(p, head_ph);
  const int gpr = Vs >> 2, total = M * gpr;
  unsigned spin = 0;
#pragma unroll 1
  for (int i0 = tid; i0 - lane < total; i0 += 4 * CSM_COMPUTE_THREADS) {   // whole warps iterate (the poll votes)
    uint4 w[4];
    int mm[4], gg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * CSM_COMPUTE_THREADS;
      mm[j] = i / gpr;
      gg[j] = i - mm[j] * gpr;
    }
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i0 + j * CSM_COMPUTE_THREADS < total) {
          w[j] = ld_tag4(p.lgt + (size_t)mm[j] * Vs + gg[j] * 4);
          const int nv = V - gg[j] * 4;   // valid words of this group (the row is padded to a multiple of 4)
          ok &= (w[j].x >> 16) == tag && (nv < 2 || (w[j].y >> 16) == tag) && (nv < 3 || (w[j].z >> 16) == tag) &&
                (nv < 4 || (w[j].w >> 16) == tag);
        }
      }
      if (!ok) poll_backoff(p, spin);
      if (!ok && spin_giveup(p, spin, head_ph, W_CAND, (unsigned)i0, 1u)) ok = true;
    } while (!__all_sync(0xffffffffu, ok));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j * CSM_COMPUTE_THREADS < total) {
        const uint32_t k0 = bf16_sort_key(w[j].x & 0xffffu), k1 = bf16_sort_key(w[j].y & 0xffffu);
        const uint32_t k2 = bf16_sort_key(w[j].z & 0xffffu), k3 = bf16_sort_key(w[j].w & 0xffffu);
        *reinterpret_cast<uint2*>(keys + (size_t)mm[j] * Vs + gg[j] * 4) = make_uint2(k0 | (k1 << 16), k2 | (k3 << 16));
      }
    }
  }
  compute_sync();
  int* hist = reinterpret_cast<int*>(sm_red(p)) + warp * 256;
  int* tok = sm_tok();
  const int k = p.topk < V ? p.topk : V;
#pragma unroll 1
  for (int m = warp; m < M; m += CSM_COMPUTE_WARPS) {
    const int idx = warp_sample_topk(keys + (size_t)m * Vs, V, k, p.inv_temp,
                                     draw_key(p.rng_seed, p.rng_frame, cb, p.seq_base + m), lane, hist);
    if (lane == 0) {
      int fedtok = idx;
      if (p.forced) fedtok = ldcg_i32(p.fed + m * CSM_NQ + cb);
      tok[m] = fedtok;
      if (c == 0) {
        p.samples[m * CSM_NQ + cb] = idx;
        if (!p.forced) p.fed[m * CSM_NQ + cb] = idx;
      }
    }
    __syncwarp();
  }
  compute_sync();
}
#endif

// ------------------------------------------------------------------ decoder attention (<= 32 positions, hd 128)
// One warp per (sequence, query head).  Lane t owns cached position t for the scores and output dims
// 4*lane.. for P.V.  q and the K/V of the position being processed come as tagged words straight from
// the qkv phase (polled here); older positions come from the cache.  q is spread to all lanes through a
// 512-byte shared-memory row of the warp.  Returns the normalised output dims 4*lane..4*lane+3.
__device__ __forceinline__ void attn_dec_unit(const StreamParams& p, int layer, int dec_pos, int b, int head, uint32_t qtag,
                                              int lane, float* qs, float (&out)[4]) {
  constexpr int HD = 128;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int kvh = head / rep;
  const int W = (nh + 2 * nk) * HD;
  const size_t kvbase = (((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)CSM_DEC_POS * HD;
  const bf16* kp = p.kc_dec + kvbase + (size_t)lane * HD;
  const bf16* vp = p.vc_dec + kvbase + lane * 4;
  // q | k | v of this position: lane l holds dims 4l..4l+3 of each
  const uint32_t* qw = p.q_dec + (size_t)b * W + head * HD + lane * 4;
  const uint32_t* kw = p.q_dec + (size_t)b * W + nh * HD + kvh * HD + lane * 4;
  const uint32_t* vw = kw + nk * HD;
  uint4 q4, k4, v4;
  bool ok;
  unsigned spin = 0;
  do {
    q4 = ld_tag4(qw);
    k4 = ld_tag4(kw);
    v4 = ld_tag4(vw);
    ok = tw_ok4(q4, qtag) & tw_ok4(k4, qtag) & tw_ok4(v4, qtag);
    if (!ok) poll_backoff(p, spin);
    if (!ok && spin_giveup(p, spin, layer, W_ATTN_DEC, (unsigned)(b * 256 + head), (unsigned)dec_pos)) ok = true;
  } while (!__all_sync(0xffffffffu, ok));
  const float sc = p.dec.scale;
  const float4 qf = make_float4(tw_val(q4.x) * sc, tw_val(q4.y) * sc, tw_val(q4.z) * sc, tw_val(q4.w) * sc);
  *reinterpret_cast<float4*>(qs + lane * 4) = qf;
  // score of the position being processed: every lane holds 4 dims of its k
  float dcur = qf.x * tw_val(k4.x) + qf.y * tw_val(k4.y) + qf.z * tw_val(k4.z) + qf.w * tw_val(k4.w);
  dcur = warp_sum(dcur);
  __syncwarp();            // qs written by all lanes before any lane reads it
  float d = 0.f;
  if (lane < dec_pos) {
#pragma unroll 8
    for (int ci = 0; ci < HD / 8; ++ci) {
      const uint4 kv = ldcg_u4(kp + ci * 8);
      const float4 a = *reinterpret_cast<const float4*>(qs + ci * 8), c4 = *reinterpret_cast<const float4*>(qs + ci * 8 + 4);
      d += a.x * bf_lo(kv.x) + a.y * bf_hi(kv.x) + a.z * bf_lo(kv.y) + a.w * bf_hi(kv.y);
      d += c4.x * bf_lo(kv.z) + c4.y * bf_hi(kv.z) + c4.z * bf_lo(kv.w) + c4.w * bf_hi(kv.w);
    }
  }
  if (lane == dec_pos) d = dcur;
  const float s = lane <= dec_pos ? d : -INFINITY;
  const float mx = warp_max(s);
  const float pe = (lane <= dec_pos) ? __expf(s - mx) : 0.f;
  const float l = warp_sum(pe);
  const float pc = __shfl_sync(0xffffffffu, pe, dec_pos);
  float o0 = pc * tw_val(v4.x), o1 = pc * tw_val(v4.y), o2 = pc * tw_val(v4.z), o3 = pc * tw_val(v4.w);
#pragma unroll 1
  for (int t0 = 0; t0 < dec_pos; t0 += 16) {
    uint2 vv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) vv[j] = (t0 + j < dec_pos) ? ldcg_u2(vp + (size_t)(t0 + j) * HD) : make_uint2(0, 0);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float pv = __shfl_sync(0xffffffffu, pe, (t0 + j) & 31);
      if (t0 + j < dec_pos) {
        o0 += pv * bf_lo(vv[j].x); o1 += pv * bf_hi(vv[j].x);
        o2 += pv * bf_lo(vv[j].y); o3 += pv * bf_hi(vv[j].y);
      }
    }
  }
  const float inv = 1.f / l;
  out[0] = o0 * inv; out[1] = o1 * inv; out[2] = o2 * inv; out[3] = o3 * inv;
}

// Separate-phase form: units spread over the CTAs, result published as tagged words.
__device__ __forceinline__ void attn_dec_phase(const StreamParams& p, const Phase& P, const Lane& L) {
  const int nh = p.dec.heads;
  const int nunits = p.B * nh;
  const uint32_t qtag = tg(p, P.src_ph), otag = tg(p, L.ph);
  float* qs = reinterpret_cast<float*>(sm_act(p)) + L.warp * 128;   // (the previous phase's MMAs are done with it)
#pragma unroll 1
  for (int unit = L.warp * L.G + L.c; unit < nunits; unit += CSM_COMPUTE_WARPS * L.G) {
    const int b = unit / nh, head = unit - b * nh;
    float o[4];
    attn_dec_unit(p, P.layer, P.dec_pos, b, head, qtag, L.lane, qs, o);
    st_tag4(p.attn_dec + (size_t)b * (nh * p.dec.hd) + head * p.dec.hd + L.lane * 4, tw_pack(o[0], otag),
            tw_pack(o[1], otag), tw_pack(o[2], otag), tw_pack(o[3], otag));
    __syncwarp();
  }
}

// Fused form (batch <= 2): the attention of every (sequence, head) is computed by EVERY CTA while it stages
// the input of o_proj -- one phase and one hand-over less per layer.  The cached K/V rows (< 32 positions)
// are copied to shared memory by all threads at the END of the qkv phase (attn_kv_copy: the loads overlap
// the hand-over of that phase's outputs; padded rows: conflict-free 16-byte reads by position), then in the
// o_proj phase each warp takes (sequence, head) units; q | k | v of the position being processed are polled
// from the qkv phase's tagged output.  Result: bf16 rows [M][astride] in the activation region = o_proj's input.
__device__ __forceinline__ bf16* attn_kvs(const StreamParams& p) {   // [b][K|V][kvh][32][hd+8] after the o_proj input rows
  return reinterpret_cast<bf16*>(sm_act(p)) + (size_t)p.m_alloc * (p.dec.heads * p.dec.hd + 8);
}

__device__ __forceinline__ void attn_kv_copy(const StreamParams& p, int layer, int dec_pos, int tid) {
  constexpr int HD = 128, RS = HD + 8;
  const int M = p.B, nk = p.dec.kv;
  bf16* kvs = attn_kvs(p);
  // item = (x = (b, K|V, kvh), position t, 16-byte chunk); asynchronous 16-byte copies (LDGSTS): the thread
  // only issues them, the o_proj phase waits for them (cp.async.wait_all) before its first CTA barrier
  // thread -> (position t0 + tid/16, chunk tid%16); loop over x and the two halves of the 32 positions
  const int nx = M * 2 * nk, nkl = nk > 1 ? 1 : 0;   // (kv heads: 1 or 2)
  const int c16 = tid & 15, tl = tid >> 4;
#pragma unroll 1
  for (int x = 0; x < nx; ++x) {
    const int kvh = x & (nk - 1), y = x >> nkl, b = y >> 1;
    const bf16* src = ((y & 1) ? p.vc_dec : p.kc_dec) + ((((size_t)layer * p.Bmax + b) * nk + kvh) * CSM_DEC_POS) * HD + c16 * 8;
    const uint32_t dsts = smem_u32(kvs + ((size_t)x * CSM_DEC_POS) * RS + c16 * 8);
#pragma unroll
    for (int t0 = 0; t0 < CSM_DEC_POS; t0 += 16) {
      const int t = t0 + tl;
      if (t < dec_pos)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts + (uint32_t)t * (RS * 2)), "l"(src + (size_t)t * HD) : "memory");
    }
  }
}

__device__ __forceinline__ void stage_attn_dec(const StreamParams& p, const Phase& P, const Lane& L, int astride) {
  constexpr int HD = 128, RS = HD + 8;
  const int M = p.B, dec_pos = P.dec_pos;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int W = (nh + 2 * nk) * HD;
  const uint32_t qtag = tg(p, P.src_ph);
  bf16* dst = reinterpret_cast<bf16*>(sm_act(p));
  const bf16* kvs = attn_kvs(p);
  float* qs = reinterpret_cast<float*>(attn_kvs(p) + (size_t)p.Bmax * 2 * nk * CSM_DEC_POS * RS) + L.warp * HD;
  const uint32_t* qbase = p.q_dec;
  const float sc = p.dec.scale;
  // first phase of a launch (stepped / bisecting runs): the qkv phase ran in another launch, copy now
  if (!CSM_KV_EARLY || L.ph == p.phase_begin) attn_kv_copy(p, P.layer, dec_pos, L.tid);
  // the K/V copies were issued at the end of the qkv phase: wait for this thread's, then one CTA barrier
  asm volatile("cp.async.wait_all;" ::: "memory");
  compute_sync();
#pragma unroll 1
  for (int unit = L.warp; unit < M * nh; unit += CSM_COMPUTE_WARPS) {
    const int b = unit / nh, head = unit - b * nh, kvh = head / rep;
    const uint32_t* qw = qbase + (size_t)b * W + head * HD + L.lane * 4;
    const uint32_t* kw = qbase + (size_t)b * W + nh * HD + kvh * HD + L.lane * 4;
    const uint32_t* vw = kw + nk * HD;
    uint4 q4, k4, v4;
    bool ok;
    unsigned spin = 0;
    do {
      q4 = ld_tag4(qw);
      k4 = ld_tag4(kw);
      v4 = ld_tag4(vw);
      ok = tw_ok4(q4, qtag) & tw_ok4(k4, qtag) & tw_ok4(v4, qtag);
      if (!ok && spin_giveup(p, spin, L.ph, W_ATTN_DEC, (unsigned)unit, (unsigned)dec_pos)) ok = true;
    } while (!__all_sync(0xffffffffu, ok));
    const float4 qf = make_float4(tw_val(q4.x) * sc, tw_val(q4.y) * sc, tw_val(q4.z) * sc, tw_val(q4.w) * sc);
    *reinterpret_cast<float4*>(qs + L.lane * 4) = qf;
    float dcur = qf.x * tw_val(k4.x) + qf.y * tw_val(k4.y) + qf.z * tw_val(k4.z) + qf.w * tw_val(k4.w);
    dcur = warp_sum(dcur);
    __syncwarp();
    const bf16* kb = kvs + (((size_t)(b * 2 + 0) * nk + kvh) * CSM_DEC_POS + L.lane) * RS;
    float d = 0.f;
    if (L.lane < dec_pos) {
#pragma unroll 4
      for (int ci = 0; ci < HD / 8; ++ci) {
        const uint4 kv = *reinterpret_cast<const uint4*>(kb + ci * 8);
        const float4 a = *reinterpret_cast<const float4*>(qs + ci * 8), c4 = *reinterpret_cast<const float4*>(qs + ci * 8 + 4);
        d += a.x * bf_lo(kv.x) + a.y * bf_hi(kv.x) + a.z * bf_lo(kv.y) + a.w * bf_hi(kv.y);
        d += c4.x * bf_lo(kv.z) + c4.y * bf_hi(kv.z) + c4.z * bf_lo(kv.w) + c4.w * bf_hi(kv.w);
      }
    }
    if (L.lane == dec_pos) d = dcur;
    const float sv = L.lane <= dec_pos ? d : -INFINITY;
    const float mx = warp_max(sv);
    const float pe = (L.lane <= dec_pos) ? __expf(sv - mx) : 0.f;
    const float l = warp_sum(pe);
    const float pc = __shfl_sync(0xffffffffu, pe, dec_pos);
    float o0 = pc * tw_val(v4.x), o1 = pc * tw_val(v4.y), o2 = pc * tw_val(v4.z), o3 = pc * tw_val(v4.w);
    const bf16* vb = kvs + (((size_t)(b * 2 + 1) * nk + kvh) * CSM_DEC_POS) * RS + L.lane * 4;
#pragma unroll 4
    for (int t = 0; t < dec_pos; ++t) {
      const float pv = __shfl_sync(0xffffffffu, pe, t);
      const uint2 vv = *reinterpret_cast<const uint2*>(vb + (size_t)t * RS);
      o0 += pv * bf_lo(vv.x); o1 += pv * bf_hi(vv.x);
      o2 += pv * bf_lo(vv.y); o3 += pv * bf_hi(vv.y);
    }
    const float inv = 1.f / l;
    *reinterpret_cast<uint2*>(dst + (size_t)b * astride + head * HD + L.lane * 4) =
        make_uint2(pack_bf16(o0 * inv, o1 * inv), pack_bf16(o2 * inv, o3 * inv));
    __syncwarp();
  }
}

// ------------------------------------------------------------------ activation staging
// Rows of the phase input -> shared memory [M][K+8] bf16.  The input is an array of tagged words written
// by the CTAs of phase P.src_ph.  Work item = 4 consecutive words (one 16-byte load); items are dealt
// round-robin to the 256 compute threads, up to U loads in flight per thread, repeated until every word
// carries the producer's tag -- the poll IS the load, one trip through L2 after the last producer's store.
// RMSNorm exactly as LlamaRMSNorm.forward (hf modeling_llama.py:62-67): fp32 x*rsqrt(mean(x^2)+eps) ->
// bf16 -> *w -> bf16.  Pass 1 stores the raw rows and per-warp partial sums of squares; after one CTA
// barrier pass 2 scales the thread's own elements in place (fixed summation order: deterministic).
// K/4 is a power of two (checked at create time): item -> (row, group) is a shift and a mask.
template <int U>
__device__ __forceinline__ void stage_poll(const StreamParams& p, const Phase& P, const Lane& L, const uint32_t* base,
                                           bf16* dst, int astride, int total, int gsh, uint32_t tag, bool norm) {
  const int gmask = (1 << gsh) - 1, ppr = 1 << (gsh - 5);
  const int act_stride = P.act_stride;
  float* scratch = sm_scratch();
  bool first = true;
#pragma unroll 1
  for (int i0 = L.tid; i0 < total; i0 += U * CSM_COMPUTE_THREADS) {
    uint4 w[U];
    bool ok;
    int iters = 0;
    unsigned spin = 0;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (U == 1 || i < total) {
          w[j] = ld_tag4(base + (size_t)(i >> gsh) * act_stride + (i & gmask) * 4);
          ok &= tw_ok4(w[j], tag);
        }
      }
      ++iters;
      if (!ok) poll_backoff(p, spin);
      if (!ok && spin_giveup(p, spin, L.ph, W_STAGE, (unsigned)i0, w[0].x)) ok = true;
    } while (!__all_sync(0xffffffffu, ok));
    if (first && L.prof) { L.prof[8] = clock64(); L.prof[9] = (unsigned long long)iters; }
    first = false;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int i = i0 + j * CSM_COMPUTE_THREADS;
      if (U == 1 || i < total) {   // warp-uniform (K/4 % 32 == 0)
        const int m = i >> gsh, g = i & gmask;
        *reinterpret_cast<uint2*>(dst + (size_t)m * astride + g * 4) =
            make_uint2(tw_pair(w[j].x, w[j].y), tw_pair(w[j].z, w[j].w));
        if (norm) {
          const float a = tw_val(w[j].x), b = tw_val(w[j].y), c = tw_val(w[j].z), d = tw_val(w[j].w);
          const float ss = warp_sum(a * a + b * b + c * c + d * d);
          if (L.lane == 0) scratch[m * ppr + (g >> 5)] = ss;
        }
      }
    }
  }
}

// SMALL (engine built for <= 2 sequences): fused attention staging, 1-2 row RMSNorm, at most 4 loads in flight;
// otherwise: many-row RMSNorm, 4 or 8 loads in flight, no fused attention.  Two kernels, each with half the code.
template <bool SMALL>
__device__ __forceinline__ void stage_act(const StreamParams& p, const Phase& P, const Lane& L, int astride) {
  const int K = P.K, M = p.B;
  bf16* dst = reinterpret_cast<bf16*>(sm_act(p));
  const int gsh = P.gsh;                         // log2(K/4)
  const bool norm = P.act_mode == ACT_NORM || P.act_mode == ACT_GATHER;
  const int gmask = (1 << gsh) - 1;
  const int total = M << gsh;
  // norm weights of this thread's (at most two) column groups, requested before the poll starts
  uint2 nw0 = make_uint2(0, 0), nw1 = make_uint2(0, 0);
  if (norm) {
    nw0 = __ldg(reinterpret_cast<const uint2*>(P.norm_w) + (L.tid & gmask));
    nw1 = __ldg(reinterpret_cast<const uint2*>(P.norm_w) + ((L.tid + CSM_COMPUTE_THREADS) & gmask));
  }
  if (P.act_mode == ACT_GATHER) {
    // decoder input of positions 1..31: projection(_embed_audio(codebook, token)) (modeling_csm.py:247-259,
    // 564-565) = row token + codebook*V of the pre-projected table (plain bf16, read-only).  The token is the
    // greedy sample of the previous head phase.  One CTA per sequence also starts the residual stream with it.
#if CSM_BUILD_STOCH
    sample_tokens(p, P.cb, P.res_ph);
#else
    reduce_candidates(p, L.warp, L.lane, L.c, L.G, P.cb, P.res_ph);
#endif
    const int* tok = sm_tok();
    const bf16* tab = P.act + (size_t)(P.cb * p.V) * K;
    const int ppr = 1 << (gsh - 5);
    float* scratch = sm_scratch();
    uint32_t* hres = reinterpret_cast<uint32_t*>(P.norm_out);
    const uint32_t otag = tg(p, L.ph);
#pragma unroll 1
    for (int i0 = L.tid; i0 < total; i0 += 4 * CSM_COMPUTE_THREADS) {
      uint2 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (i < total) v[j] = __ldg(reinterpret_cast<const uint2*>(tab + (size_t)tok[i >> gsh] * K) + (i & gmask));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * CSM_COMPUTE_THREADS;
        if (i < total) {   // warp-uniform
          const int m = i >> gsh, g = i & gmask;
          *reinterpret_cast<uint2*>(dst + (size_t)m * astride + g * 4) = v[j];
          const float a = bf_lo(v[j].x), b = bf_hi(v[j].x), c = bf_lo(v[j].y), d = bf_hi(v[j].y);
          const float ss = warp_sum(a * a + b * b + c * c + d * d);
          if (L.lane == 0) scratch[m * ppr + (g >> 5)] = ss;
          if ((m % L.G) == L.c)
            st_tag4(hres + (size_t)m * K + g * 4, (otag << 16) | (v[j].x & 0xffffu), (otag << 16) | (v[j].x >> 16),
                    (otag << 16) | (v[j].y & 0xffffu), (otag << 16) | (v[j].y >> 16));
        }
      }
    }
  } else
  if (SMALL && P.act_mode == ACT_ATTN) {
    stage_attn_dec(p, P, L, astride);
    return;
  } else {
  const uint32_t tag = tg(p, P.src_ph);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(P.act);
  if (total <= CSM_COMPUTE_THREADS) {
    if (L.tid < total) stage_poll<1>(p, P, L, base, dst, astride, total, gsh, tag, norm);   // whole warps (total % 32 == 0)
  } else if (SMALL || total < 16 * CSM_COMPUTE_THREADS) {
    stage_poll<4>(p, P, L, base, dst, astride, total, gsh, tag, norm);   // (8 in flight measured slower for few rows)
  } else {
    stage_poll<8>(p, P, L, base, dst, astride, total, gsh, tag, norm);
  }
  }
  if (!norm) return;
  compute_sync();
  const float eps = P.stack ? p.dec.eps : p.bb.eps;
  const int ppr = 1 << (gsh - 5);
  const float fK = (float)K;
  float* scratch = sm_scratch();
  const bool keep_norm = P.norm_out != nullptr && P.act_mode == ACT_NORM;   // copy of the normalised rows (last_hidden_state)
  if (SMALL || M <= 2) {
    // one or two rows: every thread derives its row's rstd itself (no further barrier)
    const bool wide = (1 << gsh) > CSM_COMPUTE_THREADS;   // two column groups per thread (K = 2048)
    int jj = 0;
#pragma unroll 2
    for (int i = L.tid; i < total; i += CSM_COMPUTE_THREADS, ++jj) {
      const int m = i >> gsh, g = i & gmask;
      // row sum: lane l reads partial l mod ppr, butterfly over the ppr-lane groups (fixed order: deterministic)
      float ss = scratch[m * ppr + (L.lane & (ppr - 1))];
      for (int o = ppr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss / fK + eps);   // mean = sum / K exactly as torch
      const uint2 nw = (wide && (jj & 1)) ? nw1 : nw0;
      uint2* px = reinterpret_cast<uint2*>(dst + (size_t)m * astride + g * 4);
      const uint2 x = *px;
      const float y0 = bfround(bf_lo(x.x) * rstd), y1 = bfround(bf_hi(x.x) * rstd);
      const float y2 = bfround(bf_lo(x.y) * rstd), y3 = bfround(bf_hi(x.y) * rstd);
      const uint2 o = make_uint2(pack_bf16(bf_lo(nw.x) * y0, bf_hi(nw.x) * y1), pack_bf16(bf_lo(nw.y) * y2, bf_hi(nw.y) * y3));
      *px = o;
      if (keep_norm && (m % L.G) == L.c) *reinterpret_cast<uint2*>(P.norm_out + (size_t)m * K + g * 4) = o;
    }
    return;
  }
  // many rows: rstd of every row once (same butterfly order as above, so a row's result does not depend on
  // the batch it is in), then a wide scaling pass over 8-element groups
  float* rstd_s = sm_rstd();
  for (int m = L.warp; m < M; m += CSM_COMPUTE_WARPS) {
    float ss = scratch[m * ppr + (L.lane & (ppr - 1))];
    for (int o = ppr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (L.lane == 0) rstd_s[m] = rsqrtf(ss / fK + eps);
  }
  compute_sync();
  {
    const int sh8 = gsh - 1, mask8 = (1 << sh8) - 1;   // 8-element groups per row
    const int total8 = M << sh8;
    const uint4 w8a = __ldg(reinterpret_cast<const uint4*>(P.norm_w) + (L.tid & mask8));
    const uint4 w8b = __ldg(reinterpret_cast<const uint4*>(P.norm_w) + ((L.tid + CSM_COMPUTE_THREADS) & mask8));
    const bool wide8 = (1 << sh8) > CSM_COMPUTE_THREADS;
    const bool keep_any = keep_norm;
    int jj = 0;
#pragma unroll 4
    for (int i = L.tid; i < total8; i += CSM_COMPUTE_THREADS, ++jj) {
      const int m = i >> sh8, g = i & mask8;
      const float rstd = rstd_s[m];
      const uint4 nw = (wide8 && (jj & 1)) ? w8b : w8a;
      uint4* px = reinterpret_cast<uint4*>(dst + (size_t)m * astride + g * 8);
      const uint4 x = *px;
      uint4 o;
      o.x = pack_bf16(bf_lo(nw.x) * bfround(bf_lo(x.x) * rstd), bf_hi(nw.x) * bfround(bf_hi(x.x) * rstd));
      o.y = pack_bf16(bf_lo(nw.y) * bfround(bf_lo(x.y) * rstd), bf_hi(nw.y) * bfround(bf_hi(x.y) * rstd));
      o.z = pack_bf16(bf_lo(nw.z) * bfround(bf_lo(x.z) * rstd), bf_hi(nw.z) * bfround(bf_hi(x.z) * rstd));
      o.w = pack_bf16(bf_lo(nw.w) * bfround(bf_lo(x.w) * rstd), bf_hi(nw.w) * bfround(bf_hi(x.w) * rstd));
      *px = o;
      if (keep_any && (m % L.G) == L.c) *reinterpret_cast<uint4*>(P.norm_out + (size_t)m * K + g * 8) = o;
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

// All chunks of one phase for this warp; partial sums -> red[kg][m][rows_pad].
// The warp owns m-tiles mt0 (and mt1 when the CTA has more m-tiles than m-tile groups) and every ks-th
// k16-tile; two k-tiles per iteration.  With a single m-tile the two k-tiles of an iteration feed the two
// accumulator sets (two independent MMA chains); otherwise accumulator set j belongs to m-tile j.
template <int NB, bool SMALL>
__device__ __forceinline__ void gemv_core(const StreamParams& p, const Phase& P, const GeoC& gc, Lane& L, bool stream_,
                                          int astride) {
  const bool stream = !SMALL && stream_;   // (engines for <= 4 sequences never stream the down_proj input)
  const int M = p.B;
  const int rows = gc.rows, mtiles = gc.mtiles, rows_pad = gc.rows_pad, ksl = gc.ksl;
  int tpc = gc.tpc, nchunks = gc.nch, ntiles = P.K >> 4;
  float acc[2][NB][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][nb][q] = 0.f;
  const int ns = 8 >> ksl;
  const int ng = L.warp & (ns - 1), kg = L.warp >> (3 - ksl);
  const int mt0 = ng, mt1 = (ng + ns < mtiles) ? ng + ns : -1;
  const bool active = mt0 < mtiles;
  const bool single = mt1 < 0;
  // per-lane ldmatrix row addresses: lane = 8*mat + r
  const int mat = L.lane >> 3, r8 = L.lane & 7;
  uint32_t tile_bytes = (uint32_t)rows * 32u;
  int ra0 = 16 * mt0 + (mat & 1) * 8 + r8, ra1 = 16 * (single ? mt0 : mt1) + (mat & 1) * 8 + r8;
  if (ra0 >= rows) ra0 = 0;
  if (ra1 >= rows) ra1 = 0;
  const uint32_t ring0 = smem_u32(sm_ring(p)), act0 = smem_u32(sm_act(p));
  // (+ kg k16-tiles: this warp's first tile of every chunk)
  uint32_t offA0 = ring0 + (uint32_t)((mat >> 1) * rows + ra0) * 16u + (uint32_t)kg * tile_bytes;
  uint32_t offA1 = ring0 + (uint32_t)((mat >> 1) * rows + ra1) * 16u + (uint32_t)kg * tile_bytes;
  uint32_t offB[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int n = nb * 8 + r8;   // batch rows past M re-read row M-1: their accumulator columns are never stored
    offB[nb] = act0 + (uint32_t)((n < M ? n : M - 1) * astride + (((mat >> 1) << ksl) + kg) * 16 + (mat & 1) * 8) * 2u;
    pin(offB[nb]);
  }
  uint32_t wsec = tile_bytes << ksl, wstep = tile_bytes << (ksl + 1), astep = 64u << ksl;
  pin(offA0); pin(offA1); pin(wsec); pin(wstep); pin(astep);
  pin(tpc); pin(nchunks); pin(ntiles);
  uint64_t* full = sm_full();
  uint64_t* empty = sm_empty();

#pragma unroll 1
  for (int ch = 0; ch < nchunks; ++ch) {
    const int T0 = ch * tpc;
    const int tiles = min(tpc, ntiles - T0);
    const uint32_t s = L.slot;
    mbar_wait_g(p, &full[s], L.slot_par, L.ph, W_FULL);
    if (ch == 0) CSM_STAMP(L, 7);   // first weight chunk of the phase is in shared memory
    uint32_t aadd;
    uint32_t as = 0;
    if (stream) {
      as = L.aslot;
      mbar_wait_g(p, &sm_afull()[as], L.aslot_par, L.ph, W_AFULL);
      aadd = as * (uint32_t)(p.act_region_bytes / 2);
    } else {
      aadd = (uint32_t)T0 * 32u;
    }
    if (active) {
      // every chunk holds a multiple of 2*ks k16-tiles (host: plan_smem), so this warp's tiles come in pairs
      // (tl, tl+ks) and the loop body has no tail predicate; T0 is a multiple of ks, so the first tile is kg
      uint32_t wa0 = offA0 + s * (uint32_t)p.slot_bytes;
      uint32_t wa1 = offA1 + s * (uint32_t)p.slot_bytes;
      uint32_t aoff = aadd;
      const int npairs = tiles >> (ksl + 1);
      if (single) {
_Pragma(CSM_STR(unroll CSM_MMA_UNROLL))
        for (int it = 0; it < npairs; ++it) {
          uint32_t aA[4], aC[4], b[NB][4];
          ldsm_x4(aA, wa0);
          ldsm_x4(aC, wa0 + wsec);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) ldsm_x4(b[nb], offB[nb] + aoff);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            mma16816(acc[0][nb], aA, b[nb][0], b[nb][1]);
            mma16816(acc[1][nb], aC, b[nb][2], b[nb][3]);
          }
          wa0 += wstep;
          aoff += astep;
        }
      } else {
#pragma unroll 2
        for (int it = 0; it < npairs; ++it) {
          uint32_t aA[4], aB[4], aC[4], aD[4], b[NB][4];
          ldsm_x4(aA, wa0);
          ldsm_x4(aB, wa1);
          ldsm_x4(aC, wa0 + wsec);
          ldsm_x4(aD, wa1 + wsec);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) ldsm_x4(b[nb], offB[nb] + aoff);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            mma16816(acc[0][nb], aA, b[nb][0], b[nb][1]);
            mma16816(acc[1][nb], aB, b[nb][0], b[nb][1]);
            mma16816(acc[0][nb], aC, b[nb][2], b[nb][3]);
            mma16816(acc[1][nb], aD, b[nb][2], b[nb][3]);
          }
          wa0 += wstep;
          wa1 += wstep;
          aoff += astep;
        }
      }
    }
    __syncwarp();
    if (L.lane == 0) {
      mbar_arrive(&empty[s]);
      if (stream) mbar_arrive(&sm_aempty()[as]);
    }
    if (++L.slot == (uint32_t)p.n_slots) { L.slot = 0; L.slot_par ^= 1u; }
    if (stream) { L.aslot ^= 1u; if (L.aslot == 0) L.aslot_par ^= 1u; }
  }
  if (!active) return;
  if (single) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[0][nb][q] += acc[1][nb][q];
  }
  // D fragment: c0,c1 = (weight row g, batch 2t, 2t+1), c2,c3 = (row g+8, same batch columns)
  const int gq = L.lane >> 2, tq = L.lane & 3;
  float* red = sm_red(p);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int mt = j == 0 ? mt0 : mt1;
    if (j == 1 && single) break;
    const int row = 16 * mt + gq;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int n0 = nb * 8 + 2 * tq, n1 = n0 + 1;
      float* r0 = red + ((size_t)kg * p.m_alloc + n0) * rows_pad + row;
      float* r1 = red + ((size_t)kg * p.m_alloc + n1) * rows_pad + row;
      if (n0 < M) { r0[0] = acc[j][nb][0]; r0[8] = acc[j][nb][2]; }
      if (n1 < M) { r1[0] = acc[j][nb][1]; r1[8] = acc[j][nb][3]; }
    }
  }
}

// ------------------------------------------------------------------ GEMV / skinny-GEMM phase
__device__ __forceinline__ float resid_poll(const StreamParams& p, const uint32_t* q, uint32_t tag, int ph) {
  uint32_t w = ld_tag(q);
  unsigned n = 0;
  while ((w >> 16) != tag) {
    if (spin_giveup(p, n, ph, W_RESID, w, tag)) break;
    w = ld_tag(q);
  }
  return tw_val(w);
}

// Returns true when the next phase's descriptor has been published inside the phase (after its first CTA
// barrier, made visible by the second), so that no barrier is needed at the end of the phase.
template <int NB, bool SMALL>
__device__ __forceinline__ bool gemv_phase(const StreamParams& p, const Phase& P, Lane& L, const uint4& nxt, bool fetch) {
  const int M = p.B, K = P.K;
  const bool stream = !SMALL && (P.act_mode == ACT_STREAM);
  const bool hi = L.c < P.r;
  const GeoC& gc = P.geo[hi ? 0 : 1];
  const int row0 = (L.c * P.q + (hi ? L.c : P.r)) * P.gran;
  const int rows = gc.rows;
  int astride;
  if (!stream) {
    astride = K + 8;
    stage_act<SMALL>(p, P, L, astride);
    compute_sync();
    CSM_STAMP(L, 4);   // activations staged
    CSM_PROGRESS(p, L.c, L.tid, 1, 1);
    // every warp is inside this phase now: the other descriptor slot is free for the next phase
    if (fetch) reinterpret_cast<uint4*>(&sm_desc()[(L.ph + 1) & 1])[L.lane] = nxt;
  } else {
    astride = gc.tpc * 16 + 8;
  }
  // epilogue mapping: thread -> (batch row m, granule u), granules padded to a power of two (host: GeoC::ush)
  const int upc = gc.upc, ush = gc.ush;
  const int u = L.tid & ((1 << ush) - 1);
  const int mstep = ush >= 8 ? 1 : CSM_COMPUTE_THREADS >> ush;
  const int m_first = ush >= 8 ? 0 : L.tid >> ush;
  const int epi = P.epi;
  uint32_t* outw = reinterpret_cast<uint32_t*>(P.out);
  const int out_stride = P.out_stride;
  // residual word of the first element this thread will update: requested now, used after the MMAs
  // (own element of the previous residual phase, or the stream's first value written by another CTA)
  uint32_t resid0 = 0;
  if (epi == EPI_RESID && u < upc && m_first < M) resid0 = ld_tag(outw + (size_t)m_first * out_stride + row0 + u);
  if (rows > 0) gemv_core<NB, SMALL>(p, P, gc, L, stream, astride);
  CSM_STAMP(L, 5);     // this warp's MMAs done
  compute_sync();
  CSM_STAMP(L, 6);     // all warps' MMAs done
  CSM_PROGRESS(p, L.c, L.tid, 1, 2);

  // ---- fused epilogues
  if (u < upc) {
    const uint32_t otag = tg(p, L.ph), rtag = tg(p, P.res_ph);
    const int gran = P.gran, ks = 1 << gc.ksl, rows_pad = gc.rows_pad;
    float* red = sm_red(p);
    const int n = u * gran;
    const int gn = row0 + n;   // packed row index
#pragma unroll 1
    for (int m = m_first; m < M; m += mstep) {
      float v0 = 0.f, v1 = 0.f;
      {
        // split-K partials: all loads issued before the adds (ks <= 8)
        const float* r = red + (size_t)m * rows_pad + n;
        const int kstride = p.m_alloc * rows_pad;
        float a0[8], a1[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          a0[kk] = 0.f;
          a1[kk] = 0.f;
          if (kk < ks) {
            a0[kk] = r[kk * kstride];
            if (gran == 2) a1[kk] = r[kk * kstride + 1];
          }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { v0 += a0[kk]; v1 += a1[kk]; }
      }
      v0 = bfround(v0);   // nn.Linear output is bf16
      v1 = bfround(v1);
      if (epi == EPI_RESID) {   // hf modeling_llama.py:325,331: residual + f(x), both bf16
        uint32_t* o = outw + (size_t)m * out_stride + gn;
        float r;
        if (m == m_first && (resid0 >> 16) == rtag) r = tw_val(resid0);
        else r = resid_poll(p, o, rtag, L.ph);
        st_tag(o, tw_pack(r + v0, otag));
      } else if (epi == EPI_SWIGLU) {  // hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16 ; rows (2j,2j+1)=(gate_j,up_j)
        const float sl = bfround(v0 / (1.f + expf(-v0)));
        if (P.flags & CSM_PF_OUT_PLAIN) P.out[(size_t)m * out_stride + (gn >> 1)] = __float2bfloat16_rn(sl * v1);
        else st_tag(outw + (size_t)m * out_stride + (gn >> 1), tw_pack(sl * v1, otag));
      } else if (epi == EPI_QKV) {     // rows (2j,2j+1) = RoPE pair (i, i+hd/2) of q/k, or two adjacent v features
        const StackDims& sd = P.stack ? p.dec : p.bb;
        const int half = sd.hd >> 1, hl = sd.hdl - 1;
        const int pidx = gn >> 1;
        const int nq = sd.heads << hl, nk = sd.kv << hl;
        const int pos = P.stack ? P.dec_pos : p.pos;
        const int cap = P.stack ? CSM_DEC_POS : p.Tcap;
        uint32_t* qrow = outw + (size_t)m * out_stride;   // tagged q | k | v of this position
        if (pidx < nq + nk) {
          const bool isq = pidx < nq;
          const int pp = isq ? pidx : pidx - nq;
          const int head = pp >> hl, i = pp & (half - 1);
          // rope tables staged in shared memory at kernel start: decoder [32][half] cos|sin, backbone row `pos`
          const bf16* rope = sm_rope();
          const bf16* ct = P.stack ? rope + pos * half + i : rope + 2 * CSM_DEC_POS * (p.dec.hd >> 1) + i;
          const bf16* st = P.stack ? ct + CSM_DEC_POS * half : ct + half;
          const float cs = __bfloat162float(*ct), sn = __bfloat162float(*st);
          // apply_rotary_pos_emb (hf modeling_llama.py:146-168): every product and the sum round to bf16
          const float o1 = bfround(bfround(v0 * cs) + bfround(-v1 * sn));
          const float o2 = bfround(bfround(v1 * cs) + bfround(v0 * sn));
          uint32_t* qd = qrow + (isq ? 0 : sd.heads * sd.hd) + head * sd.hd + i;
          st_tag(qd, tw_pack(o1, otag));
          st_tag(qd + half, tw_pack(o2, otag));
          if (!isq) {   // DynamicCache.update (hf cache_utils.py:102-121) as an in-place write at `pos`
            bf16* kc = P.stack ? p.kc_dec : p.kc_bb;
            bf16* dstp = kc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + i;
            dstp[0] = __float2bfloat16_rn(o1);
            dstp[half] = __float2bfloat16_rn(o2);
          }
        } else {
          const int f = (pidx - nq - nk) * 2;
          const int head = f >> sd.hdl, d = f & (sd.hd - 1);
          st_tag2(qrow + (sd.heads + sd.kv) * sd.hd + f, tw_pack(v0, otag), tw_pack(v1, otag));
          bf16* vc = P.stack ? p.vc_dec : p.vc_bb;
          bf16* dstp = vc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + d;
          *reinterpret_cast<uint32_t*>(dstp) = pack_bf16(v0, v1);
        }
      } else if (epi == EPI_STORE) {
        st_tag(outw + (size_t)m * out_stride + gn, tw_pack(v0, otag));
      } else {   // EPI_HEAD
        if (P.out) P.out[(size_t)m * out_stride + gn] = __float2bfloat16_rn(v0);
        if (CSM_BUILD_STOCH) st_tag(p.lgt + (size_t)m * p.lgt_stride + gn, tw_pack(v0, otag));   // for sample_tokens
        red[(size_t)m * rows_pad + n] = v0;   // kk = 0 plane, own element only
      }
    }
  }
  if (epi == EPI_HEAD) {
    // publish this CTA's best (logit, id) per sequence as one tagged 64-bit word; consumers poll and reduce
    compute_sync();
    const uint32_t otag = tg(p, L.ph);
    const int rows_pad = gc.rows_pad;
    const float* red = sm_red(p);
#pragma unroll 1
    for (int m = L.warp; m < M; m += CSM_COMPUTE_WARPS) {
      float best = -INFINITY;
      int bi = 0xffff;
#pragma unroll 1
      for (int n = L.lane; n < rows; n += 32) {
        float v = red[(size_t)m * rows_pad + n];
        if (better(v, row0 + n, best, bi)) { best = v; bi = row0 + n; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
      }
      if (L.lane == 0)
        st_tag64(p.cand + (size_t)L.c * p.Bmax + m, ((unsigned long long)otag << 32) |
                                                        ((unsigned long long)(bi & 0xffff) << 16) |
                                                        (unsigned long long)float_to_bf16_bits(best));
    }
  }
  if (SMALL && CSM_KV_EARLY && (P.flags & CSM_PF_KV_COPY)) attn_kv_copy(p, P.layer, P.dec_pos, L.tid);
  return !stream;
}

// ------------------------------------------------------------------ end of frame
// After the last head: sample codebook 31, publish the 32 ids (modeling_csm.py:657-666) and evaluate
// the stop rule torch.all(new_frame == 0) (:662).  CTA 0 only.  Out of line: once per frame.
__device__ __noinline__ void finish_phase(const StreamParams& p, int head_ph) {
  const int tid = threadIdx.x, c = blockIdx.x;
  if (c != 0) return;
  const int M = p.B;
  volatile int* sflag = sm_flag();
  compute_sync();   // (the previous phase ended without a CTA barrier)
#if CSM_BUILD_STOCH
  sample_tokens(p, CSM_NQ - 1, head_ph);
#else
  reduce_candidates(p, tid >> 5, tid & 31, c, gridDim.x, CSM_NQ - 1, head_ph);
#endif
  __threadfence_block();
  if (tid == 0) sflag[1] = 0;
  compute_sync();
  int nz = 0;
  for (int e = tid; e < M * CSM_NQ; e += CSM_COMPUTE_THREADS) {
    const int tok = p.samples[e];   // written by this CTA (this phase or earlier ones of this launch)
    nz |= (tok != 0);
    if (p.out_frames) {
      int m = e / CSM_NQ, q = e % CSM_NQ;
      p.out_frames[(size_t)m * p.out_stride + p.out_off + q] = (long long)tok;
    }
  }
  if (nz) sflag[1] = 1;
  compute_sync();
  if (tid == 0) {
    if (p.stop_on_zeros && !sflag[1]) *p.stop_flag = 1;   // all-zero frame: not kept, generation ends
    else if (p.n_frames) *p.n_frames += 1;
  }
}

// ------------------------------------------------------------------ 33-way masked embedding gather-sum
// _embed_tokens + mask multiply + sum (modeling_csm.py:261-282,327-334): fp32 accumulate in slot order,
// one bf16 rounding.  Unit = (sequence, 256-column chunk), one warp each, spread over the CTAs.
// Out of line: once per frame.
__device__ __noinline__ void embed_phase(const StreamParams& p, int ph) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.x, G = gridDim.x;
  const int H = p.bb.H;
  const int nchunk = (H + 255) / 256;
  const int nunits = p.B * nchunk;
  for (int unit = warp * G + c; unit < nunits; unit += CSM_COMPUTE_WARPS * G) {
    const int m = unit / nchunk, col = (unit - m * nchunk) * 256 + lane * 8;
    // lane l holds (id, mask) of slot l; slot 32 (text) is held by every lane
    long long my_tok, txt_tok;
    int my_mk, txt_mk;
    if (p.ids) {
      my_tok = p.ids[m * (CSM_NQ + 1) + lane];
      txt_tok = p.ids[m * (CSM_NQ + 1) + CSM_NQ];
    } else {
      my_tok = (long long)ldcg_i32(p.fed + m * CSM_NQ + lane);
      txt_tok = 0;
    }
    if (p.mask) {
      my_mk = p.mask[m * (CSM_NQ + 1) + lane];
      txt_mk = p.mask[m * (CSM_NQ + 1) + CSM_NQ];
    } else {
      my_mk = 1;
      txt_mk = 0;
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const bool incol = col < H;
#pragma unroll 1
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
      const uint32_t otag = tg(p, ph);
      uint32_t* o = p.h_bb + (size_t)m * H + col;
      st_tag4(o, tw_pack(acc[0], otag), tw_pack(acc[1], otag), tw_pack(acc[2], otag), tw_pack(acc[3], otag));
      st_tag4(o + 4, tw_pack(acc[4], otag), tw_pack(acc[5], otag), tw_pack(acc[6], otag), tw_pack(acc[7], otag));
    }
  }
}

// ------------------------------------------------------------------ backbone decode attention (split-KV, GQA)
// One unit = (sequence, kv-head, 128 cached positions); the 4 (rep) query heads of the group share
// every K/V byte read.  Units write (max, sum, o[64]) partials; the last unit of a (sequence,
// kv-head) merges them and publishes the head outputs as tagged words.  q and the K/V of the position
// being processed are polled from the qkv phase's tagged output; older positions come from the cache.
// Softmax in fp32 (sdpa_attention_forward, hf integrations/sdpa_attention.py:40-104; a decode step
// attends to every cached position).  Out of line: 16 times per frame.
template <int REP>
__device__ __noinline__ void attn_bb_phase(const StreamParams& p, int layer, int src_ph, int ph) {
  constexpr int HD = 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, c = blockIdx.x, G = gridDim.x;
  const int Ttot = p.pos + 1;
  const int nsplit = (Ttot + CSM_ATT_SPLIT - 1) / CSM_ATT_SPLIT;
  const int nk = p.bb.kv;
  const int nunits = p.B * nk * nsplit;
  float* red = sm_red(p);
  volatile int* sflag = sm_flag();
  float* sm_o = red;                         // [8 warps][REP][64]
  float* sm_m = red + 8 * REP * HD;          // [8][REP]
  float* sm_l = sm_m + 8 * REP;              // [8][REP]
  const int grp = lane >> 3, dl = lane & 7;   // 4 positions per load, 8 lanes x 8 dims each
  const uint32_t qtag = tg(p, src_ph), otag = tg(p, ph);
  const int Wq = (p.bb.heads + 2 * nk) * HD;        // tagged q | k | v row
  compute_sync();   // (the previous phase ended without a CTA barrier; this one reuses its shared memory)
  for (int unit = c; unit < nunits; unit += G) {
    const int sp = unit % nsplit;
    const int kvh = (unit / nsplit) % nk;
    const int b = unit / (nsplit * nk);
    const size_t kvbase = (((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap * HD;
    const bf16* Kp = p.kc_bb + kvbase;
    const bf16* Vp = p.vc_bb + kvbase;
    const int pbase = sp * CSM_ATT_SPLIT + warp * 16;
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
      const uint32_t* qw = p.q_bb + (size_t)b * Wq + (kvh * REP) * HD + dl * 8;
      uint4 qa[REP], qb[REP];
      bool ok;
      unsigned spin = 0;
        do {
        ok = true;
#pragma unroll
        for (int h = 0; h < REP; ++h) {
          qa[h] = ld_tag4(qw + h * HD);
          qb[h] = ld_tag4(qw + h * HD + 4);
          ok &= tw_ok4(qa[h], qtag) & tw_ok4(qb[h], qtag);
        }
        if (!ok) poll_backoff(p, spin);
        if (!ok && spin_giveup(p, spin, ph, W_ATTN_BB_Q, (unsigned)unit)) ok = true;
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
        const uint32_t* kw = p.q_bb + (size_t)b * Wq + p.bb.heads * HD + kvh * HD + dl * 8;
        const uint32_t* vw = kw + nk * HD;
        uint4 k0, k1, v0, v1;
        unsigned spin = 0;
            do {
          k0 = ld_tag4(kw); k1 = ld_tag4(kw + 4);
          v0 = ld_tag4(vw); v1 = ld_tag4(vw + 4);
          if (spin_giveup(p, spin, ph, W_ATTN_BB_KV, (unsigned)unit)) break;
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
        for (int i = 0; i < 8; ++i) sm_o[(warp * REP + h) * HD + dl * 8 + i] = o[h][i];
        if (dl == 0) { sm_m[warp * REP + h] = mx[h]; sm_l[warp * REP + h] = ls[h]; }
      }
    }
    compute_sync();
    // merge the 8 warps: thread (h, d)
    if (tid < REP * HD) {
      const int h = tid / HD, d = tid % HD;
      float Mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) Mx = fmaxf(Mx, sm_m[w * REP + h]);
      float Ls = 0.f, O = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        float mw = sm_m[w * REP + h];
        float f = (mw == -INFINITY) ? 0.f : __expf(mw - Mx);
        Ls += f * sm_l[w * REP + h];
        O += f * sm_o[(w * REP + h) * HD + d];
      }
      float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + h) * p.nsplit_max + sp) * (HD + 2);
      part[2 + d] = O;
      if (d == 0) { part[0] = Mx; part[1] = Ls; }
    }
    compute_sync();
    if (tid == 0) {
      __threadfence();
      unsigned old = atomicAdd(p.attn_cnt + b * nk + kvh, 1u);
      sflag[0] = (old == (unsigned)nsplit - 1u);
    }
    compute_sync();
    if (sflag[0]) {
      __threadfence();
      if (tid < REP * HD) {
        const int h = tid / HD, d = tid % HD;
        const float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + h) * p.nsplit_max) * (HD + 2);
        float Mx = -INFINITY;
        for (int s2 = 0; s2 < nsplit; ++s2) Mx = fmaxf(Mx, ldcg_f32(part + (size_t)s2 * (HD + 2)));
        float Ls = 0.f, O = 0.f;
        for (int s2 = 0; s2 < nsplit; ++s2) {
          const float* ps = part + (size_t)s2 * (HD + 2);
          float f = __expf(ldcg_f32(ps) - Mx);
          Ls += f * ldcg_f32(ps + 1);
          O += f * ldcg_f32(ps + 2 + d);
        }
        st_tag(p.attn_bb + (size_t)b * (p.bb.heads * HD) + (kvh * REP + h) * HD + d, tw_pack(O / Ls, otag));
      }
      if (tid == 0) p.attn_cnt[b * nk + kvh] = 0u;
    }
    compute_sync();
  }
}

}  // namespace

// STOCH only makes the kernel's NAME unique per translation unit: template instantiations have weak linkage, two
// units instantiating csm_stream_kernel<1,4,true> with different bodies would be merged into one by the linker.
template <int NB, int REP, bool SMALL, bool STOCH>
__global__ void __launch_bounds__(CSM_THREADS, 1) csm_stream_kernel(const __grid_constant__ StreamParams p) {
  if (p.stop_flag != nullptr && *p.stop_flag) return;   // generation already ended (set by an earlier launch)
  if (CSM_GUARD && *reinterpret_cast<volatile int*>(p.abort_flag) != 0) return;   // an earlier launch timed out

  Lane L;
  L.tid = threadIdx.x;
  L.warp = threadIdx.x >> 5;
  L.lane = threadIdx.x & 31;
  L.c = blockIdx.x;
  L.G = gridDim.x;
  L.slot = L.slot_par = L.aslot = L.aslot_par = 0;
  L.ph = p.phase_begin;
  L.prof = nullptr;

  if (L.tid == 0) {
    for (int s = 0; s < CSM_MAX_SLOTS; ++s) {
      mbar_init(&sm_full()[s], 1);
      mbar_init(&sm_empty()[s], CSM_COMPUTE_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm_afull()[s], 1);
      mbar_init(&sm_aempty()[s], CSM_COMPUTE_WARPS);
    }
    *sm_prog() = 0u;
    mbar_fence_init();
  }
  if (L.tid < CSM_COMPUTE_THREADS) {
    // rope tables -> shared memory: decoder cos|sin for its 32 positions, backbone row `pos`
    bf16* rope = sm_rope();
    const int hd2 = p.dec.hd >> 1, hb2 = p.bb.hd >> 1;
    const int nd = CSM_DEC_POS * hd2;
    for (int i = L.tid; i < nd; i += CSM_COMPUTE_THREADS) {
      rope[i] = p.cos_dec[i];
      rope[nd + i] = p.sin_dec[i];
    }
    if (L.tid < hb2) {
      rope[2 * nd + L.tid] = p.cos_bb[(size_t)p.pos * hb2 + L.tid];
      rope[2 * nd + hb2 + L.tid] = p.sin_bb[(size_t)p.pos * hb2 + L.tid];
    }
    // first phase descriptor
    if (L.tid < 16)
      reinterpret_cast<uint4*>(&sm_desc()[p.phase_begin & 1])[L.tid] =
          __ldg(reinterpret_cast<const uint4*>(p.phases + p.phase_begin) + L.tid);
  }
  __syncthreads();
  const int bar_base = p.phases[p.phase_begin].bar_idx;   // grid-barrier events before/at the first phase

  if (L.warp == CSM_COMPUTE_WARPS) {
    // ===================== weight stream producer =====================
    if (L.lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      uint64_t* full = sm_full();
      uint64_t* empty = sm_empty();
      unsigned char* ring = sm_ring(p);
      uint32_t s = 0, round = 0, prog = 0;
#pragma unroll 1
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type != PH_GEMV) continue;
        if (p.progress != nullptr) p.progress[L.c * 4 + 2] = ph;
        const Geom g = csm_geom(P, L.c);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2;
#pragma unroll 1
        for (int ch = 0; ch < g.nchunks; ++ch) {
          const int tiles = min(g.tpc, g.ntiles - ch * g.tpc);
          const uint32_t bytes = (uint32_t)tiles * g.rows * 32u;
          if (round > 0) mbar_wait_g(p, &empty[s], (round - 1u) & 1u, ph, W_EMPTY);
          mbar_expect_tx(&full[s], bytes);
          if (p.evict_first) bulk_g2s_hint(ring + (size_t)s * p.slot_bytes, src, bytes, &full[s], pol);
          else bulk_g2s(ring + (size_t)s * p.slot_bytes, src, bytes, &full[s]);
          src += bytes;
          prog += bytes;
          *sm_prog() = prog;
          if (++s == (uint32_t)p.n_slots) { s = 0; ++round; }
        }
      }
    }
    return;
  }
  if (L.warp == CSM_COMPUTE_WARPS + 1) {
    // ===================== activation stream producer (K=8192 phases at batch > 4) =====================
    if (!SMALL && L.lane == 0) {
      uint32_t ait = 0;
      unsigned char* actreg = sm_act(p);
#pragma unroll 1
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type != PH_GEMV || P.act_mode != ACT_STREAM) continue;
        const Geom g = csm_geom(P, L.c);
        if (g.nchunks == 0) continue;
        const bf16* actp = P.act;
        const int act_stride = P.act_stride;
        if (p.use_barrier && ph > p.phase_begin) grid_wait(p, p.bar_counter, (unsigned)(P.bar_idx - bar_base) * L.G, ph);
        fence_proxy_async();
        const int astride_b = (g.tpc * 16 + 8) * 2;
#pragma unroll 1
        for (int ch = 0; ch < g.nchunks; ++ch) {
          const int tiles = min(g.tpc, g.ntiles - ch * g.tpc);
          const uint32_t rowbytes = (uint32_t)tiles * 32u;
          const uint32_t s = ait & 1u;
          if (ait >= 2u) mbar_wait_g(p, &sm_aempty()[s], ((ait >> 1) - 1u) & 1u, ph, W_AEMPTY);
          mbar_expect_tx(&sm_afull()[s], rowbytes * (uint32_t)p.B);
          unsigned char* dst = actreg + (size_t)s * (p.act_region_bytes / 2);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(actp) + (size_t)ch * g.tpc * 32;
          for (int m = 0; m < p.B; ++m)
            bulk_g2s(dst + (size_t)m * astride_b, src + (size_t)m * act_stride * 2, rowbytes, &sm_afull()[s]);
          ++ait;
        }
      }
    }
    return;
  }
  if (L.warp == CSM_COMPUTE_WARPS + 2) {
    // ===================== L2 prefetcher =====================
    // Issues HBM->L2 prefetches for this CTA's weight stream up to l2_ahead_bytes beyond what the ring
    // has requested, so that DRAM never idles while the ring is full and the compute warps are inside
    // a latency-bound stretch (staging, epilogue, attention).  Also: the norm weights of upcoming
    // phases (one CTA each) and the K/V blocks of this CTA's first backbone attention units.
    if (L.lane == 0 && p.l2_ahead_bytes > 0) {
      uint32_t pf = 0;
#pragma unroll 1
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type == PH_ATTN_BB) {
          const int Ttot = p.pos + 1, nk = p.bb.kv;
          constexpr int CSM_ATT_SPLIT_K = SMALL ? CSM_ATT_SPLIT : CSM_ATT_SPLIT_MMA;   // positions per unit of this kernel family
          const int nsplit = (Ttot + CSM_ATT_SPLIT_K - 1) / CSM_ATT_SPLIT_K;
          const int nunits = p.B * nk * nsplit;
          int done = 0;
          // (general kernels: one warp per unit, the first round of this CTA's eight warps are units c, G + c, ...
          // as well; later rounds are prefetched by the warps themselves)
          const int maxdone = SMALL ? 4 : CSM_COMPUTE_WARPS;
          for (int unit = L.c; unit < nunits && done < maxdone; unit += L.G, ++done) {
            const int sp = unit % nsplit, kvh = (unit / nsplit) % nk, b = unit / (nsplit * nk);
            const size_t off = ((((size_t)P.layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap + (size_t)sp * CSM_ATT_SPLIT_K) * 64;
            const int npos = min(CSM_ATT_SPLIT_K, p.pos - sp * CSM_ATT_SPLIT_K);   // cached positions only
            if (npos > 0) {
              bulk_prefetch_l2(p.kc_bb + off, (uint32_t)npos * 128u);
              bulk_prefetch_l2(p.vc_bb + off, (uint32_t)npos * 128u);
            }
          }
          continue;
        }
        if (P.type != PH_GEMV) continue;
        if (p.progress != nullptr) p.progress[L.c * 4 + 3] = ph;
        if (P.norm_w != nullptr && (ph % L.G) == L.c) bulk_prefetch_l2(P.norm_w, (uint32_t)P.K * 2u);
        const Geom g = csm_geom(P, L.c);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2;
        const uint32_t total = (uint32_t)g.rows * (uint32_t)P.K * 2u;
#pragma unroll 1
        for (uint32_t off = 0; off < total; off += 32768u) {
          const uint32_t n = min(32768u, total - off);
          uint32_t prog = *sm_prog();
          while ((int)(pf - prog) > p.l2_ahead_bytes) {
            __nanosleep(500);
            prog = *sm_prog();
            if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) return;
          }
          if ((int)(pf + n - prog) > 0) bulk_prefetch_l2(src + off, n);   // skip what the ring has already asked for
          pf += n;
        }
      }
    }
    return;
  }

  // ===================== compute warps =====================
#pragma unroll 1
  for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
    unsigned long long* prof = nullptr;   // debug stamps of the first and the last CTA: [cta][phase][16]
    if (p.prof != nullptr && L.tid == 0 && (L.c == 0 || L.c == L.G - 1))
      prof = p.prof + ((size_t)(L.c == 0 ? 0 : 1) * p.n_phases_total + ph) * 16;
    L.prof = prof;
    L.ph = ph;
    CSM_PROGRESS(p, L.c, L.tid, 0, ph);
    CSM_PROGRESS(p, L.c, L.tid, 1, 0);
    // descriptor of this phase is in shared memory (read in place); fetch the next one while this phase runs
    const Phase& P = sm_desc()[ph & 1];
    if ((P.flags & CSM_PF_BAR_IN) && p.use_barrier && ph > p.phase_begin) {
      if (L.tid == 0) grid_wait(p, p.bar_counter, (unsigned)(P.bar_idx - bar_base) * L.G, ph);
      compute_sync();
    }
    if (prof) prof[0] = clock64();       // (barrier observed)
    uint4 nxt = make_uint4(0, 0, 0, 0);
    const bool fetch = L.warp == CSM_COMPUTE_WARPS - 1 && L.lane < 16 && ph + 1 < p.phase_end;
    if (fetch) nxt = __ldg(reinterpret_cast<const uint4*>(p.phases + ph + 1) + L.lane);
    if (prof) prof[1] = clock64();       // phase body starts
    const int type = P.type;
    bool published = false;
    if (type == PH_GEMV) published = gemv_phase<NB, SMALL>(p, P, L, nxt, fetch);
    else if (!SMALL && type == PH_ATTN_DEC) attn_dec_phase(p, P, L);
    else if (type == PH_ATTN_BB) attn_bb_phase<REP>(p, P.layer, P.src_ph, ph);
    else if (type == PH_EMBED) embed_phase(p, ph);
    else finish_phase(p, P.res_ph);
    if (prof) prof[2] = clock64();       // this thread's share of the body done
    if (ph + 1 < p.phase_end) {
      const bool bar_out = (sm_desc()[ph & 1].flags & CSM_PF_BAR_OUT) && p.use_barrier;
      if (!published) {
        if (fetch) reinterpret_cast<uint4*>(&sm_desc()[(ph + 1) & 1])[L.lane] = nxt;
        compute_sync();                  // next descriptor visible
      } else if (bar_out) {
        compute_sync();
      }
      if (bar_out && L.tid == 0) {
        // release: everything this CTA wrote (ordered before by the CTA barrier) becomes visible before the count
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar_counter) : "memory");
      }
    }
    if (prof) prof[3] = clock64();       // end of phase
    if (p.prof != nullptr && L.tid == 0) {   // wall-clock end of this phase for every CTA (skew between CTAs)
      unsigned long long gt;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      p.prof[(size_t)32 * p.n_phases_total + (size_t)L.c * p.n_phases_total + ph] = gt;
    }
  }
}

// ------------------------------------------------------------------ host launch
// One instantiation per (batch n-tiles, backbone GQA ratio): the kernel then carries a single GEMV and
// attention variant, which keeps its instruction footprint small.
typedef void (*StreamKernel)(const StreamParams);

template <int NB, bool SMALL>
static StreamKernel pick_rep(int rep) {
  switch (rep) {
    case 1: return csm_stream_kernel<NB, 1, SMALL, CSM_BUILD_STOCH != 0>;
    case 2: return csm_stream_kernel<NB, 2, SMALL, CSM_BUILD_STOCH != 0>;
    default: return csm_stream_kernel<NB, 4, SMALL, CSM_BUILD_STOCH != 0>;
  }
}

#if CSM_BUILD_STOCH
#define CSM_LAUNCHER csm_launch_stream_small_stoch
#else
#define CSM_LAUNCHER csm_launch_stream_small
#endif

extern "C" cudaError_t CSM_LAUNCHER(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative) {
  const int rep = p->bb.heads / p->bb.kv;
  StreamKernel k = pick_rep<1, true>(rep);
  cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  if (cooperative) {
    void* args[] = {(void*)p};
    return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(CSM_THREADS), args, smem, stream);
  }
  k<<<grid, CSM_THREADS, smem, stream>>>(*p);
  return cudaGetLastError();
}
