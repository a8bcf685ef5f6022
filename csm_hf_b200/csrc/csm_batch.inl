// The frame engine for 3..32 sequences per GPU ("general" kernel family): ONE persistent kernel executes a whole audio
// frame for the whole batch -- backbone decode step, codebook-0 head, 32 decoder positions x 4 layers, 31 codebook heads,
// sampling and the embedding gathers between them -- as a table of ~750 dependent phases.
//
// Replaces, per frame, the ~7000 ATen kernel launches of CSMModel.generate_frame
// (reference modeling_csm.py:484-589 driving hf LlamaModel.forward x32).
//
// Difference from the <= 2-sequence family (csm_stream.inl), where every vector crossing CTAs is an array of tagged
// 32-bit words polled by the consumer: with 8-32 sequences a consumer CTA needs B x K activations per phase (64-128 KB),
// and polling them as tagged words through the LSU made staging the largest share of a frame (round-1 profile:
// 14-16 us of a 22-32 us decoder phase at 32 sequences).  Here
//   * every inter-phase vector (residual streams, q|k|v, attention outputs, MLP activations) is PLAIN bf16 in global
//     memory (L2-resident);
//   * consecutive phases are separated by ONE grid barrier (release-increment / acquire-poll of a counter), which also
//     orders the KV-cache writes of a qkv phase before the attention phase that reads them -- no cache row is ever
//     read without a release/acquire edge after its write;
//   * the activation rows of a phase are copied into shared memory by the TMA engine (cp.async.bulk, one copy per
//     row, issued by a dedicated warp right after it observes the barrier), not by load instructions;
//   * RMSNorm is applied in shared memory (LlamaRMSNorm's rounding points).
//
// Structure of a CTA (one per SM, 148 on B200):
//   warps 0..7  compute: RMSNorm in shared memory, tensor-core MMA on weight chunks, fused epilogues
//   warp  8     weight stream: walks the phase table ahead of the compute warps and keeps a ring of shared-memory
//               slots full with this CTA's slice of every weight matrix (cp.async.bulk completing on mbarriers);
//               never waits for anything but a free slot, so HBM keeps streaming across barriers
//   warp  9     activation stream: after the grid barrier of a phase, bulk-copies the phase's input rows -- whole rows
//               [B, K] for K <= 2048, a ring of [B, k-chunk] tiles in lockstep with the weight chunks for K = 8192
//   warp 10     L2 prefetcher: pulls this CTA's weight slices (and norm weights, and the K/V blocks of its attention
//               units) from HBM into L2 ahead of the ring
//
// Work split of a matrix W[N,K]: rows are divided evenly over the CTAs (granule 1 or 2 rows); csm_pack.cu stores each
// CTA's rows contiguously, k16-tile major, in ldmatrix order.  The WEIGHTS are the 16-row A operand of
// mma.sync.m16n8k16, the batch rows of the activations the 8-column B operand (NB = 1, 2 or 4 column tiles).
#include <string.h>

#include "csm_common.cuh"
#include "csm_sample.cuh"

#if !defined(CSM_BUILD_STOCH)
#error "include this file from csm_stream_general{,_stoch}.cu"
#endif

#ifndef CSM_MMA_UNROLL
#define CSM_MMA_UNROLL 4
#endif
#define CSM_STR2(x) #x
#define CSM_STR(x) CSM_STR2(x)

extern __shared__ __align__(128) unsigned char csm_smem[];

namespace {

// ---- shared-memory header (CSM_SM_HDR_BYTES = 4096) ----
//   [0,64) full[8] | [64,128) empty[8] | [160,168) sflag[2] | [168,172) weight-stream progress |
//   [256,768) 2 phase descriptors | [768,2816) 512 floats scratch | [2816,2944) tok[32] | [2944,3072) rstd[32] |
//   [3072,3136) afull[8] | [3136,3200) aempty[8] | [3200,3208) dfull | [3328,3584) attention stage barriers [8 warps][4] |
//   [3584,3616) attention piece counters [8 warps] | [3616,3624) xbar (K-half partials of the peer CTA have arrived)
__device__ __forceinline__ uint64_t* sm_full() { return reinterpret_cast<uint64_t*>(csm_smem); }
__device__ __forceinline__ uint64_t* sm_empty() { return reinterpret_cast<uint64_t*>(csm_smem + 64); }
__device__ __forceinline__ volatile int* sm_flag() { return reinterpret_cast<volatile int*>(csm_smem + 160); }
__device__ __forceinline__ volatile unsigned int* sm_prog() { return reinterpret_cast<volatile unsigned int*>(csm_smem + 168); }
__device__ __forceinline__ Phase* sm_desc() { return reinterpret_cast<Phase*>(csm_smem + 256); }
__device__ __forceinline__ float* sm_scratch() { return reinterpret_cast<float*>(csm_smem + 768); }
__device__ __forceinline__ int* sm_tok() { return reinterpret_cast<int*>(csm_smem + 2816); }
__device__ __forceinline__ float* sm_rstd() { return reinterpret_cast<float*>(csm_smem + 2944); }   // 32 floats
__device__ __forceinline__ uint64_t* sm_afull() { return reinterpret_cast<uint64_t*>(csm_smem + 3072); }
__device__ __forceinline__ uint64_t* sm_aempty() { return reinterpret_cast<uint64_t*>(csm_smem + 3136); }
__device__ __forceinline__ uint64_t* sm_dfull() { return reinterpret_cast<uint64_t*>(csm_smem + 3200); }
__device__ __forceinline__ uint64_t* sm_attbar() { return reinterpret_cast<uint64_t*>(csm_smem + 3328); }
__device__ __forceinline__ uint32_t* sm_attcnt() { return reinterpret_cast<uint32_t*>(csm_smem + 3584); }
__device__ __forceinline__ uint64_t* sm_xbar() { return reinterpret_cast<uint64_t*>(csm_smem + 3616); }
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

// Per-thread state of a compute warp (plain scalars, only ever passed to inlined code: stays in registers).
struct Lane {
  int tid, warp, lane, c, G;
  uint32_t slot, slot_par;                     // weight ring position of the consumer side
  uint32_t ait;                                // activation-ring chunks consumed so far (slot = ait % a_slots)
  uint32_t dpar;                               // parity of the next whole-row staging (dfull)
  uint32_t xpar;                               // parity of the next pair exchange (xbar)
  int ph;                                      // phase being executed
  unsigned long long* prof;                    // debug stamps of this phase (thread 0 of the first / last CTA) or null
};

#define CSM_STAMP(L, i)                     \
  do {                                      \
    if ((L).prof) (L).prof[i] = clock64();  \
  } while (0)
#define CSM_PROGRESS(p, c, tid, slot, v)                                      \
  do {                                                                        \
    if ((p).progress != nullptr && (tid) == 0) (p).progress[(c) * 4 + (slot)] = (v); \
  } while (0)

// ---- hang guard ----
// Every wait in this kernel is a spin on memory another CTA (or the TMA engine) will write.  A protocol bug or a lost
// CTA would otherwise wedge the GPU for good; instead a wait that lasts longer than ~2 s records who waited for what
// and raises the abort flag, which makes every wait in the grid give up and every later launch return at once; the
// host reports it as an error (csm_frames_done / csm_generate_frame).  Cost: one counter increment per failed poll.
enum WaitId { W_STAGE = 1, W_CAND = 2, W_ATTN_DEC = 3, W_ATTN_BB_Q = 4, W_ATTN_BB_KV = 5, W_RESID = 6, W_GRID = 7,
              W_FULL = 8, W_AFULL = 9, W_EMPTY = 10, W_AEMPTY = 11, W_DFULL = 12, W_PAIR = 13 };

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
// call once per failed poll (one add and one test on the fast path); true = give up
__device__ __forceinline__ bool spin_giveup(const StreamParams& p, unsigned& n, int ph, int id, unsigned a = 0,
                                            unsigned b = 0) {
  if ((++n & 0xffffu) != 0) return false;
  return spin_slow(p, n, ph, id, a, b);
}
// Grid barrier, waiting side: the counter has been incremented (with release) by every CTA after each phase it
// finished; phase `ph` may start once all G CTAs have finished phase ph-1.
__device__ __forceinline__ void grid_wait(const StreamParams& p, unsigned target, int ph) {
  unsigned n = 0;
  // relaxed polls (no fence per round trip), then ONE acquire load that synchronises with the release increments
  while (ld_tag(p.bar_counter) < target) {
    if (spin_giveup(p, n, ph, W_GRID, target)) break;
  }
  (void)ld_acquire_u32(p.bar_counter);
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
__device__ __forceinline__ void st_bf16(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------ greedy sample of a finished head phase
// sample_topk at topk=1 (modeling_csm.py:179-189) with the canonical lowest-index tie-break: reduce the (best logit,
// id) candidates every CTA published in head phase `head_ph` for codebook `cb` (tagged 64-bit words; after the grid
// barrier they are all there, the tag check is a guard).  Result in tok[m]; CTA 0 also records samples / fed.
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
// sample_topk(logits, topk, temperature) (modeling_csm.py:179-189) for codebook `cb`: every CTA reads the tagged
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
// One warp per (sequence, query head), units spread over the CTAs.  Lane t owns cached position t for the scores and
// output dims 4*lane.. for P.V.  Every position 0..dec_pos comes from the cache: the qkv phase wrote position dec_pos
// before the grid barrier that precedes this phase.  q (plain bf16 row of the qkv phase's output) is spread to all
// lanes through a 512-byte shared-memory row of the warp.  Softmax in fp32 (sdpa_attention_forward).
__device__ __forceinline__ void attn_dec_phase(const StreamParams& p, const Phase& P, const Lane& L) {
  constexpr int HD = 128;
  const int nh = p.dec.heads, nk = p.dec.kv, rep = nh / nk;
  const int nunits = p.B * nh;
  const int W = (nh + 2 * nk) * HD;
  const int dec_pos = P.dec_pos, layer = P.layer, lane = L.lane;
  const bf16* qrow = reinterpret_cast<const bf16*>(p.q_dec);
  bf16* orow = reinterpret_cast<bf16*>(p.attn_dec);
  float* qs = reinterpret_cast<float*>(sm_act(p)) + L.warp * HD;   // (the act region is free during this phase)
  const float sc = p.dec.scale;
#pragma unroll 1
  for (int unit = L.warp * L.G + L.c; unit < nunits; unit += CSM_COMPUTE_WARPS * L.G) {
    const int b = unit / nh, head = unit - b * nh, kvh = head / rep;
    const size_t kvbase = (((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)CSM_DEC_POS * HD;
    const bf16* kp = p.kc_dec + kvbase + (size_t)lane * HD;
    const bf16* vp = p.vc_dec + kvbase + lane * 4;
    // ONE batch of independent loads -- q, this lane's K row, this lane's dims of the first 16 V rows -- so that the unit
    // costs one L2 round trip instead of three in a row (the phase is a latency chain, nothing here is bandwidth)
    const uint2 q2 = ldcg_u2(qrow + (size_t)b * W + head * HD + lane * 4);
    uint4 kk[HD / 8];
#pragma unroll
    for (int ci = 0; ci < HD / 8; ++ci) kk[ci] = ldcg_u4(kp + ci * 8);   // (rows > dec_pos: in bounds, masked below)
    uint2 vv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) vv[j] = ldcg_u2(vp + (size_t)j * HD);
    __syncwarp();            // (the previous unit's reads of qs are done)
    *reinterpret_cast<float4*>(qs + lane * 4) = make_float4(bf_lo(q2.x) * sc, bf_hi(q2.x) * sc, bf_lo(q2.y) * sc, bf_hi(q2.y) * sc);
    __syncwarp();            // qs written by all lanes before any lane reads it
    float d = 0.f;
#pragma unroll
    for (int ci = 0; ci < HD / 8; ++ci) {
      const float4 a = *reinterpret_cast<const float4*>(qs + ci * 8), c4 = *reinterpret_cast<const float4*>(qs + ci * 8 + 4);
      d += a.x * bf_lo(kk[ci].x) + a.y * bf_hi(kk[ci].x) + a.z * bf_lo(kk[ci].y) + a.w * bf_hi(kk[ci].y);
      d += c4.x * bf_lo(kk[ci].z) + c4.y * bf_hi(kk[ci].z) + c4.z * bf_lo(kk[ci].w) + c4.w * bf_hi(kk[ci].w);
    }
    const float s = lane <= dec_pos ? d : -INFINITY;
    const float mx = warp_max(s);
    const float pe = (lane <= dec_pos) ? __expf(s - mx) : 0.f;   // (0 for positions past dec_pos: stale cache rows drop out)
    const float l = warp_sum(pe);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float pv = __shfl_sync(0xffffffffu, pe, j);
      if (j <= dec_pos) {
        o0 += pv * bf_lo(vv[j].x); o1 += pv * bf_hi(vv[j].x);
        o2 += pv * bf_lo(vv[j].y); o3 += pv * bf_hi(vv[j].y);
      }
    }
    if (dec_pos >= 16) {   // (warp-uniform) second half of the positions
#pragma unroll
      for (int j = 0; j < 16; ++j) vv[j] = (16 + j <= dec_pos) ? ldcg_u2(vp + (size_t)(16 + j) * HD) : make_uint2(0, 0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pv = __shfl_sync(0xffffffffu, pe, 16 + j);
        if (16 + j <= dec_pos) {
          o0 += pv * bf_lo(vv[j].x); o1 += pv * bf_hi(vv[j].x);
          o2 += pv * bf_lo(vv[j].y); o3 += pv * bf_hi(vv[j].y);
        }
      }
    }
    const float inv = 1.f / l;
    *reinterpret_cast<uint2*>(orow + (size_t)b * (nh * HD + p.hpad) + head * HD + lane * 4) =
        make_uint2(pack_bf16(o0 * inv, o1 * inv), pack_bf16(o2 * inv, o3 * inv));
  }
}

// ------------------------------------------------------------------ RMSNorm of the staged rows, in shared memory
// Rows [M][astride] bf16 (raw residual-stream rows copied by the TMA engine) -> LlamaRMSNorm.forward
// (hf modeling_llama.py:62-67): fp32 x*rsqrt(mean(x^2)+eps) -> bf16 -> *w -> bf16, in place.  ONE WARP PER ROW: a lane
// holds elements 8*lane + 256*j of its row in registers, the sum of squares is one in-register accumulation and one
// shuffle reduction per row (fixed order: a row's result does not depend on the batch it is in), and the warp scales
// its own row right away -- one pass over shared memory, no CTA barrier inside.  K is a multiple of 256, <= 2048.
// GATHER = true: the rows do not come from shared memory but from the pre-projected embedding table --
// projection(_embed_audio(codebook, token)) (modeling_csm.py:247-259,564-565) = row token + codebook*V, the token being
// the sample of the previous head phase (tok[]); the warp of the CTA that owns the sequence also starts the residual
// stream with the raw row.
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {   // round-to-nearest bf16 products of two packed pairs
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float sumsq8(const uint4& x) {   // sum of squares of 8 bf16, as a tree (short dependency chain)
  const float a0 = bf_lo(x.x), a1 = bf_hi(x.x), a2 = bf_lo(x.y), a3 = bf_hi(x.y);
  const float a4 = bf_lo(x.z), a5 = bf_hi(x.z), a6 = bf_lo(x.w), a7 = bf_hi(x.w);
  return ((a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3)) + ((a4 * a4 + a5 * a5) + (a6 * a6 + a7 * a7));
}

template <bool GATHER>
__device__ __forceinline__ void norm_rows(const StreamParams& p, const Phase& P, const Lane& L, int astride) {
  const int K = P.K, M = p.B, nj = K >> 8;   // 256 elements per warp pass
  bf16* dst = reinterpret_cast<bf16*>(sm_act(p));
  const float eps = P.stack ? p.dec.eps : p.bb.eps;
  const float fK = (float)K;
  const bool keep_norm = !GATHER && P.norm_out != nullptr;   // copy of the normalised rows (last_hidden_state)
  const bf16* tab = GATHER ? P.act + (size_t)(P.cb * p.V) * K : nullptr;
  const int* tok = sm_tok();
  const uint4* nwp = reinterpret_cast<const uint4*>(sm_act(p) + p.normw_off) + L.lane;   // norm weights, staged with the rows
#pragma unroll 2
  for (int m = L.warp; m < M; m += CSM_COMPUTE_WARPS) {
    bf16* row = dst + (size_t)m * astride + L.lane * 8;
    float ss = 0.f;
    if (GATHER) {
      // pass 1 reads the table row (global), keeps the raw row in shared memory and, for the owner CTA of the
      // sequence, starts the residual stream with it
      const uint4* src = reinterpret_cast<const uint4*>(tab + (size_t)tok[m] * K) + L.lane;
      uint4* hres = reinterpret_cast<uint4*>(P.norm_out + (size_t)m * (K + p.hpad)) + L.lane;
      const bool own = (m % L.G) == L.c;
#pragma unroll 4
      for (int j = 0; j < nj; ++j) {
        const uint4 x = __ldg(src + j * 32);
        *reinterpret_cast<uint4*>(row + j * 256) = x;
        if (own) hres[j * 32] = x;
        ss += sumsq8(x);
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < nj; ++j) {
        const uint4 x = *reinterpret_cast<const uint4*>(row + j * 256);
        ss += sumsq8(x);
      }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / fK + eps);   // mean = sum / K exactly as torch
    // pass 2: the lane re-reads its own elements (written by itself: no barrier) and scales them
#pragma unroll 4
    for (int j = 0; j < nj; ++j) {
      const uint4 x = *reinterpret_cast<const uint4*>(row + j * 256);
      const uint4 nw = nwp[j * 32];
      uint4 o;   // bf16(x * rstd) as a packed pair, then the bf16 x bf16 -> bf16 product with the weight (one rounding each)
      o.x = mul_bf16x2(nw.x, pack_bf16(bf_lo(x.x) * rstd, bf_hi(x.x) * rstd));
      o.y = mul_bf16x2(nw.y, pack_bf16(bf_lo(x.y) * rstd, bf_hi(x.y) * rstd));
      o.z = mul_bf16x2(nw.z, pack_bf16(bf_lo(x.z) * rstd, bf_hi(x.z) * rstd));
      o.w = mul_bf16x2(nw.w, pack_bf16(bf_lo(x.w) * rstd, bf_hi(x.w) * rstd));
      *reinterpret_cast<uint4*>(row + j * 256) = o;
      if (keep_norm && (m % L.G) == L.c) *reinterpret_cast<uint4*>(P.norm_out + (size_t)m * K + j * 256 + L.lane * 8) = o;
    }
  }
}

// ------------------------------------------------------------------ tensor-core inner loop
// One ring chunk holds this CTA's rows for `tiles` k16-tiles as [tile][k-half][row][8 bf16] (csm_pack.cu), so the 16x16 A
// fragment of an m-tile is ONE ldmatrix.x4 (four conflict-free 8x8 matrices).  The B fragments (8 batch rows x 16 k) of
// two k-tiles come from the activation rows with one more ldmatrix.x4.  Rows past the CTA's last weight row and batch
// rows past M re-read a valid row: they only feed accumulator rows / columns that are never stored.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// All chunks of one phase for this warp; partial sums -> red[kg][m][rows_pad].  The warp owns m-tiles mt0 (and mt1 when
// the CTA has more m-tiles than m-tile groups) and every ks-th k16-tile; two k-tiles per iteration.  With a single
// m-tile the two k-tiles of an iteration feed the two accumulator sets (two independent MMA chains); otherwise
// accumulator set j belongs to m-tile j.
template <int NB>
__device__ __forceinline__ void gemv_core(const StreamParams& p, const Phase& P, const GeoC& gc, Lane& L, bool stream,
                                          int astride) {
  const int M = p.B;
  const int rows = gc.rows, mtiles = gc.mtiles, rows_pad = gc.rows_pad, ksl = gc.ksl;
  int tpc = gc.tpc, nchunks = gc.nch, ntiles = P.pair ? (P.K >> 5) : (P.K >> 4);   // (pair phases: this CTA's K half)
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
  const uint32_t nsa = (uint32_t)p.a_slots;

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
      as = L.ait % nsa;
      mbar_wait_g(p, &sm_afull()[as], (L.ait / nsa) & 1u, L.ph, W_AFULL);
      aadd = as * (uint32_t)p.a_slot_bytes;
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
    if (stream) ++L.ait;
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
template <int NB>
__device__ __forceinline__ void gemv_phase(const StreamParams& p, const Phase& P, Lane& L, unsigned bar_target) {
  const int M = p.B, K = P.K;
  const bool stream = P.act_mode == ACT_STREAM;
  const int cc = P.pair ? (L.c >> 1) : L.c;    // pair phases: q, r, geo are per CTA pair
  const bool hi = cc < P.r;
  const GeoC& gc = P.geo[hi ? 0 : 1];
  const int row0 = (cc * P.q + (hi ? cc : P.r)) * P.gran;
  const int rows = gc.rows;
  int astride;
  if (!stream) {
    astride = K + 8;
    if (P.act_mode == ACT_GATHER) {
      // (no TMA staging: the rows are gathered by the compute warps, which therefore observe the barrier themselves)
      if (bar_target) {
        if (L.tid == 0) grid_wait(p, bar_target, L.ph);
        compute_sync();
      }
      CSM_STAMP(L, 8);
#if CSM_BUILD_STOCH
      sample_tokens(p, P.cb, P.res_ph);
#else
      reduce_candidates(p, L.warp, L.lane, L.c, L.G, P.cb, P.res_ph);
#endif
      mbar_wait_g(p, sm_dfull(), L.dpar, L.ph, W_DFULL);   // (the norm weights, copied by the activation-stream warp)
      L.dpar ^= 1u;
      norm_rows<true>(p, P, L, astride);
    } else {
      // the activation-stream warp copies the rows after it has observed the grid barrier of this phase
      mbar_wait_g(p, sm_dfull(), L.dpar, L.ph, W_DFULL);
      L.dpar ^= 1u;
      CSM_STAMP(L, 8);
      if (P.act_mode == ACT_NORM) norm_rows<false>(p, P, L, astride);
    }
    compute_sync();
    CSM_STAMP(L, 4);   // activations staged
    CSM_PROGRESS(p, L.c, L.tid, 1, 1);
  } else {
    astride = gc.tpc * 16 + 8;
  }
  // epilogue mapping: thread -> (batch row m, granule u), granules padded to a power of two (host: GeoC::ush)
  const int upc = gc.upc, ush = gc.ush;
  const int u = L.tid & ((1 << ush) - 1);
  const int mstep = ush >= 8 ? 1 : CSM_COMPUTE_THREADS >> ush;
  const int m_first = ush >= 8 ? 0 : L.tid >> ush;
  const int epi = P.epi;
  const int out_stride = P.out_stride;
  // residual element of the first row this thread will update: requested now, used after the MMAs (its L2 round trip
  // would otherwise sit on the critical path of the epilogue).  Written >= one grid barrier ago.
  float resid0 = 0.f;
  if (epi == EPI_RESID && !P.pair && u < upc && m_first < M)
    resid0 = ldcg_bf16(P.out + (size_t)m_first * out_stride + row0 + u * P.gran);
  if (rows > 0) gemv_core<NB>(p, P, gc, L, stream, astride);
  CSM_STAMP(L, 5);     // this warp's MMAs done
  compute_sync();
  CSM_STAMP(L, 6);     // all warps' MMAs done
  CSM_PROGRESS(p, L.c, L.tid, 1, 2);

  if (P.pair) {
    // ---- CTA-pair phase (streamed down_proj): this CTA holds the partial sums of its K half for the rows of BOTH CTAs.
    // It finishes the rows [lo, hi_) of its own half of the pair's rows and sends the partials of the other rows to the
    // peer's exchange buffer through distributed shared memory (st.shared::cluster), then one release-arrive per warp on
    // the peer's mbarrier; the peer's contribution for its own rows arrives the same way.  EPI_RESID only.
    const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
    const int half = (rows + 1) >> 1;
    const int lo = rank ? half : 0, hi_ = rank ? rows : half, plo = rank ? 0 : half;
    const int ks = 1 << gc.ksl, rows_pad = gc.rows_pad, kstride = p.m_alloc * rows_pad;
    const float* red = sm_red(p);
    float* xbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sm_red(p)) + p.xbuf_off);   // [m_alloc][16]
    const uint32_t peer_xbuf = mapa_u32(smem_u32(xbuf), peer), peer_xbar = mapa_u32(smem_u32(sm_xbar()), peer);
    const bool mine = u >= lo && u < hi_;
    float vown[4] = {0.f, 0.f, 0.f, 0.f};   // (M <= 32, mstep >= 8: at most 4 rows m per thread)
    int it = 0;
    if (u < upc) {
#pragma unroll 1
      for (int m = m_first; m < M; m += mstep, ++it) {
        const float* r = red + (size_t)m * rows_pad + u;
        float v = 0.f;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (kk < ks) v += r[kk * kstride];
        if (mine) {
          if (it == 0) vown[0] = v; else if (it == 1) vown[1] = v; else if (it == 2) vown[2] = v; else vown[3] = v;
        } else {
          st_cluster_f32(peer_xbuf + (uint32_t)(m * 16 + (u - plo)) * 4u, v);
        }
      }
    }
    __syncwarp();
    if (L.lane == 0) mbar_arrive_remote(peer_xbar);           // 8 warps of the peer CTA complete my xbar
    {
      unsigned n = 0;
      while (!mbar_try_wait_cluster(sm_xbar(), L.xpar)) {
        if (spin_giveup(p, n, L.ph, W_PAIR, rank)) break;
      }
      L.xpar ^= 1u;
    }
    if (u < upc && mine) {
      it = 0;
#pragma unroll 1
      for (int m = m_first; m < M; m += mstep, ++it) {
        const float mv = it == 0 ? vown[0] : (it == 1 ? vown[1] : (it == 2 ? vown[2] : vown[3]));
        const float pv = xbuf[m * 16 + (u - lo)];
        // (K half 0) + (K half 1), in this order on both CTAs
        const float v0 = bfround(rank ? pv + mv : mv + pv);   // nn.Linear output is bf16
        bf16* o = P.out + (size_t)m * out_stride + row0 + u;
        st_bf16(o, ldcg_bf16(o) + v0);                         // hf modeling_llama.py:331: residual + f(x), both bf16
      }
    }
    return;
  }
  // ---- fused epilogues (every output is plain bf16; the grid barrier after the phase publishes it)
  if (u < upc) {
    const int gran = P.gran, ks = 1 << gc.ksl, rows_pad = gc.rows_pad;
    float* red = sm_red(p);
    const int n = u * gran;
    const int gn = row0 + n;   // packed row index
    const int kstride = p.m_alloc * rows_pad;
#pragma unroll 2
    for (int m = m_first; m < M; m += mstep) {
      float v0 = 0.f, v1 = 0.f;
      {
        // split-K partials: all loads issued before the adds (ks <= 8)
        const float* r = red + (size_t)m * rows_pad + n;
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
        bf16* o = P.out + (size_t)m * out_stride + gn;
        st_bf16(o, (m == m_first ? resid0 : ldcg_bf16(o)) + v0);
      } else if (epi == EPI_SWIGLU) {  // hf modeling_llama.py:183: bf16(silu(gate)) * up -> bf16 ; rows (2j,2j+1)=(gate_j,up_j)
        const float sl = bfround(v0 / (1.f + expf(-v0)));
        const int j = gn >> 1;
        if (P.tile_sh) {   // [k-tile][m_alloc][len+8]: the streamed down_proj copies one [B, len+8] tile per bulk copy
          const int len = 1 << P.tile_sh;
          st_bf16(P.out + ((size_t)(j >> P.tile_sh) * p.m_alloc + m) * (len + 8) + (j & (len - 1)), sl * v1);
        } else {
          st_bf16(P.out + (size_t)m * out_stride + j, sl * v1);
        }
      } else if (epi == EPI_QKV) {     // rows (2j,2j+1) = RoPE pair (i, i+hd/2) of q/k, or two adjacent v features
        const StackDims& sd = P.stack ? p.dec : p.bb;
        const int half = sd.hd >> 1, hl = sd.hdl - 1;
        const int pidx = gn >> 1;
        const int nq = sd.heads << hl, nk = sd.kv << hl;
        const int pos = P.stack ? P.dec_pos : p.pos;
        const int cap = P.stack ? CSM_DEC_POS : p.Tcap;
        bf16* qrow = P.out + (size_t)m * out_stride;   // q | k | v of this position
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
          if (isq) {
            bf16* qd = qrow + head * sd.hd + i;
            st_bf16(qd, o1);
            st_bf16(qd + half, o2);
          } else {   // DynamicCache.update (hf cache_utils.py:102-121) as an in-place write at `pos`
            bf16* kc = P.stack ? p.kc_dec : p.kc_bb;
            bf16* dstp = kc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + i;
            st_bf16(dstp, o1);
            st_bf16(dstp + half, o2);
          }
        } else {
          const int f = (pidx - nq - nk) * 2;
          const int head = f >> sd.hdl, d = f & (sd.hd - 1);
          bf16* vc = P.stack ? p.vc_dec : p.vc_bb;
          bf16* dstp = vc + ((((size_t)P.layer * p.Bmax + m) * sd.kv + head) * cap + pos) * sd.hd + d;
          *reinterpret_cast<uint32_t*>(dstp) = pack_bf16(v0, v1);
        }
      } else if (epi == EPI_STORE) {
        st_bf16(P.out + (size_t)m * out_stride + gn, v0);
      } else {   // EPI_HEAD
        if (P.out) st_bf16(P.out + (size_t)m * out_stride + gn, v0);
        if (CSM_BUILD_STOCH) st_tag(p.lgt + (size_t)m * p.lgt_stride + gn, tw_pack(v0, tg(p, L.ph)));   // for sample_tokens
        red[(size_t)m * rows_pad + n] = v0;   // kk = 0 plane, own element only
      }
    }
  }
  if (epi == EPI_HEAD) {
    // publish this CTA's best (logit, id) per sequence as one tagged 64-bit word; the next phase reduces them
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
}

// ------------------------------------------------------------------ end of frame
// After the last head: sample codebook 31, publish the 32 ids (modeling_csm.py:657-666) and evaluate
// the stop rule torch.all(new_frame == 0) (:662).  CTA 0 only.  Out of line: once per frame.
__device__ __noinline__ void finish_phase(const StreamParams& p, int head_ph) {
  const int tid = threadIdx.x, c = blockIdx.x;
  if (c != 0) return;
  const int M = p.B;
  volatile int* sflag = sm_flag();
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
  bf16* hb = reinterpret_cast<bf16*>(p.h_bb);
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
    if (incol)
      *reinterpret_cast<uint4*>(hb + (size_t)m * (H + p.hpad) + col) =
          make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

// Merge of the split-KV partials of one (sequence, kv-head) by the warp whose unit arrived last: lane (grp, dl) owns
// dims 8 dl.. of head grp.  All loads of a pass are independent and issued together: the (max, sum) pairs of every
// split first, then the 64-float partial outputs as 16-byte vectors, four splits (8 loads) at a time.
template <int REP>
__device__ __forceinline__ void merge_splits(const StreamParams& p, int b, int kvh, int nsplit, int lane, bf16* orows) {
  constexpr int HD = 64, PSTR = HD + 4;
  const int grp = lane >> 3, dl = lane & 7;
#pragma unroll
  for (int h0 = 0; h0 < REP; h0 += 4) {
    const int h = h0 + grp;
    if (h < REP) {   // (REP < 4: the upper lane groups idle)
      const float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + h) * p.nsplit_max) * PSTR;
      float M2 = -INFINITY;
#pragma unroll 1
      for (int s0 = 0; s0 < nsplit; s0 += 8) {
        float mv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mv[i] = (s0 + i < nsplit) ? ldcg_f32(part + (size_t)(s0 + i) * PSTR) : -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) M2 = fmaxf(M2, mv[i]);
      }
      float L2 = 0.f, O2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) O2[i] = 0.f;
#pragma unroll 1
      for (int s0 = 0; s0 < nsplit; s0 += 4) {
        float2 ml[4];
        uint4 oa[4], ob[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* ps = part + (size_t)min(s0 + i, nsplit - 1) * PSTR;
          ml[i] = __ldcg(reinterpret_cast<const float2*>(ps));
          oa[i] = ldcg_u4(ps + 4 + dl * 8);
          ob[i] = ldcg_u4(ps + 8 + dl * 8);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (s0 + i < nsplit) {
            const float f = __expf(ml[i].x - M2);
            L2 += f * ml[i].y;
            O2[0] += f * __uint_as_float(oa[i].x); O2[1] += f * __uint_as_float(oa[i].y);
            O2[2] += f * __uint_as_float(oa[i].z); O2[3] += f * __uint_as_float(oa[i].w);
            O2[4] += f * __uint_as_float(ob[i].x); O2[5] += f * __uint_as_float(ob[i].y);
            O2[6] += f * __uint_as_float(ob[i].z); O2[7] += f * __uint_as_float(ob[i].w);
          }
        }
      }
      const float inv = 1.f / L2;
      *reinterpret_cast<uint4*>(orows + (size_t)b * (p.bb.heads * HD + p.hpad) + (kvh * REP + h) * HD + dl * 8) =
          make_uint4(pack_bf16(O2[0] * inv, O2[1] * inv), pack_bf16(O2[2] * inv, O2[3] * inv),
                     pack_bf16(O2[4] * inv, O2[5] * inv), pack_bf16(O2[6] * inv, O2[7] * inv));
    }
  }
}

// ------------------------------------------------------------------ backbone decode attention (split-KV, GQA)
// ONE WARP per unit = (sequence, kv-head, nsub x 128 cached positions), no CTA barrier inside the phase, Q.K^T and P.V
// on mma.sync.m16n8k16; the REP query heads of the group share every K/V byte read.  Every position 0..pos comes from
// the cache: the qkv phase wrote position `pos` before the grid barrier that precedes this phase.
//
//   staging   : a unit's K rows (128 bytes per position) and V rows are contiguous in the cache, so they travel as bulk
//               copies of 32-128 positions (4-16 KB) through a per-warp ring of 1-4 stages (the activation region, free
//               during this phase), each completing on its own mbarrier; lane 0 issues piece n + stages as soon as the warp
//               has consumed piece n -- across sub-block and unit boundaries, so loads stay in flight while a unit is
//               reduced and merged.  The L2 prefetch warp pulls the K/V of the CTA's first units into L2 ahead of time.
//   S = Q K^T : A = the REP query heads of the group (rows >= REP are zero), B = 8 cached positions per n-tile, k = the
//               64 head dims in 4 steps.  The k index of an MMA is a free permutation as long as A and B agree: lane
//               (g, t) supplies dims 8t..8t+7 and 32+8t..32+8t+7 (two 16-byte shared-memory loads per K row).
//   softmax   : fp32 (sdpa_attention_forward computes its softmax in fp32); the scores of a 128-position sub-block stay
//               in registers; between sub-blocks the running (max, sum, o) are rescaled (online softmax).
//   O = P V   : the S accumulator fragments of a chunk are the A fragment of P (positions as k).  P is split into
//               bf16 hi + lo parts (two MMAs), which keeps ~16 mantissa bits of the fp32 probabilities.  B = V with the
//               output dims permuted (column g of n-tile j is dim 8g + j): lane (g, t) reads V as one 16-byte load per
//               position and builds the fragments with byte permutes.
// nsub (host: launch_frame) grows with the batch so that a launch has about one unit per warp: at 32 sequences a unit
// is 512 positions -- 4x fewer partial results, fences, counters and merges than with 128-position units.
// Units write (max, sum, o[64]) partials; the last unit of a (sequence, kv-head) to arrive merges the splits and writes
// the head outputs (plain bf16).  A sequence's result does not depend on the batch it is in for a given nsub.
template <int REP, int PS>
__device__ __noinline__ void attn_bb_phase_tma(const StreamParams& p, int layer, int ph) {
  constexpr int HD = 64, SUB = CSM_ATT_SPLIT_MMA, NCH = SUB / 16, NP = 2 * SUB / PS, PBYTES = PS * HD * 2;
  static_assert(REP <= 8, "query heads per kv head");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.x, G = gridDim.x;
  const int Ttot = p.pos + 1;
  const int nsub = p.att_nsub, split = SUB * nsub;
  const int nsplit = (Ttot + split - 1) / split;
  const int nk = p.bb.kv;
  const int nunits = p.B * nk * nsplit;
  const int g = lane >> 2, t = lane & 3;        // MMA fragment coordinates
  const int Wq = (p.bb.heads + 2 * nk) * HD;    // q | k | v row
  const float scale = p.bb.scale;
  const bf16* qrows = reinterpret_cast<const bf16*>(p.q_bb);
  bf16* orows = reinterpret_cast<bf16*>(p.attn_bb);
  const uint32_t nstg = (uint32_t)p.att_stages;
  unsigned char* stages = sm_act(p) + (size_t)warp * nstg * PBYTES;
  uint64_t* bars = sm_attbar() + warp * 4;
  const int first = warp * G + c, ustep = CSM_COMPUTE_WARPS * G;
  const uint32_t npu = (uint32_t)(NP * nsub);   // pieces per unit
  const uint32_t n0 = sm_attcnt()[warp];        // pieces this warp has streamed in earlier phases (barrier parities)
  if (lane == 0) fence_proxy_async();           // (cache rows written by generic stores before the grid barrier)
  // piece nn of this phase: unit k = nn / npu of this warp, sub-block sb, q = piece inside it: q < NP/2 -> K rows, else V rows
  auto issue = [&](uint32_t nn) {
    const int k = (int)(nn / npu), r = (int)(nn % npu), sb = r / NP, q = r % NP;
    const int unit = first + k * ustep;
    if (unit >= nunits) return;
    const int sp = unit % nsplit, kvh = (unit / nsplit) % nk, b = unit / (nsplit * nk);
    const int pp = sp * split + sb * SUB + (q % (NP / 2)) * PS;         // first position of the piece
    const int nv = min(PS, Ttot - pp);
    const uint32_t slot = (n0 + nn) % nstg;
    if (nv <= 0) { mbar_arrive(&bars[slot]); return; }                  // nothing cached there: complete the phase empty
    const bf16* src = (q < NP / 2 ? p.kc_bb : p.vc_bb) +
                      ((((size_t)layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap + pp) * HD;
    mbar_expect_tx(&bars[slot], (uint32_t)nv * (HD * 2));
    bulk_g2s(stages + (size_t)slot * PBYTES, src, (uint32_t)nv * (HD * 2), &bars[slot]);
  };
  if (lane == 0)
    for (uint32_t i = 0; i < nstg; ++i) issue(i);
  uint32_t nn = 0;
#pragma unroll 1
  for (int unit = first; unit < nunits; unit += ustep) {
    const int sp = unit % nsplit;
    const int kvh = (unit / nsplit) % nk;
    const int b = unit / (nsplit * nk);
    // Q fragment: head g (rows >= REP are zero), this lane's 16 dims as 8 packed pairs
    uint32_t qf[8];
    {
      const bf16* qw = qrows + (size_t)b * Wq + (kvh * REP + (g < REP ? g : 0)) * HD;
      const uint4 q0 = ldcg_u4(qw + 8 * t), q1 = ldcg_u4(qw + 32 + 8 * t);
      qf[0] = q0.x; qf[1] = q0.y; qf[2] = q0.z; qf[3] = q0.w; qf[4] = q1.x; qf[5] = q1.y; qf[6] = q1.z; qf[7] = q1.w;
      if (g >= REP) {
#pragma unroll
        for (int i = 0; i < 8; ++i) qf[i] = 0u;
      }
    }
    float mrun = -INFINITY, lrun = 0.f;          // running max / sum of head g (same in the 4 lanes of a row)
    float o[8][4];                               // o[jn][e]: head g, dim 8 (2 t + e) + jn
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) o[jn][0] = o[jn][1] = o[jn][2] = o[jn][3] = 0.f;
#pragma unroll 1
    for (int sb = 0; sb < nsub; ++sb) {
      const int p0 = sp * split + sb * SUB;
      const int nch = max(0, min(NCH, (Ttot - p0 + 15) >> 4));   // chunks with at least one position
      // ---- S = Q K^T; s[ch][j][e]: head g, position p0 + 16 ch + 8 j + 2 t + e
      float s[NCH][2][2];
#pragma unroll
      for (int q = 0; q < NP / 2; ++q, ++nn) {
        const uint32_t slot = (n0 + nn) % nstg;
        mbar_wait_g(p, &bars[slot], ((n0 + nn) / nstg) & 1u, ph, W_ATTN_BB_KV);
        const unsigned char* st = stages + (size_t)slot * PBYTES;
#pragma unroll
        for (int cc = 0; cc < PS / 16; ++cc) {
          const int ch = q * (PS / 16) + cc;
          if (ch < nch) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              // K row of position 16 cc + 8 j + g inside the piece: dims 8t.. and 32 + 8t..
              const unsigned char* rowp = st + (size_t)(16 * cc + 8 * j + g) * (HD * 2);
              const uint4 x0 = *reinterpret_cast<const uint4*>(rowp + 16 * t), x1 = *reinterpret_cast<const uint4*>(rowp + 64 + 16 * t);
              const uint32_t kw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t a[4] = {qf[2 * i], 0u, qf[2 * i + 1], 0u};
                mma16816(acc, a, kw[2 * i], kw[2 * i + 1]);
              }
              const int pj = p0 + 16 * ch + 8 * j + 2 * t;
              s[ch][j][0] = (pj < Ttot) ? acc[0] * scale : -INFINITY;
              s[ch][j][1] = (pj + 1 < Ttot) ? acc[1] * scale : -INFINITY;
            }
          } else {
            s[ch][0][0] = s[ch][0][1] = s[ch][1][0] = s[ch][1][1] = -INFINITY;
          }
        }
        __syncwarp();                              // every lane has read the stage
        if (lane == 0) issue(nn + nstg);
      }
      // ---- softmax of the sub-block (fp32), folded into the running state: lanes t = 0..3 of a row share a head
      float mx = mrun;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) mx = fmaxf(mx, fmaxf(fmaxf(s[ch][0][0], s[ch][0][1]), fmaxf(s[ch][1][0], s[ch][1][1])));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      // (a sub-block past the end of the context leaves mx = mrun; the first sub-block of a unit always holds a position)
      const float resc = (mrun == -INFINITY) ? 0.f : __expf(mrun - mx);
      float ls = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = __expf(s[ch][j][e] - mx);
            s[ch][j][e] = pv;
            ls += pv;
          }
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      lrun = lrun * resc + ls;
      mrun = mx;
      if (sb > 0) {
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) { o[jn][0] *= resc; o[jn][1] *= resc; }   // (rows g: c0, c1; c2, c3 are rows g + 8 = zero heads)
      }
      // ---- O += P V
#pragma unroll
      for (int q = 0; q < NP / 2; ++q, ++nn) {
        const uint32_t slot = (n0 + nn) % nstg;
        mbar_wait_g(p, &bars[slot], ((n0 + nn) / nstg) & 1u, ph, W_ATTN_BB_KV);
        unsigned char* st = stages + (size_t)slot * PBYTES;
        {
          // rows of the piece that are not cached (the tail of the context) hold whatever the stage held before:
          // zero them (their probabilities are 0, but 0 x garbage must not be NaN)
          const int nv = Ttot - (p0 + q * PS);
          if (nv < PS) {   // (warp-uniform)
            for (int r = max(nv, 0) + (lane >> 3); r < PS; r += 4)
              *reinterpret_cast<uint4*>(st + (size_t)r * (HD * 2) + (lane & 7) * 16) = make_uint4(0, 0, 0, 0);
            __syncwarp();
          }
        }
#pragma unroll
        for (int cc = 0; cc < PS / 16; ++cc) {
          const int ch = q * (PS / 16) + cc;
          if (ch < nch) {
            // A fragments of P: (a0, a2) = positions (2t, 2t+1), (2t+8, 2t+9); bf16 hi + lo parts
            const float h00 = bfround(s[ch][0][0]), h01 = bfround(s[ch][0][1]), h10 = bfround(s[ch][1][0]), h11 = bfround(s[ch][1][1]);
            const uint32_t ahi[4] = {pack_bf16(h00, h01), 0u, pack_bf16(h10, h11), 0u};
            const uint32_t alo[4] = {pack_bf16(s[ch][0][0] - h00, s[ch][0][1] - h01), 0u,
                                     pack_bf16(s[ch][1][0] - h10, s[ch][1][1] - h11), 0u};
            // V rows of positions 2t, 2t+1, 2t+8, 2t+9 of the chunk, dims 8g..8g+7
            const unsigned char* vb = st + (size_t)(16 * cc + 2 * t) * (HD * 2) + 16 * g;
            const uint4 v0 = *reinterpret_cast<const uint4*>(vb), v1 = *reinterpret_cast<const uint4*>(vb + HD * 2);
            const uint4 v2 = *reinterpret_cast<const uint4*>(vb + 8 * HD * 2), v3 = *reinterpret_cast<const uint4*>(vb + 9 * HD * 2);
            const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
            const uint32_t w2[4] = {v2.x, v2.y, v2.z, v2.w}, w3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {   // word wd of the four rows = dims 8g + 2 wd, +1
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int jn = 2 * wd + e;
                const uint32_t sel = e ? 0x7632u : 0x5410u;
                const uint32_t b0 = __byte_perm(w0[wd], w1[wd], sel);
                const uint32_t b1 = __byte_perm(w2[wd], w3[wd], sel);
                mma16816(o[jn], ahi, b0, b1);
                mma16816(o[jn], alo, b0, b1);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) issue(nn + nstg);
      }
    }
    // partial (max, sum, o[64]) of this unit: lane (g < REP, t) holds dims 16t..16t+7 (e = 0) and 16t+8.. (e = 1)
    if (g < REP) {
      float* part = p.attn_part + (((size_t)b * p.bb.heads + kvh * REP + g) * p.nsplit_max + sp) * (HD + 4);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        part[4 + 16 * t + jn] = o[jn][0];
        part[4 + 16 * t + 8 + jn] = o[jn][1];
      }
      if (t == 0) { part[0] = mrun; part[1] = lrun; }
    }
    // publish the partial: the warp's stores are ordered before lane 0's release by the warp barrier (cumulativity);
    // the unit that finds the count complete has, through the same atomic's acquire, every other unit's partial
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.attn_cnt + b * nk + kvh) : "memory");
      last = old == (unsigned)nsplit - 1u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      // last unit of this (sequence, kv-head): merge the splits and write the head outputs
      merge_splits<REP>(p, b, kvh, nsplit, lane, orows);
      if (lane == 0) p.attn_cnt[b * nk + kvh] = 0u;
    }
    __syncwarp();
  }
  if (lane == 0) sm_attcnt()[warp] = n0 + nn;
}

__device__ __forceinline__ bool phase_is_staged(const Phase& P) {   // something is copied by the activation-stream warp
  return P.type == PH_GEMV;
}

}  // namespace

// STOCH only makes the kernel's NAME unique per translation unit: template instantiations have weak linkage, two
// units instantiating csm_batch_kernel<1,4> with different bodies would be merged into one by the linker.
template <int NB, int REP, bool STOCH>
__global__ void __launch_bounds__(CSM_THREADS, 1) csm_batch_kernel(const __grid_constant__ StreamParams p) {
  if (p.stop_flag != nullptr && *p.stop_flag) return;   // generation already ended (set by an earlier launch)
  if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) return;   // an earlier launch timed out

  Lane L;
  L.tid = threadIdx.x;
  L.warp = threadIdx.x >> 5;
  L.lane = threadIdx.x & 31;
  L.c = blockIdx.x;
  L.G = gridDim.x;
  L.slot = L.slot_par = L.ait = L.dpar = L.xpar = 0;
  L.ph = p.phase_begin;
  L.prof = nullptr;

  if (L.tid == 0) {
    for (int s = 0; s < CSM_MAX_SLOTS; ++s) {
      mbar_init(&sm_full()[s], 1);
      mbar_init(&sm_empty()[s], CSM_COMPUTE_WARPS);
      mbar_init(&sm_afull()[s], 1);
      mbar_init(&sm_aempty()[s], CSM_COMPUTE_WARPS);
    }
    mbar_init(sm_dfull(), 1);
    mbar_init(sm_xbar(), CSM_COMPUTE_WARPS);
    for (int i = 0; i < CSM_COMPUTE_WARPS * 4; ++i) mbar_init(&sm_attbar()[i], 1);
    for (int i = 0; i < CSM_COMPUTE_WARPS; ++i) sm_attcnt()[i] = 0u;
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
  cluster_sync_all();   // (CTA pairs: the peer's exchange barrier is initialised before anyone arrives on it)

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
        // (pair phases: the pair's slice holds all K for the pair's rows; this CTA streams its half of the k-tiles)
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2 +
                                   (size_t)g.koff * g.rows * 32;
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
    // ===================== activation stream producer =====================
    // For every GEMV phase whose input is a plain bf16 matrix in global memory: observe the grid barrier that
    // publishes it, then copy the rows with the TMA engine -- whole rows [B, K] onto dfull for K <= 2048, a ring of
    // [B, k-chunk] tiles in lockstep with the weight chunks for the streamed (K = 8192) phases.
    if (L.lane == 0) {
      uint32_t ait = 0;
      unsigned char* actreg = sm_act(p);
      const uint32_t nsa = (uint32_t)p.a_slots;
#pragma unroll 1
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (!phase_is_staged(P)) continue;
        const bf16* actp = P.act;
        const int act_stride = P.act_stride;
        if (P.act_mode != ACT_STREAM) {
          if (p.use_barrier && ph > p.phase_begin) grid_wait(p, (unsigned)(ph - p.phase_begin) * L.G, ph);
          fence_proxy_async();
          // rows are K + 8 elements apart in global memory as in shared memory: the [B, K+8] block is ONE copy; the
          // norm weights of the phase ride on the same barrier (gathered phases: the weights only)
          const uint32_t bytes = P.act_mode == ACT_GATHER ? 0u : (uint32_t)p.B * (uint32_t)(P.K + 8) * 2u;
          const uint32_t nbytes = P.norm_w != nullptr ? (uint32_t)P.K * 2u : 0u;
          mbar_expect_tx(sm_dfull(), bytes + nbytes);
          if (bytes) bulk_g2s(actreg, actp, bytes, sm_dfull());
          if (nbytes) bulk_g2s(actreg + p.normw_off, P.norm_w, nbytes, sm_dfull());
          (void)act_stride;
          continue;
        }
        const Geom g = csm_geom(P, L.c);
        if (g.nchunks == 0) continue;
        if (p.use_barrier && ph > p.phase_begin) grid_wait(p, (unsigned)(ph - p.phase_begin) * L.G, ph);
        fence_proxy_async();
        const int astride_b = (g.tpc * 16 + 8) * 2;
#pragma unroll 1
        for (int ch = 0; ch < g.nchunks; ++ch, ++ait) {
          // the producer (gate/up phase) wrote the tiled layout [k-tile][m_alloc][len+8]: tile `ch` is ONE copy
          const uint32_t s = ait % nsa, round = ait / nsa;
          if (round > 0) mbar_wait_g(p, &sm_aempty()[s], (round - 1u) & 1u, ph, W_AEMPTY);
          const uint32_t bytes = (uint32_t)p.B * (uint32_t)astride_b;
          mbar_expect_tx(&sm_afull()[s], bytes);
          const int tile = g.koff / g.tpc + ch;   // (pair phases: the tiles of this CTA's K half)
          bulk_g2s(actreg + (size_t)s * p.a_slot_bytes, reinterpret_cast<const unsigned char*>(actp) + (size_t)tile * p.m_alloc * astride_b,
                   bytes, &sm_afull()[s]);
        }
      }
    }
    return;
  }
  if (L.warp == CSM_COMPUTE_WARPS + 2) {
    // ===================== L2 prefetcher =====================
    // Issues HBM->L2 prefetches for this CTA's weight stream up to l2_ahead_bytes beyond what the ring has requested,
    // so that DRAM never idles while the ring is full and the compute warps are inside a latency-bound stretch
    // (barrier, staging, epilogue, attention).  Also: the norm weights of upcoming phases (one CTA each) and the K/V
    // blocks of this CTA's first backbone attention units.
    if (L.lane == 0 && p.l2_ahead_bytes > 0) {
      uint32_t pf = 0;
#pragma unroll 1
      for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
        const Phase& P = p.phases[ph];
        if (P.type == PH_ATTN_BB) {
          const int Ttot = p.pos + 1, nk = p.bb.kv;
          const int split = CSM_ATT_SPLIT_MMA * p.att_nsub;
          const int nsplit = (Ttot + split - 1) / split;
          const int nunits = p.B * nk * nsplit;
          int done = 0;
          // (one warp per unit: the first round of this CTA's eight warps are units c, G + c, ...; later rounds are
          // prefetched by the warps themselves)
          for (int unit = L.c; unit < nunits && done * p.att_nsub < p.att_pf_units; unit += L.G, ++done) {
            const int sp = unit % nsplit, kvh = (unit / nsplit) % nk, b = unit / (nsplit * nk);
            const size_t off = ((((size_t)P.layer * p.Bmax + b) * nk + kvh) * (size_t)p.Tcap + (size_t)sp * split) * 64;
            const int npos = min(split, p.pos - sp * split);   // positions cached before this frame
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
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.w) + (size_t)g.row0 * P.K * 2 +
                                   (size_t)g.koff * g.rows * 32;
        const uint32_t total = (uint32_t)g.rows * (uint32_t)g.ntiles * 32u;
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
    const unsigned bar_target = (p.use_barrier && ph > p.phase_begin) ? (unsigned)(ph - p.phase_begin) * L.G : 0u;
    if (prof) prof[0] = clock64();
    uint4 nxt = make_uint4(0, 0, 0, 0);
    const bool fetch = L.warp == CSM_COMPUTE_WARPS - 1 && L.lane < 16 && ph + 1 < p.phase_end;
    if (fetch) nxt = __ldg(reinterpret_cast<const uint4*>(p.phases + ph + 1) + L.lane);
    const int type = P.type;
    if (type != PH_GEMV && bar_target) {   // (GEMV phases observe the barrier through the activation-stream warp)
      if (L.tid == 0) grid_wait(p, bar_target, ph);
      compute_sync();
    }
    if (prof) prof[1] = clock64();       // phase body starts
    if (type == PH_GEMV) gemv_phase<NB>(p, P, L, bar_target);
    else if (type == PH_ATTN_DEC) attn_dec_phase(p, P, L);
    else if (type == PH_ATTN_BB) {
      // K/V pieces of 32 / 64 / 128 positions (4 / 8 / 16 KB bulk copies): the TMA engine spends ~0.25-0.35 us per
      // bulk copy whatever its size (measured), so the largest piece the per-warp share of the activation region holds
      if (p.att_ps == 128) attn_bb_phase_tma<REP, 128>(p, P.layer, ph);
      else if (p.att_ps == 64) attn_bb_phase_tma<REP, 64>(p, P.layer, ph);
      else attn_bb_phase_tma<REP, 32>(p, P.layer, ph);
    }
    else if (type == PH_EMBED) embed_phase(p, ph);
    else finish_phase(p, P.res_ph);
    if (prof) prof[2] = clock64();       // this thread's share of the body done
    if (ph + 1 < p.phase_end) {
      if (fetch) reinterpret_cast<uint4*>(&sm_desc()[(ph + 1) & 1])[L.lane] = nxt;
      compute_sync();                    // next descriptor visible; every thread's stores of this phase are issued
      if (p.use_barrier && L.tid == 0) {
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
typedef void (*BatchKernel)(const StreamParams);

template <int NB>
static BatchKernel pick_rep(int rep) {
  switch (rep) {
    case 1: return csm_batch_kernel<NB, 1, CSM_BUILD_STOCH != 0>;
    case 2: return csm_batch_kernel<NB, 2, CSM_BUILD_STOCH != 0>;
    default: return csm_batch_kernel<NB, 4, CSM_BUILD_STOCH != 0>;
  }
}

#if CSM_BUILD_STOCH
#define CSM_LAUNCHER csm_launch_batch_stoch
#else
#define CSM_LAUNCHER csm_launch_batch
#endif

extern "C" cudaError_t CSM_LAUNCHER(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative,
                                    int cluster) {
  const int rep = p->bb.heads / p->bb.kv;
  const int nb = (p->B + 7) / 8;
  BatchKernel k = nb <= 1 ? pick_rep<1>(rep) : (nb <= 2 ? pick_rep<2>(rep) : pick_rep<4>(rep));
  cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CSM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (cooperative) {
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    ++na;
  }
  if (cluster > 1) {   // CTA pairs (streamed phases split K inside a pair and exchange partial sums through DSMEM)
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)cluster;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, k, *p);
}
