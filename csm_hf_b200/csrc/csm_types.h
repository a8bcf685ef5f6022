// Shared host/device records of the frame engine (phase table, kernel parameters).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define CSM_COMPUTE_WARPS 8
#define CSM_COMPUTE_THREADS (CSM_COMPUTE_WARPS * 32)
#define CSM_THREADS (CSM_COMPUTE_THREADS + 96)  // + weight-stream warp + activation-stream warp + L2-prefetch warp
#define CSM_MAX_SLOTS 8
#define CSM_NQ 32            // codebooks per frame (modeling_csm.py:66)
#define CSM_DEC_POS 32       // decoder positions per frame: last_h + 31 codebook embeddings
#define CSM_ATT_SPLIT 128    // backbone positions per split-KV unit (SMALL kernels: one CTA per unit, 8 warps x 16)
#define CSM_ATT_SPLIT_MMA 128 // ... of the general kernels (one warp per unit on tensor cores; the scores of a unit stay in registers)
#define CSM_MAX_ROWS 256     // at most 16 m-tiles of 16 weight rows per CTA per matrix
#define CSM_SM_HDR_BYTES 4096   // mbarriers, phase-descriptor slots, 2 KB scratch, token slots

enum PhaseType { PH_EMBED = 0, PH_GEMV = 1, PH_ATTN_BB = 2, PH_ATTN_DEC = 3, PH_FINISH = 4 };
enum ActMode { ACT_NORM = 0, ACT_PLAIN = 1, ACT_GATHER = 2, ACT_STREAM = 3, ACT_ATTN = 4 };
enum EpiMode { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_QKV = 3, EPI_HEAD = 4 };

// Phase::flags
#define CSM_PF_BAR_IN 1      // a grid barrier precedes this phase (its input is plain memory read by the TMA engine)
#define CSM_PF_BAR_OUT 2     // arrive on the grid barrier after this phase (the next phase has BAR_IN)
#define CSM_PF_OUT_PLAIN 4   // EPI_SWIGLU writes plain bf16 (consumer streams it with bulk copies) instead of tagged words
#define CSM_PF_KV_COPY 8     // decoder qkv phase, fused attention: copy this layer's cached K/V rows to shared memory
                             // at the end of the phase (the o_proj phase that follows computes the attention from them)

// One step of the per-frame program.  A frame is ~800 of these executed in order by every
// CTA of one persistent launch.  Phases are chained by DATAFLOW, not by barriers: every value that
// crosses CTAs travels as a 32-bit "tagged word" (bf16 payload | 16-bit tag of the producing phase),
// written with one relaxed store and polled by the consumer, so the hand-over costs one trip
// through L2 and needs no fence (payload and flag are the same atomic word).
// 256 bytes: the kernel copies the next descriptor into shared memory while a phase runs.
// Row split / chunking of one weight matrix for one class of CTAs, computed on the host (plan_smem) so that
// the kernel's per-phase prologue is a handful of loads instead of divisions and loops.
// Class 0: CTAs c < r (q+1 granules); class 1: the others (q granules).
struct GeoC {
  int rows;       // packed weight rows owned by a CTA of this class
  int tpc, nch;   // k16-tiles per ring chunk, chunks per CTA
  int mtiles;     // 16-row m-tiles
  int ksl;        // log2 of the split-K ways over the 8 compute warps (ks = 1<<ksl, ns = 8>>ksl m-tile groups)
  int rows_pad;   // mtiles*16+4: row stride of the split-K partial sums in shared memory
  int upc, ush;   // epilogue granules per CTA, log2 of the next power of two
};

struct __align__(16) Phase {
  int type;         // PhaseType
  int act_mode;     // ActMode   (PH_GEMV)
  int epi;          // EpiMode   (PH_GEMV)
  int gran;         // rows that must stay in one CTA (2 for RoPE pairs / gate-up pairs)
  int N, K;         // packed weight rows, reduction length
  int stack;        // 0 backbone, 1 decoder
  int layer;
  int cb;           // codebook: head index (EPI_HEAD) or embedding slice (ACT_GATHER)
  int dec_pos;      // decoder position of this pass (RoPE / attention length)
  int act_stride;   // elements between activation rows
  int out_stride;
  // row split over the grid, computed on the host for the launch grid G: CTAs c < r own q+1
  // granules, the others q
  int q, r;
  int src_ph;       // phase that produced `act` (tag to poll for); ATTN phases: the qkv phase
  int res_ph;       // EPI_RESID: phase that last wrote the residual stream; ACT_GATHER / PH_FINISH: the head phase
  int flags;        // CSM_PF_*
  int bar_idx;      // number of BAR_IN phases in table[0..this]
  int gsh;          // log2(K/4) when K/4 is a power of two (staging index math by shifts), else -1
  int tile_sh;      // general kernels, EPI_SWIGLU feeding a streamed down_proj: log2 of the k-tile length of the output's
                    // tiled layout [k-tile][m_alloc][len+8] (each [B, len+8] tile is one contiguous bulk copy); 0: row-major
  const bf16* w;        // packed weights, see csm_pack.cu
  const bf16* act;      // activation rows: tagged words (uint32) unless ACT_GATHER (embedding table, bf16) / ACT_STREAM (bf16)
  const bf16* norm_w;   // ACT_NORM weight
  bf16* out;            // STORE/RESID/SWIGLU/QKV destination (tagged words); HEAD: logits (plain bf16, nullable)
  bf16* norm_out;       // optional copy of the normalised rows (last_hidden_state)
  long long pad1_;
  GeoC geo[2];
  int pair;         // general kernels, streamed (K = 8192) phases: the two CTAs of a cluster own the rows of both and split
                    // K in halves (q, r, geo are per PAIR; weights packed per pair); partial sums meet through DSMEM
  int pad2_[15];
};
static_assert(sizeof(Phase) == 256, "Phase must be 256 bytes");

struct StackDims {
  int H, I, L, heads, kv, hd;
  int hdl;          // log2(hd) (head_dim is 64 or 128)
  float eps;
  float scale;      // hd^-0.5
};

struct StreamParams {
  const Phase* phases;
  int phase_begin, phase_end;
  int use_barrier;              // 0 in stepped mode (one launch per phase)
  unsigned int tagbase;         // tag of phase ph = (tagbase + ph) & 0xffff, never 0 (0 = "never written")
  int repl;                     // copies of every tagged vector (consumer CTA c polls copy c % repl): spreads the pollers
                                // over repl x more L2 lines, so no line is hammered by all 148 CTAs at once
  int evict_first;              // weight bulk copies carry an L2 evict-first hint
  int small;                    // engine built for <= 2 sequences (fused decoder attention): selects the SMALL kernels
  int l2_ahead_bytes;           // how far (bytes of this CTA's weight stream) the L2 prefetcher runs ahead of the ring
  int B;                        // sequences in this call
  int pos;                      // backbone position of the token being processed (= cached length)
  int Bmax, Tcap;
  int V, text_vocab;
  StackDims bb, dec;
  unsigned int* bar_counter;    // grid barrier, zeroed before every launch
  // KV caches: [L][Bmax][kv][cap][hd]
  bf16 *kc_bb, *vc_bb, *kc_dec, *vc_dec;
  const bf16 *cos_bb, *sin_bb, *cos_dec, *sin_dec;   // [n_pos][hd/2]
  uint32_t *q_bb, *q_dec;       // tagged [Bmax][(heads + 2 kv)*hd]: q | k | v of the position being processed
  uint32_t *attn_bb, *attn_dec; // tagged attention outputs [Bmax][heads*hd]
  float* attn_part;             // [Bmax][heads_bb][nsplit_max][hd+2]
  int nsplit_max;
  unsigned int* attn_cnt;       // [Bmax][kv_bb]
  unsigned long long* cand;     // [grid][Bmax] per-CTA candidate of the last head phase: tag<<32 | index<<16 | bf16 logit
  int* samples;                 // [Bmax][32] argmax of every head
  int* fed;                     // [Bmax][32] tokens fed onward (== samples unless forced)
  int forced;
  const long long* ids;         // PH_EMBED: [B][33] or null (use `fed`, audio slots only)
  const int* mask;              // [B][33] or null
  const bf16 *text_emb, *audio_emb;
  uint32_t* h_bb;               // backbone residual stream, tagged [Bmax][Hb]
  long long* out_frames;        // [B][out_stride] int64 or null; this frame at out_off
  long long out_stride, out_off;
  int* stop_flag;
  int* n_frames;
  int stop_on_zeros;
  // shared-memory plan (host: plan_smem)
  int m_alloc;                  // activation rows rounded up to 8
  int slot_bytes, n_slots;
  int rope_bytes, act_region_bytes, red_bytes;
  int a_slots, a_slot_bytes;    // general kernels: ring of [B, k-chunk] activation tiles inside the activation region (K = 8192 phases)
  int normw_off;                // general kernels: byte offset, inside the activation region, of the staged norm weights (4 KB)
  int att_nsub;                 // general kernels: 128-position sub-blocks per backbone-attention unit (about one unit per warp)
  int att_ps;                   // general kernels: positions per K/V piece of the backbone attention (32, 64 or 128)
  int att_pf_units;             // general kernels: backbone-attention units per CTA whose K/V the L2 prefetcher pulls in ahead
  int att_stages;               // general kernels: 4 KB stages per warp of the backbone-attention K/V ring (2..4)
  int xbuf_off;                 // general kernels: byte offset inside the reduction region of the [m_alloc][16] fp32 buffer the
                                // peer CTA of a pair writes its K-half partial sums into
  int hpad;                     // general kernels: elements of padding after every row of the inter-phase vectors (8): a
                                // [B, K+8] block in global memory is the shared-memory image, copied with ONE bulk copy
  unsigned long long* prof;     // debug: clock64 stamps [2 CTAs (first,last)][n_phases_total][4], see csm_stream.cu
  int n_phases_total;
  // stochastic top-k sampling (topk <= 1: greedy).  lgt: tagged logits of the last head phase [Bmax][lgt_stride]
  int topk;
  float inv_temp;
  unsigned long long rng_seed;
  unsigned int rng_frame;       // frame counter of this context (part of the noise key)
  int seq_base;                 // global index of sequence 0 (batch sharding: ranks draw independent noise)
  uint32_t* lgt;
  int lgt_stride;               // V rounded up to a multiple of 4 words
  int* abort_flag;              // [0] != 0: a wait inside a frame kernel timed out (or an earlier launch did): every wait
                                // gives up, later launches return at once; [1..7] = (cta, phase, wait id, detail...)
  int* progress;                // debug (CSM_DEBUG_PROGRESS=1): [grid][4] = phase of the compute warps, step inside it,
                                // phase of the weight stream, phase of the L2 prefetcher -- read by a watchdog on a hang
};

// One dense projection of the context prefill on the tcgen05 path (csm_gemm.cu): C[R,N] = A[R,K] * W[N,K]^T + fused tail.
struct GemmParams {
  int R, N, K;
  int epi;                      // EPI_STORE / EPI_RESID / EPI_SWIGLU / EPI_QKV
  bf16* C;                      // STORE / RESID: [R, ldc]; SWIGLU: [R, ldc] with N/2 columns; QKV: rotated q rows [R, ldc]
  int ldc;
  // EPI_QKV: row r = (sequence b0 + r / S, position pos0 + r % S); k and v go to the cache [L][Bmax][kv][Tcap][64]
  int S, pos0, b0, heads, kv, layer, Bmax, Tcap;
  bf16 *kc, *vc;
  const bf16 *cos_t, *sin_t;    // [n_pos][32]
  // operand layouts (EPI_STORE / EPI_RESID only): 0 = K-major (rows of the matrix are contiguous along the contraction),
  // 1 = MN-major (the matrix is stored [K, rows]: contiguous along its M / N dimension) -- dX = dY W reads W [N_out, K_in]
  // as an MN-major B operand, dW = dY^T X reads dY and X as MN-major operands: no transposed copies
  int a_mn, b_mn;
};

// Row split and chunking of one weight matrix for one CTA.
struct Geom {
  int row0, rows;     // packed rows owned by this CTA (pair phases: by the pair)
  int ntiles;         // k16-tiles this CTA reduces over: K/16, or K/32 in a pair phase
  int tpc;            // k16-tiles per ring slot
  int nchunks;
  int koff;           // first k16-tile of this CTA (pair phases: rank * ntiles)
};

__host__ __device__ inline Geom csm_geom(const Phase& P, int c) {
  Geom g;
  const int cc = P.pair ? (c >> 1) : c;
  const int hi = cc < P.r;
  const GeoC& gc = P.geo[hi ? 0 : 1];
  g.row0 = (cc * P.q + (hi ? cc : P.r)) * P.gran;
  g.rows = gc.rows;
  g.ntiles = P.pair ? (P.K >> 5) : (P.K >> 4);
  g.tpc = gc.tpc;
  g.nchunks = gc.nch;
  g.koff = P.pair ? (c & 1) * g.ntiles : 0;
  return g;
}
