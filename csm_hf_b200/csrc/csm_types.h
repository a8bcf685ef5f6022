// Shared host/device records of the frame engine (phase table, kernel parameters).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define CSM_COMPUTE_WARPS 8
#define CSM_COMPUTE_THREADS (CSM_COMPUTE_WARPS * 32)
#define CSM_THREADS (CSM_COMPUTE_THREADS + 64)  // + weight-stream warp + activation-stream warp
#define CSM_MAX_SLOTS 8
#define CSM_NQ 32            // codebooks per frame (modeling_csm.py:66)
#define CSM_DEC_POS 32       // decoder positions per frame: last_h + 31 codebook embeddings
#define CSM_ATT_SPLIT 128    // backbone positions per split-KV unit
#define CSM_MAX_NT 16        // at most 128 weight rows per CTA per matrix

enum PhaseType { PH_EMBED = 0, PH_GEMV = 1, PH_ATTN_BB = 2, PH_ATTN_DEC = 3 };
enum ActMode { ACT_NORM = 0, ACT_PLAIN = 1, ACT_GATHER = 2, ACT_STREAM = 3 };
enum EpiMode { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_QKV = 3, EPI_HEAD = 4 };

// One step of the per-frame program.  A frame is ~800 of these executed in order by every
// CTA of one persistent launch, with a grid-wide barrier between consecutive phases.
struct Phase {
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
  const bf16* w;        // packed weights, see csm_pack.cu
  const bf16* act;      // activation rows (ACT_GATHER: embedding table)
  const bf16* norm_w;   // ACT_NORM weight
  bf16* out;            // STORE/RESID/SWIGLU destination; QKV: q buffer; HEAD: logits (nullable)
  bf16* norm_out;       // optional copy of the normalised rows (last_hidden_state)
};

struct StackDims {
  int H, I, L, heads, kv, hd;
  float eps;
  float scale;      // hd^-0.5
};

struct StreamParams {
  const Phase* phases;
  int phase_begin, phase_end;
  int use_barrier;              // 0 in stepped mode (one launch per phase)
  int B;                        // sequences in this call
  int pos;                      // backbone position of the token being processed (= cached length)
  int Bmax, Tcap;
  int V, text_vocab;
  StackDims bb, dec;
  unsigned int* bar_counter;    // grid barrier, zeroed before every launch
  // KV caches: [L][Bmax][kv][cap][hd]
  bf16 *kc_bb, *vc_bb, *kc_dec, *vc_dec;
  const bf16 *cos_bb, *sin_bb, *cos_dec, *sin_dec;   // [n_pos][hd/2]
  bf16 *q_bb, *q_dec;           // [Bmax][heads*hd]
  bf16 *attn_bb, *attn_dec;     // attention outputs [Bmax][heads*hd]
  float* attn_part;             // [Bmax][heads_bb][nsplit_max][hd+2]
  int nsplit_max;
  unsigned int* attn_cnt;       // [Bmax][kv_bb]
  float2* head_part;            // [grid][Bmax] (value, index)
  unsigned int* head_cnt;
  int* samples;                 // [Bmax][32] argmax of every head
  int* fed;                     // [Bmax][32] tokens fed onward (== samples unless forced)
  int forced;
  const long long* ids;         // PH_EMBED: [B][33] or null (use `fed`, audio slots only)
  const int* mask;              // [B][33] or null
  const bf16 *text_emb, *audio_emb;
  bf16* h_bb;                   // backbone residual stream [Bmax][Hb]
  long long* out_frames;        // [B][out_stride] int64 or null; this frame at out_off
  long long out_stride, out_off;
  int* stop_flag;
  int* n_frames;
  int stop_on_zeros;
  // shared-memory plan
  int m_alloc;                  // activation rows rounded up to 8
  int slot_bytes, n_slots;
  int act_region_bytes, red_bytes;
  int stream_tpc_max;
  unsigned long long* prof;     // debug: CTA 0 writes clock64 at [2*ph] phase start, [2*ph+1] phase end
};

// Row split and chunking of one weight matrix for CTA `c` of `G`.
struct Geom {
  int row0, rows;     // packed rows owned by this CTA
  int ntiles;         // K/16
  int tpc;            // k16-tiles per ring slot
  int nchunks;
  int nt, ns, ks;     // n8-tiles, n-split and k-split over the 8 compute warps
};

__host__ __device__ inline Geom csm_geom(int N, int K, int gran, int G, int c, int slot_bytes, int tpc_cap) {
  Geom g;
  int U = N / gran, q = U / G, r = U % G;
  int start = c * q + (c < r ? c : r);
  int cnt = q + (c < r ? 1 : 0);
  g.row0 = start * gran;
  g.rows = cnt * gran;
  g.ntiles = K / 16;
  if (g.rows == 0) {
    g.tpc = 0; g.nchunks = 0; g.nt = 0; g.ns = 1; g.ks = 8;
    return g;
  }
  int tpc = slot_bytes / (g.rows * 32);
  if (tpc > g.ntiles) tpc = g.ntiles;
  if (tpc > tpc_cap) tpc = tpc_cap;
  if (tpc < 1) tpc = 1;
  g.tpc = tpc;
  g.nchunks = (g.ntiles + tpc - 1) / tpc;
  g.nt = (g.rows + 7) / 8;
  g.ns = g.nt >= 8 ? 8 : (g.nt >= 4 ? 4 : (g.nt >= 2 ? 2 : 1));
  g.ks = 8 / g.ns;
  return g;
}
