// C ABI of libcsm_b200.so (include/csm_b200.h): context creation (weight packing, workspace,
// phase table), prefill orchestration, per-frame persistent launches, generate loop.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/csm_b200.h"
#include "csm_types.h"

// ---- kernels / launchers defined in the other translation units
struct PackSrc {
  const bf16* ptr[3];
  int rows[3];
  long long row_stride[3];
  long long col_stride[3];
};
extern "C" {
cudaError_t csm_launch_stream_small(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative);
cudaError_t csm_launch_stream_small_stoch(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative);
cudaError_t csm_launch_batch(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative, int cluster);
cudaError_t csm_launch_batch_stoch(const StreamParams* p, int grid, size_t smem, cudaStream_t stream, int cooperative,
                                   int cluster);
cudaError_t csm_pack_launch(const PackSrc* src, const int* row_map_dev, int N, int K, int gran, int G, bf16* dst,
                            cudaStream_t stream);
cudaError_t csm_gemm_launch(const void* map_a, const void* map_w, const GemmParams* p, int sms, cudaStream_t st);
int csm_tmap_2d(void* out, const void* base, long long rows, int K, long long pitch, int box_rows);
int csm_gemm_box_rows_a();
int csm_gemm_box_rows_w();
cudaError_t csm_interleave_rows_launch(const bf16* a, const bf16* b, int rows, int K, bf16* dst, cudaStream_t st);
cudaError_t csm_embed_sum_launch(const long long* ids, const int* mask, int default_mask, const bf16* audio_emb,
                                 const bf16* text_emb, int V, int H, bf16* out, int rows, cudaStream_t st);
cudaError_t csm_rmsnorm_rows_launch(const bf16* x, const bf16* w, float eps, int H, bf16* y, int rows, cudaStream_t st);
cudaError_t csm_take_last_rows_launch(const bf16* h, int S, int H, uint32_t* dst, int b0, int nseq, uint32_t tag,
                                      int plain, cudaStream_t st);
cudaError_t csm_sample_rows_launch(const bf16* logits, int rows, int V, int topk, float inv_temp, unsigned long long seed,
                                   long long* out, cudaStream_t st);
cudaError_t csm_untag_rows_launch(const uint32_t* src, long long src_stride, int cols, int rows, bf16* dst, cudaStream_t st);
cudaError_t csm_i64_to_i32_launch(const long long* src, int* dst, int n, cudaStream_t st);
cudaError_t csm_i32_to_i64_launch(const int* src, long long* dst, int n, cudaStream_t st);
cudaError_t csm_flash_tc_prefill_launch(const bf16* qkv, int S, int b0, int nseq, int heads, int kv, const bf16* kc,
                                        const bf16* vc, int layer, int layers, int Bmax, int Tcap, float scale,
                                        const unsigned char* valid, bf16* out, cudaStream_t st);
cudaError_t csm_flash_prefill_launch(const bf16* qkv, int S, int pos0, int b0, int nseq, int heads, int kv,
                                     const bf16* kc, const bf16* vc, int layer, int Bmax, int Tcap, float scale,
                                     const unsigned char* valid, bf16* out, cudaStream_t st);
cudaError_t csm_frame_valid_launch(const int* mask, int rows, unsigned char* valid, int* any_pad, cudaStream_t st);
}

namespace {

struct LayerW {
  // natural-layout [out, in] copies for the prefill GEMMs (backbone only): q|k|v rows concatenated, gate/up rows
  // interleaved (gate_j, up_j) so that SwiGLU is a tail of the GEMM tile; TMA tensor maps of each (csm_gemm.cu)
  bf16 *qkv = nullptr, *o = nullptr, *gu = nullptr, *down = nullptr;
  CUtensorMap tm_qkv, tm_o, tm_gu, tm_down;
  bf16 *ln1 = nullptr, *ln2 = nullptr;
  // packed for the frame engine
  bf16 *p_qkv = nullptr, *p_o = nullptr, *p_gu = nullptr, *p_down = nullptr;
};

struct Stack {
  StackDims d;
  std::vector<LayerW> layers;
  bf16* norm = nullptr;
  bf16 *cos_t = nullptr, *sin_t = nullptr;
  int n_pos = 0;
};

}  // namespace

struct CsmCtx {
  bool flash_tc = getenv("CSM_FLASH_MMA") == nullptr;   // prefill attention on tcgen05 (csm_flash_tc.cu)
  int device = 0, sms = 0, G = 0;
  int Bmax = 0, Tcap = 0;
  int V = 0, text_vocab = 0;
  Stack bb, dec;
  bf16 *text_emb = nullptr, *audio_emb = nullptr;
  bf16 *p_proj = nullptr, *p_c0 = nullptr;
  bf16* proj_table = nullptr;   // projection(audio_embeddings) [32*V][Hd]: the decoder input of positions 1..31 is a row gather
  std::vector<bf16*> p_heads;
  // workspace
  bf16 *kc_bb = nullptr, *vc_bb = nullptr, *kc_dec = nullptr, *vc_dec = nullptr;
  // inter-phase vectors: tagged words (csm_common.cuh), one uint32 per element
  uint32_t *h_bb = nullptr, *h_dec = nullptr, *q_bb = nullptr, *q_dec = nullptr, *attn_bb = nullptr, *attn_dec = nullptr;
  uint32_t *mlp_bb = nullptr, *mlp_dec = nullptr;   // tagged words, or plain bf16 when the down_proj input is TMA-streamed
  bf16 *last_h = nullptr, *c0_logits = nullptr, *cb_logits = nullptr;
  std::vector<std::pair<void*, size_t>> tagged;     // buffers to clear when the 16-bit tag epoch wraps
  unsigned int tagbase = 1;                          // tag of phase ph of the next frame = tagbase + ph (1..65535)
  int l2_ahead = 256 * 1024;
  int repl = 1;                                      // copies of every tagged vector (StreamParams::repl)
  int evict_first = 1;
  float* attn_part = nullptr;
  int nsplit_max = 0;
  unsigned int *attn_cnt = nullptr, *bar_counter = nullptr;
  unsigned long long* cand = nullptr;
  int *samples = nullptr, *fed = nullptr, *stop_flag = nullptr, *n_frames = nullptr;
  unsigned long long* prof = nullptr;
  int prof_on = 0;
  int* abort_flag = nullptr; // [8], see StreamParams::abort_flag
  int* progress = nullptr;   // [sms][4], see StreamParams::progress
  int progress_on = 0;
  int bar_all = 0;           // grid barrier between all phases (CSM_BAR_ALL=1)
  // sampling (csm_set_sampling); topk <= 1 = greedy
  int topk = 1;
  float inv_temp = 1.f;
  unsigned long long rng_seed = 0;
  unsigned int rng_frame = 0;
  int seq_base = 0;
  uint32_t* lgt = nullptr;
  int lgt_stride = 0;
  // prefill workspace (lazy)
  int pf_rows = 0;
  bf16 *pf_h = nullptr, *pf_hn = nullptr, *pf_qkv = nullptr, *pf_attn = nullptr, *pf_act = nullptr;
  unsigned char* pf_valid = nullptr;   // [rows] frame-valid bytes of a padded prefill
  // host staging for csm_generate_host
  long long* st_ids = nullptr;
  int* st_mask = nullptr;
  long long* st_frames = nullptr;
  size_t st_ids_n = 0, st_frames_n = 0;
  // phase table
  std::vector<Phase> table;
  Phase* d_table = nullptr;
  int ph_head_c0 = 0;   // first phase of the "decoder part" of a frame (final norm + c0 head)
  int cache_len = 0;
  int stepped = 0;
  int fuse_attn = 0;     // decoder attention computed inside the o_proj phase (max_batch <= 2)
  int direct_mlp = 0;    // MLP activations staged whole in shared memory (max_batch <= 4)
  // shared-memory plan of the frame kernel (fixed at create time for max_batch)
  int m_alloc = 0, slot_bytes = 0, n_slots = 0, rope_bytes = 0, act_region = 0, red_bytes = 0, stream_tpc_max = 0;
  int a_slots = 2, a_slot_bytes = 0;   // activation-tile ring of the K = 8192 phases (general kernels)
  int normw_off = 0;
  int pair = 0;                        // general kernels: CTA pairs (clusters of 2) split K of the streamed phases
  int xbuf_off = 0;
  int mt2 = 0;                         // CSM_MT2=1: two m-tiles per warp wherever a CTA owns more than one (experiment)
  size_t smem_total = 0;
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_done = nullptr;
  double wait_budget_s = 120.0;   // host-side watchdog of csm_frames_done / csm_generate_host (CSM_WAIT_BUDGET_S)
  int ev_frames = 0;
  std::vector<void*> allocs;
  std::string err;
};

namespace {

int fail(CsmCtx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, CSM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                              \
  } while (0)

template <typename T>
int dalloc(CsmCtx* ctx, T** p, size_t n) {
  void* q = nullptr;
  size_t bytes = n * sizeof(T);
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc(&q, bytes));
  ctx->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}
#define DA(p, n)                                  \
  do {                                            \
    int r_ = dalloc(ctx, &(p), (size_t)(n));      \
    if (r_) return r_;                            \
  } while (0)

int copy_weight(CsmCtx* ctx, bf16** dst, const void* src, size_t n, cudaStream_t st) {
  DA(*dst, n);
  CK(cudaMemcpyAsync(*dst, src, n * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pack_matrix(CsmCtx* ctx, bf16** dst, const PackSrc& src, const std::vector<int>* row_map, int N, int K, int gran,
                cudaStream_t st, int G = 0) {
  if (G == 0) G = ctx->G;   // (pair phases: rows are split over G/2 CTA pairs)
  DA(*dst, (size_t)N * K);
  int* d_map = nullptr;
  if (row_map) {
    CK(cudaMalloc(&d_map, row_map->size() * sizeof(int)));
    CK(cudaMemcpyAsync(d_map, row_map->data(), row_map->size() * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  CK(csm_pack_launch(&src, d_map, N, K, gran, G, *dst, st));
  if (d_map) {
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_map));
  }
  // every CTA may own at most CSM_MAX_NT*8 rows of a matrix
  int U = N / gran, per = (U + G - 1) / G * gran;
  if (per > CSM_MAX_ROWS)
    return fail(ctx, CSM_EINVAL, "matrix with %d rows needs %d rows per CTA on a %d-CTA grid (max %d)", N, per, G,
                CSM_MAX_ROWS);
  return 0;
}

PackSrc one_src(const bf16* p, int rows, long long rs, long long cs) {
  PackSrc s;
  memset(&s, 0, sizeof s);
  s.ptr[0] = p; s.rows[0] = rows; s.row_stride[0] = rs; s.col_stride[0] = cs;
  s.ptr[1] = s.ptr[2] = p; s.rows[1] = s.rows[2] = 0; s.row_stride[1] = s.row_stride[2] = rs;
  s.col_stride[1] = s.col_stride[2] = cs;
  return s;
}

int build_stack(CsmCtx* ctx, Stack& S, const CsmLlamaShape& sh, const void* const* lw, const void* norm,
                bool keep_natural, cudaStream_t st) {
  StackDims& d = S.d;
  d.H = sh.hidden; d.I = sh.inter; d.L = sh.layers; d.heads = sh.heads; d.kv = sh.kv_heads;
  d.hd = sh.hidden / sh.heads; d.eps = sh.eps; d.scale = 1.0f / sqrtf((float)d.hd);
  d.hdl = d.hd == 128 ? 7 : 6;
  const int H = d.H, I = d.I, hd = d.hd, half = hd / 2;
  const int nq = d.heads * hd, nkv = d.kv * hd;
  S.layers.resize(d.L);
  // row maps
  std::vector<int> qkv_map((size_t)nq + 2 * nkv), gu_map((size_t)2 * I);
  {
    size_t n = 0;
    for (int h = 0; h < d.heads; ++h)
      for (int i = 0; i < half; ++i) { qkv_map[n++] = h * hd + i; qkv_map[n++] = h * hd + i + half; }
    for (int h = 0; h < d.kv; ++h)
      for (int i = 0; i < half; ++i) { qkv_map[n++] = nq + h * hd + i; qkv_map[n++] = nq + h * hd + i + half; }
    for (int f = 0; f < nkv; ++f) qkv_map[n++] = nq + nkv + f;
    for (int j = 0; j < I; ++j) { gu_map[2 * j] = j; gu_map[2 * j + 1] = I + j; }
  }
  for (int l = 0; l < d.L; ++l) {
    const void* const* w = lw + (size_t)l * CSM_W_PER_LAYER;
    LayerW& L = S.layers[l];
    int r;
    if ((r = copy_weight(ctx, &L.ln1, w[CSM_W_LN1], H, st))) return r;
    if ((r = copy_weight(ctx, &L.ln2, w[CSM_W_LN2], H, st))) return r;
    if (keep_natural) {
      DA(L.qkv, (size_t)(nq + 2 * nkv) * H);
      CK(cudaMemcpyAsync(L.qkv, w[CSM_W_Q], (size_t)nq * H * 2, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(L.qkv + (size_t)nq * H, w[CSM_W_K], (size_t)nkv * H * 2, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(L.qkv + (size_t)(nq + nkv) * H, w[CSM_W_V], (size_t)nkv * H * 2, cudaMemcpyDeviceToDevice, st));
      if ((r = copy_weight(ctx, &L.o, w[CSM_W_O], (size_t)H * nq, st))) return r;
      DA(L.gu, (size_t)2 * I * H);
      CK(csm_interleave_rows_launch((const bf16*)w[CSM_W_GATE], (const bf16*)w[CSM_W_UP], I, H, L.gu, st));
      if ((r = copy_weight(ctx, &L.down, w[CSM_W_DOWN], (size_t)H * I, st))) return r;
      const int bw = csm_gemm_box_rows_w();
      if (csm_tmap_2d(&L.tm_qkv, L.qkv, nq + 2 * nkv, H, H, bw) || csm_tmap_2d(&L.tm_o, L.o, H, nq, nq, bw) ||
          csm_tmap_2d(&L.tm_gu, L.gu, 2 * I, H, H, bw) || csm_tmap_2d(&L.tm_down, L.down, H, I, I, bw))
        return fail(ctx, CSM_ECUDA, "cuTensorMapEncodeTiled failed for the layer %d weights", l);
    }
    PackSrc s;
    memset(&s, 0, sizeof s);
    s.ptr[0] = (const bf16*)w[CSM_W_Q]; s.rows[0] = nq;
    s.ptr[1] = (const bf16*)w[CSM_W_K]; s.rows[1] = nkv;
    s.ptr[2] = (const bf16*)w[CSM_W_V]; s.rows[2] = nkv;
    for (int i = 0; i < 3; ++i) { s.row_stride[i] = H; s.col_stride[i] = 1; }
    if ((r = pack_matrix(ctx, &L.p_qkv, s, &qkv_map, nq + 2 * nkv, H, 2, st))) return r;
    if ((r = pack_matrix(ctx, &L.p_o, one_src((const bf16*)w[CSM_W_O], H, nq, 1), nullptr, H, nq, 1, st))) return r;
    memset(&s, 0, sizeof s);
    s.ptr[0] = (const bf16*)w[CSM_W_GATE]; s.rows[0] = I;
    s.ptr[1] = (const bf16*)w[CSM_W_UP]; s.rows[1] = I;
    s.ptr[2] = s.ptr[1]; s.rows[2] = 0;
    for (int i = 0; i < 3; ++i) { s.row_stride[i] = H; s.col_stride[i] = 1; }
    if ((r = pack_matrix(ctx, &L.p_gu, s, &gu_map, 2 * I, H, 2, st))) return r;
    if ((r = pack_matrix(ctx, &L.p_down, one_src((const bf16*)w[CSM_W_DOWN], H, I, 1), nullptr, H, I, 1, st,
                         ctx->pair ? ctx->G / 2 : 0)))
      return r;
  }
  int r;
  if ((r = copy_weight(ctx, &S.norm, norm, H, st))) return r;
  S.n_pos = sh.n_pos;
  DA(S.cos_t, (size_t)sh.n_pos * half);
  DA(S.sin_t, (size_t)sh.n_pos * half);
  CK(cudaMemcpyAsync(S.cos_t, sh.rope_cos, (size_t)sh.n_pos * half * 2, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S.sin_t, sh.rope_sin, (size_t)sh.n_pos * half * 2, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

Phase gemv(int act_mode, int epi, int gran, int N, int K, int stack, int layer, const bf16* w, const void* act,
           int act_stride, const bf16* norm_w, void* out, int out_stride, int src_ph, int res_ph = 0) {
  Phase P;
  memset(&P, 0, sizeof P);
  P.type = PH_GEMV; P.act_mode = act_mode; P.epi = epi; P.gran = gran; P.N = N; P.K = K; P.stack = stack;
  P.layer = layer; P.w = w; P.act = (const bf16*)act; P.act_stride = act_stride; P.norm_w = norm_w; P.out = (bf16*)out;
  P.out_stride = out_stride; P.src_ph = src_ph; P.res_ph = res_ph;
  return P;
}

// One transformer layer.  h_ph: index of the phase that last wrote the residual stream (in/out).
// gather_cb >= 0 (decoder layer 0 of positions 1..31): the layer input is not read from the residual stream but
// gathered from the pre-projected embedding table with the token sampled by head phase `head_ph`; the qkv phase
// also starts the residual stream (one CTA per sequence writes the gathered row as tagged words).
void add_layer_phases(CsmCtx* ctx, Stack& S, int stack, int l, int dec_pos, bool kv_only, int& h_ph, int gather_cb = -1,
                      int head_ph = 0) {
  const StackDims& d = S.d;
  LayerW& L = S.layers[l];
  uint32_t* h = stack ? ctx->h_dec : ctx->h_bb;
  uint32_t* qb = stack ? ctx->q_dec : ctx->q_bb;
  uint32_t* at = stack ? ctx->attn_dec : ctx->attn_bb;
  uint32_t* mlp = stack ? ctx->mlp_dec : ctx->mlp_bb;
  const int nq = d.heads * d.hd, nkv = d.kv * d.hd;
  const int pad = ctx->fuse_attn ? 0 : 8;   // general kernels: rows of the inter-phase vectors are K + 8 apart (csm_batch.inl)
  auto idx = [&]() { return (int)ctx->table.size(); };
  const int iq = idx();
  Phase P = gemv(ACT_NORM, EPI_QKV, 2, nq + 2 * nkv, d.H, stack, l, L.p_qkv, h, d.H + pad, L.ln1, qb, nq + 2 * nkv, h_ph);
  if (gather_cb >= 0) {
    P.act_mode = ACT_GATHER;
    P.act = ctx->proj_table;
    P.cb = gather_cb;
    P.res_ph = head_ph;                 // candidates to reduce
    P.norm_out = (bf16*)h;              // tagged residual stream to start
    h_ph = iq;
  }
  P.dec_pos = dec_pos;
  if (stack && ctx->fuse_attn && !kv_only) P.flags |= CSM_PF_KV_COPY;
  ctx->table.push_back(P);
  if (kv_only) return;
  int io;
  if (stack && ctx->fuse_attn) {
    // small batch: every CTA computes the (tiny) decoder attention itself while staging o_proj's input
    io = idx();
    P = gemv(ACT_ATTN, EPI_RESID, 1, d.H, nq, stack, l, L.p_o, qb, nq + 2 * nkv, nullptr, h, d.H, iq, h_ph);
    P.dec_pos = dec_pos;
    ctx->table.push_back(P);
  } else {
    const int ia = idx();
    memset(&P, 0, sizeof P);
    P.type = stack ? PH_ATTN_DEC : PH_ATTN_BB; P.stack = stack; P.layer = l; P.dec_pos = dec_pos; P.src_ph = iq;
    ctx->table.push_back(P);
    io = idx();
    ctx->table.push_back(gemv(ACT_PLAIN, EPI_RESID, 1, d.H, nq, stack, l, L.p_o, at, nq + pad, nullptr, h, d.H + pad, ia, h_ph));
  }
  const int ig = idx();
  P = gemv(ACT_NORM, EPI_SWIGLU, 2, 2 * d.I, d.H, stack, l, L.p_gu, h, d.H + pad, L.ln2, mlp, d.I + pad, io);
  if (!ctx->direct_mlp) P.flags |= CSM_PF_OUT_PLAIN | CSM_PF_BAR_OUT;
  ctx->table.push_back(P);
  const int id = idx();
  P = gemv(ctx->direct_mlp ? ACT_PLAIN : ACT_STREAM, EPI_RESID, 1, d.H, d.I, stack, l, L.p_down, mlp, d.I + pad, nullptr, h,
           d.H + pad, ig, io);
  if (!ctx->direct_mlp) P.flags |= CSM_PF_BAR_IN;
  P.pair = ctx->pair && !ctx->direct_mlp;
  ctx->table.push_back(P);
  h_ph = id;
}

void build_table(CsmCtx* ctx) {
  ctx->table.clear();
  const StackDims& b = ctx->bb.d;
  const StackDims& d = ctx->dec.d;
  Phase P;
  memset(&P, 0, sizeof P);
  P.type = PH_EMBED;
  ctx->table.push_back(P);
  int hb_ph = 0;
  for (int l = 0; l < b.L; ++l) add_layer_phases(ctx, ctx->bb, 0, l, 0, false, hb_ph);
  ctx->ph_head_c0 = (int)ctx->table.size();
  // final norm (-> last_hidden_state) + codebook-0 head + greedy sample (modeling_csm.py:361-365,531-532)
  int head_ph = (int)ctx->table.size();
  const int pad = ctx->fuse_attn ? 0 : 8;
  P = gemv(ACT_NORM, EPI_HEAD, 1, ctx->V, b.H, 0, 0, ctx->p_c0, ctx->h_bb, b.H + pad, ctx->bb.norm, ctx->c0_logits, ctx->V, hb_ph);
  P.cb = 0;
  P.norm_out = ctx->last_h;
  ctx->table.push_back(P);
  for (int pos = 0; pos < CSM_DEC_POS; ++pos) {
    // projection of last_h (pos 0: the backbone's final norm is recomputed from the residual stream, which is
    // bit-identical to reading last_h and saves a dependency) or of the previous codebook's embedding
    // (modeling_csm.py:535-542,564-565)
    int hd_ph = (int)ctx->table.size();
    if (pos == 0) {
      P = gemv(ACT_NORM, EPI_STORE, 1, d.H, b.H, 0, 0, ctx->p_proj, ctx->h_bb, b.H + pad, ctx->bb.norm, ctx->h_dec, d.H + pad, hb_ph);
      ctx->table.push_back(P);
    }
    // positions 1..31: projection(embedding(token)) is a row of proj_table, gathered by layer 0's qkv phase
    for (int l = 0; l < d.L; ++l)
      add_layer_phases(ctx, ctx->dec, 1, l, pos, pos == 0 && l == d.L - 1, hd_ph, (pos >= 1 && l == 0) ? pos - 1 : -1,
                       head_ph);
    if (pos >= 1) {
      // audio_head[pos-1] on the decoder's final-norm output, greedy sample (modeling_csm.py:557-560)
      head_ph = (int)ctx->table.size();
      P = gemv(ACT_NORM, EPI_HEAD, 1, ctx->V, d.H, 1, 0, ctx->p_heads[pos - 1], ctx->h_dec, d.H + pad, ctx->dec.norm,
               ctx->cb_logits + (size_t)(pos - 1) * ctx->V, (CSM_NQ - 1) * ctx->V, hd_ph);
      P.cb = pos;
      ctx->table.push_back(P);
    }
  }
  // sample codebook 31, publish the frame, stop rule (modeling_csm.py:657-666)
  memset(&P, 0, sizeof P);
  P.type = PH_FINISH;
  P.res_ph = head_ph;
  ctx->table.push_back(P);
  if (ctx->bar_all) {
    // conservative mode: a grid barrier between every two phases, on top of the tagged hand-over
    for (size_t i = 1; i < ctx->table.size(); ++i) {
      ctx->table[i].flags |= CSM_PF_BAR_IN;
      ctx->table[i - 1].flags |= CSM_PF_BAR_OUT;
    }
  }
  int nbar = 0;
  for (Phase& Q : ctx->table) {
    if (Q.flags & CSM_PF_BAR_IN) ++nbar;
    Q.bar_idx = nbar;
  }
}

// Shared-memory plan of the frame kernel for max_batch sequences, and the per-phase row split /
// ring chunking that depends on it (Phase::q, r, tpc, nch).
int plan_smem(CsmCtx* ctx) {
  const int G = ctx->G;
  const int kfull = ctx->bb.d.H > ctx->dec.d.H ? ctx->bb.d.H : ctx->dec.d.H;
  ctx->m_alloc = (ctx->Bmax + 7) / 8 * 8;
  ctx->rope_bytes = (2 * CSM_DEC_POS * (ctx->dec.d.hd / 2) + 2 * (ctx->bb.d.hd / 2)) * 2;
  ctx->rope_bytes = (ctx->rope_bytes + 255) / 256 * 256;
  // split-K partial sums: red[ks][m_alloc][mtiles*16+4] floats; attention scratch needs 9216 bytes
  int red = 9216;
  for (Phase& P : ctx->table) {
    if (P.type != PH_GEMV) continue;
    const int U = P.N / P.gran;
    const int Gp = P.pair ? G / 2 : G;   // pair phases: rows over CTA pairs
    P.q = U / Gp;
    P.r = U % Gp;
    const int g4 = P.K / 4;
    P.gsh = -1;
    if ((g4 & (g4 - 1)) == 0) { P.gsh = 0; while ((1 << P.gsh) < g4) ++P.gsh; }
    if (P.act_mode != ACT_STREAM && P.gsh < 6)
      return fail(ctx, CSM_EINVAL, "reduction length %d: the staging path needs K/4 to be a power of two >= 64", P.K);
    for (int cls = 0; cls < 2; ++cls) {
      GeoC& gc = P.geo[cls];
      memset(&gc, 0, sizeof gc);
      const int rows = (P.q + (cls == 0 ? 1 : 0)) * P.gran;
      gc.rows = rows;
      const int mt = (rows + 15) / 16;
      int ns = mt >= 5 ? 8 : (mt >= 3 ? 4 : (mt >= 2 ? 2 : 1));
      if (ctx->mt2) ns = mt >= 9 ? 8 : (mt >= 5 ? 4 : (mt >= 3 ? 2 : 1));   // every warp takes two m-tiles: half the B-fragment reads
      gc.mtiles = mt;
      gc.ksl = ns == 8 ? 0 : (ns == 4 ? 1 : (ns == 2 ? 2 : 3));
      gc.rows_pad = mt * 16 + 4;
      gc.upc = rows / P.gran;
      gc.ush = 0;
      while ((1 << gc.ush) < gc.upc) ++gc.ush;
      if (rows == 0 || (cls == 0 && P.r == 0)) continue;
      const int need = (8 / ns) * ctx->m_alloc * (mt * 16 + 4) * 4;
      if (need > red) red = need;
    }
  }
  ctx->red_bytes = (red + 255) / 256 * 256;
  if (ctx->pair) {   // exchange buffer of the CTA-pair phases: [m_alloc][16] fp32 behind the split-K partials
    ctx->xbuf_off = ctx->red_bytes;
    ctx->red_bytes += ctx->m_alloc * 16 * 4;
    ctx->red_bytes = (ctx->red_bytes + 255) / 256 * 256;
  }
  ctx->act_region = ctx->m_alloc * (kfull + 8) * 2;
  if (ctx->direct_mlp) {
    const int imax = ctx->bb.d.I > ctx->dec.d.I ? ctx->bb.d.I : ctx->dec.d.I;
    const int need = ctx->Bmax * (imax + 8) * 2;
    if (need > ctx->act_region) ctx->act_region = need;
  }
  if (ctx->fuse_attn) {
    // fused decoder attention: o_proj input rows + K|V of <= 32 cached positions (padded rows) + q scratch
    const StackDims& d = ctx->dec.d;
    const int need = ctx->m_alloc * (d.heads * d.hd + 8) * 2 + ctx->Bmax * 2 * d.kv * CSM_DEC_POS * (d.hd + 8) * 2 +
                     CSM_COMPUTE_WARPS * d.hd * 4;
    if (need > ctx->act_region) ctx->act_region = need;
  }
  if (!ctx->fuse_attn) {
    // general kernels: the norm weights of a phase are staged behind its rows; room for the activation-tile ring
    ctx->normw_off = ctx->m_alloc * (kfull + 8) * 2;
    if (ctx->act_region < ctx->normw_off + kfull * 2) ctx->act_region = ctx->normw_off + kfull * 2;
    if (ctx->act_region < 65536) ctx->act_region = 65536;
  }
  ctx->act_region = (ctx->act_region + 255) / 256 * 256;
  const int limit = 227 * 1024;
  const int avail = limit - CSM_SM_HDR_BYTES - ctx->rope_bytes - ctx->red_bytes - ctx->act_region;
  int slot = 32 * 1024;
  while (slot > 4096 && avail / slot < 3) slot /= 2;
  if (avail / slot < 2) return fail(ctx, CSM_ECAPACITY, "batch %d leaves no shared memory for the weight ring", ctx->Bmax);
  ctx->slot_bytes = slot;
  ctx->n_slots = avail / slot;
  if (ctx->n_slots > CSM_MAX_SLOTS) ctx->n_slots = CSM_MAX_SLOTS;
  const char* e = getenv("CSM_RING_SLOTS");
  if (e && atoi(e) >= 2 && atoi(e) <= ctx->n_slots) ctx->n_slots = atoi(e);
  // activation-tile ring of the streamed (K = 8192) phases: slots of [m_alloc][tpc*16+8] bf16 inside the activation
  // region, k-chunks in lockstep with the weight chunks; at least 3 slots so that two copies are in flight while one
  // tile is consumed, tiles as long as that allows (k16-tile counts in units of 16 = 2 * the widest split-K)
  {
    int tpc = 64;   // (a power of two: the tiled layout of the producer indexes by shift and mask)
    for (; tpc > 16; tpc /= 2)
      if (ctx->act_region / (ctx->m_alloc * (tpc * 16 + 8) * 2) >= 3) break;
    if (const char* e = getenv("CSM_A_TPC")) { int v = atoi(e); if (v == 16 || v == 32 || v == 64) tpc = v < tpc ? v : tpc; }
    ctx->stream_tpc_max = tpc;
    ctx->a_slot_bytes = ctx->m_alloc * (tpc * 16 + 8) * 2;
    ctx->a_slots = ctx->act_region / ctx->a_slot_bytes;
    if (ctx->a_slots > CSM_MAX_SLOTS) ctx->a_slots = CSM_MAX_SLOTS;
    if (ctx->a_slots < 2) return fail(ctx, CSM_ECAPACITY, "activation region too small");
  }
  ctx->smem_total = (size_t)CSM_SM_HDR_BYTES + ctx->rope_bytes + ctx->red_bytes + ctx->act_region +
                    (size_t)ctx->slot_bytes * ctx->n_slots;
  for (Phase& P : ctx->table) {
    if (P.type != PH_GEMV) continue;
    const int ntiles = P.pair ? P.K / 32 : P.K / 16;   // (pair phases: each CTA of a pair reduces over half of K)
    for (int cls = 0; cls < 2; ++cls) {
      GeoC& gc = P.geo[cls];
      const int rows = gc.rows;
      int tpc = 0, nch = 0;
      if (rows > 0) {
        tpc = ctx->slot_bytes / (rows * 32);
        if (tpc > ntiles) tpc = ntiles;
        if (P.act_mode == ACT_STREAM && tpc > ctx->stream_tpc_max) tpc = ctx->stream_tpc_max;
        // every chunk must hold a multiple of 2*ks k16-tiles: each warp then consumes whole (tl, tl+ks) pairs
        const int unit = 2 << gc.ksl;
        tpc = tpc / unit * unit;
        if (tpc < unit || ntiles % unit)
          return fail(ctx, CSM_EINVAL, "matrix %dx%d: %d rows per CTA do not fit a ring slot in units of %d k-tiles", P.N,
                      P.K, rows, unit);
        nch = (ntiles + tpc - 1) / tpc;
      }
      gc.tpc = tpc;
      gc.nch = nch;
    }
    if (P.act_mode == ACT_STREAM) {
      // the gate/up phase that feeds this streamed phase writes [k-tile][m_alloc][len+8] with len = the k-chunk of this
      // phase: both CTA classes chunk alike, by a power-of-two number of k16-tiles
      const bool use0 = P.r > 0, use1 = P.q > 0;
      int t = ntiles, unit = 2;
      for (int cls = 0; cls < 2; ++cls)
        if (cls == 0 ? use0 : use1) {
          if (P.geo[cls].tpc < t) t = P.geo[cls].tpc;
          if ((2 << P.geo[cls].ksl) > unit) unit = 2 << P.geo[cls].ksl;
        }
      int tp = 1;
      while (tp * 2 <= t) tp *= 2;
      if (tp < unit || ntiles % tp || ctx->a_slot_bytes < ctx->m_alloc * (tp * 16 + 8) * 2)
        return fail(ctx, CSM_EINVAL, "streamed phase %dx%d: no common k-chunk (%d tiles, unit %d)", P.N, P.K, tp, unit);
      for (int cls = 0; cls < 2; ++cls) {
        P.geo[cls].tpc = P.geo[cls].rows > 0 ? tp : 0;
        P.geo[cls].nch = P.geo[cls].rows > 0 ? ntiles / tp : 0;
      }
      Phase& Q = ctx->table[P.src_ph];
      Q.tile_sh = 4;   // log2(16 * tp)
      while ((16 << (Q.tile_sh - 4)) < tp * 16) ++Q.tile_sh;
    }
  }
  return 0;
}

// Tags are 16 bits: when the next frame's tags would pass 65535, clear every tagged buffer (tag 0 = never
// written) and restart at 1.  Happens once every ~80 frames; a few hundred KB of stream-ordered memsets.
int begin_epoch(CsmCtx* ctx, cudaStream_t st) {
  if (ctx->tagbase + ctx->table.size() > 65535u) {
    for (auto& t : ctx->tagged) CK(cudaMemsetAsync(t.first, 0, t.second, st));
    ctx->tagbase = 1;
  }
  return 0;
}
void end_epoch(CsmCtx* ctx) { ctx->tagbase += (unsigned)ctx->table.size(); }

// A wait inside a frame kernel timed out (hang guard, csm_stream.cu): report who waited for what.  Synchronises.
// Host-side watchdog: wait for the stream with a deadline instead of blocking for ever.  The frame kernels of the
// <= 2-sequence engines carry no in-kernel hang guard (it costs them 13-16 %, measured: DESIGN.md section 3.1), so a lost
// hand-over there would spin until the driver's own watchdog; the caller at least gets an error, not a hung process.
int wait_stream(CsmCtx* ctx, cudaStream_t st, double budget_s) {
  cudaEvent_t ev = ctx->ev_done;
  CK(cudaEventRecord(ev, st));
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned it = 0;; ++it) {
    cudaError_t e = cudaEventQuery(ev);
    if (e == cudaSuccess) return 0;
    if (e != cudaErrorNotReady) return fail(ctx, CSM_ECUDA, "stream failed: %s", cudaGetErrorString(e));
    if ((it & 63u) == 63u) {
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > budget_s)
        return fail(ctx, CSM_ECUDA, "frame kernels did not finish within %.0f s (a hand-over was lost?): the device needs a "
                    "reset", budget_s);
      if (dt > 0.002) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
}

int check_abort(CsmCtx* ctx, cudaStream_t st) {
  int a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CK(cudaMemcpyAsync(a, ctx->abort_flag, sizeof a, cudaMemcpyDeviceToHost, st));
  {
    int r = wait_stream(ctx, st, ctx->wait_budget_s);
    if (r) return r;
  }
  if (a[0] == 0) return 0;
  CK(cudaMemsetAsync(ctx->abort_flag, 0, sizeof a, st));   // the engine stays usable (reset + new generate)
  return fail(ctx, CSM_ECUDA,
              "frame kernel timed out: CTA %d, phase %d, wait kind %d (1 stage 2 cand 3 attn_dec 4/5 attn_bb 6 resid 7 grid "
              "barrier 8/9 ring full 10/11 ring empty), detail %d / 0x%x, thread %d", a[1], a[2], a[3], a[4], a[5], a[6]);
}

// Two kernel families: engines for <= 2 sequences run the pure-dataflow kernels of csm_stream.inl, larger engines the
// barrier-separated kernels of csm_batch.inl; each in a greedy and a stochastic-sampling build.
cudaError_t launch_kernel(CsmCtx* ctx, const StreamParams* p, cudaStream_t st, int cooperative) {
  if (ctx->fuse_attn)
    return p->topk > 1 ? csm_launch_stream_small_stoch(p, ctx->G, ctx->smem_total, st, cooperative)
                       : csm_launch_stream_small(p, ctx->G, ctx->smem_total, st, cooperative);
  const int cluster = ctx->pair ? 2 : 1;
  return p->topk > 1 ? csm_launch_batch_stoch(p, ctx->G, ctx->smem_total, st, cooperative, cluster)
                     : csm_launch_batch(p, ctx->G, ctx->smem_total, st, cooperative, cluster);
}

int launch_frame(CsmCtx* ctx, int B, int ph_begin, int ph_end, const long long* ids, const int* mask, int forced,
                 long long* out_frames, long long out_stride, long long out_off, int stop_on_zeros, int pos,
                 cudaStream_t st) {
  StreamParams p;
  memset(&p, 0, sizeof p);
  p.phases = ctx->d_table;
  p.B = B; p.pos = pos; p.Bmax = ctx->Bmax; p.Tcap = ctx->Tcap; p.V = ctx->V; p.text_vocab = ctx->text_vocab;
  p.bb = ctx->bb.d; p.dec = ctx->dec.d;
  p.bar_counter = ctx->bar_counter;
  p.kc_bb = ctx->kc_bb; p.vc_bb = ctx->vc_bb; p.kc_dec = ctx->kc_dec; p.vc_dec = ctx->vc_dec;
  p.cos_bb = ctx->bb.cos_t; p.sin_bb = ctx->bb.sin_t; p.cos_dec = ctx->dec.cos_t; p.sin_dec = ctx->dec.sin_t;
  p.q_bb = ctx->q_bb; p.q_dec = ctx->q_dec; p.attn_bb = ctx->attn_bb; p.attn_dec = ctx->attn_dec;
  p.attn_part = ctx->attn_part; p.nsplit_max = ctx->nsplit_max; p.attn_cnt = ctx->attn_cnt;
  p.cand = ctx->cand;
  p.samples = ctx->samples; p.fed = ctx->fed; p.forced = forced;
  p.ids = ids; p.mask = mask; p.text_emb = ctx->text_emb; p.audio_emb = ctx->audio_emb;
  p.h_bb = ctx->h_bb;
  p.out_frames = out_frames; p.out_stride = out_stride; p.out_off = out_off;
  p.stop_flag = ctx->stop_flag; p.n_frames = ctx->n_frames; p.stop_on_zeros = stop_on_zeros;
  p.m_alloc = ctx->m_alloc; p.slot_bytes = ctx->slot_bytes; p.n_slots = ctx->n_slots;
  p.rope_bytes = ctx->rope_bytes; p.act_region_bytes = ctx->act_region; p.red_bytes = ctx->red_bytes;
  p.a_slots = ctx->a_slots; p.a_slot_bytes = ctx->a_slot_bytes;
  p.hpad = ctx->fuse_attn ? 0 : 8;
  p.normw_off = ctx->normw_off;
  p.xbuf_off = ctx->xbuf_off;
  // backbone-attention K/V pieces: as large as the per-warp share of the activation region allows (fewer, larger
  // bulk copies), at most 128 positions; the stages that fit behind that
  {
    const int per_warp = ctx->act_region / CSM_COMPUTE_WARPS;
    int ps = per_warp >= 16384 ? 128 : (per_warp >= 8192 ? 64 : 32);
    if (const char* e = getenv("CSM_ATT_PS")) { int v = atoi(e); if ((v == 32 || v == 64 || v == 128) && v * 128 <= per_warp) ps = v; }
    p.att_ps = ps;
    p.att_stages = per_warp / (ps * 128);
    if (p.att_stages > 4) p.att_stages = 4;
    if (const char* e = getenv("CSM_ATT_STAGES")) { int v = atoi(e); if (v >= 1 && v <= p.att_stages) p.att_stages = v; }
  }
  {
    // units of the backbone attention: (sequence, kv-head, nsub x 128 positions): the smallest nsub for which every
    // compute warp of the grid gets at most ONE unit -- a unit is one long chain of dependent L2 / HBM round trips
    // (K/V pieces, partial, counter, merge), and a warp runs its units one after the other
    const int blocks = (pos + 1 + CSM_ATT_SPLIT_MMA - 1) / CSM_ATT_SPLIT_MMA;
    const long long streams = (long long)B * ctx->bb.d.kv, warps = (long long)CSM_COMPUTE_WARPS * ctx->G;
    int nsub = 1;
    while (nsub < 8 && streams * ((blocks + nsub - 1) / nsub) > warps) ++nsub;
    if (const char* e = getenv("CSM_ATT_NSUB")) nsub = atoi(e) > 0 ? atoi(e) : nsub;
    p.att_nsub = nsub;
  }
  p.att_pf_units = 12;   // x 32 KB x 148 CTAs = 57 MB of the 126 MB L2 per layer at most
  if (const char* e = getenv("CSM_ATT_PF")) p.att_pf_units = atoi(e);
  p.prof = ctx->prof_on ? ctx->prof : nullptr;
  p.n_phases_total = (int)ctx->table.size();
  p.progress = ctx->progress_on ? ctx->progress : nullptr;
  p.abort_flag = ctx->abort_flag;
  p.topk = ctx->topk; p.inv_temp = ctx->inv_temp; p.rng_seed = ctx->rng_seed; p.rng_frame = ctx->rng_frame;
  p.seq_base = ctx->seq_base; p.lgt = ctx->lgt; p.lgt_stride = ctx->lgt_stride;
  p.tagbase = ctx->tagbase;
  p.l2_ahead_bytes = ctx->l2_ahead;
  p.repl = ctx->repl;
  p.evict_first = ctx->evict_first;
  p.small = ctx->fuse_attn;
  if (!ctx->stepped) {
    p.phase_begin = ph_begin; p.phase_end = ph_end; p.use_barrier = 1;
    CK(cudaMemsetAsync(ctx->bar_counter, 0, sizeof(unsigned int), st));
    CK(launch_kernel(ctx, &p, st, 1));
    ctx->launches += 1;
  } else {
    p.use_barrier = 0;
    for (int ph = ph_begin; ph < ph_end; ++ph) {
      p.phase_begin = ph; p.phase_end = ph + 1;
      CK(launch_kernel(ctx, &p, st, 0));
      ctx->launches += 1;
    }
  }
  return 0;
}

// C[R,N] = A[R,K] * W[N,K]^T on the tcgen05 path (csm_gemm.cu).  `map_w`: tensor map of W made at create time;
// the map of A is encoded here (host-side, no driver call that touches the device).
int gemm_tc(CsmCtx* ctx, const bf16* A, int lda, const CUtensorMap* map_w, GemmParams g, cudaStream_t st) {
  CUtensorMap map_a;
  if (csm_tmap_2d(&map_a, A, g.R, g.K, lda, csm_gemm_box_rows_a()))
    return fail(ctx, CSM_ECUDA, "cuTensorMapEncodeTiled failed for an activation matrix [%d,%d]", g.R, g.K);
  CK(csm_gemm_launch(&map_a, map_w, &g, ctx->sms, st));
  ctx->launches += 1;
  return 0;
}

int ensure_prefill_ws(CsmCtx* ctx, int rows) {
  if (rows <= ctx->pf_rows) return 0;
  const StackDims& d = ctx->bb.d;
  const size_t W = (size_t)(d.heads + 2 * d.kv) * d.hd;
  // growth: the outgrown buffers are released first (nothing is in flight on them: prefill calls are stream-ordered
  // and the previous call's kernels were enqueued before this cudaFree, which synchronises)
  bf16** bufs[] = {&ctx->pf_h, &ctx->pf_hn, &ctx->pf_qkv, &ctx->pf_attn, &ctx->pf_act};
  for (bf16** b : bufs) {
    if (*b) CK(cudaFree(*b));
    *b = nullptr;
  }
  if (ctx->pf_valid) CK(cudaFree(ctx->pf_valid));
  ctx->pf_valid = nullptr;
  ctx->pf_rows = 0;
  CK(cudaMalloc(&ctx->pf_h, (size_t)rows * d.H * 2));
  CK(cudaMalloc(&ctx->pf_hn, (size_t)rows * d.H * 2));
  CK(cudaMalloc(&ctx->pf_qkv, (size_t)rows * W * 2));
  CK(cudaMalloc(&ctx->pf_attn, (size_t)rows * d.heads * d.hd * 2));
  CK(cudaMalloc(&ctx->pf_act, (size_t)rows * d.I * 2));
  CK(cudaMalloc(&ctx->pf_valid, (size_t)rows));
  ctx->pf_rows = rows;
  return 0;
}

// Backbone over S new positions for sequences [0,B): fills the KV cache and leaves the last position's
// residual-stream row (pre final-norm) in h_bb[b].  Per layer: RMSNorm rows, q|k|v GEMM with the RoPE + KV-cache
// write tail, causal GQA flash attention, o_proj GEMM with the residual tail, RMSNorm rows, gate|up GEMM with the
// SwiGLU tail, down_proj GEMM with the residual tail -- seven launches, all of them kernels of this library
// (hf LlamaDecoderLayer.forward, modeling_llama.py:303-332).
// Padding (modeling_csm.py:337-342): frames whose 33 mask entries are all zero are hidden as KEYS in this call
// (a query that sees no key gets a zero attention output, so a padded position stays exactly zero through every
// layer and caches K = V = 0); decode steps attend to every cached position, padded ones included -- the reference's
// behaviour.  Honoured for a prefill into an empty cache, which is how generate() runs it.
int prefill(CsmCtx* ctx, const long long* ids, const int* mask, int B, int S, cudaStream_t st) {
  const StackDims& d = ctx->bb.d;
  const int W = (d.heads + 2 * d.kv) * d.hd, nq = d.heads * d.hd;
  int group = 16384 / S;
  if (group < 1) group = 1;
  if (group > B) group = B;
  int r = ensure_prefill_ws(ctx, group * S);
  if (r) return r;
  const int pos0 = ctx->cache_len;
  for (int b0 = 0; b0 < B; b0 += group) {
    const int nseq = (B - b0) < group ? (B - b0) : group;
    const int R = nseq * S;
    const long long* gi = ids + (size_t)b0 * S * (CSM_NQ + 1);
    const int* gm = mask ? mask + (size_t)b0 * S * (CSM_NQ + 1) : nullptr;
    const unsigned char* valid = nullptr;
    if (gm && pos0 == 0) {
      CK(csm_frame_valid_launch(gm, R, ctx->pf_valid, nullptr, st));
      valid = ctx->pf_valid;
      ctx->launches += 1;
    }
    CK(csm_embed_sum_launch(gi, gm, 2, ctx->audio_emb, ctx->text_emb, ctx->V, d.H, ctx->pf_h, R, st));
    ctx->launches += 1;
    GemmParams g;
    memset(&g, 0, sizeof g);
    g.R = R;
    for (int l = 0; l < d.L; ++l) {
      LayerW& L = ctx->bb.layers[l];
      CK(csm_rmsnorm_rows_launch(ctx->pf_h, L.ln1, d.eps, d.H, ctx->pf_hn, R, st));
      g.N = W; g.K = d.H; g.epi = EPI_QKV; g.C = ctx->pf_qkv; g.ldc = W;
      g.S = S; g.pos0 = pos0; g.b0 = b0; g.heads = d.heads; g.kv = d.kv; g.layer = l; g.Bmax = ctx->Bmax; g.Tcap = ctx->Tcap;
      g.kc = ctx->kc_bb; g.vc = ctx->vc_bb; g.cos_t = ctx->bb.cos_t; g.sin_t = ctx->bb.sin_t;
      if ((r = gemm_tc(ctx, ctx->pf_hn, d.H, &L.tm_qkv, g, st))) return r;
      if (pos0 == 0 && d.hd == 64 && ctx->flash_tc)   // tcgen05 flash attention (csm_flash_tc.cu); CSM_FLASH_MMA=1: mma.sync
        CK(csm_flash_tc_prefill_launch(ctx->pf_qkv, S, b0, nseq, d.heads, d.kv, ctx->kc_bb, ctx->vc_bb, l, d.L, ctx->Bmax,
                                       ctx->Tcap, d.scale, valid, ctx->pf_attn, st));
      else
        CK(csm_flash_prefill_launch(ctx->pf_qkv, S, pos0, b0, nseq, d.heads, d.kv, ctx->kc_bb, ctx->vc_bb, l, ctx->Bmax,
                                    ctx->Tcap, d.scale, valid, ctx->pf_attn, st));
      g.N = d.H; g.K = nq; g.epi = EPI_RESID; g.C = ctx->pf_h; g.ldc = d.H;
      if ((r = gemm_tc(ctx, ctx->pf_attn, nq, &L.tm_o, g, st))) return r;
      CK(csm_rmsnorm_rows_launch(ctx->pf_h, L.ln2, d.eps, d.H, ctx->pf_hn, R, st));
      g.N = 2 * d.I; g.K = d.H; g.epi = EPI_SWIGLU; g.C = ctx->pf_act; g.ldc = d.I;
      if ((r = gemm_tc(ctx, ctx->pf_hn, d.H, &L.tm_gu, g, st))) return r;
      g.N = d.H; g.K = d.I; g.epi = EPI_RESID; g.C = ctx->pf_h; g.ldc = d.H;
      if ((r = gemm_tc(ctx, ctx->pf_act, d.I, &L.tm_down, g, st))) return r;
      ctx->launches += 3;
    }
    // hand the last position's residual row to the frame kernel: tagged words of the last backbone phase for the
    // <= 2-sequence kernels, plain bf16 rows for the general kernels
    for (int rr = 0; rr < ctx->repl; ++rr)
      CK(csm_take_last_rows_launch(ctx->pf_h, S, d.H, ctx->h_bb + (size_t)rr * ctx->Bmax * d.H, b0, nseq,
                                   (ctx->tagbase + (unsigned)(ctx->ph_head_c0 - 1)) & 0xffffu, !ctx->fuse_attn, st));
    ctx->launches += 1;
  }
  return 0;
}

int frame_impl(CsmCtx* ctx, const long long* ids, const int* mask, int B, int S, const long long* force_tokens,
               long long* out_frames, long long out_stride, long long out_off, int stop_on_zeros, cudaStream_t st) {
  if (B < 1 || S < 1) return fail(ctx, CSM_EINVAL, "B and S must be >= 1 (got %d, %d)", B, S);
  if (B > ctx->Bmax) return fail(ctx, CSM_ECAPACITY, "batch %d > max_batch %d", B, ctx->Bmax);
  if (ctx->cache_len + S > ctx->Tcap)
    return fail(ctx, CSM_ECAPACITY, "context %d + %d exceeds max_ctx %d", ctx->cache_len, S, ctx->Tcap);
  int forced = 0;
  if (force_tokens) {
    CK(csm_i64_to_i32_launch(force_tokens, ctx->fed, B * CSM_NQ, st));
    forced = 1;
  }
  int r;
  if ((r = begin_epoch(ctx, st))) return r;
  if (S == 1) {
    r = launch_frame(ctx, B, 0, (int)ctx->table.size(), ids, mask, forced, out_frames, out_stride, out_off, stop_on_zeros,
                     ctx->cache_len, st);
  } else {
    if (!ids) return fail(ctx, CSM_EINVAL, "prefill needs input ids");
    if ((r = prefill(ctx, ids, mask, B, S, st))) return r;
    r = launch_frame(ctx, B, ctx->ph_head_c0, (int)ctx->table.size(), nullptr, nullptr, forced, out_frames, out_stride,
                     out_off, stop_on_zeros, ctx->cache_len + S - 1, st);
  }
  if (r) return r;
  end_epoch(ctx);
  ctx->rng_frame += 1;
  ctx->cache_len += S;
  return 0;
}

}  // namespace

// =================================================================== exported C ABI
extern "C" {

int csm_create(const CsmShapes* sh, const CsmWeights* w, int max_batch, int max_ctx, void* stream, CsmCtx** out) {
  if (!out) return CSM_EINVAL;
  *out = nullptr;
  CsmCtx* ctx = new CsmCtx();
  *out = ctx;   // returned even on failure so that csm_last_error works; caller destroys it
  cudaStream_t st = (cudaStream_t)stream;
  if (!sh || !w) return fail(ctx, CSM_EINVAL, "null shapes/weights");
  if (sh->n_codebooks != CSM_NQ) return fail(ctx, CSM_EINVAL, "audio_num_codebooks must be %d", CSM_NQ);
  const CsmLlamaShape* ls[2] = {&sh->backbone, &sh->decoder};
  for (int i = 0; i < 2; ++i) {
    const CsmLlamaShape& s = *ls[i];
    if (s.hidden % 64 || s.inter % 64 || s.hidden > 2048 || s.heads % s.kv_heads || s.hidden % s.heads)
      return fail(ctx, CSM_EINVAL, "unsupported llama shape (hidden %d inter %d heads %d kv %d)", s.hidden, s.inter,
                  s.heads, s.kv_heads);
    if (!s.rope_cos || !s.rope_sin) return fail(ctx, CSM_EINVAL, "missing rope tables");
  }
  if (sh->backbone.hidden / sh->backbone.heads != 64) return fail(ctx, CSM_EINVAL, "backbone head_dim must be 64");
  if (sh->decoder.hidden / sh->decoder.heads != 128) return fail(ctx, CSM_EINVAL, "decoder head_dim must be 128");
  {
    int rep = sh->backbone.heads / sh->backbone.kv_heads;
    if (rep != 1 && rep != 2 && rep != 4) return fail(ctx, CSM_EINVAL, "backbone GQA ratio must be 1, 2 or 4");
  }
  if (max_batch < 1 || max_batch > 32) return fail(ctx, CSM_ECAPACITY, "max_batch must be in [1,32] per GPU");
  if (max_ctx < 1 || sh->backbone.n_pos < max_ctx) return fail(ctx, CSM_EINVAL, "rope table shorter than max_ctx");
  if (sh->decoder.n_pos < CSM_DEC_POS) return fail(ctx, CSM_EINVAL, "decoder rope table shorter than %d", CSM_DEC_POS);
  CK(cudaGetDevice(&ctx->device));
  CK(cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device));
  int major = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ctx->device));
  if (major != 10) return fail(ctx, CSM_EUNSUPPORTED, "libcsm_b200 is built for sm_100a only (device is sm_%d)", major);
  if (ctx->sms > 160) return fail(ctx, CSM_EUNSUPPORTED, "more than 160 SMs (%d): candidate reduction assumes <= 160 CTAs", ctx->sms);
  ctx->G = ctx->sms;
  if (const char* e = getenv("CSM_GRID")) {
    int g = atoi(e);
    if (g >= 1 && g <= ctx->sms) ctx->G = g;
  }
  ctx->Bmax = max_batch;
  ctx->Tcap = max_ctx;
  ctx->V = sh->audio_vocab;
  ctx->text_vocab = sh->text_vocab;
  const int Hb = sh->backbone.hidden, Hd = sh->decoder.hidden;
  // Two kernel families.  Engines for <= 2 sequences: pure dataflow (tagged words), decoder attention fused into o_proj,
  // MLP activations staged whole (csm_stream.inl).  Larger engines: plain bf16 hand-over with a grid barrier per phase,
  // TMA-staged activations, K = 8192 phases streamed in tiles and split over CTA pairs (csm_batch.inl).
  ctx->fuse_attn = max_batch <= 2;
  ctx->direct_mlp = max_batch <= 4;
  if (const char* e = getenv("CSM_FUSE_ATTN")) ctx->fuse_attn = atoi(e) != 0;
  if (const char* e = getenv("CSM_DIRECT_MLP")) ctx->direct_mlp = atoi(e) != 0;
  if (!ctx->direct_mlp || max_batch > 4) ctx->fuse_attn = 0;   // the <= 2-sequence kernels have no streamed-activation path
  ctx->bar_all = !ctx->fuse_attn;
  // (a pair finishes at most 32 rows of a streamed matrix: its exchange buffer holds 16 per CTA)
  ctx->pair = !ctx->fuse_attn && !ctx->direct_mlp && (ctx->G % 2 == 0) &&
              (Hb + ctx->G / 2 - 1) / (ctx->G / 2) <= 32 && (Hd + ctx->G / 2 - 1) / (ctx->G / 2) <= 32;
  if (const char* e = getenv("CSM_PAIR")) ctx->pair = ctx->pair && atoi(e) != 0;
  int r;
  if ((r = copy_weight(ctx, &ctx->text_emb, w->text_embeddings, (size_t)sh->text_vocab * Hb, st))) return r;
  if ((r = copy_weight(ctx, &ctx->audio_emb, w->audio_embeddings, (size_t)sh->audio_vocab * CSM_NQ * Hb, st))) return r;
  if ((r = build_stack(ctx, ctx->bb, sh->backbone, w->backbone_layers, w->backbone_norm, true, st))) return r;
  if ((r = build_stack(ctx, ctx->dec, sh->decoder, w->decoder_layers, w->decoder_norm, false, st))) return r;
  if ((r = pack_matrix(ctx, &ctx->p_proj, one_src((const bf16*)w->projection, Hd, Hb, 1), nullptr, Hd, Hb, 1, st))) return r;
  // projection applied once to the whole audio table (modeling_csm.py:564-565 applies it to one gathered row per
  // codebook and frame: the same function of the same row, so it is folded into a [32*V, Hd] table)
  DA(ctx->proj_table, (size_t)sh->audio_vocab * CSM_NQ * Hd);
  {
    CUtensorMap tm_proj;
    if (csm_tmap_2d(&tm_proj, w->projection, Hd, Hb, Hb, csm_gemm_box_rows_w()))
      return fail(ctx, CSM_ECUDA, "cuTensorMapEncodeTiled failed for the projection weight");
    GemmParams g;
    memset(&g, 0, sizeof g);
    g.R = sh->audio_vocab * CSM_NQ; g.N = Hd; g.K = Hb; g.epi = EPI_STORE; g.C = ctx->proj_table; g.ldc = Hd;
    if ((r = gemm_tc(ctx, ctx->audio_emb, Hb, &tm_proj, g, st))) return r;
    CK(cudaStreamSynchronize(st));   // (tm_proj lives on this stack frame)
  }
  if ((r = pack_matrix(ctx, &ctx->p_c0, one_src((const bf16*)w->codebook0_head, ctx->V, Hb, 1), nullptr, ctx->V, Hb, 1, st)))
    return r;
  ctx->p_heads.resize(CSM_NQ - 1);
  for (int i = 0; i < CSM_NQ - 1; ++i) {
    // audio_head[i] is [in=Hd, out=V]: packed row n = output n, column k = input k -> strides (1, V)
    const bf16* base = (const bf16*)w->audio_head + (size_t)i * Hd * ctx->V;
    if ((r = pack_matrix(ctx, &ctx->p_heads[i], one_src(base, ctx->V, 1, ctx->V), nullptr, ctx->V, Hd, 1, st))) return r;
  }
  // ---- workspace
  const StackDims& b = ctx->bb.d;
  const StackDims& d = ctx->dec.d;
  const size_t B = max_batch;
  DA(ctx->kc_bb, (size_t)b.L * B * b.kv * max_ctx * b.hd);
  DA(ctx->vc_bb, (size_t)b.L * B * b.kv * max_ctx * b.hd);
  DA(ctx->kc_dec, (size_t)d.L * B * d.kv * CSM_DEC_POS * d.hd);
  DA(ctx->vc_dec, (size_t)d.L * B * d.kv * CSM_DEC_POS * d.hd);
  // every tagged vector exists in `repl` copies so that no L2 line is polled by all CTAs at once
  ctx->repl = 1;   // (replicating the vectors to spread the pollers was measured: no gain on B200)
  if (const char* e = getenv("CSM_EVICT_FIRST")) ctx->evict_first = atoi(e) != 0;
  const size_t R = ctx->repl;
  DA(ctx->h_bb, R * B * b.H); DA(ctx->h_dec, R * B * d.H);
  DA(ctx->q_bb, R * B * (b.heads + 2 * b.kv) * b.hd); DA(ctx->q_dec, R * B * (d.heads + 2 * d.kv) * d.hd);
  DA(ctx->attn_bb, R * B * b.heads * b.hd); DA(ctx->attn_dec, R * B * d.heads * d.hd);
  DA(ctx->mlp_bb, R * B * b.I); DA(ctx->mlp_dec, R * B * d.I);
  ctx->tagged = {{ctx->h_bb, R * B * b.H * 4}, {ctx->h_dec, R * B * d.H * 4},
                 {ctx->q_bb, R * B * (b.heads + 2 * b.kv) * b.hd * 4}, {ctx->q_dec, R * B * (d.heads + 2 * d.kv) * d.hd * 4},
                 {ctx->attn_bb, R * B * b.heads * b.hd * 4}, {ctx->attn_dec, R * B * d.heads * d.hd * 4},
                 {ctx->mlp_bb, R * B * b.I * 4}, {ctx->mlp_dec, R * B * d.I * 4}};
  DA(ctx->last_h, B * b.H); DA(ctx->c0_logits, B * ctx->V); DA(ctx->cb_logits, B * (CSM_NQ - 1) * ctx->V);
  ctx->nsplit_max = (max_ctx + CSM_ATT_SPLIT_MMA - 1) / CSM_ATT_SPLIT_MMA;   // (sized for the smaller unit of the two kernel families)
  DA(ctx->attn_part, B * b.heads * ctx->nsplit_max * (b.hd + 4));
  DA(ctx->attn_cnt, B * b.kv);
  DA(ctx->bar_counter, 4);
  ctx->lgt_stride = (ctx->V + 3) / 4 * 4;
  DA(ctx->lgt, B * ctx->lgt_stride);
  ctx->tagged.push_back({ctx->lgt, B * ctx->lgt_stride * sizeof(uint32_t)});
  DA(ctx->cand, R * ctx->sms * B);
  ctx->tagged.push_back({ctx->cand, R * ctx->sms * B * sizeof(unsigned long long)});
  for (auto& t : ctx->tagged) CK(cudaMemsetAsync(t.first, 0, t.second, st));
  ctx->tagbase = 1;
  if (const char* e = getenv("CSM_L2_AHEAD_KB")) ctx->l2_ahead = atoi(e) * 1024;
  DA(ctx->samples, B * CSM_NQ); DA(ctx->fed, B * CSM_NQ);
  DA(ctx->stop_flag, 4); DA(ctx->n_frames, 4);
  CK(cudaMemsetAsync(ctx->attn_cnt, 0, B * b.kv * sizeof(unsigned), st));
  CK(cudaMemsetAsync(ctx->bar_counter, 0, 16, st));
  CK(cudaMemsetAsync(ctx->stop_flag, 0, 16, st));
  CK(cudaMemsetAsync(ctx->n_frames, 0, 16, st));
  CK(cudaMemsetAsync(ctx->samples, 0, B * CSM_NQ * sizeof(int), st));
  CK(cudaMemsetAsync(ctx->fed, 0, B * CSM_NQ * sizeof(int), st));
  if (const char* e = getenv("CSM_MT2")) ctx->mt2 = atoi(e) != 0;
  build_table(ctx);
  if ((r = plan_smem(ctx))) return r;
  DA(ctx->d_table, ctx->table.size());
  DA(ctx->prof, (32 + (size_t)ctx->sms) * ctx->table.size());
  DA(ctx->abort_flag, 8);
  CK(cudaMemsetAsync(ctx->abort_flag, 0, 8 * sizeof(int), st));
  DA(ctx->progress, (size_t)ctx->sms * 4);
  CK(cudaMemsetAsync(ctx->progress, 0xff, (size_t)ctx->sms * 4 * sizeof(int), st));
  if (const char* e = getenv("CSM_DEBUG_PROGRESS")) ctx->progress_on = atoi(e) != 0;
  CK(cudaMemcpyAsync(ctx->d_table, ctx->table.data(), ctx->table.size() * sizeof(Phase), cudaMemcpyHostToDevice, st));
  CK(cudaEventCreate(&ctx->ev0));
  CK(cudaEventCreate(&ctx->ev1));
  CK(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
  if (const char* e = getenv("CSM_WAIT_BUDGET_S")) ctx->wait_budget_s = atof(e) > 0 ? atof(e) : ctx->wait_budget_s;
  CK(cudaStreamSynchronize(st));
  return CSM_OK;
}

int csm_destroy(CsmCtx* ctx) {
  if (!ctx) return CSM_OK;
  cudaDeviceSynchronize();
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->st_ids) cudaFree(ctx->st_ids);
  if (ctx->st_mask) cudaFree(ctx->st_mask);
  if (ctx->st_frames) cudaFree(ctx->st_frames);
  bf16* pf[] = {ctx->pf_h, ctx->pf_hn, ctx->pf_qkv, ctx->pf_attn, ctx->pf_act};
  for (bf16* b : pf) if (b) cudaFree(b);
  if (ctx->pf_valid) cudaFree(ctx->pf_valid);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  delete ctx;
  return CSM_OK;
}

int csm_reset(CsmCtx* ctx) {
  if (!ctx) return CSM_EINVAL;
  ctx->cache_len = 0;
  return CSM_OK;
}

int csm_cache_len(const CsmCtx* ctx) { return ctx ? ctx->cache_len : CSM_EINVAL; }

int csm_set_sampling(CsmCtx* ctx, int topk, float temperature, uint64_t seed, int seq_base) {
  if (!ctx) return CSM_EINVAL;
  if (topk <= 1 || temperature == 0.f) {   // greedy (the reference's spelling: topk=1)
    ctx->topk = 1;
    ctx->inv_temp = 1.f;
    return CSM_OK;
  }
  if (!(temperature > 0.f)) return fail(ctx, CSM_EINVAL, "temperature must be >= 0 (got %g)", (double)temperature);
  // sample_tokens keeps one 16-bit key per logit of every sequence in the activation region
  if ((size_t)ctx->Bmax * ctx->lgt_stride * 2 > (size_t)ctx->act_region)
    return fail(ctx, CSM_ECAPACITY, "top-k sampling of %d sequences needs %d bytes of shared memory (have %d)", ctx->Bmax,
                ctx->Bmax * ctx->lgt_stride * 2, ctx->act_region);
  ctx->topk = topk > ctx->V ? ctx->V : topk;
  ctx->inv_temp = 1.f / temperature;
  ctx->rng_seed = seed;
  ctx->rng_frame = 0;     // frames are counted from the call that set the seed: (seed, call) reproduces its output
  ctx->seq_base = seq_base;
  return CSM_OK;
}

int csm_embed_sum(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int S, void* out, void* stream) {
  if (!ctx || !ids || !out) return fail(ctx, CSM_EINVAL, "null argument");
  CK(csm_embed_sum_launch((const long long*)ids, mask, 1, ctx->audio_emb, ctx->text_emb, ctx->V, ctx->bb.d.H, (bf16*)out,
                          B * S, (cudaStream_t)stream));
  ctx->launches += 1;
  return CSM_OK;
}

int csm_generate_frame(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int S, const int64_t* force_tokens,
                       int64_t* samples, void* last_h, void* c0_logits, void* cb_logits, void* stream) {
  if (!ctx) return CSM_EINVAL;
  if (!ids) return fail(ctx, CSM_EINVAL, "input_ids is required");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(ctx->stop_flag, 0, sizeof(int), st));
  int r = frame_impl(ctx, (const long long*)ids, mask, B, S, (const long long*)force_tokens, nullptr, 0, 0, 0, st);
  if (r) return r;
  if (samples) CK(csm_i32_to_i64_launch(ctx->samples, (long long*)samples, B * CSM_NQ, st));
  if (last_h) CK(cudaMemcpyAsync(last_h, ctx->last_h, (size_t)B * ctx->bb.d.H * 2, cudaMemcpyDeviceToDevice, st));
  if (c0_logits) CK(cudaMemcpyAsync(c0_logits, ctx->c0_logits, (size_t)B * ctx->V * 2, cudaMemcpyDeviceToDevice, st));
  if (cb_logits)
    CK(cudaMemcpyAsync(cb_logits, ctx->cb_logits, (size_t)B * (CSM_NQ - 1) * ctx->V * 2, cudaMemcpyDeviceToDevice, st));
  return CSM_OK;
}

int csm_generate(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int T, int max_new_frames,
                 int stop_on_all_zeros, int64_t* frames, void* stream) {
  if (!ctx) return CSM_EINVAL;
  if (!ids || !frames) return fail(ctx, CSM_EINVAL, "null argument");
  if (max_new_frames < 0) return fail(ctx, CSM_EINVAL, "max_new_frames < 0");
  if (T < 1) return fail(ctx, CSM_EINVAL, "empty context");
  if (T + max_new_frames > ctx->Tcap + 1)
    return fail(ctx, CSM_ECAPACITY, "context %d + %d new frames exceeds max_ctx %d", T, max_new_frames, ctx->Tcap);
  cudaStream_t st = (cudaStream_t)stream;
  ctx->cache_len = 0;
  ctx->ev_frames = 0;
  CK(cudaMemsetAsync(ctx->stop_flag, 0, sizeof(int), st));
  CK(cudaMemsetAsync(ctx->n_frames, 0, sizeof(int), st));
  if (max_new_frames == 0) return CSM_OK;
  CK(cudaMemsetAsync(frames, 0, (size_t)B * max_new_frames * CSM_NQ * sizeof(int64_t), st));
  const long long stride = (long long)max_new_frames * CSM_NQ;
  int r = frame_impl(ctx, (const long long*)ids, mask, B, T, nullptr, (long long*)frames, stride, 0, stop_on_all_zeros, st);
  if (r) return r;
  CK(cudaEventRecord(ctx->ev0, st));
  for (int f = 1; f < max_new_frames; ++f) {
    // next input row = the 32 new ids + a zero text column, audio slots unmasked (modeling_csm.py:675-690)
    r = frame_impl(ctx, nullptr, nullptr, B, 1, nullptr, (long long*)frames, stride, (long long)f * CSM_NQ,
                   stop_on_all_zeros, st);
    if (r) return r;
    ctx->ev_frames += 1;
  }
  CK(cudaEventRecord(ctx->ev1, st));
  return CSM_OK;
}

int csm_generate_more(CsmCtx* ctx, int B, int n_more, int stop_on_all_zeros, int64_t* frames, void* stream) {
  if (!ctx) return CSM_EINVAL;
  if (!frames || n_more < 0) return fail(ctx, CSM_EINVAL, "bad argument");
  if (ctx->cache_len < 1) return fail(ctx, CSM_EINVAL, "no context to continue: call csm_generate first");
  if (ctx->cache_len + n_more > ctx->Tcap)
    return fail(ctx, CSM_ECAPACITY, "context %d + %d more frames exceeds max_ctx %d", ctx->cache_len, n_more, ctx->Tcap);
  cudaStream_t st = (cudaStream_t)stream;
  ctx->ev_frames = 0;
  CK(cudaMemsetAsync(ctx->stop_flag, 0, sizeof(int), st));
  CK(cudaMemsetAsync(ctx->n_frames, 0, sizeof(int), st));
  if (n_more == 0) return CSM_OK;
  CK(cudaMemsetAsync(frames, 0, (size_t)B * n_more * CSM_NQ * sizeof(int64_t), st));
  const long long stride = (long long)n_more * CSM_NQ;
  CK(cudaEventRecord(ctx->ev0, st));
  for (int f = 0; f < n_more; ++f) {
    // next input row = the 32 ids of the last frame + a zero text column (modeling_csm.py:675-690)
    int r = frame_impl(ctx, nullptr, nullptr, B, 1, nullptr, (long long*)frames, stride, (long long)f * CSM_NQ,
                       stop_on_all_zeros, st);
    if (r) return r;
    ctx->ev_frames += 1;
  }
  CK(cudaEventRecord(ctx->ev1, st));
  return CSM_OK;
}

int csm_frames_done(CsmCtx* ctx, void* stream) {
  if (!ctx) return CSM_EINVAL;
  int n = 0;
  CK(cudaMemcpyAsync(&n, ctx->n_frames, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  int r = check_abort(ctx, (cudaStream_t)stream);   // synchronises the stream
  if (r) return r;
  return n;
}

int csm_generate_host(CsmCtx* ctx, const int64_t* ids_host, const int32_t* mask_host, int B, int T, int max_new_frames,
                      int stop_on_all_zeros, int64_t* frames_host, int* n_out, void* stream) {
  if (!ctx) return CSM_EINVAL;
  if (!ids_host || !frames_host) return fail(ctx, CSM_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_in = (size_t)B * T * (CSM_NQ + 1);
  const size_t n_fr = (size_t)B * (max_new_frames > 0 ? max_new_frames : 1) * CSM_NQ;
  if (n_in > ctx->st_ids_n) {
    if (ctx->st_ids) cudaFree(ctx->st_ids);
    if (ctx->st_mask) cudaFree(ctx->st_mask);
    ctx->st_ids = nullptr; ctx->st_mask = nullptr; ctx->st_ids_n = 0;
    CK(cudaMalloc(&ctx->st_ids, n_in * sizeof(long long)));
    CK(cudaMalloc(&ctx->st_mask, n_in * sizeof(int)));
    ctx->st_ids_n = n_in;
  }
  if (n_fr > ctx->st_frames_n) {
    if (ctx->st_frames) cudaFree(ctx->st_frames);
    ctx->st_frames = nullptr; ctx->st_frames_n = 0;
    CK(cudaMalloc(&ctx->st_frames, n_fr * sizeof(long long)));
    ctx->st_frames_n = n_fr;
  }
  CK(cudaMemcpyAsync(ctx->st_ids, ids_host, n_in * sizeof(long long), cudaMemcpyHostToDevice, st));
  if (mask_host) CK(cudaMemcpyAsync(ctx->st_mask, mask_host, n_in * sizeof(int), cudaMemcpyHostToDevice, st));
  int r = csm_generate(ctx, (const int64_t*)ctx->st_ids, mask_host ? ctx->st_mask : nullptr, B, T, max_new_frames,
                       stop_on_all_zeros, (int64_t*)ctx->st_frames, st);
  if (r) return r;
  if (max_new_frames > 0)
    CK(cudaMemcpyAsync(frames_host, ctx->st_frames, (size_t)B * max_new_frames * CSM_NQ * sizeof(long long),
                       cudaMemcpyDeviceToHost, st));
  int n = csm_frames_done(ctx, st);
  if (n < 0) return n;
  if (n_out) *n_out = n;
  return CSM_OK;
}

int64_t csm_info(const CsmCtx* ctx, int what) {
  if (!ctx) return -1;
  switch (what) {
    case CSM_INFO_SMS: return ctx->sms;
    case CSM_INFO_GRID: return ctx->G;
    case CSM_INFO_PHASES_PER_FRAME: return (int64_t)ctx->table.size();
    case CSM_INFO_SMEM_BYTES: return (int64_t)ctx->smem_total;
    case CSM_INFO_LAUNCHES: return ctx->launches;
    case CSM_INFO_STEPPED: return ctx->stepped;
  }
  return -1;
}

int csm_set_stepped(CsmCtx* ctx, int stepped) {
  if (!ctx) return CSM_EINVAL;
  ctx->stepped = stepped ? 1 : 0;
  return CSM_OK;
}

int csm_last_decode_ms(CsmCtx* ctx, float* ms, int* n) {
  if (!ctx || !ms || !n) return CSM_EINVAL;
  *ms = 0.f;
  *n = ctx->ev_frames;
  if (ctx->ev_frames <= 0) return CSM_OK;
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return CSM_OK;
}

const char* csm_last_error(const CsmCtx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// ---- debug / test hooks (see include/csm_b200.h)
int csm_debug_copy(CsmCtx* ctx, int which, void* dst_device, int64_t max_bytes, int64_t* bytes_out, void* stream) {
  // buffers 0..7 are tagged words inside the engine; they are returned as plain bf16 rows [Bmax][cols]
  if (!ctx) return CSM_EINVAL;
  const StackDims& b = ctx->bb.d;
  const StackDims& d = ctx->dec.d;
  const size_t B = ctx->Bmax;
  const void* src = nullptr;
  size_t n = 0;
  const uint32_t* tsrc = nullptr;
  long long tstride = 0;
  int tcols = 0;
  const int gpad = ctx->fuse_attn ? 0 : 8;   // general kernels: plain bf16 rows, K + 8 apart (buffers 6/7: tiled, raw)
  switch (which) {
    case 0: tsrc = ctx->h_bb; tcols = b.H; tstride = b.H + gpad; break;
    case 1: tsrc = ctx->h_dec; tcols = d.H; tstride = d.H + gpad; break;
    case 2: tsrc = ctx->q_bb; tcols = b.heads * b.hd; tstride = (b.heads + 2 * b.kv) * b.hd; break;
    case 3: tsrc = ctx->q_dec; tcols = d.heads * d.hd; tstride = (d.heads + 2 * d.kv) * d.hd; break;
    case 4: tsrc = ctx->attn_bb; tcols = b.heads * b.hd; tstride = tcols + gpad; break;
    case 5: tsrc = ctx->attn_dec; tcols = d.heads * d.hd; tstride = tcols + gpad; break;
    case 6: tsrc = ctx->mlp_bb; tcols = b.I; tstride = b.I; break;
    case 7: tsrc = ctx->mlp_dec; tcols = d.I; tstride = d.I; break;
    case 8: src = ctx->last_h; n = B * b.H * 2; break;
    case 9: src = ctx->c0_logits; n = B * ctx->V * 2; break;
    case 10: src = ctx->cb_logits; n = B * (CSM_NQ - 1) * ctx->V * 2; break;
    case 11: src = ctx->samples; n = B * CSM_NQ * 4; break;
    case 12: src = ctx->fed; n = B * CSM_NQ * 4; break;
    case 13: src = ctx->kc_bb; n = (size_t)b.L * B * b.kv * ctx->Tcap * b.hd * 2; break;
    case 14: src = ctx->vc_bb; n = (size_t)b.L * B * b.kv * ctx->Tcap * b.hd * 2; break;
    case 15: src = ctx->kc_dec; n = (size_t)d.L * B * d.kv * CSM_DEC_POS * d.hd * 2; break;
    case 16: src = ctx->vc_dec; n = (size_t)d.L * B * d.kv * CSM_DEC_POS * d.hd * 2; break;
    default: return fail(ctx, CSM_EINVAL, "unknown debug buffer %d", which);
  }
  if (tsrc) n = B * (size_t)tcols * 2;
  if (bytes_out) *bytes_out = (int64_t)n;
  if (!dst_device) return CSM_OK;
  if (tsrc) {
    if ((int64_t)n > max_bytes) return fail(ctx, CSM_EINVAL, "debug buffer %d needs %lld bytes", which, (long long)n);
    if (ctx->fuse_attn) {
      CK(csm_untag_rows_launch(tsrc, tstride, tcols, (int)B, (bf16*)dst_device, (cudaStream_t)stream));
    } else {   // general kernels: the same buffers hold plain bf16 rows
      CK(cudaMemcpy2DAsync(dst_device, (size_t)tcols * 2, tsrc, (size_t)tstride * 2, (size_t)tcols * 2, B,
                           cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return CSM_OK;
  }
  if ((int64_t)n > max_bytes) n = (size_t)max_bytes;
  CK(cudaMemcpyAsync(dst_device, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CSM_OK;
}

int csm_debug_run_phases(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int ph_begin, int ph_end,
                         int forced, void* stream) {
  if (!ctx) return CSM_EINVAL;
  if (ph_begin < 0 || ph_end > (int)ctx->table.size() || ph_begin >= ph_end)
    return fail(ctx, CSM_EINVAL, "bad phase range [%d,%d)", ph_begin, ph_end);
  int r = begin_epoch(ctx, (cudaStream_t)stream);
  if (r) return r;
  r = launch_frame(ctx, B, ph_begin, ph_end, (const long long*)ids, mask, forced, nullptr, 0, 0, 0, ctx->cache_len,
                   (cudaStream_t)stream);
  if (r == 0 && ph_end == (int)ctx->table.size()) end_epoch(ctx);   // phases of one frame share one tag epoch
  return r;
}

int csm_debug_profile_frame(CsmCtx* ctx, int B, uint64_t* clocks_host, int32_t* info_host, void* stream) {
  // one decode frame (ids from the last sampled frame) with per-phase clock64 stamps of the first and last CTA
  if (!ctx || !clocks_host) return CSM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = ctx->table.size();
  const size_t nprof = (32 + (size_t)ctx->G) * n;   // [2 CTAs][n][16] clock64 stamps + [G][n] globaltimer at phase end
  CK(cudaMemsetAsync(ctx->prof, 0, nprof * sizeof(unsigned long long), st));
  int r = begin_epoch(ctx, st);
  if (r) return r;
  ctx->prof_on = 1;
  r = launch_frame(ctx, B, 0, (int)n, nullptr, nullptr, 0, nullptr, 0, 0, 0, ctx->cache_len, st);
  ctx->prof_on = 0;
  if (r) return r;
  end_epoch(ctx);
  CK(cudaMemcpyAsync(clocks_host, ctx->prof, nprof * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (info_host)
    for (size_t i = 0; i < n; ++i) {
      const Phase& P = ctx->table[i];
      info_host[4 * i + 0] = P.type; info_host[4 * i + 1] = P.type == PH_GEMV ? P.epi : -1;
      info_host[4 * i + 2] = P.stack; info_host[4 * i + 3] = P.type == PH_GEMV ? P.act_mode : -1;
    }
  return CSM_OK;
}

int csm_sample_topk(const void* logits, int rows, int V, int topk, float temperature, uint64_t seed, int64_t* out,
                    void* stream) {
  if (!logits || !out || rows < 0 || V < 1 || topk < 1 || !(temperature > 0.f)) return CSM_EINVAL;
  if (rows == 0) return CSM_OK;
  cudaError_t e = csm_sample_rows_launch((const bf16*)logits, rows, V, topk, 1.f / temperature, seed, (long long*)out,
                                         (cudaStream_t)stream);
  return e == cudaSuccess ? CSM_OK : CSM_ECUDA;
}

int csm_linear(const void* x, int ldx, const void* W, int R, int N, int K, int tail, void* C, int ldc, void* stream) {
  if (!x || !W || !C || R < 0 || N < 64 || K < 64 || (N % 64) || (K % 64) || tail < 0 || tail > 2) return CSM_EINVAL;
  if (R == 0) return CSM_OK;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return CSM_ECUDA;
  CUtensorMap ma, mw;
  if (csm_tmap_2d(&ma, x, R, K, ldx, csm_gemm_box_rows_a()) || csm_tmap_2d(&mw, W, N, K, K, csm_gemm_box_rows_w()))
    return CSM_ECUDA;
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.R = R; g.N = N; g.K = K; g.epi = tail == 0 ? EPI_STORE : (tail == 1 ? EPI_RESID : EPI_SWIGLU);
  g.C = (bf16*)C; g.ldc = ldc;
  return csm_gemm_launch(&ma, &mw, &g, sms, (cudaStream_t)stream) == cudaSuccess ? CSM_OK : CSM_ECUDA;
}

int csm_debug_progress(CsmCtx* ctx, int32_t* host_out, void* side_stream) {
  // copy the [grid][4] progress words on `side_stream` (a non-blocking stream: works while a frame kernel hangs)
  if (!ctx || !host_out) return CSM_EINVAL;
  CK(cudaMemcpyAsync(host_out, ctx->progress, (size_t)ctx->G * 4 * sizeof(int), cudaMemcpyDeviceToHost,
                     (cudaStream_t)side_stream));
  CK(cudaStreamSynchronize((cudaStream_t)side_stream));
  return CSM_OK;
}

int csm_debug_set_cache_len(CsmCtx* ctx, int len) {
  if (!ctx) return CSM_EINVAL;
  ctx->cache_len = len;
  return CSM_OK;
}

}  // extern "C"
