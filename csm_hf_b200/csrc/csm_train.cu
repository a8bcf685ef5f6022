// Training step of CSM on B200: CSMModel.forward(labels=...) and its backward (reference modeling_csm.py:292-482,
// driven by CSMTrainer.compute_loss, train.py:303-326) -- SURVEY.md section 8(f) row N1.
//
//   csm_train_step: ids, mask, labels -> loss, backbone_loss, decoder_loss and (optionally) the gradient of `loss`
//   with respect to every parameter, in the reference's parameter layout.
//
// Every dense projection -- forward, input gradient and weight gradient -- runs on the tcgen05 GEMM of csm_gemm.cu
// (C[R,N] = A[R,K] W[N,K]^T, both operands K-major through TMA tensor maps): the input gradient dX = dY W uses a
// transposed copy of W made once per step, the weight gradient dW = dY^T X uses transposed copies of dY and X (the
// contraction then runs over the token dimension; a ragged last k-block is zero-filled by TMA).  Attention forward and
// backward are flash kernels on mma.sync (csm_train_kernels.cuh); RMSNorm, RoPE, SwiGLU, cross entropy, the decoder's
// gather / scatter and the embedding-table gradients are bandwidth kernels.  q|k|v and gate|up are fused per step into
// one weight matrix each (copies of the caller's tensors -- the caller's parameters change between steps).
//
// Numerics: bf16 tensors with the reference's rounding points in the forward (linear outputs, RMSNorm, RoPE, SwiGLU,
// residuals), fp32 accumulation inside every contraction, gradients rounded to bf16 where autograd on bf16 tensors
// rounds them (every tensor boundary); embedding-table and norm-weight gradients are summed in fp32 and rounded once.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/csm_b200.h"
#include "csm_train_kernels.cuh"

extern "C" {
cudaError_t csm_gemm_launch(const void* map_a, const void* map_w, const GemmParams* p, int sms, cudaStream_t st);
int csm_tmap_2d(void* out, const void* base, long long rows, int K, long long pitch, int box_rows);
int csm_tmap_2d_mn(void* out, const void* base, long long k_rows, long long mn, long long pitch);
int csm_gemm_box_rows_a();
int csm_gemm_box_rows_w();
cudaError_t csm_embed_sum_launch(const long long* ids, const int* mask, int default_mask, const bf16* audio_emb,
                                 const bf16* text_emb, int V, int H, bf16* out, int rows, cudaStream_t st);
cudaError_t csm_rmsnorm_rows_launch(const bf16* x, const bf16* w, float eps, int H, bf16* y, int rows, cudaStream_t st);
cudaError_t csm_frame_valid_launch(const int* mask, int rows, unsigned char* valid, int* any_pad, cudaStream_t st);
cudaError_t csm_flash_tc_launch(const bf16* qkv, int S, int nseq, int heads, int kv, float scale, const unsigned char* valid,
                                bf16* out, float* lse, cudaStream_t st);
cudaError_t csm_flash_tc_bwd_launch(const bf16* qkv, const bf16* d_out, const float* lse, const float* delta, int S, int nseq,
                                    int heads, int kv, float scale, const unsigned char* valid, bf16* dqkv, float* dq_acc,
                                    cudaStream_t st);
}

namespace {

struct TLayer {
  const bf16 *q, *k, *v, *o, *gate, *up, *down, *ln1, *ln2;          // the caller's parameters
  bf16 *gq, *gk, *gv, *go, *ggate, *gup, *gdown, *gln1, *gln2;       // the caller's gradient tensors (may be null)
  bf16 *Wqkv, *WqkvT, *WoT, *Wgu, *WguT, *WdownT;                    // per-step fused / transposed copies
  bf16 *h_in, *hn1, *qkv, *attn, *h_mid, *hn2, *gu, *act;            // saved activations [rows, .]
  float* lse;                                                        // [rows, heads]
};

struct TStack {
  int H = 0, I = 0, L = 0, heads = 0, kv = 0, hd = 0, W = 0, nq = 0, n_pos = 0;
  float eps = 0.f, scale = 0.f;
  bf16 *cos_t = nullptr, *sin_t = nullptr;
  const bf16* norm = nullptr;
  bf16* gnorm = nullptr;
  bf16 *h_out = nullptr, *hf = nullptr;   // residual stream after the last layer, and its final RMSNorm
  std::vector<TLayer> layers;
  int max_rows = 0;
};

}  // namespace

struct CsmTrain {
  std::string err;
  int sms = 148;
  int V = 0, Vp = 0, text_vocab = 0;
  int max_tokens = 0, max_frames = 0;
  TStack bb, dec;
  std::vector<void*> allocs;
  // scratch shared by both stacks (sized for the larger need)
  bf16 *tA = nullptr, *tB = nullptr;                  // transposed operands of the weight-gradient GEMMs
  bf16 *dAct = nullptr, *dGU = nullptr, *dHn = nullptr, *dAttn = nullptr, *dQKV = nullptr, *dWtmp = nullptr;
  float *dq_acc = nullptr, *delta = nullptr, *dw_acc = nullptr;
  bf16 *dh_bb = nullptr, *dh_dec = nullptr;            // residual-stream gradients
  // heads and decoder plumbing
  bf16 *Wc0p = nullptr, *Wc0pT = nullptr, *logits0 = nullptr;            // [Vp,Hb], [Hb,Vp], [tokens,Vp]
  bf16 *AHt = nullptr, *AHp = nullptr, *logits_d = nullptr;              // [31][Vp,Hd], [31][Hd,Vp], [31][frames,Vp]
  bf16 *WprojT = nullptr, *dec_in = nullptr, *d_dec_in = nullptr, *dhdf = nullptr;
  unsigned char *valid = nullptr, *fflag = nullptr;
  int *frames = nullptr, *counts = nullptr /* [0] frames, [1] backbone rows, [2] decoder rows */, *lab0 = nullptr, *labd = nullptr;
  float *row_loss = nullptr, *losses = nullptr;        // device [3]
  float *audio_acc = nullptr, *text_acc = nullptr;     // fp32 gradients of the embedding tables
  int launches = 0;
  // a step split in two calls (csm_train_step_begin / _end): what the second half needs
  struct {
    bool active = false, want = false;
    int B = 0, S = 0, split = 0;
    const long long* ids = nullptr;
    const int* mask = nullptr;
    const unsigned char* valid = nullptr;
    bf16 *g_audio = nullptr, *g_text = nullptr;
  } pend;
  bool flash_tc = true;      // head-dim-64 attention forward on tcgen05
  bool flash_tc_bwd = true;  // ... and backward
  bool mn_operands = true;   // gradients read W, dY and X as MN-major tcgen05 operands (CSM_TRAIN_TRANSPOSE=1: transposed copies)
  std::map<std::string, std::pair<const void*, size_t>> dbg;   // named intermediates of the last step (tests)
};

namespace {

int tfail(CsmTrain* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define TCK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return tfail(t, CSM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(call)          \
  do {                     \
    int r_ = (call);       \
    if (r_) return r_;     \
  } while (0)

template <typename T>
int talloc(CsmTrain* t, T** p, size_t n) {
  void* q = nullptr;
  size_t bytes = n * sizeof(T);
  if (bytes == 0) bytes = 16;
  TCK(cudaMalloc(&q, bytes));
  t->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}
inline int rup(int x, int m) { return (x + m - 1) / m * m; }
inline int nblocks(long long n, int per = 256, int cap = 148 * 16) {
  long long b = (n + per - 1) / per;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// C[R, N] (pitch ldc) = A[R, K] (pitch lda) * Wm[N, K]^T (pitch ldw);  epi: EPI_STORE or EPI_RESID (C += ...)
int gemm(CsmTrain* t, const bf16* A, long long lda, int R, int K, const bf16* Wm, long long ldw, int N, bf16* C, int ldc,
         int epi, cudaStream_t st) {
  if (R <= 0) return 0;
  CUtensorMap ma, mw;
  if (csm_tmap_2d(&ma, A, R, K, lda, csm_gemm_box_rows_a()) || csm_tmap_2d(&mw, Wm, N, K, ldw, csm_gemm_box_rows_w()))
    return tfail(t, CSM_ECUDA, "cuTensorMapEncodeTiled failed for a [%d,%d] x [%d,%d] product (pitches %lld, %lld)", R, K, N,
                 K, lda, ldw);
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.R = R; g.N = N; g.K = K; g.epi = epi; g.C = C; g.ldc = ldc;
  TCK(csm_gemm_launch(&ma, &mw, &g, t->sms, st));
  t->launches += 1;
  return 0;
}
// Input gradient: dX[R, K] = dY[R, N] W[N, K] -- W in its natural [N, K] layout is the MN-major B operand of a product
// whose contraction runs over N: no transposed weight copy.
int gemm_dgrad(CsmTrain* t, const bf16* dY, long long ldy, int R, int N, const bf16* Wm, long long ldw, int K, bf16* dX, int ldx,
               cudaStream_t st) {
  if (R <= 0) return 0;
  CUtensorMap ma, mw;
  if (csm_tmap_2d(&ma, dY, R, N, ldy, csm_gemm_box_rows_a()) || csm_tmap_2d_mn(&mw, Wm, N, K, ldw))
    return tfail(t, CSM_ECUDA, "cuTensorMapEncodeTiled failed for an input-gradient product [%d,%d] x [%d,%d]", R, N, N, K);
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.R = R; g.N = K; g.K = N; g.epi = EPI_STORE; g.C = dX; g.ldc = ldx; g.b_mn = 1;
  TCK(csm_gemm_launch(&ma, &mw, &g, t->sms, st));
  t->launches += 1;
  return 0;
}
// Weight gradient: dW[N, K] = dY[R, N]^T X[R, K] -- dY and X in their natural row-per-token layouts are both MN-major
// operands of a product whose contraction runs over the R tokens (a ragged last block of tokens is zero-filled by TMA).
int gemm_wgrad(CsmTrain* t, const bf16* dY, long long ldy, const bf16* X, long long ldx, int R, int N, int K, bf16* dW, int ldw,
               cudaStream_t st) {
  if (R <= 0) return 0;
  CUtensorMap ma, mw;
  if (csm_tmap_2d_mn(&ma, dY, R, N, ldy) || csm_tmap_2d_mn(&mw, X, R, K, ldx))
    return tfail(t, CSM_ECUDA, "cuTensorMapEncodeTiled failed for a weight-gradient product [%d,%d]^T x [%d,%d]", R, N, R, K);
  GemmParams g;
  memset(&g, 0, sizeof g);
  g.R = N; g.N = K; g.K = R; g.epi = EPI_STORE; g.C = dW; g.ldc = ldw; g.a_mn = 1; g.b_mn = 1;
  TCK(csm_gemm_launch(&ma, &mw, &g, t->sms, st));
  t->launches += 1;
  return 0;
}
// dX = dY W with W [N, K] natural; WT = its transposed copy [K, N] for the CSM_TRAIN_TRANSPOSE=1 path
int gemm(CsmTrain* t, const bf16* A, long long lda, int R, int K, const bf16* Wm, long long ldw, int N, bf16* C, int ldc,
         int epi, cudaStream_t st);
int dgrad(CsmTrain* t, const bf16* dY, long long ldy, int R, int N, const bf16* Wm, long long ldw, int K, const bf16* WT,
          bf16* dX, int ldx, cudaStream_t st) {
  if (t->mn_operands) return gemm_dgrad(t, dY, ldy, R, N, Wm, ldw, K, dX, ldx, st);
  return gemm(t, dY, ldy, R, N, WT, N, K, dX, ldx, EPI_STORE, st);
}
// dst [cols, rows] (pitch ldd) = src [rows, cols] (pitch lds) transposed
int transpose(CsmTrain* t, const bf16* src, int rows, int cols, long long lds, bf16* dst, long long ldd, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((cols + 63) / 64, (rows + 63) / 64), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(src, rows, cols, lds, dst, ldd);
  TCK(cudaGetLastError());
  t->launches += 1;
  return 0;
}
// dW[N, K] = dY[R, N]^T X[R, K]: both operands transposed into tA / tB so that the contraction (over the R rows) is the
// K-major dimension of the GEMM
int wgrad(CsmTrain* t, const bf16* dY, long long ldy, const bf16* X, long long ldx, int R, int N, int K, bf16* dW, int ldw,
          cudaStream_t st) {
  if (t->mn_operands) return gemm_wgrad(t, dY, ldy, X, ldx, R, N, K, dW, ldw, st);
  const int Rp = rup(R, 8);
  TRY(transpose(t, dY, R, N, ldy, t->tA, Rp, st));
  TRY(transpose(t, X, R, K, ldx, t->tB, Rp, st));
  return gemm(t, t->tA, Rp, N, R, t->tB, Rp, K, dW, ldw, EPI_STORE, st);
}
int copy_rows(CsmTrain* t, bf16* dst, const bf16* src, size_t n, cudaStream_t st) {
  TCK(cudaMemcpyAsync(dst, src, n * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// named intermediate for csm_train_debug; CSM_TRAIN_SNAP=1 keeps a copy taken now (buffers are reused later in the step)
int note(CsmTrain* t, const std::string& name, const void* p, size_t bytes, cudaStream_t st) {
  static const bool snap = getenv("CSM_TRAIN_SNAP") != nullptr;
  if (snap && bytes) {
    void* q = nullptr;
    TCK(cudaMalloc(&q, bytes));
    t->allocs.push_back(q);
    TCK(cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice, st));
    p = q;
  }
  t->dbg[name] = {p, bytes};
  return 0;
}

template <int HD>
int flash_fwd(CsmTrain* t, const TStack& s, const bf16* qkv, int S, int nseq, const unsigned char* valid, bf16* out, float* lse,
              cudaStream_t st) {
  const size_t smem = (size_t)4 * 64 * (HD + 8) * 2 + 128;
  TCK(cudaFuncSetAttribute((const void*)flash_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TCK(cudaFuncSetAttribute((const void*)flash_fwd_kernel<HD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared));
  dim3 grid((S + 127) / 128, s.heads, nseq);
  flash_fwd_kernel<HD><<<grid, 256, smem, st>>>(qkv, S, s.heads, s.kv, s.scale, valid, out, lse);
  TCK(cudaGetLastError());
  t->launches += 1;
  return 0;
}
template <int HD>
int flash_bwd(CsmTrain* t, const TStack& s, const bf16* qkv, const bf16* d_out, const float* lse, const float* delta, int S,
              int nseq, const unsigned char* valid, bf16* dqkv, float* dq_acc, cudaStream_t st) {
  const size_t smem = (size_t)6 * 64 * (HD + 8) * 2 + (size_t)64 * 72 * 2 + 4 * 64 * 4 + 64;
  TCK(cudaFuncSetAttribute((const void*)flash_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TCK(cudaFuncSetAttribute((const void*)flash_bwd_kernel<HD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared));
  dim3 grid((S + 63) / 64, s.kv, nseq);
  flash_bwd_kernel<HD><<<grid, 128, smem, st>>>(qkv, d_out, lse, delta, S, s.heads, s.kv, s.scale, valid, dqkv, dq_acc);
  TCK(cudaGetLastError());
  t->launches += 1;
  return 0;
}

// hf LlamaModel.forward over whole sequences (modeling_llama.py:375-425 without a cache): rows = nseq * S, the input
// is layers[0].h_in; leaves the residual stream in s.h_out and its final RMSNorm in s.hf; saves what backward needs.
int stack_forward(CsmTrain* t, TStack& s, int S, int nseq, const unsigned char* valid, cudaStream_t st) {
  const int R = nseq * S;
  for (int l = 0; l < s.L; ++l) {
    TLayer& y = s.layers[l];
    TCK(csm_rmsnorm_rows_launch(y.h_in, y.ln1, s.eps, s.H, y.hn1, R, st));
    TRY(gemm(t, y.hn1, s.H, R, s.H, y.Wqkv, s.H, s.W, y.qkv, s.W, EPI_STORE, st));
    rope_rows_kernel<false><<<nblocks((long long)R * (s.heads + s.kv) * (s.hd / 16)), 256, 0, st>>>(
        y.qkv, s.W, R, S, s.heads + s.kv, s.hd, s.cos_t, s.sin_t);
    TCK(cudaGetLastError());
    if (s.hd == 64 && t->flash_tc) {   // tcgen05 forward (csm_flash_tc.cu); CSM_FLASH_MMA=1: the mma.sync kernel
      TCK(csm_flash_tc_launch(y.qkv, S, nseq, s.heads, s.kv, s.scale, valid, y.attn, y.lse, st));
      t->launches += 1;
    } else if (s.hd == 64) {
      TRY(flash_fwd<64>(t, s, y.qkv, S, nseq, valid, y.attn, y.lse, st));
    } else {
      TRY(flash_fwd<128>(t, s, y.qkv, S, nseq, valid, y.attn, y.lse, st));
    }
    TRY(copy_rows(t, y.h_mid, y.h_in, (size_t)R * s.H, st));
    TRY(gemm(t, y.attn, s.nq, R, s.nq, y.o, s.nq, s.H, y.h_mid, s.H, EPI_RESID, st));
    TCK(csm_rmsnorm_rows_launch(y.h_mid, y.ln2, s.eps, s.H, y.hn2, R, st));
    TRY(gemm(t, y.hn2, s.H, R, s.H, y.Wgu, s.H, 2 * s.I, y.gu, 2 * s.I, EPI_STORE, st));
    swiglu_fwd_kernel<<<nblocks((long long)R * s.I / 8), 256, 0, st>>>(y.gu, R, s.I, y.act);
    TCK(cudaGetLastError());
    bf16* nxt = l + 1 < s.L ? s.layers[l + 1].h_in : s.h_out;
    TRY(copy_rows(t, nxt, y.h_mid, (size_t)R * s.H, st));
    TRY(gemm(t, y.act, s.I, R, s.I, y.down, s.I, s.H, nxt, s.H, EPI_RESID, st));
    t->launches += 4;
  }
  TCK(csm_rmsnorm_rows_launch(s.h_out, s.norm, s.eps, s.H, s.hf, R, st));
  t->launches += 1;
  return 0;
}

// fp32 accumulator of a norm-weight gradient -> the caller's bf16 tensor
int norm_bwd(CsmTrain* t, const TStack& s, const bf16* x, const bf16* w, const bf16* dy, const bf16* dres, bf16* dh, bf16* gw,
             int R, cudaStream_t st) {
  TCK(cudaMemsetAsync(t->dw_acc, 0, (size_t)s.H * 4, st));
  const int rpb = 16;   // rows per block (8 warps x 2 rows): enough blocks to fill the GPU at a few thousand rows
  rmsnorm_bwd_kernel<<<(R + rpb - 1) / rpb, 256, 0, st>>>(x, w, dy, dres, s.eps, s.H, R, rpb, dh, t->dw_acc);
  TCK(cudaGetLastError());
  if (gw) {
    f32_to_bf16_kernel<<<nblocks(s.H / 2), 256, 0, st>>>(t->dw_acc, s.H, gw);
    TCK(cudaGetLastError());
  }
  t->launches += 2;
  return 0;
}

// Adjoint of stack_forward.  dh: gradient w.r.t. the residual stream after the last layer on entry (the final norm's
// adjoint has been applied by the caller), w.r.t. the stack's input on return.  Writes the layers' weight gradients.
int stack_backward(CsmTrain* t, TStack& s, const char* tag, int S, int nseq, const unsigned char* valid, bf16* dh,
                   bool want_grads, cudaStream_t st, int l_hi = -1, int l_lo = 0) {
  const int R = nseq * S;
  if (l_hi < 0) l_hi = s.L - 1;
  for (int l = l_hi; l >= l_lo; --l) {
    TLayer& y = s.layers[l];
    const std::string pre = std::string("d.") + tag + "." + std::to_string(l) + ".";
    TRY(note(t, pre + "h_out", dh, (size_t)R * s.H * 2, st));
    // ---- MLP: h_out = h_mid + down(act), act = silu(gate) * up, gate|up = Wgu hn2, hn2 = norm(h_mid)
    TRY(dgrad(t, dh, s.H, R, s.H, y.down, s.I, s.I, y.WdownT, t->dAct, s.I, st));
    if (want_grads) TRY(wgrad(t, dh, s.H, y.act, s.I, R, s.H, s.I, y.gdown, s.I, st));
    swiglu_bwd_kernel<<<nblocks((long long)R * s.I / 8), 256, 0, st>>>(y.gu, t->dAct, R, s.I, t->dGU);
    TCK(cudaGetLastError());
    TRY(dgrad(t, t->dGU, 2 * s.I, R, 2 * s.I, y.Wgu, s.H, s.H, y.WguT, t->dHn, s.H, st));
    if (want_grads) {
      // (gate | up gradients adjacent in the caller's buffer -- the flat layout of training.py: written in place)
      bf16* dgu = y.gup == y.ggate + (size_t)s.I * s.H ? y.ggate : t->dWtmp;
      TRY(wgrad(t, t->dGU, 2 * s.I, y.hn2, s.H, R, 2 * s.I, s.H, dgu, s.H, st));
      if (dgu == t->dWtmp) {
        TRY(copy_rows(t, y.ggate, t->dWtmp, (size_t)s.I * s.H, st));
        TRY(copy_rows(t, y.gup, t->dWtmp + (size_t)s.I * s.H, (size_t)s.I * s.H, st));
      }
    }
    TRY(note(t, pre + "act", t->dAct, (size_t)R * s.I * 2, st));
    TRY(note(t, pre + "hn2", t->dHn, (size_t)R * s.H * 2, st));
    TRY(norm_bwd(t, s, y.h_mid, y.ln2, t->dHn, dh, dh, want_grads ? y.gln2 : nullptr, R, st));
    TRY(note(t, pre + "h_mid", dh, (size_t)R * s.H * 2, st));
    // ---- attention: h_mid = h_in + o(attn), attn = sdpa(rope(q), rope(k), v), q|k|v = Wqkv hn1, hn1 = norm(h_in)
    TRY(dgrad(t, dh, s.H, R, s.H, y.o, s.nq, s.nq, y.WoT, t->dAttn, s.nq, st));
    if (want_grads) TRY(wgrad(t, dh, s.H, y.attn, s.nq, R, s.H, s.nq, y.go, s.nq, st));
    attn_delta_kernel<<<nblocks((long long)R * s.heads * 32, 256, 1 << 30), 256, 0, st>>>(y.attn, t->dAttn, R, s.heads, s.hd,
                                                                                            t->delta);
    TCK(cudaGetLastError());
    TCK(cudaMemsetAsync(t->dq_acc, 0, (size_t)R * s.nq * 4, st));
    if (s.hd == 64 && t->flash_tc_bwd) {   // tcgen05 backward (csm_flash_tc_bwd.cu); CSM_FLASH_BWD_MMA=1: the mma.sync kernel
      TCK(csm_flash_tc_bwd_launch(y.qkv, t->dAttn, y.lse, t->delta, S, nseq, s.heads, s.kv, s.scale, valid, t->dQKV, t->dq_acc, st));
      t->launches += 1;
    } else if (s.hd == 64) TRY(flash_bwd<64>(t, s, y.qkv, t->dAttn, y.lse, t->delta, S, nseq, valid, t->dQKV, t->dq_acc, st));
    else TRY(flash_bwd<128>(t, s, y.qkv, t->dAttn, y.lse, t->delta, S, nseq, valid, t->dQKV, t->dq_acc, st));
    f32_to_bf16_rows_kernel<<<nblocks((long long)R * s.nq / 2), 256, 0, st>>>(t->dq_acc, R, s.nq, t->dQKV, s.W);
    TCK(cudaGetLastError());
    rope_rows_kernel<true><<<nblocks((long long)R * (s.heads + s.kv) * (s.hd / 16)), 256, 0, st>>>(
        t->dQKV, s.W, R, S, s.heads + s.kv, s.hd, s.cos_t, s.sin_t);
    TCK(cudaGetLastError());
    TRY(note(t, pre + "attn", t->dAttn, (size_t)R * s.nq * 2, st));
    TRY(note(t, pre + "qkv_raw", t->dQKV, (size_t)R * s.W * 2, st));   // gradient w.r.t. the un-rotated q | k | v
    TRY(dgrad(t, t->dQKV, s.W, R, s.W, y.Wqkv, s.H, s.H, y.WqkvT, t->dHn, s.H, st));
    TRY(note(t, pre + "hn1", t->dHn, (size_t)R * s.H * 2, st));
    if (want_grads) {
      const size_t kvw = (size_t)s.kv * s.hd;
      bf16* dqkvw = (y.gk == y.gq + (size_t)s.nq * s.H && y.gv == y.gk + kvw * s.H) ? y.gq : t->dWtmp;   // (see gate | up)
      TRY(wgrad(t, t->dQKV, s.W, y.hn1, s.H, R, s.W, s.H, dqkvw, s.H, st));
      if (dqkvw == t->dWtmp) {
        TRY(copy_rows(t, y.gq, t->dWtmp, (size_t)s.nq * s.H, st));
        TRY(copy_rows(t, y.gk, t->dWtmp + (size_t)s.nq * s.H, kvw * s.H, st));
        TRY(copy_rows(t, y.gv, t->dWtmp + ((size_t)s.nq + kvw) * s.H, kvw * s.H, st));
      }
    }
    TRY(norm_bwd(t, s, y.h_in, y.ln1, t->dHn, dh, dh, want_grads ? y.gln1 : nullptr, R, st));
    t->launches += 5;
  }
  return 0;
}

int setup_stack(CsmTrain* t, TStack& s, const CsmLlamaShape& sh, int max_rows) {
  s.H = sh.hidden; s.I = sh.inter; s.L = sh.layers; s.heads = sh.heads; s.kv = sh.kv_heads;
  s.hd = s.H / s.heads; s.nq = s.heads * s.hd; s.W = (s.heads + 2 * s.kv) * s.hd;
  s.eps = sh.eps; s.scale = 1.0f / sqrtf((float)s.hd); s.n_pos = sh.n_pos; s.max_rows = max_rows;
  if (s.hd != 64 && s.hd != 128) return tfail(t, CSM_EUNSUPPORTED, "head_dim %d (64 and 128 are built)", s.hd);
  if (s.H % 64 || s.I % 64 || s.W % 64 || s.H > 2048) return tfail(t, CSM_EUNSUPPORTED, "hidden / intermediate sizes must be multiples of 64, hidden <= 2048");
  const size_t half = (size_t)sh.n_pos * (s.hd / 2);
  TRY(talloc(t, &s.cos_t, half));
  TRY(talloc(t, &s.sin_t, half));
  TCK(cudaMemcpy(s.cos_t, sh.rope_cos, half * 2, cudaMemcpyHostToDevice));
  TCK(cudaMemcpy(s.sin_t, sh.rope_sin, half * 2, cudaMemcpyHostToDevice));
  const size_t R = (size_t)max_rows;
  s.layers.resize(s.L);
  for (TLayer& y : s.layers) {
    memset(&y, 0, sizeof y);
    TRY(talloc(t, &y.Wqkv, (size_t)s.W * s.H));
    TRY(talloc(t, &y.WqkvT, (size_t)s.W * s.H));
    TRY(talloc(t, &y.WoT, (size_t)s.nq * s.H));
    TRY(talloc(t, &y.Wgu, (size_t)2 * s.I * s.H));
    TRY(talloc(t, &y.WguT, (size_t)2 * s.I * s.H));
    TRY(talloc(t, &y.WdownT, (size_t)s.I * s.H));
    TRY(talloc(t, &y.h_in, R * s.H));
    TRY(talloc(t, &y.hn1, R * s.H));
    TRY(talloc(t, &y.qkv, R * s.W));
    TRY(talloc(t, &y.attn, R * s.nq));
    TRY(talloc(t, &y.h_mid, R * s.H));
    TRY(talloc(t, &y.hn2, R * s.H));
    TRY(talloc(t, &y.gu, R * 2 * s.I));
    TRY(talloc(t, &y.act, R * s.I));
    TRY(talloc(t, &y.lse, R * s.heads));
  }
  TRY(talloc(t, &s.h_out, R * s.H));
  TRY(talloc(t, &s.hf, R * s.H));
  return 0;
}

// the caller's parameters of this step -> fused and transposed copies
int bind_stack(CsmTrain* t, TStack& s, const void* const* w, const void* const* g, const void* norm, const void* gnorm,
               cudaStream_t st) {
  s.norm = (const bf16*)norm;
  s.gnorm = (bf16*)gnorm;
  const size_t kvw = (size_t)s.kv * s.hd;
  for (int l = 0; l < s.L; ++l) {
    TLayer& y = s.layers[l];
    const void* const* p = w + (size_t)l * CSM_W_PER_LAYER;
    y.q = (const bf16*)p[CSM_W_Q]; y.k = (const bf16*)p[CSM_W_K]; y.v = (const bf16*)p[CSM_W_V]; y.o = (const bf16*)p[CSM_W_O];
    y.gate = (const bf16*)p[CSM_W_GATE]; y.up = (const bf16*)p[CSM_W_UP]; y.down = (const bf16*)p[CSM_W_DOWN];
    y.ln1 = (const bf16*)p[CSM_W_LN1]; y.ln2 = (const bf16*)p[CSM_W_LN2];
    if (g) {
      const void* const* q = g + (size_t)l * CSM_W_PER_LAYER;
      y.gq = (bf16*)q[CSM_W_Q]; y.gk = (bf16*)q[CSM_W_K]; y.gv = (bf16*)q[CSM_W_V]; y.go = (bf16*)q[CSM_W_O];
      y.ggate = (bf16*)q[CSM_W_GATE]; y.gup = (bf16*)q[CSM_W_UP]; y.gdown = (bf16*)q[CSM_W_DOWN];
      y.gln1 = (bf16*)q[CSM_W_LN1]; y.gln2 = (bf16*)q[CSM_W_LN2];
    }
    TRY(copy_rows(t, y.Wqkv, y.q, (size_t)s.nq * s.H, st));
    TRY(copy_rows(t, y.Wqkv + (size_t)s.nq * s.H, y.k, kvw * s.H, st));
    TRY(copy_rows(t, y.Wqkv + ((size_t)s.nq + kvw) * s.H, y.v, kvw * s.H, st));
    TRY(copy_rows(t, y.Wgu, y.gate, (size_t)s.I * s.H, st));
    TRY(copy_rows(t, y.Wgu + (size_t)s.I * s.H, y.up, (size_t)s.I * s.H, st));
    if (g && !t->mn_operands) {
      TRY(transpose(t, y.Wqkv, s.W, s.H, s.H, y.WqkvT, s.W, st));      // [W, H] -> [H, W]
      TRY(transpose(t, y.o, s.H, s.nq, s.nq, y.WoT, s.H, st));         // [H, nq] -> [nq, H]
      TRY(transpose(t, y.Wgu, 2 * s.I, s.H, s.H, y.WguT, 2 * s.I, st)); // [2I, H] -> [H, 2I]
      TRY(transpose(t, y.down, s.H, s.I, s.I, y.WdownT, s.H, st));     // [H, I] -> [I, H]
    }
  }
  return 0;
}

}  // namespace

// =================================================================== exported C ABI
extern "C" {

int csm_train_create(const CsmShapes* sh, int max_tokens, int max_frames, CsmTrain** out) {
  if (!out) return CSM_EINVAL;
  *out = nullptr;
  CsmTrain* t = new CsmTrain();
  *out = t;
  if (!sh || max_tokens < 2 || max_frames < 1) return tfail(t, CSM_EINVAL, "bad arguments");
  if (sh->n_codebooks != 32) return tfail(t, CSM_EINVAL, "audio_num_codebooks must be 32");
  int dev = 0;
  TCK(cudaGetDevice(&dev));
  TCK(cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, dev));
  t->V = sh->audio_vocab; t->Vp = rup(sh->audio_vocab, 64); t->text_vocab = sh->text_vocab;
  t->max_tokens = max_tokens; t->max_frames = max_frames;
  t->mn_operands = getenv("CSM_TRAIN_TRANSPOSE") == nullptr;
  t->flash_tc = getenv("CSM_FLASH_MMA") == nullptr;
  t->flash_tc_bwd = getenv("CSM_FLASH_MMA") == nullptr && getenv("CSM_FLASH_BWD_MMA") == nullptr;
  if (sh->backbone.n_pos < 1 || sh->decoder.n_pos < 33) return tfail(t, CSM_EINVAL, "rope tables: the decoder needs 33 positions");
  const int Rd = max_frames * 33;
  TRY(setup_stack(t, t->bb, sh->backbone, max_tokens));
  TRY(setup_stack(t, t->dec, sh->decoder, Rd));
  const TStack &b = t->bb, &d = t->dec;
  const size_t Rm = (size_t)(max_tokens > Rd ? max_tokens : Rd), Rp = (size_t)rup((int)Rm, 8);
  auto mx = [](size_t a, size_t c) { return a > c ? a : c; };
  const size_t wide = mx(mx((size_t)2 * b.I, (size_t)2 * d.I), mx((size_t)t->Vp, mx((size_t)b.W, (size_t)d.W)));
  const size_t Hm = mx(b.H, d.H), Im = mx(b.I, d.I), Wm = mx(b.W, d.W), nqm = mx(b.nq, d.nq), hm = mx(b.heads, d.heads);
  TRY(talloc(t, &t->tA, wide * Rp));
  TRY(talloc(t, &t->tB, mx(wide, Hm) * Rp));
  TRY(talloc(t, &t->dAct, Rm * Im));
  TRY(talloc(t, &t->dGU, Rm * 2 * Im));
  TRY(talloc(t, &t->dHn, Rm * Hm));
  TRY(talloc(t, &t->dAttn, Rm * nqm));
  TRY(talloc(t, &t->dQKV, Rm * Wm));
  TRY(talloc(t, &t->dWtmp, mx(mx(2 * Im, Wm), (size_t)t->Vp) * Hm + (size_t)d.H * t->Vp));
  TRY(talloc(t, &t->dq_acc, Rm * nqm));
  TRY(talloc(t, &t->delta, Rm * hm));
  TRY(talloc(t, &t->dw_acc, (size_t)2048));
  TRY(talloc(t, &t->dh_bb, (size_t)max_tokens * b.H));
  TRY(talloc(t, &t->dh_dec, (size_t)Rd * d.H));
  TRY(talloc(t, &t->Wc0p, (size_t)t->Vp * b.H));
  TRY(talloc(t, &t->Wc0pT, (size_t)t->Vp * b.H));
  TRY(talloc(t, &t->logits0, (size_t)max_tokens * t->Vp));
  TRY(talloc(t, &t->AHt, (size_t)31 * t->Vp * d.H));
  TRY(talloc(t, &t->AHp, (size_t)31 * t->Vp * d.H));
  TRY(talloc(t, &t->logits_d, (size_t)31 * max_frames * t->Vp));
  TRY(talloc(t, &t->WprojT, (size_t)b.H * d.H));
  TRY(talloc(t, &t->dec_in, (size_t)Rd * b.H));
  TRY(talloc(t, &t->d_dec_in, (size_t)Rd * b.H));
  TRY(talloc(t, &t->dhdf, (size_t)Rd * d.H));
  TRY(talloc(t, &t->valid, (size_t)max_tokens));
  TRY(talloc(t, &t->fflag, (size_t)max_tokens));
  TRY(talloc(t, &t->frames, (size_t)max_tokens));
  TRY(talloc(t, &t->counts, (size_t)4));
  TRY(talloc(t, &t->lab0, (size_t)max_tokens));
  TRY(talloc(t, &t->labd, (size_t)31 * max_frames));
  TRY(talloc(t, &t->row_loss, mx((size_t)max_tokens, (size_t)31 * max_frames)));
  TRY(talloc(t, &t->losses, (size_t)4));
  TRY(talloc(t, &t->audio_acc, (size_t)t->V * 32 * b.H));
  TRY(talloc(t, &t->text_acc, (size_t)t->text_vocab * b.H));
  return 0;
}

int csm_train_destroy(CsmTrain* t) {
  if (!t) return 0;
  for (void* p : t->allocs) cudaFree(p);
  delete t;
  return 0;
}

const char* csm_train_last_error(const CsmTrain* t) { return t ? t->err.c_str() : "null training context"; }

int csm_train_launches(const CsmTrain* t) { return t ? t->launches : 0; }

/* Tests: host copy of a named intermediate of the last step ("bb.0.qkv", "dec.1.attn", "d.bb.x0", ...).  host == NULL:
 * only the size is returned. */
int csm_train_debug(CsmTrain* t, const char* name, void* host, long long cap, long long* bytes) {
  if (!t || !name || !bytes) return CSM_EINVAL;
  auto it = t->dbg.find(name);
  if (it == t->dbg.end()) return tfail(t, CSM_EINVAL, "no intermediate named %s", name);
  *bytes = (long long)it->second.second;
  if (host) {
    if (cap < *bytes) return tfail(t, CSM_EINVAL, "buffer of %lld bytes for %lld", cap, *bytes);
    TCK(cudaDeviceSynchronize());
    TCK(cudaMemcpy(host, it->second.first, (size_t)*bytes, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int csm_train_step_begin(CsmTrain* t, const CsmWeights* w, const CsmWeights* g, const int64_t* ids, const int32_t* mask,
                         const int64_t* labels, int B, int S, int split_layer, int* n_frames_host, void* last_h_out,
                         void* c0_logits_out, void* stream) {
  if (!t) return CSM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  t->pend.active = false;
  if (!w || !ids || !labels) return tfail(t, CSM_EINVAL, "null weights / ids / labels");
  if (split_layer < 0 || split_layer > t->bb.L) return tfail(t, CSM_EINVAL, "split_layer %d outside [0, %d]", split_layer, t->bb.L);
  const int R = B * S;
  if (B < 1 || S < 2 || R > t->max_tokens) return tfail(t, CSM_ECAPACITY, "B*S = %d exceeds max_tokens %d (or S < 2)", R, t->max_tokens);
  if (S > t->bb.n_pos) return tfail(t, CSM_ECAPACITY, "sequence length %d exceeds the %d rope positions", S, t->bb.n_pos);
  TStack &b = t->bb, &d = t->dec;
  const int V = t->V, Vp = t->Vp;
  const bool want = g != nullptr;
  const long long* ids_ll = reinterpret_cast<const long long*>(ids);
  const long long* lab_ll = reinterpret_cast<const long long*>(labels);
  t->dbg.clear();

  // ---- this step's parameters
  TRY(bind_stack(t, b, w->backbone_layers, want ? g->backbone_layers : nullptr, w->backbone_norm, want ? g->backbone_norm : nullptr, st));
  TRY(bind_stack(t, d, w->decoder_layers, want ? g->decoder_layers : nullptr, w->decoder_norm, want ? g->decoder_norm : nullptr, st));
  const bf16* text_emb = (const bf16*)w->text_embeddings;
  const bf16* audio_emb = (const bf16*)w->audio_embeddings;
  const bf16* proj = (const bf16*)w->projection;           // [Hd, Hb]
  const bf16* c0 = (const bf16*)w->codebook0_head;         // [V, Hb]
  const bf16* ah = (const bf16*)w->audio_head;             // [31, Hd, V]
  // codebook0_head padded to Vp rows (zero rows: their logits are 0 and are never read); audio_head[c] as [Vp, Hd]
  // (the W operand of the logits GEMM) and as [Hd, Vp] (the W operand of its input gradient)
  TCK(cudaMemsetAsync(t->Wc0p, 0, (size_t)Vp * b.H * 2, st));
  TRY(copy_rows(t, t->Wc0p, c0, (size_t)V * b.H, st));
  TCK(cudaMemsetAsync(t->AHt, 0, (size_t)31 * Vp * d.H * 2, st));
  for (int c = 0; c < 31; ++c) TRY(transpose(t, ah + (size_t)c * d.H * V, d.H, V, V, t->AHt + (size_t)c * Vp * d.H, d.H, st));
  if (want && !t->mn_operands) {
    TRY(transpose(t, t->Wc0p, Vp, b.H, b.H, t->Wc0pT, Vp, st));
    TRY(transpose(t, proj, d.H, b.H, b.H, t->WprojT, d.H, st));
    TCK(cudaMemsetAsync(t->AHp, 0, (size_t)31 * Vp * d.H * 2, st));
    TCK(cudaMemcpy2DAsync(t->AHp, (size_t)Vp * 2, ah, (size_t)V * 2, (size_t)V * 2, (size_t)31 * d.H, cudaMemcpyDeviceToDevice, st));
  }

  // ---- backbone forward (modeling_csm.py:319-361)
  TCK(cudaMemsetAsync(t->counts, 0, 4 * sizeof(int), st));
  const unsigned char* valid = nullptr;
  if (mask) {
    TCK(csm_frame_valid_launch(mask, R, t->valid, nullptr, st));
    valid = t->valid;
  }
  TCK(csm_embed_sum_launch(ids_ll, mask, 1, audio_emb, text_emb, V, b.H, b.layers[0].h_in, R, st));
  t->launches += 2;
  TRY(stack_forward(t, b, S, B, valid, st));
  TRY(gemm(t, b.hf, b.H, R, b.H, t->Wc0p, b.H, Vp, t->logits0, Vp, EPI_STORE, st));
  // the last position's hidden state and codebook-0 logits (modeling_csm.py:363-365), before the logits are overwritten
  if (last_h_out)
    TCK(cudaMemcpy2DAsync(last_h_out, (size_t)b.H * 2, b.hf + (size_t)(S - 1) * b.H, (size_t)S * b.H * 2, (size_t)b.H * 2, B,
                          cudaMemcpyDeviceToDevice, st));
  if (c0_logits_out)
    TCK(cudaMemcpy2DAsync(c0_logits_out, (size_t)V * 2, t->logits0 + (size_t)(S - 1) * Vp, (size_t)S * Vp * 2, (size_t)V * 2, B,
                          cudaMemcpyDeviceToDevice, st));
  // codebook-0 loss with the causal shift on float32 logits (:376-389); the logits become their own gradient
  shift_labels_kernel<<<(R + 255) / 256, 256, 0, st>>>(lab_ll, B, S, t->lab0, t->counts + 1);
  TCK(cudaGetLastError());
  ce_rows_kernel<false><<<(int)(((long long)R * 32 + 255) / 256), 256, 0, st>>>(t->logits0, Vp, V, R, t->lab0, t->counts + 1, t->row_loss);
  TCK(cudaGetLastError());
  mean_loss_kernel<<<1, 1024, 0, st>>>(t->row_loss, R, t->counts + 1, 0, t->losses + 1);
  TCK(cudaGetLastError());
  // ---- frames with all 32 audio labels (:392-399); their number is needed on the host (the reference's nonzero syncs too)
  frame_flag_kernel<<<(R + 255) / 256, 256, 0, st>>>(lab_ll, R, t->fflag);
  TCK(cudaGetLastError());
  frame_list_kernel<<<1, 1024, 0, st>>>(t->fflag, R, t->frames, t->counts);
  TCK(cudaGetLastError());
  t->launches += 5;
  int F = 0;
  TCK(cudaMemcpyAsync(&F, t->counts, sizeof(int), cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  if (F > t->max_frames) return tfail(t, CSM_ECAPACITY, "%d frames carry decoder labels, max_frames is %d", F, t->max_frames);
  if (n_frames_host) *n_frames_host = F;
  const int Rd = F * 33;
  TCK(cudaMemsetAsync(t->losses + 2, 0, sizeof(float), st));
  if (F > 0) {
    // ---- decoder forward on the selected frames (:405-457)
    decoder_gather_kernel<<<Rd, 256, 0, st>>>(b.hf, audio_emb, ids_ll, t->frames, S, V, b.H, t->dec_in);
    TCK(cudaGetLastError());
    TRY(gemm(t, t->dec_in, b.H, Rd, b.H, proj, b.H, d.H, d.layers[0].h_in, d.H, EPI_STORE, st));
    TRY(stack_forward(t, d, 33, F, nullptr, st));
    for (int c = 0; c < 31; ++c)   // einsum fcd,cdv->fcv: position c + 1 against audio_head[c]
      TRY(gemm(t, d.hf + (size_t)(c + 1) * d.H, (long long)33 * d.H, F, d.H, t->AHt + (size_t)c * Vp * d.H, d.H, Vp,
               t->logits_d + (size_t)c * F * Vp, Vp, EPI_STORE, st));
    decoder_labels_kernel<<<(F * 31 + 255) / 256, 256, 0, st>>>(lab_ll, t->frames, F, t->labd, t->counts + 2);
    TCK(cudaGetLastError());
    ce_rows_kernel<true><<<(int)(((long long)F * 31 * 32 + 255) / 256), 256, 0, st>>>(t->logits_d, Vp, V, F * 31, t->labd,
                                                                                      t->counts + 2, t->row_loss);
    TCK(cudaGetLastError());
    mean_loss_kernel<<<1, 1024, 0, st>>>(t->row_loss, F * 31, t->counts + 2, 1, t->losses + 2);
    TCK(cudaGetLastError());
    t->launches += 4;
  }
  t->dbg["bb.hf"] = {b.hf, (size_t)R * b.H * 2};
  t->dbg["bb.h_out"] = {b.h_out, (size_t)R * b.H * 2};
  t->dbg["dec.hf"] = {d.hf, (size_t)Rd * d.H * 2};
  t->dbg["dec_in"] = {t->dec_in, (size_t)Rd * b.H * 2};
  for (int k = 0; k < 2; ++k) {
    TStack& s = k ? d : b;
    const size_t rows = k ? Rd : R;
    for (int l = 0; l < s.L; ++l) {
      const std::string p = std::string(k ? "dec." : "bb.") + std::to_string(l) + ".";
      const TLayer& y = s.layers[l];
      t->dbg[p + "h_in"] = {y.h_in, rows * s.H * 2};
      t->dbg[p + "hn1"] = {y.hn1, rows * s.H * 2};
      t->dbg[p + "qkv"] = {y.qkv, rows * s.W * 2};
      t->dbg[p + "attn"] = {y.attn, rows * s.nq * 2};
      t->dbg[p + "h_mid"] = {y.h_mid, rows * s.H * 2};
      t->dbg[p + "hn2"] = {y.hn2, rows * s.H * 2};
      t->dbg[p + "act"] = {y.act, rows * s.I * 2};
    }
  }

  if (want) {
    // ================= backward =================
    TCK(cudaMemsetAsync(t->audio_acc, 0, (size_t)V * 32 * b.H * 4, st));
    TCK(cudaMemsetAsync(t->text_acc, 0, (size_t)t->text_vocab * b.H * 4, st));
    // codebook-0 head: d hf = dlogits0 Wc0, d Wc0 = dlogits0^T hf
    TRY(dgrad(t, t->logits0, Vp, R, Vp, t->Wc0p, b.H, b.H, t->Wc0pT, t->dh_bb, b.H, st));
    TRY(wgrad(t, t->logits0, Vp, b.hf, b.H, R, Vp, b.H, t->dWtmp, b.H, st));
    TRY(copy_rows(t, (bf16*)g->codebook0_head, t->dWtmp, (size_t)V * b.H, st));
    if (F > 0) {
      // audio heads: d hdf[:, c+1] = dlogits_c audio_head[c]^T, d audio_head[c] = hdf[:, c+1]^T dlogits_c
      TCK(cudaMemsetAsync(t->dhdf, 0, (size_t)Rd * d.H * 2, st));
      bf16* dAH = t->dWtmp + (size_t)Vp * b.H;    // [Hd, Vp] scratch behind the codebook-0 gradient
      for (int c = 0; c < 31; ++c) {
        const bf16* dl = t->logits_d + (size_t)c * F * Vp;
        TRY(dgrad(t, dl, Vp, F, Vp, t->AHt + (size_t)c * Vp * d.H, d.H, d.H, t->AHp + (size_t)c * d.H * Vp,
                  t->dhdf + (size_t)(c + 1) * d.H, 33 * d.H, st));
        TRY(wgrad(t, d.hf + (size_t)(c + 1) * d.H, (long long)33 * d.H, dl, Vp, F, d.H, Vp, dAH, Vp, st));
        TCK(cudaMemcpy2DAsync((bf16*)g->audio_head + (size_t)c * d.H * V, (size_t)V * 2, dAH, (size_t)Vp * 2, (size_t)V * 2,
                              (size_t)d.H, cudaMemcpyDeviceToDevice, st));
      }
      TRY(norm_bwd(t, d, d.h_out, d.norm, t->dhdf, nullptr, t->dh_dec, d.gnorm, Rd, st));
      TRY(stack_backward(t, d, "dec", 33, F, nullptr, t->dh_dec, true, st));
      TRY(note(t, "d.dec_x0", t->dh_dec, (size_t)Rd * d.H * 2, st));
      // projection: d dec_in = d x0 Wproj, d Wproj = d x0^T dec_in
      TRY(dgrad(t, t->dh_dec, d.H, Rd, d.H, proj, b.H, b.H, t->WprojT, t->d_dec_in, b.H, st));
      TRY(wgrad(t, t->dh_dec, d.H, t->dec_in, b.H, Rd, d.H, b.H, (bf16*)g->projection, b.H, st));
      decoder_scatter_kernel<<<Rd, 256, 0, st>>>(t->d_dec_in, ids_ll, t->frames, S, V, b.H, t->dh_bb, t->audio_acc);
      TCK(cudaGetLastError());
      t->launches += 1;
    } else {
      TCK(cudaMemsetAsync((void*)g->audio_head, 0, (size_t)31 * d.H * V * 2, st));   // no frame carries decoder labels
      TCK(cudaMemsetAsync((void*)g->projection, 0, (size_t)d.H * b.H * 2, st));
      TCK(cudaMemsetAsync((void*)g->decoder_norm, 0, (size_t)d.H * 2, st));
      for (int l = 0; l < d.L; ++l) {
        const TLayer& y = d.layers[l];
        const size_t kvw = (size_t)d.kv * d.hd;
        TCK(cudaMemsetAsync(y.gq, 0, (size_t)d.nq * d.H * 2, st));
        TCK(cudaMemsetAsync(y.gk, 0, kvw * d.H * 2, st));
        TCK(cudaMemsetAsync(y.gv, 0, kvw * d.H * 2, st));
        TCK(cudaMemsetAsync(y.go, 0, (size_t)d.nq * d.H * 2, st));
        TCK(cudaMemsetAsync(y.ggate, 0, (size_t)d.I * d.H * 2, st));
        TCK(cudaMemsetAsync(y.gup, 0, (size_t)d.I * d.H * 2, st));
        TCK(cudaMemsetAsync(y.gdown, 0, (size_t)d.I * d.H * 2, st));
        TCK(cudaMemsetAsync(y.gln1, 0, (size_t)d.H * 2, st));
        TCK(cudaMemsetAsync(y.gln2, 0, (size_t)d.H * 2, st));
      }
    }
    TRY(note(t, "d.bb.hf", t->dh_bb, (size_t)R * b.H * 2, st));
    // backbone: final norm, then the layers from the last one down to split_layer (the rest in csm_train_step_end)
    TRY(norm_bwd(t, b, b.h_out, b.norm, t->dh_bb, nullptr, t->dh_bb, b.gnorm, R, st));
    if (split_layer < b.L) TRY(stack_backward(t, b, "bb", S, B, valid, t->dh_bb, true, st, b.L - 1, split_layer));
    t->pend.g_audio = (bf16*)g->audio_embeddings;
    t->pend.g_text = (bf16*)g->text_embeddings;
  }
  t->pend.active = true; t->pend.want = want; t->pend.B = B; t->pend.S = S; t->pend.split = split_layer;
  t->pend.ids = ids_ll; t->pend.mask = mask; t->pend.valid = valid;
  return 0;
}

/* Second half of a step started with csm_train_step_begin: the backbone layers below split_layer, the embedding-table
 * gradients, and the losses (synchronises the stream). */
int csm_train_step_end(CsmTrain* t, float* losses_host, void* stream) {
  if (!t || !losses_host) return CSM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!t->pend.active) return tfail(t, CSM_EINVAL, "csm_train_step_end without csm_train_step_begin");
  t->pend.active = false;
  TStack& b = t->bb;
  const int B = t->pend.B, S = t->pend.S, R = B * S, V = t->V;
  if (t->pend.want) {
    if (t->pend.split > 0) TRY(stack_backward(t, b, "bb", S, B, t->pend.valid, t->dh_bb, true, st, t->pend.split - 1, 0));
    TRY(note(t, "d.bb.x0", t->dh_bb, (size_t)R * b.H * 2, st));
    embed_bwd_kernel<<<R, 256, 0, st>>>(t->dh_bb, t->pend.ids, t->pend.mask, V, b.H, t->audio_acc, t->text_acc);
    TCK(cudaGetLastError());
    f32_to_bf16_kernel<<<nblocks((long long)V * 32 * b.H / 2), 256, 0, st>>>(t->audio_acc, (long long)V * 32 * b.H, t->pend.g_audio);
    TCK(cudaGetLastError());
    f32_to_bf16_kernel<<<nblocks((long long)t->text_vocab * b.H / 2), 256, 0, st>>>(t->text_acc, (long long)t->text_vocab * b.H,
                                                                                   t->pend.g_text);
    TCK(cudaGetLastError());
    t->launches += 3;
  }
  float l3[4];
  TCK(cudaMemcpyAsync(l3, t->losses, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  TCK(cudaStreamSynchronize(st));
  losses_host[1] = l3[1];
  losses_host[2] = l3[2];
  losses_host[0] = l3[1] + l3[2];   // modeling_csm.py:471: loss = backbone_loss + decoder_loss
  return 0;
}

int csm_train_step(CsmTrain* t, const CsmWeights* w, const CsmWeights* g, const int64_t* ids, const int32_t* mask,
                   const int64_t* labels, int B, int S, float* losses_host, int* n_frames_host, void* last_h_out,
                   void* c0_logits_out, void* stream) {
  if (!losses_host) return t ? tfail(t, CSM_EINVAL, "null losses") : CSM_EINVAL;
  int r = csm_train_step_begin(t, w, g, ids, mask, labels, B, S, 0, n_frames_host, last_h_out, c0_logits_out, stream);
  if (r) return r;
  return csm_train_step_end(t, losses_host, stream);
}

}  // extern "C"
