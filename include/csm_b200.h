/*
 * csm_b200.h -- C ABI of libcsm_b200.so, the B200 (sm_100a) engine behind the reference's
 * CSMModel Python API.
 *
 * The reference (thomasgauthier/csm-hf) has no FFI: its boundary is the Python class
 * CSMModel in modeling_csm.py.  Each entry point below replaces one arrow of that
 * class's call graph; the Python mirror in
 * csm_hf_b200/modeling.py binds them with ctypes (INTEGRATION.md shows the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types; no exceptions cross the ABI;
 *   - return 0 on success, a negative CSM_E* code on failure; csm_last_error() gives text;
 *   - unless the name ends in _host, every data pointer is a DEVICE pointer and all
 *     work is enqueued on the caller's `stream` (a cudaStream_t passed as void*);
 *     nothing synchronises the device except csm_generate_host and csm_frames_done;
 *   - weights are bf16 in the reference state_dict layouts ([out,in] row-major for every
 *     nn.Linear, audio_head [31, in, out]); the engine packs private copies at create
 *     time, the caller may free its tensors afterwards;
 *   - one CsmCtx per device, not thread-safe (the reference is single-threaded Python).
 */
#ifndef CSM_B200_H
#define CSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CsmCtx CsmCtx;

enum {
  CSM_OK = 0,
  CSM_EINVAL = -1,      /* bad argument / unsupported shape (maps to ValueError)   */
  CSM_ECUDA = -2,       /* CUDA runtime error                                       */
  CSM_ECAPACITY = -3,   /* batch > max_batch or context > max_ctx                   */
  CSM_EUNSUPPORTED = -4 /* feature of the reference not on the accelerated path     */
};

/* One Llama stack: LlamaConfig fields read by the path (modeling_csm.py:68-109). */
typedef struct {
  int32_t hidden;       /* hidden_size                */
  int32_t inter;        /* intermediate_size          */
  int32_t layers;       /* num_hidden_layers          */
  int32_t heads;        /* num_attention_heads        */
  int32_t kv_heads;     /* num_key_value_heads        */
  float eps;            /* rms_norm_eps               */
  /* HOST pointers, bf16 bits, [n_pos][head_dim/2]: cos/sin of inv_freq*pos rounded to bf16
   * exactly as LlamaRotaryEmbedding.forward does (hf modeling_llama.py:122-135). */
  const uint16_t* rope_cos;
  const uint16_t* rope_sin;
  int32_t n_pos;
} CsmLlamaShape;

typedef struct {
  int32_t text_vocab;   /* CSMConfig.text_vocab_size      (modeling_csm.py:64) */
  int32_t audio_vocab;  /* CSMConfig.audio_vocab_size     (:65)                */
  int32_t n_codebooks;  /* CSMConfig.audio_num_codebooks  (:66), must be 32    */
  CsmLlamaShape backbone;
  CsmLlamaShape decoder;
} CsmShapes;

/* Per-layer tensors in this order (state_dict key suffixes). */
enum { CSM_W_Q = 0, CSM_W_K, CSM_W_V, CSM_W_O, CSM_W_GATE, CSM_W_UP, CSM_W_DOWN, CSM_W_LN1, CSM_W_LN2, CSM_W_PER_LAYER };

typedef struct {
  const void* text_embeddings;   /* [text_vocab, Hb]                (modeling_csm.py:222) */
  const void* audio_embeddings;  /* [audio_vocab*32, Hb]            (:223-225)            */
  const void* projection;        /* [Hd, Hb]                        (:228)                */
  const void* codebook0_head;    /* [audio_vocab, Hb]               (:231-233)            */
  const void* audio_head;        /* [31, Hd, audio_vocab]           (:236-240)            */
  const void* backbone_norm;     /* [Hb]                                                  */
  const void* decoder_norm;      /* [Hd]                                                  */
  const void* const* backbone_layers; /* HOST array [layers*CSM_W_PER_LAYER] of device ptrs */
  const void* const* decoder_layers;  /* HOST array [layers*CSM_W_PER_LAYER] of device ptrs */
} CsmWeights;

/* CSMModel.__init__ + from_pretrained (modeling_csm.py:214-245): pack weights, size the
 * backbone KV cache for max_batch x max_ctx positions. */
int csm_create(const CsmShapes* shapes, const CsmWeights* weights, int max_batch, int max_ctx,
               void* stream, CsmCtx** out);
int csm_destroy(CsmCtx* ctx);

/* CSMModel.reset_caches (modeling_csm.py:288-290): forget the cached context. */
int csm_reset(CsmCtx* ctx);
/* Number of backbone positions currently cached (== DynamicCache.get_seq_length()). */
int csm_cache_len(const CsmCtx* ctx);

/* _embed_tokens + mask multiply + sum over the 33 slots (modeling_csm.py:261-282,327-334).
 * ids int64 [B,S,33]; mask int32 [B,S,33] or NULL (= all ones); out bf16 [B,S,Hb]. */
int csm_embed_sum(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int S,
                  void* out_bf16, void* stream);

/* CSMModel.generate_frame (modeling_csm.py:484-589) at temperature 0 (reference:
 * topk=1, ties -> lowest index).  S == cached-context continuation: S > 1 is the
 * prefill call (causal), S == 1 a decode step attending to every cached position.
 *   ids int64 [B,S,33], mask int32 [B,S,33].  mask NULL means "audio columns set, text column clear" (the row
 *       generate() builds for a decode step, modeling_csm.py:684-687) -- NOT the reference's attention_mask=None,
 *       which sums all 33 slots (:328-332); the Python mirror passes an explicit all-ones mask for that case.
 *       A prefill into an empty cache honours padding: frames whose 33 mask entries are all zero are hidden as
 *       keys (modeling_csm.py:337-342); decode steps attend to every cached position, as the reference does.
 *   force_tokens int64 [B,32] or NULL: teacher forcing -- the decoder is fed these
 *       tokens instead of its own argmax (samples still report the argmax)
 *   samples int64 [B,32]; last_h bf16 [B,Hb]; c0_logits bf16 [B,V];
 *   cb_logits bf16 [B,31,V] (logits given to the sampler for codebooks 1..31); any
 *   output may be NULL. */
int csm_generate_frame(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int S,
                       const int64_t* force_tokens, int64_t* samples, void* last_h,
                       void* c0_logits, void* cb_logits, void* stream);

/* CSMModel.generate (modeling_csm.py:591-702), use_cache=True, temperature 0.
 * Resets the cache, prefills T context frames, emits up to max_new_frames frames into
 * frames int64 [B,max_new_frames,32].  Asynchronous: the number of frames kept
 * (stop_on_all_zeros, :662-663) is read back with csm_frames_done. */
int csm_generate(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int T,
                 int max_new_frames, int stop_on_all_zeros, int64_t* frames, void* stream);
/* Continues the generate() loop (modeling_csm.py:644-690) of the last csm_generate / csm_generate_more call for up to
 * n_more further frames into frames int64 [B,n_more,32] -- the same decode steps, issued in chunks, so that a caller
 * that shards a batch over several contexts can exchange the stop flag (:662-663) between chunks. */
int csm_generate_more(CsmCtx* ctx, int B, int n_more, int stop_on_all_zeros, int64_t* frames, void* stream);
/* Synchronises `stream` and returns the frame count of the last csm_generate / csm_generate_more (<0: error). */
int csm_frames_done(CsmCtx* ctx, void* stream);

/* Same as csm_generate with HOST buffers: copies ids/mask in, frames out, synchronises,
 * stores the frame count in *n_out.  This is the end-to-end call bench.py times. */
int csm_generate_host(CsmCtx* ctx, const int64_t* ids_host, const int32_t* mask_host, int B, int T,
                      int max_new_frames, int stop_on_all_zeros, int64_t* frames_host, int* n_out,
                      void* stream);

/* Introspection used by tests and bench.py. */
enum { CSM_INFO_SMS = 0, CSM_INFO_GRID, CSM_INFO_PHASES_PER_FRAME, CSM_INFO_SMEM_BYTES,
       CSM_INFO_LAUNCHES, CSM_INFO_STEPPED };
int64_t csm_info(const CsmCtx* ctx, int what);
/* stepped != 0: launch one kernel per phase instead of one persistent launch per frame
 * (debug/bisect aid; results are identical). */
int csm_set_stepped(CsmCtx* ctx, int stepped);
/* Device time of the decode-frame launches of the last csm_generate, in ms, measured
 * with CUDA events on `stream` (synchronises); *n = launches covered. */
int csm_last_decode_ms(CsmCtx* ctx, float* ms, int* n);

const char* csm_last_error(const CsmCtx* ctx);

/* Sampling mode of the following csm_generate_frame / csm_generate calls (sample_topk, modeling_csm.py:179-189).
 * topk <= 1 or temperature == 0: greedy, lowest index on ties (the default).  Otherwise: keep the logits >= the
 * k-th largest of logits/temperature, softmax, draw -- by Gumbel-max with counter-based noise keyed by
 * (seed, frames generated since this call, codebook, seq_base + sequence index, vocabulary index); the reference
 * draws from torch's global generator, so parity is distributional.  seq_base: global index of this context's
 * sequence 0 when a batch is sharded over several contexts. */
int csm_set_sampling(CsmCtx* ctx, int topk, float temperature, uint64_t seed, int seq_base);

/* The same sampler on stand-alone rows (no context): logits [rows][V] bf16 on the device -> out[rows] int64;
 * row r uses the noise key (seed, 0, 0, r).  Used by the distribution tests. */
int csm_sample_topk(const void* logits, int rows, int V, int topk, float temperature, uint64_t seed, int64_t* out,
                    void* stream);

/* One nn.Linear (no bias) on the tcgen05 tensor cores, the kernel the context prefill runs its projections on
 * (hf modeling_llama.py:183,262-264,288): C[R,N] = x[R,K] * W[N,K]^T, bf16 in, fp32 accumulate, bf16 out, with the
 * element-wise tail that follows the projection in the reference fused in:
 *   tail 0: none; tail 1: C += y (residual add, C holds the residual on entry); tail 2: SwiGLU -- W rows interleaved
 *   (gate_j, up_j), C [R, N/2] = bf16(silu(gate)) * up.
 * x row pitch ldx, C row pitch ldc (elements); K % 64 == 0, N % 64 == 0; all device pointers, 16-byte aligned. */
int csm_linear(const void* x, int ldx, const void* W, int R, int N, int K, int tail, void* C, int ldc, void* stream);

/* ---- debug / test hooks: not part of the drop-in surface ------------------------------------
 * csm_debug_copy: device-to-device copy of an internal buffer (0 h_bb, 1 h_dec, 2 q_bb, 3 q_dec,
 *   4 attn_bb, 5 attn_dec, 6 mlp_bb, 7 mlp_dec, 8 last_h, 9 c0_logits, 10 cb_logits, 11 samples(i32),
 *   12 fed(i32), 13/14 backbone K/V cache, 15/16 decoder K/V cache); dst NULL = size query.
 * csm_debug_run_phases: execute phases [begin,end) of the frame program at the current cache
 *   length without advancing it (bisecting a frame against the oracle).                        */
int csm_debug_copy(CsmCtx* ctx, int which, void* dst_device, int64_t max_bytes, int64_t* bytes_out, void* stream);
int csm_debug_run_phases(CsmCtx* ctx, const int64_t* ids, const int32_t* mask, int B, int ph_begin, int ph_end,
                         int forced, void* stream);
int csm_debug_set_cache_len(CsmCtx* ctx, int len);
/* Watchdog aid (engine created with CSM_DEBUG_PROGRESS=1): host_out[grid][4] = (phase of the compute warps, step
 * inside it, phase of the weight stream, phase of the L2 prefetcher) of every CTA, copied on `side_stream`
 * (must be a non-blocking stream) so that it can be read while a frame kernel is still running. */
int csm_debug_progress(CsmCtx* ctx, int32_t* host_out, void* side_stream);
/* One decode frame with per-phase stamps.  clocks_host holds (32 + grid) * phases uint64:
 *   [2][phases][16] clock64 of the first and the last CTA: (phase top, body start, body end, phase end,
 *     activations staged, own MMAs done, all MMAs done, first weight chunk resident, input poll succeeded,
 *     poll iterations, unused...), then
 *   [grid][phases] %globaltimer (ns) at the end of every phase of every CTA (skew between CTAs).
 * info_host[4*phases] = (type, epilogue, stack, act_mode) per phase. Synchronises. */
int csm_debug_profile_frame(CsmCtx* ctx, int B, uint64_t* clocks_host, int32_t* info_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step: CSMModel.forward(input_ids, attention_mask, labels=...) and the backward
 * of its loss -- reference modeling_csm.py:292-482 (loss branch :367-465), called by CSMTrainer.compute_loss
 * (train.py:303-326) and differentiated by torch autograd there.
 *
 * The training context owns only workspace (saved activations, per-step fused / transposed weight copies); the
 * parameters and the gradient tensors stay the caller's and are passed on every step, in the reference's layout. */
typedef struct CsmTrain CsmTrain;

/* max_tokens: largest B*S of a step; max_frames: largest number of frames carrying all 32 audio labels in a step
 * (the decoder runs on those only: 1/16 amortisation, processor.py:340-360).  rope tables of `shapes` must cover the
 * longest sequence (backbone) and 33 positions (decoder). */
int csm_train_create(const CsmShapes* shapes, int max_tokens, int max_frames, CsmTrain** out);
int csm_train_destroy(CsmTrain* t);
const char* csm_train_last_error(const CsmTrain* t);
int csm_train_launches(const CsmTrain* t);

/* weights: the current parameters (device, bf16).  grads: where to WRITE d loss / d parameter (same layout, device,
 * bf16; not accumulated), or NULL for the loss alone.
 * ids int64 [B,S,33]; mask int32 [B,S,33] or NULL (= attention_mask=None: all 33 slots summed, nothing padded);
 * labels int64 [B,S,33] with -100 = ignore.  losses: HOST float[3] = loss, backbone_loss, decoder_loss
 * (modeling_csm.py:389,467,471); n_frames: HOST int, frames the decoder ran on (may be NULL).
 * last_h bf16 [B,Hb] and c0_logits bf16 [B,V]: the last position's outputs (:363-365), device, may be NULL.
 * Synchronises the stream (the frame count is needed on the host, as the reference's nonzero() :399). */
int csm_train_step(CsmTrain* t, const CsmWeights* weights, const CsmWeights* grads, const int64_t* ids,
                   const int32_t* mask, const int64_t* labels, int B, int S, float* losses, int* n_frames,
                   void* last_h, void* c0_logits, void* stream);

/* The same step in two calls, so that a data-parallel caller can start averaging the gradients that are already final
 * while the rest of the backward runs: _begin enqueues the forward, the losses and the backward down to (and including)
 * backbone layer split_layer -- the gradients of the decoder, the heads, the projection, the backbone norm and the backbone
 * layers >= split_layer are then final in stream order; _end enqueues the backbone layers below split_layer and the
 * embedding-table gradients, reads the losses back and synchronises.  split_layer in [0, layers]; the weight / gradient
 * pointer arrays of _begin must stay alive until _end returns. */
int csm_train_step_begin(CsmTrain* t, const CsmWeights* weights, const CsmWeights* grads, const int64_t* ids,
                         const int32_t* mask, const int64_t* labels, int B, int S, int split_layer, int* n_frames,
                         void* last_h, void* c0_logits, void* stream);
int csm_train_step_end(CsmTrain* t, float* losses, void* stream);

/* Tests: host copy of a named intermediate of the last step ("bb.0.qkv", "dec.1.attn", "d.bb.x0", ...); host == NULL
 * returns only its size in bytes. */
int csm_train_debug(CsmTrain* t, const char* name, void* host, long long cap, long long* bytes);

#ifdef __cplusplus
}
#endif
#endif /* CSM_B200_H */
