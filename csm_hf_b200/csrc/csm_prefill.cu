// Context prefill kernels: the first CSMModel.forward call of generate() (modeling_csm.py:508-517 with
// S = T context frames).  Row-parallel kernels around the dense projections: masked 33-way
// embedding gather-sum, RMSNorm, causal GQA flash attention on mma.sync tensor cores (padding-aware).
// The projections themselves, with RoPE + KV-cache write, SwiGLU and the residual adds fused into their
// tails, are the tcgen05 GEMMs of csm_gemm.cu.
#include "csm_common.cuh"
#include "csm_sample.cuh"

// ---------------------------------------------------------------- K1: masked 33-way gather-sum
// out[r] = sum_slot mask[r][slot] * emb(slot, ids[r][slot])   (modeling_csm.py:261-282, 327-334)
// One CTA per frame row; each thread owns 8 contiguous features (one 16-byte load per slot),
// fp32 accumulation in slot order (audio 0..31, then text), one bf16 rounding.
__global__ void csm_embed_sum_kernel(const long long* __restrict__ ids, const int* __restrict__ mask, int default_mask,
                                     const bf16* __restrict__ audio_emb, const bf16* __restrict__ text_emb, int V, int H,
                                     bf16* __restrict__ out, int rows) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  __shared__ long long s_tok[CSM_NQ + 1];
  __shared__ int s_mk[CSM_NQ + 1];
  if (threadIdx.x <= CSM_NQ) {
    int slot = threadIdx.x;
    s_tok[slot] = ids[(size_t)r * (CSM_NQ + 1) + slot];
    // default_mask 1: all slots present (attention_mask=None, modeling_csm.py:330-331); 2: audio slots only
    s_mk[slot] = mask ? mask[(size_t)r * (CSM_NQ + 1) + slot] : (default_mask == 1 ? 1 : (slot < CSM_NQ ? 1 : 0));
  }
  __syncthreads();
  for (int c8 = threadIdx.x; c8 < H / 8; c8 += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int slot = 0; slot <= CSM_NQ; ++slot) {
      const int mk = s_mk[slot];
      if (mk == 0) continue;
      const bf16* row = slot < CSM_NQ ? audio_emb + (size_t)(s_tok[slot] + (long long)slot * V) * H
                                      : text_emb + (size_t)s_tok[slot] * H;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + c8);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
      const float f = (float)mk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += bf_lo(u[i]) * f;
        acc[2 * i + 1] += bf_hi(u[i]) * f;
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]);
    o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]);
    o.w = pack_bf16(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out + (size_t)r * H)[c8] = o;
  }
}

// ---------------------------------------------------------------- RMSNorm over rows (hf modeling_llama.py:62-67)
__global__ void csm_rmsnorm_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, float eps, int H,
                                        bf16* __restrict__ y, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  const bf16* src = x + (size_t)r * H;
  uint4 v[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = (i * 32 + lane) * 8;
    if (idx < H) {
      v[i] = *reinterpret_cast<const uint4*>(src + idx);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int q = 0; q < 4; ++q) { float a = bf_lo(u[q]), b = bf_hi(u[q]); ss += a * a + b * b; }
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = (i * 32 + lane) * 8;
    if (idx < H) {
      uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + idx));
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
      const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
      uint4 o;
      uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float y0 = bfround(bf_lo(u[q]) * rstd), y1 = bfround(bf_hi(u[q]) * rstd);
        ou[q] = pack_bf16(bf_lo(ww[q]) * y0, bf_hi(ww[q]) * y1);
      }
      *reinterpret_cast<uint4*>(y + (size_t)r * H + idx) = o;
    }
  }
}

// ---------------------------------------------------------------- gate / up rows interleaved (create time)
// dst row 2j = a row j (gate_j), dst row 2j+1 = b row j (up_j): a GEMM tile of the interleaved matrix holds both halves
// of SwiGLU for its columns (csm_gemm.cu, EPI_SWIGLU).
__global__ void csm_interleave_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, int K,
                                           bf16* __restrict__ dst) {
  const int j = blockIdx.x;
  const uint4* sa = reinterpret_cast<const uint4*>(a + (size_t)j * K);
  const uint4* sb = reinterpret_cast<const uint4*>(b + (size_t)j * K);
  uint4* da = reinterpret_cast<uint4*>(dst + (size_t)(2 * j) * K);
  uint4* db = reinterpret_cast<uint4*>(dst + (size_t)(2 * j + 1) * K);
  for (int i = threadIdx.x; i < K / 8; i += blockDim.x) { da[i] = sa[i]; db[i] = sb[i]; }
}

// ---------------------------------------------------------------- frame-valid bytes of a padded batch
// valid[r] = any(mask[r][0..32] != 0)  (modeling_csm.py:337-342: hf_attention_mask = mask.sum(-1) > 0)
__global__ void csm_frame_valid_kernel(const int* __restrict__ mask, int rows, unsigned char* __restrict__ valid) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int any = 0;
  for (int s = 0; s <= CSM_NQ; ++s) any |= mask[(size_t)r * (CSM_NQ + 1) + s];
  valid[r] = any != 0;
}

// ---------------------------------------------------------------- the last position's hidden row per sequence,
// handed to the frame kernel as tagged words (bf16 | tag of the last backbone phase, csm_common.cuh)
__global__ void csm_take_last_rows_kernel(const bf16* __restrict__ h, int S, int H, uint32_t* __restrict__ dst, int b0,
                                          uint32_t tag, int plain) {
  const int b = blockIdx.x;
  const unsigned short* src = reinterpret_cast<const unsigned short*>(h + ((size_t)b * S + (S - 1)) * H);
  if (plain) {   // general kernels: the residual stream is a plain bf16 row
    unsigned short* d = reinterpret_cast<unsigned short*>(dst) + (size_t)(b0 + b) * (H + 8);   // rows H + 8 apart
    for (int i = threadIdx.x; i < H; i += blockDim.x) d[i] = src[i];
    return;
  }
  uint32_t* d = dst + (size_t)(b0 + b) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) d[i] = (tag << 16) | (uint32_t)src[i];
}

// tagged rows -> plain bf16 rows (debug / tests)
__global__ void csm_untag_rows_kernel(const uint32_t* __restrict__ src, long long src_stride, int cols,
                                      bf16* __restrict__ dst) {
  const int r = blockIdx.x;
  unsigned short* d = reinterpret_cast<unsigned short*>(dst + (size_t)r * cols);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) d[i] = (unsigned short)(src[(size_t)r * src_stride + i] & 0xffffu);
}

__global__ void csm_i64_to_i32_kernel(const long long* __restrict__ src, int* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int)src[i];
}
__global__ void csm_i32_to_i64_kernel(const int* __restrict__ src, long long* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (long long)src[i];
}

// ---------------------------------------------------------------- causal GQA flash attention (prefill), hd = 64
// grid (ceil(S/128), heads, nseq); 8 warps x 16 query rows.  K/V come from the cache (already rotated), q from the qkv
// rows.  K/V blocks of 64 keys are double-buffered in shared memory with cp.async (the next block is in flight while
// this one is multiplied); K and V fragments by ldmatrix.  Online softmax in fp32; P is rounded to bf16 for the PV MMA
// as every flash kernel (incl. the SDPA kernels the reference dispatches to) does.
// Padding (valid != null): keys of padded frames are hidden; a query that sees no key gets a zero output.
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_plain(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool pred) {
  const int n = pred ? 16 : 0;   // (src-size 0: the 16 bytes are zero-filled)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(256) csm_flash_prefill_kernel(const bf16* __restrict__ qkv, int S, int pos0, int b0,
                                                                int heads, int kv, const bf16* __restrict__ kc,
                                                                const bf16* __restrict__ vc, int layer, int Bmax,
                                                                int Tcap, float scale,
                                                                const unsigned char* __restrict__ valid,
                                                                bf16* __restrict__ out) {
  constexpr int HD = 64, BQ = 128, BK = 64, LDS = 72;
  __shared__ __align__(16) bf16 sK[2][BK * LDS];
  __shared__ __align__(16) bf16 sV[2][BK * LDS];
  __shared__ unsigned char sOk[2][BK];   // padded batches: key of the block visible (valid == null: all visible)
  // the heaviest query blocks (longest causal prefix) are scheduled first
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, bl = blockIdx.z;
  const int b = b0 + bl;
  const int kvh = head / (heads / kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int width = (heads + 2 * kv) * HD;
  const int q0 = qt * BQ + warp * 16;        // first query row (sequence-local) of this warp
  const int r_lo = q0 + g, r_hi = q0 + g + 8;
  // Q fragments (A operand), 4 k16-tiles
  uint32_t qa[4][4];
  {
    const bf16* qlo = qkv + ((size_t)bl * S + r_lo) * width + head * HD;
    const bf16* qhi = qkv + ((size_t)bl * S + r_hi) * width + head * HD;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int col = kt * 16 + 2 * t;
      qa[kt][0] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qlo + col) : 0u;
      qa[kt][1] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qhi + col) : 0u;
      qa[kt][2] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qlo + col + 8) : 0u;
      qa[kt][3] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qhi + col + 8) : 0u;
    }
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[j][q] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const int p_lo = pos0 + r_lo, p_hi = pos0 + r_hi;              // absolute positions of this lane's rows
  const int last_key = min(pos0 + qt * BQ + BQ - 1, pos0 + S - 1);  // causal limit of the CTA
  const size_t kvbase = (((size_t)layer * Bmax + b) * kv + kvh) * (size_t)Tcap * HD;
  const float sl2 = scale * 1.4426950408889634f;
  const int nblk = last_key / BK + 1;
  // asynchronous copy of key block kb into buffer kb & 1: 64 rows x 8 chunks of 16 bytes for K and for V
  auto load_block = [&](int kb) {
    const int k0 = kb * BK, buf = kb & 1;
#pragma unroll
    for (int i = threadIdx.x; i < BK * HD / 8; i += 256) {
      const int kr = i >> 3, c8 = i & 7;
      const bool ok = k0 + kr <= last_key;
      const size_t off = kvbase + (size_t)(ok ? k0 + kr : 0) * HD + c8 * 8;
      cp_async16(&sK[buf][kr * LDS + c8 * 8], kc + off, ok);
      cp_async16(&sV[buf][kr * LDS + c8 * 8], vc + off, ok);
    }
    if (threadIdx.x < BK) {
      const int kk = k0 + (int)threadIdx.x - pos0;   // sequence-local index of the key (valid covers this call's rows)
      sOk[buf][threadIdx.x] = (valid == nullptr || kk < 0 || kk >= S) ? 1 : valid[(size_t)bl * S + kk];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_block(0);
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * BK, buf = kb & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                       // block kb has landed for every thread; everyone is done with buffer buf ^ 1
    if (kb + 1 < nblk) load_block(kb + 1);
    if (k0 > pos0 + q0 + 15) continue;     // whole block is in this warp's future
    const bf16* cK = sK[buf];
    const bf16* cV = sV[buf];
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sc[j][q] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        // four 8x8 matrices: keys 8j..8j+7 x dims 32 kp + {0, 8, 16, 24} = the B fragments of k16-tiles 2 kp, 2 kp + 1
        uint32_t kf[4];
        ldmatrix_x4_plain(kf, cK + (8 * j + (lane & 7)) * LDS + 32 * kp + 8 * (lane >> 3));
        mma16816(sc[j], qa[2 * kp], kf[0], kf[1]);
        mma16816(sc[j], qa[2 * kp + 1], kf[2], kf[3]);
      }
    }
    // mask + online softmax (base-2 exponent with the scale folded in)
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = k0 + 8 * j + 2 * t;
      const bool ok0 = sOk[buf][8 * j + 2 * t] != 0, ok1 = sOk[buf][8 * j + 2 * t + 1] != 0;
      if (key > p_lo || !ok0) sc[j][0] = -INFINITY;
      if (key + 1 > p_lo || !ok1) sc[j][1] = -INFINITY;
      if (key > p_hi || !ok0) sc[j][2] = -INFINITY;
      if (key + 1 > p_hi || !ok1) sc[j][3] = -INFINITY;
      mx_lo = fmaxf(mx_lo, fmaxf(sc[j][0], sc[j][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(sc[j][2], sc[j][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float c_lo = (mx_lo == -INFINITY) ? 1.f : exp2f((m_lo - mx_lo) * sl2);
    const float c_hi = (mx_hi == -INFINITY) ? 1.f : exp2f((m_hi - mx_hi) * sl2);
    m_lo = mx_lo;
    m_hi = mx_hi;
    l_lo *= c_lo;
    l_hi *= c_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= c_lo; o[j][1] *= c_lo; o[j][2] *= c_hi; o[j][3] *= c_hi; }
    const float ms_lo = (mx_lo == -INFINITY) ? 0.f : mx_lo * sl2, ms_hi = (mx_hi == -INFINITY) ? 0.f : mx_hi * sl2;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(sc[j][0] * sl2 - ms_lo), p1 = exp2f(sc[j][1] * sl2 - ms_lo);
      const float p2 = exp2f(sc[j][2] * sl2 - ms_hi), p3 = exp2f(sc[j][3] * sl2 - ms_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      const int kk = j >> 1;
      if ((j & 1) == 0) { pa[kk][0] = pack_bf16(p0, p1); pa[kk][1] = pack_bf16(p2, p3); }
      else { pa[kk][2] = pack_bf16(p0, p1); pa[kk][3] = pack_bf16(p2, p3); }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jd = 0; jd < 8; ++jd) {
        uint32_t b0r, b1r;
        ldmatrix_x2_trans(b0r, b1r, cV + (16 * kk + (lane & 15)) * LDS + 8 * jd);
        mma16816(o[jd], pa[kk], b0r, b1r);
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  // a query that saw no key at all (a padded frame) gets a zero output, as torch's SDPA returns for a fully masked row
  const float i_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, i_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  bf16* olo = out + ((size_t)bl * S + r_lo) * (heads * HD) + head * HD;
  bf16* ohi = out + ((size_t)bl * S + r_hi) * (heads * HD) + head * HD;
#pragma unroll
  for (int jd = 0; jd < 8; ++jd) {
    const int col = 8 * jd + 2 * t;
    if (r_lo < S) *reinterpret_cast<uint32_t*>(olo + col) = pack_bf16(o[jd][0] * i_lo, o[jd][1] * i_lo);
    if (r_hi < S) *reinterpret_cast<uint32_t*>(ohi + col) = pack_bf16(o[jd][2] * i_hi, o[jd][3] * i_hi);
  }
}

// ---------------------------------------------------------------- stand-alone top-k sampler (tests)
// One warp per logits row, the same device code the frame kernel runs (csm_sample.cuh): row r is drawn with the
// noise key (seed, frame 0, codebook 0, sequence r).
__global__ void __launch_bounds__(256) csm_sample_rows_kernel(const bf16* __restrict__ logits, int rows, int V, int Vs,
                                                              int topk, float inv_temp, unsigned long long seed,
                                                              long long* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned short* keys = reinterpret_cast<unsigned short*>(smem) + (size_t)warp * Vs;
  int* hist = reinterpret_cast<int*>(smem + (size_t)8 * Vs * 2) + warp * 256;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const unsigned short* src = reinterpret_cast<const unsigned short*>(logits + (size_t)r * V);
  for (int i = lane; i < V; i += 32) keys[i] = (unsigned short)bf16_sort_key(src[i]);
  __syncwarp();
  const int idx = warp_sample_topk(keys, V, topk < V ? topk : V, inv_temp, draw_key(seed, 0u, 0, r), lane, hist);
  if (lane == 0) out[r] = idx;
}

// ---------------------------------------------------------------- host launchers
extern "C" {

cudaError_t csm_sample_rows_launch(const bf16* logits, int rows, int V, int topk, float inv_temp, unsigned long long seed,
                                   long long* out, cudaStream_t st) {
  const int Vs = (V + 3) / 4 * 4;
  const size_t smem = (size_t)8 * Vs * 2 + 8 * 256 * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute((const void*)csm_sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  csm_sample_rows_kernel<<<(rows + 7) / 8, 256, smem, st>>>(logits, rows, V, Vs, topk, inv_temp, seed, out);
  return cudaGetLastError();
}

cudaError_t csm_embed_sum_launch(const long long* ids, const int* mask, int default_mask, const bf16* audio_emb,
                                 const bf16* text_emb, int V, int H, bf16* out, int rows, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  csm_embed_sum_kernel<<<rows, 256, 0, st>>>(ids, mask, default_mask, audio_emb, text_emb, V, H, out, rows);
  return cudaGetLastError();
}
cudaError_t csm_rmsnorm_rows_launch(const bf16* x, const bf16* w, float eps, int H, bf16* y, int rows, cudaStream_t st) {
  csm_rmsnorm_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, w, eps, H, y, rows);
  return cudaGetLastError();
}
cudaError_t csm_interleave_rows_launch(const bf16* a, const bf16* b, int rows, int K, bf16* dst, cudaStream_t st) {
  csm_interleave_rows_kernel<<<rows, 256, 0, st>>>(a, b, K, dst);
  return cudaGetLastError();
}
cudaError_t csm_frame_valid_launch(const int* mask, int rows, unsigned char* valid, int* any_pad, cudaStream_t st) {
  (void)any_pad;
  csm_frame_valid_kernel<<<(rows + 255) / 256, 256, 0, st>>>(mask, rows, valid);
  return cudaGetLastError();
}
cudaError_t csm_take_last_rows_launch(const bf16* h, int S, int H, uint32_t* dst, int b0, int nseq, uint32_t tag,
                                      int plain, cudaStream_t st) {
  csm_take_last_rows_kernel<<<nseq, 256, 0, st>>>(h, S, H, dst, b0, tag, plain);
  return cudaGetLastError();
}
cudaError_t csm_untag_rows_launch(const uint32_t* src, long long src_stride, int cols, int rows, bf16* dst, cudaStream_t st) {
  csm_untag_rows_kernel<<<rows, 256, 0, st>>>(src, src_stride, cols, dst);
  return cudaGetLastError();
}
cudaError_t csm_i64_to_i32_launch(const long long* src, int* dst, int n, cudaStream_t st) {
  csm_i64_to_i32_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t csm_i32_to_i64_launch(const int* src, long long* dst, int n, cudaStream_t st) {
  csm_i32_to_i64_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t csm_flash_prefill_launch(const bf16* qkv, int S, int pos0, int b0, int nseq, int heads, int kv,
                                     const bf16* kc, const bf16* vc, int layer, int Bmax, int Tcap, float scale,
                                     const unsigned char* valid, bf16* out, cudaStream_t st) {
  dim3 grid((S + 127) / 128, heads, nseq);
  csm_flash_prefill_kernel<<<grid, 256, 0, st>>>(qkv, S, pos0, b0, heads, kv, kc, vc, layer, Bmax, Tcap, scale, valid, out);
  return cudaGetLastError();
}

}  // extern "C"
