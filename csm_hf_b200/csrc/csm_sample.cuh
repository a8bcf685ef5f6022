// Stochastic top-k sampling of one logits row by one warp (reference: sample_topk +
// _multinomial_sample_one_no_sync, modeling_csm.py:170-189):
//     logits / temperature -> keep every entry >= the k-th largest -> softmax -> argmax(p / Exp(1)).
// argmax(p_i / E_i) with E_i ~ Exp(1) is the Gumbel-max trick, argmax(logit_i / T + G_i), G_i = -log(E_i), so
// no softmax is materialised.  The reference draws E from torch's global generator; here the noise is a
// counter-based hash of (seed, frame, codebook, sequence, index), so every CTA of the frame kernel draws the
// same sample without communicating and a (seed, call sequence) pair reproduces its output.  Parity with the
// reference is therefore distributional (tests/test_gpu_parity.py::test_topk_sampling_distribution).
#pragma once
#include "csm_common.cuh"

// bf16 bits -> 16-bit key that sorts like the value (unsigned compare)
__device__ __forceinline__ uint32_t bf16_sort_key(uint32_t b) { return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u); }
__device__ __forceinline__ float bf16_from_sort_key(uint32_t s) {
  const uint32_t b = (s & 0x8000u) ? (s & 0x7fffu) : (~s & 0xffffu);
  return __uint_as_float(b << 16);
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// standard Gumbel noise for element i of the draw identified by `rkey`
__device__ __forceinline__ float gumbel_noise(unsigned long long rkey, int i) {
  const unsigned long long h = splitmix64(rkey ^ ((unsigned long long)(unsigned)i * 0xD1342543DE82EF95ull));
  const float u = ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  return -__logf(-__logf(u));
}
__device__ __forceinline__ unsigned long long draw_key(unsigned long long seed, unsigned frame, int cb, int seq) {
  return splitmix64(seed ^ splitmix64(((unsigned long long)frame << 32) | ((unsigned long long)(unsigned)cb << 20) |
                                      (unsigned long long)(unsigned)seq));
}

// In a 256-bin histogram (shared memory, this warp's), find the highest bin B such that the number of entries in
// bins > B is < k <= the number in bins >= B.  Returns B; `above` = entries in bins > B.
__device__ __forceinline__ int warp_find_bin(const int* hist, int k, int lane, int& above) {
  int c[8], t = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; t += c[j]; }
  // suffix sum over lanes: entries in bins of lanes > lane
  int suf = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  const int hi = suf - t;                       // entries strictly above this lane's bins
  const bool mine = hi < k && k <= suf;         // the k-th largest falls into one of my 8 bins
  int bin = -1, ab = 0;
  if (mine) {
    int acc = hi;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      if (bin < 0 && acc + c[j] >= k) { bin = lane * 8 + j; ab = acc; }
      acc += c[j];
    }
  }
  const unsigned who = __ballot_sync(0xffffffffu, mine);
  const int src = who ? (__ffs(who) - 1) : 0;   // k > number of entries: fall back to bin 0 (keep everything)
  bin = __shfl_sync(0xffffffffu, bin, src);
  above = __shfl_sync(0xffffffffu, ab, src);
  if (!who) { bin = 0; above = 0; }
  return bin;
}

// keys: V sort keys (16 bit) of one row in shared memory; hist: 256 ints of shared memory owned by this warp.
// Returns the sampled index (same value in every lane).
__device__ __forceinline__ int warp_sample_topk(const unsigned short* keys, int V, int k, float inv_temp,
                                                unsigned long long rkey, int lane, int* hist) {
  // ---- k-th largest key by a two-pass radix select (high byte, then low byte inside that bin)
  for (int i = lane; i < 256; i += 32) hist[i] = 0;
  __syncwarp();
  for (int i = lane; i < V; i += 32) atomicAdd(&hist[keys[i] >> 8], 1);
  __syncwarp();
  int above;
  const int hb = warp_find_bin(hist, k, lane, above);
  __syncwarp();
  for (int i = lane; i < 256; i += 32) hist[i] = 0;
  __syncwarp();
  for (int i = lane; i < V; i += 32) {
    const uint32_t s = keys[i];
    if ((int)(s >> 8) == hb) atomicAdd(&hist[s & 255u], 1);
  }
  __syncwarp();
  int above2;
  const int lb = warp_find_bin(hist, k - above, lane, above2);
  const uint32_t thr = ((uint32_t)hb << 8) | (uint32_t)lb;   // every key >= thr is kept (ties with the k-th included)
  __syncwarp();
  // ---- Gumbel-max over the kept entries
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < V; i += 32) {
    const uint32_t s = keys[i];
    if (s >= thr) {
      const float v = bf16_from_sort_key(s) * inv_temp + gumbel_noise(rkey, i);
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  return bi;
}
