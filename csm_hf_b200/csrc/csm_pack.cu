// Weight packing for the frame engine.
//
// Source: reference state_dict tensors, bf16 (nn.Linear [out,in] row-major; audio_head slices are
// [in,out], modeling_csm.py:236-240,557).  Destination for a matrix with N packed rows and
// reduction length K, split over G CTAs (csm_geom in csm_types.h):
//
//   for CTA c (rows row0..row0+rows): for k16-tile T: for local row r: 32 bytes =
//       k-permuted 16 bf16  [k0 k1 k8 k9 | k2 k3 k10 k11 | k4 k5 k12 k13 | k6 k7 k14 k15]
//
// so that (a) a CTA's whole slice, and any k-range of it, is one contiguous byte range for a
// bulk copy, and (b) lane (g,t) of a warp reads its mma.m16n8k16 B fragment {b0b1,b2b3} for
// row 8i+g with one 8-byte shared-memory load at [(T*rows + 8i+g)*32 + 8t].
//
// `row_map[n]` gives, for packed row n, the row of the virtual concatenation of up to three
// sources (q|k|v or gate|up), which is how RoPE pairs (i, i+hd/2) and (gate_j, up_j) pairs
// are made adjacent.
#include "csm_common.cuh"

struct PackSrc {
  const bf16* ptr[3];
  int rows[3];
  long long row_stride[3];   // elements between rows
  long long col_stride[3];   // elements between columns (1, or N for a transposed source)
};

__global__ void csm_pack_kernel(PackSrc src, const int* __restrict__ row_map, int N, int K, int gran, int G,
                                bf16* __restrict__ dst) {
  const int ntiles = K / 16;
  const long long total = (long long)N * ntiles;
  for (long long rec = (long long)blockIdx.x * blockDim.x + threadIdx.x; rec < total;
       rec += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(rec / ntiles);   // packed row
    const int T = (int)(rec % ntiles);
    // owner CTA of packed row n
    const int U = N / gran, q = U / G, r = U % G;
    const int unit = n / gran;
    int c;
    if (unit < r * (q + 1)) c = unit / (q + 1);
    else c = r + (q > 0 ? (unit - r * (q + 1)) / q : 0);
    const int start = c * q + (c < r ? c : r);
    const int cnt = q + (c < r ? 1 : 0);
    const int row0 = start * gran, rows = cnt * gran;
    const int lr = n - row0;
    int sr = row_map ? row_map[n] : n;
    int si = 0;
    while (si < 2 && sr >= src.rows[si]) { sr -= src.rows[si]; ++si; }
    const bf16* sp = src.ptr[si] + (long long)sr * src.row_stride[si] + (long long)T * 16 * src.col_stride[si];
    const long long cs = src.col_stride[si];
    bf16 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = sp[k * cs];
    bf16 o[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      o[4 * t + 0] = v[2 * t];
      o[4 * t + 1] = v[2 * t + 1];
      o[4 * t + 2] = v[2 * t + 8];
      o[4 * t + 3] = v[2 * t + 9];
    }
    bf16* dp = dst + (long long)row0 * K + ((long long)T * rows + lr) * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) dp[k] = o[k];
  }
}

extern "C" cudaError_t csm_pack_launch(const PackSrc* src, const int* row_map_dev, int N, int K, int gran, int G,
                                       bf16* dst, cudaStream_t stream) {
  long long total = (long long)N * (K / 16);
  int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  csm_pack_kernel<<<(int)blocks, threads, 0, stream>>>(*src, row_map_dev, N, K, gran, G, dst);
  return cudaGetLastError();
}
