// Weight packing for the frame engine.
//
// Source: reference state_dict tensors, bf16 (nn.Linear [out,in] row-major; audio_head slices are
// [in,out], modeling_csm.py:236-240,557).  Destination for a matrix with N packed rows and
// reduction length K, split over G CTAs (csm_geom in csm_types.h):
//
//   for CTA c (rows row0..row0+rows): for k16-tile T: for k-half h (k 0-7, 8-15): for local row r:
//       16 bytes = the 8 bf16 W[row0+r][16T+8h .. 16T+8h+7]
//
// so that (a) a CTA's whole slice, and any k-range of it, is one contiguous byte range for a
// bulk copy, and (b) the mma.m16n8k16 A fragment of 16 rows x 16 k is one ldmatrix.x4 whose four
// 8x8 matrices (8 consecutive rows of one k-half) are 128 contiguous, bank-conflict-free bytes.
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
    // [tile][k-half][row][8]: each 8x8 ldmatrix tile (8 rows x one k-half) is 128 contiguous bytes
    bf16* dp = dst + (long long)row0 * K + ((long long)T * 2 * rows + lr) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) dp[k] = v[k];
    dp += (long long)rows * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) dp[k] = v[8 + k];
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
