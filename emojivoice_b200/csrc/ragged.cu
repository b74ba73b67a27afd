// Compact tile lists for ragged batches (see RaggedPlanner in conv.cuh): one small kernel per distinct tile geometry
// turns the per-item lengths into the list of (item, m-tile) pairs that hold rows an utterance's valid audio depends on.
#include <algorithm>

#include "conv.cuh"

namespace ev {
namespace {

// every block recomputes the (tiny) prefix sum over the items and fills its own slice of the list
__global__ void __launch_bounds__(256) ragged_table_kernel(const int* __restrict__ lens, int B, int margin, int rpf, int shift,
                                                           int tile_rows, int M, int* __restrict__ table) {
  __shared__ int start[kRaggedMaxB + 1];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    // frames valid at this level: t with (t << shift) < len  (the decoder's half-rate mask is mask[:, :, ::2])
    const long long frames = ((long long)max(lens[b], 0) + (1 << shift) - 1) >> shift;
    const long long need = min((long long)M, (frames + margin) * rpf);
    start[b + 1] = (int)((need + tile_rows - 1) / tile_rows);
  }
  __syncthreads();
  if (threadIdx.x < 32) {                 // warp scan, 32 items per step
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + threadIdx.x;
      int v = b < B ? start[b + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if ((int)threadIdx.x >= o) v += n; }
      if (b < B) start[b + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (threadIdx.x == 0) { start[0] = 0; if (blockIdx.x == 0) table[0] = carry; }
  }
  __syncthreads();
  const int total = start[B];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = B - 1;              // last b with start[b] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (start[mid] <= i) lo = mid; else hi = mid - 1;
    }
    table[1 + i] = (lo << 16) | (i - start[lo]);
  }
}

}  // namespace

const int* RaggedPlanner::table(int tile_rows, int M, cudaStream_t s) {
  if (!lens || !arena || B <= 0 || B > kRaggedMaxB || B >= 32768 || tile_rows <= 0) return nullptr;
  const int m_tiles = ceil_div(M, tile_rows);
  if (m_tiles >= 65536) return nullptr;
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].rpf == rows_per_frame && cache[i].shift == len_shift && cache[i].tile_rows == tile_rows && cache[i].M == M && cache[i].margin == margin) return cache[i].table;
  const size_t ints = align_up((size_t)B * m_tiles + 1, 64);
  if (n_cache >= 32 || arena_off + ints > arena_ints) return nullptr;
  int* t = arena + arena_off;
  // an ordinary launch (no programmatic serialization): every later kernel of the stream sees the finished table
  cudaStream_t ts = (side && ready) ? side : s;
  ragged_table_kernel<<<std::max(1, std::min(64, ceil_div(B * m_tiles, 2048))), 256, 0, ts>>>(lens, B, margin, rows_per_frame, len_shift, tile_rows, M, t);
  if (cudaGetLastError() != cudaSuccess) return nullptr;
  if (ts != s && (cudaEventRecord(ready, ts) != cudaSuccess || cudaStreamWaitEvent(s, ready, 0) != cudaSuccess)) return nullptr;
  arena_off += ints;
  if (launch_counter) ++*launch_counter;
  cache[n_cache++] = Entry{rows_per_frame, len_shift, tile_rows, M, margin, t};
  return t;
}

cudaError_t ragged_build_table(const int* lens, int B, int margin, int rows_per_frame, int len_shift, int tile_rows, int M, int* table,
                               cudaStream_t s) {
  if (B <= 0 || B > kRaggedMaxB || B >= 32768 || ceil_div(M, tile_rows) >= 65536) return cudaErrorInvalidValue;
  ragged_table_kernel<<<std::max(1, std::min(64, ceil_div(B * ceil_div(M, tile_rows), 2048))), 256, 0, s>>>(lens, B, margin, rows_per_frame, len_shift, tile_rows, M, table);
  return cudaGetLastError();
}

}  // namespace ev
