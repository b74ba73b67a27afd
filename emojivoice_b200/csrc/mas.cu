// Monotonic alignment search: the reference's only native component,
// Matcha-TTS/matcha/utils/monotonic_align/core.pyx:11-47 (maximum_path_each / maximum_path_c), as a CUDA kernel.
//
//   forward :  for y in [0, t_y):  for x in [max(0, t_x + y - t_y), min(t_x, y + 1)):
//                 value[x, y] += max( x == y ? NEG : value[x, y-1],
//                                     x == 0 ? (y == 0 ? 0 : NEG) : value[x-1, y-1] )
//   backward:  index = t_x - 1;  for y = t_y-1 .. 0:  path[index, y] = 1;
//                 if index != 0 and (index == y or value[index, y-1] < value[index-1, y-1]): index -= 1
//
// Column y depends on column y-1 only, so one CTA per utterance sweeps the columns with one thread per text position x
// (two shared-memory columns, one __syncthreads per column).  The backward pass needs nothing but the outcome of the
// comparison it repeats, so the forward pass keeps ONE BIT per cell -- "coming from x-1 wins" -- instead of writing the
// value matrix back (t_x * t_y / 8 bytes, in shared memory).  Every cell does exactly the reference's one float32
// add on the same operands (no reassociation, no FMA), so the path is bit-identical to the Cython code.
#include <algorithm>

#include "ctx.cuh"

namespace ev {
namespace {

constexpr int MAS_THREADS = 1024;

__global__ void __launch_bounds__(MAS_THREADS, 1)
mas_kernel(const float* __restrict__ value, const int* __restrict__ t_xs, const int* __restrict__ t_ys, int Tx, int Ty, float neg,
           int* __restrict__ path, uint32_t* __restrict__ gbits, int words, int bits_in_smem) {
  extern __shared__ uint32_t smem_u[];
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int t_x = min(max(t_xs[b], 0), Tx), t_y = min(max(t_ys[b], 0), Ty);
  const float* v = value + (size_t)b * Tx * Ty;
  int* out = path + (size_t)b * Tx * Ty;
  // smem: col[2][Tx] floats | idx[Ty] ints | bits[Tx][words] (when they fit)
  float* col = reinterpret_cast<float*>(smem_u);
  int* idx = reinterpret_cast<int*>(smem_u + 2 * Tx);
  uint32_t* bits = bits_in_smem ? smem_u + 2 * Tx + Ty : gbits + (size_t)b * Tx * words;
  if (t_x <= 0 || t_y <= 0) return;      // the reference loops do nothing useful for empty items (path stays zero)
  for (int i = tid; i < t_x * words; i += nth) bits[i] = 0u;     // rows [0, t_x) of the [Tx][words] bit matrix
  __syncthreads();

  // each thread owns rows x = tid, tid + nth, ... (Tx <= 1024 in practice: one row per thread)
  for (int y = 0; y < t_y; ++y) {
    const float* prev = col + ((y + 1) & 1) * Tx;
    float* cur = col + (y & 1) * Tx;
    const int x_lo = max(0, t_x + y - t_y), x_hi = min(t_x, y + 1);
    for (int x = x_lo + tid; x < x_hi; x += nth) {
      const float v_cur = (x == y) ? neg : prev[x];
      const float v_prev = (x == 0) ? (y == 0 ? 0.0f : neg) : prev[x - 1];
      // Cython's max(v_cur, v_prev): the second argument wins only when strictly greater
      const float m = (v_prev > v_cur) ? v_prev : v_cur;
      cur[x] = __fadd_rn(m, v[(size_t)x * Ty + y]);
      // what the backward pass will ask at (index = x, column y): index == y or value[x, y-1] < value[x-1, y-1]
      const bool dec = x != 0 && (x == y || prev[x] < prev[x - 1]);
      // only thread (x mod nth) ever touches row x: plain read-modify-write
      if (dec) bits[(size_t)x * words + (y >> 5)] |= 1u << (y & 31);
    }
    __syncthreads();
  }
  if (tid == 0) {
    int index = t_x - 1;
    for (int y = t_y - 1; y >= 0; --y) {
      idx[y] = index;
      // rows outside the forward band never set a bit; the reference reads values the band guarantees to exist
      if (index != 0 && ((bits[(size_t)index * words + (y >> 5)] >> (y & 31)) & 1u)) --index;
    }
  }
  __syncthreads();
  for (int y = tid; y < t_y; y += nth) out[(size_t)idx[y] * Ty + y] = 1;
}

}  // namespace
}  // namespace ev

using namespace ev;

extern "C" size_t ev_maximum_path_workspace_bytes(const ev_ctx* ctx, int B, int Tx, int Ty) {
  if (!ctx || B <= 0 || Tx <= 0 || Ty <= 0) return 0;
  return (size_t)B * Tx * ((Ty + 31) / 32) * sizeof(uint32_t) + 256;
}

extern "C" int ev_maximum_path(ev_ctx* ctx, const float* value, const int32_t* t_xs, const int32_t* t_ys, int B, int Tx, int Ty,
                               float max_neg_val, int32_t* path, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return EV_ERR_INVALID;
  if (!value || !t_xs || !t_ys || !path || B <= 0 || Tx <= 0 || Ty <= 0) return fail(ctx, EV_ERR_INVALID, "ev_maximum_path: null argument or empty shape");
  EV_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int words = (Ty + 31) / 32;
  const size_t base = ((size_t)2 * Tx + Ty) * 4, with_bits = base + (size_t)Tx * words * 4;
  const int bits_in_smem = with_bits <= (size_t)200 * 1024;
  if (!bits_in_smem && (!workspace || workspace_bytes < ev_maximum_path_workspace_bytes(ctx, B, Tx, Ty)))
    return fail(ctx, EV_ERR_STATE, "ev_maximum_path: workspace too small");
  if (base > (size_t)200 * 1024) return fail(ctx, EV_ERR_INVALID, "ev_maximum_path: t_x / t_y too large");
  const size_t smem = bits_in_smem ? with_bits : base;
  static DeviceOnce once;
  EV_CUDA(ctx, once.run([&]() { return cudaFuncSetAttribute(mas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
  EV_CUDA(ctx, cudaMemsetAsync(path, 0, (size_t)B * Tx * Ty * sizeof(int32_t), s));
  const int threads = std::min(MAS_THREADS, ((Tx + 31) / 32) * 32);
  EV_LAUNCH(ctx, s, "maximum_path", 0, (double)B * Tx * Ty * 8.0,
            (mas_kernel<<<B, threads, smem, s>>>(value, t_xs, t_ys, Tx, Ty, max_neg_val, path, reinterpret_cast<uint32_t*>(workspace), words,
                                                 bits_in_smem), cudaGetLastError()));
  return EV_OK;
}
