// Bit-exact emulation of ATen's CPU float32 inner-dimension sum (cascade_sum, 4-way ILP, 8-lane vectors).
//
// Why: the reference computes y_lengths = trunc(max(torch.sum(w_ceil, [1,2]), 1)) (matcha_tts.py:124) in
// float32 on whatever device it runs; with length_scale != 1 the terms are not exactly representable and the
// result depends on the summation order.  The CPU oracle is the parity target, so
// the GPU length stage reproduces ATen's order add for add.  Usable from host and device code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EV_HD __host__ __device__ __forceinline__
#else
#define EV_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define __fadd_rn_compat(a, b) __fadd_rn((a), (b))
#else
#define __fadd_rn_compat(a, b) ((a) + (b))
#endif

namespace evsum {

constexpr int kLevels = 4;
constexpr int kIlp = 4;

EV_HD int ceil_log2(int64_t x) {
  if (x <= 2) return 1;
  int n = 0;
  uint64_t v = (uint64_t)x - 1;
  while (v) { ++n; v >>= 1; }
  return n;
}

// row_sum over `size` vectors of W lanes; vector i lane l lives at x[(i*W + l)].  Result: W lane sums.
template <int W>
EV_HD void row_sum_lanes(const float* x, int64_t size, float* out /*[W]*/) {
  const int64_t n4 = size / kIlp;
  int lp = ceil_log2(n4) / kLevels;
  if (lp < 4) lp = 4;
  const int64_t step = (int64_t)1 << lp;
  const int64_t mask = step - 1;
  float acc[kLevels][kIlp][W];
  for (int L = 0; L < kLevels; ++L)
    for (int k = 0; k < kIlp; ++k)
      for (int l = 0; l < W; ++l) acc[L][k][l] = 0.0f;
  int64_t i = 0;
  while (i + step <= n4) {
    for (int64_t j = 0; j < step; ++j, ++i)
      for (int k = 0; k < kIlp; ++k)
        for (int l = 0; l < W; ++l) acc[0][k][l] = __fadd_rn_compat(acc[0][k][l], x[((i * kIlp) + k) * W + l]);
    for (int L = 1; L < kLevels; ++L) {
      for (int k = 0; k < kIlp; ++k)
        for (int l = 0; l < W; ++l) {
          acc[L][k][l] = __fadd_rn_compat(acc[L][k][l], acc[L - 1][k][l]);
          acc[L - 1][k][l] = 0.0f;
        }
      if ((i & (mask << (L * lp))) != 0) break;
    }
  }
  for (; i < n4; ++i)
    for (int k = 0; k < kIlp; ++k)
      for (int l = 0; l < W; ++l) acc[0][k][l] = __fadd_rn_compat(acc[0][k][l], x[((i * kIlp) + k) * W + l]);
  for (int L = 1; L < kLevels; ++L)
    for (int k = 0; k < kIlp; ++k)
      for (int l = 0; l < W; ++l) acc[0][k][l] = __fadd_rn_compat(acc[0][k][l], acc[L][k][l]);
  for (int64_t r = n4 * kIlp; r < size; ++r)
    for (int l = 0; l < W; ++l) acc[0][0][l] = __fadd_rn_compat(acc[0][0][l], x[r * W + l]);
  for (int k = 1; k < kIlp; ++k)
    for (int l = 0; l < W; ++l) acc[0][0][l] = __fadd_rn_compat(acc[0][0][l], acc[0][k][l]);
  for (int l = 0; l < W; ++l) out[l] = acc[0][0][l];
}

// torch.sum of a contiguous float32 row of n elements, as the CPU kernel orders it.
EV_HD float sum_f32(const float* x, int64_t n) {
  constexpr int W = 8;
  if (n < W) {
    float s;
    row_sum_lanes<1>(x, n, &s);
    return s;
  }
  const int64_t nv = n / W;
  float lanes[W];
  row_sum_lanes<W>(x, nv, lanes);
  float s = 0.0f;
  for (int64_t k = nv * W; k < n; ++k) s = __fadd_rn_compat(s, x[k]);
  for (int l = 0; l < W; ++l) s = __fadd_rn_compat(s, lanes[l]);
  return s;
}

}  // namespace evsum
