// fp32 CUDA-core implicit-GEMM conv1d (channel-last), fixed summation order: taps outer, channels inner.
// Used for the text encoder / duration predictor (fp32 protects ceil(exp(logw))) and as the
// EV_PREC_FP32 parity mode of the decoder and vocoder.
#include "conv.cuh"

namespace ev {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

__global__ void __launch_bounds__(THREADS)
conv_simt_kernel(ConvGeom g, const float* __restrict__ x, long long x_ld, long long x_bs,
                 const float* __restrict__ w, int N_pad, Epilogue e) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int b = blockIdx.z, m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int a_row = tid >> 2, a_c = (tid & 3) << 2;
  const int b_k = tid >> 4, b_n = (tid & 15) << 2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const float* xb = x + (long long)b * x_bs;
  for (int tap = 0; tap < g.taps; ++tap) {
    const int t_in = (m0 + a_row) * g.conv_stride + g.tap_off[tap];
    const bool row_ok = (m0 + a_row < g.M) && t_in >= 0 && t_in < g.T_in;
    const float* xrow = xb + (long long)t_in * x_ld;
    const float* wt = w + (size_t)tap * g.C_in * N_pad;
    for (int c0 = 0; c0 < g.C_in; c0 += BK) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && c0 + a_c < g.C_in) av = *reinterpret_cast<const float4*>(xrow + c0 + a_c);
      if (c0 + b_k < g.C_in && n0 + b_n < N_pad)
        bv = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(c0 + b_k) * N_pad + n0 + b_n));
      __syncthreads();
      As[a_c + 0][a_row] = av.x; As[a_c + 1][a_row] = av.y; As[a_c + 2][a_row] = av.z; As[a_c + 3][a_row] = av.w;
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
  }

  float* out_act = reinterpret_cast<float*>(e.out_act);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      int t, co;
      if (!ep_coord(e, r, n, t, co)) continue;
      const float mv = e.mask.at(b, t);
      const float v = ep_value(e, b, t, co, acc[i][j], mv);
      if (e.out_f32) e.out_f32[b * e.f32_bs + (long long)t * e.f32_ld + co] = v;
      if (out_act) out_act[b * e.act_bs + (long long)t * e.act_ld + co] = ep_act(e, co, v, mv);
    }
  }
}
}  // namespace

cudaError_t conv_simt_launch(const ConvGeom& g, const float* x, long long x_ld, long long x_bs, const ConvWeights& w,
                             const Epilogue& e, cudaStream_t stream) {
  if ((g.C_in & 3) || (x_ld & 3) || (w.N_pad & 3)) return cudaErrorInvalidValue;
  dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN), g.B);
  conv_simt_kernel<<<grid, THREADS, 0, stream>>>(g, x, x_ld, x_bs, w.w_f32, w.N_pad, e);
  return cudaGetLastError();
}

}  // namespace ev
