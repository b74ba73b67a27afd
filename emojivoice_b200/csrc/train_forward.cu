// The training-time forward pass of MatchaTTS (models/matcha_tts.py:154-245) after the text encoder, as loss VALUES (no
// backward pass is built): Gaussian log-prior of every (token, frame) pair -> monotonic alignment search (mas.cu) ->
// duration loss -> optional segment cut -> mu_y -> conditional-flow-matching loss through ONE evaluation of the estimator
// with a time per item (components/flow_matching.py:87-118) -> prior loss.  The random draws of the reference (t ~ U[0,1),
// z ~ N(0,1), the cut offsets) are inputs, so a run can be compared with the reference under the same draws.
#include <cmath>

#include "ctx.cuh"

using namespace ev;

extern "C" size_t ev_estimator_workspace_bytes(const ev_ctx* ctx, int B, int T_pad);

namespace {

constexpr int TF_TI = 16;     // tokens per block of the log-prior kernel

// log_prior[b,i,j] * mask, models/matcha_tts.py:190-196 with monotonic_align/__init__.py:13 (value * mask).  The three sums
// run over the feature axis in float32 and are combined in the reference's order: y_square - y_mu_double + mu_square + const.
__global__ void __launch_bounds__(128) log_prior_kernel(const float* __restrict__ mu_x, const float* __restrict__ y, const int* __restrict__ xl,
                                                        const int* __restrict__ yl, int F, int Tx, int Ty, float cst, float* __restrict__ value) {
  extern __shared__ float mu_s[];                      // [TF_TI][F] of this block's tokens, then their mu_square
  const int b = blockIdx.z, i0 = blockIdx.y * TF_TI, j = blockIdx.x * blockDim.x + threadIdx.x;
  float* musq = mu_s + TF_TI * F;
  for (int e = threadIdx.x; e < TF_TI * F; e += blockDim.x) {
    const int ii = e / F, c = e - ii * F;
    mu_s[e] = (i0 + ii < Tx) ? mu_x[((long long)b * F + c) * Tx + i0 + ii] : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x < TF_TI) {
    float s = 0.0f;
    for (int c = 0; c < F; ++c) { const float m = mu_s[threadIdx.x * F + c]; s = fmaf(-0.5f * m, m, s); }   // sum(factor * mu^2)
    musq[threadIdx.x] = s;
  }
  __syncthreads();
  if (j >= Ty) return;
  const float* yc = y + (long long)b * F * Ty + j;
  float ysq = 0.0f;
  for (int c = 0; c < F; ++c) { const float v = yc[(long long)c * Ty]; ysq = fmaf(-0.5f, v * v, ysq); }      // factor^T (y^2)
  const int nx = xl[b], ny = yl[b];
  for (int ii = 0; ii < TF_TI && i0 + ii < Tx; ++ii) {
    float dbl = 0.0f;
    for (int c = 0; c < F; ++c) dbl = fmaf(2.0f * (-0.5f * mu_s[ii * F + c]), yc[(long long)c * Ty], dbl);  // (2 factor mu)^T y
    const float lp = ysq - dbl + musq[ii] + cst;
    value[((long long)b * Tx + i0 + ii) * Ty + j] = (i0 + ii < nx && j < ny) ? lp : 0.0f;
  }
}

// one warp per (b, i): attn row as float, logw_ = log(1e-8 + sum_j attn) * x_mask (matcha_tts.py:204), squared error vs logw
__global__ void __launch_bounds__(256) duration_loss_kernel(const int* __restrict__ path, const float* __restrict__ logw, const int* __restrict__ xl,
                                                            int B, int Tx, int Ty, double* __restrict__ acc) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B * Tx) return;
  const int b = row / Tx, i = row - b * Tx;
  const int* p = path + (long long)row * Ty;
  int n = 0;
  for (int j = lane; j < Ty; j += 32) n += p[j];
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) {
    const float m = i < xl[b] ? 1.0f : 0.0f;
    const float lw_ = logf(1e-8f + (float)n) * m;
    const float d = logw[row] - lw_;
    atomicAdd(acc + 0, (double)(d * d));
  }
}

// one thread per (b, j') frame of the (cut) segment: the token it is aligned to, and the attn column of the output
__global__ void __launch_bounds__(128) frame_tokens_kernel(const int* __restrict__ path, const int* __restrict__ yl_cut, const long long* __restrict__ off,
                                                           int Tx, int Ty, int Tc, int* __restrict__ tok, float* __restrict__ attn_out) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Tc) return;
  const int src = j + (off ? (int)off[b] : 0);
  const bool valid = j < yl_cut[b] && src < Ty;
  int t = -1;
  for (int i = 0; i < Tx; ++i) {
    const int v = valid ? path[((long long)b * Tx + i) * Ty + src] : 0;
    if (v) t = i;
    attn_out[((long long)b * Tx + i) * Tc + j] = (float)v;
  }
  tok[(long long)b * Tc + j] = t;
}

// elementwise over (b, c, j'): x1 (cut target), mu_y (matcha_tts.py:235-236: attn^T mu_x with a 0/1 attn = a gather), the
// flow-matching pair y_t / u (flow_matching.py:108-112) and the prior-loss terms (matcha_tts.py:241)
__global__ void __launch_bounds__(256) cfm_pair_kernel(const float* __restrict__ y, const float* __restrict__ mu_x, const float* __restrict__ z,
                                                       const float* __restrict__ t_rand, const int* __restrict__ tok, const int* __restrict__ yl_cut,
                                                       const long long* __restrict__ off, int F, int Tx, int Ty, int Tc, float sigma_min,
                                                       float* __restrict__ mu_y, float* __restrict__ y_t, float* __restrict__ u, double* __restrict__ acc) {
  const int b = blockIdx.z, c = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  double part = 0.0;
  if (j < Tc) {
    const int src = j + (off ? (int)off[b] : 0);
    const bool valid = j < yl_cut[b];
    // without a cut the target is y as the caller padded it; a cut segment is zero beyond its length (matcha_tts.py:219-227)
    const float x1 = off ? ((valid && src < Ty) ? y[((long long)b * F + c) * Ty + src] : 0.0f) : y[((long long)b * F + c) * Ty + j];
    const int tk = tok[(long long)b * Tc + j];
    const float m = tk >= 0 ? mu_x[((long long)b * F + c) * Tx + tk] : 0.0f;
    const long long o = ((long long)b * F + c) * Tc + j;
    const float t = t_rand[b], zz = z[o];
    mu_y[o] = m;
    y_t[o] = (1.0f - (1.0f - sigma_min) * t) * zz + t * x1;
    u[o] = x1 - (1.0f - sigma_min) * zz;
    if (valid) { const float d = x1 - m; part = (double)(0.5f * (d * d + 1.8378770664093453f)); }   // log(2 pi)
  }
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0 && part != 0.0) atomicAdd(acc + 1, part);
}

// sum over ALL frames, padded ones included, of (v - u)^2: F.mse_loss(..., reduction="sum") at flow_matching.py:114
__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ v, const float* __restrict__ u, long long n, double* __restrict__ acc) {
  double part = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = v[i] - u[i];
    part += (double)(d * d);
  }
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc + 2, part);
}

__global__ void lengths_kernel(const long long* __restrict__ x_lengths, const long long* __restrict__ y_lengths, int B, int out_size, int* xl, int* yl,
                               int* yl_cut, double* acc) {
  // single block: int32 copies of the lengths, the cut lengths out_size + min(y_len - out_size, 0) (matcha_tts.py:221), their sums
  __shared__ long long sx, sy;
  if (threadIdx.x == 0) { sx = 0; sy = 0; }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int nx = (int)x_lengths[b], ny = (int)y_lengths[b];
    const int nc = out_size > 0 ? min(ny, out_size) : ny;
    xl[b] = nx; yl[b] = ny; yl_cut[b] = nc;
    atomicAdd(reinterpret_cast<unsigned long long*>(&sx), (unsigned long long)nx);
    atomicAdd(reinterpret_cast<unsigned long long*>(&sy), (unsigned long long)nc);
  }
  __syncthreads();
  if (threadIdx.x == 0) { acc[0] = acc[1] = acc[2] = 0.0; acc[3] = (double)sx; acc[4] = (double)sy; }
}

__global__ void finish_losses_kernel(const double* acc, int F, int prior_loss, float* losses) {
  losses[0] = (float)(acc[0] / acc[3]);                               // utils/model.py:44-46
  losses[1] = prior_loss ? (float)(acc[1] / (acc[4] * F)) : 0.0f;     // matcha_tts.py:241-242
  losses[2] = (float)(acc[2] / (acc[4] * F));                         // flow_matching.py:114-116
}

__global__ void f32_to_i32_kernel(const float* __restrict__ a, int* __restrict__ o, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = (int)a[i];
}
__global__ void i32_to_i64_kernel(const int* __restrict__ a, long long* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i];
}

struct TfBuffers {
  int *xl, *yl, *yl_cut, *tok, *path;
  float *value, *mu_y, *y_t, *u, *v;
  double* acc;
  void* mas_ws; size_t mas_bytes;
  void* est_ws; size_t est_bytes;
};

void plan_tf(const ev_ctx* ctx, int B, int Tx, int Ty, int Tc, Workspace& w, TfBuffers* t) {
  const int F = ctx->matcha.cfg.n_feats;
  t->xl = w.take<int>(B); t->yl = w.take<int>(B); t->yl_cut = w.take<int>(B);
  t->tok = w.take<int>((size_t)B * Tc);
  t->path = w.take<int>((size_t)B * Tx * Ty);
  t->value = w.take<float>((size_t)B * Tx * Ty);
  t->mu_y = w.take<float>((size_t)B * F * Tc); t->y_t = w.take<float>((size_t)B * F * Tc);
  t->u = w.take<float>((size_t)B * F * Tc); t->v = w.take<float>((size_t)B * F * Tc);
  t->acc = w.take<double>(8);
  t->mas_bytes = ev_maximum_path_workspace_bytes(ctx, B, Tx, Ty);
  t->mas_ws = w.take<char>(t->mas_bytes);
  t->est_bytes = ev_estimator_workspace_bytes(ctx, B, Tc);
  t->est_ws = w.take<char>(t->est_bytes);
}

}  // namespace

extern "C" size_t ev_train_forward_workspace_bytes(const ev_ctx* ctx, int B, int Tx, int Ty, int out_size) {
  if (!ctx || !ctx->matcha.loaded || B <= 0 || Tx <= 0 || Ty <= 0) return 0;
  Workspace w(nullptr, 0);
  TfBuffers t;
  plan_tf(ctx, B, Tx, Ty, out_size > 0 ? out_size : Ty, w, &t);
  return w.off + 256;
}

extern "C" int ev_train_forward(ev_ctx* ctx, const float* mu_x, const float* logw, const int64_t* x_lengths, const float* y,
                                const int64_t* y_lengths, const float* spk_emb, const float* t_rand, const float* z,
                                const float* durations, int out_size, const int64_t* out_offset, int B, int Tx, int Ty,
                                float sigma_min, int prior_loss, int precision, float* losses, float* attn, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!ctx) return EV_ERR_INVALID;
  if (!ctx->matcha.loaded) return fail(ctx, EV_ERR_STATE, "ev_train_forward: matcha weights not loaded");
  if (!mu_x || !logw || !x_lengths || !y || !y_lengths || !t_rand || !z || !losses || !attn || B <= 0 || Tx <= 0 || Ty <= 0)
    return fail(ctx, EV_ERR_INVALID, "ev_train_forward: null argument or empty shape");
  if (out_size > 0 && !out_offset) return fail(ctx, EV_ERR_INVALID, "ev_train_forward: a segment cut needs its offsets");
  const int Tc = out_size > 0 ? out_size : Ty;
  if (Tc % 4) return fail(ctx, EV_ERR_INVALID, "ev_train_forward: the decoder's length must be a multiple of 4 (utils/model.py:14-20)");
  const ev_matcha_cfg& c = ctx->matcha.cfg;
  if (c.n_spks > 1 && !spk_emb) return fail(ctx, EV_ERR_INVALID, "ev_train_forward: spk_emb required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  EV_CUDA(ctx, cudaSetDevice(ctx->device));
  Workspace w(workspace, workspace_bytes);
  TfBuffers t;
  plan_tf(ctx, B, Tx, Ty, Tc, w, &t);
  if (w.overflow || !workspace) return fail(ctx, EV_ERR_STATE, "ev_train_forward: workspace too small");
  ctx->prof_tag = "/train";
  const int F = c.n_feats;
  const long long* off = out_size > 0 ? reinterpret_cast<const long long*>(out_offset) : nullptr;
  EV_LAUNCH(ctx, s, "train_lengths", 0, 24.0 * B,
            (lengths_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const long long*>(x_lengths), reinterpret_cast<const long long*>(y_lengths), B, out_size,
                                              t.xl, t.yl, t.yl_cut, t.acc), cudaGetLastError()));
  if (durations) {            // use_precomputed_durations (matcha_tts.py:185-186): the length regulator of synthesise
    float* attn_full = out_size > 0 ? t.value : attn;
    EV_LAUNCH(ctx, s, "generate_path", 0, 4.0 * B * (double)Tx * Ty, generate_path(durations, t.xl, t.yl, B, Tx, Ty, attn_full, t.tok, s));
    EV_LAUNCH(ctx, s, "f32_to_i32", 0, 8.0 * B * (double)Tx * Ty,
              (f32_to_i32_kernel<<<tc_sm_count() * 4, 256, 0, s>>>(attn_full, t.path, (long long)B * Tx * Ty), cudaGetLastError()));
  } else {
    const float cst = (float)(-0.5 * std::log(2.0 * M_PI) * F);
    dim3 grid(ceil_div(Ty, 128), ceil_div(Tx, TF_TI), B);
    EV_LAUNCH(ctx, s, "log_prior", 4.0 * B * (double)Tx * Ty * F, 4.0 * B * (double)Tx * Ty,
              (log_prior_kernel<<<grid, 128, (size_t)(TF_TI * F + TF_TI) * sizeof(float), s>>>(mu_x, y, t.xl, t.yl, F, Tx, Ty, cst, t.value), cudaGetLastError()));
    EV_TRY(ev_maximum_path(ctx, t.value, t.xl, t.yl, B, Tx, Ty, -1e9f, t.path, t.mas_ws, t.mas_bytes, stream));
    ctx->prof_tag = "/train";
  }
  EV_LAUNCH(ctx, s, "duration_loss", 0, 4.0 * B * (double)Tx * Ty,
            (duration_loss_kernel<<<ceil_div(B * Tx, 8), 256, 0, s>>>(t.path, logw, t.xl, B, Tx, Ty, t.acc), cudaGetLastError()));
  EV_LAUNCH(ctx, s, "frame_tokens", 0, 8.0 * B * (double)Tx * Tc,
            (frame_tokens_kernel<<<dim3(ceil_div(Tc, 128), B), 128, 0, s>>>(t.path, t.yl_cut, off, Tx, Ty, Tc, t.tok, attn), cudaGetLastError()));
  EV_LAUNCH(ctx, s, "cfm_pair", 0, 24.0 * B * (double)F * Tc,
            (cfm_pair_kernel<<<dim3(ceil_div(Tc, 256), F, B), 256, 0, s>>>(y, mu_x, z, t_rand, t.tok, t.yl_cut, off, F, Tx, Ty, Tc, sigma_min, t.mu_y, t.y_t,
                                                                            t.u, t.acc), cudaGetLastError()));
  // int64 lengths of the (cut) segment for the estimator's masks
  long long* yl64 = reinterpret_cast<long long*>(t.value);     // the log-prior is dead by now
  EV_LAUNCH(ctx, s, "i32_to_i64", 0, 12.0 * B, (i32_to_i64_kernel<<<ceil_div(B, 256), 256, 0, s>>>(t.yl_cut, yl64, B), cudaGetLastError()));
  EV_TRY(ev_estimator(ctx, t.y_t, reinterpret_cast<const int64_t*>(yl64), t.mu_y, t_rand, spk_emb, B, Tc, precision, t.v, t.est_ws, t.est_bytes, stream));
  ctx->prof_tag = "/train";
  EV_LAUNCH(ctx, s, "cfm_sq_err", 0, 8.0 * B * (double)F * Tc,
            (sq_err_kernel<<<tc_sm_count() * 4, 256, 0, s>>>(t.v, t.u, (long long)B * F * Tc, t.acc), cudaGetLastError()));
  EV_LAUNCH(ctx, s, "finish_losses", 0, 64.0, (finish_losses_kernel<<<1, 1, 0, s>>>(t.acc, F, prior_loss, losses), cudaGetLastError()));
  return EV_OK;
}
