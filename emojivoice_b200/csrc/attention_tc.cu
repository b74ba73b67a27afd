// Decoder self-attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands by TMA).
//
// Semantics: diffusers `Attention` as BasicTransformerBlock calls it (transformer.py:266-271): softmax over
// ALL T_pad keys of  q.k/sqrt(d) + mask_j  where the 0/1 frame mask is a float attn_mask, i.e. valid keys get +1.
//
// One CTA = 128 queries of one (batch item, head); keys are walked in blocks of 64, twice:
//   pass 1   S = Q K^T (tcgen05, 128x64 fp32 tile in TMEM)  ->  running row maximum (thread = query row)
//   pass 2   S again -> p = exp2(s - max) -> row sums, P as a bf16 K-major tile in shared memory (manual 128B swizzle,
//            generic->async proxy fence) -> O += P V (tcgen05; V is consumed MN-major straight from its TMA tile)
// The two-pass form needs no accumulator rescaling (the score MMA is ~6 % of the softmax cost, so recomputing it is
// cheaper than a TMEM read-modify-write of O) and is exact with respect to the maximum.
// Warps: 0..3 softmax (TMEM lane quarter = warp), 4 = TMA producer, 5 = TMEM allocator + MMA issuer.
// Two CTAs fit one SM (97 KB shared memory, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
#include <cudaTypedefs.h>

#include "conv.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace ev {
using namespace tc;
namespace {

constexpr int BQ = 128, BKV = 64, HD = 64, STAGES = 3;
constexpr int THREADS = 192;
constexpr int Q_BYTES = BQ * HD * 2, KV_BYTES = BKV * HD * 2, P_BYTES = BQ * BKV * 2;
constexpr int SMEM_BYTES = 1024 + Q_BYTES + STAGES * 2 * KV_BYTES + 2 * P_BYTES;
constexpr uint32_t TMEM_COLS = 256;   // S0 [0,64) S1 [64,128) O [128,192)

struct Params {
  int T, H, inner;
  float c1, c2;            // scale*log2(e), log2(e)
  const int* lens; int len_shift;
  bf16* out; long long out_ld, out_bs;
  int skip_padded_queries;   // query blocks without a valid row are left unwritten (the caller never reads those rows)
};

__global__ void __launch_bounds__(THREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tm, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, o_full, kv_full[STAGES], kv_empty[STAGES], s_full[2], s_empty[2], p_full[2], p_empty[2];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_s = base, kv_s = base + Q_BYTES, p_s = kv_s + STAGES * 2 * KV_BYTES;
  uint8_t* p_gen = smem_raw + (p_s - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int n_kb = (p.T + BKV - 1) / BKV, n_it = 2 * n_kb;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    mbar_init(&q_full, 1); mbar_init(&o_full, 1);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_trigger();
  pdl_wait();
  // CTA-uniform: the whole query block lies in the padding and its output is not wanted
  const bool skip = p.skip_padded_queries && p.lens && (q0 << p.len_shift) >= __ldg(p.lens + b);

  if (skip) {
  } else if (warp == 4) {
    if (lane == 0) {
      // ---------------- TMA producer: Q once, then K (pass 1) and K+V (pass 2) through a 3-stage ring
      mbar_expect_tx(&q_full, Q_BYTES);
      tma_load_3d(q_s, &tm, &q_full, h * HD, q0, b);
      tma_load_3d(q_s + Q_BYTES / 2, &tm, &q_full, h * HD, q0 + 64, b);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % STAGES, j = it < n_kb ? it : it - n_kb;
        mbar_wait(&kv_empty[st], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
        const uint32_t k_t = kv_s + (uint32_t)(st * 2 * KV_BYTES);
        if (it < n_kb) {
          mbar_expect_tx(&kv_full[st], KV_BYTES);
          tma_load_3d(k_t, &tm, &kv_full[st], p.inner + h * HD, j * BKV, b);
        } else {
          mbar_expect_tx(&kv_full[st], 2 * KV_BYTES);
          tma_load_3d(k_t, &tm, &kv_full[st], p.inner + h * HD, j * BKV, b);
          tma_load_3d(k_t + KV_BYTES, &tm, &kv_full[st], 2 * p.inner + h * HD, j * BKV, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------- MMA issuer: warp-uniform control flow, one elected lane issues (descriptors stay in uniform registers)
    constexpr uint32_t idesc_s = make_idesc(BQ, BKV, 0, 0);   // S = Q K^T : both operands K-major
    constexpr uint32_t idesc_o = make_idesc(BQ, HD, 0, 1);    // O += P V  : V is MN-major (rows = keys)
    auto issue_pv = [&](int it) {
      const int j = it - n_kb, pb = j & 1, st = it % STAGES;
      mbar_wait(&p_full[pb], (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t p_t = p_s + (uint32_t)(pb * P_BYTES), v_t = kv_s + (uint32_t)(st * 2 * KV_BYTES + KV_BYTES);
      const uint64_t da = make_smem_desc(p_t), db = make_smem_desc_ex(v_t, 1024, 8192, 2);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)   // A: +32 B along K (keys); B: +16 key rows of 128 B
          umma_bf16(tmem_base + 128, da + (uint64_t)(2 * k), db + (uint64_t)(k * 128), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(&p_empty[pb]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    for (int it = 0; it < n_it; ++it) {
      const int st = it % STAGES, sb = it & 1;
      mbar_wait(&kv_full[st], (uint32_t)(it / STAGES) & 1u);
      mbar_wait(&s_empty[sb], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t k_t = kv_s + (uint32_t)(st * 2 * KV_BYTES);
      const uint64_t dq = make_smem_desc(q_s), dk = make_smem_desc(k_t);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(sb * BKV), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[sb]);
        if (it < n_kb) umma_commit(&kv_empty[st]);
      }
      __syncwarp();
      if (it > n_kb) issue_pv(it - 1);
    }
    issue_pv(n_it - 1);
    if (elect_one()) umma_commit(&o_full);
    __syncwarp();
  } else {
    // ---------------- softmax warps: thread = query row
    const int row = warp * 32 + lane, t = q0 + row;
    const int len = p.lens ? __ldg(p.lens + b) : 0x7fffffff;
    const int n_valid = p.lens ? min(p.T, (len + (1 << p.len_shift) - 1) >> p.len_shift) : p.T;   // valid keys form a prefix
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float m = -INFINITY;
    for (int it = 0; it < n_kb; ++it) {
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (uint32_t)(it >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(sb * BKV + half * 32), raw);
        const int k0 = it * BKV + half * 32, lim_valid = n_valid - k0, lim_in = p.T - k0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = fmaf(__uint_as_float(raw[c]), p.c1, c < lim_valid ? p.c2 : 0.0f);
          m = fmaxf(m, c < lim_in ? x : -INFINITY);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    float l = 0.0f;
    for (int j = 0; j < n_kb; ++j) {
      const int it = n_kb + j, sb = it & 1, pb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)(it >> 1) & 1u);
      mbar_wait(&p_empty[pb], ((uint32_t)(j >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      uint8_t* prow = p_gen + pb * P_BYTES + row * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(sb * BKV + half * 32), raw);
        const int k0 = j * BKV + half * 32, lim_valid = n_valid - k0, lim_in = p.T - k0;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float x0 = fmaf(__uint_as_float(raw[c]), p.c1, c < lim_valid ? p.c2 : 0.0f);
          const float x1 = fmaf(__uint_as_float(raw[c + 1]), p.c1, (c + 1) < lim_valid ? p.c2 : 0.0f);
          const float p0 = c < lim_in ? ex2f(x0 - m) : 0.0f;
          const float p1 = (c + 1) < lim_in ? ex2f(x1 - m) : 0.0f;
          l += p0 + p1;
          __nv_bfloat162 v2 = __floats2bfloat162_rn(p0, p1);
          pk[c >> 1] = *reinterpret_cast<uint32_t*>(&v2);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // 16-byte chunk (half*4 + i) of the row, 128B-swizzled like a TMA tile
          const int ch = (half * 4 + i) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + ch * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
      }
      tcgen05_fence_before();
      fence_proxy_async();        // the P tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_empty[sb]); mbar_arrive(&p_full[pb]); }
    }
    mbar_wait(&o_full, 0);
    tcgen05_fence_after();
    const float inv = 1.0f / l;
    bf16* orow = p.out + b * p.out_bs + (long long)t * p.out_ld + h * HD;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + (uint32_t)(128 + half * 32), raw);
      if (t < p.T) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(raw[8 * i + 2 * e]) * inv, __uint_as_float(raw[8 * i + 2 * e + 1]) * inv);
            w[e] = *reinterpret_cast<uint32_t*>(&v2);
          }
          *reinterpret_cast<uint4*>(orow + half * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

cudaError_t attention_tc(const AttnTcArgs& a, cudaStream_t s, std::string* err) {
  if (a.D != HD || a.inner != a.H * HD) { if (err) *err = "attention_tc: head_dim must be 64"; return cudaErrorInvalidValue; }
  if ((a.ld & 7) || (a.bs & 7) || (reinterpret_cast<uintptr_t>(a.qkv) & 15) || (a.out_ld & 7) || (a.out_bs & 7) ||
      (reinterpret_cast<uintptr_t>(a.out) & 15)) {
    if (err) *err = "attention_tc: tensors must be 16-byte aligned / strided";
    return cudaErrorInvalidValue;
  }
  CUtensorMap tm;
  if (!tc_encode_bf16_map(&tm, a.qkv, (uint64_t)(3 * a.inner), (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.ld * 2, (uint64_t)a.bs * 2,
                          HD, BKV, 128, err))
    return cudaErrorInvalidValue;
  static DeviceOnce once;
  cudaError_t ce_attr = once.run([&]() { return cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); });
  if (ce_attr != cudaSuccess) return ce_attr;
  Params p;
  p.T = a.T; p.H = a.H; p.inner = a.inner;
  p.c1 = a.scale * 1.4426950408889634f; p.c2 = 1.4426950408889634f;
  p.lens = a.lens; p.len_shift = a.len_shift;
  p.out = a.out; p.out_ld = a.out_ld; p.out_bs = a.out_bs;
  p.skip_padded_queries = a.skip_padded_queries;
  dim3 grid(ceil_div(a.T, BQ), a.H, a.B);
  return launch_pdl(attn_tc_kernel, grid, dim3(THREADS), (size_t)SMEM_BYTES, s, tm, p);
}

}  // namespace ev
