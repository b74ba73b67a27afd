// One ResnetBlock1D of the flow-matching U-Net (decoder.py:32-61) per launch, on tcgen05, plus the pre-LayerNorm of the
// transformer block that follows (transformer.py:262):
//
//     h1 = conv1(x)                    -> GroupNorm(8) statistics over ALL T_pad frames of the item
//     a  = (Mish(GN1(h1)) * m + temb) * m                                   (bf16 operand of conv2)
//     h2 = conv2(a)                    -> GroupNorm statistics
//     xr = Mish(GN2(h2)) * m + res_conv(x)                                  (fp32 residual stream of the transformer block)
//     n  = LayerNorm(xr)                                                    (bf16 operand of the QKV projection)
//
// Round 1 ran this as five launches (res_conv, conv1, gn_apply, conv2, gn_apply_ln), each ~15-20 us of which ~3 us were
// tensor work: the decoder was launch / prologue / epilogue-drain bound ("Where a small decoder conv spends its
// time").  Here every CTA owns ONE tile of mb x 128 frames x all 256 channels for the whole block:
//   * the conv accumulators (mb x 256 fp32 columns) stay in TMEM across the GroupNorm: pass 1 reads them for the
//     statistics (fp64 atomics per (item, group)), a GRID BARRIER makes the statistics of every CTA visible, pass 2 reads
//     them again and normalises -- the fp32 conv output never touches HBM;
//   * conv2's operand `a` goes through global memory (L2) because its taps need the neighbour tiles' rows: a second grid
//     barrier, then TMA loads it back (generic-proxy stores -> fence.proxy.async -> barrier -> TMA);
//   * Mish(GN2(h2)) * m is written back INTO the accumulator (tcgen05.st) and res_conv(x) is accumulated onto it by the
//     tensor core (accumulate = 1), so the residual add costs nothing and `r` is never materialised;
//   * LayerNorm statistics are taken per row from the same TMEM tile (thread = row), exchanged through shared memory.
// All CTAs must be co-resident (grid = B * ceil(T / (128 mb)) <= SM count, one CTA per SM): the launch is cooperative.
// `mode 1` stops after the first apply (final_block: conv -> GN -> Mish -> mask, decoder.py:431).
//
// Warp roles (19 warps): 0 activation-tile TMA producer, 1 weight-tile TMA producer (weights are constants: it runs free),
// 2 MMA issuer (owns TMEM), 3-18 sixteen epilogue warps (TMEM lane quadrant = warp & 3, four column slots per quadrant).
// Shared memory: weight ring 4 x 32 KB ([64 k x 256 n] bf16, 128B swizzle), activation ring (2-4 haloed tiles) ALIASED
// with the epilogue's transpose buffers (they are never live together: an epilogue phase starts when its GEMM's last MMA
// has completed, and the next GEMM's activation loads wait for the grid barrier behind that epilogue phase).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>

#include "conv.cuh"
#include "tc_ptx.cuh"

namespace ev {
using namespace tc;
namespace {

constexpr int RN_C = 256;
constexpr int RN_EPI_WARPS = 16;
constexpr int RN_ROLE_WARPS = 3;
constexpr int RN_THREADS = 32 * (RN_ROLE_WARPS + RN_EPI_WARPS);
constexpr int RN_EPI_THREADS = 32 * RN_EPI_WARPS;
constexpr int RN_B_TILE = RN_C * 128;            // [64 k x 256 n] bf16
constexpr int RN_B_SLOTS = 4;
constexpr int RN_MAX_A_SLOTS = 4;
constexpr int RN_STAGE_LD = 36;
constexpr int RN_STAGE_WARP = 32 * RN_STAGE_LD * 4;
constexpr int RN_AREGION = RN_EPI_WARPS * RN_STAGE_WARP;   // 73728 B: activation ring / transpose buffers
constexpr int RN_ROWSTAT = 2 * 128 * 4 * 2 * 4;            // LayerNorm partial sums [m-block][row][slot][2]
constexpr int RN_SMEM = 1024 + RN_B_SLOTS * RN_B_TILE + RN_AREGION + RN_ROWSTAT;
constexpr int RN_ACT_PITCH = 80;                           // bytes per staged bf16 row (64 + 16: conflict-free 16-byte access)

struct RnMaps { CUtensorMap x, a, w1, w2, wr; };

struct RnParams {
  int B, T, mb, m_tiles, n_cta;
  int kc_in;                                   // 64-channel K-chunks of the block input (conv1 and res_conv)
  int a_boxes, a_box_rows, a_slot_bytes, a_slots;
  int mode;
  const int* lens; int len_shift;
  const float *bias1, *bias2, *bias_r, *g1, *b1, *g2, *b2, *temb, *ln_g, *ln_b;
  double* gn1; double* gn2;                    // [B][8][2] (sum, sum of squares), zeroed by the caller
  unsigned int* bar;                           // three grid-barrier counters, zeroed by the caller
  bf16* a_buf; long long a_ld, a_bs;           // conv2's operand (mode 0) / the block's output (mode 1)
  float* xr; bf16* n_out;                      // (b, t, 256) dense
  float eps_gn, eps_ln;
  int trace;
};

__device__ unsigned long long g_rn_trace[32];
#define RN_TR(i) do { if (p.trace && blockIdx.x == 0 && lane == 0) g_rn_trace[(i)] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ uint32_t rd_hi(uint32_t sbo, uint32_t layout) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t rd_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t rd_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __noinline__ void grid_barrier_timeout() {
  printf("emojivoice_b200: resnet_tc grid barrier timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
// one thread per CTA: wait until `n` CTAs have arrived at `ctr` (a bug or a non-resident CTA must trap, not hang the device)
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int n) {
  const long long t0 = clock64();
  while (ld_acquire_gpu(ctr) < n) {
    __nanosleep(32);
    if (clock64() - t0 > (2ll << 30)) grid_barrier_timeout();     // ~1 s
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(RN_EPI_THREADS) : "memory"); }

// x * tanh(softplus(x)) = x * w / (w + 2), w = e^x (e^x + 2)   (same formulation as gn_apply256, elementwise.cu)
__device__ __forceinline__ float mish_fast(float x) {
  const float n = __expf(fminf(x, 20.0f));
  const float w = n * (n + 2.0f);
  return x > 20.0f ? x : x * __fdividef(w, w + 2.0f);
}

__global__ void __launch_bounds__(RN_THREADS, 1)
resnet_tc_kernel(const __grid_constant__ RnMaps maps, const __grid_constant__ RnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[RN_MAX_A_SLOTS], a_empty[RN_MAX_A_SLOTS], b_full[RN_B_SLOTS], b_empty[RN_B_SLOTS];
  __shared__ __align__(8) uint64_t acc_full, epi_done;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_s = base, a_s = base + RN_B_SLOTS * RN_B_TILE;
  uint8_t* stage_gen = base_gen + RN_B_SLOTS * RN_B_TILE;                       // aliases the activation ring
  float* rowstat = reinterpret_cast<float*>(base_gen + RN_B_SLOTS * RN_B_TILE + RN_AREGION);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = (int)blockIdx.x / p.m_tiles, mt = (int)blockIdx.x - b * p.m_tiles;
  const int m0 = mt * p.mb * 128;
  const int vmb = min(p.mb, (p.T - m0 + 127) >> 7);                             // m-blocks that hold a frame
  const bool full = p.mode == 0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1) : "memory");
    for (int s = 0; s < RN_MAX_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < RN_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&epi_done, RN_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_trigger();
  pdl_wait();
  RN_TR(0);

  if (warp == 0) {
    // ---------------- activation tiles: conv1 reads x, conv2 reads `a` (behind grid barrier 2), res_conv reads x again
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 1;
      const uint32_t a_bytes = (uint32_t)(p.a_boxes * p.a_box_rows * 128), box_bytes = (uint32_t)(p.a_box_rows * 128);
      auto load_a = [&](const CUtensorMap* map, int kc) {
        mbar_wait(&a_empty[sa], pa);
        mbar_expect_tx(&a_full[sa], a_bytes);
        const uint32_t dst = a_s + (uint32_t)(sa * p.a_slot_bytes);
        tma_load_3d(dst, map, &a_full[sa], kc * 64, m0 - 1, b);
        if (p.a_boxes > 1) tma_load_3d(dst + box_bytes, map, &a_full[sa], kc * 64, m0 - 1 + p.a_box_rows, b);
        if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
      };
      for (int kc = 0; kc < p.kc_in; ++kc) load_a(&maps.x, kc);
      if (full) {
        // `a` of every CTA is complete (and this CTA's transpose buffers are idle) once all CTAs passed barrier 2
        grid_wait(p.bar + 1, (unsigned)p.n_cta);
        asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy stores of the other CTAs -> this thread's TMA reads
        RN_TR(6);
        for (int kc = 0; kc < RN_C / 64; ++kc) load_a(&maps.a, kc);
        for (int kc = 0; kc < p.kc_in; ++kc) load_a(&maps.x, kc);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- weight tiles in the issuer's order (constants: no dependency on anything)
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 1;
      auto load_b = [&](const CUtensorMap* map, int kc, int tap) {
        mbar_wait(&b_empty[sb], pb);
        mbar_expect_tx(&b_full[sb], (uint32_t)RN_B_TILE);
        tma_load_3d(b_s + (uint32_t)(sb * RN_B_TILE), map, &b_full[sb], kc * 64, 0, tap);
        if (++sb == RN_B_SLOTS) { sb = 0; pb ^= 1u; }
      };
      for (int kc = 0; kc < p.kc_in; ++kc)
        for (int j = 0; j < 3; ++j) load_b(&maps.w1, kc, j);
      if (full) {
        for (int kc = 0; kc < RN_C / 64; ++kc)
          for (int j = 0; j < 3; ++j) load_b(&maps.w2, kc, j);
        for (int kc = 0; kc < p.kc_in; ++kc) load_b(&maps.wr, kc, 0);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---------------- MMA issuer: warp-uniform control flow, one elected lane issues.  M = 128, N = 256, K = 16.
    constexpr uint32_t idesc = make_idesc(128, RN_C);
    const uint32_t hi = rd_hi(1024u, 2u);
    const uint32_t a_lo0 = rd_lo(a_s), b_lo0 = rd_lo(b_s);
    const uint32_t a_step16 = (uint32_t)p.a_slot_bytes >> 4;
    constexpr uint32_t b_step16 = (uint32_t)RN_B_TILE >> 4, mb_step16 = (128u * 128u) >> 4, row16 = 128u >> 4;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    // one GEMM: `kchunks` activation tiles, `taps` weight tiles each; tap j reads the haloed tile from row `row0 + j` on
    auto gemm = [&](int kchunks, int taps, int row0, uint32_t acc_first) {
      uint32_t acc = acc_first;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&a_full[sa], pa);
        tcgen05_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_step16 + (uint32_t)row0 * row16;
        for (int j = 0; j < taps; ++j) {
          mbar_wait(&b_full[sb], pb);
          tcgen05_fence_after();
          const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_step16;
          if (elect_one()) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              if (m < vmb) {
                const uint32_t d = tmem_base + (uint32_t)(m * RN_C), am = a_lo + (uint32_t)j * row16 + (uint32_t)m * mb_step16;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16(d, rd_join(hi, am + 2u * ks), rd_join(hi, b_lo + 2u * ks), idesc, (acc | (uint32_t)ks) ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[sb]);
          }
          __syncwarp();
          acc = 1;
          if (++sb == RN_B_SLOTS) { sb = 0; pb ^= 1u; }
        }
        if (elect_one()) umma_commit(&a_empty[sa]);
        __syncwarp();
        if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    };
    gemm(p.kc_in, 3, 0, 0u);                       // conv1: taps at tile rows 0, 1, 2 (tile row 0 = frame m0 - 1)
    RN_TR(1);
    if (full) {
      mbar_wait(&epi_done, 0);                     // apply-1 has read the accumulator for the last time
      tcgen05_fence_after();
      gemm(RN_C / 64, 3, 0, 0u);                   // conv2
      RN_TR(7);
      mbar_wait(&epi_done, 1);                     // Mish(GN2(h2)) * m sits in the accumulator
      tcgen05_fence_after();
      gemm(p.kc_in, 1, 1, 1u);                     // + res_conv(x): the centre row of the haloed tile, accumulated on top
      RN_TR(10);
    }
  } else {
    // ---------------- sixteen epilogue warps
    const int ew = warp - RN_ROLE_WARPS, q = warp & 3, slot = ew >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_blk = vmb * 8;                                         // 32-column blocks of this lane quadrant
    float* wstage = reinterpret_cast<float*>(stage_gen + ew * RN_STAGE_WARP);
    const int len_b = p.lens ? __ldg(p.lens + b) : 0x7fffffff;
    const bool leader = ew == 0 && lane == 0;
    const double inv_n = 1.0 / (32.0 * (double)p.T);

    // GroupNorm statistics of (accumulator + bias) over the tile's frames t < T: one fp64 atomic pair per warp and block
    auto stats_pass = [&](const float* bias, double* gn) {
#pragma unroll 1
      for (int blk = slot; blk < n_blk; blk += 4) {
        const int m = blk >> 3, cb = blk & 7;
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
        const bool ok = m0 + m * 128 + q * 32 + lane < p.T;
        float s = 0.0f, qq = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cb * 32 + j));
          const float v0 = __uint_as_float(raw[j]) + bv.x, v1 = __uint_as_float(raw[j + 1]) + bv.y;
          const float v2 = __uint_as_float(raw[j + 2]) + bv.z, v3 = __uint_as_float(raw[j + 3]) + bv.w;
          s += (v0 + v1) + (v2 + v3);
          qq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
        }
        if (!ok) { s = 0.0f; qq = 0.0f; }
        s = warp_sum(s); qq = warp_sum(qq);
        if (lane == 0) {
          double* dst = gn + ((long long)b * 8 + cb) * 2;
          atomicAdd(dst, (double)s);
          atomicAdd(dst + 1, (double)qq);
        }
      }
    };
    // every CTA has added its statistics once `ctr` reaches n_cta
    auto grid_sync = [&](unsigned int* ctr) {
      epi_bar();
      if (leader) {
        __threadfence();
        atomicAdd(ctr, 1u);
        grid_wait(ctr, (unsigned)p.n_cta);
        __threadfence();
      }
      epi_bar();
    };
    // mean / rstd of group `cb` of this item (same arithmetic as gn_apply256: fp64 sums -> fp32 mean, rstd)
    auto group_stat = [&](const double* gn, int cb, float& mean, float& rstd) {
      const double* src = gn + ((long long)b * 8 + cb) * 2;
      const double s = __ldcg(src), qq = __ldcg(src + 1);
      const double mu = s * inv_n;
      double var = qq * inv_n - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + (double)p.eps_gn));
    };
    // 32 bf16 values per lane (thread = row) -> coalesced 16-byte stores of the 32 x 32 block at dst (row stride ld elements)
    auto store_bf16_block = [&](const uint32_t (&pk)[16], bf16* dst, long long ld, int t0) {
      uint8_t* brow = reinterpret_cast<uint8_t*>(wstage) + lane * RN_ACT_PITCH;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(brow + i * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      __syncwarp();
      const int rsub = lane >> 2, ch = lane & 3;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(wstage) + ch * 16;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rl = rsub + it * 8, t = t0 + rl;
        if (t < p.T) *reinterpret_cast<uint4*>(dst + (long long)t * ld + ch * 8) = *reinterpret_cast<const uint4*>(src + rl * RN_ACT_PITCH);
      }
      __syncwarp();
    };

    // ======== conv1 done: statistics -> grid barrier 1 -> apply 1
    mbar_wait(&acc_full, 0);
    tcgen05_fence_after();
    if (ew == 0) RN_TR(2);
    stats_pass(p.bias1, p.gn1);
    grid_sync(p.bar + 0);
    if (ew == 0) RN_TR(3);
#pragma unroll 1
    for (int blk = slot; blk < n_blk; blk += 4) {
      const int m = blk >> 3, cb = blk & 7;
      float mean, rstd;
      group_stat(p.gn1, cb, mean, rstd);
      uint32_t raw[32];
      tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
      const int t = m0 + m * 128 + q * 32 + lane;
      const bool valid = t < p.T && (t << p.len_shift) < len_b;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int c = cb * 32 + j;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias1 + c));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.g1 + c)), be = __ldg(reinterpret_cast<const float4*>(p.b1 + c));
        float4 te = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full) te = *reinterpret_cast<const float4*>(p.temb + c);
        const float s0 = rstd * ga.x, s1 = rstd * ga.y, s2 = rstd * ga.z, s3 = rstd * ga.w;
        float v0 = mish_fast(fmaf(__uint_as_float(raw[j]) + bv.x, s0, be.x - mean * s0)) + te.x;
        float v1 = mish_fast(fmaf(__uint_as_float(raw[j + 1]) + bv.y, s1, be.y - mean * s1)) + te.y;
        float v2 = mish_fast(fmaf(__uint_as_float(raw[j + 2]) + bv.z, s2, be.z - mean * s2)) + te.z;
        float v3 = mish_fast(fmaf(__uint_as_float(raw[j + 3]) + bv.w, s3, be.w - mean * s3)) + te.w;
        if (!valid) { v0 = v1 = v2 = v3 = 0.0f; }                      // (Mish * m + temb) * m: a select, the row may hold anything
        __nv_bfloat162 lo2 = __floats2bfloat162_rn(v0, v1), hi2 = __floats2bfloat162_rn(v2, v3);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
        pk[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hi2);
      }
      store_bf16_block(pk, p.a_buf + b * p.a_bs + cb * 32, p.a_ld, m0 + m * 128 + q * 32);
    }
    if (ew == 0) RN_TR(4);
    if (!full) {                                                       // final_block: done
      tcgen05_fence_before();
    } else {
      tcgen05_fence_before();
      asm volatile("fence.proxy.async;" ::: "memory");                 // this thread's stores of `a` -> later TMA reads (any CTA)
      __syncwarp();
      if (lane == 0) mbar_arrive(&epi_done);                           // conv2 may overwrite the accumulator
      // grid barrier 2 (arrive only): the activation producers wait for it before loading `a`
      epi_bar();
      if (leader) { __threadfence(); atomicAdd(p.bar + 1, 1u); }
      if (ew == 0) RN_TR(5);

      // ======== conv2 done: statistics -> grid barrier 3 -> Mish(GN2) * m back into the accumulator
      mbar_wait(&acc_full, 1);
      tcgen05_fence_after();
      if (ew == 0) RN_TR(8);
      stats_pass(p.bias2, p.gn2);
      grid_sync(p.bar + 2);
#pragma unroll 1
      for (int blk = slot; blk < n_blk; blk += 4) {
        const int m = blk >> 3, cb = blk & 7;
        float mean, rstd;
        group_stat(p.gn2, cb, mean, rstd);
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
        const int t = m0 + m * 128 + q * 32 + lane;
        const bool valid = t < p.T && (t << p.len_shift) < len_b;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = cb * 32 + j;
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias2 + c));
          const float4 ga = __ldg(reinterpret_cast<const float4*>(p.g2 + c)), be = __ldg(reinterpret_cast<const float4*>(p.b2 + c));
          const float s0 = rstd * ga.x, s1 = rstd * ga.y, s2 = rstd * ga.z, s3 = rstd * ga.w;
          const float v0 = mish_fast(fmaf(__uint_as_float(raw[j]) + bv.x, s0, be.x - mean * s0));
          const float v1 = mish_fast(fmaf(__uint_as_float(raw[j + 1]) + bv.y, s1, be.y - mean * s1));
          const float v2 = mish_fast(fmaf(__uint_as_float(raw[j + 2]) + bv.z, s2, be.z - mean * s2));
          const float v3 = mish_fast(fmaf(__uint_as_float(raw[j + 3]) + bv.w, s3, be.w - mean * s3));
          raw[j] = valid ? __float_as_uint(v0) : 0u; raw[j + 1] = valid ? __float_as_uint(v1) : 0u;
          raw[j + 2] = valid ? __float_as_uint(v2) : 0u; raw[j + 3] = valid ? __float_as_uint(v3) : 0u;
        }
        tmem_st32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&epi_done);                           // res_conv may accumulate on top
      if (ew == 0) RN_TR(9);

      // ======== xr = accumulator + res bias -> fp32 stream (coalesced through the transpose buffer) + LayerNorm partial sums
      mbar_wait(&acc_full, 0);
      tcgen05_fence_after();
      if (ew == 0) RN_TR(11);
      const int sub = lane >> 3, cl = (lane & 7) * 4;
      float ls[2] = {0.0f, 0.0f}, lq[2] = {0.0f, 0.0f};
#pragma unroll 1
      for (int blk = slot; blk < n_blk; blk += 4) {
        const int m = blk >> 3, cb = blk & 7;
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
        float s = 0.0f, qq = 0.0f;
        float* srow_w = wstage + lane * RN_STAGE_LD;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias_r + cb * 32 + j));
          const float v0 = __uint_as_float(raw[j]) + bv.x, v1 = __uint_as_float(raw[j + 1]) + bv.y;
          const float v2 = __uint_as_float(raw[j + 2]) + bv.z, v3 = __uint_as_float(raw[j + 3]) + bv.w;
          s += (v0 + v1) + (v2 + v3);
          qq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
          *reinterpret_cast<float4*>(srow_w + j) = make_float4(v0, v1, v2, v3);
        }
        if (m == 0) { ls[0] += s; lq[0] += qq; } else { ls[1] += s; lq[1] += qq; }
        __syncwarp();
        float* dst = p.xr + ((long long)b * p.T) * RN_C + cb * 32 + cl;
        const int t0 = m0 + m * 128 + q * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rl = u * 4 + sub, t = t0 + rl;
          if (t < p.T) *reinterpret_cast<float4*>(dst + (long long)t * RN_C) = *reinterpret_cast<const float4*>(wstage + rl * RN_STAGE_LD + cl);
        }
        __syncwarp();
      }
      for (int m = 0; m < vmb; ++m) {
        float* rs = rowstat + (((m * 128) + q * 32 + lane) * 4 + slot) * 2;
        rs[0] = ls[m]; rs[1] = lq[m];
      }
      epi_bar();
      float mean_r[2], rstd_r[2];
      for (int m = 0; m < vmb; ++m) {
        const float4* rs = reinterpret_cast<const float4*>(rowstat + ((m * 128) + q * 32 + lane) * 8);
        const float4 u0 = rs[0], u1 = rs[1];
        const float mu = ((u0.x + u0.z) + (u1.x + u1.z)) * (1.0f / RN_C);
        const float var = fmaxf(((u0.y + u0.w) + (u1.y + u1.w)) * (1.0f / RN_C) - mu * mu, 0.0f);
        mean_r[m] = mu;
        rstd_r[m] = 1.0f / sqrtf(var + p.eps_ln);
      }
      // ======== n = LayerNorm(xr) -> bf16 operand of the QKV projection
#pragma unroll 1
      for (int blk = slot; blk < n_blk; blk += 4) {
        const int m = blk >> 3, cb = blk & 7;
        uint32_t raw[32];
        tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
        const float mu = m == 0 ? mean_r[0] : mean_r[1], rs = m == 0 ? rstd_r[0] : rstd_r[1];
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = cb * 32 + j;
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias_r + c));
          const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_g + c)), be = __ldg(reinterpret_cast<const float4*>(p.ln_b + c));
          const float v0 = (__uint_as_float(raw[j]) + bv.x - mu) * rs * ga.x + be.x;
          const float v1 = (__uint_as_float(raw[j + 1]) + bv.y - mu) * rs * ga.y + be.y;
          const float v2 = (__uint_as_float(raw[j + 2]) + bv.z - mu) * rs * ga.z + be.z;
          const float v3 = (__uint_as_float(raw[j + 3]) + bv.w - mu) * rs * ga.w + be.w;
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(v0, v1), hi2 = __floats2bfloat162_rn(v2, v3);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
          pk[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hi2);
        }
        store_bf16_block(pk, p.n_out + ((long long)b * p.T) * RN_C + cb * 32, RN_C, m0 + m * 128 + q * 32);
      }
      tcgen05_fence_before();
      if (ew == 0) RN_TR(12);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int g_rn_mode = -1;    // EV_RN_FUSE: 0 = five launches per ResNet block (round-1 path), 1 = fused (default)
int g_rn_coop = 1;     // EV_RN_COOP=0: plain launch (co-residency then rests on grid <= SM count alone)

}  // namespace

cudaError_t resnet_tc_read_trace(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_rn_trace, sizeof(unsigned long long) * std::min(n, 32));
}

// m-blocks per CTA for B items of T frames, or 0 when the tiles cannot all be co-resident
int resnet_tc_plan(int B, int T) {
  const int sms = tc_sm_count();
  if (B <= 0 || T <= 0) return 0;
  if ((long long)B * ceil_div(T, 128) <= sms) return 1;
  if ((long long)B * ceil_div(T, 256) <= sms) return 2;
  return 0;
}

bool resnet_tc_supported(const ConvWeights& conv1, const ConvWeights* conv2, const ConvWeights* res, int B, int T) {
  if (g_rn_mode < 0) {
    const char* v = getenv("EV_RN_FUSE");
    g_rn_mode = v ? atoi(v) : 1;
    const char* c = getenv("EV_RN_COOP");
    g_rn_coop = c ? atoi(c) : 1;
  }
  auto ok3 = [](const ConvWeights& w) {
    return w.w_bf16 && w.bias && w.taps == 3 && w.N == RN_C && w.N_pad_tc == RN_C && w.conv_stride == 1 && w.dilation == 1 && w.pad == 1 &&
           !w.transposed && w.K_pad % 64 == 0 && w.K_pad >= w.C_in;
  };
  if (g_rn_mode == 0 || !ok3(conv1) || resnet_tc_plan(B, T) == 0) return false;
  if (conv2) {
    if (!ok3(*conv2) || conv2->C_in != RN_C || !res) return false;
    if (!res->w_bf16 || !res->bias || res->taps != 1 || res->N != RN_C || res->N_pad_tc != RN_C || res->C_in != conv1.C_in || res->K_pad != conv1.K_pad)
      return false;
  }
  return true;
}

cudaError_t resnet_tc_launch(const ResnetTcArgs& a, cudaStream_t s, std::string* err) {
  const ConvWeights& w1 = *a.conv1;
  const bool full = a.conv2 != nullptr;
  if (!resnet_tc_supported(w1, a.conv2, a.res, a.B, a.T)) {
    if (err) *err = "resnet_tc: unsupported layer shapes or too many tiles for one co-resident wave";
    return cudaErrorInvalidValue;
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if ((a.x_ld & 7) || (a.x_bs & 7) || (a.a_ld & 7) || (a.a_bs & 7) || !al16(a.x) || !al16(a.a_buf) || !al16(a.gn_g1) || !al16(a.gn_b1) ||
      (full && (!al16(a.xr) || !al16(a.n_out) || !al16(a.temb) || !al16(a.gn_g2) || !al16(a.gn_b2) || !al16(a.ln_g) || !al16(a.ln_b)))) {
    if (err) *err = "resnet_tc: tensors must be 16-byte aligned";
    return cudaErrorInvalidValue;
  }
  RnParams p{};
  p.B = a.B; p.T = a.T;
  p.mb = resnet_tc_plan(a.B, a.T);
  p.m_tiles = ceil_div(a.T, 128 * p.mb);
  p.n_cta = a.B * p.m_tiles;
  p.kc_in = ceil_div(w1.C_in, 64);
  const int a_rows = p.mb * 128 + 2;
  p.a_boxes = a_rows > 256 ? 2 : 1;
  p.a_box_rows = (int)align_up((size_t)ceil_div(a_rows, p.a_boxes), 8);
  p.a_slot_bytes = (int)align_up((size_t)p.a_boxes * p.a_box_rows * 128, 1024);
  p.a_slots = std::min(RN_MAX_A_SLOTS, RN_AREGION / p.a_slot_bytes);
  p.mode = full ? 0 : 1;
  p.lens = a.lens; p.len_shift = a.len_shift;
  p.bias1 = w1.bias; p.g1 = a.gn_g1; p.b1 = a.gn_b1; p.temb = a.temb;
  p.gn1 = a.gn_sum1; p.gn2 = a.gn_sum2; p.bar = a.barriers;
  p.a_buf = a.a_buf; p.a_ld = a.a_ld; p.a_bs = a.a_bs;
  p.eps_gn = 1e-5f; p.eps_ln = 1e-5f;
  { static const int tr = []() { const char* v = getenv("EV_RN_TRACE"); return v ? atoi(v) : 0; }(); p.trace = tr; }
  RnMaps maps;
  bool ok = tc_encode_bf16_map(&maps.x, a.x, (uint64_t)w1.C_in, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.x_ld * 2, (uint64_t)a.x_bs * 2,
                               64u, (uint32_t)p.a_box_rows, 128, err);
  ok = ok && tc_encode_bf16_map(&maps.w1, w1.w_bf16, (uint64_t)w1.K_pad, (uint64_t)w1.N_pad_tc, 3, (uint64_t)w1.K_pad * 2,
                                (uint64_t)w1.K_pad * w1.N_pad_tc * 2, 64u, (uint32_t)RN_C, 128, err);
  if (full) {
    const ConvWeights& w2 = *a.conv2;
    const ConvWeights& wr = *a.res;
    p.bias2 = w2.bias; p.bias_r = wr.bias; p.g2 = a.gn_g2; p.b2 = a.gn_b2; p.ln_g = a.ln_g; p.ln_b = a.ln_b;
    p.xr = a.xr; p.n_out = a.n_out;
    ok = ok && tc_encode_bf16_map(&maps.a, a.a_buf, (uint64_t)RN_C, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.a_ld * 2, (uint64_t)a.a_bs * 2,
                                  64u, (uint32_t)p.a_box_rows, 128, err);
    ok = ok && tc_encode_bf16_map(&maps.w2, w2.w_bf16, (uint64_t)w2.K_pad, (uint64_t)w2.N_pad_tc, 3, (uint64_t)w2.K_pad * 2,
                                  (uint64_t)w2.K_pad * w2.N_pad_tc * 2, 64u, (uint32_t)RN_C, 128, err);
    ok = ok && tc_encode_bf16_map(&maps.wr, wr.w_bf16, (uint64_t)wr.K_pad, (uint64_t)wr.N_pad_tc, 1, (uint64_t)wr.K_pad * 2,
                                  (uint64_t)wr.K_pad * wr.N_pad_tc * 2, 64u, (uint32_t)RN_C, 128, err);
  } else {
    maps.a = maps.x; maps.w2 = maps.w1; maps.wr = maps.w1;
  }
  if (!ok) return cudaErrorInvalidValue;
  static DeviceOnce once;
  cudaError_t ce = once.run([&]() { return cudaFuncSetAttribute(resnet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RN_SMEM); });
  if (ce != cudaSuccess) return ce;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_cta); cfg.blockDim = dim3(RN_THREADS); cfg.dynamicSmemBytes = RN_SMEM; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = g_rn_coop ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, resnet_tc_kernel, maps, p);
}

}  // namespace ev
