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
// tensor work.  Here every CTA owns ONE tile of R = 128 * mb consecutive frames x all 256 channels for the whole block:
//   * tiles overlap by two frames: tile row 0 and row R - 1 are HALO rows (the neighbour tiles own those frames), so conv2's
//     operand `a` -- its taps reach one row to either side -- never leaves the SM: the epilogue warps write it, hand-swizzled,
//     straight into the shared-memory K-major operand planes conv2's MMAs read.  No global round trip, no grid barrier there;
//   * the tiles of one item form a THREAD-BLOCK CLUSTER (<= 8 CTAs).  The conv accumulators (R x 256 fp32) stay in TMEM across
//     the GroupNorm: pass 1 reads them for the statistics of the OWNED rows, every CTA then
//     pushes its 16 partial sums into the shared memory of every CTA of the cluster (st.async + mbarrier complete_tx over
//     DSMEM: one ~0.5 us hop instead of global atomics + a grid barrier), sums them in rank order (deterministic) and pass 2
//     reads the accumulators again and normalises.  The fp32 conv outputs never touch HBM, nothing is zeroed beforehand,
//     and the grid may be any number of clusters (no co-residency requirement);
//   * normalisation constants are folded per channel once per CTA (scale = rstd * gamma, shift = beta + (bias - mean) * scale),
//     so the per-element work of an apply pass is one FMA + Mish (ex2 + rcp) + one FMA;
//   * Mish(GN2(h2)) * m + res_bias is written back INTO the accumulator (tcgen05.st) and res_conv(x) is accumulated onto it by
//     the tensor core, per 128-column half: the MMAs of half 0 run under the apply pass of half 1, those of half 1 under the
//     output pass of half 0;
//   * with one m-block per CTA (mb = 1) conv2 accumulates into the spare 256 TMEM columns and starts on K-chunks 0-1 while the
//     epilogue warps are still producing K-chunks 2-3 of `a`;
//   * LayerNorm statistics are taken per row from the same TMEM tile (thread = row), combined across the four column-slot warps through shared memory (one writer per partial sum, fixed order).
//   * optionally the block that follows starts here too: n is written into the (by then dead) operand planes instead of global
//     memory and the attention's stacked q|k|v projection (256 -> 384, no bias) runs on it in the same launch -- issuer 0, the
//     weight ring continues with W_qkv's 12 tiles; with two m-blocks the 768 output columns take two rounds through TMEM (q|k, then v);
// `mode 1` stops after the first apply (final_block: conv -> GN -> Mish -> mask, decoder.py:431) and writes bf16 to global.
//
// Warp roles (20 warps): 0 activation-tile TMA producer, 1 weight-tile TMA producer (weights are constants: it runs free),
// 2-3 MMA issuers, one per 128-column half of the accumulator (a weight tile only feeds 4-8 MMAs: one issuer's barrier
// bookkeeping runs under the other's MMAs; warp 2 owns TMEM), 4-19 sixteen epilogue warps (TMEM lane quadrant = warp & 3,
// column slot = (warp - 4) >> 2: a warp works on the 32-column blocks 4h + slot of both 128-column halves h).
// Shared memory: weight ring (5-7 slots of [64 k x 128 n] bf16 = 16 KB, 128B swizzle), the A region (block-input tiles for
// conv1, then the four operand planes of `a`, then block-input tiles for res_conv + the output transpose buffers), constants.
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
constexpr int RN_ROLE_WARPS = 4;
constexpr int RN_THREADS = 32 * (RN_ROLE_WARPS + RN_EPI_WARPS);
constexpr int RN_EPI_THREADS = 32 * RN_EPI_WARPS;
constexpr int RN_W_TILE = 128 * 128;             // [64 k x 128 n] bf16
constexpr int RN_MAX_W_SLOTS = 8;
constexpr int RN_X1_SLOTS = 3, RN_X3_SLOTS = 2;
constexpr int RN_STAGE_WARP = 4096;              // 32 x 32 fp32 (XOR-swizzled) or 32 rows x 80 B of bf16
constexpr int RN_STAGE = RN_EPI_WARPS * RN_STAGE_WARP;
constexpr int RN_ACT_PITCH = 80;                 // bytes per staged bf16 row (64 + 16: conflict-free 16-byte access)
constexpr int RN_MAX_CLUSTER = 8;
// constants: scale, shift, extra (temb / res bias), conv bias, LN gamma, LN beta [256 each], LayerNorm partial row sums
// [256 rows][4 column slots][2], GroupNorm partial sums [8 groups][4 lane quadrants][2], the cluster's group sums
// [2 GroupNorms][8 ranks][16].  Every partial sum has ONE writer and is added in a fixed order: results are reproducible.
constexpr int RN_CONST_FLOATS = 6 * RN_C + 256 * 8 + 64 + 2 * RN_MAX_CLUSTER * 16;
constexpr int RN_SMEM_LIMIT = 227 * 1024 - 1024; // dynamic + ~0.5 KB of static barriers must stay below 227 KB

struct RnMaps { CUtensorMap x1, x3, w1, w2, wr, wq; };

struct RnParams {
  int B, T, mb, m_tiles, n_cta;
  int kc_in;                                   // 64-channel K-chunks of the block input (conv1 and res_conv)
  int x1_boxes, x1_box_rows, x1_slot_bytes;    // conv1's haloed input tile: R + 2 rows from frame m0 - 2
  int a_bytes, w_slots;
  int mode;
  const int* lens; int len_shift;
  const float *bias1, *bias2, *bias_r, *g1, *b1, *g2, *b2, *temb, *ln_g, *ln_b;
  long long temb_bs;                           // item stride of temb (0: one time step for the whole batch)
  bf16* a_buf; long long a_ld, a_bs;           // mode 1: the block's output; mode 0: optional copy of conv2's operand (tests)
  float* xr; bf16* n_out;                      // (b, t, 256) dense
  float* xr_cf;                                // instead of xr: the same stream CHANNEL-FIRST (b, 256, t), for ff_tc's tail mode
  bf16* qkv_out;                               // != nullptr: q|k|v = n W_qkv^T (b, t, 384) computed here; n_out is not written
  float eps_gn, eps_ln;
  int trace;
};

__device__ unsigned long long g_rn_trace[32];
#define RN_TR(i) do { if (p.trace && blockIdx.x == 0 && lane == 0) g_rn_trace[(i)] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ uint32_t rd_hi(uint32_t sbo, uint32_t layout) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t rd_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t rd_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

static __device__ __noinline__ void rn_wait_timeout(int line) {
  printf("emojivoice_b200: resnet_tc mbarrier wait at line %d timed out (block %d thread %d)\n", line, blockIdx.x, threadIdx.x);
  __trap();
}
// mbar_wait with the source line in the diagnostic (a pipeline bug must trap, not hang the device)
__device__ __forceinline__ void rn_wait_at(uint64_t* bar, uint32_t parity, int line) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) rn_wait_timeout(line);
  }
}
#define RN_WAIT(bar, parity) rn_wait_at((bar), (parity), __LINE__)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(RN_EPI_THREADS) : "memory"); }

// x * tanh(softplus(x)) + add.  With n = e^x: tanh(log(1 + n)) = 1 - 2 / (n (n + 2) + 2); x is clamped at 20 inside the
// exponential only (there the quotient is 1 to fp32 precision, torch's softplus threshold: decoder.py:38)
__device__ __forceinline__ float mish_add(float x, float add) {
  const float n = ex2f(fminf(x * 1.4426950408889634f, 28.853900817779268f));
  const float d = fmaf(n, n + 2.0f, 2.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(x, fmaf(-2.0f, r, 1.0f), add);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
// 4 bytes into the shared memory of a CTA of the cluster; completes 4 bytes of transaction count on THAT CTA's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(RN_THREADS, 1)
resnet_tc_kernel(const __grid_constant__ RnMaps maps, const __grid_constant__ RnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x1_full[RN_X1_SLOTS], x1_empty[RN_X1_SLOTS], x3_full[RN_X3_SLOTS], x3_empty[RN_X3_SLOTS];
  __shared__ __align__(8) uint64_t w_full[RN_MAX_W_SLOTS], w_empty[RN_MAX_W_SLOTS];
  __shared__ __align__(8) uint64_t acc_full, plane_ready[2], tm_ready[2], res_full[2], xch_bar[2];
  __shared__ __align__(8) uint64_t qkv_ready, qkv_full[2], qkv_drained;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_s = base, a_s = base + (uint32_t)(p.w_slots * RN_W_TILE);
  uint8_t* a_gen = base_gen + p.w_slots * RN_W_TILE;
  uint8_t* stage_gen = a_gen + (p.a_bytes - RN_STAGE);                          // the last 64 KB of the A region
  float* cst = reinterpret_cast<float*>(a_gen + p.a_bytes);
  float *c_scale = cst, *c_shift = cst + RN_C, *c_extra = cst + 2 * RN_C, *c_bias = cst + 3 * RN_C, *c_lng = cst + 4 * RN_C, *c_lnb = cst + 5 * RN_C;
  float* rowsum = cst + 6 * RN_C;                                               // [256 rows][4 slots][2]
  float* gsum = rowsum + 256 * 8;                                               // [8 groups][4 quadrants][2]
  float* xch = gsum + 64;                                                       // [2][8 ranks][16]: partial sums pushed by the cluster's CTAs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.mb * 128, U = R - 2;
  const int b = (int)blockIdx.x / p.m_tiles, mt = (int)blockIdx.x - b * p.m_tiles;
  const int m0 = mt * U;                                                        // first owned frame; tile row r holds frame m0 - 1 + r
  const int vmb = min(p.mb, (p.T - m0 + 1 + 127) >> 7);                         // m-blocks that hold a frame < T
  const bool full = p.mode == 0;
  const uint32_t acc2_col = (p.mb == 1 && full) ? 256u : 0u;                    // conv2 / res_conv / output accumulator columns

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1) : "memory");
    for (int s = 0; s < RN_X1_SLOTS; ++s) { mbar_init(&x1_full[s], 1); mbar_init(&x1_empty[s], 2); }
    for (int s = 0; s < RN_X3_SLOTS; ++s) { mbar_init(&x3_full[s], 1); mbar_init(&x3_empty[s], 1); }
    for (int s = 0; s < RN_MAX_W_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(&acc_full, 2);
    for (int h = 0; h < 2; ++h) {
      mbar_init(&plane_ready[h], RN_EPI_WARPS); mbar_init(&tm_ready[h], RN_EPI_WARPS); mbar_init(&res_full[h], 1); mbar_init(&xch_bar[h], 1);
      mbar_init(&qkv_full[h], 1);
    }
    mbar_init(&qkv_ready, RN_EPI_WARPS); mbar_init(&qkv_drained, RN_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                             // every CTA's mbarriers exist before a peer pushes its sums at them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_trigger();
  pdl_wait();
  RN_TR(0);

  if (warp == 0) {
    // ---------------- block-input tiles: conv1 reads R + 2 rows from frame m0 - 2; res_conv reads R rows from frame m0 - 1
    if (lane == 0) {
      int sx = 0;
      uint32_t px = 1;
      const uint32_t x1_bytes = (uint32_t)(p.x1_boxes * p.x1_box_rows * 128), box_bytes = (uint32_t)(p.x1_box_rows * 128);
      for (int kc = 0; kc < p.kc_in; ++kc) {
        RN_WAIT(&x1_empty[sx], px);
        mbar_expect_tx(&x1_full[sx], x1_bytes);
        const uint32_t dst = a_s + (uint32_t)(sx * p.x1_slot_bytes);
        tma_load_3d(dst, &maps.x1, &x1_full[sx], kc * 64, m0 - 2, b);
        if (p.x1_boxes > 1) tma_load_3d(dst + box_bytes, &maps.x1, &x1_full[sx], kc * 64, m0 - 2 + p.x1_box_rows, b);
        if (++sx == RN_X1_SLOTS) { sx = 0; px ^= 1u; }
      }
      if (full) {
        RN_WAIT(&acc_full, 0);
        RN_WAIT(&acc_full, 1);                  // conv2's MMAs have completed: the operand planes are dead
        sx = 0; px = 1;
        const uint32_t x3_bytes = (uint32_t)(R * 128);
        for (int h = 0; h < 2; ++h)
          for (int kc = 0; kc < p.kc_in; ++kc) {
            RN_WAIT(&x3_empty[sx], px);
            mbar_expect_tx(&x3_full[sx], x3_bytes);
            const uint32_t dst = a_s + (uint32_t)sx * x3_bytes;
            for (int j = 0; j < p.mb; ++j) tma_load_3d(dst + (uint32_t)(j * 128 * 128), &maps.x3, &x3_full[sx], kc * 64, m0 - 1 + j * 128, b);
            if (++sx == RN_X3_SLOTS) { sx = 0; px ^= 1u; }
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- weight tiles in the issuer's order (constants: no dependency on anything)
    if (lane == 0) {
      int sw = 0;
      uint32_t pw = 1;
      auto load_w = [&](const CUtensorMap* map, int kc, int h, int tap) {
        RN_WAIT(&w_empty[sw], pw);
        mbar_expect_tx(&w_full[sw], (uint32_t)RN_W_TILE);
        tma_load_3d(w_s + (uint32_t)(sw * RN_W_TILE), map, &w_full[sw], kc * 64, h * 128, tap);
        if (++sw == p.w_slots) { sw = 0; pw ^= 1u; }
      };
      for (int kc = 0; kc < p.kc_in; ++kc)
        for (int j = 0; j < 3; ++j)
          for (int h = 0; h < 2; ++h) load_w(&maps.w1, kc, h, j);
      if (full) {
        for (int kc = 0; kc < RN_C / 64; ++kc)
          for (int j = 0; j < 3; ++j)
            for (int h = 0; h < 2; ++h) load_w(&maps.w2, kc, h, j);
        for (int h = 0; h < 2; ++h)
          for (int kc = 0; kc < p.kc_in; ++kc) load_w(&maps.wr, kc, h, 0);
        if (p.qkv_out)
          for (int n3 = 0; n3 < 3; ++n3)
            for (int kc = 0; kc < RN_C / 64; ++kc) load_w(&maps.wq, kc, n3, 0);
      }
    }
    __syncwarp();
  } else if (warp < RN_ROLE_WARPS) {
    // ---------------- two MMA issuers, h = 0 / 1: the 128-column half of every accumulator.  Warp-uniform control flow, one
    // elected lane issues.  M = 128, N = 128, K = 16.  Weight tiles are numbered in the producer's order; tile i sits in ring
    // slot i % w_slots.  conv1 / conv2: tile 2 (3 kc + j) + h belongs to issuer h; res_conv: issuer 0 takes every tile.
    const int h = warp - 2;
    constexpr uint32_t idesc = make_idesc(128, 128);
    const uint32_t hi = rd_hi(1024u, 2u);
    const uint32_t a_lo0 = rd_lo(a_s), w_lo0 = rd_lo(w_s);
    constexpr uint32_t w_step16 = (uint32_t)RN_W_TILE >> 4, row16 = 128u >> 4;
    int sw = h;
    uint32_t pw = 0;
    auto ring_advance = [&](int k) { sw += k; while (sw >= p.w_slots) { sw -= p.w_slots; pw ^= 1u; } };
    // the MMAs of one weight tile: m-blocks j < vmb, A rows from `a_lo` + j * 128 rows, D columns d0 + j * 256
    // The two issuers share the weight ring.  The slot count is EVEN, so every slot only ever holds tiles of one issuer: an
    // issuer asks for fill k of a slot after it consumed fill k - 1 itself.  (With an odd count fill k - 1 would be the other
    // issuer's tile, possibly still in flight -- weight tiles that miss L2 land out of order -- and a parity wait cannot tell
    // fill k from fill k - 2: the ring protocol would break.)
    auto tile_mmas = [&](uint32_t a_lo, uint32_t d0, uint32_t acc) {
      RN_WAIT(&w_full[sw], pw);
      tcgen05_fence_after();
      const uint32_t w_lo = w_lo0 + (uint32_t)sw * w_step16;
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          if (m < vmb) {
            const uint32_t d = d0 + (uint32_t)(m * RN_C), am = a_lo + (uint32_t)m * (128u * row16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(d, rd_join(hi, am + 2u * ks), rd_join(hi, w_lo + 2u * ks), idesc, (acc | (uint32_t)ks) ? 1u : 0u);
          }
        }
        umma_commit(&w_empty[sw]);
      }
      __syncwarp();
    };
    {  // conv1: tap j of tile row r reads input-tile row r + j (input-tile row 0 = frame m0 - 2)
      int sx = 0;
      uint32_t px = 0;
      for (int kc = 0; kc < p.kc_in; ++kc) {
        RN_WAIT(&x1_full[sx], px);
        tcgen05_fence_after();
        if (h == 0 && kc == 0) RN_TR(15);
        const uint32_t a_lo = a_lo0 + (uint32_t)(sx * p.x1_slot_bytes >> 4);
        for (int j = 0; j < 3; ++j) {
          tile_mmas(a_lo + (uint32_t)j * row16, tmem_base + (uint32_t)(h * 128), (kc | j) ? 1u : 0u);
          ring_advance(2);
        }
        if (elect_one()) umma_commit(&x1_empty[sx]);
        __syncwarp();
        if (++sx == RN_X1_SLOTS) { sx = 0; px ^= 1u; }
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
    if (h == 0) RN_TR(1);
    if (full) {
      // conv2: operand planes written by the epilogue warps; tap j of tile row r reads plane row r + j - 1 (rows -1 and R only
      // feed the two halo rows of the output, which nobody reads).  K-chunks 0-1 are the channels of half 0, 2-3 of half 1.
      // With two m-blocks conv2 overwrites conv1's accumulator: every apply-1 read must have happened (both halves ready).
      const uint32_t plane16 = (uint32_t)(R * 128) >> 4;
      RN_WAIT(&plane_ready[0], 0);
      if (p.mb == 2) RN_WAIT(&plane_ready[1], 0);
      tcgen05_fence_after();
      for (int kc = 0; kc < RN_C / 64; ++kc) {
        if (kc == 2 && p.mb == 1) { RN_WAIT(&plane_ready[1], 0); tcgen05_fence_after(); }
        const uint32_t a_lo = a_lo0 + (uint32_t)kc * plane16;
        for (int j = 0; j < 3; ++j) {
          tile_mmas(a_lo + (uint32_t)j * row16 - row16, tmem_base + acc2_col + (uint32_t)(h * 128), (kc | j) ? 1u : 0u);
          ring_advance(2);
        }
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
      if (h == 0) RN_TR(7);
      // res_conv, accumulated on top of Mish(GN2(h2)) * m + bias, half by half: input-tile row r = frame m0 - 1 + r.  One
      // issuer does both halves (the other one would have to join the rings mid-sequence: mbarrier parities alias); the
      // input tiles stream twice, in the order h = 0: kc = 0.., h = 1: kc = 0..
      if (h == 0) {
        int sx = 0;
        uint32_t px = 0;
        const uint32_t x3_16 = (uint32_t)(R * 128) >> 4;
        for (int hh = 0; hh < 2; ++hh) {
          RN_WAIT(&tm_ready[hh], 0);
          tcgen05_fence_after();
          for (int kc = 0; kc < p.kc_in; ++kc) {
            RN_WAIT(&x3_full[sx], px);
            tcgen05_fence_after();
            tile_mmas(a_lo0 + (uint32_t)sx * x3_16, tmem_base + acc2_col + (uint32_t)(hh * 128), 1u);
            ring_advance(1);
            if (elect_one()) umma_commit(&x3_empty[sx]);
            __syncwarp();
            if (++sx == RN_X3_SLOTS) { sx = 0; px ^= 1u; }
          }
          if (elect_one()) umma_commit(&res_full[hh]);
          __syncwarp();
        }
        if (p.qkv_out) {
          // q|k|v = n W_qkv^T: n sits in the operand planes (written by the LayerNorm pass), one 128-column output tile per n3.
          // One m-block: q, k, v -> columns 0 / 128 / 256.  Two m-blocks (m at + 256): q, k first, v after the epilogue has drained them.
          const uint32_t plane16 = (uint32_t)(R * 128) >> 4;
          RN_WAIT(&qkv_ready, 0);
          tcgen05_fence_after();
          for (int n3 = 0; n3 < 3; ++n3) {
            if (n3 == 2 && p.mb == 2) {
              if (elect_one()) umma_commit(&qkv_full[0]);
              __syncwarp();
              RN_WAIT(&qkv_drained, 0);
              tcgen05_fence_after();
            }
            const uint32_t d0 = tmem_base + (uint32_t)((p.mb == 2 ? (n3 & 1) : n3) * 128);
            for (int kc = 0; kc < RN_C / 64; ++kc) {
              tile_mmas(a_lo0 + (uint32_t)kc * plane16, d0, kc ? 1u : 0u);
              ring_advance(1);
            }
          }
          if (elect_one()) umma_commit(&qkv_full[p.mb == 2 ? 1 : 0]);
          __syncwarp();
        }
      }
      if (h == 0) RN_TR(10);
    }
  } else {
    // ---------------- sixteen epilogue warps
    const int ew = warp - RN_ROLE_WARPS, q = warp & 3, slot = ew >> 2, te = threadIdx.x - 32 * RN_ROLE_WARPS;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* wstage = stage_gen + ew * RN_STAGE_WARP;
    const int len_b = p.lens ? __ldg(p.lens + b) : 0x7fffffff;
    const double inv_n = 1.0 / (32.0 * (double)p.T);
    const int rl = q * 32 + lane;                                      // row of this thread inside an m-block

    // constants that do not depend on a barrier
    if (te < RN_C) {
      c_bias[te] = __ldg(p.bias1 + te);
      if (full) { c_lng[te] = __ldg(p.ln_g + te); c_lnb[te] = __ldg(p.ln_b + te); }
    }
    epi_bar();

    // GroupNorm statistics of (accumulator + bias) over the OWNED rows of the tile
    auto stats_pass = [&](uint32_t col0) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int cb = 4 * h + slot;
        float s = 0.0f, qq = 0.0f;
#pragma unroll 1
        for (int m = 0; m < vmb; ++m) {
          uint32_t raw[32];
          tmem_ld32(lane_addr + col0 + (uint32_t)(m * RN_C + cb * 32), raw);
          const int r = m * 128 + rl;
          const bool own = r >= 1 && r <= R - 2 && m0 - 1 + r < p.T;
          float s1 = 0.0f, q1 = 0.0f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(c_bias + cb * 32 + j);
            const float v0 = __uint_as_float(raw[j]) + bv.x, v1 = __uint_as_float(raw[j + 1]) + bv.y;
            const float v2 = __uint_as_float(raw[j + 2]) + bv.z, v3 = __uint_as_float(raw[j + 3]) + bv.w;
            s1 += (v0 + v1) + (v2 + v3);
            q1 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, q1))));
          }
          if (own) { s += s1; qq += q1; }
        }
        s = warp_sum(s); qq = warp_sum(qq);
        if (lane == 0) *reinterpret_cast<float2*>(gsum + (cb * 4 + q) * 2) = make_float2(s, qq);
      }
    };
    // exchange k: push this CTA's 16 partial sums into slot `mt` of every CTA of the cluster (this one included); the
    // receiving CTA's mbarrier counts the bytes, so nobody waits for anything but the data itself
    auto exchange = [&](int k) {
      epi_bar();                                                       // this CTA's shared-memory atomics have landed
      if (ew == 0) {
        if (lane == 0) mbar_expect_tx(&xch_bar[k], (uint32_t)(p.m_tiles * 64));
        if (lane < 16) {
          const float* gp = gsum + (lane >> 1) * 8 + (lane & 1);        // group lane / 2, sum or sum of squares: quadrants in order
          const float v = (gp[0] + gp[2]) + (gp[4] + gp[6]);
          const uint32_t slot_addr = smem_u32(xch + (k * RN_MAX_CLUSTER + mt) * 16 + lane), bar_addr = smem_u32(&xch_bar[k]);
          for (int c = 0; c < p.m_tiles; ++c) st_async_f32(map_to_cta(slot_addr, (uint32_t)c), v, map_to_cta(bar_addr, (uint32_t)c));
        }
      }
      RN_WAIT(&xch_bar[k], 0);
    };
    // per-channel scale / shift of this item's GroupNorm (partial sums added in rank order, mean / variance in fp64), bias folded in
    auto fold_consts = [&](int k, const float* gamma, const float* beta, const float* extra, const float* next_bias) {
      if (te < RN_C) {
        const float* part = xch + (k * RN_MAX_CLUSTER) * 16 + (te >> 5) * 2;
        double s = 0.0, qq = 0.0;
        for (int c = 0; c < p.m_tiles; ++c) { s += (double)part[c * 16]; qq += (double)part[c * 16 + 1]; }
        const double mu = s * inv_n;
        double var = qq * inv_n - mu * mu;
        if (var < 0.0) var = 0.0;
        const float mean = (float)mu, ve = (float)(var + (double)p.eps_gn);
        float rstd = rsqrtf(ve);
        rstd = rstd * fmaf(-0.5f * ve, rstd * rstd, 1.5f);             // one Newton step: fp32-exact to the last bit or two
        const float sc = rstd * __ldg(gamma + te);
        const float bias = c_bias[te];
        c_scale[te] = sc;
        c_shift[te] = fmaf(bias - mean, sc, __ldg(beta + te));
        c_extra[te] = extra ? __ldg(extra + te) : 0.0f;
        if (next_bias) c_bias[te] = __ldg(next_bias + te);
      }
      epi_bar();
    };
    // 32 bf16 values per lane (thread = row) -> coalesced 16-byte stores of the owned rows of the 32 x 32 block
    auto store_bf16_block = [&](const uint32_t (&pk)[16], bf16* dst, long long ld, int m) {
      uint8_t* brow = wstage + lane * RN_ACT_PITCH;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(brow + i * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      __syncwarp();
      const int rsub = lane >> 2, ch = lane & 3;
      const uint8_t* src = wstage + ch * 16;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = rsub + it * 8, r = m * 128 + q * 32 + rr, t = m0 - 1 + r;
        if (r >= 1 && r <= R - 2 && t < p.T) *reinterpret_cast<uint4*>(dst + (long long)t * ld + ch * 8) = *reinterpret_cast<const uint4*>(src + rr * RN_ACT_PITCH);
      }
      __syncwarp();
    };

    // ======== conv1 done: statistics -> exchange within the cluster -> apply 1
    RN_WAIT(&acc_full, 0);
    tcgen05_fence_after();
    if (ew == 0) RN_TR(2);
    stats_pass(0u);
    exchange(0);
    if (ew == 0) RN_TR(3);
    fold_consts(0, p.g1, p.b1, full ? p.temb + b * p.temb_bs : nullptr, full ? p.bias2 : nullptr);
    if (ew == 0) RN_TR(13);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int cb = 4 * h + slot;
      if (ew == 0 && h == 1) RN_TR(14);
#pragma unroll 1
      for (int m = 0; m < p.mb; ++m) {
        const int r = m * 128 + rl, t = m0 - 1 + r;
        const bool valid = t >= 0 && t < p.T && (t << p.len_shift) < len_b;   // (Mish * m + temb) * m; rows outside [0, T) are conv2's zero padding
        uint32_t pk[16];
        // the pass is MUFU-bound (ex2 + rcp per element: ~2.4 k clk per 32 x 32 block with 16 warps): a warp whose 32 frames are all
        // padding -- a third of the rows of a config-2 batch -- writes its zeros without the arithmetic (warp-uniform branch)
        if (__any_sync(0xffffffffu, valid)) {
          uint32_t raw[32];
          tmem_ld32(lane_addr + (uint32_t)(m * RN_C + cb * 32), raw);
          if (ew == 0 && h == 0 && m == 0) RN_TR(16);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int c = cb * 32 + j;
            const float4 sc = *reinterpret_cast<const float4*>(c_scale + c), sh = *reinterpret_cast<const float4*>(c_shift + c);
            const float4 ex = *reinterpret_cast<const float4*>(c_extra + c);
            const float v0 = mish_add(fmaf(__uint_as_float(raw[j]), sc.x, sh.x), ex.x);
            const float v1 = mish_add(fmaf(__uint_as_float(raw[j + 1]), sc.y, sh.y), ex.y);
            const float v2 = mish_add(fmaf(__uint_as_float(raw[j + 2]), sc.z, sh.z), ex.z);
            const float v3 = mish_add(fmaf(__uint_as_float(raw[j + 3]), sc.w, sh.w), ex.w);
            pk[j >> 1] = valid ? pack_bf16(v0, v1) : 0u;                 // a select: the row may hold anything
            pk[(j >> 1) + 1] = valid ? pack_bf16(v2, v3) : 0u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        if (ew == 0 && h == 0 && m == 0) { asm volatile("" ::"r"(pk[0]), "r"(pk[15])); RN_TR(17); }
        if (full) {
          // K-major operand plane (cb >> 1), row r, 16-byte chunks 4 (cb & 1) + i, 128B swizzle: chunk ^= r & 7
          uint8_t* prow = a_gen + (cb >> 1) * (R * 128) + r * 128;
          const int c0 = 4 * (cb & 1), x7 = r & 7;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(prow + (((c0 + i) ^ x7) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          if (p.a_buf && r >= 1 && r <= R - 2 && t < p.T) {            // tests only: a copy of the operand
            uint4* dst = reinterpret_cast<uint4*>(p.a_buf + b * p.a_bs + (long long)t * p.a_ld + cb * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        } else {
          store_bf16_block(pk, p.a_buf + b * p.a_bs + cb * 32, p.a_ld, m);
        }
      }
      if (ew == 0 && h == 0) RN_TR(18);
      if (full) {
        tcgen05_fence_before();
        fence_proxy_async();                                           // generic-proxy smem stores -> the MMAs' async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&plane_ready[h]);
      }
    }
    if (ew == 0) RN_TR(4);
    if (full) {
      // ======== conv2 done: statistics -> exchange -> Mish(GN2(h2)) * m + res bias back into the accumulator
      RN_WAIT(&acc_full, 1);
      tcgen05_fence_after();
      if (ew == 0) RN_TR(8);
      stats_pass(acc2_col);
      exchange(1);
      if (ew == 0) RN_TR(5);
      fold_consts(1, p.g2, p.b2, p.bias_r, nullptr);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int cb = 4 * h + slot;
#pragma unroll 1
        for (int m = 0; m < vmb; ++m) {
          uint32_t raw[32];
          const uint32_t taddr = lane_addr + acc2_col + (uint32_t)(m * RN_C + cb * 32);
          const int t = m0 - 1 + m * 128 + rl;
          const bool valid = t >= 0 && t < p.T && (t << p.len_shift) < len_b;
          if (__any_sync(0xffffffffu, valid)) {
            tmem_ld32(taddr, raw);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int c = cb * 32 + j;
              const float4 sc = *reinterpret_cast<const float4*>(c_scale + c), sh = *reinterpret_cast<const float4*>(c_shift + c);
              const float4 ex = *reinterpret_cast<const float4*>(c_extra + c);
              const float v0 = mish_add(fmaf(__uint_as_float(raw[j]), sc.x, sh.x), ex.x);
              const float v1 = mish_add(fmaf(__uint_as_float(raw[j + 1]), sc.y, sh.y), ex.y);
              const float v2 = mish_add(fmaf(__uint_as_float(raw[j + 2]), sc.z, sh.z), ex.z);
              const float v3 = mish_add(fmaf(__uint_as_float(raw[j + 3]), sc.w, sh.w), ex.w);
              raw[j] = __float_as_uint(valid ? v0 : ex.x); raw[j + 1] = __float_as_uint(valid ? v1 : ex.y);
              raw[j + 2] = __float_as_uint(valid ? v2 : ex.z); raw[j + 3] = __float_as_uint(valid ? v3 : ex.w);
            }
          } else {           // 32 padded frames: the masked block output is the res_conv bias alone
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ex = *reinterpret_cast<const float4*>(c_extra + cb * 32 + j);
              raw[j] = __float_as_uint(ex.x); raw[j + 1] = __float_as_uint(ex.y); raw[j + 2] = __float_as_uint(ex.z); raw[j + 3] = __float_as_uint(ex.w);
            }
          }
          tmem_st32(taddr, raw);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tm_ready[h]);                      // res_conv may accumulate on top of this half
      }
      if (ew == 0) RN_TR(9);

      // ======== xr = accumulator -> fp32 stream (coalesced through the transpose buffer) + LayerNorm row sums
      const int sub = lane >> 3, cl = lane & 7;
      float ls[2] = {0.0f, 0.0f}, lq[2] = {0.0f, 0.0f};
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int cb = 4 * h + slot;
        RN_WAIT(&res_full[h], 0);
        tcgen05_fence_after();
        if (ew == 0 && h == 0) RN_TR(11);
#pragma unroll 1
        for (int m = 0; m < vmb; ++m) {
          uint32_t raw[32];
          tmem_ld32(lane_addr + acc2_col + (uint32_t)(m * RN_C + cb * 32), raw);
          float s = 0.0f, qq = 0.0f;
          if (p.xr_cf) {
            // channel-first stream: the 32 rows of a warp are consecutive addresses of one channel -- coalesced straight from
            // the registers, no transpose (the consumer, ff_tc's tail mode, reads it thread = row as well)
            const int r = m * 128 + rl, t = m0 - 1 + r;
            const bool own = r >= 1 && r <= R - 2 && t < p.T;
            float* dst = p.xr_cf + ((long long)b * RN_C + cb * 32) * p.T + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(raw[j]);
              s += v;
              qq = fmaf(v, v, qq);
              if (own) dst[(long long)j * p.T] = v;
            }
            if (m == 0) { ls[0] += s; lq[0] += qq; } else { ls[1] += s; lq[1] += qq; }
            continue;
          }
          uint8_t* srow = wstage + lane * 128;
          const int x7 = lane & 7;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float v0 = __uint_as_float(raw[j]), v1 = __uint_as_float(raw[j + 1]), v2 = __uint_as_float(raw[j + 2]), v3 = __uint_as_float(raw[j + 3]);
            s += (v0 + v1) + (v2 + v3);
            qq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, qq))));
            *reinterpret_cast<float4*>(srow + ((((j >> 2) ^ x7)) << 4)) = make_float4(v0, v1, v2, v3);
          }
          if (m == 0) { ls[0] += s; lq[0] += qq; } else { ls[1] += s; lq[1] += qq; }
          __syncwarp();
          float* dst = p.xr + ((long long)b * p.T) * RN_C + cb * 32 + cl * 4;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int rr = u * 4 + sub, r = m * 128 + q * 32 + rr, t = m0 - 1 + r;
            if (r >= 1 && r <= R - 2 && t < p.T)
              *reinterpret_cast<float4*>(dst + (long long)t * RN_C) = *reinterpret_cast<const float4*>(wstage + rr * 128 + ((cl ^ (rr & 7)) << 4));
          }
          __syncwarp();
        }
      }
      *reinterpret_cast<float2*>(rowsum + (rl * 4 + slot) * 2) = make_float2(ls[0], lq[0]);
      if (vmb > 1) *reinterpret_cast<float2*>(rowsum + ((128 + rl) * 4 + slot) * 2) = make_float2(ls[1], lq[1]);
      epi_bar();
      if (ew == 0) RN_TR(6);
      // ======== n = LayerNorm(xr) -> bf16 operand of the QKV projection
#pragma unroll 1
      for (int m = 0; m < vmb; ++m) {
        const float4 ra = *reinterpret_cast<const float4*>(rowsum + (m * 128 + rl) * 8), rb = *reinterpret_cast<const float4*>(rowsum + (m * 128 + rl) * 8 + 4);
        const float mu = ((ra.x + ra.z) + (rb.x + rb.z)) * (1.0f / RN_C);
        const float var = fmaxf(((ra.y + ra.w) + (rb.y + rb.w)) * (1.0f / RN_C) - mu * mu, 0.0f);
        const float rs = rsqrtf(var + p.eps_ln);
        const float nmu = -mu * rs;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int cb = 4 * h + slot;
          uint32_t raw[32];
          tmem_ld32(lane_addr + acc2_col + (uint32_t)(m * RN_C + cb * 32), raw);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int c = cb * 32 + j;
            const float4 ga = *reinterpret_cast<const float4*>(c_lng + c), be = *reinterpret_cast<const float4*>(c_lnb + c);
            const float v0 = fmaf(fmaf(__uint_as_float(raw[j]), rs, nmu), ga.x, be.x);
            const float v1 = fmaf(fmaf(__uint_as_float(raw[j + 1]), rs, nmu), ga.y, be.y);
            const float v2 = fmaf(fmaf(__uint_as_float(raw[j + 2]), rs, nmu), ga.z, be.z);
            const float v3 = fmaf(fmaf(__uint_as_float(raw[j + 3]), rs, nmu), ga.w, be.w);
            pk[j >> 1] = pack_bf16(v0, v1);
            pk[(j >> 1) + 1] = pack_bf16(v2, v3);
          }
          if (p.qkv_out) {
            // operand plane (cb >> 1), row r, 128B swizzle -- the layout conv2's operand used; those planes are dead by now
            const int r = m * 128 + rl;
            uint8_t* prow = a_gen + (cb >> 1) * (R * 128) + r * 128;
            const int c0 = 4 * (cb & 1), x7 = r & 7;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(prow + (((c0 + i) ^ x7) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          } else {
            store_bf16_block(pk, p.n_out + ((long long)b * p.T) * RN_C + cb * 32, RN_C, m);
          }
        }
      }
      if (p.qkv_out) {
        tcgen05_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qkv_ready);
        // drain: thread = row writes its 32 bf16 (64 contiguous bytes = two whole sectors) per 32-column block straight from registers
        // (the transpose buffers overlap the planes the second round's MMAs still read)
        const int n_rounds = p.mb == 2 ? 2 : 1;
        for (int round = 0; round < n_rounds; ++round) {
          RN_WAIT(&qkv_full[round], 0);
          tcgen05_fence_after();
          const int nb = (p.mb == 2 ? (round == 0 ? 8 : 4) : 12);           // 32-column blocks per m-block in this round
          for (int m = 0; m < vmb; ++m) {
            const int r = m * 128 + rl, t = m0 - 1 + r;
            const bool own = r >= 1 && r <= R - 2 && t < p.T;
            bf16* drow = p.qkv_out + ((long long)b * p.T + t) * 384 + (round == 1 ? 256 : 0);
#pragma unroll 1
            for (int blk = slot; blk < nb; blk += 4) {
              uint32_t raw[32];
              tmem_ld32(lane_addr + (uint32_t)(m * RN_C + blk * 32), raw);
              if (own) {
                uint4* dst = reinterpret_cast<uint4*>(drow + blk * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  dst[i] = make_uint4(pack_bf16(__uint_as_float(raw[8 * i]), __uint_as_float(raw[8 * i + 1])),
                                      pack_bf16(__uint_as_float(raw[8 * i + 2]), __uint_as_float(raw[8 * i + 3])),
                                      pack_bf16(__uint_as_float(raw[8 * i + 4]), __uint_as_float(raw[8 * i + 5])),
                                      pack_bf16(__uint_as_float(raw[8 * i + 6]), __uint_as_float(raw[8 * i + 7])));
              }
            }
          }
          if (round + 1 < n_rounds) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&qkv_drained);
          }
        }
      }
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

// A region: conv1's input ring, the four operand planes, or res_conv's input ring + the 64 KB of transpose buffers
int rn_a_bytes(int mb, int x1_slot_bytes) {
  const int R = 128 * mb;
  return (int)align_up((size_t)std::max({RN_X1_SLOTS * x1_slot_bytes, 4 * R * 128, RN_X3_SLOTS * R * 128 + RN_STAGE}), 1024);
}

}  // namespace

cudaError_t resnet_tc_read_trace(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_rn_trace, sizeof(unsigned long long) * std::min(n, 32));
}

namespace {
struct RnPlan { int mb, m_tiles, x1_boxes, x1_box_rows, x1_slot_bytes, a_bytes, w_slots, smem_bytes; };
RnPlan rn_plan_for(int mb, int T) {
  RnPlan q{};
  q.mb = mb;
  const int R = 128 * mb, x1_rows = R + 2;
  q.m_tiles = ceil_div(T, R - 2);
  q.x1_boxes = x1_rows > 256 ? 2 : 1;
  q.x1_box_rows = (int)align_up((size_t)ceil_div(x1_rows, q.x1_boxes), 8);
  q.x1_slot_bytes = (int)align_up((size_t)q.x1_boxes * q.x1_box_rows * 128, 1024);
  q.a_bytes = rn_a_bytes(mb, q.x1_slot_bytes);
  const int fixed = 1024 + q.a_bytes + RN_CONST_FLOATS * 4;
  q.w_slots = std::min(RN_MAX_W_SLOTS, (RN_SMEM_LIMIT - fixed) / RN_W_TILE) & ~1;   // even: see the issuers' ring protocol
  q.smem_bytes = fixed + q.w_slots * RN_W_TILE;
  return q;
}
cudaError_t rn_set_attributes() {
  static DeviceOnce once;
  return once.run([&]() { return cudaFuncSetAttribute(resnet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RN_SMEM_LIMIT); });
}
// CTAs of `cluster`-sized clusters (one m-block per CTA) the device holds at once; clusters do not straddle GPCs, so this
// can be less than the SM count
int rn_resident_ctas(int cluster) {
  static std::atomic<int> cache[RN_MAX_CLUSTER + 1];
  int v = cache[cluster].load(std::memory_order_relaxed);
  if (v > 0) return v;
  v = tc_sm_count() / cluster * cluster * 3 / 4;                     // conservative guess if the query is unavailable
  if (rn_set_attributes() == cudaSuccess) {
    const RnPlan q = rn_plan_for(1, 126 * cluster);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster * 64); cfg.blockDim = dim3(RN_THREADS); cfg.dynamicSmemBytes = q.smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, resnet_tc_kernel, &cfg) == cudaSuccess && n > 0) v = n * cluster;
    else (void)cudaGetLastError();
  }
  cache[cluster].store(v, std::memory_order_relaxed);
  return v;
}
}  // namespace

// m-blocks per CTA for B items of T frames (a CTA owns 128 * mb - 2 frames; the tiles of an item form a cluster of <= 8 CTAs):
// one m-block when the whole batch then fits one wave, else two; 0 when an item needs more than 8 tiles
int resnet_tc_plan(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  const int t1 = ceil_div(T, 126), t2 = ceil_div(T, 254);
  if (t1 <= RN_MAX_CLUSTER && (long long)B * t1 <= rn_resident_ctas(t1)) return 1;
  if (t2 <= RN_MAX_CLUSTER) return 2;
  return 0;
}

bool resnet_tc_supported(const ConvWeights& conv1, const ConvWeights* conv2, const ConvWeights* res, int B, int T) {
  if (g_rn_mode < 0) {
    const char* v = getenv("EV_RN_FUSE");
    g_rn_mode = v ? atoi(v) : 1;
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

bool resnet_tc_qkv_supported(const ConvWeights& w) {
  static const bool on = []() { const char* v = getenv("EV_QKV_FUSE"); return !(v && atoi(v) == 0); }();
  return on && w.w_bf16 && !w.bias && w.taps == 1 && w.N == 384 && w.N_pad_tc == 384 && w.C_in == RN_C && w.K_pad == RN_C && w.conv_stride == 1 &&
         !w.transposed;
}

cudaError_t resnet_tc_launch(const ResnetTcArgs& a, cudaStream_t s, std::string* err) {
  const ConvWeights& w1 = *a.conv1;
  const bool full = a.conv2 != nullptr;
  if (!resnet_tc_supported(w1, a.conv2, a.res, a.B, a.T)) {
    if (err) *err = "resnet_tc: unsupported layer shapes, or more than 8 tiles per item";
    return cudaErrorInvalidValue;
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if ((a.x_ld & 7) || (a.x_bs & 7) || !al16(a.x) || (a.a_buf && ((a.a_ld & 7) || (a.a_bs & 7) || !al16(a.a_buf))) || (!full && !a.a_buf) ||
      (full && (!al16(a.xr) || !al16(a.n_out) || !al16(a.qkv_out) || (!a.xr && !a.xr_cf) || (!a.n_out && !a.qkv_out) || !a.temb || !a.ln_g || !a.ln_b)) ||
      (a.qkv_out && (!full || !a.qkv || !resnet_tc_qkv_supported(*a.qkv)))) {
    if (err) *err = "resnet_tc: tensors must be 16-byte aligned";
    return cudaErrorInvalidValue;
  }
  const RnPlan plan = rn_plan_for(resnet_tc_plan(a.B, a.T), a.T);
  if (plan.w_slots < 4) {
    if (err) *err = "resnet_tc: shared-memory plan does not fit";
    return cudaErrorInvalidValue;
  }
  RnParams p{};
  p.B = a.B; p.T = a.T;
  p.mb = plan.mb; p.m_tiles = plan.m_tiles;
  p.n_cta = a.B * p.m_tiles;
  p.kc_in = ceil_div(w1.C_in, 64);
  p.x1_boxes = plan.x1_boxes; p.x1_box_rows = plan.x1_box_rows; p.x1_slot_bytes = plan.x1_slot_bytes;
  p.a_bytes = plan.a_bytes; p.w_slots = plan.w_slots;
  const int smem_bytes = plan.smem_bytes;
  p.mode = full ? 0 : 1;
  p.lens = a.lens; p.len_shift = a.len_shift;
  p.bias1 = w1.bias; p.g1 = a.gn_g1; p.b1 = a.gn_b1; p.temb = a.temb; p.temb_bs = a.temb_bs;
  p.a_buf = a.a_buf; p.a_ld = a.a_ld; p.a_bs = a.a_bs;
  p.eps_gn = 1e-5f; p.eps_ln = 1e-5f;
  { static const int tr = []() { const char* v = getenv("EV_RN_TRACE"); return v ? atoi(v) : 0; }(); p.trace = tr; }
  RnMaps maps;
  bool ok = tc_encode_bf16_map(&maps.x1, a.x, (uint64_t)w1.C_in, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.x_ld * 2, (uint64_t)a.x_bs * 2,
                               64u, (uint32_t)p.x1_box_rows, 128, err);
  ok = ok && tc_encode_bf16_map(&maps.w1, w1.w_bf16, (uint64_t)w1.K_pad, (uint64_t)w1.N_pad_tc, 3, (uint64_t)w1.K_pad * 2,
                                (uint64_t)w1.K_pad * w1.N_pad_tc * 2, 64u, 128u, 128, err);
  if (full) {
    const ConvWeights& w2 = *a.conv2;
    const ConvWeights& wr = *a.res;
    p.bias2 = w2.bias; p.bias_r = wr.bias; p.g2 = a.gn_g2; p.b2 = a.gn_b2; p.ln_g = a.ln_g; p.ln_b = a.ln_b;
    p.xr = a.xr; p.n_out = a.n_out; p.xr_cf = a.xr_cf; p.qkv_out = a.qkv_out;
    ok = ok && tc_encode_bf16_map(&maps.x3, a.x, (uint64_t)w1.C_in, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.x_ld * 2, (uint64_t)a.x_bs * 2,
                                  64u, 128u, 128, err);
    ok = ok && tc_encode_bf16_map(&maps.w2, w2.w_bf16, (uint64_t)w2.K_pad, (uint64_t)w2.N_pad_tc, 3, (uint64_t)w2.K_pad * 2,
                                  (uint64_t)w2.K_pad * w2.N_pad_tc * 2, 64u, 128u, 128, err);
    ok = ok && tc_encode_bf16_map(&maps.wr, wr.w_bf16, (uint64_t)wr.K_pad, (uint64_t)wr.N_pad_tc, 1, (uint64_t)wr.K_pad * 2,
                                  (uint64_t)wr.K_pad * wr.N_pad_tc * 2, 64u, 128u, 128, err);
    if (a.qkv_out) {
      const ConvWeights& wq = *a.qkv;
      ok = ok && tc_encode_bf16_map(&maps.wq, wq.w_bf16, (uint64_t)wq.K_pad, (uint64_t)wq.N_pad_tc, 1, (uint64_t)wq.K_pad * 2,
                                    (uint64_t)wq.K_pad * wq.N_pad_tc * 2, 64u, 128u, 128, err);
    } else {
      maps.wq = maps.w1;
    }
  } else {
    maps.x3 = maps.x1; maps.w2 = maps.w1; maps.wr = maps.w1; maps.wq = maps.w1;
  }
  if (!ok) return cudaErrorInvalidValue;
  cudaError_t ce = rn_set_attributes();
  if (ce != cudaSuccess) return ce;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_cta); cfg.blockDim = dim3(RN_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;                      // one cluster = the tiles of one item
  at[0].val.clusterDim.x = p.m_tiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, resnet_tc_kernel, maps, p);
}

}  // namespace ev
