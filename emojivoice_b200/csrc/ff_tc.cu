// Fused feed-forward of the decoder's BasicTransformerBlock (transformer.py:296-316 with the SnakeBeta activation,
// transformer.py:13-83) on tcgen05: per 128-row tile ONE kernel does
//     n = LayerNorm3(x);  h = SnakeBeta(n W1^T + b1);  y = (x + h W2^T + b2) * mask   -> bf16 operand of the next conv
// The 1024-wide hidden activation never leaves the SM: it is produced 128 channels at a time into TMEM (double-buffered
// accumulator), activated by the epilogue warps into a bf16 K-major operand tile in shared memory, and consumed by the
// second GEMM, whose 256-wide accumulator stays in TMEM for the whole tile.  Replaces three launches (layer_norm,
// conv_tc ff1, conv_tc ff2) and the 2 x 2 KB per row HBM round trip of the hidden tensor.
//
// Per CTA (persistent over 128-row tiles, one per SM):
//   shared memory  A  [4 planes x 128 rows x 128 B]  LN output, bf16, 128B-swizzled K-major (written by hand)     64 KB
//                  P  [2 buffers x 2 planes x 128 rows x 128 B]  SnakeBeta output of one 128-channel chunk          64 KB
//                  W  ring of 6 x 16 KB weight tiles (TMA, box 64 k x 128 n): W1 chunk = 4 tiles, W2 chunk = 4 tiles  96 KB
//   TMEM           acc1 [2 x 128 columns] (hidden chunk c in buffer c & 1), acc2 [256 columns]
//   warp 0  : TMA producer -- weights only, which no kernel of the stream writes: with the dense grid it does NOT wait for the
//             programmatic dependency, so the first six weight tiles land while the previous kernel is still draining (with a
//             tile list, which is data of the stream, it waits like everybody else)
//   warp 1  : MMA issuer (owns TMEM).  Order  G1(0) G1(1) | G1(2) G2(0) | G1(3) G2(1) | ... | G2(6) | G2(7): the tensor pipe
//             always has the next chunk's first GEMM to run while the epilogue warps activate the current one
//   warps 2-17 : LayerNorm prologue (warp = row), per-chunk activation (thread = row, 32 channels), final epilogue.
// Work per tile: 2 x 128 x 256 x 1024 MACs = 16.4 k clk of tcgen05 at N = 128 (64 clk per MMA); weight stream 1 MB per
// tile from L2.
//
// Attention tail mode (FfTcArgs::att != nullptr): the attention's out-projection and its residual add (transformer.py:
// 283-294) run in front, in the same launch:  x = xr + att Wo^T + bo.  The producer loads the tile's 128 x 128 bf16
// attention output into the idle P[0] buffer, the issuer multiplies it with Wo (16 MMAs) INTO acc2, the workers add the
// bias and the fp32 residual stream -- stored CHANNEL-FIRST by resnet_tc, so that thread = row reads of a column are
// coalesced -- write x back into acc2 (tcgen05.st) and normalise it from there (thread = row, row sums combined across
// the four column-slot warps through the idle P[1] buffer).  GEMM 2 then accumulates ON TOP of x, so the final residual
// add is free and x is read from HBM once.  Replaces the out-projection launch and its fp32 read-modify-write of xr.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>

#include "conv.cuh"
#include "tc_ptx.cuh"

namespace ev {
using namespace tc;
namespace {

constexpr int FF_D = 256;                   // model width (K of GEMM 1, N of GEMM 2)
constexpr int FF_CHUNK = 128;               // hidden channels per chunk
constexpr int FF_SLOTS = 6;
constexpr int FF_PLANE = 128 * 128;         // 128 rows x 64 bf16
constexpr int FF_A_BYTES = 4 * FF_PLANE;
constexpr int FF_P_BYTES = 2 * FF_PLANE;
constexpr int FF_W_TILE = FF_PLANE;
constexpr int FF_SMEM = FF_A_BYTES + 2 * FF_P_BYTES + FF_SLOTS * FF_W_TILE + 1024;
constexpr int FF_EPI_WARPS = 16;
constexpr int FF_THREADS = 32 * (2 + FF_EPI_WARPS);

struct FfMaps { CUtensorMap w1, w2, wo, att; };

struct FfParams {
  const float* x; long long x_bs;           // fp32 residual stream (b, t, 256), dense rows (att_mode == 0)
  const float* xr_cf; const float* bo;      // att_mode: fp32 residual stream (b, 256, t) channel-first, out-projection bias [256]
  int att_mode;
  const float* ln_g; const float* ln_b; float eps;
  const float* b1; const float* snake_a; const float* snake_invb;   // [n_chunks * 128]
  const float* b2;                          // [256]
  bf16* out; long long out_ld, out_bs;
  const int* lens; int len_shift;
  const int* tiles;                         // compact (item, m-tile) list or nullptr = dense grid
  int B, T, m_tiles, total_tiles, n_chunks;
  int debug;                                // EV_FF_DEBUG (timing experiments only, results are wrong): 1 no SnakeBeta math, 2 no weight
                                            // loads, 4 no LayerNorm loads, 8 no output pass
};

__device__ __forceinline__ uint32_t d_hi(uint32_t sbo, uint32_t layout) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t d_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t d_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// EV_FF_DEBUG & 16: clock64 stamps of CTA 0's first tile (issuer: slots 0.., worker warp 0: slots 64..), read by ev_test_ff_trace
__device__ unsigned long long g_ff_trace[192];
#define FF_TR(i) do { if ((p.debug & 16) && blockIdx.x == 0 && lane == 0 && it == 0) g_ff_trace[(i)] = (unsigned long long)clock64(); } while (0)

__global__ void __launch_bounds__(FF_THREADS, 1)
ff_tc_kernel(const __grid_constant__ FfMaps maps, const __grid_constant__ FfParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[FF_SLOTS], w_empty[FF_SLOTS];
  __shared__ __align__(8) uint64_t a_ready, a_free, acc1_full[2], acc1_free[2], p_ready[2], p_free[2], acc2_full, acc2_empty;
  __shared__ __align__(8) uint64_t att_full, att_free, oproj_full;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_s = base, p_s = base + FF_A_BYTES, w_s = p_s + 2 * FF_P_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = p.n_chunks;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w2) : "memory");
    for (int s = 0; s < FF_SLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(&a_ready, FF_EPI_WARPS); mbar_init(&a_free, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc1_full[s], 1); mbar_init(&acc1_free[s], FF_EPI_WARPS); mbar_init(&p_ready[s], FF_EPI_WARPS); mbar_init(&p_free[s], 1); }
    mbar_init(&acc2_full, 1); mbar_init(&acc2_empty, FF_EPI_WARPS);
    mbar_init(&att_full, 1); mbar_init(&att_free, 1); mbar_init(&oproj_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base + 256u;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- weight producer (no dependency on the previous kernel: weights are constant)
      int sl = 0;
      uint32_t ph = 1;
      auto load_w1 = [&](int c) {
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(&w_empty[sl], ph);
          if (p.debug & 2) { mbar_arrive(&w_full[sl]); if (++sl == FF_SLOTS) { sl = 0; ph ^= 1u; } continue; }
          mbar_expect_tx(&w_full[sl], (uint32_t)FF_W_TILE);
          tma_load_3d(w_s + (uint32_t)(sl * FF_W_TILE), &maps.w1, &w_full[sl], kc * 64, c * FF_CHUNK, 0);
          if (++sl == FF_SLOTS) { sl = 0; ph ^= 1u; }
        }
      };
      auto load_w2 = [&](int c) {
        for (int kc = 0; kc < 2; ++kc)
          for (int nh = 0; nh < 2; ++nh) {
            mbar_wait(&w_empty[sl], ph);
            if (p.debug & 2) { mbar_arrive(&w_full[sl]); if (++sl == FF_SLOTS) { sl = 0; ph ^= 1u; } continue; }
            mbar_expect_tx(&w_full[sl], (uint32_t)FF_W_TILE);
            tma_load_3d(w_s + (uint32_t)(sl * FF_W_TILE), &maps.w2, &w_full[sl], c * FF_CHUNK + kc * 64, nh * 128, 0);
            if (++sl == FF_SLOTS) { sl = 0; ph ^= 1u; }
          }
      };
      // weights are constants, so the dense grid does not wait for the previous kernel at all; a tile list is data of the
      // stream and is read behind the programmatic dependency
      if (p.tiles || p.att_mode) pdl_wait();
      const int total_tiles = p.tiles ? __ldg(p.tiles) : p.total_tiles;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        if (p.att_mode) {
          // the tile's attention output -> P[0] (free once the previous tile's last GEMM 2 on that buffer has completed),
          // then the four out-projection weight tiles
          int b, m0;
          if (p.tiles) { const int pair = __ldg(p.tiles + 1 + tile); b = pair >> 16; m0 = (pair & 0xffff) * 128; }
          else { b = tile / p.m_tiles; m0 = (tile - b * p.m_tiles) * 128; }
          mbar_wait(&att_free, ((uint32_t)it & 1u) ^ 1u);
          mbar_expect_tx(&att_full, (uint32_t)FF_P_BYTES);
          tma_load_3d(p_s, &maps.att, &att_full, 0, m0, b);
          tma_load_3d(p_s + (uint32_t)FF_PLANE, &maps.att, &att_full, 64, m0, b);
          for (int kc = 0; kc < 2; ++kc)
            for (int nh = 0; nh < 2; ++nh) {
              mbar_wait(&w_empty[sl], ph);
              mbar_expect_tx(&w_full[sl], (uint32_t)FF_W_TILE);
              tma_load_3d(w_s + (uint32_t)(sl * FF_W_TILE), &maps.wo, &w_full[sl], kc * 64, nh * 128, 0);
              if (++sl == FF_SLOTS) { sl = 0; ph ^= 1u; }
            }
        }
        load_w1(0);
        if (n_chunks > 1) load_w1(1);
        for (int c = 0; c < n_chunks; ++c) {
          if (c + 2 < n_chunks) load_w1(c + 2);
          load_w2(c);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer: warp-uniform control flow, one elected lane issues
    constexpr uint32_t idesc = make_idesc(128, 128);
    const uint32_t hi = d_hi(1024u, 2u);
    const uint32_t a_lo0 = d_lo(a_s), p_lo0 = d_lo(p_s), w_lo0 = d_lo(w_s);
    constexpr uint32_t plane16 = (uint32_t)FF_PLANE >> 4, pbuf16 = (uint32_t)FF_P_BYTES >> 4;
    int sl = 0;
    uint32_t wph = 0;
    int it = 0;
    auto gemm1 = [&](int c, bool last) {      // acc1[c & 1] = A (128 x 256) x W1[c]^T
      const uint32_t d = acc1 + (uint32_t)((c & 1) * FF_CHUNK);
      // the workers have read this buffer's previous contents (chunk c - 2, or the previous tile's last chunks)
      mbar_wait(&acc1_free[c & 1], ((uint32_t)(it * (n_chunks >> 1) + (c >> 1)) & 1u) ^ 1u);
      tcgen05_fence_after();
      for (int kc = 0; kc < 4; ++kc) {
        mbar_wait(&w_full[sl], wph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)kc * plane16, b_lo = w_lo0 + (uint32_t)sl * plane16;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(d, d_join(hi, a_lo + 2u * ks), d_join(hi, b_lo + 2u * ks), idesc, (kc | ks) ? 1u : 0u);
          umma_commit(&w_empty[sl]);
          if (kc == 3) {
            umma_commit(&acc1_full[c & 1]);
            if (last) umma_commit(&a_free);   // every GEMM-1 MMA of the tile has read A: the next tile's LayerNorm may overwrite it
          }
        }
        __syncwarp();
        if (++sl == FF_SLOTS) { sl = 0; wph ^= 1u; }
      }
    };
    auto gemm2 = [&](int c, bool last) {      // acc2 (128 x 256) += P[c & 1] (128 x 128) x W2[:, chunk c]^T
      const uint32_t pb = p_lo0 + (uint32_t)(c & 1) * pbuf16;
      for (int kc = 0; kc < 2; ++kc)
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&w_full[sl], wph);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = pb + (uint32_t)kc * plane16, b_lo = w_lo0 + (uint32_t)sl * plane16;
            const uint32_t d = acc2 + (uint32_t)(nh * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(d, d_join(hi, a_lo + 2u * ks), d_join(hi, b_lo + 2u * ks), idesc, (p.att_mode | c | kc | ks) ? 1u : 0u);
            umma_commit(&w_empty[sl]);
            if (kc == 1 && nh == 1) {
              umma_commit(&p_free[c & 1]);
              if (c + 2 == n_chunks || n_chunks == 1) umma_commit(&att_free);   // last use of P[0] in this tile
              if (last) umma_commit(&acc2_full);
            }
          }
          __syncwarp();
          if (++sl == FF_SLOTS) { sl = 0; wph ^= 1u; }
        }
    };
    if (p.tiles) pdl_wait();
    const int total_tiles = p.tiles ? __ldg(p.tiles) : p.total_tiles;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      FF_TR(0);
      if (p.att_mode) {
        // acc2 = att (128 x 128, in P[0]) x Wo^T: the out-projection; GEMM 2 accumulates on top of it later
        mbar_wait(&att_full, (uint32_t)it & 1u);
        mbar_wait(&acc2_empty, ((uint32_t)it & 1u) ^ 1u);                // the previous tile's output has left acc2
        tcgen05_fence_after();
        for (int kc = 0; kc < 2; ++kc)
          for (int nh = 0; nh < 2; ++nh) {
            mbar_wait(&w_full[sl], wph);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = p_lo0 + (uint32_t)kc * plane16, b_lo = w_lo0 + (uint32_t)sl * plane16;
              const uint32_t d = acc2 + (uint32_t)(nh * 128);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(d, d_join(hi, a_lo + 2u * ks), d_join(hi, b_lo + 2u * ks), idesc, (kc | ks) ? 1u : 0u);
              umma_commit(&w_empty[sl]);
              if (kc == 1 && nh == 1) umma_commit(&oproj_full);
            }
            __syncwarp();
            if (++sl == FF_SLOTS) { sl = 0; wph ^= 1u; }
          }
      }
      mbar_wait(&a_ready, (uint32_t)it & 1u);
      tcgen05_fence_after();
      FF_TR(1);
      const int half = (n_chunks + 1) >> 1;   // uses of each acc1 / P buffer per tile
      gemm1(0, n_chunks == 1);
      FF_TR(2);
      if (n_chunks > 1) gemm1(1, n_chunks == 2);
      FF_TR(3);
      for (int c = 0; c < n_chunks; ++c) {
        // G1(c+2) only needs chunk c's accumulator to have been READ: it runs under the workers' activation of chunk c
        if (c + 2 < n_chunks) gemm1(c + 2, c + 3 == n_chunks);
        FF_TR(10 + 4 * c);
        const uint32_t u = (uint32_t)(it * half + (c >> 1));
        mbar_wait(&p_ready[c & 1], u & 1u);
        if (c == 0 && !p.att_mode) mbar_wait(&acc2_empty, ((uint32_t)it & 1u) ^ 1u);   // the previous tile's output has left acc2
        tcgen05_fence_after();
        FF_TR(8 + 4 * c);
        gemm2(c, c == n_chunks - 1);
        FF_TR(9 + 4 * c);
      }
    }
  } else {
    // ---------------- 16 worker warps
    const int ew = warp - 2, q = warp & 3, j = ew >> 2;   // TMEM lane quadrant q, 32-column slot j of a 128-column chunk
    const uint32_t lane_q = (uint32_t)(q * 32) << 16;
    uint8_t* a_gen = base_gen;
    uint8_t* p_gen = base_gen + FF_A_BYTES;
    // LayerNorm affine parameters of this lane's 8 channels (4 at 4*lane, 4 at 128 + 4*lane): constants, fetched before the
    // programmatic-dependency wait
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + 32 + lane);
    const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + lane), be1 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 32 + lane);
    // b1 / snake parameters of chunk c are three 128-byte lines per warp; lanes 0-2 pull the next chunk's lines into L1 while
    // the current chunk is processed (a cold miss per chunk otherwise sits on the workers' critical path)
    auto prefetch_params = [&](int c) {
      if (lane < 3) {
        const float* base_p = lane == 0 ? p.b1 : (lane == 1 ? p.snake_a : p.snake_invb);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(base_p + c * FF_CHUNK + j * 32));
      }
    };
    prefetch_params(0);
    pdl_wait();                                            // x is the previous kernel's output
    { const int it = 0; if (ew == 0) FF_TR(64); }
    const int half = (n_chunks + 1) >> 1;
    int it = 0;
    const int total_tiles = p.tiles ? __ldg(p.tiles) : p.total_tiles;
    if (p.tiles) {
      // tiles without a valid row are not in the list: their output rows are zero (y * mask); warp per row, 16 B per lane
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const int frames = (__ldg(p.lens + b) + (1 << p.len_shift) - 1) >> p.len_shift;
        const int first = ((min(max(frames, 0), p.T) + 127) >> 7) << 7;
        for (int t = first + ew; t < p.T; t += FF_EPI_WARPS)
          reinterpret_cast<uint4*>(p.out + b * p.out_bs + (long long)t * p.out_ld)[lane] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int b, m0;
      if (p.tiles) { const int pair = __ldg(p.tiles + 1 + tile); b = pair >> 16; m0 = (pair & 0xffff) * 128; }
      else { b = tile / p.m_tiles; m0 = (tile - b * p.m_tiles) * 128; }
      const float* xb = p.x + b * p.x_bs;
      if (p.att_mode) {
        // ---- x = out-projection (acc2) + bo + residual stream -> back into acc2; LayerNorm3(x) -> operand A.  thread = row.
        const int rowq = q * 32 + lane, t = m0 + rowq;
        const bool in_seq = t < p.T;
        const float* xcol = p.xr_cf + ((long long)b * FF_D) * p.T + t;
        float* rpart = reinterpret_cast<float*>(p_gen + FF_P_BYTES);          // idle P[1]: [128 rows][4 slots][2]
        mbar_wait(&oproj_full, (uint32_t)it & 1u);
        tcgen05_fence_after();
        float s = 0.0f, qq = 0.0f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col0 = (j + 4 * h) * 32;
          float xv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) xv[i] = in_seq ? xcol[(long long)(col0 + i) * p.T] : 0.0f;   // a column of 32 rows: 128 B per warp
          uint32_t raw[32];
          tmem_ld32(acc2 + lane_q + (uint32_t)col0, raw);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const int c = hh * 16 + i;
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bo + col0 + c));
              const float v0 = __uint_as_float(raw[c]) + bb.x + xv[i], v1 = __uint_as_float(raw[c + 1]) + bb.y + xv[i + 1];
              const float v2 = __uint_as_float(raw[c + 2]) + bb.z + xv[i + 2], v3 = __uint_as_float(raw[c + 3]) + bb.w + xv[i + 3];
              s += (v0 + v1) + (v2 + v3);
              qq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, qq))));
              raw[c] = __float_as_uint(v0); raw[c + 1] = __float_as_uint(v1); raw[c + 2] = __float_as_uint(v2); raw[c + 3] = __float_as_uint(v3);
            }
            if (hh == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] = in_seq ? xcol[(long long)(col0 + 16 + i) * p.T] : 0.0f;
            }
          }
          tmem_st32(acc2 + lane_q + (uint32_t)col0, raw);
        }
        *reinterpret_cast<float2*>(rpart + (rowq * 4 + j) * 2) = make_float2(s, qq);
        tmem_st_wait();
        tcgen05_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * FF_EPI_WARPS) : "memory");
        tcgen05_fence_after();
        const float4 pa = *reinterpret_cast<const float4*>(rpart + rowq * 8), pb = *reinterpret_cast<const float4*>(rpart + rowq * 8 + 4);
        const float mean = ((pa.x + pa.z) + (pb.x + pb.z)) * (1.0f / FF_D);
        const float var = fmaxf(((pa.y + pa.w) + (pb.y + pb.w)) * (1.0f / FF_D) - mean * mean, 0.0f);
        const float rstd = in_seq ? rsqrtf(var + p.eps) : 0.0f;             // rows past the sequence: finite values, never stored
        const float nmu = -mean * rstd;
        mbar_wait(&a_free, ((uint32_t)it & 1u) ^ 1u);                        // the previous tile's GEMM 1 has finished reading A
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int cb = j + 4 * h, col0 = cb * 32;
          uint32_t raw[32];
          tmem_ld32(acc2 + lane_q + (uint32_t)col0, raw);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_g + col0 + i)), be = __ldg(reinterpret_cast<const float4*>(p.ln_b + col0 + i));
            const float v0 = fmaf(fmaf(__uint_as_float(raw[i]), rstd, nmu), ga.x, be.x), v1 = fmaf(fmaf(__uint_as_float(raw[i + 1]), rstd, nmu), ga.y, be.y);
            const float v2 = fmaf(fmaf(__uint_as_float(raw[i + 2]), rstd, nmu), ga.z, be.z), v3 = fmaf(fmaf(__uint_as_float(raw[i + 3]), rstd, nmu), ga.w, be.w);
            __nv_bfloat162 e0 = __floats2bfloat162_rn(v0, v1), e1 = __floats2bfloat162_rn(v2, v3);
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&e0);
            pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&e1);
          }
          uint8_t* rp = a_gen + (cb >> 1) * FF_PLANE + rowq * 128;
          const int c16 = (cb & 1) * 4, sw = rowq & 7;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rp + (((c16 + i) ^ sw) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        tcgen05_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready);
      } else
      // ---- LayerNorm: warp ew normalises rows ew*8 .. ew*8+7 of the tile into the swizzled bf16 operand A
      {
#pragma unroll 1
        for (int hb4 = 0; hb4 < 8; hb4 += 4) {
        float4 v0[4], v1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = m0 + ew * 8 + hb4 + i;
          if (t < p.T && !(p.debug & 4)) {
            const float4* row = reinterpret_cast<const float4*>(xb + (long long)t * FF_D);
            v0[i] = row[lane]; v1[i] = row[32 + lane];
          } else { v0[i] = make_float4(0.f, 0.f, 0.f, 0.f); v1[i] = v0[i]; }
        }
        if (hb4 == 0) mbar_wait(&a_free, ((uint32_t)it & 1u) ^ 1u);      // the previous tile's GEMM 1 has finished reading A
        if (ew == 0) FF_TR(hb4 == 0 ? 65 : 67);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = ew * 8 + hb4 + i;
          const float4 a = v0[i], c = v1[i];
          const float mean = warp_sum((a.x + a.y) + (a.z + a.w) + (c.x + c.y) + (c.z + c.w)) * (1.0f / FF_D);
          const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
          const float d4 = c.x - mean, d5 = c.y - mean, d6 = c.z - mean, d7 = c.w - mean;
          const float var = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) + (d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7)) * (1.0f / FF_D);
          const float rstd = (m0 + r < p.T) ? 1.0f / sqrtf(var + p.eps) : 0.0f;
          const float k = (m0 + r < p.T) ? 1.0f : 0.0f;    // rows past the sequence: exact zeros (their outputs are never stored)
          __nv_bfloat162 l0 = __floats2bfloat162_rn((d0 * rstd * g0.x + be0.x) * k, (d1 * rstd * g0.y + be0.y) * k);
          __nv_bfloat162 l1 = __floats2bfloat162_rn((d2 * rstd * g0.z + be0.z) * k, (d3 * rstd * g0.w + be0.w) * k);
          __nv_bfloat162 h0 = __floats2bfloat162_rn((d4 * rstd * g1.x + be1.x) * k, (d5 * rstd * g1.y + be1.y) * k);
          __nv_bfloat162 h1 = __floats2bfloat162_rn((d6 * rstd * g1.z + be1.z) * k, (d7 * rstd * g1.w + be1.w) * k);
          // channel ch = 4*lane (+128): plane ch/64, 16-byte chunk (ch%64)/8, 8-byte half (lane & 1)
          const int plane = lane >> 4, c16 = (lane & 15) >> 1, hb = (lane & 1) * 8;
          const uint32_t off = (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4) + hb);
          uint2 lo2, hi2;
          lo2.x = *reinterpret_cast<uint32_t*>(&l0); lo2.y = *reinterpret_cast<uint32_t*>(&l1);
          hi2.x = *reinterpret_cast<uint32_t*>(&h0); hi2.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(a_gen + plane * FF_PLANE + off) = lo2;
          *reinterpret_cast<uint2*>(a_gen + (plane + 2) * FF_PLANE + off) = hi2;
        }
        }
        fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready);
        if (ew == 0) FF_TR(66);
      }
      // ---- per chunk: acc1 -> + b1 -> SnakeBeta -> bf16 operand P
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        const uint32_t u = (uint32_t)(it * half + (c >> 1));
        prefetch_params(c + 1 < n_chunks ? c + 1 : 0);
        mbar_wait(&acc1_full[buf], u & 1u);
        tcgen05_fence_after();
        if (ew == 0) FF_TR(72 + 6 * c);
        uint32_t raw[32];
        tmem_ld32(acc1 + lane_q + (uint32_t)(buf * FF_CHUNK + j * 32), raw);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_free[buf]);       // the issuer may start chunk c+2's first GEMM into this buffer
        if (ew == 0) FF_TR(73 + 6 * c);
        const int ch0 = c * FF_CHUNK + j * 32;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + ch0 + i));
          const float4 sa = __ldg(reinterpret_cast<const float4*>(p.snake_a + ch0 + i));
          const float4 sb = __ldg(reinterpret_cast<const float4*>(p.snake_invb + ch0 + i));
          float v0 = __uint_as_float(raw[i]) + bb.x, v1 = __uint_as_float(raw[i + 1]) + bb.y;
          float v2 = __uint_as_float(raw[i + 2]) + bb.z, v3 = __uint_as_float(raw[i + 3]) + bb.w;
          if (!(p.debug & 1)) {
          const float s0 = __sinf(v0 * sa.x), s1 = __sinf(v1 * sa.y), s2 = __sinf(v2 * sa.z), s3 = __sinf(v3 * sa.w);
          v0 = fmaf(sb.x, s0 * s0, v0); v1 = fmaf(sb.y, s1 * s1, v1); v2 = fmaf(sb.z, s2 * s2, v2); v3 = fmaf(sb.w, s3 * s3, v3);
          }
          __nv_bfloat162 e0 = __floats2bfloat162_rn(v0, v1), e1 = __floats2bfloat162_rn(v2, v3);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&e0);
          pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&e1);
        }
        if (ew == 0) FF_TR(74 + 6 * c);
        mbar_wait(&p_free[buf], (u & 1u) ^ 1u);            // GEMM 2 of chunk c-2 has finished reading this buffer
        if (ew == 0) FF_TR(75 + 6 * c);
        {
          uint8_t* rp = p_gen + buf * FF_P_BYTES + (j >> 1) * FF_PLANE + row * 128;
          const int c16 = (j & 1) * 4, sw = row & 7;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rp + (((c16 + i) ^ sw) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        tcgen05_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[buf]);
        if (ew == 0) FF_TR(76 + 6 * c);
      }
      // ---- output: y = (x + acc2 + b2) * mask -> bf16; each warp owns two 32-column blocks (j and j + 4) of its rows
      mbar_wait(&acc2_full, (uint32_t)it & 1u);
      tcgen05_fence_after();
      if (ew == 0) FF_TR(130);
      {
        // Each warp owns two 32 x 32 blocks (columns (j + 4h) * 32).  thread = row after tcgen05.ld; a private 4 KB tile in
        // the idle A region (every GEMM 1 of the tile has completed; the next tile's LayerNorm is these warps' own next step;
        // P[0] may already hold the next tile's attention output) -- XOR-swizzled 16-byte chunks, conflict-free both ways --
        // turns that into a row-wise pass: 8 lanes x float4 per row, so the residual reads (128 B per row) and the bf16
        // stores (64 B per row) are coalesced.
        uint8_t* stg = a_gen + ew * 4096;
        const int sub = lane >> 3, cl = lane & 7;
        const int len_b = p.lens ? __ldg(p.lens + b) : 0x7fffffff;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col = (j + 4 * h) * 32 + cl * 4;
          float4 rr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {                      // residual loads fly while the accumulator is fetched and staged
            const int t = m0 + q * 32 + u * 4 + sub;
            rr[u] = (t < p.T && !p.att_mode && !(p.debug & 8)) ? *reinterpret_cast<const float4*>(xb + (long long)t * FF_D + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const float4 bias = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
          if (ew == 0) FF_TR(132 + 4 * h);
          uint32_t raw[32];
          tmem_ld32(acc2 + lane_q + (uint32_t)((j + 4 * h) * 32), raw);
          if (ew == 0) FF_TR(133 + 4 * h);
          if (h == 1) {                                      // last TMEM read of the tile: hand acc2 back to the issuer
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc2_empty);
          }
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((c8 ^ (lane & 7)) << 4)) = make_uint4(raw[4 * c8], raw[4 * c8 + 1], raw[4 * c8 + 2], raw[4 * c8 + 3]);
          __syncwarp();
          if (ew == 0) FF_TR(134 + 4 * h);
          bf16* dst = p.out + b * p.out_bs + col;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int rl = u * 4 + sub, t = m0 + q * 32 + rl;
            if (t >= p.T || (p.debug & 8)) continue;
            const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((cl ^ (rl & 7)) << 4));
            const bool valid = (t << p.len_shift) < len_b;     // a select, not a multiply: padded rows may hold anything
            __nv_bfloat162 o0 = __floats2bfloat162_rn(a.x + bias.x + rr[u].x, a.y + bias.y + rr[u].y);
            __nv_bfloat162 o1 = __floats2bfloat162_rn(a.z + bias.z + rr[u].z, a.w + bias.w + rr[u].w);
            uint2 pk2;
            pk2.x = valid ? *reinterpret_cast<uint32_t*>(&o0) : 0u; pk2.y = valid ? *reinterpret_cast<uint32_t*>(&o1) : 0u;
            *reinterpret_cast<uint2*>(dst + (long long)t * p.out_ld) = pk2;
          }
          __syncwarp();                                      // the staging tile is rewritten by the next block
          if (ew == 0) FF_TR(135 + 4 * h);
        }
      }
      if (ew == 0) FF_TR(131);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

cudaError_t ff_tc_read_trace(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_ff_trace, sizeof(unsigned long long) * std::min(n, 192));
}

bool ff_tc_supported(const ConvWeights& ff1, const ConvWeights& ff2) {
  static const int mode = []() { const char* v = getenv("EV_FF_FUSE"); return v ? atoi(v) : 1; }();   // EV_FF_FUSE=0: three launches
  return mode != 0 && ff1.w_bf16 && ff2.w_bf16 && ff1.taps == 1 && ff2.taps == 1 && ff1.C_in == FF_D && ff2.N == FF_D &&
         ff1.N == ff2.C_in && ff1.N % (2 * FF_CHUNK) == 0 && ff1.K_pad == FF_D && ff2.K_pad == ff1.N &&
         ff1.N_pad_tc == ff1.N && ff2.N_pad_tc == FF_D && ff1.bias && ff2.bias;
}

// the attention out-projection the tail mode can run in front: Linear(128 -> 256) with a bias
bool ff_tc_oproj_supported(const ConvWeights& wo) {
  return wo.w_bf16 && wo.bias && wo.taps == 1 && wo.C_in == 128 && wo.K_pad == 128 && wo.N == FF_D && wo.N_pad_tc == FF_D;
}

cudaError_t ff_tc_launch(const FfTcArgs& a, cudaStream_t s, std::string* err) {
  const ConvWeights& w1 = *a.ff1;
  const ConvWeights& w2 = *a.ff2;
  const bool att_mode = a.att != nullptr;
  if (!ff_tc_supported(w1, w2) || (a.out_ld & 7) || (a.out_bs & 7) || (reinterpret_cast<uintptr_t>(a.out) & 15) ||
      (reinterpret_cast<uintptr_t>(a.x) & 15) || (!att_mode && !a.x) ||
      (att_mode && (!a.oproj || !ff_tc_oproj_supported(*a.oproj) || !a.xr_cf || (a.att_ld & 7) || (a.att_bs & 7) ||
                    (reinterpret_cast<uintptr_t>(a.att) & 15)))) {
    if (err) *err = "ff_tc: unsupported layer shape or alignment";
    return cudaErrorInvalidValue;
  }
  FfMaps maps;
  if (!tc_encode_bf16_map(&maps.w1, w1.w_bf16, (uint64_t)w1.K_pad, (uint64_t)w1.N_pad_tc, 1, (uint64_t)w1.K_pad * 2,
                          (uint64_t)w1.K_pad * w1.N_pad_tc * 2, 64u, 128u, 128, err))
    return cudaErrorInvalidValue;
  if (!tc_encode_bf16_map(&maps.w2, w2.w_bf16, (uint64_t)w2.K_pad, (uint64_t)w2.N_pad_tc, 1, (uint64_t)w2.K_pad * 2,
                          (uint64_t)w2.K_pad * w2.N_pad_tc * 2, 64u, 128u, 128, err))
    return cudaErrorInvalidValue;
  if (att_mode) {
    const ConvWeights& wo = *a.oproj;
    if (!tc_encode_bf16_map(&maps.wo, wo.w_bf16, (uint64_t)wo.K_pad, (uint64_t)wo.N_pad_tc, 1, (uint64_t)wo.K_pad * 2,
                            (uint64_t)wo.K_pad * wo.N_pad_tc * 2, 64u, 128u, 128, err))
      return cudaErrorInvalidValue;
    if (!tc_encode_bf16_map(&maps.att, a.att, 128u, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.att_ld * 2, (uint64_t)a.att_bs * 2, 64u, 128u, 128, err))
      return cudaErrorInvalidValue;
  } else {
    maps.wo = maps.w2; maps.att = maps.w2;
  }
  FfParams p{};
  p.x = a.x; p.x_bs = (long long)a.T * FF_D;
  p.att_mode = att_mode ? 1 : 0; p.xr_cf = a.xr_cf; p.bo = att_mode ? a.oproj->bias : nullptr;
  p.ln_g = a.ln_g; p.ln_b = a.ln_b; p.eps = a.eps;
  p.b1 = w1.bias; p.snake_a = a.snake_a; p.snake_invb = a.snake_invb; p.b2 = w2.bias;
  p.out = a.out; p.out_ld = a.out_ld; p.out_bs = a.out_bs;
  p.lens = a.lens; p.len_shift = a.len_shift;
  p.tiles = a.lens ? a.tiles : nullptr; p.B = a.B;
  { static const int dbg = []() { const char* v = getenv("EV_FF_DEBUG"); return v ? atoi(v) : 0; }(); p.debug = dbg; }
  p.T = a.T; p.m_tiles = ceil_div(a.T, 128); p.total_tiles = p.m_tiles * a.B; p.n_chunks = w1.N / FF_CHUNK;
  static DeviceOnce once;
  cudaError_t ce_attr = once.run([&]() { return cudaFuncSetAttribute(ff_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM); });
  if (ce_attr != cudaSuccess) return ce_attr;
  const int grid = std::min(p.total_tiles, tc_sm_count());
  return launch_pdl(ff_tc_kernel, dim3(grid), dim3(FF_THREADS), (size_t)FF_SMEM, s, maps, p);
}

}  // namespace ev
