// bf16 implicit-GEMM conv1d on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.  sm_100a only.
//
// Tile: 128 output time steps (UMMA M = 128 = TMEM lanes) x BN output channels (UMMA N = BN TMEM columns),
// reduced over ceil(C_in/64) K-chunks x taps (4 x UMMA K=16 per weight tile).
//   A (activations) : 3-D tensor map (channel, time, batch) over the channel-last bf16 tensor.  ONE haloed box
//                     [64 ch x (128 + (taps-1)*dilation) rows] is fetched per K-chunk at row m0 - pad; every tap's MMA
//                     reads the same smem tile through a descriptor whose start address is advanced by
//                     tap*dilation rows (the 128B swizzle is a function of absolute smem address bits, so no
//                     base-offset is needed -- verified on hardware).  Rows outside [0,T) are zero-filled by TMA,
//                     which IS the convolution's zero padding (no im2col, no halo copies in HBM).
//                     The stride-2 conv reads a (T/2, 2*ld) view of the same memory (tap_col selects even/odd rows).
//   B (weights)     : 3-D tensor map (c_in, n, tap) over [taps][N_pad][K_pad] bf16, box [64 x BN].
// PERSISTENT CTAs (one per SM) walk the tile list; TMEM holds two accumulators so the epilogue of tile i overlaps
// the TMA/MMA main loop of tile i+1.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue: tcgen05.ld -> fp32 staging tile in smem -> row-wise coalesced pass with the fused
// bias / mask / Euler-or-residual / MRF-mean / activation and vectorised fp32 + bf16 stores.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>

#include "conv.cuh"
#include "tc_ptx.cuh"

namespace ev {
using namespace tc;
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int MAX_A_SLOTS = 4, MAX_B_SLOTS = 8;
constexpr int kMaxGroups = kMaxTaps;

struct TcParams {
  ConvGeom g;
  Epilogue e;
  // taps are processed in groups that share one activation tile: with halo reuse all taps form one group whose
  // tile covers rows [m0 + grp_row0, m0 + grp_row0 + a_rows); otherwise every tap is its own 128-row group.
  int n_groups;
  int grp_row0[kMaxGroups], grp_col0[kMaxGroups], grp_first[kMaxGroups], grp_count[kMaxGroups];
  int tap_byte_off[kMaxTaps];   // byte offset of tap j's first row inside its group's tile (rows are 128 B)
  int a_rows;                   // TMA box rows of the activation tile (128 + halo, multiple of 8)
  int a_slot_bytes;             // a_rows*128 rounded up to 1024
  int a_slots, b_slots;
  int kchunks;
  int vec_ok;
  int m_tiles, n_tiles, total_tiles;
  int resident;                 // all weight tiles stay in smem for the CTA's lifetime (narrow layers)
};

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_LD = BN + 4;                         // floats per staged accumulator row
  static constexpr int STAGING_BYTES = BM * STAGE_LD * 4;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulators
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[MAX_A_SLOTS], a_empty[MAX_A_SLOTS], b_full[MAX_B_SLOTS], b_empty[MAX_B_SLOTS];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t tiles0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = tiles0, b_base = tiles0 + (uint32_t)(p.a_slots * p.a_slot_bytes);
  float* stage = reinterpret_cast<float*>(smem_raw + (tiles0 - smem_u32(smem_raw)) + p.a_slots * p.a_slot_bytes + p.b_slots * C::B_TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int A_SLOTS = p.a_slots, B_SLOTS = p.b_slots;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < MAX_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < MAX_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: per tile, per K-chunk, per tap group: one (haloed) activation tile, then one
      // weight tile per tap.  Ring positions run on across tiles, so the next tile's loads start while this one computes.
      int ai = 0, bi = 0;
      if (p.resident) {   // narrow layers: every (K-chunk, tap) weight tile is fetched once and kept
        const int n_w = p.kchunks * p.g.taps;
        mbar_expect_tx(&b_full[0], (uint32_t)(n_w * C::B_TILE_BYTES));
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int j = 0; j < p.g.taps; ++j)
            tma_load_3d(b_base + (uint32_t)((kc * p.g.taps + j) * C::B_TILE_BYTES), &tmB, &b_full[0], kc * BK, 0, j);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = (tile / p.n_tiles) % p.m_tiles, b = tile / (p.n_tiles * p.m_tiles);
        const int m0 = mt * BM, n0 = nt * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int g = 0; g < p.n_groups; ++g, ++ai) {
            const int sa = ai % A_SLOTS;
            mbar_wait(&a_empty[sa], ((uint32_t)(ai / A_SLOTS) & 1u) ^ 1u);
            mbar_expect_tx(&a_full[sa], (uint32_t)(p.a_rows * BK * 2));
            tma_load_3d(a_base + (uint32_t)(sa * p.a_slot_bytes), &tmA, &a_full[sa], p.grp_col0[g] + kc * BK, m0 + p.grp_row0[g], b);
            if (p.resident) continue;
            for (int j = 0; j < p.grp_count[g]; ++j, ++bi) {
              const int sb = bi % B_SLOTS;
              mbar_wait(&b_empty[sb], ((uint32_t)(bi / B_SLOTS) & 1u) ^ 1u);
              mbar_expect_tx(&b_full[sb], (uint32_t)C::B_TILE_BYTES);
              tma_load_3d(b_base + (uint32_t)(sb * C::B_TILE_BYTES), &tmB, &b_full[sb], kc * BK, n0, p.grp_first[g] + j);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (single thread): every tap re-reads the same smem tile at a row offset
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int ai = 0, bi = 0, ti = 0;
      if (p.resident) { mbar_wait(&b_full[0], 0); tcgen05_fence_after(); }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const int buf = ti & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(ti >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        uint32_t first = 1;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int g = 0; g < p.n_groups; ++g, ++ai) {
            const int sa = ai % A_SLOTS;
            mbar_wait(&a_full[sa], (uint32_t)(ai / A_SLOTS) & 1u);
            tcgen05_fence_after();
            const uint32_t a_tile = a_base + (uint32_t)(sa * p.a_slot_bytes);
            for (int j = 0; j < p.grp_count[g]; ++j) {
              const int tap = p.grp_first[g] + j;
              uint32_t b_tile;
              int sb = 0;
              if (p.resident) {
                b_tile = b_base + (uint32_t)((kc * p.g.taps + tap) * C::B_TILE_BYTES);
              } else {
                sb = bi % B_SLOTS;
                mbar_wait(&b_full[sb], (uint32_t)(bi / B_SLOTS) & 1u);
                tcgen05_fence_after();
                b_tile = b_base + (uint32_t)(sb * C::B_TILE_BYTES);
                ++bi;
              }
              const uint64_t da = make_smem_desc(a_tile + (uint32_t)p.tap_byte_off[tap]);
              const uint64_t db = make_smem_desc(b_tile);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 B along K inside the swizzle atom = +2 in the (addr>>4) field
                umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                first = 0;
              }
              if (!p.resident) umma_commit(&b_empty[sb]);   // weight slot is free once these MMAs have read it
            }
            umma_commit(&a_empty[sa]);     // activation tile is free once every tap of the group has read it
          }
        }
        umma_commit(&acc_full[buf]);       // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 2..9): (q, half) = TMEM lane quarter, column half
    // phase 1: TMEM -> registers -> fp32 staging tile in smem (thread = accumulator row), then the accumulator is
    //          handed back to the MMA warp;  phase 2: warps walk the staged tile row-wise: coalesced residual loads
    //          issued ahead of the stores, fused epilogue arithmetic, coalesced fp32 + bf16 stores.
    const int ew = warp - 2;
    const int q = warp & 3, half = ew >> 2;
    const int et = threadIdx.x - 64;
    const Epilogue& e = p.e;
    bf16* out_act = reinterpret_cast<bf16*>(e.out_act);
    const bool has_res = e.res != nullptr, has_res2 = e.res2 != nullptr, has_f32 = e.out_f32 != nullptr, has_act = e.out_act != nullptr;
    const bool use_div = e.div != 1.0f, snake = e.act == ACT_SNAKE;
    const float slope = e.act == ACT_LRELU ? e.slope : (e.act == ACT_RELU ? 0.0f : 1.0f);
    const float alpha = e.alpha;
    constexpr int G = (BN / 4) < 32 ? (BN / 4) : 32;   // lanes per staged row
    constexpr int RPI = 32 / G;                        // rows per warp iteration
    constexpr int ITERS = 16 / RPI;                    // iterations per warp (16 rows each)
    constexpr int U = ITERS < 8 ? ITERS : 8;           // iterations per batch: loads of a batch are issued together
    constexpr int NB = ITERS / U;                      // batches per tile (1 or 2)
    const int sub = lane / G, cl = (lane % G) * 4;
    const float* srow0 = stage + (ew * 16 + sub) * C::STAGE_LD + cl;

    // residual loads of one batch of one tile (issued long before they are consumed; tt[u] < 0 marks an invalid row)
    auto issue = [&](int tile, int it0, float4 (&rr)[U], float4 (&rr2)[U], int (&tt)[U]) {
      const int nt = tile % p.n_tiles, mt = (tile / p.n_tiles) % p.m_tiles, b = tile / (p.n_tiles * p.m_tiles);
      const int n = nt * BN + cl;
      const bool n_ok = n < p.g.N;
      int co = 0, phase = 0;
      if (n_ok) { phase = n / e.phase_cout; co = n - phase * e.phase_cout; }
      const int r_base = mt * BM + ew * 16 + sub;
      const float* res_p = has_res ? e.res + b * e.res_bs + co : nullptr;
      const float* res2_p = has_res2 ? e.res2 + b * e.res2_bs + co : nullptr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r_base + (it0 + u) * RPI;
        int t = e.up_s * r + phase - e.up_p;
        if (!n_ok || r >= p.g.M || t >= e.T_out) t = -1;
        tt[u] = t;
        if (has_res) rr[u] = t >= 0 ? *reinterpret_cast<const float4*>(res_p + (long long)t * e.res_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res2) rr2[u] = t >= 0 ? *reinterpret_cast<const float4*>(res2_p + (long long)t * e.res2_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // fused epilogue arithmetic + coalesced stores of one batch
    auto finish = [&](int tile, int it0, const float4 (&rr)[U], const float4 (&rr2)[U], const int (&tt)[U]) {
      const int nt = tile % p.n_tiles, b = tile / (p.n_tiles * p.m_tiles);
      const int n = nt * BN + cl;
      const bool n_ok = n < p.g.N;
      int co = 0;
      if (n_ok) co = n - (n / e.phase_cout) * e.phase_cout;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sa4 = bias4, sb4 = bias4;
      if (n_ok && e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + co));
      if (n_ok && snake) {
        sa4 = __ldg(reinterpret_cast<const float4*>(e.snake_a + co));
        sb4 = __ldg(reinterpret_cast<const float4*>(e.snake_invb + co));
      }
      const int len_b = e.mask.lens ? __ldg(e.mask.lens + b) : 0x7fffffff;
      float* f32_p = has_f32 ? e.out_f32 + b * e.f32_bs + co : nullptr;
      bf16* act_p = has_act ? out_act + b * e.act_bs + co : nullptr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = tt[u];
        if (t < 0) continue;
        const float4 a4 = *reinterpret_cast<const float4*>(srow0 + (it0 + u) * RPI * C::STAGE_LD);
        float v0 = a4.x + bias4.x, v1 = a4.y + bias4.y, v2 = a4.z + bias4.z, v3 = a4.w + bias4.w;
        const float mv = ((t << e.mask.shift) < len_b) ? 1.0f : 0.0f;
        if (e.mask_pre) { v0 *= mv; v1 *= mv; v2 *= mv; v3 *= mv; }
        v0 *= alpha; v1 *= alpha; v2 *= alpha; v3 *= alpha;
        if (has_res) { v0 += rr[u].x; v1 += rr[u].y; v2 += rr[u].z; v3 += rr[u].w; }
        if (has_res2) { v0 += rr2[u].x; v1 += rr2[u].y; v2 += rr2[u].z; v3 += rr2[u].w; }
        if (use_div) { v0 = v0 / e.div; v1 = v1 / e.div; v2 = v2 / e.div; v3 = v3 / e.div; }
        if (has_f32) *reinterpret_cast<float4*>(f32_p + (long long)t * e.f32_ld) = make_float4(v0, v1, v2, v3);
        if (has_act) {
          float a0, a1, a2, a3;
          if (snake) {   // y + sin^2(y*e^alpha) / (e^beta + 1e-9); fast sine is ample for bf16 operands
            const float s0 = __sinf(v0 * sa4.x), s1 = __sinf(v1 * sa4.y), s2 = __sinf(v2 * sa4.z), s3 = __sinf(v3 * sa4.w);
            a0 = fmaf(sb4.x, s0 * s0, v0); a1 = fmaf(sb4.y, s1 * s1, v1); a2 = fmaf(sb4.z, s2 * s2, v2); a3 = fmaf(sb4.w, s3 * s3, v3);
          } else {       // LeakyReLU(slope) for slope in [0,1]: max(v, v*slope); slope = 1 is the identity
            a0 = fmaxf(v0, v0 * slope); a1 = fmaxf(v1, v1 * slope); a2 = fmaxf(v2, v2 * slope); a3 = fmaxf(v3, v3 * slope);
          }
          if (e.mask_act) { a0 *= mv; a1 *= mv; a2 *= mv; a3 *= mv; }
          __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(act_p + (long long)t * e.act_ld) = pk;
        }
      }
    };

    float4 P[U], P2[U];
    int Pt[U];
    if (p.vec_ok && (int)blockIdx.x < p.total_tiles) issue((int)blockIdx.x, 0, P, P2, Pt);   // first tile's residuals
    int ti = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int nt = tile % p.n_tiles, mt = (tile / p.n_tiles) % p.m_tiles, b = tile / (p.n_tiles * p.m_tiles);
      const int m0 = mt * BM, n0 = nt * BN;
      const int buf = ti & 1;
      mbar_wait(&acc_full[buf], (uint32_t)(ti >> 1) & 1u);
      tcgen05_fence_after();
      {
        float* srow = stage + (q * 32 + lane) * C::STAGE_LD;
        constexpr int CHUNKS = BN / 32;                      // 32-column chunks of the tile
        constexpr int PER = CHUNKS >= 2 ? CHUNKS / 2 : 1;    // chunks per warp (BN=32: only half 0 works)
        if (CHUNKS >= 2 || half == 0) {
#pragma unroll 1
          for (int ci = 0; ci < PER; ++ci) {
            const int c = (CHUNKS >= 2 ? half * PER : 0) + ci;
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), raw);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(srow + c * 32 + j) = make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
          }
        }
      }
      tcgen05_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");           // staging complete, TMEM reads retired
      if (et == 0) mbar_arrive(&acc_empty[buf]);
      if (p.vec_ok) {
        const int next = tile + (int)gridDim.x;
        if (NB == 2) {
          float4 Q[U], Q2[U];
          int Qt[U];
          issue(tile, U, Q, Q2, Qt);         // second half's residuals fly while the first half is finished
          finish(tile, 0, P, P2, Pt);
          if (next < p.total_tiles) issue(next, 0, P, P2, Pt);
          finish(tile, U, Q, Q2, Qt);
        } else {
          finish(tile, 0, P, P2, Pt);
          if (next < p.total_tiles) issue(next, 0, P, P2, Pt);   // next tile's residuals fly during its main loop
        }
      } else {
        // generic scalar path (unaligned strides / channel counts that are not multiples of 4)
        for (int idx = et; idx < BM * BN; idx += EPI_THREADS) {
          const int rl = idx / BN, nl = idx - rl * BN;
          const int r = m0 + rl, n = n0 + nl;
          if (r >= p.g.M || n >= p.g.N) continue;
          int t, co;
          if (!ep_coord(e, r, n, t, co)) continue;
          const float mv = e.mask.at(b, t);
          const float v = ep_value(e, b, t, co, stage[rl * C::STAGE_LD + nl], mv);
          if (e.out_f32) e.out_f32[b * e.f32_bs + (long long)t * e.f32_ld + co] = v;
          if (out_act) out_act[b * e.act_bs + (long long)t * e.act_ld + co] = __float2bfloat16_rn(ep_act(e, co, v, mv));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");           // staging may be overwritten by the next tile
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_halo_mode = -1;   // EV_TC_HALO=0 disables halo reuse (one activation tile per tap) for debugging

bool encode_map(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                uint64_t s2_bytes, uint32_t b0, uint32_t b1, std::string* err, int swizzle_bytes = 128) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u",
               (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
               (unsigned long long)s1_bytes, (unsigned long long)s2_bytes, b0, b1);
      *err = buf;
    }
    return false;
  }
  return true;
}

int g_sm_count = 0;
int g_resident_mode = 1;   // EV_TC_RESIDENT=0 disables the weights-resident variant (debugging)

template <int BN>
cudaError_t launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  // smem: 1 KiB alignment slack + activation ring + weight ring (or all weights) + fp32 staging tile
  const int budget = 200 * 1024 - 1024 - C::STAGING_BYTES;
  p.m_tiles = ceil_div(p.g.M, BM);
  p.n_tiles = ceil_div(p.g.N, BN);
  p.total_tiles = p.m_tiles * p.n_tiles * p.g.B;
  const int w_tiles = p.kchunks * p.g.taps;
  int a_slots, b_slots;
  p.resident = (g_resident_mode != 0) && p.n_tiles == 1 && p.total_tiles >= 2 * g_sm_count &&
               (w_tiles * C::B_TILE_BYTES + 2 * p.a_slot_bytes <= budget);
  if (p.resident) {
    b_slots = w_tiles;
    a_slots = (budget - w_tiles * C::B_TILE_BYTES) / p.a_slot_bytes;
    if (a_slots > MAX_A_SLOTS) a_slots = MAX_A_SLOTS;
  } else {
    a_slots = 3;
    b_slots = (budget - a_slots * p.a_slot_bytes) / C::B_TILE_BYTES;
    if (b_slots < 3) { a_slots = 2; b_slots = (budget - a_slots * p.a_slot_bytes) / C::B_TILE_BYTES; }
    if (b_slots > MAX_B_SLOTS) b_slots = MAX_B_SLOTS;
    if (b_slots < 2) return cudaErrorInvalidConfiguration;
  }
  p.a_slots = a_slots;
  p.b_slots = b_slots;
  const int smem = 1024 + a_slots * p.a_slot_bytes + b_slots * C::B_TILE_BYTES + C::STAGING_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (ce != cudaSuccess) return ce;
    configured = true;
  }
  const int grid = p.total_tiles < g_sm_count ? p.total_tiles : g_sm_count;
  conv_tc_kernel<BN><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace

bool tc_encode_bf16_map(::CUtensorMap_st* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                        uint64_t s2_bytes, uint32_t b0, uint32_t b1, int swizzle_bytes, std::string* err) {
  if (!g_encode && !conv_tc_init(err)) return false;
  return encode_map(map, base, d0, d1, d2, s1_bytes, s2_bytes, b0, b1, err, swizzle_bytes);
}
int tc_sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

int conv_tc_pick_bn(int N) {
  // two CTAs stay resident per SM at BN = 128 (one drains its accumulator while the other feeds the tensor core)
  if (N % 128 == 0) return 128;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return 128;
}

bool conv_tc_init(std::string* err) {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (ce != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
    if (err) *err = std::string("cuTensorMapEncodeTiled entry point unavailable: ") + cudaGetErrorString(ce);
    return false;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (g_sm_count <= 0) g_sm_count = 148;
  return true;
}

// x: channel-last bf16 activations (b, t, c) at x + b*x_bs + t*x_ld + c, `x_rows` addressable rows per item.
cudaError_t conv_tc_launch(const ConvGeom& g, const bf16* x, long long x_ld, long long x_bs, int x_rows,
                           const ConvWeights& w, const Epilogue& e, cudaStream_t stream, std::string* err) {
  if (!g_encode && !conv_tc_init(err)) return cudaErrorNotSupported;
  if ((x_ld & 7) || (x_bs & 7) || (reinterpret_cast<uintptr_t>(x) & 15)) {
    if (err) *err = "conv_tc: activation tensor is not 16-byte aligned / strided";
    return cudaErrorInvalidValue;
  }
  if (g_halo_mode < 0) {
    const char* env = getenv("EV_TC_HALO");
    g_halo_mode = env ? atoi(env) : 1;
    const char* env2 = getenv("EV_TC_RESIDENT");
    g_resident_mode = env2 ? atoi(env2) : 1;
  }
  TcParams p;
  p.g = g;
  p.e = e;
  p.kchunks = ceil_div(g.C_in, BK);
  int tap_row[kMaxTaps], tap_col[kMaxTaps];
  CUtensorMap tmA, tmB;
  uint64_t d0, d1, s1;
  if (g.conv_stride == 1) {
    for (int j = 0; j < g.taps; ++j) { tap_row[j] = g.tap_off[j]; tap_col[j] = 0; }
    d0 = (uint64_t)g.C_in; d1 = (uint64_t)g.T_in; s1 = (uint64_t)x_ld * 2;
  } else if (g.conv_stride == 2) {
    // (T, ld) viewed as (T/2, 2*ld): time 2j+h is row j, columns [h*ld, h*ld + C_in)
    if (g.T_in & 1) { if (err) *err = "conv_tc: stride-2 conv needs an even input length"; return cudaErrorInvalidValue; }
    for (int j = 0; j < g.taps; ++j) {
      const int off = g.tap_off[j];
      const int h = ((off % 2) + 2) % 2;
      tap_row[j] = (off - h) / 2;
      tap_col[j] = h * (int)x_ld;
    }
    d0 = (uint64_t)(x_ld + g.C_in); d1 = (uint64_t)(g.T_in / 2); s1 = (uint64_t)x_ld * 4;
  } else {
    if (err) *err = "conv_tc: unsupported stride";
    return cudaErrorInvalidValue;
  }
  // halo reuse: all taps read one tile when they share the column origin and the row span fits a TMA box
  int lo = tap_row[0], hi = tap_row[0];
  bool same_col = true;
  for (int j = 1; j < g.taps; ++j) { lo = std::min(lo, tap_row[j]); hi = std::max(hi, tap_row[j]); same_col &= tap_col[j] == tap_col[0]; }
  const bool halo = g_halo_mode != 0 && g.taps > 1 && same_col && (BM + hi - lo) <= 256;
  if (halo) {
    p.n_groups = 1;
    p.grp_row0[0] = lo; p.grp_col0[0] = tap_col[0]; p.grp_first[0] = 0; p.grp_count[0] = g.taps;
    for (int j = 0; j < g.taps; ++j) p.tap_byte_off[j] = (tap_row[j] - lo) * BK * 2;
    p.a_rows = (int)align_up(BM + hi - lo, 8);
  } else {
    p.n_groups = g.taps;
    for (int j = 0; j < g.taps; ++j) {
      p.grp_row0[j] = tap_row[j]; p.grp_col0[j] = tap_col[j]; p.grp_first[j] = j; p.grp_count[j] = 1; p.tap_byte_off[j] = 0;
    }
    p.a_rows = BM;
  }
  p.a_slot_bytes = (int)align_up((size_t)p.a_rows * BK * 2, 1024);
  bool ok = encode_map(&tmA, x, d0, d1, (uint64_t)g.B, s1, (uint64_t)x_bs * 2, BK, (uint32_t)p.a_rows, err);
  if (!ok) return cudaErrorInvalidValue;
  (void)x_rows;
  const int BN = conv_tc_pick_bn(g.N);
  if (w.N_pad_tc % BN != 0 || w.K_pad % BK != 0 || w.K_pad < p.kchunks * BK) {
    if (err) *err = "conv_tc: packed weight padding does not match the tile shape";
    return cudaErrorInvalidValue;
  }
  ok = encode_map(&tmB, w.w_bf16, (uint64_t)w.K_pad, (uint64_t)w.N_pad_tc, (uint64_t)w.taps, (uint64_t)w.K_pad * 2,
                  (uint64_t)w.K_pad * w.N_pad_tc * 2, BK, (uint32_t)BN, err);
  if (!ok) return cudaErrorInvalidValue;
  if (e.act != ACT_NONE && e.act != ACT_RELU && e.act != ACT_LRELU && e.act != ACT_SNAKE) {
    if (err) *err = "conv_tc: the tensor-core epilogue implements identity / (leaky) ReLU / SnakeBeta only";
    return cudaErrorInvalidValue;
  }
  auto al4 = [](long long v) { return (v & 3) == 0; };
  p.vec_ok = (e.phase_cout % 4 == 0) && (g.N % 4 == 0) && al4(e.res_ld) && al4(e.res_bs) && al4(e.res2_ld) && al4(e.res2_bs) &&
             al4(e.f32_ld) && al4(e.f32_bs) && al4(e.act_ld) && al4(e.act_bs) &&
             ((reinterpret_cast<uintptr_t>(e.res) | reinterpret_cast<uintptr_t>(e.res2) |
               reinterpret_cast<uintptr_t>(e.out_f32) | reinterpret_cast<uintptr_t>(e.bias)) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(e.out_act) & 7) == 0;
  switch (BN) {
    case 32: return launch_bn<32>(tmA, tmB, p, stream);
    case 64: return launch_bn<64>(tmA, tmB, p, stream);
    default: return launch_bn<128>(tmA, tmB, p, stream);
  }
}

}  // namespace ev
