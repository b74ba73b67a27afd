// bf16 implicit-GEMM conv1d on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.  sm_100a only.
//
// Tile: 128 output time steps (UMMA M = 128 = TMEM lanes) x BN output channels (UMMA N = BN TMEM columns),
// reduced over taps x ceil(C_in/64) stages of K = 64 bf16 channels (4 x UMMA K=16 per stage).
//   A (activations) : 3-D tensor map (channel, time, batch) over the channel-last bf16 tensor; the box
//                     [64 ch x 128 rows] for tap j is fetched at row  m0 + tap_row[j]  -- rows outside [0,T) are
//                     zero-filled by TMA, which IS the convolution's zero padding (no im2col, no halo copies);
//                     the stride-2 conv reads a (T/2, 2*ld) view of the same memory (tap_col selects even/odd).
//   B (weights)     : 3-D tensor map (c_in, n, tap) over [taps][N_pad][K_pad] bf16, box [64 x BN].
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld 32x32b, fused bias / mask / Euler-or-residual / activation, vector stores).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "conv.cuh"

namespace ev {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB

struct TcParams {
  ConvGeom g;
  Epilogue e;
  int tap_row[kMaxTaps];
  int tap_col[kMaxTaps];
  int kchunks;
  int vec_ok;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
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
    if (clock64() - t0 > 8000000000LL) {  // ~4 s: a pipeline bug must not hang the device
      printf("conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major), canonical value 1
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : 3;  // 1 CTA/SM at BN=256; 2-3 co-resident CTAs for the narrow tiles
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack for the 1024-B alignment
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int n_iters = p.g.taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], (uint32_t)C::STAGE_BYTES);
        const uint32_t sa = tiles + (uint32_t)s * C::STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full_bar[s], p.tap_col[tap] + kc * BK, m0 + p.tap_row[tap], b);
        tma_load_3d(sa + A_TILE_BYTES, &tmB, &full_bar[s], kc * BK, n0, tap);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = tiles + (uint32_t)s * C::STAGE_BYTES;
        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)  // +32 B along K inside the swizzle atom = +2 in the (addr>>4) field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((it | k) != 0));
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(&accum_bar);       // accumulator complete
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: warp q = warp%4 owns TMEM lanes [32q, 32q+32) = GEMM rows m0+32q+lane
    const int q = warp & 3;
    const int r = m0 + q * 32 + lane;
    const Epilogue& e = p.e;
    mbar_wait(&accum_bar, 0);
    tcgen05_fence_after();
    bf16* out_act = reinterpret_cast<bf16*>(e.out_act);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.g.N) break;  // warp-uniform
      __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge lanes that skipped the previous chunk
      uint32_t raw[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
      if (r >= p.g.M) continue;
      const int phase = nb / e.phase_cout;
      const int co0 = nb - phase * e.phase_cout;
      const bool one_phase = (co0 + 32 <= e.phase_cout) && (nb + 32 <= p.g.N);
      if (one_phase && p.vec_ok) {
        const int t = e.up_s * r + phase - e.up_p;
        if (t < 0 || t >= e.T_out) continue;
        const float mv = e.mask.at(b, t);
        const float* bias = e.bias ? e.bias + co0 : nullptr;
        const float* res = e.res ? e.res + b * e.res_bs + (long long)t * e.res_ld + co0 : nullptr;
        const float* res2 = e.res2 ? e.res2 + b * e.res2_bs + (long long)t * e.res2_ld + co0 : nullptr;
        float* of = e.out_f32 ? e.out_f32 + b * e.f32_bs + (long long)t * e.f32_ld + co0 : nullptr;
        bf16* oa = out_act ? out_act + b * e.act_bs + (long long)t * e.act_ld + co0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float v[4] = {__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3])};
          if (bias) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(bias + j)); v[0] += t4.x; v[1] += t4.y; v[2] += t4.z; v[3] += t4.w; }
          if (e.mask_pre) { v[0] *= mv; v[1] *= mv; v[2] *= mv; v[3] *= mv; }
          v[0] *= e.alpha; v[1] *= e.alpha; v[2] *= e.alpha; v[3] *= e.alpha;
          if (res) { const float4 t4 = *reinterpret_cast<const float4*>(res + j); v[0] += t4.x; v[1] += t4.y; v[2] += t4.z; v[3] += t4.w; }
          if (res2) { const float4 t4 = *reinterpret_cast<const float4*>(res2 + j); v[0] += t4.x; v[1] += t4.y; v[2] += t4.z; v[3] += t4.w; }
          if (e.div != 1.0f) { v[0] = v[0] / e.div; v[1] = v[1] / e.div; v[2] = v[2] / e.div; v[3] = v[3] / e.div; }
          if (of) *reinterpret_cast<float4*>(of + j) = make_float4(v[0], v[1], v[2], v[3]);
          if (oa) {
            const float a0 = ep_act(e, co0 + j, v[0], mv), a1 = ep_act(e, co0 + j + 1, v[1], mv);
            const float a2 = ep_act(e, co0 + j + 2, v[2], mv), a3 = ep_act(e, co0 + j + 3, v[3], mv);
            __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(oa + j) = pk;
          }
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
          const int n = nb + j;
          if (n >= p.g.N) break;
          int t, co;
          if (!ep_coord(e, r, n, t, co)) continue;
          const float mv = e.mask.at(b, t);
          // dynamic register-array index: keep it simple, this path only serves ragged tails (e.g. N = 80)
          float accv = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) accv = (jj == j) ? __uint_as_float(raw[jj]) : accv;
          const float v = ep_value(e, b, t, co, accv, mv);
          if (e.out_f32) e.out_f32[b * e.f32_bs + (long long)t * e.f32_ld + co] = v;
          if (out_act) out_act[b * e.act_bs + (long long)t * e.act_ld + co] = __float2bfloat16_rn(ep_act(e, co, v, mv));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

bool encode_map(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                uint64_t s2_bytes, uint32_t b0, uint32_t b1, std::string* err) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

template <int BN>
cudaError_t launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (ce != cudaSuccess) return ce;
    configured = true;
  }
  dim3 grid(ceil_div(p.g.M, BM), ceil_div(p.g.N, BN), p.g.B);
  conv_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

}  // namespace

int conv_tc_pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
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
  TcParams p;
  p.g = g;
  p.e = e;
  p.kchunks = ceil_div(g.C_in, BK);
  CUtensorMap tmA, tmB;
  bool ok;
  if (g.conv_stride == 1) {
    for (int j = 0; j < g.taps; ++j) { p.tap_row[j] = g.tap_off[j]; p.tap_col[j] = 0; }
    ok = encode_map(&tmA, x, (uint64_t)g.C_in, (uint64_t)g.T_in, (uint64_t)g.B, (uint64_t)x_ld * 2, (uint64_t)x_bs * 2,
                    BK, BM, err);
  } else if (g.conv_stride == 2) {
    // (T, ld) viewed as (T/2, 2*ld): time 2j+h is row j, columns [h*ld, h*ld + C_in)
    if (g.T_in & 1) { if (err) *err = "conv_tc: stride-2 conv needs an even input length"; return cudaErrorInvalidValue; }
    for (int j = 0; j < g.taps; ++j) {
      const int off = g.tap_off[j];
      const int h = ((off % 2) + 2) % 2;
      p.tap_row[j] = (off - h) / 2;
      p.tap_col[j] = h * (int)x_ld;
    }
    ok = encode_map(&tmA, x, (uint64_t)(x_ld + g.C_in), (uint64_t)(g.T_in / 2), (uint64_t)g.B, (uint64_t)x_ld * 4,
                    (uint64_t)x_bs * 2, BK, BM, err);
  } else {
    if (err) *err = "conv_tc: unsupported stride";
    return cudaErrorInvalidValue;
  }
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
  auto al4 = [](long long v) { return (v & 3) == 0; };
  p.vec_ok = (e.phase_cout % 32 == 0) && al4(e.res_ld) && al4(e.res_bs) && al4(e.res2_ld) && al4(e.res2_bs) &&
             al4(e.f32_ld) && al4(e.f32_bs) && al4(e.act_ld) && al4(e.act_bs) &&
             ((reinterpret_cast<uintptr_t>(e.res) | reinterpret_cast<uintptr_t>(e.res2) |
               reinterpret_cast<uintptr_t>(e.out_f32) | reinterpret_cast<uintptr_t>(e.bias)) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(e.out_act) & 7) == 0;
  switch (BN) {
    case 32: return launch_bn<32>(tmA, tmB, p, stream);
    case 64: return launch_bn<64>(tmA, tmB, p, stream);
    case 128: return launch_bn<128>(tmA, tmB, p, stream);
    default: return launch_bn<256>(tmA, tmB, p, stream);
  }
}

}  // namespace ev
