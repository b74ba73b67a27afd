// Micro-benchmark: issue rate of tcgen05.mma.kind::f16 (M = 128, K = 16) on sm_100a as a function of N, of the
// shared-memory swizzle mode of the operands, of the number of accumulators rotated, of the row alignment of the A
// start address (conv taps are row offsets into a haloed tile) and of the number of issuing warps.
// Operands are zero-filled shared memory; only timing matters.  Diagnostic tool (not part of the library):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I emojivoice_b200/csrc scripts/mma_rate.cu -o scripts/_bin/mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

using namespace ev::tc;

struct Cfg { int N, layout, n_acc, n_iter, row_step, n_issuers, ks_per_tap, n_spin, tf32, fill, commit_each, b_tiles; };

__global__ void __launch_bounds__(640, 1) mma_rate_kernel(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done[4];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] =
        c.fill ? make_uint4(0x3f8ccccdu + 7919u * i, 0xbf19999au + 104729u * i, 0x3e4ccccdu + 31u * i, 0xbdcccccdu + 17u * i) : make_uint4(0, 0, 0, 0);   // |x| ~ 0.1 .. 2
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&done[i], i == 2 ? 1 << 20 : 1);     // done[2]: a sink for per-tap commits
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const int RB = c.layout == 2 ? 128 : (c.layout == 4 ? 64 : 32);
  const uint32_t hi = ((uint32_t)(8 * RB) >> 4) | (1u << 14) | ((uint32_t)c.layout << 29);
  const uint32_t a_lo0 = ((base & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo0 = (((base + 96 * 1024) & 0x3FFFFu) >> 4) | (1u << 16);
  long long t0 = 0, t1 = 0;
  if (warp < c.n_issuers) {
    const uint32_t idesc = c.tf32 ? make_idesc_tf32(128, c.N) : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
    __syncwarp();
    t0 = clock64();
    if (elect_one()) {
      // same shape as the fused ResBlock issuer: per tap, MB m-blocks x KS k-steps, straight-line code, descriptors by adds
      uint32_t a_lo = a_lo0;
      const uint32_t a_step = (uint32_t)((c.row_step * RB) >> 4);
      const uint32_t d0 = tmem + (uint32_t)(warp * c.n_acc * c.N);
      int wrap = 0;
      for (int it = 0; it < c.n_iter; ++it) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          if (m >= c.n_acc) break;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks >= c.ks_per_tap) break;
            const uint64_t da = ((uint64_t)hi << 32) | (a_lo + (uint32_t)((m * 128 * RB) >> 4) + 2u * ks);
            const uint64_t db = ((uint64_t)hi << 32) | (b_lo0 + (uint32_t)((it % c.b_tiles) * ((c.N * RB) >> 4)) + 2u * ks);   // b_tiles > 1: a new weight tile per tap
            if (c.tf32) umma_tf32(d0 + (uint32_t)(m * c.N), da, db, idesc, 1u);      // kind::tf32: K = 8 fp32 elements = the same 32 bytes per row
            else umma_bf16(d0 + (uint32_t)(m * c.N), da, db, idesc, 1u);
          }
        }
        if (c.commit_each) umma_commit(&done[2]);      // like the conv kernels: one commit per (tap, weight tile)
        a_lo += a_step;
        if (++wrap == 25) { wrap = 0; a_lo = a_lo0; }
      }
      umma_commit(&done[warp]);
      if (warp == 0) umma_commit(&done[3]);
    }
    __syncwarp();
    mbar_wait(&done[warp], 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  else if (warp >= 2 && warp < 2 + c.n_spin) {
    mbar_wait(&done[3], 0);          // like epilogue warps waiting for an MMA phase: do they slow the issuers down?
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 4 * grid);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<long long> h(4 * grid);
  printf("grid %d CTAs, 128x Nx16 bf16 MMAs, clk per MMA (max over CTAs / issuers)\n", grid);
  printf("%5s %7s %6s %9s %9s %8s | %10s %12s\n", "N", "swizzle", "mblk", "row_step", "issuers", "ks/tap", "clk/MMA", "floor(N/2)");
  printf("(mblk = m-blocks per tap, each with ks k-steps; spin = warps waiting on an mbarrier meanwhile)\n");
  if (argc > 2 && atoi(argv[2]) == 1) {       // kind::tf32 (the text encoder's 3xTF32 path): N = 128, 128B swizzle
    printf("%s, 128 x 128 MMAs, %d weight tile(s) cycled:\n", (argc > 5 ? atoi(argv[5]) : 1) ? "kind::tf32 (K = 8)" : "kind::f16 (K = 16)", argc > 6 ? atoi(argv[6]) : 1);
    for (int issuers = 1; issuers <= 2; ++issuers)
      for (int n_acc = 1; n_acc <= 2; ++n_acc) {
        Cfg c{128, 2, n_acc, 4096 / (4 * n_acc), 1, issuers, 4, 0, argc > 5 ? atoi(argv[5]) : 1, argc > 3 ? atoi(argv[3]) : 0, argc > 4 ? atoi(argv[4]) : 0, argc > 6 ? atoi(argv[6]) : 1};
        cudaMemset(out, 0, sizeof(long long) * 4 * grid);
        mma_rate_kernel<<<grid, 640, 200 * 1024>>>(c, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
        cudaMemcpy(h.data(), out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (long long v : h) mx = v > mx ? v : mx;
        printf("  issuers %d m-blocks %d: %.1f clk per MMA\n", issuers, n_acc, (double)mx / (double)(c.n_iter * 4 * n_acc * issuers));
      }
    return 0;
  }
  const int Ns[] = {32, 64, 128};
  for (int N : Ns)
    for (int layout : {2, 4})
      for (int half = 0; half < (layout == 2 ? 2 : 1); ++half)
        for (int issuers = 1; issuers <= 2; ++issuers)
          for (int n_acc = 1; n_acc <= 2; ++n_acc)
            for (int row_step : {0, 1, 5})
              for (int n_spin : {0, 16}) {
                if (issuers * n_acc * N > 512) continue;
                const int ks = layout == 2 ? (half ? 2 : 4) : 2;
                Cfg c{N, layout, n_acc, 4096 / (ks * n_acc), row_step, issuers, ks, n_spin, 0, 0, 0, 1};
                cudaMemset(out, 0, sizeof(long long) * 4 * grid);
                mma_rate_kernel<<<grid, 640, 200 * 1024>>>(c, out);
                cudaError_t ce = cudaDeviceSynchronize();
                if (ce != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(ce)); return 1; }
                cudaMemcpy(h.data(), out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (long long v : h) mx = v > mx ? v : mx;
                const double per = (double)mx / (double)(c.n_iter * ks * n_acc * issuers);
                printf("%5d %7s %6d %9d %9d %8d spin %2d | %10.1f %12.1f\n", N, layout == 2 ? "128B" : "64B", n_acc, row_step, issuers, ks, n_spin, per,
                       N / 2.0);
              }
  return 0;
}
