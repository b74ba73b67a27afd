// Shared definitions of the emojivoice_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && !defined(__CUDA_ARCH_FEAT_SM103_ALL)
#error "emojivoice_b200 kernels are written for sm_100a (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif

namespace ev {

// 3xFP16 split operands (weights.cu, attention_enc_tc.cu): activations are split as fp16 halves of x * 8, so that lo stays a normal
// fp16 number down to |x| ~ 0.016
constexpr float kF16ActScale = 8.0f;

typedef __nv_bfloat16 bf16;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3, ACT_SNAKE = 4, ACT_SILU = 5, ACT_MISH = 6 };

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Row validity ("mask") is never materialised: rows of batch item b are valid while (t << shift) < lens[b]
// (sequence_mask, utils/model.py:7-11; the decoder's half-rate mask is mask[:, :, ::2], decoder.py:407).
struct RowMask {
  const int* lens;  // [B] or nullptr (= all rows valid)
  int shift;
  __device__ __forceinline__ float at(int b, int t) const {
    return (lens == nullptr || (t << shift) < lens[b]) ? 1.0f : 0.0f;
  }
};

__device__ __forceinline__ float mish_f(float x) {
  // torch.nn.functional.mish = x * tanh(softplus(x)), softplus threshold 20 (decoder.py:38)
  float sp = x > 20.0f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope, float snake_a, float snake_invb) {
  switch (act) {
    case ACT_RELU: return v > 0.0f ? v : 0.0f;
    case ACT_LRELU: return v > 0.0f ? v : v * slope;
    case ACT_TANH: return tanhf(v);
    case ACT_SNAKE: { float s = sinf(v * snake_a); return v + snake_invb * (s * s); }  // transformer.py:78
    case ACT_SILU: return silu_f(v);
    case ACT_MISH: return mish_f(v);
    default: return v;
  }
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

// Programmatic dependent launch (PDL): a kernel launched with the attribute may begin while its predecessor on the
// stream is still draining; pdl_wait() blocks until that predecessor has completed and its writes are visible, so it
// must precede the kernel's first global-memory access.  pdl_trigger() lets the successor start its own prologue.
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}


// Host: one-time per-DEVICE setup (cudaFuncSetAttribute is a per-device property; a process may hold contexts on several
// devices, and several host threads may race here: the setter runs until it has succeeded once for the current device).
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <typename F> cudaError_t run(F&& setter) {
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    ce = setter();
    if (ce == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return ce;
  }
};

// Host: launch `kernel` allowing programmatic stream serialization (only for kernels that call pdl_wait()).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ev
