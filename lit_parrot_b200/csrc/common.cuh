// Shared device/host helpers for liblitparrot_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "lp_abi.h"

namespace lp {

// ---------------------------------------------------------------------------------------------
// host side: error plumbing + launch helper with Programmatic Dependent Launch (PDL)
// ---------------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* what);

// One-time set-up PER DEVICE (cudaFuncSetAttribute is per device; a process may drive several GPUs): `mask` holds one bit per
// device ordinal.  Callers test, configure, then mark — two threads racing configure twice, which is harmless.
inline bool needs_device_setup(const std::atomic<unsigned long long>& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  return !(mask.load(std::memory_order_acquire) & (1ull << (dev & 63)));
}
inline void mark_device_setup(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
}
bool pdl_enabled();
int num_sms();
void count_launch();

#define LP_CUDA_TRY(expr)                                 \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) {                              \
      ::lp::set_cuda_error(_e, #expr);                    \
      return LP_ERR_CUDA;                                 \
    }                                                     \
  } while (0)

// Every kernel of this library executes griddepcontrol.wait before it touches activations, so a chain
// of PDL launches keeps full stream-order semantics while the next kernel's prologue (weight prefetch)
// overlaps the previous kernel's tail.
template <typename... KArgs, typename... Args>
inline int launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  LP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  count_launch();
  return LP_OK;
}

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float maybe_round(float v, int round_bf16) { return round_bf16 ? bf16_round(v) : v; }

// adapter-v2 output affine (lit_gpt/adapter_v2.py:34-35): adapter_scale * (linear(x) + adapter_bias), `y` = linear(x) incl. its own
// bias (already rounded in bf16 mode); each step rounded where the reference's bf16-true run rounds
__device__ __forceinline__ float out_affine(float y, const float* __restrict__ ob, const float* __restrict__ os, int n, int round_bf16) {
  if (ob) y = maybe_round(y + ob[n], round_bf16);
  if (os) y = maybe_round(os[n] * y, round_bf16);
  return y;
}

// 2^x as ONE MUFU instruction (ex2.approx.ftz: results below 2^-126 flush to zero, 2 ulp).  exp2f() wraps the same instruction in a
// range fix-up (compare, halve, square: 3 more instructions per element) that a softmax probability does not need.
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit load that does not allocate in L1 (weights are read exactly once)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float v) { return v / (1.0f + expf(-v)); }

}  // namespace lp
