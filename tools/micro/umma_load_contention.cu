// Does TMA traffic into shared memory slow SS-mode tcgen05.mma down?  Thread 0 issues back-to-back MMAs (M 128, N 256, K 16) on fixed
// operand tiles while thread 32 streams `KB_PER_MMA_x4` KB of bulk copies per 4 MMAs (the GEMM's ratio is 48 KB per 4 MMAs) into a ring.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(long long* out, const unsigned char* src, size_t src_bytes, int iters, int load_kb) {
  extern __shared__ __align__(1024) unsigned char smem[];  // [48 KB operands][3 x 48 KB ring]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bars[8];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 128;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int kk = i & 3;
      const uint64_t da = desc(a0 + kk * 32), db = desc(b0 + kk * 32);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da),
                   "l"(db), "r"(idesc), "r"(i ? 1 : 0) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bars[7])) : "memory");
    mbar_wait(smem_u32(&bars[7]), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x == 32 && load_kb > 0) {
    // stream: per 4 MMAs (512 cycles) load_kb KB -> total bytes = iters / 4 * load_kb KB, 4 copies in flight
    const long long total = (long long)(iters / 4) * load_kb * 1024;
    const uint32_t chunk = (uint32_t)load_kb * 1024;
    const size_t span = (src_bytes / gridDim.x) & ~(size_t)65535;
    const unsigned char* g = src + (size_t)blockIdx.x * span;
    long long done = 0;
    int s = 0;
    int uses[3] = {0, 0, 0};
    size_t off = 0;
    while (done < total) {
      if (uses[s] > 0) mbar_wait(smem_u32(&bars[s]), (uses[s] - 1) & 1);  // the previous copy into this slot has landed
      const uint32_t dst = smem_u32(smem) + 49152 + s * 49152;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(g + off), "r"(chunk),
                   "r"(smem_u32(&bars[s])) : "memory");
      ++uses[s];
      off += chunk;
      if (off + chunk > span) off = 0;
      done += chunk;
      if (++s == 3) s = 0;
    }
    for (int i = 0; i < 3; ++i)
      if (uses[i] > 0) mbar_wait(smem_u32(&bars[i]), (uses[i] - 1) & 1);
    if (blockIdx.x == 0) out[1] = clock64();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  unsigned char* src;
  const size_t src_bytes = (size_t)1 << 30;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 1024);
  const int iters = 16384;
  for (int kb : {0, 16, 32, 48}) {
    k<<<148, 128, 4 * 49152 + 1024>>>(d, src, src_bytes, iters, kb);
    cudaDeviceSynchronize();
    k<<<148, 128, 4 * 49152 + 1024>>>(d, src, src_bytes, iters, kb);
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("loads %2d KB per 4 MMAs (all 148 SMs): %.1f cycles per MMA  err=%s\n", kb, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
