// How fast can one SM pull L2-resident data into shared memory with bulk copies while all 148 SMs do the same?  (The prefill GEMM
// needs 48 KB per k-block of 512 MMA cycles = 94 B/clk/SM with single-CTA 128x256 tiles, 64 B/clk/SM with 2-CTA W multicast.)
// One loader thread per CTA keeps `depth` copies of `chunk` bytes in flight from a window of `span` bytes (per SM, or shared by all
// SMs); optionally thread 0 issues back-to-back tcgen05.mma on fixed operand tiles at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingress tma_ingress.cu && ./tma_ingress
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

constexpr int RING = 160 * 1024;

__global__ void __launch_bounds__(128, 1) k(long long* out, const unsigned char* src, size_t span, int shared_src, long long total,
                                            uint32_t chunk, int depth, int mma_iters) {
  extern __shared__ __align__(1024) unsigned char smem[];  // [48 KB operands][ring]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bars[16];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])) : "memory");
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
  if (threadIdx.x == 0 && mma_iters > 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 128;
    const long long t0 = clock64();
    for (int i = 0; i < mma_iters; ++i) {
      const int kk = i & 3;
      const uint64_t da = desc(a0 + kk * 32), db = desc(b0 + kk * 32);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da),
                   "l"(db), "r"(idesc), "r"(i ? 1 : 0) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bars[15])) : "memory");
    mbar_wait(smem_u32(&bars[15]), 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    const unsigned char* g = src + (shared_src ? 0 : (size_t)blockIdx.x * span);
    long long done = 0;
    int s = 0;
    int uses[12] = {0};
    size_t off = shared_src ? ((size_t)blockIdx.x * chunk) % span : 0;  // SMs start at different places of a shared window
    const long long t0 = clock64();
    while (done < total) {
      if (uses[s] > 0) mbar_wait(smem_u32(&bars[s]), (uses[s] - 1) & 1);
      const uint32_t dst = smem_u32(smem) + 49152 + s * chunk;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(g + off), "r"(chunk),
                   "r"(smem_u32(&bars[s])) : "memory");
      ++uses[s];
      off += chunk;
      if (off + chunk > span) off = 0;
      done += chunk;
      if (++s == depth) s = 0;
    }
    for (int i = 0; i < depth; ++i)
      if (uses[i] > 0) mbar_wait(smem_u32(&bars[i]), (uses[i] - 1) & 1);
    const long long t1 = clock64();
    out[2 + blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 160);
  unsigned char* src;
  const size_t src_bytes = (size_t)2 << 30;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  const int smem = 49152 + RING + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const long long total = 64ll << 20;  // per SM
  struct Case { const char* name; size_t span; int shared; uint32_t chunk; int depth; int mma; };
  const Case cases[] = {
      {"L2 window 192 KB / SM (28 MB), 16 KB x 8", 192 << 10, 0, 16384, 8, 0},
      {"L2 window 192 KB / SM (28 MB), 32 KB x 4", 192 << 10, 0, 32768, 4, 0},
      {"L2 window 192 KB / SM, 16 KB x 8 + MMAs", 192 << 10, 0, 16384, 8, 1},
      {"shared 8 MB window (all SMs), 16 KB x 8", 8 << 20, 1, 16384, 8, 0},
      {"shared 8 MB window (all SMs), 32 KB x 4", 8 << 20, 1, 32768, 4, 0},
      {"shared 8 MB window, 16 KB x 8 + MMAs", 8 << 20, 1, 16384, 8, 1},
      {"shared 48 MB window (all SMs), 16 KB x 8", 48 << 20, 1, 16384, 8, 0},
      {"DRAM stream 13.8 MB / SM (2 GB), 16 KB x 8", (size_t)13 << 20, 0, 16384, 8, 0},
  };
  for (const Case& c : cases) {
    const int mma_iters = c.mma ? 1 << 17 : 0;
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, smem>>>(d, src, c.span, c.shared, total, c.chunk, c.depth, mma_iters);
    cudaDeviceSynchronize();
    long long h[160] = {0};
    cudaMemcpy(h, d, 8 * 150, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 62;
    for (int i = 0; i < 148; ++i) { mx = h[2 + i] > mx ? h[2 + i] : mx; mn = h[2 + i] < mn ? h[2 + i] : mn; }
    printf("%-48s: %.1f .. %.1f B/clk/SM", c.name, (double)total / mx, (double)total / mn);
    if (c.mma) printf("   (%.1f cycles per MMA)", (double)h[0] / mma_iters);
    printf("  err=%s\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
