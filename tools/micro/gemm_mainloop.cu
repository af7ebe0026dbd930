// The prefill GEMM's main loop alone (no epilogue): TMA producer + tcgen05.mma consumer over a ring, accumulating forever.
// Derived from tma_tile_ingress.cu (load pattern: per k-block one {64 x 128 rows} X box (16 KB) and
// one {64 x 256 rows} W box (32 KB), 128-byte swizzle, `depth` k-blocks in flight, CTA tiles walking an [M, K] x [N, K] problem
// that is L2 resident.  Reports bytes per clock per SM; the GEMM needs 94 B/clk/SM for the tensor pipe to run at full rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_mainloop gemm_mainloop.cu && ./gemm_mainloop
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst), "l"(m),
               "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
               "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// mode 0: 2-D boxes, one k-block (64 columns) per stage: X 16 KB + W 32 KB.  mode 1: 3-D boxes with 2 k-blocks per copy (the W box is
// split into two 128-row copies to stay within 256 x ...): same bytes per stage pair, half the copies per byte for X.
__global__ void __launch_bounds__(128, 1) k(long long* out, const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mw,
                                            int M, int N, int K, int depth, int rows_w, int tiles_per_cta, int do_mma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[8], ebars[8], fin;
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&ebars[i])) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&fin)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 64 && do_mma) {
    const int bn = rows_w > 0 ? rows_w : 128;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t stage_bytes = 16384 + (uint32_t)rows_w * 128;
    int s = 0, ph = 0;
    const int total_kb = tiles_per_cta * (K / 64);
    const long long t0 = clock64();
    for (int i = 0; i < total_kb; ++i) {
      mbar_wait(smem_u32(&bars[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t a0 = smem_u32(smem) + s * stage_bytes, b0 = rows_w > 0 ? a0 + 16384 : a0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = desc(a0 + kk * 32), db = desc(b0 + kk * 32);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da),
                     "l"(db), "r"(idesc), "r"(i | kk ? 1 : 0) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&ebars[s])) : "memory");
      if (++s == depth) { s = 0; ph ^= 1; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&fin)) : "memory");
    mbar_wait(smem_u32(&fin), 0);
    out[2 * blockIdx.x] = clock64() - t0;
    out[2 * blockIdx.x + 1] = (long long)total_kb * stage_bytes;
  }
  if (threadIdx.x == 32) {
    const uint32_t stage_bytes = 16384 + (uint32_t)rows_w * 128;
    const int mt = M / 128, nt = N / 256;
    int uses[8] = {0};
    int s = 0;
    long long bytes = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int tile = (blockIdx.x + t * gridDim.x) % (mt * nt);
      const int m0 = (tile % mt) * 128, n0 = (tile / mt) * 256;
      for (int kb = 0; kb < K / 64; ++kb) {
        if (do_mma) { if (uses[s] > 0) mbar_wait(smem_u32(&ebars[s]), (uses[s] - 1) & 1); }
        else if (uses[s] > 0) mbar_wait(smem_u32(&bars[s]), (uses[s] - 1) & 1);
        const uint32_t dst = smem_u32(smem) + s * stage_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bars[s])), "r"(stage_bytes) : "memory");
        tma2d(dst, &mx, kb * 64, m0, smem_u32(&bars[s]));
        if (rows_w > 0) tma2d(dst + 16384, &mw, kb * 64, n0, smem_u32(&bars[s]));
        ++uses[s];
        bytes += stage_bytes;
        if (++s == depth) s = 0;
      }
    }
    if (!do_mma) {
      for (int i = 0; i < depth; ++i)
        if (uses[i] > 0) mbar_wait(smem_u32(&bars[i]), (uses[i] - 1) & 1);
      const long long t1 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = bytes;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}

__global__ void fill_random(unsigned short* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // bf16 in roughly N(0, 1)-like range: random sign and mantissa, exponent 120..127
    p[i] = (unsigned short)((h & 0x807fu) | ((120u + ((h >> 16) & 7u)) << 7));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(f);
  const int M = argc > 1 ? atoi(argv[1]) : 2048, N = argc > 2 ? atoi(argv[2]) : 4096, K = argc > 3 ? atoi(argv[3]) : 4096;  // default X 16 MB + W 32 MB: L2 resident
  const int tpc = (M / 128) * (N / 256) / 148 + 1;
  printf("M %d N %d K %d: X %.0f MB, W %.0f MB, %d tiles per CTA\n", M, N, K, M * (double)K * 2e-6, N * (double)K * 2e-6, tpc);
  unsigned short *x, *w;
  cudaMalloc(&x, (size_t)M * K * 2);
  cudaMalloc(&w, (size_t)N * K * 2);
  const bool rnd = argc > 4 && atoi(argv[4]) != 0;  // operand VALUES matter: tensor-core power depends on the data
  if (rnd) {
    fill_random<<<1024, 256>>>(x, (size_t)M * K, 1u);
    fill_random<<<1024, 256>>>(w, (size_t)N * K, 2u);
  } else {
    cudaMemset(x, 0, (size_t)M * K * 2);
    cudaMemset(w, 0, (size_t)N * K * 2);
  }
  printf("operands: %s\n", rnd ? "random bf16" : "zeros");
  long long* d;
  cudaMalloc(&d, 16 * 148);
  auto make = [&](void* p, int rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    return m;
  };
  const CUtensorMap mx = make(x, M, 128);
  const int smem = 4 * 49152 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { const char* name; int rows_w, depth, mma; };
  const Case cases[] = {{"loads only, N 256, 4 stages", 256, 4, 0},        {"loads + MMA, N 256, 4 stages", 256, 4, 1},
                        {"loads + MMA, N 256, 3 stages", 256, 3, 1},       {"loads + MMA, N 256, 2 stages", 256, 2, 1},
                        {"loads + MMA, N 128, 4 stages", 128, 4, 1},       {"loads + MMA, N 128, 6 stages", 128, 6, 1}};
  for (const Case& c : cases) {
    const CUtensorMap mw = make(w, N, c.rows_w > 0 ? c.rows_w : 128);
    for (int rep = 0; rep < (argc > 5 ? atoi(argv[5]) : 2); ++rep) k<<<148, 128, smem>>>(d, mx, mw, M, N, K, c.depth, c.rows_w, tpc, c.mma);
    cudaDeviceSynchronize();
    long long h[2 * 148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mn = 1e30, mxr = 0;
    for (int i = 0; i < 148; ++i) {
      const double r = (double)h[2 * i + 1] / (double)h[2 * i];
      mn = r < mn ? r : mn;
      mxr = r > mxr ? r : mxr;
    }
    printf("%-32s: %.1f .. %.1f B/clk/SM  (%.0f cycles per k-block; MMA floor %d)  err=%s\n", c.name, mn, mxr, (16384.0 + c.rows_w * 128) / mn, c.rows_w * 2,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
