// Does the HBM layout of a weight stage matter?  All 148 SMs stream a [N, K] bf16 matrix (>> L2) in 16 KB stages, 10 in flight per SM:
//   mode 0: the step kernel's pattern — one cp.async.bulk.tensor.3d per stage, box {128 B, 16 rows, 8 K-blocks} of the row-major
//           matrix: 16 separate 1 KB runs, one per row, 2 K bytes apart (K = 4096: 8 KB);
//   mode 1: stage-major storage — the same 16 KB as ONE contiguous run, one cp.async.bulk per stage;
//   mode 2: as 0 with 32 KB stages (16 rows x 16 K-blocks: 2 KB runs), 5 in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_layout stream_layout.cu && ./stream_layout
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap map, const unsigned char* w, int N, int K, int mode, int depth,
                                           uint32_t stage_bytes, int rows) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int kb_per_stage = (int)(stage_bytes / (128 * rows));  // K-blocks of 128 B x `rows` rows
  const int nks = (K * 2 / 128) / kb_per_stage;              // stages per 16-row tile
  const long long ntiles = N / rows, total = ntiles * nks;
  const long long s0 = total * blockIdx.x / gridDim.x, s1 = total * (blockIdx.x + 1) / gridDim.x;  // contiguous ranges, like the step kernel
  int uses[16] = {0};
  int s = 0;
  for (long long u = s0; u < s1; ++u) {
    if (uses[s] > 0) mbar_wait(smem_u32(&bars[s]), (uses[s] - 1) & 1);
    const uint32_t dst = smem_u32(smem) + s * stage_bytes, bar = smem_u32(&bars[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(stage_bytes) : "memory");
    if (mode == 1) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                   "l"(w + (size_t)u * stage_bytes), "r"(stage_bytes), "r"(bar) : "memory");
    } else {
      const int tile = (int)(u / nks), ks = (int)(u % nks);
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
                   "l"(&map), "r"(0), "r"(tile * rows), "r"(ks * kb_per_stage), "r"(bar) : "memory");
    }
    ++uses[s];
    if (++s == depth) s = 0;
  }
  for (int i = 0; i < depth; ++i)
    if (uses[i] > 0) mbar_wait(smem_u32(&bars[i]), (uses[i] - 1) & 1);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(f);
  const int K = argc > 1 ? atoi(argv[1]) : 4096, N = argc > 2 ? atoi(argv[2]) : 131072;  // 1 GiB of bf16 at K = 4096
  unsigned char* w;
  cudaMalloc(&w, (size_t)N * K * 2);
  cudaMemset(w, 1, (size_t)N * K * 2);
  const int smem = 200 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Case { const char* name; int mode, depth; uint32_t stage; int l2; int rows = 16; int ctas = 148; };
  const Case cases[] = {{"3-D tensor box, 16 rows x 1 KB runs, 16 KB x 10", 0, 10, 16384, 1}, {"contiguous 16 KB x 10", 1, 10, 16384, 1},
                        {"3-D tensor box, 16 rows x 2 KB runs, 32 KB x 5", 2, 5, 32768, 1}, {"contiguous 32 KB x 5", 1, 5, 32768, 1},
                        {"3-D tensor box 16 KB x 10, no L2 promotion", 0, 10, 16384, 0},
                        {"3-D tensor box 16 KB x 8", 0, 8, 16384, 1}, {"3-D tensor box 16 KB x 7", 0, 7, 16384, 1},
                        {"3-D tensor box 16 KB x 6", 0, 6, 16384, 1}, {"3-D tensor box 16 KB x 5", 0, 5, 16384, 1},
                        {"3-D tensor box 16 KB x 4", 0, 4, 16384, 1}, {"3-D tensor box 16 KB x 3", 0, 3, 16384, 1},
                        // the swap-AB GEMM's weight stages: 128 rows x 512 B runs
                        {"128 rows x 512 B (64 KB) x 2, 148 CTAs", 0, 2, 65536, 1, 128, 148}, {"128 rows x 512 B (64 KB) x 2, 96 CTAs", 0, 2, 65536, 1, 128, 96},
                        {"128 rows x 512 B (64 KB) x 3, 96 CTAs", 0, 3, 65536, 1, 128, 96}, {"128 rows x 256 B (32 KB) x 4, 96 CTAs", 0, 4, 32768, 1, 128, 96},
                        {"128 rows x 256 B (32 KB) x 6, 96 CTAs", 0, 6, 32768, 1, 128, 96}, {"64 rows x 1 KB (64 KB) x 2, 96 CTAs", 0, 2, 65536, 1, 64, 96},
                        {"contiguous 64 KB x 2, 96 CTAs", 1, 2, 65536, 1, 128, 96}, {"contiguous 64 KB x 3, 148 CTAs", 1, 3, 65536, 1, 128, 148}};
  printf("matrix %d x %d bf16 = %.0f MB, 148 CTAs\n", N, K, (double)N * K * 2e-6);
  for (const Case& c : cases) {
    CUtensorMap m;
    cuuint64_t dims[3] = {64, (cuuint64_t)N, (cuuint64_t)(K / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)c.rows, c.stage / (128 * c.rows)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     c.l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      k<<<c.ctas, 64, smem>>>(m, w, N, K, c.mode, c.depth, c.stage, c.rows);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("%-52s: %.1f us  %.2f TB/s  err=%s\n", c.name, best * 1e3, (double)N * K * 2 / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
