// Micro-benchmark: issue rate of legacy mma.sync (HMMA bf16 m16n8k16, IMMA s8 m16n8k32) on sm_100a, per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int CHAINS>
__global__ void rate(int iters, int* out) {
  int acc[CHAINS][4];
  float facc[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) { acc[c][i] = 0; facc[c][i] = 0.f; }
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(facc[c][0]), "+f"(facc[c][1]), "+f"(facc[c][2]), "+f"(facc[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  int s = 0;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += acc[c][i] + (int)facc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int CHAINS>
void run(const char* name, int warps) {
  int* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<KIND, CHAINS><<<148, warps * 32>>>(16, out);
  cudaEventRecord(e0);
  rate<KIND, CHAINS><<<148, warps * 32>>>(iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas_per_sm = (double)iters * CHAINS * warps;
  double cyc = ms * 1e-3 * 1.965e9;
  printf("%s warps/SM=%2d chains=%d: %.2f cycles per MMA per SM (%.1f per SMSP), %.1f dense TFLOP/s-equivalent\n", name, warps, CHAINS,
         cyc / mmas_per_sm, 4 * cyc / mmas_per_sm, 148.0 * mmas_per_sm * (KIND == 0 ? 4096 : 8192) * 2 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  run<0, 1>("HMMA.16816 bf16", 4); run<0, 4>("HMMA.16816 bf16", 4); run<0, 4>("HMMA.16816 bf16", 16); run<0, 8>("HMMA.16816 bf16", 16);
  run<1, 1>("IMMA.16832 u8s8", 4); run<1, 4>("IMMA.16832 u8s8", 4); run<1, 4>("IMMA.16832 u8s8", 16); run<1, 8>("IMMA.16832 u8s8", 16);
  return 0;
}
