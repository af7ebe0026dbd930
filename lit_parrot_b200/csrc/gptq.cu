// GPTQ quantiser on the device — the step BEFORE the int4 path: reference quantize/gptq.py:267-431
// (GPTQQuantizer: collect_input_stats 349-363, find_params_weight 318-347, quantize 365-431).
//
//   Hessian   H <- H * n / (n + b) + (sqrt(2 / (n + b)) X)^T (sqrt(2 / (n + b)) X)            (gptq.py:349-363)
//     X fp32 [rows, K] are the inputs of the linear layer for one batch of calibration tokens.  The product is a GEMM with the
//     TOKENS as the reduction dimension; it runs on the tcgen05 GEMM of gemm_tc.cu (TMA-fed, accumulator in tensor memory):
//     gptq_terms_kernel transposes the scaled X and splits every value into bf16 terms x = t0 + t1 (+ t2); the term products
//     t_a^T t_b with a + b <= terms - 1 are laid side by side along the reduction dimension, so ONE GEMM call with
//     reduction length P * rows (P = 3 or 6) and residual H accumulates the update in fp32 with the products exact.
//   Cholesky  H^-1 and its upper factor (gptq.py:387-391) are cuSOLVER calls issued by the host side (torch.linalg): a one-off
//     O(K^3) factorisation per layer, not restated here.
//   Sweep     per block of <= 128 columns (gptq.py:393-424): rows are independent, so ONE WARP owns a row: its 128 block values live
//     in 4 registers per lane, the 128 x 128 block of the factor in shared memory; per column: quantise (round-to-nearest on the
//     group's grid), error / d, rank-1 update of the remaining columns of the block — all in registers, with the reference's
//     fp32 operation order (divide, rint, product rounded, then subtracted), so given the same inputs the integer codes of a
//     block are bit-identical to the reference's.
//   Trailing  W[:, i2:] -= Err . Hinv[i1:i2, i2:] (gptq.py:424): fp32 CUDA-core GEMM with a 128-long reduction (the error feedback
//     wants fp32 products; 0.1-0.5 TFLOP per layer, a few milliseconds).
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace lp {

// ------------------------------------------------------------------------------------------------ Hessian operands
// x [rows, K] -> A, B bf16 [K, P * rows_pad]: segment p holds term ta[p] of (s x)^T in A and term tb[p] in B.
// 32 x 32 tiles through shared memory: coalesced reads along K, coalesced writes along the token dimension.
__global__ void __launch_bounds__(256) gptq_terms_kernel(const float* __restrict__ x, int rows, int K, int rows_pad, float s, int terms,
                                                         __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ B) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, k = k0 + tx;
    tile[j][tx] = (r < rows && k < K) ? __fmul_rn(x[(size_t)r * K + k], s) : 0.f;
  }
  __syncthreads();
  const int P = terms == 2 ? 3 : 6;
  // (a, b) pairs in order of magnitude: 2 terms: (0,0) (0,1) (1,0); 3 terms: + (0,2) (2,0) (1,1)
  const int ta[6] = {0, 0, 1, 0, 2, 1}, tb[6] = {0, 1, 0, 2, 0, 1};
  for (int j = ty; j < 32; j += 8) {
    const int k = k0 + j, r = r0 + tx;
    if (k >= K || r >= rows_pad) continue;
    float v = tile[tx][j];
    __nv_bfloat16 t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      t[i] = __float2bfloat16_rn(v);
      v -= __bfloat162float(t[i]);  // exact
    }
    const size_t base = (size_t)k * P * rows_pad + r;
    for (int p = 0; p < P; ++p) {
      A[base + (size_t)p * rows_pad] = t[ta[p]];
      B[base + (size_t)p * rows_pad] = t[tb[p]];
    }
  }
}

__global__ void __launch_bounds__(256) gptq_scale_kernel(float* __restrict__ H, size_t n, float a) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) H[i] = __fmul_rn(H[i], a);
}

// ------------------------------------------------------------------------------------------------ grid parameters
// find_params_weight (gptq.py:318-347), perchannel: one warp per (row, group) over columns [c, c + group) clipped to K.
__global__ void __launch_bounds__(256) gptq_find_params_kernel(const float* __restrict__ W, int N, int K, int g_first, int n_groups, int group,
                                                               float maxq, int sym, float* __restrict__ scales, float* __restrict__ zeros,
                                                               int ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N * n_groups) return;
  const int row = warp / n_groups, gi = g_first + warp % n_groups;
  const int c0 = gi * group, c1 = min(K, c0 + group);
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (int c = c0 + lane; c < c1; c += 32) {
    const float v = W[(size_t)row * K + c];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    float xmin = fminf(mn, 0.f), xmax = fmaxf(mx, 0.f);
    if (sym) {
      xmax = fmaxf(fabsf(xmin), xmax);
      if (xmin < 0.f) xmin = -xmax;
    }
    if (xmin == 0.f && xmax == 0.f) {
      xmin = -1.f;
      xmax = 1.f;
    }
    const float scale = __fdiv_rn(__fsub_rn(xmax, xmin), maxq);
    const float zero = sym ? (maxq + 1.f) * 0.5f : rintf(__fdiv_rn(-xmin, scale));
    scales[(size_t)row * ld + gi] = scale;
    zeros[(size_t)row * ld + gi] = zero;
  }
}

// ------------------------------------------------------------------------------------------------ block sweep
constexpr int GQ_BLOCK = 128;  // columns per block (the reference's blocksize)
constexpr int GQ_WARPS = 8;    // rows per CTA

__global__ void __launch_bounds__(GQ_WARPS * 32) gptq_sweep_kernel(const float* __restrict__ W, int N, int K, int i1, int count,
                                                                    const float* __restrict__ Hinv, const float* __restrict__ scales,
                                                                    const float* __restrict__ zeros, int ld_sz, int group, float maxq,
                                                                    float* __restrict__ Q, float* __restrict__ Err, float* __restrict__ loss) {
  extern __shared__ float Hs[];  // [count][GQ_BLOCK]: rows i1 .. i1+count of the factor, columns of the block
  for (int e = threadIdx.x; e < count * GQ_BLOCK; e += blockDim.x) {
    const int i = e / GQ_BLOCK, c = e % GQ_BLOCK;
    Hs[e] = c < count ? Hinv[(size_t)(i1 + i) * K + i1 + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * GQ_WARPS + warp;
  if (row >= N) return;
  float w[4], qv[4], ev[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = r * 32 + lane;
    w[r] = c < count ? W[(size_t)row * K + i1 + c] : 0.f;
    qv[r] = ev[r] = 0.f;
  }
  float lsum = 0.f;
  float scale = 1.f, zero = 0.f;
  int cur_group = -1;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    for (int li = 0; li < 32; ++li) {
      const int i = r * 32 + li;
      if (i >= count) break;  // warp uniform
      const int gi = (i1 + i) / group;
      if (gi != cur_group) {  // the group's grid (computed by gptq_find_params_kernel from the block-start values, gptq.py:409-412)
        cur_group = gi;
        scale = scales[(size_t)row * ld_sz + gi];
        zero = zeros[(size_t)row * ld_sz + gi];
      }
      const float wi = __shfl_sync(0xffffffffu, w[r], li);
      const float d = Hs[i * GQ_BLOCK + i];
      // quantize_weight (gptq.py:313-316): q = clamp(round(x / scale) + zero, 0, maxq); scale * (q - zero)
      const float qi = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(wi, scale)), zero), 0.f), maxq);
      const float q = __fmul_rn(scale, __fsub_rn(qi, zero));
      const float diff = __fsub_rn(wi, q);
      const float err = __fdiv_rn(diff, d);
      lsum += __fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d));  // Losses1[:, i] = (w - q) ** 2 / d ** 2
      if (lane == li) {
        qv[r] = q;
        ev[r] = err;
      }
      // W1[:, i:] -= err (outer) Hinv1[i, i:]: product rounded, then subtracted (matmul of a column with a row, then `-=`)
#pragma unroll
      for (int r2 = 0; r2 < 4; ++r2) {
        const int c = r2 * 32 + lane;
        if (r2 >= r && c >= i && c < count) w[r2] = __fsub_rn(w[r2], __fmul_rn(err, Hs[i * GQ_BLOCK + c]));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = r * 32 + lane;
    if (c < count) {
      Q[(size_t)row * K + i1 + c] = qv[r];
      Err[(size_t)row * GQ_BLOCK + c] = ev[r];
    }
  }
  if (lane == 0) loss[row] += 0.5f * lsum;  // Losses[:, i1:i2] = Losses1 / 2 (summed per row here; one warp owns the row)
}

// ------------------------------------------------------------------------------------------------ trailing update
// W[:, i2 + c] -= sum_i Err[:, i] * Hinv[i1 + i, i2 + c]: 64 x 64 output tile per CTA, 256 threads x (4 x 4), reduction in chunks of 16.
__global__ void __launch_bounds__(256) gptq_trailing_kernel(float* __restrict__ W, int N, int K, int i1, int count, int i2,
                                                            const float* __restrict__ Hinv, const float* __restrict__ Err) {
  __shared__ float As[16][65];  // Err chunk, transposed: [i][row]
  __shared__ float Bs[16][64];  // Hinv chunk: [i][col]
  const int R = K - i2;
  const int n0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < count; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int rr = e / 16, ii = e % 16;
      As[ii][rr] = (n0 + rr < N && k0 + ii < count) ? Err[(size_t)(n0 + rr) * GQ_BLOCK + k0 + ii] : 0.f;
      const int i = e / 64, cc = e % 64;
      Bs[i][cc] = (c0 + cc < R && k0 + i < count) ? Hinv[(size_t)(i1 + k0 + i) * K + i2 + c0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = As[i][ty * 4 + u];
        b[u] = Bs[i][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int n = n0 + ty * 4 + u;
    if (n >= N) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = c0 + tx * 4 + v;
      if (c < R) W[(size_t)n * K + i2 + c] -= acc[u][v];
    }
  }
}

// out = silu(a) * b (LLaMAMLP, model.py:298-300) for the calibration forward, where fc_1 and fc_2 are separate layers (one of them may
// already be quantised, so the interleaved SwiGLU weight of the inference engine does not exist)
__global__ void __launch_bounds__(256) swiglu_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                     size_t n, int round_bf16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = maybe_round(maybe_round(silu(a[i]), round_bf16) * b[i], round_bf16);
}

}  // namespace lp

extern "C" {

int lp_swiglu(const float* a, const float* b, float* out, size_t n, int round_bf16, void* stream) {
  if (!a || !b || !out || n == 0) return LP_ERR_INVALID_ARG;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)lp::num_sms() * 8);
  lp::swiglu_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, out, n, round_bf16);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

size_t lp_gptq_hessian_workspace_bytes(int rows, int K, int terms) {
  if (rows <= 0 || K <= 0 || (terms != 2 && terms != 3)) return 0;
  const size_t rows_pad = (size_t)(rows + 7) / 8 * 8, P = terms == 2 ? 3 : 6;
  return 2 * ((size_t)K * P * rows_pad * sizeof(__nv_bfloat16) + 256);
}

int lp_gptq_hessian_update(float* H, int K, const float* x, int rows, float keep, float x_scale, int terms, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!H || !x || !workspace || K <= 0 || rows <= 0 || (terms != 2 && terms != 3)) return LP_ERR_INVALID_ARG;
  if (K % 8) return LP_ERR_UNSUPPORTED;  // 16-byte row stride of the GEMM operands
  if (workspace_bytes < lp_gptq_hessian_workspace_bytes(rows, K, terms)) return LP_ERR_WORKSPACE;
  const int rows_pad = (rows + 7) / 8 * 8, P = terms == 2 ? 3 : 6;
  const size_t half = ((size_t)K * P * rows_pad * sizeof(__nv_bfloat16) + 255) / 256 * 256;
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + half);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  lp::gptq_scale_kernel<<<lp::num_sms() * 4, 256, 0, st>>>(H, (size_t)K * K, keep);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  lp::gptq_terms_kernel<<<dim3((rows_pad + 31) / 32, (K + 31) / 32), 256, 0, st>>>(x, rows, K, rows_pad, x_scale, terms, A, B);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  // H += A . B^T over the P * rows_pad reduction: the tcgen05 GEMM with the residual epilogue, in place
  return lp_gemm_bf16_tc(A, 1, K, B, K, P * rows_pad, nullptr, LP_EPI_RESIDUAL, H, H, nullptr, 0, 0, stream);
}

int lp_gptq_find_params(const float* W, int N, int K, int group_first, int n_groups, int group, int maxq, int sym, float* scales,
                        float* zeros, int ld, void* stream) {
  if (!W || !scales || !zeros || N <= 0 || K <= 0 || group <= 0 || n_groups <= 0 || group_first < 0 || maxq <= 0) return LP_ERR_INVALID_ARG;
  if ((long long)(group_first + n_groups - 1) * group >= K || group_first + n_groups > ld) return LP_ERR_INVALID_ARG;
  const long long warps = (long long)N * n_groups;
  lp::gptq_find_params_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, N, K, group_first, n_groups, group, (float)maxq, sym, scales, zeros, ld);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

int lp_gptq_block_sweep(const float* W, int N, int K, int i1, int count, const float* Hinv, const float* scales, const float* zeros,
                        int ld, int group, int maxq, float* Q, float* Err, float* loss_rows, void* stream) {
  if (!W || !Hinv || !scales || !zeros || !Q || !Err || !loss_rows || N <= 0 || K <= 0 || i1 < 0 || count <= 0 || i1 + count > K ||
      group <= 0 || maxq <= 0)
    return LP_ERR_INVALID_ARG;
  if (count > lp::GQ_BLOCK) return LP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)count * lp::GQ_BLOCK * sizeof(float);
  static bool attr_set[64] = {};
  int dev = 0;
  LP_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    LP_CUDA_TRY(cudaFuncSetAttribute(lp::gptq_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lp::GQ_BLOCK * lp::GQ_BLOCK * 4));
    attr_set[dev] = true;
  }
  lp::gptq_sweep_kernel<<<(N + lp::GQ_WARPS - 1) / lp::GQ_WARPS, lp::GQ_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, N, K, i1, count, Hinv, scales, zeros, ld, group, (float)maxq, Q, Err, loss_rows);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

int lp_gptq_trailing_update(float* W, int N, int K, int i1, int count, const float* Hinv, const float* Err, void* stream) {
  if (!W || !Hinv || !Err || N <= 0 || K <= 0 || i1 < 0 || count <= 0 || count > lp::GQ_BLOCK || i1 + count > K) return LP_ERR_INVALID_ARG;
  const int i2 = i1 + count, R = K - i2;
  if (R == 0) return LP_OK;
  lp::gptq_trailing_kernel<<<dim3((R + 63) / 64, (N + 63) / 64), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, N, K, i1, count, i2, Hinv,
                                                                                                                   Err);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

}  // extern "C"
