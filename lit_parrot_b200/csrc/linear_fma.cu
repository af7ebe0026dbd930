// Weight-streaming linear layer for small M (decode): y = x[M,K] . W[N,K]^T with fused epilogue.
//
// "FMA family": exact fp32 multiply-accumulate on CUDA cores, every weight format.  HBM-bound by
// construction: each warp owns two output rows and streams them with 128-bit L1-bypassing loads,
// several batches in flight (the first batch is issued BEFORE griddepcontrol.wait, so under PDL the weight
// stream of layer n+1 starts while layer n drains); x (a few KB) is read through L1.
// The tensor-core ("MMA") family in linear_mma.cu takes over where the FMA issue rate would become the
// limit (int4 at full HBM rate, batch 32).
#include "common.cuh"

namespace lp {

__constant__ float c_nf4_code[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                     -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                     0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                     0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

template <int FMT> struct FmtTraits;
template <> struct FmtTraits<LP_W_F32> { static constexpr int ELEMS = 4; };    // elements per 16-byte chunk
template <> struct FmtTraits<LP_W_BF16> { static constexpr int ELEMS = 8; };
template <> struct FmtTraits<LP_W_INT8> { static constexpr int ELEMS = 16; };
template <> struct FmtTraits<LP_W_INT4> { static constexpr int ELEMS = 32; };
template <> struct FmtTraits<LP_W_NF4> { static constexpr int ELEMS = 32; };

constexpr int FMA_THREADS = 128;
constexpr int FMA_ROWS_PER_CTA = (FMA_THREADS / 32) * 2;
constexpr int FMA_U = 4;  // 16-byte loads in flight per lane per row

// decode `EL` weights of one 16-byte chunk into fp32.  `k0` = first column of the chunk.
template <int FMT>
__device__ __forceinline__ void decode_chunk(const uint4& v, float* w, const lp_weight& W, int row, int k0, int round_bf16,
                                             const float* s_nf4) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
  if constexpr (FMT == LP_W_F32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __uint_as_float(u[i]);
  } else if constexpr (FMT == LP_W_BF16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[2 * i] = bf16lo(u[i]);
      w[2 * i + 1] = bf16hi(u[i]);
    }
  } else if constexpr (FMT == LP_W_INT8) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int b = 0; b < 4; ++b) w[4 * i + b] = (float)(int)(signed char)((u[i] >> (8 * b)) & 0xff);
  } else if constexpr (FMT == LP_W_INT4) {
    // w = (q - zero) * scale evaluated like the reference (gptq.py:246-251), optionally rounded to bf16
    const int ngroups = (W.K + W.group - 1) / W.group;
    const int gi = k0 / W.group;
    const float sc = W.aux0[(size_t)row * ngroups + gi], ze = W.aux1[(size_t)row * ngroups + gi];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float q = (float)((u[i] >> (4 * n)) & 0xf);
        float d = __fmul_rn(__fsub_rn(q, ze), sc);
        if (round_bf16) d = bf16_round(__fmul_rn(bf16_round(__fsub_rn(q, ze)), sc));
        w[8 * i + n] = d;
      }
  } else {  // NF4: first element of a byte pair in the HIGH nibble; absmax per `group` flattened elements
    const size_t flat = (size_t)row * W.K + k0;
    const float am = W.aux0[flat / W.group];  // a 32-element chunk never straddles a 64-element block
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t byte = (u[i] >> (8 * b)) & 0xff;
        float hi = s_nf4[byte >> 4] * am, lo = s_nf4[byte & 0xf] * am;
        w[8 * i + 2 * b] = round_bf16 ? bf16_round(hi) : hi;
        w[8 * i + 2 * b + 1] = round_bf16 ? bf16_round(lo) : lo;
      }
  }
}

template <int FMT, int M>
__global__ void __launch_bounds__(FMA_THREADS)
linear_fma_kernel(const float* __restrict__ x, int m_actual, lp_weight W, int epi, const float* __restrict__ residual,
                  float* __restrict__ out, int round_bf16) {
  constexpr int EL = FmtTraits<FMT>::ELEMS;
  __shared__ float s_nf4[16];
  if constexpr (FMT == LP_W_NF4) {
    if (threadIdx.x < 16) s_nf4[threadIdx.x] = c_nf4_code[threadIdx.x];
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (blockIdx.x * (FMA_THREADS / 32) + warp) * 2;
  const int N = W.N, K = W.K;
  if (r0 >= N) {
    pdl_wait();
    return;
  }
  const bool has1 = (r0 + 1) < N;
  size_t row_bytes;
  if constexpr (FMT == LP_W_F32) row_bytes = (size_t)K * 4;
  else if constexpr (FMT == LP_W_BF16) row_bytes = (size_t)K * 2;
  else if constexpr (FMT == LP_W_INT8) row_bytes = (size_t)K;
  else if constexpr (FMT == LP_W_INT4) row_bytes = (size_t)((K + 255) / 256 * 256) / 2;
  else row_bytes = (size_t)K / 2;
  const char* p0 = reinterpret_cast<const char*>(W.w) + (size_t)r0 * row_bytes;
  const char* p1 = p0 + (has1 ? row_bytes : 0);
  const int nchunks = K / EL;

  float acc[2][M];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) acc[r][m] = 0.f;

  uint4 cur[FMA_U][2];
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < FMA_U; ++u) {
    const int c = lane + 32 * u;
    cur[u][0] = (c < nchunks) ? ldg_stream(p0 + (size_t)c * 16) : zero4;
    cur[u][1] = (c < nchunks) ? ldg_stream(p1 + (size_t)c * 16) : zero4;
  }
  pdl_wait();  // activations of the previous kernel are visible from here on
  pdl_launch_dependents();

  for (int c0 = lane; c0 < nchunks; c0 += 32 * FMA_U) {
    uint4 nxt[FMA_U][2];
#pragma unroll
    for (int u = 0; u < FMA_U; ++u) {
      const int c = c0 + 32 * (FMA_U + u);
      nxt[u][0] = (c < nchunks) ? ldg_stream(p0 + (size_t)c * 16) : zero4;
      nxt[u][1] = (c < nchunks) ? ldg_stream(p1 + (size_t)c * 16) : zero4;
    }
#pragma unroll
    for (int u = 0; u < FMA_U; ++u) {
      const int c = c0 + 32 * u;
      if (c < nchunks) {
        const int k0 = c * EL;
        float w0[EL], w1[EL];
        decode_chunk<FMT>(cur[u][0], w0, W, r0, k0, round_bf16, s_nf4);
        decode_chunk<FMT>(cur[u][1], w1, W, has1 ? r0 + 1 : r0, k0, round_bf16, s_nf4);
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if (m < m_actual) {
            const float4* xp = reinterpret_cast<const float4*>(x + (size_t)m * K + k0);
#pragma unroll
            for (int q = 0; q < EL / 4; ++q) {
              const float4 xv = __ldg(xp + q);
              acc[0][m] = fmaf(w0[4 * q + 0], xv.x, acc[0][m]);
              acc[0][m] = fmaf(w0[4 * q + 1], xv.y, acc[0][m]);
              acc[0][m] = fmaf(w0[4 * q + 2], xv.z, acc[0][m]);
              acc[0][m] = fmaf(w0[4 * q + 3], xv.w, acc[0][m]);
              acc[1][m] = fmaf(w1[4 * q + 0], xv.x, acc[1][m]);
              acc[1][m] = fmaf(w1[4 * q + 1], xv.y, acc[1][m]);
              acc[1][m] = fmaf(w1[4 * q + 2], xv.z, acc[1][m]);
              acc[1][m] = fmaf(w1[4 * q + 3], xv.w, acc[1][m]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < FMA_U; ++u) {
      cur[u][0] = nxt[u][0];
      cur[u][1] = nxt[u][1];
    }
  }

#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);

  if (lane < M && lane < m_actual) {
    const int m = lane;
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int mm = 0; mm < M; ++mm)
      if (mm == m) {
        y0 = acc[0][mm];
        y1 = acc[1][mm];
      }
    if constexpr (FMT == LP_W_INT8) {
      y0 *= W.aux0[r0];
      y1 *= W.aux0[has1 ? r0 + 1 : r0];
    }
    if (W.bias) {
      y0 += W.bias[r0];
      y1 += W.bias[has1 ? r0 + 1 : r0];
    }
    y0 = maybe_round(y0, round_bf16);
    y1 = maybe_round(y1, round_bf16);
    y0 = out_affine(y0, W.out_bias, W.out_scale, r0, round_bf16);
    y1 = out_affine(y1, W.out_bias, W.out_scale, has1 ? r0 + 1 : r0, round_bf16);
    if (epi == LP_EPI_SWIGLU) {
      // rows (2i, 2i+1) = (fc_1 row i, fc_2 row i): silu(a) * b, each step rounded in bf16 mode (model.py:300)
      const float a = maybe_round(silu(y0), round_bf16);
      out[(size_t)m * (N / 2) + (r0 >> 1)] = maybe_round(a * y1, round_bf16);
    } else {
      if (epi == LP_EPI_GELU) {
        y0 = maybe_round(gelu_erf(y0), round_bf16);
        y1 = maybe_round(gelu_erf(y1), round_bf16);
      } else if (epi == LP_EPI_RESIDUAL) {
        y0 = maybe_round(residual[(size_t)m * N + r0] + y0, round_bf16);
        if (has1) y1 = maybe_round(residual[(size_t)m * N + r0 + 1] + y1, round_bf16);
      }
      out[(size_t)m * N + r0] = y0;
      if (has1) out[(size_t)m * N + r0 + 1] = y1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// W (any stored format) -> dense bf16 [N, K] row-major, for the tensor-core prefill GEMM.  One thread per 16-byte chunk.
// int4 / NF4 are dequantised exactly like the reference does for bf16 activations: (q - zero) * scale, resp.
// code * absmax, rounded to bf16 (quantize/gptq.py:243-252).
// ---------------------------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) dequant_bf16_kernel(lp_weight W, __nv_bfloat16* __restrict__ out) {
  constexpr int EL = FmtTraits<FMT>::ELEMS;
  __shared__ float s_nf4[16];
  if (threadIdx.x < 16) s_nf4[threadIdx.x] = c_nf4_code[threadIdx.x];
  __syncthreads();
  const int K = W.K, nchunks = K / EL;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)W.N * nchunks) return;
  const int row = (int)(idx / nchunks), c = (int)(idx % nchunks);
  size_t row_bytes;
  if constexpr (FMT == LP_W_F32) row_bytes = (size_t)K * 4;
  else if constexpr (FMT == LP_W_BF16) row_bytes = (size_t)K * 2;
  else if constexpr (FMT == LP_W_INT8) row_bytes = (size_t)K;
  else if constexpr (FMT == LP_W_INT4) row_bytes = (size_t)((K + 255) / 256 * 256) / 2;
  else row_bytes = (size_t)K / 2;
  const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(W.w) + (size_t)row * row_bytes + (size_t)c * 16);
  float w[EL];
  decode_chunk<FMT>(v, w, W, row, c * EL, 0, s_nf4);
  const float rs = (FMT == LP_W_INT8) ? W.aux0[row] : 1.0f;
  __nv_bfloat16* dst = out + (size_t)row * K + (size_t)c * EL;
#pragma unroll
  for (int i = 0; i < EL; ++i) dst[i] = __float2bfloat16_rn(w[i] * rs);
}

int dequant_bf16(const lp_weight& W, void* out, void* stream) {
  const int K = W.K;
  auto go = [&](auto kern, int el) {
    if (K % el) return (int)LP_ERR_UNSUPPORTED;
    const long long n = (long long)W.N * (K / el);
    return launch(kern, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, W, reinterpret_cast<__nv_bfloat16*>(out));
  };
  switch (W.fmt) {
    case LP_W_F32: return go(dequant_bf16_kernel<LP_W_F32>, 4);
    case LP_W_BF16: return go(dequant_bf16_kernel<LP_W_BF16>, 8);
    case LP_W_INT8: return W.aux0 ? go(dequant_bf16_kernel<LP_W_INT8>, 16) : (int)LP_ERR_INVALID_ARG;
    case LP_W_INT4: return (W.aux0 && W.aux1 && W.group > 0 && W.group % 32 == 0) ? go(dequant_bf16_kernel<LP_W_INT4>, 32) : (int)LP_ERR_UNSUPPORTED;
    case LP_W_NF4: return (W.aux0 && W.group > 0 && W.group % 32 == 0) ? go(dequant_bf16_kernel<LP_W_NF4>, 32) : (int)LP_ERR_UNSUPPORTED;
    default: return LP_ERR_INVALID_ARG;
  }
}

template <int FMT>
static int launch_fma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16,
                      void* stream) {
  dim3 grid((W.N + FMA_ROWS_PER_CTA - 1) / FMA_ROWS_PER_CTA), block(FMA_THREADS);
  if (M == 1) return launch(linear_fma_kernel<FMT, 1>, grid, block, 0, stream, x, M, W, epi, residual, out, round_bf16);
  if (M == 2) return launch(linear_fma_kernel<FMT, 2>, grid, block, 0, stream, x, M, W, epi, residual, out, round_bf16);
  if (M <= 4) return launch(linear_fma_kernel<FMT, 4>, grid, block, 0, stream, x, M, W, epi, residual, out, round_bf16);
  return launch(linear_fma_kernel<FMT, 8>, grid, block, 0, stream, x, M, W, epi, residual, out, round_bf16);
}

int linear_fma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16,
               void* stream) {
  const int K = W.K;
  switch (W.fmt) {
    case LP_W_F32:
      if (K % 4) return LP_ERR_UNSUPPORTED;
      return launch_fma<LP_W_F32>(x, M, W, epi, residual, out, round_bf16, stream);
    case LP_W_BF16:
      if (K % 8) return LP_ERR_UNSUPPORTED;
      return launch_fma<LP_W_BF16>(x, M, W, epi, residual, out, round_bf16, stream);
    case LP_W_INT8:
      if (K % 16 || !W.aux0) return LP_ERR_UNSUPPORTED;
      return launch_fma<LP_W_INT8>(x, M, W, epi, residual, out, round_bf16, stream);
    case LP_W_INT4:
      if (K % 32 || W.group <= 0 || W.group % 32 || !W.aux0 || !W.aux1) return LP_ERR_UNSUPPORTED;
      return launch_fma<LP_W_INT4>(x, M, W, epi, residual, out, round_bf16, stream);
    case LP_W_NF4:
      if (K % 32 || W.group <= 0 || W.group % 32 || !W.aux0) return LP_ERR_UNSUPPORTED;
      return launch_fma<LP_W_NF4>(x, M, W, epi, residual, out, round_bf16, stream);
    default:
      return LP_ERR_INVALID_ARG;
  }
}

}  // namespace lp
