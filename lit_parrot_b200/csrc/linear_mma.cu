// Weight-streaming linear layer, "MMA family": the multiply-accumulates run on the tensor cores
// (mma.sync.m16n8k16, bf16 x bf16 -> fp32) so that the instruction issue slots are left for what an HBM-bound
// int4 GEMV is actually limited by — nibble unpacking.  The kernel is judged on GB/s, not on tensor utilisation.
//
//   * the WEIGHTS are the 16-row A operand, streamed straight from global memory into registers with 128-bit
//     L1-bypassing loads (each lane owns 16 bytes of row g and of row g+8 per chunk); the k-order inside an MMA is a
//     free permutation, so no shuffles are needed: it is the x operand that is stored permuted;
//   * int4: nibble -> bf16 with ONE lop3 per two weights ((w & 0x000F000F) | 0x43004300 = bf16x2 {128+q_lo, 128+q_hi});
//     the +128 and the GPTQ zero point are removed algebraically per 128-column chunk:
//        sum (q - z) s x = s * (sum (128+q) x  -  (128 + z) * sum x)
//   * x (one or a few activation rows) is the 8-column B operand.  To keep fp32 ACTIVATION accuracy every row is
//     split into bf16 terms x = hi + mid (+ lo) occupying separate columns — the products q*hi are exact in the
//     fp32 accumulator — and the columns are added in the epilogue.  In bf16-faithful mode x is already bf16: 1 column;
//   * 8 warps split K; the first weight chunks are requested BEFORE griddepcontrol.wait (PDL), then x is staged to
//     shared memory (optionally through a fused LayerNorm / RMSNorm), then the main loop runs.
#include "common.cuh"

namespace lp {

constexpr int MMA_WARPS = 8;
constexpr int MMA_THREADS = MMA_WARPS * 32;
constexpr int MMA_ROWS = 16;
constexpr int MMA_U = 4;  // weight chunks in flight per warp

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t nib_to_bf16x2(uint32_t w) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;\n" : "=r"(r) : "r"(w));  // (w & m) | magic
  return r;
}

__device__ __forceinline__ uint16_t bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

struct NormArgs {
  const float* w;
  const float* b;
  float eps;
  int kind;  // -1: none, else lp_norm_kind
};

// chunk geometry: elements of K covered by one warp iteration
template <int FMT> struct MmaTraits;
template <> struct MmaTraits<LP_W_BF16> { static constexpr int CH = 64; static constexpr int NLD = 4; };  // 4 x 16B per lane
template <> struct MmaTraits<LP_W_INT4> { static constexpr int CH = 128; static constexpr int NLD = 2; };

template <int FMT> struct Chunk {
  uint4 w[MmaTraits<FMT>::NLD];
  float s0, s1, z0, z1;  // int4 only
};

template <int FMT>
__device__ __forceinline__ void load_chunk(Chunk<FMT>& ch, const char* base, size_t row_bytes, int c, int g, int t,
                                           const lp_weight& W, int row0, int ngroups) {
  if constexpr (FMT == LP_W_BF16) {
    const char* p0 = base + (size_t)g * row_bytes + (size_t)c * 128 + t * 16;
    const char* p1 = p0 + 8 * row_bytes;
    ch.w[0] = ldg_stream(p0);
    ch.w[1] = ldg_stream(p1);
    ch.w[2] = ldg_stream(p0 + 64);
    ch.w[3] = ldg_stream(p1 + 64);
  } else {
    const char* p0 = base + (size_t)g * row_bytes + (size_t)c * 64 + t * 16;
    ch.w[0] = ldg_stream(p0);
    ch.w[1] = ldg_stream(p0 + 8 * row_bytes);
    const int gi = (c * 128) / W.group;
    const size_t i0 = (size_t)(row0 + g) * ngroups + gi, i1 = (size_t)(row0 + g + 8) * ngroups + gi;
    ch.s0 = __ldg(W.aux0 + i0);
    ch.s1 = __ldg(W.aux0 + i1);
    ch.z0 = __ldg(W.aux1 + i0);
    ch.z1 = __ldg(W.aux1 + i1);
  }
}

template <int FMT>
__global__ void __launch_bounds__(MMA_THREADS)
linear_mma_kernel(const float* __restrict__ x, int M, int split, lp_weight W, NormArgs nrm, int epi,
                  const float* __restrict__ residual, float* __restrict__ out, int round_bf16, int ldx, int nchunks) {
  constexpr int CH = MmaTraits<FMT>::CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ncols = M * split;
  uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);                       // [ncols][ldx] bf16 bits
  float* xsum = reinterpret_cast<float*>(smem_raw + (size_t)ncols * ldx * 2);  // [nchunks][8]   (int4 only)
  float* red = xsum + (FMT == LP_W_INT4 ? nchunks * 8 : 0);                    // [MMA_WARPS][16][8]
  float* ys = red + MMA_WARPS * 16 * 8;                                        // [8][16]
  __shared__ float s_stat[2][MMA_WARPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * MMA_ROWS;
  const int K = W.K, N = W.N;
  const size_t row_bytes = (FMT == LP_W_BF16) ? (size_t)K * 2 : (size_t)((K + 127) / 128 * 128) / 2;
  const char* base = reinterpret_cast<const char*>(W.w) + (size_t)row0 * row_bytes;
  const int ngroups = (FMT == LP_W_INT4) ? (K + W.group - 1) / W.group : 0;

  // ---- 1. weight prefetch: independent of the previous kernel, so it is issued before the PDL wait --------------
  Chunk<FMT> ch[MMA_U];
#pragma unroll
  for (int u = 0; u < MMA_U; ++u) {
    const int c = warp + u * MMA_WARPS;
    if (c < nchunks) load_chunk<FMT>(ch[u], base, row_bytes, c, g, t, W, row0, ngroups);
  }
  pdl_wait();
  pdl_launch_dependents();

  // ---- 2. stage x: (optional norm) -> bf16 split terms -> shared memory ----------------------------------------
  for (int m = 0; m < M; ++m) {
    const float* xr = x + (size_t)m * K;
    float mean = 0.f, rstd = 1.f;
    if (nrm.kind >= 0) {
      // statistics over the full row (every CTA recomputes them: K floats out of L2)
      float s = 0.f, ss = 0.f;
      for (int k = threadIdx.x * 4; k < K; k += MMA_THREADS * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + k);
        s += v.x + v.y + v.z + v.w;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      s = warp_sum(s);
      ss = warp_sum(ss);
      __syncthreads();
      if (lane == 0) {
        s_stat[0][warp] = s;
        s_stat[1][warp] = ss;
      }
      __syncthreads();
      s = ss = 0.f;
#pragma unroll
      for (int w = 0; w < MMA_WARPS; ++w) {
        s += s_stat[0][w];
        ss += s_stat[1][w];
      }
      if (nrm.kind == LP_NORM_LAYERNORM) {
        mean = s / (float)K;
        // two-pass variance for accuracy
        float v2 = 0.f;
        for (int k = threadIdx.x * 4; k < K; k += MMA_THREADS * 4) {
          const float4 v = *reinterpret_cast<const float4*>(xr + k);
          const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
          v2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        v2 = warp_sum(v2);
        __syncthreads();
        if (lane == 0) s_stat[0][warp] = v2;
        __syncthreads();
        v2 = 0.f;
#pragma unroll
        for (int w = 0; w < MMA_WARPS; ++w) v2 += s_stat[0][w];
        rstd = 1.0f / sqrtf(v2 / (float)K + nrm.eps);
      } else {
        rstd = 1.0f / sqrtf(ss / (float)K + nrm.eps);
      }
    }
    // one warp stages 128 consecutive columns per iteration (= one int4 chunk, so its x-sum is a warp reduction)
    const int n128 = (FMT == LP_W_INT4) ? nchunks : (K + 127) / 128;
    for (int c = warp; c < n128; c += MMA_WARPS) {
      const int k = c * 128 + lane * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < K) {  // K % 4 == 0
        const float4 xv = *reinterpret_cast<const float4*>(xr + k);
        v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
        if (nrm.kind >= 0) {
          const float4 wv = *reinterpret_cast<const float4*>(nrm.w + k);
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nrm.b) bv = *reinterpret_cast<const float4*>(nrm.b + k);
          if (nrm.kind == LP_NORM_LAYERNORM) {
            v[0] = (v[0] - mean) * rstd * wv.x + bv.x; v[1] = (v[1] - mean) * rstd * wv.y + bv.y;
            v[2] = (v[2] - mean) * rstd * wv.z + bv.z; v[3] = (v[3] - mean) * rstd * wv.w + bv.w;
          } else {
            v[0] = wv.x * (v[0] * rstd); v[1] = wv.y * (v[1] * rstd);
            v[2] = wv.z * (v[2] * rstd); v[3] = wv.w * (v[3] * rstd);
          }
        }
      }
      const bool in_smem = (FMT == LP_W_INT4) || (k < ldx);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (s < split) {
          uint16_t hb[4];
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            hb[i] = bf16_bits(v[i]);
            const float hv = __uint_as_float((uint32_t)hb[i] << 16);
            part += hv;
            v[i] -= hv;  // exact: next term of the split
          }
          uint16_t* dst = xs + (size_t)(m * split + s) * ldx;
          if (in_smem) {
            if constexpr (FMT == LP_W_INT4) {
              // within each group of 8 columns the order is [0,4,1,5,2,6,3,7] (matches the lop3 nibble pairs)
              const int k8 = k & ~7, p0 = (k & 4) ? 1 : 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[k8 + 2 * i + p0] = hb[i];
            } else {
              *reinterpret_cast<uint2*>(dst + k) = make_uint2(hb[0] | ((uint32_t)hb[1] << 16), hb[2] | ((uint32_t)hb[3] << 16));
            }
          }
          if constexpr (FMT == LP_W_INT4) {
            part = warp_sum(part);
            if (lane == 0) xsum[c * 8 + m * split + s] = part;
          }
        }
      }
      if constexpr (FMT == LP_W_INT4) {
        if (lane >= ncols && lane < 8) xsum[c * 8 + lane] = 0.f;
      }
    }
  }
  __syncthreads();

  // ---- 3. main loop: each warp walks its chunks (warp, warp + 8, ...) -------------------------------------------
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const uint16_t* xrow = xs + (size_t)(g < ncols ? g : 0) * ldx;
  const bool bvalid = g < ncols;
  for (int cb = warp; cb < nchunks; cb += MMA_WARPS * MMA_U) {
#pragma unroll
    for (int u = 0; u < MMA_U; ++u) {
      const int c = cb + u * MMA_WARPS;
      if (c < nchunks) {
        const Chunk<FMT> cur = ch[u];
        const int cn = c + MMA_U * MMA_WARPS;
        if (cn < nchunks) load_chunk<FMT>(ch[u], base, row_bytes, cn, g, t, W, row0, ngroups);
        if constexpr (FMT == LP_W_BF16) {
          // lane's columns: k0 = c*64 + t*8 (first 16 bytes) and k0 + 32 (second)
          const uint4 z4 = make_uint4(0, 0, 0, 0);
          const uint4 xa = bvalid ? *reinterpret_cast<const uint4*>(xrow + c * 64 + t * 8) : z4;
          const uint4 xb = bvalid ? *reinterpret_cast<const uint4*>(xrow + c * 64 + 32 + t * 8) : z4;
          mma_bf16_16816(acc, cur.w[0].x, cur.w[1].x, cur.w[0].y, cur.w[1].y, xa.x, xa.y);
          mma_bf16_16816(acc, cur.w[0].z, cur.w[1].z, cur.w[0].w, cur.w[1].w, xa.z, xa.w);
          mma_bf16_16816(acc, cur.w[2].x, cur.w[3].x, cur.w[2].y, cur.w[3].y, xb.x, xb.y);
          mma_bf16_16816(acc, cur.w[2].z, cur.w[3].z, cur.w[2].w, cur.w[3].w, xb.z, xb.w);
        } else {
          float cacc[4] = {0.f, 0.f, 0.f, 0.f};
          const uint32_t wa[4] = {cur.w[0].x, cur.w[0].y, cur.w[0].z, cur.w[0].w};
          const uint32_t wb[4] = {cur.w[1].x, cur.w[1].y, cur.w[1].z, cur.w[1].w};
          const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // 8 columns k0 = c*128 + t*32 + 8j of rows g (wa) and g+8 (wb); x holds them as [0,4,1,5,2,6,3,7]
            const uint4 xv = bvalid ? *reinterpret_cast<const uint4*>(xrow + c * 128 + t * 32 + j * 8) : z4;
            mma_bf16_16816(cacc, nib_to_bf16x2(wa[j]), nib_to_bf16x2(wb[j]), nib_to_bf16x2(wa[j] >> 4), nib_to_bf16x2(wb[j] >> 4),
                           xv.x, xv.y);
            mma_bf16_16816(cacc, nib_to_bf16x2(wa[j] >> 8), nib_to_bf16x2(wb[j] >> 8), nib_to_bf16x2(wa[j] >> 12),
                           nib_to_bf16x2(wb[j] >> 12), xv.z, xv.w);
          }
          const float2 xsv = *reinterpret_cast<const float2*>(xsum + c * 8 + 2 * t);
          const float o0 = 128.f + cur.z0, o1 = 128.f + cur.z1;
          acc[0] = fmaf(cur.s0, fmaf(-o0, xsv.x, cacc[0]), acc[0]);
          acc[1] = fmaf(cur.s0, fmaf(-o0, xsv.y, cacc[1]), acc[1]);
          acc[2] = fmaf(cur.s1, fmaf(-o1, xsv.x, cacc[2]), acc[2]);
          acc[3] = fmaf(cur.s1, fmaf(-o1, xsv.y, cacc[3]), acc[3]);
        }
      }
    }
  }

  // ---- 4. cross-warp reduction, column recombination, epilogue --------------------------------------------------
  {
    float* r = red + warp * 128;
    r[g * 8 + 2 * t] = acc[0];
    r[g * 8 + 2 * t + 1] = acc[1];
    r[(g + 8) * 8 + 2 * t] = acc[2];
    r[(g + 8) * 8 + 2 * t + 1] = acc[3];
  }
  __syncthreads();
  if (threadIdx.x < 16 * M) {
    const int r = threadIdx.x & 15, m = threadIdx.x >> 4;
    float y = 0.f;
    for (int s = split - 1; s >= 0; --s) {  // smallest terms first
      float part = 0.f;
#pragma unroll
      for (int w = 0; w < MMA_WARPS; ++w) part += red[w * 128 + r * 8 + m * split + s];
      y += part;
    }
    const int row = row0 + r;
    if (W.bias) y += W.bias[row];
    ys[m * 16 + r] = maybe_round(y, round_bf16);
  }
  __syncthreads();
  if (threadIdx.x < 16 * M) {
    const int r = threadIdx.x & 15, m = threadIdx.x >> 4;
    const int row = row0 + r;
    float y = ys[m * 16 + r];
    if (epi == LP_EPI_SWIGLU) {
      if ((r & 1) == 0) {
        const float a = maybe_round(silu(y), round_bf16);
        out[(size_t)m * (N / 2) + (row >> 1)] = maybe_round(a * ys[m * 16 + r + 1], round_bf16);
      }
    } else {
      if (epi == LP_EPI_GELU) y = maybe_round(gelu_erf(y), round_bf16);
      else if (epi == LP_EPI_RESIDUAL) y = maybe_round(residual[(size_t)m * N + row] + y, round_bf16);
      out[(size_t)m * N + row] = y;
    }
  }
}

constexpr size_t MMA_MAX_SMEM = 100 * 1024;

int init_linear_mma() {
  LP_CUDA_TRY(cudaFuncSetAttribute(linear_mma_kernel<LP_W_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_MAX_SMEM));
  LP_CUDA_TRY(cudaFuncSetAttribute(linear_mma_kernel<LP_W_INT4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_MAX_SMEM));
  return LP_OK;
}

int linear_mma_max_m() { return 8; }

// `nrm`: fused norm prologue (kind -1 = none)
int linear_mma_norm(const float* x, int M, const lp_weight& W, const NormArgs& nrm, int epi, const float* residual, float* out,
                    int round_bf16, void* stream) {
  if (W.fmt != LP_W_BF16 && W.fmt != LP_W_INT4) return LP_ERR_UNSUPPORTED;
  if (W.N % MMA_ROWS) return LP_ERR_UNSUPPORTED;
  int split;
  if (round_bf16 && nrm.kind < 0) {
    if (M > 8) return LP_ERR_UNSUPPORTED;
    split = 1;  // x is bf16-valued already
  } else {
    if (M > 4) return LP_ERR_UNSUPPORTED;
    split = M <= 2 ? 3 : 2;
  }
  const int K = W.K;
  if (K % 4) return LP_ERR_UNSUPPORTED;
  int nchunks, kpad;
  if (W.fmt == LP_W_BF16) {
    if (K % 64) return LP_ERR_UNSUPPORTED;
    nchunks = K / 64;
    kpad = (K + 127) / 128 * 128;
  } else {
    if (W.group <= 0 || W.group % 128 || !W.aux0 || !W.aux1) return LP_ERR_UNSUPPORTED;
    nchunks = (K + 127) / 128;
    kpad = nchunks * 128;
  }
  // row stride = 64 bytes mod 128, so the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups
  int ldx = kpad + 32;
  if ((ldx * 2) % 128 != 64) ldx += 32;
  const int ncols = M * split;
  const size_t smem = (size_t)ncols * ldx * 2 + (W.fmt == LP_W_INT4 ? (size_t)nchunks * 8 * 4 : 0) + MMA_WARPS * 128 * 4 + 8 * 16 * 4;
  if (smem > MMA_MAX_SMEM) return LP_ERR_UNSUPPORTED;
  dim3 grid(W.N / MMA_ROWS), block(MMA_THREADS);
  if (W.fmt == LP_W_BF16)
    return launch(linear_mma_kernel<LP_W_BF16>, grid, block, smem, stream, x, M, split, W, nrm, epi, residual, out, round_bf16, ldx, nchunks);
  return launch(linear_mma_kernel<LP_W_INT4>, grid, block, smem, stream, x, M, split, W, nrm, epi, residual, out, round_bf16, ldx, nchunks);
}

int linear_mma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16, void* stream) {
  NormArgs none = {nullptr, nullptr, 0.f, -1};
  return linear_mma_norm(x, M, W, none, epi, residual, out, round_bf16, stream);
}

}  // namespace lp
