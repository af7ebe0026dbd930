// Embedding gather, LayerNorm / RMSNorm, fused RoPE + KV-cache append, GPTQ int4 repack.
#include "common.cuh"

namespace lp {

// ---------------------------------------------------------------------------------------------
// embedding gather: out[r, :] = wte[idx[r], :]        (reference: model.py:99)
// ---------------------------------------------------------------------------------------------
template <typename WT>
__global__ void embed_kernel(const void* __restrict__ idx, int idx64, const int* __restrict__ idx_offset,
                             const WT* __restrict__ wte, float* __restrict__ out, int E, int round_bf16) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x + (idx_offset ? *idx_offset : 0);
  const long long tok = idx64 ? reinterpret_cast<const long long*>(idx)[r] : (long long)reinterpret_cast<const int*>(idx)[r];
  const WT* src = wte + (size_t)tok * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float v;
    if constexpr (sizeof(WT) == 2) v = __bfloat162float(src[e]);
    else v = src[e];
    out[(size_t)blockIdx.x * E + e] = maybe_round(v, round_bf16);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics; torch.nn.LayerNorm) and RMSNorm (lit_gpt/rmsnorm.py:17-21)
// one CTA per row
// ---------------------------------------------------------------------------------------------
constexpr int NORM_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous use
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (NORM_THREADS >> 5)) ? red[l] : 0.f;
  return warp_sum(t);
}

__global__ void __launch_bounds__(NORM_THREADS) norm_kernel(int kind, const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float eps, float* __restrict__ y, int E,
                                                            int round_bf16) {
  __shared__ float red[32];
  pdl_wait();
  pdl_launch_dependents();
  const float* xr = x + (size_t)blockIdx.x * E;
  float* yr = y + (size_t)blockIdx.x * E;
  if (kind == LP_NORM_LAYERNORM) {
    float s = 0.f;
    for (int e = threadIdx.x; e < E; e += NORM_THREADS) s += xr[e];
    const float mean = block_sum(s, red) / (float)E;
    float v = 0.f;
    for (int e = threadIdx.x; e < E; e += NORM_THREADS) {
      const float d = xr[e] - mean;
      v += d * d;
    }
    const float rstd = 1.0f / sqrtf(block_sum(v, red) / (float)E + eps);
    for (int e = threadIdx.x; e < E; e += NORM_THREADS) {
      float o = (xr[e] - mean) * rstd * w[e] + (b ? b[e] : 0.f);
      yr[e] = maybe_round(o, round_bf16);
    }
  } else {
    float s = 0.f;
    if (!round_bf16) {
      for (int e = threadIdx.x; e < E; e += NORM_THREADS) s += xr[e] * xr[e];
      const float r = 1.0f / sqrtf(block_sum(s, red) / (float)E + eps);
      for (int e = threadIdx.x; e < E; e += NORM_THREADS) yr[e] = w[e] * (xr[e] * r);
    } else {
      // the reference evaluates every step in bf16: x*x, mean (fp32 accumulate, rounded), +eps, rsqrt, x*r, w*xn
      for (int e = threadIdx.x; e < E; e += NORM_THREADS) s += bf16_round(xr[e] * xr[e]);
      const float ms = bf16_round(block_sum(s, red) / (float)E);
      const float r = bf16_round(1.0f / sqrtf(bf16_round(ms + bf16_round(eps))));
      for (int e = threadIdx.x; e < E; e += NORM_THREADS) yr[e] = bf16_round(w[e] * bf16_round(xr[e] * r));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RoPE + KV append.  grid (B*T, H + 2G), one thread per head dim.
// qkv row layout: for group g: [q_{g,0..qpk-1}, k_g, v_g] x hs   (model.py:210-214)
// rotate-half on the first n_elem dims: out = x*cos + rot*sin with rot = (-x2, x1)   (model.py:330-336)
// ---------------------------------------------------------------------------------------------
template <typename KV>
__global__ void rope_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ cosT, const float* __restrict__ sinT,
                               const int* __restrict__ pos, float* __restrict__ q_out, KV* __restrict__ kc, KV* __restrict__ vc,
                               int T, int H, int G, int hs, int n_elem, int max_seq, int round_bf16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row % T;
  const int qpk = H / G;
  const int p = pos[t];
  // one CTA per token row; its threads walk the (head slot, dim) pairs of the row
  for (int e = threadIdx.x; e < (H + 2 * G) * hs; e += blockDim.x) {
  const int slot_in_row = e / hs;  // 0 .. H+2G-1, in qkv row order
  const int d = e % hs;
  const int g = slot_in_row / (qpk + 2), j = slot_in_row % (qpk + 2);
  const float* src = qkv + (size_t)row * (H + 2 * G) * hs + (size_t)slot_in_row * hs;
  float v = src[d];
  if (j <= qpk && d < n_elem) {  // q heads and the k head are rotated, v is not
    const int half = n_elem >> 1;
    const float partner = (d < half) ? -src[d + half] : src[d - half];
    const float c = cosT[(size_t)p * n_elem + d], s = sinT[(size_t)p * n_elem + d];
    v = __fadd_rn(__fmul_rn(v, c), __fmul_rn(partner, s));
    v = maybe_round(v, round_bf16);
  }
  if (j < qpk) {
    q_out[(size_t)row * H * hs + (size_t)(g * qpk + j) * hs + d] = v;
  } else {
    const int slot = p % max_seq;
    KV* dst = (j == qpk ? kc : vc) + (((size_t)b * G + g) * max_seq + slot) * hs + d;
    if constexpr (sizeof(KV) == 2) *dst = __float2bfloat16_rn(v);
    else *dst = v;
  }
  }
}

// The same, four dims per thread (hs % 4 == 0, n_elem % 8 == 0: the rotation partner of an aligned group of four is an aligned
// group of four): 16-byte loads / stores, one index division per four elements.  Same arithmetic per element (bit-identical);
// the scalar kernel above took 44 us per falcon-7b layer at T = 1792 (4 % of the prefill) for 70 MB of traffic.
template <typename KV>
__global__ void __launch_bounds__(256) rope_kv_vec4_kernel(const float* __restrict__ qkv, const float* __restrict__ cosT,
                                                           const float* __restrict__ sinT, const int* __restrict__ pos,
                                                           float* __restrict__ q_out, KV* __restrict__ kc, KV* __restrict__ vc, int T, int H,
                                                           int G, int hs, int n_elem, int max_seq, int round_bf16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;
  const int b = row / T, t = row % T;
  const int qpk = H / G;
  const int p = pos[t];
  const int hs4 = hs >> 2, half = n_elem >> 1;
  const float* rowp = qkv + (size_t)row * (H + 2 * G) * hs;
  for (int e = threadIdx.x; e < (H + 2 * G) * hs4; e += blockDim.x) {
    const int slot_in_row = e / hs4, d = (e - slot_in_row * hs4) << 2;
    const int g = slot_in_row / (qpk + 2), j = slot_in_row - g * (qpk + 2);
    const float* src = rowp + (size_t)slot_in_row * hs;
    float4 v = *reinterpret_cast<const float4*>(src + d);
    if (j <= qpk && d < n_elem) {  // q heads and the k head are rotated, v is not
      const bool lo = d < half;
      float4 pr = *reinterpret_cast<const float4*>(src + (lo ? d + half : d - half));
      if (lo) pr = make_float4(-pr.x, -pr.y, -pr.z, -pr.w);
      const float4 c = *reinterpret_cast<const float4*>(cosT + (size_t)p * n_elem + d);
      const float4 s = *reinterpret_cast<const float4*>(sinT + (size_t)p * n_elem + d);
      v.x = maybe_round(__fadd_rn(__fmul_rn(v.x, c.x), __fmul_rn(pr.x, s.x)), round_bf16);
      v.y = maybe_round(__fadd_rn(__fmul_rn(v.y, c.y), __fmul_rn(pr.y, s.y)), round_bf16);
      v.z = maybe_round(__fadd_rn(__fmul_rn(v.z, c.z), __fmul_rn(pr.z, s.z)), round_bf16);
      v.w = maybe_round(__fadd_rn(__fmul_rn(v.w, c.w), __fmul_rn(pr.w, s.w)), round_bf16);
    }
    if (j < qpk) {
      *reinterpret_cast<float4*>(q_out + (size_t)row * H * hs + (size_t)(g * qpk + j) * hs + d) = v;
    } else {
      const int slot = p % max_seq;
      KV* dst = (j == qpk ? kc : vc) + (((size_t)b * G + g) * max_seq + slot) * hs + d;
      if constexpr (sizeof(KV) == 2) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), c2 = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&a);
        pk.y = *reinterpret_cast<const uint32_t*>(&c2);
        *reinterpret_cast<uint2*>(dst) = pk;
      } else {
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GPTQ storage (uint8 (N, K/2), strides (1, N)) -> row-major [N, Kp/2], zero padded
// ---------------------------------------------------------------------------------------------
__global__ void repack_int4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int N, int kb, int row_bytes) {
  __shared__ uint8_t tile[32][33];
  const int n0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  // read: consecutive threads walk n (contiguous in the source)
  {
    const int n = n0 + threadIdx.x;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
      const int j = j0 + jj;
      tile[jj][threadIdx.x] = (n < N && j < kb) ? src[(size_t)j * N + n] : 0;
    }
  }
  __syncthreads();
  {
    const int j = j0 + threadIdx.x;
    for (int nn = threadIdx.y; nn < 32; nn += blockDim.y) {
      const int n = n0 + nn;
      if (n < N && j < row_bytes) dst[(size_t)n * row_bytes + j] = tile[threadIdx.x][nn];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// range check of token ids and positions on the device (nn.Embedding / index_select raise in the reference, model.py:88-99)
// ---------------------------------------------------------------------------------------------
__global__ void validate_inputs_kernel(void* idx, int idx64, int n_idx, int vocab, int* pos, int n_pos, int block_size, int* flag) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += gridDim.x * blockDim.x) {
    if (idx64) {
      long long& t = reinterpret_cast<long long*>(idx)[i];
      if (t < 0 || t >= vocab) { bad |= 1; t = t < 0 ? 0 : vocab - 1; }
    } else {
      int& t = reinterpret_cast<int*>(idx)[i];
      if (t < 0 || t >= vocab) { bad |= 1; t = t < 0 ? 0 : vocab - 1; }
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += gridDim.x * blockDim.x) {
    int& q = pos[i];
    if (q < 0 || q >= block_size) { bad |= 2; q = q < 0 ? 0 : block_size - 1; }
  }
  if (bad) atomicOr(flag, bad);
}

// The same check fused with the copy of a step's inputs into the static buffers a captured graph reads: idx (int32 / int64) ->
// int64 [n_idx], pos (int32 / int64) -> int32 [n_pos].  One launch instead of two device copies and a check.
__global__ void stage_inputs_kernel(const void* __restrict__ idx_src, int idx64, int n_idx, const void* __restrict__ pos_src, int pos64,
                                    int n_pos, long long* __restrict__ idx_dst, int* __restrict__ pos_dst, int vocab, int block_size,
                                    int* flag) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += gridDim.x * blockDim.x) {
    long long t = idx64 ? reinterpret_cast<const long long*>(idx_src)[i] : (long long)reinterpret_cast<const int*>(idx_src)[i];
    if (t < 0 || t >= vocab) { bad |= 1; t = t < 0 ? 0 : vocab - 1; }
    idx_dst[i] = t;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += gridDim.x * blockDim.x) {
    long long q = pos64 ? reinterpret_cast<const long long*>(pos_src)[i] : (long long)reinterpret_cast<const int*>(pos_src)[i];
    if (q < 0 || q >= block_size) { bad |= 2; q = q < 0 ? 0 : block_size - 1; }
    pos_dst[i] = (int)q;
  }
  if (bad) atomicOr(flag, bad);
}

}  // namespace lp

extern "C" {

int lp_stage_inputs(const void* idx_src, int idx_is_int64, int n_idx, const void* pos_src, int pos_is_int64, int n_pos, int64_t* idx_dst,
                    int32_t* pos_dst, int vocab, int block_size, int32_t* flag, void* stream) {
  if (!flag || vocab <= 0 || block_size <= 0 || n_idx <= 0 || n_pos <= 0 || !idx_src || !pos_src || !idx_dst || !pos_dst)
    return LP_ERR_INVALID_ARG;
  const int n = n_idx > n_pos ? n_idx : n_pos;
  const int grid = (n + 255) / 256 < 64 ? (n + 255) / 256 : 64;
  lp::stage_inputs_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(idx_src, idx_is_int64, n_idx, pos_src, pos_is_int64, n_pos,
                                                                                    reinterpret_cast<long long*>(idx_dst), pos_dst, vocab,
                                                                                    block_size, flag);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

int lp_validate_inputs(void* idx, int idx_is_int64, int n_idx, int vocab, int32_t* pos, int n_pos, int block_size, int32_t* flag,
                       void* stream) {
  if (!flag || vocab <= 0 || block_size <= 0 || n_idx < 0 || n_pos < 0 || (n_idx && !idx) || (n_pos && !pos)) return LP_ERR_INVALID_ARG;
  const int n = n_idx > n_pos ? n_idx : n_pos;
  if (n == 0) return LP_OK;
  const int grid = (n + 255) / 256 < 64 ? (n + 255) / 256 : 64;
  lp::validate_inputs_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(idx, idx_is_int64, n_idx, vocab, pos, n_pos,
                                                                                       block_size, flag);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

int lp_embed(const void* idx, int idx_is_int64, const int32_t* idx_offset, const void* wte, int wte_dtype, float* out, int rows,
             int E, int round_bf16, void* stream) {
  if (!idx || !wte || !out || rows <= 0 || E <= 0) return LP_ERR_INVALID_ARG;
  if (wte_dtype == LP_F32)
    return lp::launch(lp::embed_kernel<float>, dim3(rows), dim3(256), 0, stream, idx, idx_is_int64, idx_offset, (const float*)wte, out, E,
                      round_bf16);
  if (wte_dtype == LP_BF16)
    return lp::launch(lp::embed_kernel<__nv_bfloat16>, dim3(rows), dim3(256), 0, stream, idx, idx_is_int64, idx_offset,
                      (const __nv_bfloat16*)wte, out, E, round_bf16);
  return LP_ERR_INVALID_ARG;
}

int lp_norm(int kind, const float* x, const float* weight, const float* bias, float eps, float* y, int rows, int E,
            int round_bf16, void* stream) {
  if (!x || !weight || !y || rows <= 0 || E <= 0) return LP_ERR_INVALID_ARG;
  if (kind != LP_NORM_LAYERNORM && kind != LP_NORM_RMS) return LP_ERR_INVALID_ARG;
  return lp::launch(lp::norm_kernel, dim3(rows), dim3(lp::NORM_THREADS), 0, stream, kind, x, weight, bias, eps, y, E, round_bf16);
}

int lp_rope_kv_append(const float* qkv, const float* cos, const float* sin, const int32_t* pos, float* q_out, void* k_cache,
                      void* v_cache, int kv_dtype, int B, int T, int H, int G, int hs, int n_elem, int max_seq, int round_bf16,
                      void* stream) {
  if (!qkv || !pos || !q_out || !k_cache || !v_cache) return LP_ERR_INVALID_ARG;
  if (n_elem > 0 && (!cos || !sin)) return LP_ERR_INVALID_ARG;
  if (B <= 0 || T <= 0 || H <= 0 || G <= 0 || H % G || hs <= 0 || hs > 1024 || n_elem < 0 || n_elem > hs || ((n_elem & 1) && n_elem != 1) || max_seq <= 0)
    return LP_ERR_INVALID_ARG;
  dim3 grid(B * T), block(256);
  const bool vec4 = hs % 4 == 0 && n_elem % 8 == 0 && !((uintptr_t)qkv & 15) && !((uintptr_t)q_out & 15) && !((uintptr_t)k_cache & 15) &&
                    !((uintptr_t)v_cache & 15) && (n_elem == 0 || (!((uintptr_t)cos & 15) && !((uintptr_t)sin & 15)));
  if (vec4 && kv_dtype == LP_F32)
    return lp::launch(lp::rope_kv_vec4_kernel<float>, grid, block, 0, stream, qkv, cos, sin, pos, q_out, (float*)k_cache, (float*)v_cache,
                      T, H, G, hs, n_elem, max_seq, round_bf16);
  if (vec4 && kv_dtype == LP_BF16)
    return lp::launch(lp::rope_kv_vec4_kernel<__nv_bfloat16>, grid, block, 0, stream, qkv, cos, sin, pos, q_out,
                      (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, T, H, G, hs, n_elem, max_seq, round_bf16);
  if (kv_dtype == LP_F32)
    return lp::launch(lp::rope_kv_kernel<float>, grid, block, 0, stream, qkv, cos, sin, pos, q_out, (float*)k_cache, (float*)v_cache,
                      T, H, G, hs, n_elem, max_seq, round_bf16);
  if (kv_dtype == LP_BF16)
    return lp::launch(lp::rope_kv_kernel<__nv_bfloat16>, grid, block, 0, stream, qkv, cos, sin, pos, q_out,
                      (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, T, H, G, hs, n_elem, max_seq, round_bf16);
  return LP_ERR_INVALID_ARG;
}

size_t lp_int4_row_bytes(int K) { return (size_t)((K + 255) / 256 * 256) / 2; }

int lp_repack_gptq_int4(const uint8_t* src, uint8_t* dst, int N, int K, void* stream) {
  if (!src || !dst || N <= 0 || K <= 0 || (K & 1)) return LP_ERR_INVALID_ARG;
  const int row_bytes = (int)lp_int4_row_bytes(K);
  dim3 grid((N + 31) / 32, (row_bytes + 31) / 32), block(32, 8);
  lp::repack_int4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, N, K / 2, row_bytes);
  LP_CUDA_TRY(cudaGetLastError());
  return LP_OK;
}

}  // extern "C"
