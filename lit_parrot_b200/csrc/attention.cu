// Attention of query rows against the KV cache (decode, and small-T prefill through the same path).
//
// grid (rows = B*T, G, n_splits).  One CTA owns ALL q heads of one KV group for one block of the
// sequence, so K/V are read once per group (MHA: 1 head, GQA 70B: 8 heads, Falcon MQA: 71 heads).
// Keys are staged through shared memory in 64-key tiles; each warp runs an online softmax with warp
// shuffles (lane = key for q.k, lane = head-dim for p.v); split partials (m, l, o[hs]) go to a
// workspace and a small combine kernel merges them.  Only min(pos+1, max_seq) keys are read: there
// is no mask tensor (the reference attends a masked, zero-filled max_seq-long cache, model.py:91-92,
// 130-144, 273-275).
#include <math_constants.h>

#include "common.cuh"

namespace lp {

constexpr int ATT_THREADS = 128;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_TILE = 64;     // keys per shared-memory tile
constexpr int ATT_MAX_DPL = 8;   // head dims per lane -> hs <= 256

struct AttnSmem {
  // [ATT_TILE][hs + 4] K, same for V, then q [qpk][hs], then merge scratch
  float* k;
  float* v;
  float* q;
  float* merge;  // [qpk][WT][hs + 2]
};

__device__ __forceinline__ float kv_to_float(float v) { return v; }
__device__ __forceinline__ float kv_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename KV>
__global__ void __launch_bounds__(ATT_THREADS)
attn_split_kernel(const float* __restrict__ q, const KV* __restrict__ kc, const KV* __restrict__ vc, const int* __restrict__ pos,
                  float* __restrict__ out, float* __restrict__ ws, int T, int H, int G, int hs, int max_seq, float scale,
                  int n_splits, int round_bf16) {
  extern __shared__ __align__(16) float smem[];
  const int row = blockIdx.x, g = blockIdx.y, split = blockIdx.z;
  const int b = row / T, t = row % T;
  const int qpk = H / G;
  const int ldk = hs + 4;
  const bool vec4 = (hs & 3) == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps split heads first (WH), then key tiles (WT)
  const int WH = qpk >= 4 ? 4 : (qpk >= 2 ? 2 : 1);
  const int WT = ATT_WARPS / WH;
  const int wh = warp % WH, wt = warp / WH;
  const int heads_per_warp = (qpk + WH - 1) / WH;

  float* sk = smem;
  float* sv = sk + ATT_TILE * ldk;
  float* sq = sv + ATT_TILE * ldk;
  float* sm = sq + qpk * hs;  // merge scratch [qpk][WT][hs+2]

  pdl_wait();
  pdl_launch_dependents();

  int kv_len = pos[t] + 1;
  if (kv_len > max_seq) kv_len = max_seq;
  // keys of this split: contiguous block, multiple of the tile size
  int chunk = (kv_len + n_splits - 1) / n_splits;
  chunk = (chunk + ATT_TILE - 1) / ATT_TILE * ATT_TILE;
  const int k_begin = split * chunk;
  const int k_end = min(kv_len, k_begin + chunk);

  // stage q (pre-scaled) for all heads of the group
  for (int i = threadIdx.x; i < qpk * hs; i += ATT_THREADS)
    sq[i] = q[(size_t)row * H * hs + (size_t)g * qpk * hs + i] * scale;

  const int dpl = (hs + 31) / 32;
  // per-warp running state for the heads it owns is kept in registers one head at a time:
  // loop heads outermost, tiles innermost would re-stage K/V per head, so instead keep state in smem-free
  // registers for up to HPW heads by iterating heads inside the tile loop and spilling state to `sm`.
  // State layout in sm: [head][wt][0..hs-1]=o, [hs]=m, [hs+1]=l
  const int ldm = hs + 2;
  for (int hh = 0; hh < heads_per_warp; ++hh) {
    const int h = wh + hh * WH;
    if (h < qpk) {
      float* st = sm + ((size_t)h * WT + wt) * ldm;
      for (int d = lane; d < hs; d += 32) st[d] = 0.f;
      if (lane == 0) {
        st[hs] = -CUDART_INF_F;
        st[hs + 1] = 0.f;
      }
    }
  }

  const KV* kbase = kc + ((size_t)b * G + g) * max_seq * hs;
  const KV* vbase = vc + ((size_t)b * G + g) * max_seq * hs;

  for (int tile0 = k_begin; tile0 < k_end; tile0 += ATT_TILE) {
    const int nk = min(ATT_TILE, k_end - tile0);
    __syncthreads();  // previous tile fully consumed (also orders the sq / sm initialisation)
    // cooperative, coalesced copy of nk x hs K and V elements into padded fp32 tiles
    if (!vec4) {  // odd head sizes (the reference's unit tests use hs = 2 and 4): scalar staging
      for (int i = threadIdx.x; i < nk * hs; i += ATT_THREADS) {
        const int kk = i / hs, d = i % hs;
        sk[kk * ldk + d] = kv_to_float(kbase[(size_t)(tile0 + kk) * hs + d]);
        sv[kk * ldk + d] = kv_to_float(vbase[(size_t)(tile0 + kk) * hs + d]);
      }
    }
    for (int i = threadIdx.x * 4; vec4 && i < nk * hs; i += ATT_THREADS * 4) {
      const int kk = i / hs, d = i % hs;  // hs % 4 == 0, so a 4-vector never crosses a key
      const KV* ks = kbase + (size_t)(tile0 + kk) * hs + d;
      const KV* vs = vbase + (size_t)(tile0 + kk) * hs + d;
      float4 kf, vf;
      if constexpr (sizeof(KV) == 2) {
        const uint2 ku = *reinterpret_cast<const uint2*>(ks);
        const uint2 vu = *reinterpret_cast<const uint2*>(vs);
        kf = make_float4(bf16lo(ku.x), bf16hi(ku.x), bf16lo(ku.y), bf16hi(ku.y));
        vf = make_float4(bf16lo(vu.x), bf16hi(vu.x), bf16lo(vu.y), bf16hi(vu.y));
      } else {
        kf = *reinterpret_cast<const float4*>(ks);
        vf = *reinterpret_cast<const float4*>(vs);
      }
      *reinterpret_cast<float4*>(sk + kk * ldk + d) = kf;
      *reinterpret_cast<float4*>(sv + kk * ldk + d) = vf;
    }
    __syncthreads();

    // this warp's 32-key sub-tiles within the tile: sub-tile index s with s % WT == wt
    for (int s = wt; s * 32 < nk; s += WT) {
      const int key = s * 32 + lane;
      const bool valid = key < nk;
      const float* krow = sk + (valid ? key : 0) * ldk;
      for (int hh = 0; hh < heads_per_warp; ++hh) {
        const int h = wh + hh * WH;
        if (h >= qpk) break;
        const float* qh = sq + h * hs;
        float sc = 0.f;
        if (!vec4)
          for (int d = 0; d < hs; ++d) sc = fmaf(qh[d], krow[d], sc);
        for (int d = 0; vec4 && d < hs; d += 4) {
          const float4 kv4 = *reinterpret_cast<const float4*>(krow + d);
          const float4 q4 = *reinterpret_cast<const float4*>(qh + d);
          sc = fmaf(q4.x, kv4.x, sc);
          sc = fmaf(q4.y, kv4.y, sc);
          sc = fmaf(q4.z, kv4.z, sc);
          sc = fmaf(q4.w, kv4.w, sc);
        }
        if (!valid) sc = -CUDART_INF_F;
        float* st = sm + ((size_t)h * WT + wt) * ldm;
        const float m_old = st[hs], l_old = st[hs + 1];
        const float m_new = fmaxf(m_old, warp_max(sc));
        const float p = valid ? expf(sc - m_new) : 0.f;
        const float corr = (m_old == -CUDART_INF_F) ? 0.f : expf(m_old - m_new);
        const float l_new = l_old * corr + warp_sum(p);
        float o[ATT_MAX_DPL];
#pragma unroll
        for (int i = 0; i < ATT_MAX_DPL; ++i) {
          const int d = lane + 32 * i;
          o[i] = (i < dpl && d < hs) ? st[d] * corr : 0.f;
        }
        const int nvalid = min(32, nk - s * 32);
        for (int j = 0; j < nvalid; ++j) {
          const float pj = __shfl_sync(0xffffffffu, p, j);
          const float* vrow = sv + (s * 32 + j) * ldk;
#pragma unroll
          for (int i = 0; i < ATT_MAX_DPL; ++i) {
            const int d = lane + 32 * i;
            if (i < dpl && d < hs) o[i] = fmaf(pj, vrow[d], o[i]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < ATT_MAX_DPL; ++i) {
          const int d = lane + 32 * i;
          if (i < dpl && d < hs) st[d] = o[i];
        }
        if (lane == 0) {
          st[hs] = m_new;
          st[hs + 1] = l_new;
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // merge the WT partial states of each head and emit either the final output (n_splits == 1) or a partial
  for (int h = warp; h < qpk; h += ATT_WARPS) {
    const float* st0 = sm + (size_t)h * WT * ldm;
    float m = -CUDART_INF_F;
    for (int w = 0; w < WT; ++w) m = fmaxf(m, st0[w * ldm + hs]);
    float l = 0.f;
    float o[ATT_MAX_DPL];
#pragma unroll
    for (int i = 0; i < ATT_MAX_DPL; ++i) o[i] = 0.f;
    for (int w = 0; w < WT; ++w) {
      const float mw = st0[w * ldm + hs];
      const float c = (mw == -CUDART_INF_F) ? 0.f : expf(mw - m);
      l += st0[w * ldm + hs + 1] * c;
#pragma unroll
      for (int i = 0; i < ATT_MAX_DPL; ++i) {
        const int d = lane + 32 * i;
        if (i < dpl && d < hs) o[i] = fmaf(st0[w * ldm + d], c, o[i]);
      }
    }
    const int head = g * qpk + h;
    if (n_splits == 1) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int i = 0; i < ATT_MAX_DPL; ++i) {
        const int d = lane + 32 * i;
        if (i < dpl && d < hs) out[(size_t)row * H * hs + (size_t)head * hs + d] = maybe_round(o[i] * inv, round_bf16);
      }
    } else {
      float* p = ws + (((size_t)row * H + head) * n_splits + split) * ldm;
#pragma unroll
      for (int i = 0; i < ATT_MAX_DPL; ++i) {
        const int d = lane + 32 * i;
        if (i < dpl && d < hs) p[d] = o[i];
      }
      if (lane == 0) {
        p[hs] = m;
        p[hs + 1] = l;
      }
    }
  }
}

// grid (rows, H), one warp per (row, head)
__global__ void attn_combine_kernel(const float* __restrict__ ws, float* __restrict__ out, int H, int hs, int n_splits,
                                    int round_bf16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x, head = blockIdx.y, lane = threadIdx.x;
  const int ldm = hs + 2;
  const float* p = ws + (((size_t)row * H + head) * n_splits) * ldm;
  float m = -CUDART_INF_F;
  for (int s = 0; s < n_splits; ++s) m = fmaxf(m, p[s * ldm + hs]);
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float ms = p[s * ldm + hs];
    l += (ms == -CUDART_INF_F) ? 0.f : p[s * ldm + hs + 1] * expf(ms - m);
  }
  const float inv = 1.0f / l;
  for (int d = lane; d < hs; d += 32) {
    float o = 0.f;
    for (int s = 0; s < n_splits; ++s) {
      const float ms = p[s * ldm + hs];
      if (ms != -CUDART_INF_F) o = fmaf(p[s * ldm + d], expf(ms - m), o);
    }
    out[(size_t)row * H * hs + (size_t)head * hs + d] = maybe_round(o * inv, round_bf16);
  }
}

static size_t attn_smem_bytes(int qpk, int hs) {
  const int WH = qpk >= 4 ? 4 : (qpk >= 2 ? 2 : 1);
  const int WT = ATT_WARPS / WH;
  return sizeof(float) * ((size_t)2 * ATT_TILE * (hs + 4) + (size_t)qpk * hs + (size_t)qpk * WT * (hs + 2));
}

static int attn_splits(int rows, int G, int max_seq) {
  const int max_splits = (max_seq + ATT_TILE - 1) / ATT_TILE;
  int want = (2 * num_sms() + rows * G - 1) / (rows * G);
  if (want < 1) want = 1;
  if (want > max_splits) want = max_splits;
  if (want > 64) want = 64;
  return want;
}

int init_attention() {
  LP_CUDA_TRY(cudaFuncSetAttribute(attn_split_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  LP_CUDA_TRY(cudaFuncSetAttribute(attn_split_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return LP_OK;
}

}  // namespace lp

extern "C" {

size_t lp_attn_workspace_bytes(int B, int T, int H, int hs, int max_seq) {
  const int max_splits = (max_seq + lp::ATT_TILE - 1) / lp::ATT_TILE;
  const int s = max_splits > 64 ? 64 : max_splits;
  return sizeof(float) * (size_t)B * T * H * s * (hs + 2);
}

int lp_attn_decode(const float* q, const void* k_cache, const void* v_cache, int kv_dtype, const int32_t* pos, float* out,
                   void* workspace, size_t workspace_bytes, int B, int T, int H, int G, int hs, int max_seq, float scale,
                   int round_bf16, void* stream) {
  if (!q || !k_cache || !v_cache || !pos || !out) return LP_ERR_INVALID_ARG;
  if (B <= 0 || T <= 0 || H <= 0 || G <= 0 || H % G || hs <= 0 || max_seq <= 0) return LP_ERR_INVALID_ARG;
  if (hs > 32 * lp::ATT_MAX_DPL) return LP_ERR_UNSUPPORTED;
  const int rows = B * T, qpk = H / G;
  const size_t smem = lp::attn_smem_bytes(qpk, hs);
  if (smem > 200 * 1024) return LP_ERR_UNSUPPORTED;
  const int n_splits = lp::attn_splits(rows, G, max_seq);
  if (n_splits > 1) {
    if (!workspace) return LP_ERR_INVALID_ARG;
    if (workspace_bytes < sizeof(float) * (size_t)rows * H * n_splits * (hs + 2)) return LP_ERR_WORKSPACE;
  }
  dim3 grid(rows, G, n_splits), block(lp::ATT_THREADS);
  int rc;
  if (kv_dtype == LP_F32)
    rc = lp::launch(lp::attn_split_kernel<float>, grid, block, smem, stream, q, (const float*)k_cache, (const float*)v_cache, pos,
                    out, (float*)workspace, T, H, G, hs, max_seq, scale, n_splits, round_bf16);
  else if (kv_dtype == LP_BF16)
    rc = lp::launch(lp::attn_split_kernel<__nv_bfloat16>, grid, block, smem, stream, q, (const __nv_bfloat16*)k_cache,
                    (const __nv_bfloat16*)v_cache, pos, out, (float*)workspace, T, H, G, hs, max_seq, scale, n_splits, round_bf16);
  else
    return LP_ERR_INVALID_ARG;
  if (rc != LP_OK || n_splits == 1) return rc;
  return lp::launch(lp::attn_combine_kernel, dim3(rows, H), dim3(32), 0, stream, (const float*)workspace, out, H, hs, n_splits,
                    round_bf16);
}

}  // extern "C"
