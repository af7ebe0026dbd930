// Adapter / LoRA inference support: what the fork's own inference scripts run on top of the base model.
//
//   lp_adapter_attn — the gated attention over the adaption prompt of LLaMA-Adapter (lit_gpt/adapter.py:234-254):
//       y = y + gating_factor * SDPA(q, ak, av, mask = all ones)
//     with q the RoPE'd queries of CausalSelfAttention.forward (adapter.py:207-212) and ak / av the k / v parts of
//     attn.attn(adapter_wte.weight) — input independent, cached by the reference as `adapter_kv_cache` and by the host side here.
//     aT is 10 keys: one warp per (token row, head), q rotated in registers, scores / softmax / P.V with warp shuffles.  The
//     work is ~H * aT * hs FMAs per token next to the layer's weight streaming: a CUDA-core kernel, launched right behind the
//     causal attention kernel (PDL), adding into its output in place.
//   lp_lora_merge — `weight.data += (lora_B @ lora_A) * scaling` (lit_gpt/lora.py:154-164) and the scattered form of
//     LoRAQKVLinear (lora.py:296-361: the update only touches the rows listed in lora_ind).  Load time.
#include <math_constants.h>

#include "common.cuh"

namespace lp {

constexpr int AD_MAX_HS = 256;  // head dims handled: 8 per lane
constexpr int AD_MAX_T = 64;    // adaption prompt length handled (reference default: 10)

// grid (B*T, ceil(H / 4)), block 128: warp w handles head 4 blockIdx.y + w of token row blockIdx.x.
__global__ void __launch_bounds__(128) adapter_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ cosT,
                                                           const float* __restrict__ sinT, const int* __restrict__ pos,
                                                           const float* __restrict__ ak, const float* __restrict__ av,
                                                           const float* __restrict__ gating, float* __restrict__ out, int T, int H, int G,
                                                           int hs, int n_elem, int aT, float scale, int round_bf16) {
  pdl_wait();  // qkv and the causal attention output come from the kernels before
  pdl_launch_dependents();
  const int row = blockIdx.x, t = row % T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y * 4 + warp;
  if (h >= H) return;
  const int qpk = H / G, g = h / qpk, j = h % qpk;
  const float* src = qkv + (size_t)row * (H + 2 * G) * hs + (size_t)(g * (qpk + 2) + j) * hs;
  const int p = pos[t];
  const int half = n_elem >> 1;
  // rotated q (same arithmetic and rounding points as lp_rope_kv_append), dims lane, lane + 32, ...
  float q[AD_MAX_HS / 32];
#pragma unroll
  for (int i = 0; i < AD_MAX_HS / 32; ++i) {
    const int d = lane + 32 * i;
    float v = 0.f;
    if (d < hs) {
      v = src[d];
      if (d < n_elem) {
        const float partner = (d < half) ? -src[d + half] : src[d - half];
        const float c = cosT[(size_t)p * n_elem + d], s = sinT[(size_t)p * n_elem + d];
        v = maybe_round(__fadd_rn(__fmul_rn(v, c), __fmul_rn(partner, s)), round_bf16);
      }
    }
    q[i] = v;
  }
  const float* kg = ak + (size_t)g * aT * hs;
  const float* vg = av + (size_t)g * aT * hs;
  // scores of the aT prefix keys: lane k keeps score k (and k + 32)
  float s0 = -CUDART_INF_F, s1 = -CUDART_INF_F;
  for (int k = 0; k < aT; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < AD_MAX_HS / 32; ++i) {
      const int d = lane + 32 * i;
      if (d < hs) acc = fmaf(q[i], kg[(size_t)k * hs + d], acc);
    }
    acc = warp_sum(acc) * scale;
    if ((k & 31) == lane) {
      if (k < 32) s0 = acc;
      else s1 = acc;
    }
  }
  const float m = warp_max(fmaxf(s0, s1));
  const float e0 = (lane < aT) ? expf(s0 - m) : 0.f, e1 = (lane + 32 < aT) ? expf(s1 - m) : 0.f;
  const float inv = 1.0f / warp_sum(e0 + e1);
  const float gate = gating[h];
  float o[AD_MAX_HS / 32];
#pragma unroll
  for (int i = 0; i < AD_MAX_HS / 32; ++i) o[i] = 0.f;
  for (int k = 0; k < aT; ++k) {
    const float pk = __shfl_sync(0xffffffffu, k < 32 ? e0 : e1, k & 31) * inv;
#pragma unroll
    for (int i = 0; i < AD_MAX_HS / 32; ++i) {
      const int d = lane + 32 * i;
      if (d < hs) o[i] = fmaf(pk, vg[(size_t)k * hs + d], o[i]);
    }
  }
  float* dst = out + (size_t)row * H * hs + (size_t)h * hs;
#pragma unroll
  for (int i = 0; i < AD_MAX_HS / 32; ++i) {
    const int d = lane + 32 * i;
    if (d < hs) {
      // bf16-true reference: ay is a bf16 tensor, gating * ay is rounded, the sum is rounded (adapter.py:254)
      const float ay = maybe_round(o[i], round_bf16);
      dst[d] = maybe_round(dst[d] + maybe_round(gate * ay, round_bf16), round_bf16);
    }
  }
}

// One thread per (row, 4 columns): delta = sum_r B[i, r] * A[r, k], accumulated in r order (r <= 64), W += scaling * delta.
template <typename WT>
__global__ void __launch_bounds__(256) lora_merge_kernel(WT* __restrict__ W, int K, const float* __restrict__ Brows, const float* __restrict__ A,
                                                         int r, const int* __restrict__ rows, int n_rows, float scaling) {
  const int i = blockIdx.x;
  const int k = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (i >= n_rows || k >= K) return;
  const int n = rows ? rows[i] : i;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  for (int rr = 0; rr < r; ++rr) {
    const float b = Brows[(size_t)i * r + rr];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (k + c < K) d[c] = fmaf(b, A[(size_t)rr * K + k + c], d[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (k + c < K) {
      WT* w = W + (size_t)n * K + k + c;
      if constexpr (sizeof(WT) == 2) {
        // reference in bf16: (B @ A) is a bf16 tensor, * scaling is rounded, += rounds once more (lora.py:160-163)
        const float delta = bf16_round(bf16_round(d[c]) * scaling);
        *w = __float2bfloat16_rn(__bfloat162float(*w) + delta);
      } else {
        *w = __fadd_rn(*w, __fmul_rn(d[c], scaling));  // two roundings, like the reference's `delta * scaling` then `+=`
      }
    }
  }
}

}  // namespace lp

extern "C" {

int lp_adapter_attn(const float* qkv, const float* cos, const float* sin, const int32_t* pos, const float* ak, const float* av,
                    const float* gating, float* out, int B, int T, int H, int G, int hs, int n_elem, int aT, float scale,
                    int round_bf16, void* stream) {
  if (!qkv || !pos || !ak || !av || !gating || !out) return LP_ERR_INVALID_ARG;
  if (n_elem > 0 && (!cos || !sin)) return LP_ERR_INVALID_ARG;
  if (B <= 0 || T <= 0 || H <= 0 || G <= 0 || H % G || hs <= 0 || n_elem < 0 || n_elem > hs || ((n_elem & 1) && n_elem != 1) || aT <= 0)
    return LP_ERR_INVALID_ARG;
  if (hs > lp::AD_MAX_HS || aT > lp::AD_MAX_T) return LP_ERR_UNSUPPORTED;
  return lp::launch(lp::adapter_attn_kernel, dim3(B * T, (H + 3) / 4), dim3(128), 0, stream, qkv, cos, sin, pos, ak, av, gating, out, T, H,
                    G, hs, n_elem, aT, scale, round_bf16);
}

int lp_lora_merge(void* W, int w_dtype, int N, int K, const float* B_rows, const float* A, int r, const int32_t* rows, int n_rows,
                  float scaling, void* stream) {
  if (!W || !B_rows || !A || N <= 0 || K <= 0 || r <= 0 || n_rows <= 0 || n_rows > N) return LP_ERR_INVALID_ARG;
  if (w_dtype != LP_F32 && w_dtype != LP_BF16) return LP_ERR_INVALID_ARG;
  dim3 grid(n_rows, (K + 1023) / 1024), block(256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (w_dtype == LP_F32) lp::lora_merge_kernel<float><<<grid, block, 0, st>>>(reinterpret_cast<float*>(W), K, B_rows, A, r, rows, n_rows, scaling);
  else lp::lora_merge_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(W), K, B_rows, A, r, rows, n_rows, scaling);
  LP_CUDA_TRY(cudaGetLastError());
  lp::count_launch();
  return LP_OK;
}

}  // extern "C"
