// Single-token decode attention, fused with RoPE, the KV-cache append and the split-K merge (one launch per layer).
//
//   reference: CausalSelfAttention.forward (model.py:208-249): regroup q/k/v, apply_rope (330-336), cache index_copy_
//   (236-245), scaled_dot_product_attention over the masked max_seq-long cache (256-275).
//
// grid (B*G, n_splits), 128 threads.  One CTA owns ALL q heads of one KV group for one contiguous block of the
// sequence, so K/V are read from HBM exactly once per group (MHA 1 head, GQA-8 for Llama-2-70b, MQA-71 for Falcon).
//   * K/V tiles (64 keys, bf16) stream through a 3-stage cp.async ring in shared memory (16-byte chunks, XOR
//     swizzled so ldmatrix is conflict free); the first tiles are requested BEFORE griddepcontrol.wait, i.e. while
//     the QKV projection of the same layer is still draining (PDL);
//   * q.K^T and P.V run on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate): the heads of the group are the
//     16 rows of the A operand.  To keep fp32-activation accuracy q and P are split into bf16 hi + lo terms (two MMAs
//     each) — the products with the bf16-stored K/V are then exact in the fp32 accumulator;
//   * softmax is the usual online form, row max / row sum with warp shuffles inside the 4-lane quads;
//   * the CTA that owns the new token's slot writes the rotated k and v to the cache (bf16) and patches them into its
//     shared-memory tile, so no other kernel has to run between the QKV projection and attention;
//   * split partials (m, l, o[hs]) go to the workspace; the last CTA of a (batch, group) to arrive (atomic ticket)
//     merges them in split order — deterministic — and resets the ticket for the next launch.
#include <math_constants.h>

#include "common.cuh"

namespace lp {

constexpr int AD_THREADS = 128;
constexpr int AD_TILE = 64;
constexpr int AD_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// hi/lo split of two floats: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  __nv_bfloat162 h;
  h.x = h0;
  h.y = h1;
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_bf16(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
}

struct AttnDecParams {
  const float* qkv;   // [B, (H + 2G) * hs], rows group-interleaved (model.py:210-214)
  const float* cosT;  // [block_size, n_elem]
  const float* sinT;
  const int* pos;     // [1]
  __nv_bfloat16* kc;  // [B, G, max_seq, hs]
  __nv_bfloat16* vc;
  float* out;         // [B, H * hs]
  float* part;        // split partials [(b*H + head) * n_splits + split][hs + 2]
  int* tickets;       // [B * G], zero between launches
  int H, G, n_elem, max_seq, n_splits, round_bf16;
  float scale_log2;   // softmax scale * log2(e)
};

// grid (B*G, n_splits, head tiles): a CTA handles up to 16 q heads of one KV group (MQA-71 -> 5 head tiles that re-read
// the small shared K/V block out of L2); its 4 warps split the keys of every tile.
template <int HS>
__global__ void __launch_bounds__(AD_THREADS)
attn_decode_fused_kernel(const AttnDecParams p) {
  constexpr int WH = 1, MT = 1;
  constexpr int WT = 4 / WH;            // key groups
  constexpr int KPW = AD_TILE / WT;     // keys per warp per tile
  constexpr int NT = KPW / 8;           // score n-tiles per warp
  constexpr int CH = HS / 8;            // 16-byte chunks per key row
  constexpr int LDQ = HS + 8;           // padded q row (bf16 elements): conflict-free A-fragment loads
  constexpr int QROWS = WH * MT * 16;
  constexpr int STAGE_BYTES = AD_TILE * HS * 2;  // one of K or V
  constexpr int LDM = HS + 2;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sKV = smem;                                                      // [STAGES][2][TILE][HS] bf16, swizzled
  __nv_bfloat16* sQh = reinterpret_cast<__nv_bfloat16*>(smem + AD_STAGES * 2 * STAGE_BYTES);  // [QROWS][LDQ]
  __nv_bfloat16* sQl = sQh + QROWS * LDQ;
  __nv_bfloat16* sNew = sQl + QROWS * LDQ;                                        // [2][HS] new k, v
  float* sMerge = reinterpret_cast<float*>(smem);                                 // reuses the ring after the loop
  __shared__ int s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int wh = warp % WH, wt = warp / WH;
  const int b = blockIdx.x / p.G, g = blockIdx.x % p.G;
  const int split = blockIdx.y;
  const int qpk_all = p.H / p.G;
  const int h_first = blockIdx.z * 16;               // first head (inside the group) of this CTA
  const int qpk = min(16, qpk_all - h_first);        // heads of this CTA
  const int ticket_id = blockIdx.x * gridDim.z + blockIdx.z;

  // position of the token being decoded.  Written by the sampler of the previous step, i.e. by a kernel that had
  // completed before the kernel preceding this one passed its own griddepcontrol.wait: safe to read before ours.
  const int pos = p.pos[0];
  const int kv_len = min(pos + 1, p.max_seq);
  const int slot = pos % p.max_seq;
  int chunk = (kv_len + p.n_splits - 1) / p.n_splits;
  chunk = (chunk + AD_TILE - 1) / AD_TILE * AD_TILE;
  const int k_begin = split * chunk;
  const int k_end = min(kv_len, k_begin + chunk);
  const int ntiles = k_end > k_begin ? (k_end - k_begin + AD_TILE - 1) / AD_TILE : 0;
  const bool owner = slot >= k_begin && slot < k_end && blockIdx.z == 0;
  const bool patch = slot >= k_begin && slot < k_end;

  const __nv_bfloat16* kbase = p.kc + ((size_t)b * p.G + g) * p.max_seq * HS;
  const __nv_bfloat16* vbase = p.vc + ((size_t)b * p.G + g) * p.max_seq * HS;
  const uint32_t sKV_u32 = smem_u32(sKV);

  auto issue_tile = [&](int t) {
    if (t < ntiles) {
      const int tile0 = k_begin + t * AD_TILE;
      const uint32_t sk = sKV_u32 + (t % AD_STAGES) * 2 * STAGE_BYTES, sv = sk + STAGE_BYTES;
#pragma unroll
      for (int i = tid; i < AD_TILE * CH; i += AD_THREADS) {
        const int row = i / CH, c = i % CH;
        const int key = tile0 + row;
        const bool valid = key < k_end;
        const size_t off = (size_t)(valid ? key : k_begin) * HS + c * 8;
        const uint32_t d = (row * CH + (c ^ (row & 7))) * 16;
        cp_async16(sk + d, kbase + off, valid ? 16 : 0);  // rows past the end are zero-filled
        cp_async16(sv + d, vbase + off, valid ? 16 : 0);
      }
    }
    cp_async_commit();
  };

  // ---- 1. K/V prefetch, independent of the QKV projection that precedes this kernel -------------------------------
#pragma unroll
  for (int s = 0; s < AD_STAGES - 1; ++s) issue_tile(s);

  pdl_wait();
  pdl_launch_dependents();

  // ---- 2. q (all heads of the group), new k, new v: RoPE, scale, bf16 hi/lo split ----------------------------------
  {
    const float* src0 = p.qkv + (size_t)b * (p.H + 2 * p.G) * HS + (size_t)g * (qpk_all + 2) * HS;
    const int half = p.n_elem >> 1;
    const int nrows = qpk + (patch ? 2 : 0);
    for (int i = tid; i < nrows * HS; i += AD_THREADS) {
      const int j = i / HS, d = i % HS;
      // rows 0 .. qpk-1: this CTA's q heads; then the group's new k and v
      const float* src = src0 + (size_t)(j < qpk ? h_first + j : qpk_all + (j - qpk)) * HS;
      float v = src[d];
      if (j <= qpk && d < p.n_elem) {
        const float partner = (d < half) ? -src[d + half] : src[d - half];
        const float c = p.cosT[(size_t)pos * p.n_elem + d], s = p.sinT[(size_t)pos * p.n_elem + d];
        v = __fadd_rn(__fmul_rn(v, c), __fmul_rn(partner, s));  // same op order as apply_rope (model.py:330-336)
        v = maybe_round(v, p.round_bf16);
      }
      if (j < qpk) {
        const float qs = v * p.scale_log2;
        const __nv_bfloat16 h = __float2bfloat16_rn(qs);
        sQh[j * LDQ + d] = h;
        sQl[j * LDQ + d] = __float2bfloat16_rn(qs - __bfloat162float(h));
      } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        sNew[(j - qpk) * HS + d] = h;
        if (owner) {  // one CTA per (batch, group) appends to the cache
          __nv_bfloat16* dst = (j == qpk ? p.kc : p.vc) + (((size_t)b * p.G + g) * p.max_seq + slot) * HS + d;
          *dst = h;
        }
      }
    }
    for (int i = qpk * LDQ + tid; i < QROWS * LDQ; i += AD_THREADS) {  // unused head rows
      sQh[i] = __float2bfloat16_rn(0.f);
      sQl[i] = __float2bfloat16_rn(0.f);
    }
  }

  // ---- 3. main loop over key tiles ---------------------------------------------------------------------------------
  float O[MT][HS / 8][4];
  float mrow[MT][2], lrow[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    mrow[mt][0] = mrow[mt][1] = -CUDART_INF_F;
    lrow[mt][0] = lrow[mt][1] = 0.f;
#pragma unroll
    for (int dn = 0; dn < HS / 8; ++dn)
#pragma unroll
      for (int c = 0; c < 4; ++c) O[mt][dn][c] = 0.f;
  }
  const int kw0 = wt * KPW;  // this warp's first key inside a tile

  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<AD_STAGES - 2>();
    __syncthreads();  // tile t has landed for everyone; stage (t-1) % STAGES is free (also orders the q staging)
    issue_tile(t + AD_STAGES - 1);
    const int tile0 = k_begin + t * AD_TILE;
    const uint32_t sk = sKV_u32 + (t % AD_STAGES) * 2 * STAGE_BYTES, sv = sk + STAGE_BYTES;
    if (patch && slot >= tile0 && slot < tile0 + AD_TILE) {  // CTA-uniform
      if (tid < 2 * CH) {
        const int which = tid / CH, c = tid % CH, row = slot - tile0;
        const uint4 val = *reinterpret_cast<const uint4*>(sNew + which * HS + c * 8);
        unsigned char* dst = sKV + (t % AD_STAGES) * 2 * STAGE_BYTES + which * STAGE_BYTES + (row * CH + (c ^ (row & 7))) * 16;
        *reinterpret_cast<uint4*>(dst) = val;
      }
      __syncthreads();
    }

    // ---- S = q . K^T for this warp's KPW keys ----
    float S[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) S[mt][nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HS / 16; ks += 2) {
      uint32_t qh[MT][2][4], ql[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int off = ((wh + mt * WH) * 16 + g8) * LDQ + (ks + kk) * 16 + 2 * t4;
          qh[mt][kk][0] = *reinterpret_cast<const uint32_t*>(sQh + off);
          qh[mt][kk][1] = *reinterpret_cast<const uint32_t*>(sQh + off + 8 * LDQ);
          qh[mt][kk][2] = *reinterpret_cast<const uint32_t*>(sQh + off + 8);
          qh[mt][kk][3] = *reinterpret_cast<const uint32_t*>(sQh + off + 8 * LDQ + 8);
          ql[mt][kk][0] = *reinterpret_cast<const uint32_t*>(sQl + off);
          ql[mt][kk][1] = *reinterpret_cast<const uint32_t*>(sQl + off + 8 * LDQ);
          ql[mt][kk][2] = *reinterpret_cast<const uint32_t*>(sQl + off + 8);
          ql[mt][kk][3] = *reinterpret_cast<const uint32_t*>(sQl + off + 8 * LDQ + 8);
        }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t kb[4];
        const int row = kw0 + nt * 8 + (lane & 7), c = 2 * ks + (lane >> 3);
        ldsm_x4(kb, sk + (row * CH + (c ^ (row & 7))) * 16);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(S[mt][nt], qh[mt][0], kb[0], kb[1]);
          mma16816(S[mt][nt], ql[mt][0], kb[0], kb[1]);
          mma16816(S[mt][nt], qh[mt][1], kb[2], kb[3]);
          mma16816(S[mt][nt], ql[mt][1], kb[2], kb[3]);
        }
      }
    }

    // ---- online softmax (rows g8 and g8 + 8 of each head tile) ----
    uint32_t ph[MT][NT][2], pl[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int key = tile0 + kw0 + nt * 8 + 2 * t4;
        if (key >= k_end) S[mt][nt][0] = S[mt][nt][2] = -CUDART_INF_F;
        if (key + 1 >= k_end) S[mt][nt][1] = S[mt][nt][3] = -CUDART_INF_F;
        mx0 = fmaxf(mx0, fmaxf(S[mt][nt][0], S[mt][nt][1]));
        mx1 = fmaxf(mx1, fmaxf(S[mt][nt][2], S[mt][nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(mrow[mt][0], mx0), mn1 = fmaxf(mrow[mt][1], mx1);
      const float base0 = (mn0 == -CUDART_INF_F) ? 0.f : mn0, base1 = (mn1 == -CUDART_INF_F) ? 0.f : mn1;
      const float corr0 = exp2f(mrow[mt][0] - base0), corr1 = exp2f(mrow[mt][1] - base1);
      mrow[mt][0] = mn0;
      mrow[mt][1] = mn1;
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float p0 = exp2f(S[mt][nt][0] - base0), p1 = exp2f(S[mt][nt][1] - base0);
        const float p2 = exp2f(S[mt][nt][2] - base1), p3 = exp2f(S[mt][nt][3] - base1);
        ls0 += p0 + p1;
        ls1 += p2 + p3;
        split2(p0, p1, ph[mt][nt][0], pl[mt][nt][0]);
        split2(p2, p3, ph[mt][nt][1], pl[mt][nt][1]);
      }
      lrow[mt][0] = lrow[mt][0] * corr0 + ls0;  // per-lane partial sums; the quad reduction happens once at the end
      lrow[mt][1] = lrow[mt][1] * corr1 + ls1;
#pragma unroll
      for (int dn = 0; dn < HS / 8; ++dn) {
        O[mt][dn][0] *= corr0;
        O[mt][dn][1] *= corr0;
        O[mt][dn][2] *= corr1;
        O[mt][dn][3] *= corr1;
      }
    }

    // ---- O += P . V ----
#pragma unroll
    for (int kk = 0; kk < KPW / 16; ++kk) {
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        ah[mt][0] = ph[mt][2 * kk][0];
        ah[mt][1] = ph[mt][2 * kk][1];
        ah[mt][2] = ph[mt][2 * kk + 1][0];
        ah[mt][3] = ph[mt][2 * kk + 1][1];
        al[mt][0] = pl[mt][2 * kk][0];
        al[mt][1] = pl[mt][2 * kk][1];
        al[mt][2] = pl[mt][2 * kk + 1][0];
        al[mt][3] = pl[mt][2 * kk + 1][1];
      }
#pragma unroll
      for (int dn = 0; dn < HS / 8; dn += 2) {
        uint32_t vb[4];
        const int row = kw0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = dn + (lane >> 4);
        ldsm_x4_trans(vb, sv + (row * CH + (c ^ (row & 7))) * 16);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(O[mt][dn], ah[mt], vb[0], vb[1]);
          mma16816(O[mt][dn], al[mt], vb[0], vb[1]);
          mma16816(O[mt][dn + 1], ah[mt], vb[2], vb[3]);
          mma16816(O[mt][dn + 1], al[mt], vb[2], vb[3]);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // ring memory is free: reuse it as the merge buffer (also orders q staging when ntiles == 0)

  // ---- 4. merge the WT key groups of the CTA -----------------------------------------------------------------------
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float l0 = lrow[mt][0], l1 = lrow[mt][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float* r0 = sMerge + ((size_t)wt * QROWS + (wh + mt * WH) * 16 + g8) * LDM;
    float* r1 = r0 + 8 * LDM;
#pragma unroll
    for (int dn = 0; dn < HS / 8; ++dn) {
      *reinterpret_cast<float2*>(r0 + dn * 8 + 2 * t4) = make_float2(O[mt][dn][0], O[mt][dn][1]);
      *reinterpret_cast<float2*>(r1 + dn * 8 + 2 * t4) = make_float2(O[mt][dn][2], O[mt][dn][3]);
    }
    if (t4 == 0) {
      r0[HS] = mrow[mt][0];
      r0[HS + 1] = l0;
      r1[HS] = mrow[mt][1];
      r1[HS + 1] = l1;
    }
  }
  __syncthreads();
  for (int i = tid; i < qpk * HS; i += AD_THREADS) {
    const int h = i / HS, d = i % HS;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < WT; ++w) m = fmaxf(m, sMerge[((size_t)w * QROWS + h) * LDM + HS]);
    float l = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < WT; ++w) {
      const float* r = sMerge + ((size_t)w * QROWS + h) * LDM;
      const float c = (r[HS] == -CUDART_INF_F) ? 0.f : exp2f(r[HS] - m);
      l = fmaf(r[HS + 1], c, l);
      o = fmaf(r[d], c, o);
    }
    const int head = g * qpk_all + h_first + h;
    if (p.n_splits == 1) {
      p.out[(size_t)b * p.H * HS + (size_t)head * HS + d] = maybe_round(o / l, p.round_bf16);
    } else {
      float* dst = p.part + (((size_t)b * p.H + head) * p.n_splits + split) * LDM;
      dst[d] = o;
      if (d == 0) {
        dst[HS] = m;
        dst[HS + 1] = l;
      }
    }
  }
  if (p.n_splits == 1) return;

  // ---- 5. last CTA of this (batch, group) merges the splits --------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.tickets + ticket_id, 1) == p.n_splits - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // (a) every split's (m, l) for every head of the group -> shared memory (one L2 round trip, all loads independent)
  const int ns = p.n_splits;
  float* sM = sMerge;             // [qpk][ns]  running max, then the normalised weight of the split
  float* sL = sMerge + qpk * ns;  // [qpk][ns]
  for (int i = tid; i < qpk * ns; i += AD_THREADS) {
    const int h = i / ns, s = i % ns;
    const float* src = p.part + (((size_t)b * p.H + g * qpk_all + h_first + h) * ns + s) * LDM;
    sM[i] = __ldcg(src + HS);
    sL[i] = __ldcg(src + HS + 1);
  }
  __syncthreads();
  // (b) one warp per head: global max, split weights exp2(m_s - m) / l
  for (int h = warp; h < qpk; h += AD_THREADS / 32) {
    float m = -CUDART_INF_F;
    for (int s = lane; s < ns; s += 32) m = fmaxf(m, sM[h * ns + s]);
    m = warp_max(m);
    float l = 0.f;
    for (int s = lane; s < ns; s += 32) {
      const float ms = sM[h * ns + s];
      const float c = (ms == -CUDART_INF_F) ? 0.f : exp2f(ms - m);
      l = fmaf(sL[h * ns + s], c, l);
      sM[h * ns + s] = c;
    }
    l = warp_sum(l);
    const float inv = 1.0f / l;
    for (int s = lane; s < ns; s += 32) sM[h * ns + s] *= inv;
  }
  __syncthreads();
  // (c) weighted sum of the partial outputs in split order; the loads do not depend on the accumulation chain
  for (int i = tid; i < qpk * HS; i += AD_THREADS) {
    const int h = i / HS, d = i % HS;
    const int head = g * qpk_all + h_first + h;
    const float* src = p.part + ((size_t)b * p.H + head) * ns * LDM + d;
    const float* w = sM + h * ns;
    float o = 0.f;
#pragma unroll 8
    for (int s = 0; s < ns; ++s) o = fmaf(__ldcg(src + (size_t)s * LDM), w[s], o);
    p.out[(size_t)b * p.H * HS + (size_t)head * HS + d] = maybe_round(o, p.round_bf16);
  }
  if (tid == 0) p.tickets[ticket_id] = 0;
}

template <int HS>
static size_t ad_smem_bytes() {
  return (size_t)AD_STAGES * 2 * AD_TILE * HS * 2 + (size_t)2 * 16 * (HS + 8) * 2 + (size_t)2 * HS * 2;
}

template <int HS>
static int ad_launch(const AttnDecParams& p, int B, int head_tiles, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = attn_decode_fused_kernel<HS>;
  const size_t smem = ad_smem_bytes<HS>();
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    mark_device_setup(attr_set);
  }
  return launch(kern, dim3(B * p.G, p.n_splits, head_tiles), dim3(AD_THREADS), smem, stream, p);
}

static int ad_splits(int BG, int max_seq) {
  const int ntiles = (max_seq + AD_TILE - 1) / AD_TILE;
  int want = (2 * num_sms() + BG - 1) / BG;
  if (want < 1) want = 1;
  const int tiles_per_split = (ntiles + want - 1) / want;
  int n = (ntiles + tiles_per_split - 1) / tiles_per_split;
  if (n > 64) n = 64;
  return n;
}

static size_t ad_ticket_bytes(int n) { return ((size_t)n * sizeof(int) + 255) / 256 * 256; }

}  // namespace lp

extern "C" {

size_t lp_attn_fused_workspace_bytes(int B, int H, int G, int hs, int max_seq) {
  if (B <= 0 || H <= 0 || G <= 0 || hs <= 0 || max_seq <= 0) return 0;
  const int head_tiles = (H / G + 15) / 16;
  const int n = lp::ad_splits(B * G * head_tiles, max_seq);
  return lp::ad_ticket_bytes(B * G * head_tiles) + sizeof(float) * (size_t)B * H * n * (hs + 2);
}

int lp_attn_decode_fused(const float* qkv, const float* cos, const float* sin, const int32_t* pos, float* out, void* k_cache,
                         void* v_cache, int kv_dtype, void* workspace, size_t workspace_bytes, int B, int H, int G, int hs,
                         int n_elem, int max_seq, float scale, int round_bf16, void* stream) {
  if (!qkv || !pos || !out || !k_cache || !v_cache) return LP_ERR_INVALID_ARG;
  if (B <= 0 || H <= 0 || G <= 0 || H % G || hs <= 0 || max_seq <= 0 || n_elem < 0 || n_elem > hs || ((n_elem & 1) && n_elem != 1)) return LP_ERR_INVALID_ARG;
  if (n_elem > 0 && (!cos || !sin)) return LP_ERR_INVALID_ARG;
  if (kv_dtype != LP_BF16 || (hs != 64 && hs != 128)) return LP_ERR_UNSUPPORTED;
  const int head_tiles = (H / G + 15) / 16;
  lp::AttnDecParams p;
  p.qkv = qkv;
  p.cosT = cos;
  p.sinT = sin;
  p.pos = pos;
  p.kc = reinterpret_cast<__nv_bfloat16*>(k_cache);
  p.vc = reinterpret_cast<__nv_bfloat16*>(v_cache);
  p.out = out;
  p.H = H;
  p.G = G;
  p.n_elem = n_elem;
  p.max_seq = max_seq;
  p.round_bf16 = round_bf16;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.n_splits = lp::ad_splits(B * G * head_tiles, max_seq);
  const size_t tb = lp::ad_ticket_bytes(B * G * head_tiles);
  if (p.n_splits > 1) {
    if (!workspace) return LP_ERR_INVALID_ARG;
    if (workspace_bytes < tb + sizeof(float) * (size_t)B * H * p.n_splits * (hs + 2)) return LP_ERR_WORKSPACE;
  }
  p.tickets = reinterpret_cast<int*>(workspace);
  p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + tb);
  if (hs == 128) return lp::ad_launch<128>(p, B, head_tiles, stream);
  return lp::ad_launch<64>(p, B, head_tiles, stream);
}

}  // extern "C"

// =====================================================================================================================
// Prefill attention (T > 1 query rows against the bf16 KV cache, causal): FlashAttention-2 style on mma.sync.
//
//   reference: scaled_dot_product_attention with the lower-triangular mask rows of `input_pos` (model.py:91-92, 256-275).
//
// grid (ceil(T / 64), H, B), 128 threads: a CTA owns 64 consecutive query positions of one head, each warp 16 of them.
// K/V tiles (64 keys) of the head's KV group stream through a cp.async ring exactly as in the decode kernel; tiles entirely
// above the causal diagonal are never loaded.  q (already rotated, fp32) is split into bf16 hi + lo, P likewise, unless
// `exact` is 0 (bf16-faithful mode: q is bf16-valued and P is rounded to bf16 like the reference's bf16 SDPA).
// Query row t sits at position pos0 + t and attends keys [0, pos0 + t]; requires pos0 + T <= max_seq (no ring wrap).
// =====================================================================================================================
namespace lp {

template <int HS>
__global__ void __launch_bounds__(AD_THREADS)
attn_prefill_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kc, const __nv_bfloat16* __restrict__ vc,
                    const int* __restrict__ pos, float* __restrict__ out, int T, int H, int G, int max_seq, float scale_log2, int exact,
                    int round_bf16) {
  constexpr int CH = HS / 8;
  constexpr int STAGE_BYTES = AD_TILE * HS * 2;
  constexpr int NT = AD_TILE / 8;  // score n-tiles per warp: every warp sees the whole 64-key tile
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (H / G);
  const uint32_t sKV_u32 = smem_u32(smem);

  pdl_wait();
  pdl_launch_dependents();
  const int pos0 = pos[0];
  const int q_first = qt * 64;                       // first query row (time index) of the CTA
  const int q_last = min(T, q_first + 64) - 1;
  const int k_end = pos0 + q_last + 1;               // keys [0, k_end) are visible to at least one row
  const int ntiles = (k_end + AD_TILE - 1) / AD_TILE;
  const __nv_bfloat16* kbase = kc + ((size_t)b * G + g) * max_seq * HS;
  const __nv_bfloat16* vbase = vc + ((size_t)b * G + g) * max_seq * HS;

  auto issue_tile = [&](int t) {
    if (t < ntiles) {
      const int tile0 = t * AD_TILE;
      const uint32_t sk = sKV_u32 + (t % AD_STAGES) * 2 * STAGE_BYTES, sv = sk + STAGE_BYTES;
#pragma unroll
      for (int i = tid; i < AD_TILE * CH; i += AD_THREADS) {
        const int row = i / CH, c = i % CH;
        const int key = tile0 + row;
        const bool valid = key < k_end;
        const size_t off = (size_t)(valid ? key : 0) * HS + c * 8;
        const uint32_t d = (row * CH + (c ^ (row & 7))) * 16;
        cp_async16(sk + d, kbase + off, valid ? 16 : 0);
        cp_async16(sv + d, vbase + off, valid ? 16 : 0);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < AD_STAGES - 1; ++s) issue_tile(s);

  // ---- this warp's 16 query rows as A fragments (hi / lo), straight from global memory ----
  const int r0 = q_first + warp * 16 + g8, r1 = r0 + 8;  // time indices of the two rows this lane holds
  uint32_t qh[HS / 16][4], ql[HS / 16][4];
  {
    const float* q0 = q + ((size_t)b * T + min(r0, T - 1)) * H * HS + (size_t)h * HS;
    const float* q1 = q + ((size_t)b * T + min(r1, T - 1)) * H * HS + (size_t)h * HS;
#pragma unroll
    for (int ks = 0; ks < HS / 16; ++ks) {
      const float2 a = *reinterpret_cast<const float2*>(q0 + ks * 16 + 2 * t4);
      const float2 c = *reinterpret_cast<const float2*>(q1 + ks * 16 + 2 * t4);
      const float2 a8 = *reinterpret_cast<const float2*>(q0 + ks * 16 + 8 + 2 * t4);
      const float2 c8 = *reinterpret_cast<const float2*>(q1 + ks * 16 + 8 + 2 * t4);
      split2(a.x * scale_log2, a.y * scale_log2, qh[ks][0], ql[ks][0]);
      split2(c.x * scale_log2, c.y * scale_log2, qh[ks][1], ql[ks][1]);
      split2(a8.x * scale_log2, a8.y * scale_log2, qh[ks][2], ql[ks][2]);
      split2(c8.x * scale_log2, c8.y * scale_log2, qh[ks][3], ql[ks][3]);
    }
  }
  const int lim0 = pos0 + r0, lim1 = pos0 + r1;  // last visible key of each row

  float O[HS / 8][4];
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int dn = 0; dn < HS / 8; ++dn) O[dn][0] = O[dn][1] = O[dn][2] = O[dn][3] = 0.f;

  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<AD_STAGES - 2>();
    __syncthreads();
    issue_tile(t + AD_STAGES - 1);
    const int tile0 = t * AD_TILE;
    const uint32_t sk = sKV_u32 + (t % AD_STAGES) * 2 * STAGE_BYTES, sv = sk + STAGE_BYTES;
    if (tile0 > pos0 + q_first + warp * 16 + 15) continue;  // whole tile above this warp's diagonal (warp-uniform)

    float S[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HS / 16; ks += 2) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t kb[4];
        const int row = nt * 8 + (lane & 7), c = 2 * ks + (lane >> 3);
        ldsm_x4(kb, sk + (row * CH + (c ^ (row & 7))) * 16);
        mma16816(S[nt], qh[ks], kb[0], kb[1]);
        mma16816(S[nt], qh[ks + 1], kb[2], kb[3]);
        if (exact) {
          mma16816(S[nt], ql[ks], kb[0], kb[1]);
          mma16816(S[nt], ql[ks + 1], kb[2], kb[3]);
        }
      }
    }
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int key = tile0 + nt * 8 + 2 * t4;
      if (key > lim0) S[nt][0] = -CUDART_INF_F;
      if (key + 1 > lim0) S[nt][1] = -CUDART_INF_F;
      if (key > lim1) S[nt][2] = -CUDART_INF_F;
      if (key + 1 > lim1) S[nt][3] = -CUDART_INF_F;
      mx0 = fmaxf(mx0, fmaxf(S[nt][0], S[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(S[nt][2], S[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float base0 = (mn0 == -CUDART_INF_F) ? 0.f : mn0, base1 = (mn1 == -CUDART_INF_F) ? 0.f : mn1;
    const float corr0 = fast_ex2(m0 - base0), corr1 = fast_ex2(m1 - base1);
    m0 = mn0;
    m1 = mn1;
    uint32_t ph[NT][2], pl[NT][2];
    float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float p0 = fast_ex2(S[nt][0] - base0), p1 = fast_ex2(S[nt][1] - base0);
      float p2 = fast_ex2(S[nt][2] - base1), p3 = fast_ex2(S[nt][3] - base1);
      split2(p0, p1, ph[nt][0], pl[nt][0]);
      split2(p2, p3, ph[nt][1], pl[nt][1]);
      if (!exact) {  // the row sum uses the probabilities that are actually multiplied with V
        const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&ph[nt][0]);
        const __nv_bfloat162 h1 = *reinterpret_cast<const __nv_bfloat162*>(&ph[nt][1]);
        p0 = __bfloat162float(h0.x); p1 = __bfloat162float(h0.y);
        p2 = __bfloat162float(h1.x); p3 = __bfloat162float(h1.y);
      }
      ls0 += p0 + p1;
      ls1 += p2 + p3;
    }
    l0 = l0 * corr0 + ls0;
    l1 = l1 * corr1 + ls1;
#pragma unroll
    for (int dn = 0; dn < HS / 8; ++dn) {
      O[dn][0] *= corr0;
      O[dn][1] *= corr0;
      O[dn][2] *= corr1;
      O[dn][3] *= corr1;
    }
#pragma unroll
    for (int kk = 0; kk < AD_TILE / 16; ++kk) {
      const uint32_t ah[4] = {ph[2 * kk][0], ph[2 * kk][1], ph[2 * kk + 1][0], ph[2 * kk + 1][1]};
      const uint32_t al[4] = {pl[2 * kk][0], pl[2 * kk][1], pl[2 * kk + 1][0], pl[2 * kk + 1][1]};
#pragma unroll
      for (int dn = 0; dn < HS / 8; dn += 2) {
        uint32_t vb[4];
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = dn + (lane >> 4);
        ldsm_x4_trans(vb, sv + (row * CH + (c ^ (row & 7))) * 16);
        mma16816(O[dn], ah, vb[0], vb[1]);
        mma16816(O[dn + 1], ah, vb[2], vb[3]);
        if (exact) {
          mma16816(O[dn], al, vb[0], vb[1]);
          mma16816(O[dn + 1], al, vb[2], vb[3]);
        }
      }
    }
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  if (r0 < T) {
    float* o = out + ((size_t)b * T + r0) * H * HS + (size_t)h * HS;
#pragma unroll
    for (int dn = 0; dn < HS / 8; ++dn)
      *reinterpret_cast<float2*>(o + dn * 8 + 2 * t4) = make_float2(maybe_round(O[dn][0] * i0, round_bf16), maybe_round(O[dn][1] * i0, round_bf16));
  }
  if (r1 < T) {
    float* o = out + ((size_t)b * T + r1) * H * HS + (size_t)h * HS;
#pragma unroll
    for (int dn = 0; dn < HS / 8; ++dn)
      *reinterpret_cast<float2*>(o + dn * 8 + 2 * t4) = make_float2(maybe_round(O[dn][2] * i1, round_bf16), maybe_round(O[dn][3] * i1, round_bf16));
  }
}

template <int HS>
static int ap_launch(const float* q, const void* kc, const void* vc, const int* pos, float* out, int B, int T, int H, int G, int max_seq,
                     float scale, int exact, int round_bf16, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = attn_prefill_kernel<HS>;
  const size_t smem = (size_t)AD_STAGES * 2 * AD_TILE * HS * 2;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mark_device_setup(attr_set);
  }
  return launch(kern, dim3((T + 63) / 64, H, B), dim3(AD_THREADS), smem, stream, q, reinterpret_cast<const __nv_bfloat16*>(kc),
                reinterpret_cast<const __nv_bfloat16*>(vc), pos, out, T, H, G, max_seq, scale * 1.4426950408889634f, exact, round_bf16);
}

}  // namespace lp

namespace lp {
int attn_prefill_tc(const float* q, const void* k_cache, const void* v_cache, const int32_t* pos, float* out, int B, int T, int H, int G,
                    int hs, int max_seq, float scale, int round_bf16, void* stream);  // attention_tc.cu
static int g_prefill_path = 0;  // 0 auto (tcgen05 for T >= 1024, measured break-even), 1 mma.sync kernel only, 2 tcgen05 kernel only
}  // namespace lp

extern "C" int lp_set_attn_prefill_path(int path) {
  if (path < 0 || path > 2) return LP_ERR_INVALID_ARG;
  lp::g_prefill_path = path;
  return LP_OK;
}

extern "C" int lp_attn_prefill(const float* q, const void* k_cache, const void* v_cache, int kv_dtype, const int32_t* pos, float* out, int B,
                               int T, int H, int G, int hs, int max_seq, float scale, int round_bf16, void* stream) {
  if (!q || !k_cache || !v_cache || !pos || !out) return LP_ERR_INVALID_ARG;
  if (B <= 0 || T <= 0 || H <= 0 || G <= 0 || H % G || max_seq <= 0) return LP_ERR_INVALID_ARG;
  if (kv_dtype != LP_BF16 || (hs != 64 && hs != 128) || T > max_seq) return LP_ERR_UNSUPPORTED;
  if (lp::g_prefill_path == 2 || (lp::g_prefill_path == 0 && T >= 1024)) {
    const int rc = lp::attn_prefill_tc(q, k_cache, v_cache, pos, out, B, T, H, G, hs, max_seq, scale, round_bf16, stream);
    if (rc != LP_ERR_UNSUPPORTED || lp::g_prefill_path == 2) return rc;
  }
  const int exact = round_bf16 ? 0 : 1;
  if (hs == 128) return lp::ap_launch<128>(q, k_cache, v_cache, pos, out, B, T, H, G, max_seq, scale, exact, round_bf16, stream);
  return lp::ap_launch<64>(q, k_cache, v_cache, pos, out, B, T, H, G, max_seq, scale, exact, round_bf16, stream);
}
