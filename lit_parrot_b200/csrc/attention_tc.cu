// Causal prefill attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TENSOR MEMORY).
//
//   reference: scaled_dot_product_attention over the KV cache for T > 1 consecutive positions (model.py:247, 256-275 with the
//   mask rows of model.py:91-92); q is already rotated, k / v are already in the cache (lp_rope_kv_append).
//
// One CTA = 128 query rows of one head (grid: query tiles x heads x batch), 9 warps:
//   warp 8 (one thread): TMA producer AND MMA issuer.  K / V tiles of 64 keys come straight from the bf16 cache
//            [B, G, max_seq, hs] through 2-D tensor maps (box 64 keys x 64 dims = 128-byte rows, hardware 128-byte swizzle, two
//            stages).  S = Q . K^T : A = Q (K-major), B = K tile (K-major), D = S[128 x 64] fp32 in TMEM.  O += P . V : A = P
//            (K-major, written by the softmax warps), B = V tile used AS LOADED — keys are the rows, i.e. an MN-major operand
//            (instruction-descriptor bit 16, leading-dimension byte offset = distance of the two 64-dim slabs), D = O[128 x hs];
//   warps 0-7 (query row r = TMEM lane r is shared by warps w and w + 4: keys 0-31 / 32-63 of a tile, dims 0..hs/2 / hs/2..hs of
//            O): convert q (fp32, pre-scaled by softmax scale x log2 e) to the swizzled bf16 operand, then per key tile:
//            tcgen05.ld half a row of S, causal mask, row maximum exchanged through shared memory, exp2, rescale of O in TMEM
//            (tcgen05.ld / st) only when some row maximum of the warp moved, P -> shared memory, signal.
// fp32-ACTIVATION accuracy: q and P are split into bf16 hi + lo terms (two MMAs each into the same accumulator); products with
// the bf16 K / V are exact in fp32.  bf16-faithful mode uses one term.
// S is double buffered in TMEM: the scores of key block j+1 are computed while the softmax warps work on block j, and P.V of
// block j runs while they work on block j+1.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"

namespace lp {

const CUtensorMap* tc_cached_map(const void* ptr, int rows, int K, int box_rows);  // gemm_tc.cu

constexpr int AT_BM = 128;      // query rows per CTA
constexpr int AT_BN = 64;       // keys per tile
constexpr int AT_SM_WARPS = 8;                       // softmax warps: two per TMEM lane quarter, 32 of the 64 key columns each
constexpr int AT_SM_THREADS = AT_SM_WARPS * 32;
constexpr int AT_THREADS = AT_SM_THREADS + 32;      // + the TMA / MMA control warp

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void at_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void at_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void at_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void at_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void at_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void at_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// shared-memory operand descriptor as {lo, hi}: lo = start address >> 4 | leading byte offset >> 4 << 16; hi = stride byte offset
// (8-row group distance, 1024 B) | version 1 | 128-byte swizzle
constexpr uint32_t AT_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t at_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFF) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ void at_mma(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_c), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(AT_DESC_HI) : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void at_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

struct AtParams {
  const float* q;   // [B*T, H*hs] rotated queries
  float* out;       // [B*T, H*hs]
  const int* pos;   // [T] consecutive positions, pos[0] + T <= max_seq
  int B, T, H, G, max_seq, nterms, round_bf16;
  float scale_log2;
};

template <int HS>
__global__ void __launch_bounds__(AT_THREADS)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const AtParams p) {
  constexpr int SL = HS / 64;                  // 64-dim (128-byte) slabs per row
  constexpr int Q_SLAB = AT_BM * 128;          // [128 rows][128 B]
  constexpr int KV_SLAB = AT_BN * 128;         // [64 keys][128 B]
  constexpr int KV_TILE = SL * KV_SLAB;        // one of K or V
  constexpr int TM_S = 0, TM_O = 128;          // TMEM columns: S[2][128 x 64] (double buffered), O[128 x HS]
  constexpr int TM_COLS = 256;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nterms = p.nterms;
  unsigned char* sQ = smem;                                   // [nterms][SL][128][128 B]
  unsigned char* sK = sQ + nterms * SL * Q_SLAB;              // [2 stages][SL][64][128 B]
  unsigned char* sV = sK + 2 * KV_TILE;
  unsigned char* sP = sV + 2 * KV_TILE;                       // [nterms][128][128 B]
  __shared__ __align__(8) uint64_t bars[9];                   // kv_full[2], kv_empty[2], s_full[2], p_ready, pv_done, q_ready
  __shared__ uint32_t s_tmem;
  __shared__ float s_mx[2][2][AT_BM];   // [tile parity][column half][row]: partial row maxima
  __shared__ float s_l[AT_BM];          // row sums of the upper column half (final normalisation)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * AT_BM, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (p.H / p.G);
  const uint32_t bar0 = at_smem_u32(bars);
  const uint32_t kv_full0 = bar0, kv_empty0 = bar0 + 16, s_full0 = bar0 + 32, p_ready = bar0 + 48, pv_done = bar0 + 56, q_ready = bar0 + 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      at_mbar_init(kv_full0 + 8 * s, 1);
      at_mbar_init(kv_empty0 + 8 * s, 1);
    }
    at_mbar_init(s_full0, 1);
    at_mbar_init(s_full0 + 8, 1);
    at_mbar_init(p_ready, AT_SM_THREADS);
    at_mbar_init(pv_done, 1);
    at_mbar_init(q_ready, AT_SM_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(at_smem_u32(&s_tmem)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  at_fence_before();
  __syncthreads();
  at_fence_after();
  const uint32_t tmem = s_tmem;

  pdl_wait();  // q and the cache rows of this prefill are written by the preceding kernels
  pdl_launch_dependents();
  const int p0 = p.pos[0];
  const int rows = min(AT_BM, p.T - t0);                      // valid query rows of this tile
  const int ntiles = (p0 + t0 + rows - 1) / AT_BN + 1;        // key tiles up to the diagonal of the last valid row
  const int kv_row0 = (b * p.G + g) * p.max_seq;              // first cache row of this (batch, group)

  if (warp == AT_SM_WARPS) {
    if (lane == 0) {
      // ======================= TMA producer + MMA issuer =======================
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HS >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      auto load_tile = [&](int j) {
        const int s = j & 1;
        at_mbar_expect_tx(kv_full0 + 8 * s, 2 * KV_TILE);
#pragma unroll
        for (int sl = 0; sl < SL; ++sl) {
          at_tma_2d(at_smem_u32(sK) + s * KV_TILE + sl * KV_SLAB, &map_k, sl * 64, kv_row0 + j * AT_BN, kv_full0 + 8 * s);
          at_tma_2d(at_smem_u32(sV) + s * KV_TILE + sl * KV_SLAB, &map_v, sl * 64, kv_row0 + j * AT_BN, kv_full0 + 8 * s);
        }
      };
      auto issue_s = [&](int j) {  // S = Q . K_j^T (all terms of q into the same accumulator)
        const int s = j & 1;
        at_mbar_wait(kv_full0 + 8 * s, (j >> 1) & 1);
        at_fence_after();
        uint32_t acc = 0;
        for (int t = 0; t < nterms; ++t) {
#pragma unroll
          for (int ks = 0; ks < HS / 16; ++ks) {
            const uint32_t qa = at_smem_u32(sQ) + (t * SL + ks / 4) * Q_SLAB + (ks % 4) * 32;
            const uint32_t ka = at_smem_u32(sK) + s * KV_TILE + (ks / 4) * KV_SLAB + (ks % 4) * 32;
            at_mma(tmem + TM_S + s * AT_BN, at_desc_lo(qa, 16), at_desc_lo(ka, 16), idesc_s, acc);
            acc = 1;
          }
        }
        at_commit(s_full0 + 8 * s);
      };
      load_tile(0);
      if (ntiles > 1) load_tile(1);
      at_mbar_wait(q_ready, 0);
      at_fence_after();
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        // S is double buffered: the scores of block j+1 are computed while the softmax warps work on block j (its buffer was
        // read by the softmax of block j-1, which p_ready(j-1) has confirmed)
        if (j + 1 < ntiles) issue_s(j + 1);
        at_mbar_wait(p_ready, j & 1);  // P_j is in shared memory, S_j has been read, O has been rescaled
        at_fence_after();
        uint32_t acc = j > 0 ? 1u : 0u;
        for (int t = 0; t < nterms; ++t) {
#pragma unroll
          for (int kk = 0; kk < AT_BN / 16; ++kk) {  // 16 keys per MMA: 32 B along P's rows, 16 rows (2 KB) down V
            const uint32_t pa = at_smem_u32(sP) + t * Q_SLAB + kk * 32;
            const uint32_t va = at_smem_u32(sV) + s * KV_TILE + kk * 16 * 128;
            at_mma(tmem + TM_O, at_desc_lo(pa, 16), at_desc_lo(va, KV_SLAB), idesc_o, acc);
            acc = 1;
          }
        }
        at_commit(pv_done);
        at_commit(kv_empty0 + 8 * s);
        if (j + 2 < ntiles) {
          at_mbar_wait(kv_empty0 + 8 * s, (j >> 1) & 1);
          load_tile(j + 2);
        }
      }
    }
  } else {
    // ======================= softmax warps: row r = TMEM lane r, column half `ch` =======================
    const int wq = warp & 3, ch = warp >> 2;
    const int r = wq * 32 + lane;
    constexpr int HC = AT_BN / 2;   // key columns of a tile per thread
    constexpr int HO = HS / 2;      // O columns per thread
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    // ---- q -> bf16 terms in the swizzled K-major operand layout (coalesced: a warp reads whole rows) ----
    {
      constexpr int LPR = HS / 4;  // lanes per row (float4 each)
      for (int rr = warp * 16 + lane / LPR; rr < warp * 16 + 16; rr += 32 / LPR) {
        const int d = (lane % LPR) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < rows) v = *reinterpret_cast<const float4*>(p.q + ((size_t)(b * p.T + t0 + rr) * p.H + h) * HS + d);
        float x[4] = {v.x * p.scale_log2, v.y * p.scale_log2, v.z * p.scale_log2, v.w * p.scale_log2};
        const int c = d / 8, slab = c / 8;  // 16-byte chunk of the row; 64-dim slab
        for (int t = 0; t < nterms; ++t) {
          uint32_t w[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
            x[2 * i] -= __bfloat162float(h0);
            x[2 * i + 1] -= __bfloat162float(h1);
            w[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          }
          unsigned char* dst = sQ + (t * SL + slab) * Q_SLAB + rr * 128 + (((c % 8) ^ (rr & 7)) << 4) + (d % 8) * 2;
          *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
        }
      }
    }
    at_fence_async_smem();
    at_mbar_arrive(q_ready);

    const int qpos = p0 + t0 + r;  // absolute position of this row: attends keys <= qpos
    float m_run = -CUDART_INF_F, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      at_mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
      at_fence_after();
      float sc[HC];
      {
        uint32_t v[32];
        at_ld32(tmem + lane_base + TM_S + (j & 1) * AT_BN + ch * HC, v);
#pragma unroll
        for (int i = 0; i < HC; ++i) sc[i] = __uint_as_float(v[i]);
      }
      const int key0 = j * AT_BN + ch * HC;
      if (key0 + HC - 1 > p0 + t0 + wq * 32) {  // these columns touch the diagonal for some row of this warp
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (key0 + i > qpos) sc[i] = -CUDART_INF_F;
      }
      float mx = sc[0];
#pragma unroll
      for (int i = 1; i < HC; ++i) mx = fmaxf(mx, sc[i]);
      // row maximum over both column halves (the partner warp has the same rows): shared memory, double buffered by tile
      s_mx[j & 1][ch][r] = mx;
      asm volatile("bar.sync 1, %0;\n" ::"n"(AT_SM_THREADS) : "memory");
      mx = fmaxf(mx, s_mx[j & 1][ch ^ 1][r]);
      const float m_new = fmaxf(m_run, mx);             // finite from tile 0 on (key 0 is visible to every row)
      const float base = m_new == -CUDART_INF_F ? 0.f : m_new;
      const float corr = fast_ex2(m_run - base);           // 0 for the first tile
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < HC; ++i) {
        sc[i] = fast_ex2(sc[i] - base);
        ls += sc[i];
      }
      l_run = fmaf(l_run, corr, ls);  // partial row sum of this column half; the halves are added at the end
      m_run = m_new;
      if (j > 0) {
        at_mbar_wait(pv_done, (j - 1) & 1);  // P.V of the previous block is complete: O is stable, the P buffer is free
        at_fence_after();
        if (__any_sync(0xffffffffu, corr != 1.0f)) {  // some row maximum of this warp moved: rescale this warp's half of O
#pragma unroll 1
          for (int c0 = 0; c0 < HO; c0 += 32) {
            uint32_t v[32];
            at_ld32(tmem + lane_base + TM_O + ch * HO + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
            at_st32(tmem + lane_base + TM_O + ch * HO + c0, v);
          }
        }
      }
      // ---- P -> bf16 terms, swizzled K-major rows of 64 keys (this thread: chunks ch*4 .. ch*4+3) ----
      for (int t = 0; t < nterms; ++t) {
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(sc[8 * c + 2 * i]), h1 = __float2bfloat16_rn(sc[8 * c + 2 * i + 1]);
            sc[8 * c + 2 * i] -= __bfloat162float(h0);
            sc[8 * c + 2 * i + 1] -= __bfloat162float(h1);
            w[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          }
          const int cc = ch * (HC / 8) + c;
          *reinterpret_cast<uint4*>(sP + t * Q_SLAB + r * 128 + ((cc ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      at_fence_async_smem();
      at_fence_before();
      at_mbar_arrive(p_ready);
    }
    // ---- O / l -> out (this thread: its half of the columns; l = sum of both halves) ----
    if (ch == 1) s_l[r] = l_run;
    asm volatile("bar.sync 1, %0;\n" ::"n"(AT_SM_THREADS) : "memory");
    if (ch == 0) s_l[r] += l_run;
    asm volatile("bar.sync 1, %0;\n" ::"n"(AT_SM_THREADS) : "memory");
    at_mbar_wait(pv_done, (ntiles - 1) & 1);
    at_fence_after();
    const float inv = 1.0f / s_l[r];
    float* dst = p.out + ((size_t)(b * p.T + t0 + r) * p.H + h) * HS + ch * HO;
#pragma unroll 1
    for (int c0 = 0; c0 < HO; c0 += 32) {
      uint32_t v[32];
      at_ld32(tmem + lane_base + TM_O + ch * HO + c0, v);
      if (r < rows) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + c0 + i) =
              make_float4(maybe_round(__uint_as_float(v[i]) * inv, p.round_bf16), maybe_round(__uint_as_float(v[i + 1]) * inv, p.round_bf16),
                          maybe_round(__uint_as_float(v[i + 2]) * inv, p.round_bf16), maybe_round(__uint_as_float(v[i + 3]) * inv, p.round_bf16));
      }
    }
    at_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    at_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

template <int HS>
static int at_launch(const CUtensorMap& mk, const CUtensorMap& mv, const AtParams& p, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = attn_prefill_tc_kernel<HS>;
  constexpr int SL = HS / 64;
  // q terms + 2 stages of K and V + P terms: hs 128 fp32 mode 161 KB (one CTA per SM), bf16 mode 113 KB (two per SM: one CTA's
  // softmax overlaps the other's MMAs); hs 64: 97 / 65 KB (two / three per SM)
  const size_t smem = (size_t)p.nterms * SL * AT_BM * 128 + 4 * SL * AT_BN * 128 + (size_t)p.nterms * AT_BM * 128 + 1024;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * SL * AT_BM * 128 + 4 * SL * AT_BN * 128 + 2 * AT_BM * 128 + 1024)));
    mark_device_setup(attr_set);
  }
  return launch(kern, dim3((p.T + AT_BM - 1) / AT_BM, p.H, p.B), dim3(AT_THREADS), smem, stream, mk, mv, p);
}

// LP_ERR_UNSUPPORTED -> the caller falls back to the mma.sync kernel (attention_decode.cu)
int attn_prefill_tc(const float* q, const void* k_cache, const void* v_cache, const int32_t* pos, float* out, int B, int T, int H, int G,
                    int hs, int max_seq, float scale, int round_bf16, void* stream) {
  if (hs != 64 && hs != 128) return LP_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(k_cache) | reinterpret_cast<uintptr_t>(v_cache) | reinterpret_cast<uintptr_t>(q)) & 15) return LP_ERR_UNSUPPORTED;
  const CUtensorMap* mk = tc_cached_map(k_cache, B * G * max_seq, hs, AT_BN);
  const CUtensorMap* mv = tc_cached_map(v_cache, B * G * max_seq, hs, AT_BN);
  if (!mk || !mv) return LP_ERR_UNSUPPORTED;
  AtParams p;
  p.q = q;
  p.out = out;
  p.pos = pos;
  p.B = B;
  p.T = T;
  p.H = H;
  p.G = G;
  p.max_seq = max_seq;
  p.nterms = round_bf16 ? 1 : 2;
  p.round_bf16 = round_bf16;
  p.scale_log2 = scale * 1.4426950408889634f;
  return hs == 128 ? at_launch<128>(*mk, *mv, p, stream) : at_launch<64>(*mk, *mv, p, stream);
}

}  // namespace lp
