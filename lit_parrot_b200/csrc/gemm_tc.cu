// Dense projections for prefill (and wide decode batches) on the 5th-generation tensor cores:
//   C[M,N] = epilogue( X[M,K] . W[N,K]^T + bias ),  X and W bf16 K-major, fp32 accumulation in TENSOR MEMORY.
//
//   reference ops: every nn.Linear of GPT.forward applied to T > 1 tokens (model.py:111, 205, 252, 285-301).
//
// PERSISTENT CTAs (one per SM) walk the 128 x BN output tiles (BN = 256 or 128), M tiles fastest; the accumulator is
// double buffered in TMEM (2 x BN of the 512 columns), so the epilogue of tile i overlaps the MMAs of tile i + 1:
//   warp 0   : TMA producer — cp.async.bulk.tensor.2d loads of a 128 x 64 X tile (per activation term, see below) and a
//              BN x 64 W tile per stage into a 128-byte-swizzled ring, full/empty mbarriers;
//   warp 1   : MMA issuer — one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M 128, N BN, K 16) from
//              shared-memory descriptors, accumulating into TMEM (BN columns x 128 lanes of fp32); tcgen05.commit
//              releases ring slots and finally signals the epilogue;
//   warp 2   : allocates / frees the TMEM columns;
//   warps 2-5: epilogue — tcgen05.ld (32 lanes x 32 columns per instruction), bias / GELU(erf) / SwiGLU / residual,
//              then either fp32 rows or the bf16 hi/mid/lo split of the result (the next GEMM's operand).
// fp32-ACTIVATION accuracy on bf16 tensor cores: the activation is passed as `nterms` bf16 arrays with
// x = t0 + t1 (+ t2) exactly (lp_split_bf16); every term is multiplied with the same W tile (nterms MMAs per k-step), the
// products are exact in the fp32 accumulator.  bf16-faithful mode uses one term.
// Grid: (M tiles, N tiles) with the M tiles fastest, so the CTAs that share a W tile run together and W is streamed from
// HBM once while the (small) activation tiles stay in L2.
#include <cuda.h>

#include <atomic>
#include <type_traits>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace lp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;       // 128 bytes of bf16: one swizzle span
constexpr int TC_THREADS = 192; // 6 warps

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
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
// Wait used by the FOUR epilogue warps for `accumulator full`: they wait for a whole tile's main loop (tens of thousands of
// cycles); 128 threads re-issuing try_wait back to back compete with the TMA writes and the MMA's operand reads for shared
// memory, so they sleep between polls (LP_GEMM_DEBUG bit 2 restores the hot spin for A/B measurements).
__device__ __forceinline__ void tc_mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "nanosleep.u32 %2;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// multicast: the box lands at the same shared-memory offset of every CTA in `mask`, and completes on the mbarrier at the same
// offset in each of them
__device__ __forceinline__ void tc_tma_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_prefetch_2d(const CUtensorMap* map, int c0, int c1) {  // one box into L2 only
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tc_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// K-major, 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// The same descriptor as {lo, hi} words: only the 14-bit start-address field (bytes >> 4) changes from MMA to MMA, so the issuing
// thread adds to `lo` instead of rebuilding 64-bit values (it is the bottleneck for the narrow decode-batch MMAs).
constexpr uint32_t TC_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFF) | (1u << 16); }
__device__ __forceinline__ void tc_mma_lo(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_c), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI) : "memory");
}

// tuning aid (LP_GEMM_DEBUG bit 1): the MMA issuer of CTA 0 records {SM cycles, ns, k-blocks} of its main loops
__device__ long long g_tc_debug[4];

struct TcParams {
  const float* bias;      // [N] or NULL
  const float* out_bias;  // adapter-v2 output affine (lp_weight.out_bias / out_scale), [N] or NULL
  const float* out_scale;
  const float* residual;  // [M, N] or NULL
  float* out_f32;         // [M, Nout] or NULL
  __nv_bfloat16* out_bf;  // [out_terms][M, Nout] bf16 split of the result, or NULL
  int M, N, K, epi, round_bf16, nterms, out_terms;
  int l2_ahead;  // k-blocks the producer prefetches into L2 ahead of its TMA loads (first-touch W tiles come from HBM)
  int debug;  // tuning aid (LP_GEMM_DEBUG): bit 0 = skip the epilogue arithmetic and stores (timing experiments only)
};

// Epilogue of one 128 x BN accumulator tile held in TMEM at `tacc`: warp quarter q owns lanes [32 q, +32) = rows m0 + 32 q + lane.
template <int BN, bool AFF>
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, uint32_t tacc, int m0, int n0, int q, int lane, bool swiglu, int nout) {
  const int row = m0 + q * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    tc_ld32(tacc + ((uint32_t)(q * 32) << 16) + c0, v);
    if (row < p.M) {
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        float t = __uint_as_float(v[j]);
        if (p.bias && n < p.N) t += p.bias[n];
        y[j] = maybe_round(t, p.round_bf16);
        if (AFF && n < p.N) y[j] = out_affine(y[j], p.out_bias, p.out_scale, n, p.round_bf16);
      }
      int ncols = 32, ocol = n0 + c0;
      if (swiglu) {  // W rows interleaved: column 2i = fc_1 row i, 2i+1 = fc_2 row i (model.py:298-300)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = maybe_round(silu(y[2 * j]), p.round_bf16);
          y[j] = maybe_round(a * y[2 * j + 1], p.round_bf16);
        }
        ncols = 16;
        ocol = (n0 + c0) / 2;
      } else if (p.epi == LP_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = maybe_round(gelu_erf(y[j]), p.round_bf16);
      } else if (p.epi == LP_EPI_RESIDUAL) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ocol + j < nout) y[j] = maybe_round(p.residual[(size_t)row * nout + ocol + j] + y[j], p.round_bf16);
      }
      if (p.out_f32) {
        float* dst = p.out_f32 + (size_t)row * nout + ocol;
        if (ocol + ncols <= nout && (nout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < ncols) *reinterpret_cast<float4*>(dst + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols && ocol + j < nout) dst[j] = y[j];
        }
      }
      if (p.out_bf) {
        const bool vec = ocol + ncols <= nout && (nout & 7) == 0;  // 16-byte stores of 8 bf16
        for (int t = 0; t < p.out_terms; ++t) {
          __nv_bfloat16* dst = p.out_bf + ((size_t)t * p.M + row) * nout + ocol;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(y[j]), h1 = __float2bfloat16_rn(y[j + 1]);
            y[j] -= __bfloat162float(h0);
            y[j + 1] -= __bfloat162float(h1);
            pk[j / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          }
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < ncols) *reinterpret_cast<uint4*>(dst + j) = make_uint4(pk[j / 2], pk[j / 2 + 1], pk[j / 2 + 2], pk[j / 2 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols && ocol + j < nout)
                dst[j] = __ushort_as_bfloat16((unsigned short)((j & 1) ? (pk[j / 2] >> 16) : (pk[j / 2] & 0xffffu)));
          }
        }
      }
    }
  }
}

// CL = 2: the CTAs of a 2-CTA cluster work on neighbouring M tiles of the SAME N tile (tiles_m even) in lock step and share the W
// tile: each loads one half of its rows and multicasts it to both, which cuts the L2 -> SM traffic the kernel is bound by from
// (nterms * 16 + BN / 8) KB to (nterms * 16 + BN / 16) KB per k-block.  A ring slot is free once BOTH CTAs' MMAs have read it.
template <int BN, int CL, bool AFF>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p, int nstages) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int A_BYTES = TC_BM * TC_BK * 2;           // per term
  constexpr int B_BYTES = BN * TC_BK * 2;
  const int stage_bytes = p.nterms * A_BYTES + B_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * 12 + 4];  // full[12], empty[12], accumulator full[2], accumulator empty[2]
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM, tiles_n = (p.N + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;
  const uint32_t bar0 = tc_smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (12 + s); };
  auto acc_full = [&](int a) { return bar0 + 8 * (24 + a); };
  auto acc_empty = [&](int a) { return bar0 + 8 * (26 + a); };
  const uint32_t ring = tc_smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      tc_mbar_init(full_bar(s), 1);
      tc_mbar_init(empty_bar(s), CL);  // CL = 2: this CTA's and the partner's MMAs have both read the slot
    }
    for (int a = 0; a < 2; ++a) {
      tc_mbar_init(acc_full(a), 1);
      tc_mbar_init(acc_empty(a), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {  // TMEM allocation: two accumulators of BN fp32 columns (power of two >= 32), by one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&s_tmem)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  uint32_t cta_rank = 0;
  if (CL == 2) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(cta_rank));
    tc_cluster_sync();  // the partner's barriers are initialised before anything is multicast to them
  }

  pdl_wait();  // the activation terms are written by the preceding kernel
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * BN;
        for (int kb = 0; kb < nk; ++kb) {
          if (p.l2_ahead > 0) {  // the operands of k-block kb + l2_ahead (possibly of this CTA's next tile) -> L2
            int pk = kb + p.l2_ahead, ptile = tile;
            while (pk >= nk) { pk -= nk; ptile += gridDim.x; }
            if (ptile < ntiles) {
              const int pm0 = (ptile % tiles_m) * TC_BM, pn0 = (ptile / tiles_m) * BN;
              if (CL == 1) tc_prefetch_2d(&map_w, pk * TC_BK, pn0);
              else tc_prefetch_2d(&map_w, pk * TC_BK, pn0 + (int)cta_rank * (BN / 2));
              for (int t = 0; t < p.nterms; ++t) tc_prefetch_2d(&map_x, pk * TC_BK, t * p.M + pm0);
            }
          }
          tc_mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t dst = ring + (uint32_t)s * stage_bytes;
          tc_mbar_expect_tx(full_bar(s), stage_bytes);
          for (int t = 0; t < p.nterms; ++t)  // term t of X: rows [t*M + m0, +128) of the stacked [nterms*M, K] tensor
            tc_tma_2d(dst + t * A_BYTES, &map_x, kb * TC_BK, t * p.M + m0, full_bar(s));
          if (CL == 1)
            tc_tma_2d(dst + p.nterms * A_BYTES, &map_w, kb * TC_BK, n0, full_bar(s));
          else  // this CTA's half of the W tile's rows, to both CTAs (map_w has a BN / 2 row box)
            tc_tma_2d_mc(dst + p.nterms * A_BYTES + cta_rank * (B_BYTES / 2), &map_w, kb * TC_BK, n0 + (int)cta_rank * (BN / 2), full_bar(s),
                         (uint16_t)3);
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0, ph = 0, it = 0;
      long long dbg_c0 = 0, dbg_t0 = 0, dbg_kb = 0;
      if (p.debug & 2) {
        dbg_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(dbg_t0));
      }
      // The issuing thread is the pace maker: with the term count a run-time value and 64-bit descriptors rebuilt per MMA the
      // loop below cost ~175 cycles per k-step + ~58 per MMA (measured: 936 cycles per k-block of four 128-cycle MMAs, whatever
      // the ring depth, cluster mode or epilogue did).  Hence: term count as a template constant, k-steps and terms fully
      // unrolled, and only the 14-bit start-address field of the descriptors advanced, in 32-bit arithmetic.
      const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
      const uint32_t ring_lo = tc_desc_lo(ring);
      auto issue_tiles = [&](auto nt_tag) {
        constexpr int NT = decltype(nt_tag)::value;
        const uint32_t b_off = (uint32_t)(NT * TC_BM * TC_BK * 2) >> 4;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
          const int a = it & 1;  // TMEM accumulator of this tile
          tc_mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1);  // the epilogue has drained it (first use: passes immediately)
          tc_fence_after();
          const uint32_t tacc = tmem + a * BN;
          for (int kb = 0; kb < nk; ++kb) {
            tc_mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_lo = ring_lo + (uint32_t)s * stage_lo;
            const uint32_t b_lo = a_lo + b_off;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
#pragma unroll
              for (int t = 0; t < NT; ++t)
                tc_mma_lo(tacc, a_lo + (uint32_t)(t * (TC_BM * TC_BK * 2 >> 4) + k * 2), b_lo + (uint32_t)(k * 2), idesc,
                          (k | t) ? 1u : (uint32_t)(kb != 0));
            }
            if (CL == 1) tc_commit(empty_bar(s));  // the slot is free once these MMAs have read it
            else tc_commit_mc(empty_bar(s), (uint16_t)3);  // ... in both CTAs: the partner multicasts into this slot too
            if (++s == nstages) { s = 0; ph ^= 1; }
          }
          tc_commit(acc_full(a));  // accumulator complete
          dbg_kb += nk;
        }
      };
      if (p.nterms == 1) issue_tiles(std::integral_constant<int, 1>{});
      else if (p.nterms == 2) issue_tiles(std::integral_constant<int, 2>{});
      else issue_tiles(std::integral_constant<int, 3>{});
      if ((p.debug & 2) && blockIdx.x == 0) {
        tc_mbar_wait(acc_full((it - 1) & 1), ((it - 1) >> 1) & 1);  // the last MMAs have completed
        long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
        g_tc_debug[0] = clock64() - dbg_c0;
        g_tc_debug[1] = t1 - dbg_t0;
        g_tc_debug[2] = dbg_kb;
      }
    }
  }
  if (warp >= 2) {
    // ===== epilogue: warp w may touch TMEM lanes [32 (w % 4), +32) =====
    const int q = warp & 3;
    const bool swiglu = p.epi == LP_EPI_SWIGLU;
    const int nout = swiglu ? p.N / 2 : p.N;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int a = it & 1;
    const int m0 = (tile % tiles_m) * TC_BM, n0 = (tile / tiles_m) * BN;
    if (p.debug & 4) tc_mbar_wait(acc_full(a), (it >> 1) & 1);
    else tc_mbar_wait_sleep(acc_full(a), (it >> 1) & 1, 256);
    tc_fence_after();
    if (!(p.debug & 1)) tc_epilogue_tile<BN, AFF>(p, tmem + a * BN, m0, n0, q, lane, swiglu, nout);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(acc_empty(a)) : "memory");  // TMEM buffer may be reused
    }
  }
  __syncthreads();
  if (CL == 2) tc_cluster_sync();  // no CTA leaves while the partner may still multicast into it / arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(2 * BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA PAIRS (cta_group::2): one 256 x 256 output tile per 2-CTA cluster.  With one CTA per tile the SS-mode MMA is bound by
// shared-memory bandwidth (per k-block the TMA writes and the MMA reads nterms * 16 + 32 KB each: ~190 B/clk of 128); as a
// pair each SM holds its own 128 rows of X and HALF of the W tile (128 of its 256 rows) and the tensor cores of both SMs see
// both halves, so per SM the traffic drops to nterms * 16 + 16 KB per k-block at the same MMA time.
//   both CTAs : TMA producer (own X rows, own half of W) — every load completes on the LEADER's `full` barrier;
//   leader    : issues tcgen05.mma.cta_group::2 (M 256, N 256, K 16) from its own descriptors (same offsets in both CTAs);
//               tcgen05.commit multicast frees the ring slot / publishes the accumulator in BOTH CTAs;
//   both CTAs : epilogue of their 128 accumulator rows; the `accumulator drained` arrivals go to the leader's barrier.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void tc_tma_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_mma2_lo(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_c), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {  // arrives on the barrier at this offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <bool AFF>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p, int nstages) {
  constexpr int BN = 256;                        // N of the pair's tile; each CTA stages BN / 2 rows of W
  extern __shared__ __align__(1024) unsigned char smem[];
  const int A_BYTES = TC_BM * TC_BK * 2;         // per term: this CTA's 128 rows of X
  constexpr int B_BYTES = (BN / 2) * TC_BK * 2;  // this CTA's half of the W tile
  const int stage_bytes = p.nterms * A_BYTES + B_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * 12 + 4];
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const bool leader = rank == 0;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
  const int tiles_m = (p.M + 2 * TC_BM - 1) / (2 * TC_BM), tiles_n = (p.N + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const uint32_t bar0 = tc_smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (12 + s); };
  auto acc_full = [&](int a) { return bar0 + 8 * (24 + a); };
  auto acc_empty = [&](int a) { return bar0 + 8 * (26 + a); };
  const uint32_t ring = tc_smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      tc_mbar_init(full_bar(s), 1);   // leader: its own arrive.expect_tx; the bytes of both CTAs' loads complete on it
      tc_mbar_init(empty_bar(s), 1);  // one multicast commit of the leader
    }
    for (int a = 0; a < 2; ++a) {
      tc_mbar_init(acc_full(a), 1);   // one multicast commit of the leader
      tc_mbar_init(acc_empty(a), 8);  // leader: one arrival per epilogue warp of BOTH CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc_cluster_sync();  // both CTAs' barriers exist
  if (warp == 2) {    // TMEM of both SMs: two accumulators of BN fp32 columns; one warp of each CTA takes part
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&s_tmem)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_cluster_sync();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int tile = cluster; tile < ntiles; tile += nclusters) {
        const int m0 = (tile % tiles_m) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile / tiles_m) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < nk; ++kb) {
          tc_mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t dst = ring + (uint32_t)s * stage_bytes;
          const uint32_t lfull = tc_mapa(full_bar(s), 0);  // the leader's barrier
          if (leader) tc_mbar_expect_tx(full_bar(s), 2 * stage_bytes);
          for (int t = 0; t < p.nterms; ++t)
            tc_tma_2d_pair(dst + t * A_BYTES, &map_x, kb * TC_BK, t * p.M + m0, lfull);
          tc_tma_2d_pair(dst + p.nterms * A_BYTES, &map_w, kb * TC_BK, n0, lfull);
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===== MMA issuer (leader only) =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      int s = 0, ph = 0, it = 0;
      // issue loop as in gemm_tc_kernel: template term count, unrolled, 32-bit descriptor arithmetic
      const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
      const uint32_t ring_lo = tc_desc_lo(ring);
      auto issue_tiles = [&](auto nt_tag) {
        constexpr int NT = decltype(nt_tag)::value;
        const uint32_t b_off = (uint32_t)(NT * TC_BM * TC_BK * 2) >> 4;
        for (int tile = cluster; tile < ntiles; tile += nclusters, ++it) {
          const int a = it & 1;
          tc_mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem + a * BN;
          for (int kb = 0; kb < nk; ++kb) {
            tc_mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a_lo = ring_lo + (uint32_t)s * stage_lo;
            const uint32_t b_lo = a_lo + b_off;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
#pragma unroll
              for (int t = 0; t < NT; ++t)
                tc_mma2_lo(tacc, a_lo + (uint32_t)(t * (TC_BM * TC_BK * 2 >> 4) + k * 2), b_lo + (uint32_t)(k * 2), idesc,
                           (k | t) ? 1u : (uint32_t)(kb != 0));
            }
            tc_commit2(empty_bar(s));
            if (++s == nstages) { s = 0; ph ^= 1; }
          }
          tc_commit2(acc_full(a));
        }
      };
      if (p.nterms == 1) issue_tiles(std::integral_constant<int, 1>{});
      else if (p.nterms == 2) issue_tiles(std::integral_constant<int, 2>{});
      else issue_tiles(std::integral_constant<int, 3>{});
    }
  }
  if (warp >= 2) {
    // ===== epilogue (both CTAs): this CTA's 128 rows x 256 columns =====
    const int q = warp & 3;
    const bool swiglu = p.epi == LP_EPI_SWIGLU;
    const int nout = swiglu ? p.N / 2 : p.N;
    int it = 0;
    for (int tile = cluster; tile < ntiles; tile += nclusters, ++it) {
      const int a = it & 1;
      const int m0 = (tile % tiles_m) * 2 * TC_BM + (int)rank * TC_BM, n0 = (tile / tiles_m) * BN;
      tc_mbar_wait(acc_full(a), (it >> 1) & 1);
      tc_fence_after();
      tc_epilogue_tile<BN, AFF>(p, tmem + a * BN, m0, n0, q, lane, swiglu, nout);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator may be reused: tell the leader
        const uint32_t lb = tc_mapa(acc_empty(a), 0);
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(lb) : "memory");
      }
    }
  }
  __syncthreads();
  tc_cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(2 * BN) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Decode batches (9 <= M <= 64 rows): "swap-AB".  The WEIGHTS are the 128-row A operand (M_mma = 128 weight rows per tile),
// the batch is the N dimension of the MMA (NB = M rounded up to 16), so a stage is KB x (16 KB of weights + NB x 128 B per
// activation term) — weight-streaming bound, like the GEMV — instead of 16 KB of zero-padded activations per 4 KB of weights.
// Both operands are fetched through 3-D tensor maps {128 B, rows, K-blocks} (as in linear_stream.cu) with KB = 4 (2) K-blocks per
// copy: 512 (256) contiguous bytes per weight row.  With one 128-byte segment per row per copy (2-D box) HBM delivered 1.5 TB/s.  TMEM holds
// C^T: lane = weight row n, column = batch row m; the epilogue thread of lane n walks the batch rows, so for a fixed m the
// warp stores 32 consecutive n: coalesced.  In-place residual projections (N = n_embd: only N / 128 tiles) are additionally
// split along K over `ksplit` CTAs and accumulate with atomic adds (x += partial).
// ---------------------------------------------------------------------------------------------------------------------
struct TcSwapParams {
  const float* bias;
  const float* out_bias;  // adapter-v2 output affine, [N] or NULL
  const float* out_scale;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_bf;
  int M, N, K, epi, round_bf16, nterms, out_terms, NB, ksplit, KB;
  int fuse;  // M % 16 == 0: the terms are one stacked B operand (N_mma = nterms * NB), their columns are added in the epilogue
  int streamk;  // accumulate-in-place ops: the (tile, K-stage) sequence is cut into gridDim.x EQUAL ranges, one per CTA (a range may
                // end in the middle of a tile and continue in the next one); partial tiles are added atomically
};

// The work of one CTA of gemm_tc_swap_kernel as a list of segments (tile, K-stage range).  Unit mode: units u = tile * ksplit + ks,
// u = blockIdx.x, + gridDim.x, ...  Stream mode (p.streamk): the CTA's contiguous range of the tile-major stage sequence.
template <bool SK>
struct SwapWork {
  int nk, ksplit, nunits, u;
  long long cur, end;
  __device__ __forceinline__ void init(const TcSwapParams& p, int nk_, int tiles_n) {
    nk = nk_;
    ksplit = p.ksplit;
    nunits = tiles_n * p.ksplit;
    u = blockIdx.x;
    if (SK) {
      const long long S = (long long)tiles_n * nk_;
      cur = S * blockIdx.x / gridDim.x;
      end = S * (blockIdx.x + 1) / gridDim.x;
    }
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (SK) {
      if (cur >= end) return false;
      tile = (int)(cur / nk);
      kb0 = (int)(cur - (long long)tile * nk);
      const long long e = min(end, (long long)(tile + 1) * nk);
      kb1 = kb0 + (int)(e - cur);
      cur = e;
      return true;
    }
    if (u >= nunits) return false;
    tile = u / ksplit;
    const int ks = u % ksplit;
    kb0 = (int)((long long)nk * ks / ksplit);
    kb1 = (int)((long long)nk * (ks + 1) / ksplit);
    u += gridDim.x;
    return true;
  }
};
constexpr int TC_SWAP_ACC = 256;  // TMEM columns per accumulator (nterms * NB <= 192), two accumulators

// out-of-line activations for one-shot epilogues (code size, see gemm_tc_swap_kernel)
__device__ __noinline__ float tc_gelu(float v) { return gelu_erf(v); }
__device__ __noinline__ float tc_silu(float v) { return silu(v); }

// Tuning aid, compiled in with -DLP_SWAP_TRACE only (tools/trace_swap.py): globaltimer stamps per CTA
// [entry, set-up done, first loads issued, first stage landed, last MMA committed, accumulator seen by the epilogue, epilogue done, exit]
#ifdef LP_SWAP_TRACE
__device__ unsigned long long g_swap_trace[4 * 256 * 12];  // the last four launches
__device__ unsigned int g_swap_count[256];
__device__ __forceinline__ unsigned long long tc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
#define SWT(i) g_swap_trace[(s_hist * 256 + blockIdx.x) * 12 + (i)] = tc_now()
#else
#define SWT(i) (void)0
#endif

template <bool AFF, bool SK>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_swap_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcSwapParams p, int nstages) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int A_BYTES = TC_BM * TC_BK * 2;  // weight slab: 128 rows x 64 k
  const int XB = p.NB * TC_BK * 2;            // activation slab of one term: NB rows x 64 k
  const int KB = p.KB;                        // K-blocks (slabs) per stage
  const int stage_bytes = KB * (A_BYTES + p.nterms * XB);
  __shared__ __align__(8) uint64_t bars[2 * 12 + 4];
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef LP_SWAP_TRACE
  __shared__ unsigned int s_hist_sm;
  if (threadIdx.x == 0) s_hist_sm = atomicAdd(&g_swap_count[blockIdx.x], 1u) & 3u;
  __syncthreads();
  const unsigned int s_hist = s_hist_sm;
#endif
  if (threadIdx.x == 0) SWT(0);
  const int nk = (p.K / TC_BK + KB - 1) / KB;  // stages along K (K % 64 == 0; slabs past the end are zero-filled by TMA)
  const int tiles_n = (p.N + TC_BM - 1) / TC_BM;
  const uint32_t bar0 = tc_smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (12 + s); };
  auto acc_full = [&](int a) { return bar0 + 8 * (24 + a); };
  auto acc_empty = [&](int a) { return bar0 + 8 * (26 + a); };
  const uint32_t ring = tc_smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      tc_mbar_init(full_bar(s), 1);
      tc_mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc_mbar_init(acc_full(a), 1);
      tc_mbar_init(acc_empty(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&s_tmem)), "r"(2 * TC_SWAP_ACC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) SWT(1);

  if (warp == 0) {
    // ===== TMA producer: the weights do not depend on the preceding kernel, the activation terms do =====
    if (lane == 0) {
      int s = 0, ph = 0;
      bool waited = false;
      SwapWork<SK> wk;
      wk.init(p, nk, tiles_n);
      int tile, kb0, kb1;
      while (wk.next(tile, kb0, kb1)) {
        const int n0 = tile * TC_BM;
        // every CTA needs the same activation slabs: start the K loop at a CTA-dependent offset (and wrap) so that the CTAs read
        // different slabs at any one time instead of the same few L2 lines
        const int nkb = kb1 - kb0, rot = (int)((blockIdx.x * 5u) % (unsigned)nkb);
        for (int i = 0; i < nkb; ++i) {
          const int kb = kb0 + (i + rot) % nkb;
          tc_mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t dst = ring + (uint32_t)s * stage_bytes;
          tc_mbar_expect_tx(full_bar(s), stage_bytes);
          tc_tma_3d(dst, &map_w, 0, n0, kb * KB, full_bar(s));
          if (!waited) {
            pdl_wait();
            waited = true;
            SWT(2);
          }
          if (p.fuse) {  // all terms in one copy: rows [0, nterms*M) of the stacked [nterms*M, K] tensor, slabs [kk][nterms*NB rows]
            tc_tma_3d(dst + KB * A_BYTES, &map_x, 0, 0, kb * KB, full_bar(s));
          } else {
            for (int t = 0; t < p.nterms; ++t)  // term t: rows [t*M, t*M + NB), slabs [t][kk][NB rows]
              tc_tma_3d(dst + KB * A_BYTES + t * KB * XB, &map_x, 0, t * p.M, kb * KB, full_bar(s));
          }
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[128 weight rows, NB batch rows] += W_slab . X_term_slab^T =====
    if (lane == 0) {
      const int nmma = p.fuse ? p.nterms * p.NB : p.NB;  // N of one MMA
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const int nb_ops = p.fuse ? 1 : p.nterms;                         // B operands per k-step
      const uint32_t bslab = (uint32_t)(p.fuse ? p.nterms * XB : XB) >> 4;  // distance of consecutive K slabs of one B operand
      const uint32_t bterm = (uint32_t)(KB * XB) >> 4;                  // distance of the terms (unfused layout)
      int s = 0, ph = 0, it = 0;
      SwapWork<SK> wk;
      wk.init(p, nk, tiles_n);
      int tile, kb0, kb1;
      for (; wk.next(tile, kb0, kb1); ++it) {
        const int a = it & 1;
        tc_mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem + a * TC_SWAP_ACC;
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc_mbar_wait(full_bar(s), ph);
          tc_fence_after();
#ifdef LP_SWAP_TRACE
          if (it == 0 && kb == kb0) SWT(3);
#endif
          const uint32_t st = ring + (uint32_t)s * stage_bytes;
          uint32_t a_lo = tc_desc_lo(st);
          uint32_t b_lo = tc_desc_lo(st + KB * A_BYTES);
          if (p.fuse && KB == 4) {
            // the common decode-batch case, fully unrolled: with run-time loop bounds the 16 MMAs of a stage cost ~95 cycles of
            // issue work each (they execute in 48), and with only two 96 KB stages that time is not hidden behind the loads
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                tc_mma_lo(tacc, a_lo + (uint32_t)(kk * (A_BYTES >> 4) + 2 * k), b_lo + (uint32_t)kk * bslab + (uint32_t)(2 * k), idesc,
                          (kk | k) ? 1u : accumulate);
            }
            accumulate = 1;
          } else
          for (int kk = 0; kk < KB; ++kk) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {  // 32 bytes (16 bf16) per MMA along K: +2 in the address field
              uint32_t bt = b_lo + 2 * k;
              for (int t = 0; t < nb_ops; ++t) {
                tc_mma_lo(tacc + (p.fuse ? 0 : t * p.NB), a_lo + 2 * k, bt, idesc, accumulate);
                bt += bterm;
              }
              accumulate = 1;
            }
            a_lo += A_BYTES >> 4;
            b_lo += bslab;
          }
          tc_commit(empty_bar(s));
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
        tc_commit(acc_full(a));
        SWT(4);
      }
    }
  }
  if (warp >= 2) {
    // ===== epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = weight rows; columns = batch rows =====
    pdl_wait();  // residual / bias of earlier kernels
    pdl_launch_dependents();
    const int q = warp & 3;
    const bool swiglu = p.epi == LP_EPI_SWIGLU;
    const int nout = swiglu ? p.N / 2 : p.N;
    const bool atomic = p.ksplit > 1 || SK;
    int it = 0;
    SwapWork<SK> wk;
    wk.init(p, nk, tiles_n);
    int tile, kb0, kb1;
    for (; wk.next(tile, kb0, kb1); ++it) {
      const int a = it & 1;
      const int n = tile * TC_BM + q * 32 + lane;  // weight row = output column
      const bool first = kb0 == 0;  // this CTA's part of the tile starts at K = 0: it adds the biases
      tc_mbar_wait(acc_full(a), (it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) SWT(5);
      const float bias = (p.bias && n < p.N && first) ? p.bias[n] : 0.f;
      // adapter-v2 (compile-time variant): scale * ((acc + bias) + out_bias); K-split partials: the biases once
      const float obv = (AFF && p.out_bias && n < p.N && first) ? p.out_bias[n] : 0.f;
      const float osv = (AFF && p.out_scale && n < p.N) ? p.out_scale[n] : 1.f;
#pragma unroll 1
      for (int c0 = 0; c0 < p.NB; c0 += 32) {
        // the accumulator holds one column block per activation term (term t of batch row m: column t*NB + m): add them
        uint32_t v[32];
        float acc[32];
        tc_ld32(tmem + a * TC_SWAP_ACC + ((uint32_t)(q * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        for (int t = 1; t < p.nterms; ++t) {
          tc_ld32(tmem + a * TC_SWAP_ACC + ((uint32_t)(q * 32) << 16) + t * p.NB + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
        }
#ifdef LP_SWAP_TRACE
        if (threadIdx.x == 64) SWT(8);
#endif
        // The tail runs ONCE per tile, as straight-line code over the 32 accumulator registers.  It is written as a sequence of
        // short uniform passes (one per step of the epilogue, the activation out of line) instead of one unrolled body holding
        // every variant: that body was ~16 k instructions, and a launch whose CTAs own a single tile spent as long fetching
        // them (all SMs walking the same cold instruction lines in lock step: 20 us) as streaming its weights (18 us).
        const bool rb = p.round_bf16 != 0;
#define SWAP_PASS(expr)                   \
  _Pragma("unroll") for (int j = 0; j < 32; ++j) { expr; }
#define SWAP_ROUND() \
  if (rb) SWAP_PASS(acc[j] = bf16_round(acc[j]))
        SWAP_PASS(acc[j] += bias)
        SWAP_ROUND()
        if (AFF) {
          SWAP_PASS(acc[j] += obv)
          SWAP_ROUND()
          SWAP_PASS(acc[j] *= osv)
          SWAP_ROUND()
        }
        int oc = n;
        bool store = n < p.N;
        if (swiglu) {  // W rows interleaved: row 2i = fc_1 row i, 2i+1 = fc_2 row i (model.py:298-300): neighbouring lanes
          SWAP_PASS(const float other = __shfl_xor_sync(0xffffffffu, acc[j], 1); acc[j] = maybe_round(tc_silu(acc[j]), p.round_bf16) * other)
          SWAP_ROUND()
          oc = n >> 1;
          store = store && (n & 1) == 0;
        } else if (p.epi == LP_EPI_GELU) {
          SWAP_PASS(acc[j] = tc_gelu(acc[j]))
          SWAP_ROUND()
        } else if (p.epi == LP_EPI_RESIDUAL && !atomic) {
          if (store) SWAP_PASS(if (c0 + j < p.M) acc[j] += p.residual[(size_t)(c0 + j) * nout + oc])
          SWAP_ROUND()
        }
        if (store) {
          if (atomic) {
            SWAP_PASS(if (c0 + j < p.M) atomicAdd(p.out_f32 + (size_t)(c0 + j) * nout + oc, acc[j]))  // in place: x += partial
          } else {
            if (p.out_f32) SWAP_PASS(if (c0 + j < p.M) p.out_f32[(size_t)(c0 + j) * nout + oc] = acc[j])
            if (p.out_bf) {
#pragma unroll 1
              for (int t = 0; t < p.out_terms; ++t) {
                __nv_bfloat16* dst = p.out_bf + (size_t)t * p.M * nout + oc;
                SWAP_PASS(if (c0 + j < p.M) {
                  const __nv_bfloat16 hb = __float2bfloat16_rn(acc[j]);
                  dst[(size_t)(c0 + j) * nout] = hb;
                  acc[j] -= __bfloat162float(hb);
                })
              }
            }
          }
        }
#undef SWAP_PASS
#undef SWAP_ROUND
      }
#ifdef LP_SWAP_TRACE
      if (threadIdx.x == 64) SWT(9);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(acc_empty(a)) : "memory");
      if (threadIdx.x == 64) SWT(6);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(2 * TC_SWAP_ACC) : "memory");
  }
  if (threadIdx.x == 0) SWT(7);
}

// ---------------------------------------------------------------------------------------------------------------------
// x fp32 [rows, K] (optionally through LayerNorm / RMSNorm) -> nterms bf16 arrays [nterms][rows, K] with x = sum of terms
// ---------------------------------------------------------------------------------------------------------------------
// One block of 512 threads per row; the row is read ONCE (float4 per thread and pass, up to 8 passes = 16384 columns, kept in
// registers through the statistics and the split; longer rows are re-read).  Decode batches call this with a handful of rows, so
// the kernel is latency-bound: one pass and two block reductions instead of three passes.
constexpr int SPL_THREADS = 512;
constexpr int SPL_CACHE = 8;
__global__ void __launch_bounds__(SPL_THREADS) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int K,
                                                                 int nterms, int norm_kind, const float* __restrict__ nw,
                                                                 const float* __restrict__ nb, float eps, int round_bf16) {
  __shared__ float red[2][SPL_THREADS / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x;
  const float* xr = x + (size_t)r * K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (K & 3) == 0 && K <= SPL_CACHE * SPL_THREADS * 4;  // register-cached path
  float4 xc[SPL_CACHE];
  if (vec) {
#pragma unroll
    for (int i = 0; i < SPL_CACHE; ++i) {
      const int k = (i * SPL_THREADS + threadIdx.x) * 4;
      xc[i] = k < K ? *reinterpret_cast<const float4*>(xr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  auto block_sum2 = [&](float& a, float& b) {  // both sums with one pair of barriers
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) {
      red[0][warp] = a;
      red[1][warp] = b;
    }
    __syncthreads();
    a = 0.f;
    b = 0.f;
#pragma unroll
    for (int w = 0; w < SPL_THREADS / 32; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
  };
  float mean = 0.f, rstd = 1.f;
  if (norm_kind >= 0) {
    float s = 0.f, ss = 0.f;
    if (vec) {
#pragma unroll
      for (int i = 0; i < SPL_CACHE; ++i) {
        s += (xc[i].x + xc[i].y) + (xc[i].z + xc[i].w);
        ss += xc[i].x * xc[i].x + xc[i].y * xc[i].y + xc[i].z * xc[i].z + xc[i].w * xc[i].w;
      }
    } else {
      for (int k = threadIdx.x; k < K; k += SPL_THREADS) {
        const float v = xr[k];
        s += v;
        ss += v * v;
      }
    }
    block_sum2(s, ss);
    if (norm_kind == LP_NORM_LAYERNORM) {
      mean = s / (float)K;
      float v2 = 0.f, dummy = 0.f;  // two-pass variance
      if (vec) {
#pragma unroll
        for (int i = 0; i < SPL_CACHE; ++i) {
          if ((i * SPL_THREADS + threadIdx.x) * 4 < K) {
            const float d0 = xc[i].x - mean, d1 = xc[i].y - mean, d2 = xc[i].z - mean, d3 = xc[i].w - mean;
            v2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        }
      } else {
        for (int k = threadIdx.x; k < K; k += SPL_THREADS) {
          const float d = xr[k] - mean;
          v2 += d * d;
        }
      }
      block_sum2(v2, dummy);
      rstd = 1.0f / sqrtf(v2 / (float)K + eps);
    } else {
      rstd = 1.0f / sqrtf(ss / (float)K + eps);
    }
  }
  auto emit = [&](int k, float v) {
    if (norm_kind == LP_NORM_LAYERNORM) v = (v - mean) * rstd * nw[k] + (nb ? nb[k] : 0.f);
    else if (norm_kind == LP_NORM_RMS) v = nw[k] * (v * rstd);
    v = maybe_round(v, round_bf16);
    for (int t = 0; t < nterms; ++t) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out[((size_t)t * rows + r) * K + k] = h;
      v -= __bfloat162float(h);
    }
  };
  if (vec) {
#pragma unroll
    for (int i = 0; i < SPL_CACHE; ++i) {
      const int k = (i * SPL_THREADS + threadIdx.x) * 4;
      if (k < K) {
        float v[4] = {xc[i].x, xc[i].y, xc[i].z, xc[i].w};
        float w4[4] = {1.f, 1.f, 1.f, 1.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (norm_kind >= 0) {
          const float4 t = *reinterpret_cast<const float4*>(nw + k);
          w4[0] = t.x; w4[1] = t.y; w4[2] = t.z; w4[3] = t.w;
          if (norm_kind == LP_NORM_LAYERNORM && nb) {
            const float4 u = *reinterpret_cast<const float4*>(nb + k);
            b4[0] = u.x; b4[1] = u.y; b4[2] = u.z; b4[3] = u.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (norm_kind == LP_NORM_LAYERNORM) v[j] = (v[j] - mean) * rstd * w4[j] + b4[j];
          else if (norm_kind == LP_NORM_RMS) v[j] = w4[j] * (v[j] * rstd);
          v[j] = maybe_round(v[j], round_bf16);
        }
        for (int t = 0; t < nterms; ++t) {  // 8-byte stores of 4 bf16
          uint32_t pk[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
            v[2 * j] -= __bfloat162float(h0);
            v[2 * j + 1] -= __bfloat162float(h1);
            pk[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          }
          *reinterpret_cast<uint2*>(out + ((size_t)t * rows + r) * K + k) = make_uint2(pk[0], pk[1]);
        }
      }
    }
  } else {
    for (int k = threadIdx.x; k < K; k += SPL_THREADS) emit(k, xr[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TcEncodeFn tc_encode_fn() {
  static TcEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<TcEncodeFn>(f);
  }();
  return fn;
}

// bf16 row-major [rows, K] -> 2-D map with a {64, box_rows} box, 128-byte swizzle; out-of-range rows / columns read as zero
static bool tc_make_map(CUtensorMap* m, const void* ptr, int rows, int K, int box_rows) {
  TcEncodeFn enc = tc_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {TC_BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

const CUtensorMap* tc_cached_map(const void* ptr, int rows, int K, int box_rows) {  // also used by attention_tc.cu
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(ptr, rows, K, box_rows);
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  CUtensorMap m;
  if (!tc_make_map(&m, ptr, rows, K, box_rows)) return nullptr;
  if (cache.size() > 8192) cache.clear();
  return &cache.emplace(key, m).first->second;
}

// bf16 row-major [rows, K] (K % 64 == 0) as a 3-D tensor {64 elements = 128 B, rows, K / 64 blocks}: one copy brings `kb` slabs
// of box_rows x 128 B, laid out [slab][row][128 B] under the 128-byte swizzle, from kb * 128 contiguous bytes of every row
static const CUtensorMap* tc_cached_map3(const void* ptr, int rows, int K, int box_rows, int kb) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(ptr, rows, K, box_rows, kb);
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  TcEncodeFn enc = tc_encode_fn();
  if (!enc) return nullptr;
  CUtensorMap m;
  cuuint64_t dims[3] = {TC_BK, (cuuint64_t)rows, (cuuint64_t)(K / TC_BK)};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {TC_BK, (cuuint32_t)box_rows, (cuuint32_t)kb};
  cuuint32_t es[3] = {1, 1, 1};
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return nullptr;
  if (cache.size() > 8192) cache.clear();
  return &cache.emplace(key, m).first->second;
}

template <int BN, int CL, bool AFF>
static int tc_launch_v(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = gemm_tc_kernel<BN, CL, AFF>;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    mark_device_setup(attr_set);
  }
  const int stage_bytes = p.nterms * TC_BM * TC_BK * 2 + BN * TC_BK * 2;
  int nstages = (212 * 1024) / stage_bytes;
  if (nstages > 12) nstages = 12;
  static const int cap = [] {  // tuning aid: LP_GEMM_STAGES caps the ring depth
    const char* e = getenv("LP_GEMM_STAGES");
    return e ? atoi(e) : 0;
  }();
  if (cap > 1 && nstages > cap) nstages = cap;
  if (nstages < 2) return LP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)nstages * stage_bytes + 1024;
  const int ntiles = ((p.M + TC_BM - 1) / TC_BM) * ((p.N + BN - 1) / BN);
  int grid = ntiles < num_sms() ? ntiles : num_sms();
  if (CL == 1) return launch(kern, dim3(grid), dim3(TC_THREADS), smem, stream, mx, mw, p, nstages);
  // 2-CTA clusters: even grid; tiles_m is even, so CTAs 2i and 2i+1 always get tiles t, t+1 of one N tile and the same number
  // of them (the pair runs in lock step)
  grid &= ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  LP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mx, mw, p, nstages));
  count_launch();
  return LP_OK;
}

static std::atomic<int> g_gemm_pair{[] {
  const char* e = getenv("LP_GEMM_PAIR");  // CTA pairs are the default for prefill-sized problems; LP_GEMM_PAIR=0 disables
  return e ? atoi(e) : 1;
}()};

template <bool AFF>
static int tc_launch_pair_v(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = gemm_tc_pair_kernel<AFF>;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    mark_device_setup(attr_set);
  }
  const int stage_bytes = p.nterms * TC_BM * TC_BK * 2 + 128 * TC_BK * 2;
  int nstages = (212 * 1024) / stage_bytes;
  if (nstages > 12) nstages = 12;
  const size_t smem = (size_t)nstages * stage_bytes + 1024;
  const int ntiles = ((p.M + 2 * TC_BM - 1) / (2 * TC_BM)) * ((p.N + 255) / 256);
  int grid = 2 * ntiles < num_sms() ? 2 * ntiles : (num_sms() & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  LP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mx, mw, p, nstages));
  count_launch();
  return LP_OK;
}

template <bool AFF, bool SK>
static int tc_launch_swap_v(const CUtensorMap& mx, const CUtensorMap& mw, const TcSwapParams& p, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = gemm_tc_swap_kernel<AFF, SK>;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    mark_device_setup(attr_set);
  }
  const int stage_bytes = p.KB * (TC_BM * TC_BK * 2 + p.nterms * p.NB * TC_BK * 2);
  int nstages = (212 * 1024) / stage_bytes;
  if (nstages > 12) nstages = 12;
  if (nstages < 2) return LP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)nstages * stage_bytes + 1024;
  const int tiles = (p.N + TC_BM - 1) / TC_BM;
  const long long nunits = p.streamk ? (long long)tiles * ((p.K / TC_BK + p.KB - 1) / p.KB) : (long long)tiles * p.ksplit;
  const int grid = nunits < num_sms() ? (int)nunits : num_sms();
  return launch(kern, dim3(grid), dim3(TC_THREADS), smem, stream, mx, mw, p, nstages);
}

// Run-time selection of the compile-time variants: AFF = adapter-v2 output affine in the epilogue, SK = stream-K work split.
static int tc_launch_swap(const CUtensorMap& mx, const CUtensorMap& mw, const TcSwapParams& p, void* stream) {
  const bool aff = p.out_bias || p.out_scale;
  if (p.streamk) return aff ? tc_launch_swap_v<true, true>(mx, mw, p, stream) : tc_launch_swap_v<false, true>(mx, mw, p, stream);
  return aff ? tc_launch_swap_v<true, false>(mx, mw, p, stream) : tc_launch_swap_v<false, false>(mx, mw, p, stream);
}
template <int BN, int CL>
static int tc_launch(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, void* stream) {
  return (p.out_bias || p.out_scale) ? tc_launch_v<BN, CL, true>(mx, mw, p, stream) : tc_launch_v<BN, CL, false>(mx, mw, p, stream);
}
static int tc_launch_pair(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, void* stream) {
  return (p.out_bias || p.out_scale) ? tc_launch_pair_v<true>(mx, mw, p, stream) : tc_launch_pair_v<false>(mx, mw, p, stream);
}

}  // namespace lp

extern "C" {

#ifdef LP_SWAP_TRACE
int lp_debug_swap_trace(unsigned long long* out, int n) {  // tuning builds only (tools/trace_swap.py)
  return cudaMemcpyFromSymbol(out, lp::g_swap_trace, sizeof(unsigned long long) * (n < 12288 ? n : 12288)) == cudaSuccess ? LP_OK : LP_ERR_CUDA;
}
#endif

int lp_debug_gemm_stats(long long* out4) {  // {cycles, ns, k-blocks, -} of the last LP_GEMM_DEBUG=2 launch (CTA 0)
  return cudaMemcpyFromSymbol(out4, lp::g_tc_debug, 4 * sizeof(long long)) == cudaSuccess ? LP_OK : LP_ERR_CUDA;
}

int lp_set_gemm_pair(int enabled) {
  lp::g_gemm_pair.store(enabled ? 1 : 0);
  return LP_OK;
}

int lp_split_bf16(const float* x, void* out_bf16, int rows, int K, int nterms, int norm_kind, const float* norm_w, const float* norm_b,
                  float eps, int round_bf16, void* stream) {
  if (!x || !out_bf16 || rows <= 0 || K <= 0 || nterms < 1 || nterms > 3) return LP_ERR_INVALID_ARG;
  if (norm_kind >= 0 && !norm_w) return LP_ERR_INVALID_ARG;
  return lp::launch(lp::split_bf16_kernel, dim3(rows), dim3(lp::SPL_THREADS), 0, stream, x, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, K, nterms,
                    norm_kind, norm_w, norm_b, eps, round_bf16);
}

int lp_gemm_bf16_tc(const void* x_terms, int nterms, int M, const void* w_bf16, int N, int K, const float* bias, int epilogue,
                    const float* residual, float* out_f32, void* out_bf16, int out_terms, int round_bf16, void* stream) {
  return lp_gemm_bf16_tc_affine(x_terms, nterms, M, w_bf16, N, K, bias, nullptr, nullptr, epilogue, residual, out_f32, out_bf16, out_terms,
                                round_bf16, stream);
}

int lp_gemm_bf16_tc_affine(const void* x_terms, int nterms, int M, const void* w_bf16, int N, int K, const float* bias,
                           const float* out_bias, const float* out_scale, int epilogue, const float* residual, float* out_f32,
                           void* out_bf16, int out_terms, int round_bf16, void* stream) {
  if (!x_terms || !w_bf16 || M <= 0 || N <= 0 || K <= 0 || nterms < 1 || nterms > 3) return LP_ERR_INVALID_ARG;
  if (!out_f32 && !out_bf16) return LP_ERR_INVALID_ARG;
  if (out_bf16 && (out_terms < 1 || out_terms > 3)) return LP_ERR_INVALID_ARG;
  if (epilogue < LP_EPI_NONE || epilogue > LP_EPI_RESIDUAL) return LP_ERR_INVALID_ARG;
  if (epilogue == LP_EPI_RESIDUAL && !residual) return LP_ERR_INVALID_ARG;
  if (K % 8 || N % 8) return LP_ERR_UNSUPPORTED;  // 16-byte global strides (ragged N / K tiles are zero-filled by TMA)
  if ((reinterpret_cast<uintptr_t>(x_terms) & 15) || (reinterpret_cast<uintptr_t>(w_bf16) & 15)) return LP_ERR_UNSUPPORTED;
  if (M <= 64 && K % lp::TC_BK == 0) {
    // decode batches: swap-AB (weights = 128-row A operand, batch = N of the MMA), see gemm_tc_swap_kernel
    lp::TcSwapParams q;
    q.bias = bias;
    q.out_bias = out_bias;
    q.out_scale = out_scale;
    q.residual = residual;
    q.out_f32 = out_f32;
    q.out_bf = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    q.M = M;
    q.N = N;
    q.K = K;
    q.epi = epilogue;
    q.round_bf16 = round_bf16;
    q.nterms = nterms;
    q.out_terms = out_terms;
    q.NB = (M + 15) / 16 * 16;
    q.fuse = (M % 16 == 0) ? 1 : 0;
    // K-blocks per copy: 4 (512 contiguous bytes per weight row) if two stages fit, else 2
    q.KB = 2 * 4 * (lp::TC_BM * lp::TC_BK * 2 + nterms * q.NB * lp::TC_BK * 2) <= 212 * 1024 ? 4 : 2;
    {
      static const int kb_env = [] {  // tuning aid: LP_SWAP_KB = 1 | 2 | 4 forces the K-blocks per stage
        const char* e = getenv("LP_SWAP_KB");
        return e ? atoi(e) : 0;
      }();
      if (kb_env == 1 || kb_env == 2 || kb_env == 4) q.KB = kb_env;
    }
    q.ksplit = 1;
    q.streamk = 0;
    const int tiles_n = (N + lp::TC_BM - 1) / lp::TC_BM, nk = (K / lp::TC_BK + q.KB - 1) / q.KB;
    if (epilogue == LP_EPI_RESIDUAL && residual == out_f32 && !out_bf16 && !round_bf16) {
      // x += W . u in place: partial sums are added atomically, so the work can be cut anywhere along K.  The (tile, K-stage)
      // sequence is split into one equal range per SM (stream-K): every SM streams the same number of weight bytes whatever the
      // tile count (32 tiles of n_embd rows, 96 of a QKV matrix, ...).  LP_SWAP_STREAMK=0: the older unit split (A/B).
      static const bool sk_env = [] {
        const char* e = getenv("LP_SWAP_STREAMK");
        return !(e && e[0] == '0');
      }();
      if (sk_env && tiles_n < lp::num_sms() && (long long)tiles_n * nk >= lp::num_sms()) {
        q.streamk = 1;
      } else if (tiles_n < lp::num_sms()) {
        int ks = (lp::num_sms() + tiles_n - 1) / tiles_n;
        while (ks > 1 && nk / ks < 2) --ks;
        q.ksplit = ks > 8 ? 8 : ks;
      }
    }
    const CUtensorMap* mx = lp::tc_cached_map3(x_terms, nterms * M, K, q.fuse ? nterms * q.NB : q.NB, q.KB);
    const CUtensorMap* mw = lp::tc_cached_map3(w_bf16, N, K, lp::TC_BM, q.KB);
    if (mx && mw) {
      const int rc = lp::tc_launch_swap(*mx, *mw, q, stream);
      if (rc != LP_ERR_UNSUPPORTED) return rc;
    }
  }
  // Tile width: 256 halves the activation re-reads (prefill); decode batches (one M tile) are weight-streaming bound and
  // need ~one tile per SM, so the width shrinks until the grid fills the chip.  Ragged last tiles are zero-filled by TMA.
  const int tiles_m = (M + lp::TC_BM - 1) / lp::TC_BM;
  int BN = 32;
  for (int bn : {256, 128, 64}) {
    if (bn == 256 && nterms > 2) continue;  // shared-memory budget of a stage
    if ((long long)tiles_m * ((N + bn - 1) / bn) >= (long long)lp::num_sms() * 4 / 5) {
      BN = bn;
      break;
    }
  }
  const CUtensorMap* mx = lp::tc_cached_map(x_terms, nterms * M, K, lp::TC_BM);
  // W-tile multicast over CTA pairs: prefill-sized problems whose M tiles pair up inside one N tile
  static const bool cl_env = [] {
    const char* e = getenv("LP_GEMM_CLUSTER");
    return !(e && e[0] == '0');
  }();
  // CTA pairs (default; lp_set_gemm_pair(0) / LP_GEMM_PAIR=0 turn them off): 1393 vs 1302 TFLOP/s on 2048 x 18176 x 4544
  if (lp::g_gemm_pair.load() && BN == 256 && (long long)((M + 255) / 256) * ((N + 255) / 256) * 2 >= lp::num_sms()) {
    // prefill-sized: CTA pairs (cta_group::2), 256 x 256 tiles
    const CUtensorMap* mx2 = lp::tc_cached_map(x_terms, nterms * M, K, lp::TC_BM);
    const CUtensorMap* mw2 = lp::tc_cached_map(w_bf16, N, K, 128);
    if (mx2 && mw2) {
      lp::TcParams q;
      q.bias = bias; q.out_bias = out_bias; q.out_scale = out_scale; q.residual = residual; q.out_f32 = out_f32; q.out_bf = reinterpret_cast<__nv_bfloat16*>(out_bf16);
      q.M = M; q.N = N; q.K = K; q.epi = epilogue; q.round_bf16 = round_bf16; q.nterms = nterms; q.out_terms = out_terms;
      q.debug = 0; q.l2_ahead = 0;
      return lp::tc_launch_pair(*mx2, *mw2, q, stream);
    }
  }
  const int tiles_n = (N + BN - 1) / BN;
  const bool pair = cl_env && BN >= 128 && tiles_m % 2 == 0 && (long long)tiles_m * tiles_n >= 2LL * lp::num_sms();
  const CUtensorMap* mw = lp::tc_cached_map(w_bf16, N, K, pair ? BN / 2 : BN);
  if (!mx || !mw) return LP_ERR_UNSUPPORTED;
  lp::TcParams p;
  p.bias = bias;
  p.out_bias = out_bias;
  p.out_scale = out_scale;
  p.residual = residual;
  p.out_f32 = out_f32;
  p.out_bf = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.M = M;
  p.N = N;
  p.K = K;
  p.epi = epilogue;
  p.round_bf16 = round_bf16;
  p.nterms = nterms;
  p.out_terms = out_terms;
  {
    static const int dbg = [] {
      const char* e = getenv("LP_GEMM_DEBUG");
      return e ? atoi(e) : 0;
    }();
    p.debug = dbg;
    static const int ahead = [] {
      const char* e = getenv("LP_GEMM_L2AHEAD");
      return e ? atoi(e) : 0;
    }();
    p.l2_ahead = ahead;
  }
  if (pair) return BN == 256 ? lp::tc_launch<256, 2>(*mx, *mw, p, stream) : lp::tc_launch<128, 2>(*mx, *mw, p, stream);
  switch (BN) {
    case 256: return lp::tc_launch<256, 1>(*mx, *mw, p, stream);
    case 128: return lp::tc_launch<128, 1>(*mx, *mw, p, stream);
    case 64: return lp::tc_launch<64, 1>(*mx, *mw, p, stream);
    default: return lp::tc_launch<32, 1>(*mx, *mw, p, stream);
  }
}

}  // extern "C"
