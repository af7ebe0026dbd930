// Weight-streaming linear layer, "stream family": y = x[M,K] . W[N,K]^T for the decode step (M <= 8), built to keep
// HBM busy across kernel boundaries.  Judged on GB/s.
//
//   * PERSISTENT grid (one CTA per SM); each CTA owns a contiguous range of 16-row weight tiles;
//   * a PRODUCER thread streams the weights with TMA tensor copies (cp.async.bulk.tensor, completion on an mbarrier)
//     into a ring of stages.  The row-major weight matrix [N, K] is described to the TMA unit as a 3-D tensor
//     {128 bytes, N rows, K-blocks} so that ONE copy brings a whole stage (16 rows x 8 K-blocks = 16 KB) and lays it out
//     as [K-block][row][128 B] with the hardware 128-byte swizzle: every ldmatrix / 128-bit fragment load is bank-conflict
//     free.  (Measured: 1 KB bulk copies cap an SM at ~18 GB/s whatever the ring depth — the copy count, not the bytes in
//     flight, is the limit; 16 KB copies lift it.)  The ring is filled BEFORE griddepcontrol.wait: under PDL the first
//     stages of layer n+1's weights are already in flight while layer n drains;
//   * sixteen CONSUMER warps split K inside a stage.  The multiply-accumulates run on the tensor cores
//     (mma.sync.m16n8k16, bf16 x bf16 -> fp32) so the issue slots are left for what an int4 GEMV is limited by — nibble
//     unpacking: nibble -> bf16 is ONE lop3 per two weights ((w & 0x000F000F) | 0x43004300 = {128+q_lo, 128+q_hi}); the
//     +128 and the GPTQ zero point are removed algebraically per 128-column group:
//          sum (q - z) s x = s * (sum (128+q) x - (128 + z) * sum x);
//   * x (the M activation rows) is the 8- or 16-column B operand, staged once per CTA in shared memory, optionally
//     through a fused LayerNorm / RMSNorm.  To keep fp32 ACTIVATION accuracy every row is split into bf16 terms
//     x = hi + mid (+ lo) in separate columns (products with the bf16 / int4 weights are exact in the fp32
//     accumulator); the columns are re-added in the epilogue.  bf16-faithful mode has bf16-valued x: one column per row;
//   * per tile: cross-warp reduction through shared memory, then the fused epilogue (bias, GELU, SwiGLU, residual).
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "stream_common.cuh"

namespace lp {

struct GsParams {
  const float* x;
  const float* residual;
  float* out;
  lp_weight W;
  NormArgs nrm;
  int M, split, epi, round_bf16;
  int ldx;          // bf16 elements per staged x row
  int nkb;          // K-blocks per row = ceil(row bytes / 128)
  int nks;          // stages per tile = ceil(nkb / GS_KB)
  int ntiles;       // N / 16
  int nstages;      // ring depth
  int stage_stride; // bytes (multiple of 1024: the 128-byte swizzle pattern is anchored at 1 KB)
  int tma_rank;     // 3: one copy per stage; 2: one copy per K-block (fallback if the 3-D map was refused)
  int ngroups;      // int4 groups per row
  int gp128;        // 128-column chunks per scale group (group / 128)
  unsigned long long* trace;  // debug: per-CTA phase timestamps (globaltimer ns), NULL in production
};

// One activation row -> shared-memory B-operand columns (see the call site).  s_stat: [2][GS_CWARPS] floats.
template <int FMT, bool CACHED, int NIC = 4>
__device__ __forceinline__ void gs_stage_row(const GsParams& p, int m, int NCOL, int ncols, float* s_stat, uint16_t* xs, signed char* xs8,
                                             float* xsum, float* colscale) {
  constexpr int STRIDE = GS_CWARPS * 32 * 4;
  constexpr int NI = CACHED ? NIC : 1;
  const int K = p.W.K;
  const int nch128 = (K + 127) / 128;
  const int niter = (nch128 * 128 + STRIDE - 1) / STRIDE;  // covers the zero padding up to a whole 128-column chunk
  const int ctid = threadIdx.x, warp = ctid >> 5, lane = ctid & 31;
  const float* xr = p.x + (size_t)m * K;
  const bool has_norm = p.nrm.kind >= 0;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 xc[NI], wc[NI], bc[NI];
  auto ld = [&](const float* base, int i) { const int k = ctid * 4 + i * STRIDE; return (base && k < K) ? *reinterpret_cast<const float4*>(base + k) : zero4; };
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      xc[i] = i < niter ? ld(xr, i) : zero4;
      wc[i] = (has_norm && i < niter) ? ld(p.nrm.w, i) : zero4;
      bc[i] = (has_norm && i < niter) ? ld(p.nrm.b, i) : zero4;
    }
  }
#define GS_ITERS(i) _Pragma("unroll") for (int i = 0; i < (CACHED ? NI : niter); ++i) if (!CACHED || i < niter)
  auto raw = [&](int i) { return CACHED ? xc[CACHED ? i : 0] : ld(xr, i); };
  auto block_reduce = [&](float a, float b, bool is_max, float& ra, float& rb) {
    a = is_max ? warp_max(a) : warp_sum(a);
    b = warp_sum(b);
    gs_bar_consumers();
    if (lane == 0) {
      s_stat[warp] = a;
      s_stat[GS_CWARPS + warp] = b;
    }
    gs_bar_consumers();
    ra = is_max ? 0.f : 0.f;
    rb = 0.f;
#pragma unroll
    for (int w = 0; w < GS_CWARPS; ++w) {
      ra = is_max ? fmaxf(ra, s_stat[w]) : ra + s_stat[w];
      rb += s_stat[GS_CWARPS + w];
    }
  };
  float mean = 0.f, rstd = 1.f;
  if (has_norm) {
    float sm = 0.f, ss = 0.f;
    GS_ITERS(i) {
      const float4 v = raw(i);
      sm += v.x + v.y + v.z + v.w;
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    block_reduce(sm, ss, false, sm, ss);
    if (p.nrm.kind == LP_NORM_LAYERNORM) {
      mean = sm / (float)K;
      float v2 = 0.f, dummy;  // two-pass variance
      GS_ITERS(i) {
        const int k = ctid * 4 + i * STRIDE;
        if (k < K) {
          const float4 v = raw(i);
          const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
          v2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      }
      block_reduce(v2, 0.f, false, v2, dummy);
      rstd = 1.0f / sqrtf(v2 / (float)K + p.nrm.eps);
    } else {
      rstd = 1.0f / sqrtf(ss / (float)K + p.nrm.eps);
    }
  }
  // normalised values of this thread's 4 columns of iteration i (zeros beyond K)
  auto nv = [&](int i, float (&v)[4]) {
    const float4 x4 = raw(i);
    v[0] = x4.x; v[1] = x4.y; v[2] = x4.z; v[3] = x4.w;
    if (has_norm) {
      const float4 w4 = CACHED ? wc[CACHED ? i : 0] : ld(p.nrm.w, i);
      const float4 b4 = CACHED ? bc[CACHED ? i : 0] : ld(p.nrm.b, i);
      if (p.nrm.kind == LP_NORM_LAYERNORM) {
        v[0] = (v[0] - mean) * rstd * w4.x + b4.x; v[1] = (v[1] - mean) * rstd * w4.y + b4.y;
        v[2] = (v[2] - mean) * rstd * w4.z + b4.z; v[3] = (v[3] - mean) * rstd * w4.w + b4.w;
      } else {
        v[0] = w4.x * (v[0] * rstd); v[1] = w4.y * (v[1] * rstd);
        v[2] = w4.z * (v[2] * rstd); v[3] = w4.w * (v[3] * rstd);
      }
      if (ctid * 4 + i * STRIDE >= K) v[0] = v[1] = v[2] = v[3] = 0.f;  // LayerNorm bias must not leak into the padding
    }
  };
  if constexpr (FMT == LP_W_INT4) {
    // int4 weights run on the INTEGER tensor cores (IMMA u8 x s8 -> s32).  The activation row becomes block fixed
    // point: X_k = rint(x_k * 2^22 / max|x|), written as three balanced base-256 digits (int8) in three B columns.
    // Integer dot products are exact; the digits' weights (max|x| / 2^22 * 256^i) are applied in the epilogue, so the
    // result carries ~22 bits of the largest activation: fp32-activation accuracy.
    float amax = 0.f, dummy;
    GS_ITERS(i) {
      float v[4];
      nv(i, v);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
    }
    block_reduce(amax, 0.f, true, amax, dummy);
    const float inv = amax > 0.f ? 4194304.0f / amax : 0.f;
    if (ctid < 3) colscale[m * 3 + ctid] = (amax / 4194304.0f) * (ctid == 0 ? 1.0f : (ctid == 1 ? 256.0f : 65536.0f));
    GS_ITERS(i) {
      const int c = warp + GS_CWARPS * i;  // 128-column chunk: columns 128 c + 4 lane = 4 ctid + STRIDE i
      if (c < nch128) {
        float v[4];
        nv(i, v);
        // The 8 columns a lane feeds to one IMMA are kc = 32 tt + 8 j + e of the chunk (tt = MMA k-lane, j = word).  They
        // are stored at byte 32 tt + 8 j + (e >> 1) + 4 (e & 1): even nibbles (operand a0/a1) first, odd nibbles (a2/a3) after.
        const int kc = lane * 4, tt = kc >> 5, j = (kc >> 3) & 3, e0 = kc & 7;
        int sum[3] = {0, 0, 0};
        signed char* d8 = xs8 + c * 128 + tt * 32 + j * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int X = __float2int_rn(v[q] * inv);
          const int e = e0 + q, pos = (e >> 1) + 4 * (e & 1);
#pragma unroll
          for (int dgt = 0; dgt < 3; ++dgt) {
            const int tdig = dgt < 2 ? (((X + 128) & 255) - 128) : X;
            X = (X - tdig) >> 8;
            sum[dgt] += tdig;
            d8[(size_t)(m * 3 + dgt) * p.ldx + pos] = (signed char)tdig;
          }
        }
#pragma unroll
        for (int dgt = 0; dgt < 3; ++dgt) {
          const float ps = warp_sum((float)sum[dgt]);
          if (lane == 0) xsum[c * NCOL + m * 3 + dgt] = ps;
        }
        if (m == 0 && lane >= ncols && lane < NCOL) xsum[c * NCOL + lane] = 0.f;
      }
    }
  } else {
    GS_ITERS(i) {
      const int k = ctid * 4 + i * STRIDE;
      if (k < K) {
        float v[4];
        nv(i, v);
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
          if (sp < p.split) {
            uint16_t hb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              hb[q] = gs_bf16_bits(v[q]);
              v[q] -= __uint_as_float((uint32_t)hb[q] << 16);  // exact: next term of the split
            }
            uint16_t* dst = xs + (size_t)(m * p.split + sp) * p.ldx;
            *reinterpret_cast<uint2*>(dst + k) = make_uint2(hb[0] | ((uint32_t)hb[1] << 16), hb[2] | ((uint32_t)hb[3] << 16));
          }
        }
      }
    }
    if (ctid < p.split) colscale[m * p.split + ctid] = 1.0f;
  }
#undef GS_ITERS
}

template <int FMT, int NB>
__global__ void __launch_bounds__(GS_THREADS, 1) linear_stream_kernel(const __grid_constant__ CUtensorMap tmap, const GsParams p) {
  constexpr int NCOL = 8 * NB;
  constexpr int COLS_PER_BLK = (FMT == LP_W_BF16) ? 64 : 256;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int K = p.W.K, N = p.W.N;
  const int ncols = p.M * p.split;
  const int nch128 = (K + 127) / 128;
  // ---- shared memory carve-up: [ring (1 KB aligned)] [barriers] [red] [xsum] [xs] ----
  unsigned char* ring = smem;
  unsigned char* after = smem + (size_t)p.nstages * p.stage_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after);                      // full[nstages], empty[nstages]
  volatile int* done = reinterpret_cast<volatile int*>(after + 240);         // [2] last finalised local tile per parity
  float* red = reinterpret_cast<float*>(after + 256);                       // [2][CWARPS][16][NCOL]
  float* colscale = red + 2 * GS_CWARPS * 16 * NCOL;                        // [NCOL] weight of each B column in the epilogue
  float* xsum = colscale + NCOL;                                            // [nch128][NCOL]   (int4 only)
  // bf16 weights: [ncols][ldx] bf16 bits;  int4 weights: [ncols][ldx] int8 digits (ldx in elements of either kind)
  uint16_t* xs = reinterpret_cast<uint16_t*>(xsum + (FMT == LP_W_INT4 ? nch128 * NCOL : 0));
  signed char* xs8 = reinterpret_cast<signed char*>(xs);
  __shared__ float s_stat[2][GS_CWARPS];
  (void)xs8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = gs_smem_u32(bars);
  const uint32_t ring_u32 = gs_smem_u32(ring);
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (p.nstages + s); };

  // Tiles of this CTA: contiguous range of WHOLE tiles.  A stage-granular (stream-K) split with cross-CTA partial sums was
  // tried (git history: "stream-K (opt-in)"): perfectly balanced, but 2.03 vs 1.74 ms per stablelm-3b step and 2.70 vs 2.13 ms
  // per 7B-int4 step on B200 — a CTA that starts late then also delays its neighbour's tile — so it was dropped.  A
  // dynamic tile scheduler (static tiles for the ring depth, atomic claims afterwards) was tried too: no gain (1.78 vs 1.74
  // ms), because for the small layers the whole per-CTA work fits inside the prefetch ring, i.e. is static anyway.
  const int tile_begin = (int)(((long long)p.ntiles * blockIdx.x) / gridDim.x);
  const int tile_end = (int)(((long long)p.ntiles * (blockIdx.x + 1)) / gridDim.x);
  const int nunits = (tile_end - tile_begin) * p.nks;
  const int aux_bytes = (p.W.flags & LP_WF_AUX_PACKED) ? 4 : 8;

  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 0] = gs_now();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), GS_CWARPS);
    }
    done[0] = done[1] = -1;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == GS_CWARPS) {
    // =========================== PRODUCER: weights do not depend on the previous kernel -> no griddepcontrol.wait ====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmap) : "memory");
      int s = 0, ph = 0, ks = 0, tile = tile_begin;
      for (int u = 0; u < nunits; ++u) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t dst = ring_u32 + (uint32_t)s * p.stage_stride;
        uint32_t aux_len = 0;
        int g_begin = 0;
        if constexpr (FMT == LP_W_INT4) {
          // scale/zero words of this tile for the groups that intersect this stage: [tile][group][16 rows]
          const int c_begin = ks * GS_KB * 2, c_end = min(nch128, (ks + 1) * GS_KB * 2);
          g_begin = c_begin / p.gp128;
          aux_len = (uint32_t)((c_end + p.gp128 - 1) / p.gp128 - g_begin) * 16 * aux_bytes;
        }
        mbar_expect_tx(full_bar(s), GS_KB * GS_BLK_BYTES + aux_len);
        if (p.tma_rank == 3) {
          tma_load_3d(dst, &tmap, 0, tile * GS_ROWS, ks * GS_KB, full_bar(s));
        } else {
          for (int kb = 0; kb < GS_KB; ++kb)
            tma_load_2d(dst + kb * GS_BLK_BYTES, &tmap, (ks * GS_KB + kb) * (FMT == LP_W_BF16 ? 64 : 128), tile * GS_ROWS, full_bar(s));
        }
        if constexpr (FMT == LP_W_INT4) {
          const char* src = reinterpret_cast<const char*>(p.W.aux2) + ((size_t)tile * p.ngroups + g_begin) * 16 * aux_bytes;
          bulk_g2s(dst + GS_KB * GS_BLK_BYTES, src, aux_len, full_bar(s));
        }
        if (++s == p.nstages) { s = 0; ph ^= 1; }
        if (++ks == p.nks) { ks = 0; ++tile; }
      }
    }
    return;
  }

  // =============================== CONSUMERS ========================================================================
  pdl_wait();  // activations of the previous kernel are visible from here on
  pdl_launch_dependents();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 1] = gs_now();
  const int g = lane >> 2, t = lane & 3;

  // ---- stage x: (optional norm) -> B-operand columns in shared memory ----------------------------------------------
  // Thread `ctid` owns columns 4 ctid + 2048 i (.. + 3) of every pass (statistics, max, conversion), so for rows of up to
  // 8192 columns the activations (and norm parameters) are read from global memory ONCE and stay in registers; every
  // later pass costs only a named barrier.
  if ((nch128 * 128 + 2047) / 2048 <= 4) {
    for (int m = 0; m < p.M; ++m) gs_stage_row<FMT, true>(p, m, NCOL, ncols, &s_stat[0][0], xs, xs8, xsum, colscale);
  } else {  // (a 6-deep register cache for 11008-column rows was tried: it spills and slows the main loop)
    for (int m = 0; m < p.M; ++m) gs_stage_row<FMT, false>(p, m, NCOL, ncols, &s_stat[0][0], xs, xs8, xsum, colscale);
  }
  gs_bar_consumers();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 2] = gs_now();

  // ---- main loop ----------------------------------------------------------------------------------------------------
  float acc[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] = 0.f;
  const uint16_t* xrow[NB];
  const signed char* xrow8[NB];
  bool bvalid[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    bvalid[nb] = (g + 8 * nb) < ncols;
    xrow[nb] = xs + (size_t)(bvalid[nb] ? g + 8 * nb : 0) * p.ldx;
    xrow8[nb] = xs8 + (size_t)(bvalid[nb] ? g + 8 * nb : 0) * p.ldx;
  }
  (void)xrow; (void)xrow8;
  // int4: MMA row g <-> tile row pr0 (and g + 8 <-> pr0 + 8).  Neighbouring g get rows 4 apart, so under the 128-byte
  // swizzle (16-byte unit ^= row & 7) the two rows of a quarter warp read opposite halves of a 128-byte line.
  const int pr0 = (FMT == LP_W_INT4) ? ((g >> 1) + 4 * (g & 1)) : g;

  int s = 0, ph = 0, ks = 0, tile = tile_begin;
  for (int u = 0; u < nunits; ++u) {
    mbar_wait(full_bar(s), ph);
    if (p.trace && threadIdx.x == 0 && u < 4) p.trace[blockIdx.x * 8 + 3 + u] = gs_now();
    const uint32_t st = ring_u32 + (uint32_t)s * p.stage_stride;
    constexpr int NSUB = 2 * GS_KB / GS_CWARPS;  // halves of a K-block per warp: 1 (16 warps) or 2 (8 warps)
    const int kbl = NSUB == 1 ? warp >> 1 : warp;  // K-block inside the stage
    const int sub0 = NSUB == 1 ? (warp & 1) : 0;
    const int kb = ks * GS_KB + kbl;
    if (kb < p.nkb) {
      if constexpr (FMT == LP_W_BF16) {
        const int kcol = kb * COLS_PER_BLK;
        const uint32_t blk = st + kbl * GS_BLK_BYTES;
        const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int kk2 = 0; kk2 < 2 * NSUB; ++kk2) {
          const int kk = sub0 * 2 + kk2;
          uint32_t a[4];
          gs_ldsm_x4(a, blk + row * 128 + (((kk * 2 + (lane >> 4)) ^ (row & 7)) << 4));
          const int k0 = kcol + kk * 16;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            uint32_t b0 = 0, b1 = 0;
            if (bvalid[nb]) {
              b0 = *reinterpret_cast<const uint32_t*>(xrow[nb] + k0 + 2 * t);
              b1 = *reinterpret_cast<const uint32_t*>(xrow[nb] + k0 + 8 + 2 * t);
            }
            gs_mma(acc[nb], a[0], a[1], a[2], a[3], b0, b1);
          }
        }
      } else {
        const unsigned char* blk = ring + (size_t)s * p.stage_stride + kbl * GS_BLK_BYTES;
        const unsigned char* auxp = ring + (size_t)s * p.stage_stride + GS_KB * GS_BLK_BYTES;
        const int g_begin = p.gp128 == 1 ? ks * GS_KB * 2 : (ks * GS_KB * 2) / p.gp128;
#pragma unroll
        for (int cs = 0; cs < NSUB; ++cs) {
          const int cc = sub0 + cs;
          const int c = kb * 2 + cc;  // 128-column chunk inside the row
          if (c < nch128) {
            const uint4 wa4 = *reinterpret_cast<const uint4*>(blk + pr0 * 128 + (((cc * 4 + t) ^ (pr0 & 7)) << 4));
            const uint4 wb4 = *reinterpret_cast<const uint4*>(blk + (pr0 + 8) * 128 + (((cc * 4 + t) ^ (pr0 & 7)) << 4));
            const uint32_t wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w};
            const uint32_t wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
            const int gl = (p.gp128 == 1 ? c : c / p.gp128) - g_begin;  // group inside the stage's aux block
            float s0, s1, z0, z1;
            if (p.W.flags & LP_WF_AUX_PACKED) {  // bf16 scale << 16 | bf16 zero
              const uint32_t u0 = *reinterpret_cast<const uint32_t*>(auxp + ((size_t)gl * 16 + pr0) * 4);
              const uint32_t u1 = *reinterpret_cast<const uint32_t*>(auxp + ((size_t)gl * 16 + pr0 + 8) * 4);
              s0 = __uint_as_float(u0 & 0xffff0000u);
              s1 = __uint_as_float(u1 & 0xffff0000u);
              z0 = __uint_as_float(u0 << 16);
              z1 = __uint_as_float(u1 << 16);
            } else {
              const float2 a0 = *reinterpret_cast<const float2*>(auxp + ((size_t)gl * 16 + pr0) * 8);
              const float2 a1 = *reinterpret_cast<const float2*>(auxp + ((size_t)gl * 16 + pr0 + 8) * 8);
              s0 = a0.x; z0 = a0.y; s1 = a1.x; z1 = a1.y;
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              int ca[4] = {0, 0, 0, 0}, cb[4] = {0, 0, 0, 0};
              uint4 xv0 = make_uint4(0, 0, 0, 0), xv1 = xv0;
              if (bvalid[nb]) {  // 32 digit bytes: this k-lane's 32 columns of the chunk, in operand order
                xv0 = *reinterpret_cast<const uint4*>(xrow8[nb] + c * 128 + t * 32);
                xv1 = *reinterpret_cast<const uint4*>(xrow8[nb] + c * 128 + t * 32 + 16);
              }
              const uint32_t M4 = 0x0F0F0F0Fu;
              gs_imma(ca, wa[0] & M4, wb[0] & M4, (wa[0] >> 4) & M4, (wb[0] >> 4) & M4, xv0.x, xv0.y);
              gs_imma(cb, wa[1] & M4, wb[1] & M4, (wa[1] >> 4) & M4, (wb[1] >> 4) & M4, xv0.z, xv0.w);
              gs_imma(ca, wa[2] & M4, wb[2] & M4, (wa[2] >> 4) & M4, (wb[2] >> 4) & M4, xv1.x, xv1.y);
              gs_imma(cb, wa[3] & M4, wb[3] & M4, (wa[3] >> 4) & M4, (wb[3] >> 4) & M4, xv1.z, xv1.w);
              const float2 xsv = *reinterpret_cast<const float2*>(xsum + c * NCOL + nb * 8 + 2 * t);
              // sum (q - z) s x = s * (sum q X - z * sum X), all integers exact in fp32 (|.| < 2^24)
              acc[nb][0] = fmaf(s0, fmaf(-z0, xsv.x, (float)(ca[0] + cb[0])), acc[nb][0]);
              acc[nb][1] = fmaf(s0, fmaf(-z0, xsv.y, (float)(ca[1] + cb[1])), acc[nb][1]);
              acc[nb][2] = fmaf(s1, fmaf(-z1, xsv.x, (float)(ca[2] + cb[2])), acc[nb][2]);
              acc[nb][3] = fmaf(s1, fmaf(-z1, xsv.y, (float)(ca[3] + cb[3])), acc[nb][3]);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));  // this warp is done with the stage
    if (++s == p.nstages) { s = 0; ph ^= 1; }

    if (++ks == p.nks) {
      ks = 0;
      // ---- tile finished: cross-warp reduction, column recombination, epilogue ----
      // Only ONE warp (rotating with the tile) waits for the others and finalises; everybody else drops its partial sums,
      // arrives on the named barrier and moves on to the next tile.  `red` is double buffered by tile parity: before
      // overwriting a buffer a warp checks that the tile that used it two tiles ago has been finalised (`done`).
      const int lt = tile - tile_begin, par = lt & 1;
      if (lt >= 2) {
        while (done[par] < lt - 2) {}
      }
      float* r = red + ((size_t)par * GS_CWARPS + warp) * 16 * NCOL;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        r[pr0 * NCOL + nb * 8 + 2 * t] = acc[nb][0];
        r[pr0 * NCOL + nb * 8 + 2 * t + 1] = acc[nb][1];
        r[(pr0 + 8) * NCOL + nb * 8 + 2 * t] = acc[nb][2];
        r[(pr0 + 8) * NCOL + nb * 8 + 2 * t + 1] = acc[nb][3];
        acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
      }
      const int fin = lt % GS_CWARPS;        // first finalising warp of this tile (rotates)
      const int nfin = p.M < GS_CWARPS ? p.M : GS_CWARPS;  // one finalising warp per activation row (or several rows each)
      const int fidx = (warp - fin + GS_CWARPS) % GS_CWARPS;
      if (fidx >= nfin) {
        // barrier ids are immediates on purpose: a register id makes ptxas reserve all 16 named barriers of the SM for
        // one CTA, which forbids two of these kernels to be co-resident (PDL overlap)
        if (par == 0) asm volatile("bar.arrive 2, %0;\n" ::"n"(GS_CWARPS * 32) : "memory");
        else asm volatile("bar.arrive 3, %0;\n" ::"n"(GS_CWARPS * 32) : "memory");
      } else {
        if (par == 0) asm volatile("bar.sync 2, %0;\n" ::"n"(GS_CWARPS * 32) : "memory");
        else asm volatile("bar.sync 3, %0;\n" ::"n"(GS_CWARPS * 32) : "memory");
        // lane = (row of the tile, half): each half adds the partial sums of 8 warps per B column, then one shuffle
        const int rr = lane & 15, half = lane >> 4;
        for (int m = fidx; m < p.M; m += nfin) {
        const float* rb = red + ((size_t)par * GS_CWARPS + half * (GS_CWARPS / 2)) * 16 * NCOL + rr * NCOL + m * p.split;
        float cs[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < GS_CWARPS / 2; ++w) {
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2)
            if (s2 < p.split) cs[s2] += rb[(size_t)w * 16 * NCOL + s2];
        }
        float y = 0.f;
#pragma unroll
        for (int si = 0; si < 3; ++si) {  // smallest terms first (bf16 split: last term; int8 digits: digit 0)
          if (si < p.split) {
            const int s2 = (FMT == LP_W_INT4) ? si : p.split - 1 - si;
            float c = 0.f;
#pragma unroll
            for (int q = 0; q < 3; ++q) c = (q == s2) ? cs[q] : c;
            c += __shfl_xor_sync(0xffffffffu, c, 16);
            y = fmaf(c, colscale[m * p.split + s2], y);
          }
        }
        const int row = tile * GS_ROWS + rr;
        if (p.W.bias) y += p.W.bias[row];
        y = maybe_round(y, p.round_bf16);
        y = out_affine(y, p.W.out_bias, p.W.out_scale, row, p.round_bf16);
        if (p.epi == LP_EPI_SWIGLU) {
          const float other = __shfl_xor_sync(0xffffffffu, y, 1);  // fc_2 row of the pair
          if (half == 0 && (rr & 1) == 0) {
            const float a = maybe_round(silu(y), p.round_bf16);
            p.out[(size_t)m * (N / 2) + (row >> 1)] = maybe_round(a * other, p.round_bf16);
          }
        } else if (half == 0) {
          if (p.epi == LP_EPI_GELU) y = maybe_round(gelu_erf(y), p.round_bf16);
          else if (p.epi == LP_EPI_RESIDUAL) y = maybe_round(p.residual[(size_t)m * N + row] + y, p.round_bf16);
          p.out[(size_t)m * N + row] = y;
        }
        }
        if (nfin > 1) {  // all finalisers have read `red`
          if (par == 0) asm volatile("bar.sync 4, %0;\n" ::"r"(nfin * 32) : "memory");
          else asm volatile("bar.sync 5, %0;\n" ::"r"(nfin * 32) : "memory");
        }
        __syncwarp();
        if (fidx == 0 && lane == 0) {
          __threadfence_block();
          done[par] = lt;
        }
      }
      ++tile;
    }
  }
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 8 + 7] = gs_now();
}

// ------------------------------------------------------------------------------------------------ host side
// TMA descriptors of the weight matrices, built once per (pointer, shape) and cached.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

const GsMap& gs_tensor_map(const lp_weight& W, size_t row_bytes) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int>, GsMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(W.w, W.N, W.K, W.fmt);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  GsMap m;
  memset(&m, 0, sizeof(m));
  EncodeTiledFn enc = encode_fn();
  if (enc) {
    const bool bf = W.fmt == LP_W_BF16;
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    const cuuint64_t inner = bf ? 64 : 128;  // elements per 128 bytes
    const cuuint64_t nkb = (row_bytes + 127) / 128;
    {  // 3-D view {128 B, N rows, K-blocks}: the K-block stride (128 B) is smaller than the row stride on purpose
      cuuint64_t dims[3] = {inner, (cuuint64_t)W.N, nkb};
      cuuint64_t strides[2] = {(cuuint64_t)row_bytes, 128};
      cuuint32_t box[3] = {(cuuint32_t)inner, GS_ROWS, GS_KB};
      cuuint32_t es[3] = {1, 1, 1};
      if (row_bytes % 128 == 0 &&
          enc(&m.map, dt, 3, const_cast<void*>(W.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        m.rank = 3;
    }
    if (m.rank == 0) {  // plain 2-D {row bytes, N}: one 2 KB copy per K-block
      cuuint64_t dims[2] = {(cuuint64_t)(bf ? W.K : row_bytes), (cuuint64_t)W.N};
      cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
      cuuint32_t box[2] = {(cuuint32_t)inner, GS_ROWS};
      cuuint32_t es[2] = {1, 1};
      if (enc(&m.map, dt, 2, const_cast<void*>(W.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        m.rank = 2;
    }
  }
  if (cache.size() > 4096) cache.clear();
  return cache.emplace(key, m).first->second;
}

static unsigned long long* g_trace = nullptr;
void set_stream_trace(void* buf) { g_trace = reinterpret_cast<unsigned long long*>(buf); }

static size_t gs_tail_smem(int fmt, int NB, int K, int ncols, int ldx) {
  const int NCOL = 8 * NB;
  return 256 + (size_t)2 * GS_CWARPS * 16 * NCOL * 4 + (size_t)NCOL * 4 + (fmt == LP_W_INT4 ? (size_t)((K + 127) / 128) * NCOL * 4 : 0) +
         (size_t)ncols * ldx * (fmt == LP_W_INT4 ? 1 : 2);
}

template <int FMT, int NB>
static int gs_launch(const CUtensorMap& map, const GsParams& p, size_t smem, int grid, void* stream) {
  static std::atomic<unsigned long long> attr_set{0};  // one bit per device
  auto kern = linear_stream_kernel<FMT, NB>;
  if (needs_device_setup(attr_set)) {
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    // keep the SM's L1/shared split at maximum shared memory: otherwise the carve-out is sized for ONE CTA of this kernel
    // and the next kernel's CTAs (PDL) cannot become resident before this one drains
    LP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    mark_device_setup(attr_set);
  }
  return launch(kern, dim3(grid), dim3(GS_THREADS), smem, stream, map, p);
}

// `nrm.kind` -1: no fused norm
int linear_stream(const float* x, int M, const lp_weight& W, const NormArgs& nrm, int epi, const float* residual, float* out,
                  int round_bf16, void* stream) {
  static_assert(2 * GS_KB == GS_CWARPS || GS_KB == GS_CWARPS, "one or two consumer warps per K-block of a stage");
  if (W.fmt != LP_W_BF16 && W.fmt != LP_W_INT4) return LP_ERR_UNSUPPORTED;
  if (W.N % GS_ROWS) return LP_ERR_UNSUPPORTED;
  const int K = W.K;
  if (K % 16) return LP_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(W.w) & 15) != 0) return LP_ERR_UNSUPPORTED;
  int split;
  if (W.fmt == LP_W_INT4) {
    // integer path: three int8 digits per activation row in every mode (a bf16-valued x is covered exactly as well)
    if (round_bf16 && nrm.kind >= 0) return LP_ERR_UNSUPPORTED;  // bf16-faithful norm roundings live in lp_norm
    if (M > 5) return LP_ERR_UNSUPPORTED;
    split = 3;
  } else if (round_bf16 && nrm.kind < 0) {
    if (M > 16) return LP_ERR_UNSUPPORTED;
    split = 1;  // x is bf16-valued already
  } else {
    if (round_bf16) return LP_ERR_UNSUPPORTED;  // bf16-faithful norm roundings live in lp_norm
    if (M > 8) return LP_ERR_UNSUPPORTED;
    split = M <= 2 ? 3 : 2;
    if (split == 3 && (size_t)M * 3 * K * 2 > 48 * 1024) split = 2;  // long rows: 16 mantissa bits per term pair are plenty
  }
  const int ncols = M * split;
  const int NB = ncols <= 8 ? 1 : 2;
  GsParams p = {};
  p.x = x; p.residual = residual; p.out = out; p.W = W; p.nrm = nrm;
  p.M = M; p.split = split; p.epi = epi; p.round_bf16 = round_bf16;
  p.trace = g_trace;
  const int kpad = (K + 255) / 256 * 256;
  size_t row_bytes;
  int aux_stage = 0;
  if (W.fmt == LP_W_BF16) {
    if (K % 64) return LP_ERR_UNSUPPORTED;  // 128-byte K-blocks
    row_bytes = (size_t)K * 2;
    p.ldx = kpad + 8;  // row stride = 16 bytes mod 128: B-fragment loads of a warp touch 32 distinct banks
  } else {
    if (W.group <= 0 || W.group % 128 || !W.aux2) return LP_ERR_UNSUPPORTED;
    row_bytes = (size_t)kpad / 2;  // lp_int4_row_bytes(K)
    p.ldx = kpad + 16;  // int8 digits; row stride = odd multiple of 16 bytes: neighbouring columns read opposite-parity units
    p.gp128 = W.group / 128;
    p.ngroups = (K + W.group - 1) / W.group;
    const int groups_per_stage = (GS_KB * 2 + p.gp128 - 1) / p.gp128 + 1;
    aux_stage = groups_per_stage * 16 * ((W.flags & LP_WF_AUX_PACKED) ? 4 : 8);
  }
  const GsMap& gm = gs_tensor_map(W, row_bytes);
  if (gm.rank == 0) return LP_ERR_UNSUPPORTED;
  p.tma_rank = gm.rank;
  p.nkb = (int)((row_bytes + 127) / 128);
  p.nks = (p.nkb + GS_KB - 1) / GS_KB;
  p.ntiles = W.N / GS_ROWS;
  p.stage_stride = (GS_KB * GS_BLK_BYTES + aux_stage + 1023) / 1024 * 1024;
  const size_t tail = gs_tail_smem(W.fmt, NB, K, ncols, p.ldx);
  static const size_t env_budget = [] {
    const char* e = getenv("LP_GS_BUDGET_KB");
    return e ? (size_t)atoi(e) * 1024 : GS_SMEM_BUDGET;
  }();
  size_t budget = env_budget;
  if (tail + 3 * (size_t)p.stage_stride + 1024 > budget) budget = 224 * 1024;  // long activation rows: give up co-residency
  if (tail + 3 * (size_t)p.stage_stride + 1024 > budget) return LP_ERR_UNSUPPORTED;
  int ns = (int)((budget - tail - 1024) / p.stage_stride);
  if (ns > GS_MAX_STAGES) ns = GS_MAX_STAGES;
  p.nstages = ns;
  const size_t smem = (size_t)ns * p.stage_stride + tail + 1024;  // + slack for the 1 KB alignment of the ring
  const int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
  if (W.fmt == LP_W_BF16)
    return NB == 1 ? gs_launch<LP_W_BF16, 1>(gm.map, p, smem, grid, stream) : gs_launch<LP_W_BF16, 2>(gm.map, p, smem, grid, stream);
  return NB == 1 ? gs_launch<LP_W_INT4, 1>(gm.map, p, smem, grid, stream) : gs_launch<LP_W_INT4, 2>(gm.map, p, smem, grid, stream);
}

}  // namespace lp
