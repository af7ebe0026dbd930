// The whole single-token decode step (batch 1) as ONE persistent kernel: every linear layer of every block, the attention
// of every block and the lm_head, behind a single TMA ring that never stops streaming.
//
//   reference: GPT.forward (model.py:63-111) -> Block.forward (158-180) -> CausalSelfAttention.forward (194-254, apply_rope
//   330-336, cache index_copy_ 236-245, scaled_dot_product_attention 256-275) -> MLP (284-301), for T == 1.
//
// Why: with one launch per op a replayed int4 Llama-2-7b layer spends 21 us streaming weights and 45 us on what surrounds a
// launch (activation staging, ramp-up / drain of the ring, tile skew, attention start-up; profiles/r1_results.md).  Weights
// and old K/V rows do not depend on the activations, so here
//   * one PRODUCER thread per CTA walks a static op table (device memory: TMA descriptor + shapes per op) and keeps a ring of
//     16 KB stages full: `cp.async.bulk.tensor.3d` weight stages (linear_stream.cu layout), 16 KB `cp.async.bulk` K / V tiles
//     for attention.  It only ever waits for a free ring slot: while the consumers sit at a grid-wide dependency the ring
//     fills with the NEXT ops' bytes, HBM stays busy across op boundaries;
//   * sixteen CONSUMER warps execute the ops in order.  An op that reads another op's output waits for that op's arrival
//     counter (one `red.release.gpu` per CTA, polled with `ld.acquire.gpu`), re-stages its activation row (fused LayerNorm /
//     RMSNorm, bf16 term split or int8 digit split as in linear_stream.cu) and drains the ring: mma.sync bf16 / IMMA int4;
//   * attention (MHA / GQA / MQA): q head h is split over P = min(4, #CTAs / H) CTAs along the sequence (the heads of a group
//     stream the same K / V tiles: L2 hits); RoPE(q, k_new), the cache append (global
//     + patch of the landed tile), q.K^T, online softmax and P.V run on the CUDA cores straight from the ring stages; the
//     (m, l, o) partials of the P splits are merged by the attention-projection's activation staging — no extra pass.
// The op table is built once per (model, cache) by lp_decode_step_plan; a step is lp_decode_step: a tiny prologue kernel
// (embedding row, counter reset) + this kernel.  Judged on whole-step GB/s against the HBM roofline.
#include <math_constants.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "stream_common.cuh"

namespace lp {

constexpr int DS_MAXP = 4;           // max sequence splits per head
constexpr int DS_CTHREADS = GS_CWARPS * 32;
constexpr int DS_EWARPS = 2;                                   // epilogue warps (tile parity 0 / 1)
constexpr int DS_THREADS = (GS_CWARPS + 1 + DS_EWARPS + 1) * 32;   // consumers + producer warp + epilogue warps + watcher warp
constexpr int DS_TILE_BAR_THREADS = DS_CTHREADS + 32;          // named barriers 2 / 3: the consumers arrive, one epilogue warp waits
constexpr int DS_OPEND_THREADS = DS_CTHREADS + DS_EWARPS * 32; // named barrier 4: end of an op
constexpr int DS_KIND_LINEAR = 0, DS_KIND_EXCHANGE = 2, DS_KIND_SLAB = 3;
constexpr int DS_SLAB_ROWS = 256;  // output rows per slab stage: one m16 tile per consumer warp
constexpr int DS_SLAB_MAXU = 16;   // 8-column units of a CTA's slab (<= 128 input columns)
constexpr int DS_MAX_TP = 8;
constexpr int DS_NREC = 3;  // op records resident in shared memory: previous (its signal may still be pending), current, next
constexpr int DS_RED_FLOATS = 2 * GS_CWARPS * 16 * 4;  // [2 parities][warps][16 rows][4 B-columns]

struct alignas(128) DsOp {
  CUtensorMap map;  // linear: weight matrix {128 B, N rows, K-blocks}
  const float* x;
  const float* residual;
  float* out;
  const float* bias;
  const float* nw;
  const float* nb;
  const void* aux2;
  const float* qkv;       // attention
  __nv_bfloat16* kc;
  __nv_bfloat16* vc;
  float eps;
  int kind, dep, signal;
  int norm_kind, epi, fmt, N, K, split, ldx, nkb, nks, ntiles, ngroups, gp128, aux_bytes, x_attn;
  int save_x, reuse_x;  // LayerNorm ops reading the same row back to back (parallel residual: QKV then FC): the first keeps the raw
                        // row and its statistics in shared memory, the second stages from there (no L2 round trip, no reductions)
  int streamk;  // in-place residual op: stages (not tiles) are split evenly over the CTAs, partial tiles are added atomically
  // fused column->row pairs (DS_KIND_SLAB): this CTA multiplies the input columns it produced itself (its SwiGLU outputs, or the
  // attention output of its head) with the matching K-slab of the following projection and adds the partial row into `out`
  const unsigned char* slab_img;  // per-CTA slab images (lp_decode_step_slab_build)
  int slab_src;    // 0: the CTA's outputs of the preceding LINEAR op (keep_local); 1: the attention output of the CTA's head
  int keep_local;  // LINEAR: the epilogue leaves this CTA's outputs in shared memory (s_loc) instead of writing them to `out`
  int hsync;       // ATTENTION feeding a slab / SLAB fed by attention: first of the H per-head arrival counters, else -1
  // tensor-parallel exchange (kind 2): out = residual + sum over ranks of the partial at buf_off of every rank's symmetric buffer
  const unsigned long long* tp_bufs;  // [tp] peer-mapped buffer addresses
  const unsigned long long* tp_pads;  // [tp] peer-mapped signal pads
  unsigned int* tp_state;             // [2] of the slot: epoch (shared with lp_tp_allreduce_residual), unused
  unsigned long long tp_buf_off;
  int tp_pad_base, tp_rank, tp_size, tp_use, tp_uses;  // tp_use: index of this exchange among the tp_uses of its slot per step
};

struct DsSlabMeta {  // == lp_slab_meta: what one CTA streams for a slab op
  long long off;    // byte offset of the CTA's image
  int nunits;       // 8-column units of the CTA (<= DS_SLAB_MAXU)
  int units_a;      // units that belong to the first scale group (the rest to the next one)
  int nseg;         // scale groups touched: 1 or 2
  int row0, nrb;    // first output row, number of 256-row stages (0: no work)
  int unit0;        // first unit (global index: column / 8)
  int group_a;      // scale group of the first segment
  int stage_bytes;  // (nunits + nseg) KB
};
static_assert(sizeof(DsSlabMeta) == 40 && sizeof(DsSlabMeta) == sizeof(lp_slab_meta), "lp_slab_meta layout");

struct DsParams {
  const DsOp* ops;
  const DsSlabMeta* slab_meta[2];  // [0]: MLP slabs, [1]: attention-projection slabs; one record per CTA, or NULL
  int ncounters;                   // arrival counters: one per op + H per fused attention op
  int dbg;                         // LP_DS_DEBUG bits (timing experiments, WRONG results): 1 no reds, 2 no stagger, 4 stores instead of reds
  unsigned* counters;  // [nops], zeroed by the prologue kernel of every step
  const int* pos;
  const float* cosT;
  const float* sinT;
  float* part;         // attention partials [H][P][hs + 4]
  unsigned long long* trace;
  float scale_log2;
  int nops, H, G, n_elem, max_seq, P;
  int nstages, stage_stride, xsum_floats;
  int xs_bytes;  // size of the activation-column area; the raw-row buffer of save_x / reuse_x ops follows it
  int i4pair;  // int4 ops use the paired main loop (even stage count); 2: arithmetic skipped (timing experiment)
  int l2_ahead;  // weight stages the producer may prefetch into L2 beyond its TMA cursor while it is blocked on a full ring
  const int* deps;      // [nops] dep of every op, contiguous (the watcher thread walks it)
  unsigned* err;        // [8] sticky error record of the watchdog: {code, op, dep, CTA, counter value, ...}; 0 = healthy
  unsigned long long timeout_ns;  // bound of every cross-CTA / cross-GPU wait (0: unbounded)
  unsigned int* tp_state0;  // epoch counters of the (up to two) tensor-parallel exchange slots, or NULL
  unsigned int* tp_state1;
};

__device__ __forceinline__ unsigned ds_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ds_ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ds_red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
__device__ __forceinline__ float4 ds_ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ int ds_lds_acquire(const int* p) {  // CTA-scope acquire of a shared-memory flag
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void ds_sts_release(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// Watchdog: every wait on ANOTHER CTA (dependency counters) or another GPU (exchange flags) is bounded.  The first waiter whose
// bound expires writes a record into the sticky error word of the plan and stops waiting; everybody else sees the word within a
// few polls and stops waiting too, so the kernel runs to its end (results of that step are garbage) and exits: a dead dependency
// becomes LP_ERR_TIMEOUT from lp_decode_step_status() instead of a hung GPU.  Waits inside a CTA (mbarriers, named barriers) only
// depend on that CTA's own warps and need no bound.
constexpr unsigned DS_ERR_DEP_TIMEOUT = 1u, DS_ERR_EXCHANGE_TIMEOUT = 2u;
__device__ __forceinline__ void ds_report(unsigned* err, unsigned code, int op, int dep, unsigned seen) {
  if (atomicCAS(err, 0u, code) == 0u) {
    err[1] = (unsigned)op;
    err[2] = (unsigned)dep;
    err[3] = blockIdx.x;
    err[4] = seen;
    __threadfence();
  }
}

// `bytes` (multiple of 16) of constants into L2: issued by the producer thread well ahead of the consumers' need
__device__ __forceinline__ void ds_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

// one weight stage (same tensor map and box as the TMA load) into L2 only
__device__ __forceinline__ void ds_prefetch_stage_l2(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct DsRing {
  uint32_t bar0, ring_u32;
  unsigned char* ring;
  int nstages, stage_stride;
  int s, ph;
  __device__ __forceinline__ uint32_t full() const { return bar0 + 8 * s; }
  __device__ __forceinline__ uint32_t empty() const { return bar0 + 8 * (nstages + s); }
  __device__ __forceinline__ void advance() {
    if (++s == nstages) { s = 0; ph ^= 1; }
  }
};

// Stages [sb, se) of a linear op that belong to this CTA, in (tile, K-stage) order.  Whole 16-row tiles normally; for in-place
// residual ops (N = n_embd: 256 tiles for 148 CTAs would mean 1 or 2 tiles each) the STAGES are split evenly — a tile that
// straddles two CTAs is finished by both with `x += partial` as an atomic add, nobody waits for anybody.
__device__ __forceinline__ void ds_stage_range(const DsOp& o, int& sb, int& se) {
  const long long G = gridDim.x, c = blockIdx.x;
  if (o.streamk) {
    const long long T = (long long)o.ntiles * o.nks;
    sb = (int)(T * c / G);
    se = (int)(T * (c + 1) / G);
  } else {
    sb = (int)((long long)o.ntiles * c / G) * o.nks;
    se = (int)((long long)o.ntiles * (c + 1) / G) * o.nks;
  }
}

// Attention geometry of this step (depends on the device-side position).
template <int HS>
struct DsAttnGeo {
  static constexpr int TPK = HS / 16;             // threads per key in q.K^T (16 dims each)
  static constexpr int AT = DS_CTHREADS / TPK;    // keys per 16 KB tile
  static constexpr int DCH = HS / 8;              // 16-byte chunks per row
  static constexpr int NSL = DS_CTHREADS / DCH;   // key slices in P.V (2 keys of a tile each)
  static constexpr int LDM = HS + 4;              // partial record: o[HS], m, l, pad
  int pos, kv_len, slot, nblk, bpp, slot_blk;
  __device__ __forceinline__ void init(const DsParams& p) {
    pos = p.pos[0];
    kv_len = min(pos + 1, p.max_seq);
    slot = pos % p.max_seq;
    nblk = (kv_len + AT - 1) / AT;
    bpp = (nblk + p.P - 1) / p.P;
    slot_blk = slot / AT;
  }
};

// ------------------------------------------------------------------------------------------------ shared-memory access
// The hot loops address shared memory through 32-bit shared-space addresses (generic pointers cost an address
// conversion per access).
__device__ __forceinline__ uint4 ds_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 ds_lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t ds_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 ds_lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 ds_lds64f(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void ds_sts64f(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void ds_sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};\n" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void ds_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Loop invariants that ptxas would otherwise rematerialise from %tid / kernel parameters in every iteration of the stage
// loop (measured: ~40 of 133 instructions per warp-stage): an opaque move keeps them in their register.
__device__ __forceinline__ uint32_t ds_pin(uint32_t v) {
  asm volatile("mov.b32 %0, %1;\n" : "=r"(v) : "r"(v));
  return v;
}
__device__ __forceinline__ int ds_pin(int v) { return (int)ds_pin((uint32_t)v); }

// ------------------------------------------------------------------------------------------------ activation staging
// One activation row -> shared-memory B-operand columns.  Thread `ctid` owns the 8 consecutive columns 8 ctid + 4096 i of
// every pass, so the row is read from L2 once (activations are produced by other SMs inside this kernel: __ldcg) and kept
// in registers through the norm statistics, max|x| and the conversion; every term / digit row is written with one 8- or
// 16-byte store per 8 columns.  NI = 2 (K <= 8192; norm parameters cached as well), 4 (K <= 16384) or 6 (K <= 24576).
//   bf16 weights: x = hi + mid (+ lo) as bf16 rows [split][ldx];
//   int4 weights: block fixed point X = rint(x * 2^22 / max|x|) as three balanced base-256 int8 digit rows, bytes of a 16-column
//   pair of groups in IMMA operand order (even columns of both groups, then odd columns), plus the per-128-column digit sums
//   (zero-point term).
// `xmode` 1: LayerNorm op that also leaves the raw row in `xraw` and (mean, rstd) in `xstat` for the next op; 2: the op that
// stages from there — same arithmetic on the same values, so the result is bit-identical to staging from global memory.
template <int NI, class LoadX, class WaitDep>
__device__ __forceinline__ void ds_stage_row(const DsOp& o, LoadX load_x, WaitDep wait_dep, float* s_stat, uint32_t xs_u32, float* xsum,
                                             unsigned long long* tr, int xmode = 0, float* xraw = nullptr, float* xstat = nullptr) {
  constexpr int STRIDE = DS_CTHREADS * 8;
  constexpr bool WCACHE = NI <= 2;
  const int K = o.K, fmt = o.fmt, split = o.split, ldx = o.ldx, norm_kind = o.norm_kind;
  const float* nw = o.nw;
  const float* nb = o.nb;
  const int kpad = (K + 127) / 128 * 128;
  const int ctid = threadIdx.x, warp = ctid >> 5, lane = ctid & 31;
  const bool has_norm = norm_kind >= 0, has_bias = has_norm && norm_kind == LP_NORM_LAYERNORM;
  float x[NI][8];
  float wq[WCACHE ? NI : 1][8], bq[WCACHE ? NI : 1][8];
  auto ld8 = [&](const float* base, int k, float (&d)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(base + k), b = *reinterpret_cast<const float4*>(base + k + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  };
  // norm parameters are constants: every warp fetches them BEFORE waiting for the producer op (the dependency is polled by the
  // watcher thread of the producer warp, so nobody here has a reason to hold back), off the critical path behind the dependency
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int k = ctid * 8 + i * STRIDE;
    if (WCACHE && has_norm && k < K) {
      ld8(nw, k, wq[WCACHE ? i : 0]);
      if (has_bias) ld8(nb, k, bq[WCACHE ? i : 0]);
    }
  }
  wait_dep();
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int k = ctid * 8 + i * STRIDE;
    if (k < K) {
      float4 a, b;
      if (xmode == 2) {
        a = *reinterpret_cast<const float4*>(xraw + k);
        b = *reinterpret_cast<const float4*>(xraw + k + 4);
      } else {
        a = load_x(k);
        b = load_x(k + 4);
        if (xmode == 1) {
          *reinterpret_cast<float4*>(xraw + k) = a;
          *reinterpret_cast<float4*>(xraw + k + 4) = b;
        }
      }
      x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w; x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) x[i][q] = 0.f;
    }
  }
  if (tr && ctid == 0) {
    float acc0 = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) acc0 += x[i][0] + (WCACHE && has_norm ? wq[WCACHE ? i : 0][0] : 0.f);
    if (acc0 != 12345.678f) tr[4] = gs_now();  // the loads have returned
  }
  int sbuf = 0;
  // block-wide (sum of a, sum of b); one barrier per reduction: the statistics buffer is double buffered
  auto block_reduce = [&](float a, float b, float& ra, float& rb) {
    a = warp_sum(a);
    b = warp_sum(b);
    float* st = s_stat + sbuf * 2 * GS_CWARPS;
    sbuf ^= 1;
    if (lane == 0) {
      st[warp] = a;
      st[GS_CWARPS + warp] = b;
    }
    gs_bar_consumers();
    const float4* s4 = reinterpret_cast<const float4*>(st);
    ra = 0.f;
    rb = 0.f;
#pragma unroll
    for (int w = 0; w < GS_CWARPS / 4; ++w) {
      const float4 va = s4[w], vb = s4[GS_CWARPS / 4 + w];
      ra += (va.x + va.y) + (va.z + va.w);
      rb += (vb.x + vb.y) + (vb.z + vb.w);
    }
  };
  if (has_norm && !has_bias) {
    // RMSNorm: W . (w x rstd) = rstd * (W . (w x)): the scalar rstd is applied by the tile epilogue, so the row is converted
    // without waiting for a reduction; sum x^2 travels through s_stat[2 * GS_CWARPS ..] and the barrier that ends the staging
    // (ds_linear turns it into the post scale).  No block reduction on this path.
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int k = ctid * 8 + i * STRIDE;
      float wl[8];
      if (!WCACHE && k < K) ld8(nw, k, wl);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ss = fmaf(x[i][q], x[i][q], ss);
        x[i][q] *= (WCACHE ? wq[WCACHE ? i : 0][q] : (k < K ? wl[q] : 0.f));
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) s_stat[2 * GS_CWARPS + warp] = ss;
  } else if (has_norm) {
    float sm = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        sm += x[i][q];
        ss = fmaf(x[i][q], x[i][q], ss);
      }
    float mean = 0.f, rstd;
    if (xmode == 2) {
      mean = xstat[0];
      rstd = xstat[1];
    } else {
      block_reduce(sm, ss, sm, ss);
      mean = sm / (float)K;
      float v2 = 0.f, dummy;  // two-pass variance
#pragma unroll
      for (int i = 0; i < NI; ++i)
        if (ctid * 8 + i * STRIDE < K) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float d = x[i][q] - mean;
            v2 = fmaf(d, d, v2);
          }
        }
      block_reduce(v2, 0.f, v2, dummy);
      rstd = 1.0f / sqrtf(v2 / (float)K + o.eps);
      if (xmode == 1 && ctid == 0) {
        xstat[0] = mean;
        xstat[1] = rstd;
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int k = ctid * 8 + i * STRIDE;
      if (k < K) {
        float wl[8], bl[8];
        if (!WCACHE) {
          ld8(nw, k, wl);
          ld8(nb, k, bl);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          x[i][q] = (x[i][q] - mean) * rstd * (WCACHE ? wq[WCACHE ? i : 0][q] : wl[q]) + (WCACHE ? bq[WCACHE ? i : 0][q] : bl[q]);
      }
    }
  }
  if (tr && ctid == 0) tr[5] = gs_now();  // normalised
  if (fmt == LP_W_INT4 || fmt == LP_W_INT8) {
    const bool natural = fmt == LP_W_INT8;  // int8 weights: digit bytes in column order; int4: even / odd columns split (nibble IMMAs)
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int k = ctid * 8 + i * STRIDE;
      if (k < kpad) {  // half-warp uniform (a 128-column chunk = 16 threads): the zero padding of the last chunk is written too
        // block fixed point PER 128-COLUMN GROUP (the 16 lanes that own it): X = rint(x * 2^22 / max|x_group|); the group's
        // scale a_g = max / 2^22 rides in the 4th float of the group's digit-sum record and is folded into the weight scale by
        // the main loop.  No block-wide reduction, and small groups keep all 22 bits.
        float am = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) am = fmaxf(am, fabsf(x[i][q]));
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
        const float inv = am > 0.f ? 4194304.0f / am : 0.f;
        // X = d0 + 256 d1 + 65536 d2 with balanced digits d0, d1 in [-128, 127]: the bytes of Z = (X + 0x8080) ^ 0x8080 ARE those digits
        // as int8 (X + 0x8080 = (d0 + 128) + 256 (d1 + 128) + 65536 d2 without carries; ^ 0x80 takes the 128 off again)
        uint32_t z[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) z[q] = (uint32_t)(__float2int_rn(x[i][q] * inv) + 0x8080) ^ 0x8080u;
        // 4 x 4 byte transposes: digit d of columns 0,2,4,6 (IMMA a0/a1 side) and of columns 1,3,5,7 (a2/a3 side)
        uint32_t lo[3], hi[3];
        if (natural) {  // words of columns 0..3 and 4..7
          const uint32_t u0 = __byte_perm(z[0], z[1], 0x5140), u1 = __byte_perm(z[2], z[3], 0x5140);
          const uint32_t u2 = __byte_perm(z[0], z[1], 0x7362), u3 = __byte_perm(z[2], z[3], 0x7362);
          lo[0] = __byte_perm(u0, u1, 0x5410);
          lo[1] = __byte_perm(u0, u1, 0x7632);
          lo[2] = __byte_perm(u2, u3, 0x5410);
          const uint32_t v0 = __byte_perm(z[4], z[5], 0x5140), v1 = __byte_perm(z[6], z[7], 0x5140);
          const uint32_t v2 = __byte_perm(z[4], z[5], 0x7362), v3 = __byte_perm(z[6], z[7], 0x7362);
          hi[0] = __byte_perm(v0, v1, 0x5410);
          hi[1] = __byte_perm(v0, v1, 0x7632);
          hi[2] = __byte_perm(v2, v3, 0x5410);
        } else {
          const uint32_t u0 = __byte_perm(z[0], z[2], 0x5140), u1 = __byte_perm(z[4], z[6], 0x5140);
          const uint32_t u2 = __byte_perm(z[0], z[2], 0x7362), u3 = __byte_perm(z[4], z[6], 0x7362);
          lo[0] = __byte_perm(u0, u1, 0x5410);
          lo[1] = __byte_perm(u0, u1, 0x7632);
          lo[2] = __byte_perm(u2, u3, 0x5410);
          const uint32_t v0 = __byte_perm(z[1], z[3], 0x5140), v1 = __byte_perm(z[5], z[7], 0x5140);
          const uint32_t v2 = __byte_perm(z[1], z[3], 0x7362), v3 = __byte_perm(z[5], z[7], 0x7362);
          hi[0] = __byte_perm(v0, v1, 0x5410);
          hi[1] = __byte_perm(v0, v1, 0x7632);
          hi[2] = __byte_perm(v2, v3, 0x5410);
        }
        float ps[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          // 16-column pair of groups (A, B) -> [A.even | B.even | A.odd | B.odd]: the B operands of the low-nibble and the
          // high-nibble IMMA are then adjacent registers of one 128-bit load
          const uint32_t a16 = natural ? xs_u32 + (uint32_t)(d * ldx + k) : xs_u32 + (uint32_t)(d * ldx + (k & ~15) + ((k & 8) >> 1));
          asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a16), "r"(lo[d]) : "memory");
          asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a16 + (natural ? 4 : 8)), "r"(hi[d]) : "memory");
          int sd = __dp4a((int)lo[d], 0x01010101, 0);  // digit sum of the thread's 8 columns
          sd = __dp4a((int)hi[d], 0x01010101, sd);
#pragma unroll
          for (int off = 8; off > 0; off >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, off);
          ps[d] = (float)sd;
        }
        if ((lane & 15) == 0) *reinterpret_cast<float4*>(xsum + (k >> 7) * 4) = make_float4(ps[0], ps[1], ps[2], am * (1.0f / 4194304.0f));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int k = ctid * 8 + i * STRIDE;
      if (k < K) {
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
          if (sp < split) {
            uint32_t w2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t h0 = gs_bf16_bits(x[i][2 * q]), h1 = gs_bf16_bits(x[i][2 * q + 1]);
              x[i][2 * q] -= __uint_as_float(h0 << 16);  // exact: next term of the split
              x[i][2 * q + 1] -= __uint_as_float(h1 << 16);
              w2[q] = h0 | (h1 << 16);
            }
            ds_sts128(xs_u32 + (uint32_t)(sp * ldx + k) * 2, w2[0], w2[1], w2[2], w2[3]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ one linear op (consumers)
// Main loop of one linear op, specialised per weight format.  Per 16 KB stage a warp handles half a K-block of the 16-row
// tile: bf16: 2 x (ldmatrix.x4 + mma.sync.m16n8k16) against the bf16 term columns; int4: 4 x IMMA m16n8k32 u8 x s8 against the
// int8 digit columns, then the per-group scale / zero point.  Everything that does not change from stage to stage (lane
// offsets inside a stage, the position of the warp's K-slice) lives in registers; per stage: wait, loads, math, arrive.
template <int FMT, bool PACKED>
__device__ __forceinline__ void ds_linear_main(const DsOp& o, DsRing& rg, uint32_t red_u32, uint32_t xsum_u32,
                                               uint32_t xs_u32, volatile int* done, int sb, int se, int& gt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nks = o.nks, split = o.split, ldx = o.ldx, gp128 = o.gp128;
  const int kbl = warp >> 1, sub0 = warp & 1;
  const int bcol = g < split ? g : split - 1;  // B columns >= split only feed accumulator columns nobody reads
  const int nunits = se - sb;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};

  // lane-constant byte offsets inside a stage, and the K position of this warp's slice in the first stage of a tile
  uint32_t woff0, woff1, xpos0, xstep;
  int slice0, nslices;  // this warp's slice index in stage 0 and the number of slices per row (bounds check of the last stage)
  const int pr0 = (g >> 1) + 4 * (g & 1);  // int4: MMA row g <-> tile row pr0 (bank-conflict-free under the 128-byte swizzle)
  if (FMT == LP_W_BF16) {
    const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
    woff0 = kbl * GS_BLK_BYTES + row * 128 + ((((sub0 * 2 + 0) * 2 + (lane >> 4)) ^ (row & 7)) << 4);
    woff1 = kbl * GS_BLK_BYTES + row * 128 + ((((sub0 * 2 + 1) * 2 + (lane >> 4)) ^ (row & 7)) << 4);
    xpos0 = xs_u32 + (uint32_t)(bcol * ldx + kbl * 64 + sub0 * 32 + 2 * t) * 2;
    xstep = GS_KB * 64 * 2;
    slice0 = kbl;
    nslices = o.nkb;
  } else {
    woff0 = kbl * GS_BLK_BYTES + pr0 * 128 + (((sub0 * 4 + t) ^ (pr0 & 7)) << 4);
    woff1 = woff0 + 8 * 128;
    xpos0 = xs_u32 + (uint32_t)(bcol * ldx + (kbl * 2 + sub0) * 128 + t * 32);
    xstep = GS_KB * 2 * 128;
    slice0 = kbl * 2 + sub0;
    nslices = (o.K + 127) / 128;
  }
  const int slice_step = FMT == LP_W_BF16 ? GS_KB : GS_KB * 2;
  constexpr int AUXB = PACKED ? 4 : 8;
  const uint32_t xsum0 = ds_pin(xsum_u32 + (uint32_t)((kbl * 2 + sub0) * 4) * 4);  // group record {digit sums 0..2, a_g}
  const bool todd = ds_pin(t & 1) != 0;  // accumulator columns 2t, 2t+1: digits (0, 1) for even t, (2, -) for odd t
  const uint32_t aux0 = ds_pin((uint32_t)(GS_KB * GS_BLK_BYTES + (slice0 * 16 + pr0) * AUXB));  // group 128: scale slot of this slice
  const uint32_t auxr = ds_pin((uint32_t)(GS_KB * GS_BLK_BYTES + pr0 * AUXB));
  woff0 = ds_pin(woff0);
  woff1 = ds_pin(woff1);
  xpos0 = ds_pin(xpos0);
  slice0 = ds_pin(slice0);
  const bool lane0 = ds_pin(lane) == 0;
  const uint32_t ring_u32 = ds_pin(rg.ring_u32), bar0 = ds_pin(rg.bar0);
  const int nstages = ds_pin(rg.nstages), stage_stride = ds_pin(rg.stage_stride);
  int rs = rg.s, rph = rg.ph;
  int ks = sb % nks;  // the first tile may start in the middle (stream-K ops)
  uint32_t xpos = xpos0 + (uint32_t)ks * xstep, xsp = xsum0 + (uint32_t)ks * (GS_KB * 2 * 4 * 4);
  int slice = slice0 + ks * slice_step;

  for (int u = 0; u < nunits; ++u) {
    mbar_wait(bar0 + 8 * rs, rph);
    const uint32_t st = ring_u32 + (uint32_t)(rs * stage_stride);
    if (slice < nslices) {
      if (FMT == LP_W_BF16) {
        uint32_t a[4], c[4];
        gs_ldsm_x4(a, st + woff0);
        gs_ldsm_x4(c, st + woff1);
        const uint32_t b0 = ds_lds32(xpos), b1 = ds_lds32(xpos + 16), b2 = ds_lds32(xpos + 32), b3 = ds_lds32(xpos + 48);
        gs_mma(acc, a[0], a[1], a[2], a[3], b0, b1);
        gs_mma(acc, c[0], c[1], c[2], c[3], b2, b3);
      } else {
        const uint4 wa = ds_lds128(st + woff0), wb = ds_lds128(st + woff1);
        const uint4 xv0 = ds_lds128(xpos), xv1 = ds_lds128(xpos + 16);
        const float4 xg = ds_lds128f(xsp);
        float2 xsv;
        xsv.x = todd ? xg.z : xg.x;
        xsv.y = todd ? xg.w : xg.y;
        // scale / zero of this row pair for the slice's group (aux block behind the weights of the stage)
        const uint32_t ap = st + (gp128 == 1 ? aux0 : auxr + (uint32_t)(slice / gp128 - (ks * GS_KB * 2) / gp128) * 16 * AUXB);
        float s0, s1, z0, z1;
        if (PACKED) {  // bf16 scale << 16 | bf16 zero
          const uint32_t u0 = ds_lds32(ap), u1 = ds_lds32(ap + 8 * AUXB);
          s0 = __uint_as_float(u0 & 0xffff0000u);
          s1 = __uint_as_float(u1 & 0xffff0000u);
          z0 = __uint_as_float(u0 << 16);
          z1 = __uint_as_float(u1 << 16);
        } else {
          const float2 a0 = ds_lds64f(ap), a1 = ds_lds64f(ap + 8 * AUXB);
          s0 = a0.x; z0 = a0.y; s1 = a1.x; z1 = a1.y;
        }
        s0 *= xg.w;  // activation scale of this 128-column group (block fixed point per group, ds_stage_row)
        s1 *= xg.w;
        // low nibbles (even columns) and high nibbles (odd columns, left in place: 16 q) go through separate IMMAs, so the
        // unpack is one LOP3 per operand register; 16 q d sums are exact multiples of 16 and rescaled in fp32.
        int cl[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
        const uint32_t ML = 0x0F0F0F0Fu, MH = 0xF0F0F0F0u;
        gs_imma(cl, wa.x & ML, wb.x & ML, wa.y & ML, wb.y & ML, xv0.x, xv0.y);
        gs_imma(ch, wa.x & MH, wb.x & MH, wa.y & MH, wb.y & MH, xv0.z, xv0.w);
        gs_imma(cl, wa.z & ML, wb.z & ML, wa.w & ML, wb.w & ML, xv1.x, xv1.y);
        gs_imma(ch, wa.z & MH, wb.z & MH, wa.w & MH, wb.w & MH, xv1.z, xv1.w);
        // sum (q - z) s x = s * (sum q X - z * sum X), all integers exact in fp32 (|.| < 2^24)
        acc[0] = fmaf(s0, fmaf((float)ch[0], 0.0625f, fmaf(-z0, xsv.x, (float)cl[0])), acc[0]);
        acc[1] = fmaf(s0, fmaf((float)ch[1], 0.0625f, fmaf(-z0, xsv.y, (float)cl[1])), acc[1]);
        acc[2] = fmaf(s1, fmaf((float)ch[2], 0.0625f, fmaf(-z1, xsv.x, (float)cl[2])), acc[2]);
        acc[3] = fmaf(s1, fmaf((float)ch[3], 0.0625f, fmaf(-z1, xsv.y, (float)cl[3])), acc[3]);
      }
    }
    __syncwarp();
    if (lane0) mbar_arrive(bar0 + 8 * (nstages + rs));
    if (++rs == nstages) { rs = 0; rph ^= 1; }
    xpos += xstep;
    xsp += GS_KB * 2 * 4 * 4;
    slice += slice_step;

    if (++ks == nks || u == nunits - 1) {  // end of the tile, or of this CTA's part of it
      ks = 0;
      xpos = xpos0;
      xsp = xsum0;
      slice = slice0;
      // tile finished: every warp drops its partial sums (columns 0..3) and moves on; the epilogue warp of this parity
      // finalises.  `gt` numbers the tiles of this CTA across ALL ops: parity = gt & 1, and a `red` buffer may be rewritten once
      // the previous tile of the same parity has been finalised (done[par] counts finalised tiles).
      const int par = gt & 1;
      while (done[par] < (gt >> 1)) {}
      if (t < 2) {
        const int prow = FMT == LP_W_INT4 ? pr0 : g;  // tile row held by accumulator row g
        const uint32_t r = red_u32 + (uint32_t)(((par * GS_CWARPS + warp) * 16 + prow) * 4 + 2 * t) * 4;
        ds_sts64f(r, acc[0], acc[1]);
        ds_sts64f(r + 8 * 4 * 4, acc[2], acc[3]);
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      // barrier ids are immediates on purpose (a register id makes ptxas reserve all 16 named barriers)
      if (par == 0) asm volatile("bar.arrive 2, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
      else asm volatile("bar.arrive 3, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
      ++gt;
    }
  }
  rg.s = rs;
  rg.ph = rph;
}

// int4 main loop, "paired" form (ring with an even number of stages): the stage loop of ds_linear_main costs ~60 SASS
// instructions per warp and stage around ~55 of arithmetic (mbarrier wait / arrive, ring and K bookkeeping, bounds and tile
// checks), and at int4 width the issue slots are what paces the stream.  Here a warp visits only every other stage — warps
// with (warp & 1) == h own the ring slots of parity h — and works on a whole K-block (both 128-column groups, 2 KB) there, so
// the fixed part is paid once per 2 KB.  A slot is released by its 8 warps with an arrival count of 2 each (the barrier keeps
// its count of GS_CWARPS: attention and the bf16 loops share the ring).  Slot parity is fixed per warp because the stage
// count is even, so every warp sees every phase of the barriers it waits on.
__device__ __forceinline__ void mbar_arrive2(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 2;\n" ::"r"(bar) : "memory");
}

template <bool PACKED>
__device__ __forceinline__ void ds_linear_main_i4pair(const DsOp& o, DsRing& rg, uint32_t red_u32, uint32_t xsum_u32, uint32_t xs_u32,
                                                      volatile int* done, int sb, int se, int& gt, bool nomath) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nks = o.nks, split = o.split, ldx = o.ldx, gp128 = o.gp128;
  const int kbl = warp >> 1, h = warp & 1;
  const int bcol = g < split ? g : split - 1;
  const int nunits = se - sb;
  const int nslices = (o.K + 127) / 128;
  const int pr0 = (g >> 1) + 4 * (g & 1);
  constexpr int AUXB = PACKED ? 4 : 8;
  constexpr uint32_t XSTEP = GS_KB * 2 * 128, XSUMSTEP = GS_KB * 2 * 4 * 4;
  // lane-constant offsets of the FIRST group of the warp's K-block; the second group is + 128 B (x digits), + 16 B (digit
  // sums), ^ 64 B (weights: chunk (4 + t) ^ (row & 7) under the 128-byte swizzle), + 16 AUXB (scales)
  const uint32_t woff = ds_pin((uint32_t)(kbl * GS_BLK_BYTES + pr0 * 128 + ((t ^ (pr0 & 7)) << 4)));
  const uint32_t xpos0 = ds_pin(xs_u32 + (uint32_t)(bcol * ldx + kbl * 256 + t * 32));
  const uint32_t xsum0 = ds_pin(xsum_u32 + (uint32_t)(kbl * 8) * 4);
  const bool todd = ds_pin(t & 1) != 0;
  const uint32_t aux0 = ds_pin((uint32_t)(GS_KB * GS_BLK_BYTES + (kbl * 32 + pr0) * AUXB));
  const uint32_t auxr = ds_pin((uint32_t)(GS_KB * GS_BLK_BYTES + pr0 * AUXB));
  const int slice0 = ds_pin(kbl * 2);
  const bool lane0 = ds_pin(lane) == 0;
  const uint32_t ring_u32 = ds_pin(rg.ring_u32), bar0 = ds_pin(rg.bar0);
  const int nstages = ds_pin(rg.nstages), stage_stride = ds_pin(rg.stage_stride);
  // ring position of this warp's first unit (local unit h)
  int rs = rg.s + h, rph = rg.ph;
  if (rs >= nstages) { rs -= nstages; rph ^= 1; }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};

  auto group = [&](uint32_t st, const uint4& wa, const uint4& wb, const uint4& xv0, const uint4& xv1, const float4& xg, int slice,
                   int ks, uint32_t aux_first) {
    float2 xsv;
    xsv.x = todd ? xg.z : xg.x;
    xsv.y = todd ? xg.w : xg.y;
    const uint32_t ap = st + (gp128 == 1 ? aux_first : auxr + (uint32_t)(slice / gp128 - (ks * GS_KB * 2) / gp128) * 16 * AUXB);
    float s0, s1, z0, z1;
    if (PACKED) {
      const uint32_t u0 = ds_lds32(ap), u1 = ds_lds32(ap + 8 * AUXB);
      s0 = __uint_as_float(u0 & 0xffff0000u);
      s1 = __uint_as_float(u1 & 0xffff0000u);
      z0 = __uint_as_float(u0 << 16);
      z1 = __uint_as_float(u1 << 16);
    } else {
      const float2 a0 = ds_lds64f(ap), a1 = ds_lds64f(ap + 8 * AUXB);
      s0 = a0.x; z0 = a0.y; s1 = a1.x; z1 = a1.y;
    }
    s0 *= xg.w;
    s1 *= xg.w;
    int cl[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
    const uint32_t ML = 0x0F0F0F0Fu, MH = 0xF0F0F0F0u;
    gs_imma(cl, wa.x & ML, wb.x & ML, wa.y & ML, wb.y & ML, xv0.x, xv0.y);
    gs_imma(ch, wa.x & MH, wb.x & MH, wa.y & MH, wb.y & MH, xv0.z, xv0.w);
    gs_imma(cl, wa.z & ML, wb.z & ML, wa.w & ML, wb.w & ML, xv1.x, xv1.y);
    gs_imma(ch, wa.z & MH, wb.z & MH, wa.w & MH, wb.w & MH, xv1.z, xv1.w);
    acc[0] = fmaf(s0, fmaf((float)ch[0], 0.0625f, fmaf(-z0, xsv.x, (float)cl[0])), acc[0]);
    acc[1] = fmaf(s0, fmaf((float)ch[1], 0.0625f, fmaf(-z0, xsv.y, (float)cl[1])), acc[1]);
    acc[2] = fmaf(s1, fmaf((float)ch[2], 0.0625f, fmaf(-z1, xsv.x, (float)cl[2])), acc[2]);
    acc[3] = fmaf(s1, fmaf((float)ch[3], 0.0625f, fmaf(-z1, xsv.y, (float)cl[3])), acc[3]);
  };

  int ub = 0, ks0 = sb % nks;  // first local unit of the current tile, its K-stage
  while (ub < nunits) {
    const int ue = min(nunits, ub + nks - ks0);
    int m = ub + ((ub ^ h) & 1);
    int ks = ks0 + (m - ub);
    uint32_t xpos = xpos0 + (uint32_t)ks * XSTEP, xsp = xsum0 + (uint32_t)ks * XSUMSTEP;
    int slice = slice0 + ks * (GS_KB * 2);
    for (; m < ue; m += 2) {
      mbar_wait(bar0 + 8 * rs, rph);
      const uint32_t st = ring_u32 + (uint32_t)(rs * stage_stride);
      if (nomath) {
      } else if (slice + 1 < nslices) {
        const uint4 wa0 = ds_lds128(st + woff), wb0 = ds_lds128(st + woff + 8 * 128);
        const uint4 wa1 = ds_lds128(st + (woff ^ 64u)), wb1 = ds_lds128(st + (woff ^ 64u) + 8 * 128);
        const uint4 xa0 = ds_lds128(xpos), xa1 = ds_lds128(xpos + 16);
        const uint4 xb0 = ds_lds128(xpos + 128), xb1 = ds_lds128(xpos + 144);
        const float4 xs0 = ds_lds128f(xsp), xs1 = ds_lds128f(xsp + 16);
        group(st, wa0, wb0, xa0, xa1, xs0, slice, ks, aux0);
        group(st, wa1, wb1, xb0, xb1, xs1, slice + 1, ks, aux0 + 16 * AUXB);
      } else if (slice < nslices) {
        const uint4 wa0 = ds_lds128(st + woff), wb0 = ds_lds128(st + woff + 8 * 128);
        const uint4 xa0 = ds_lds128(xpos), xa1 = ds_lds128(xpos + 16);
        const float4 xs0 = ds_lds128f(xsp);
        group(st, wa0, wb0, xa0, xa1, xs0, slice, ks, aux0);
      }
      __syncwarp();
      if (lane0) mbar_arrive2(bar0 + 8 * (nstages + rs));
      rs += 2;
      if (rs >= nstages) { rs -= nstages; rph ^= 1; }
      ks += 2;
      xpos += 2 * XSTEP;
      xsp += 2 * XSUMSTEP;
      slice += 2 * GS_KB * 2;
    }
    // end of the tile (or of this CTA's part of it): as in ds_linear_main
    const int par = gt & 1;
    while (done[par] < (gt >> 1)) {}
    if (t < 2) {
      const uint32_t r = red_u32 + (uint32_t)(((par * GS_CWARPS + warp) * 16 + pr0) * 4 + 2 * t) * 4;
      ds_sts64f(r, acc[0], acc[1]);
      ds_sts64f(r + 8 * 4 * 4, acc[2], acc[3]);
    }
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    if (par == 0) asm volatile("bar.arrive 2, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
    else asm volatile("bar.arrive 3, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
    ++gt;
    ub = ue;
    ks0 = 0;
  }
  // common ring position after the op
  const int adv = rg.s + nunits;
  rg.ph ^= (adv / nstages) & 1;
  rg.s = adv % nstages;
}

// Main loops of the bitsandbytes formats (quantize/bnb.py:18-75; arithmetic restated from the published formats, parity unpinned):
//   INT8  row-wise int8 weights: IMMA m16n8k32 s8 x s8 against the three int8 digit planes of the activations (no unpack at
//         all: the weight bytes ARE the A operand); per 128-column K-block the integer sums are scaled by the block's activation
//         scale, the row scale (SCB / 127) is applied by the tile epilogue.
//   NF4   4-bit codes into the 16-entry normal-float table, fp32 absmax per 64 weights: every code goes through a shared-memory
//         table (two copies, lanes alternate: bank-conflict free) that yields the bf16 hi and lo terms of its value
//         (hi + lo = the fp32 table entry to 2^-17); the HMMA loop of the bf16 format then runs on those terms against the bf16
//         terms of the activations, one accumulator per 64-weight block, scaled by the block's absmax.
__device__ __forceinline__ void gs_imma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int FMT>
__device__ __forceinline__ void ds_linear_main_bnb(const DsOp& o, DsRing& rg, uint32_t red_u32, uint32_t xsum_u32, uint32_t xs_u32,
                                                   uint32_t tbl_u32, volatile int* done, int sb, int se, int& gt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nks = o.nks, split = o.split, ldx = o.ldx, nkb = o.nkb;
  const int kbl = warp >> 1, sub0 = warp & 1;
  const int bcol = g < split ? g : split - 1;  // B columns >= split only feed accumulator columns nobody reads
  const int nunits = se - sb;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool lane0 = ds_pin(lane) == 0;
  const uint32_t ring_u32 = ds_pin(rg.ring_u32), bar0 = ds_pin(rg.bar0);
  const int nstages = ds_pin(rg.nstages), stage_stride = ds_pin(rg.stage_stride);
  const uint32_t wrow = ds_pin((uint32_t)(kbl * GS_BLK_BYTES + g * 128));  // tile row g of this warp's K-block (row g + 8: + 1024)
  const uint32_t sw = ds_pin((uint32_t)(g & 7));                           // 128-byte swizzle: 16-byte chunk index ^ (row & 7)
  const uint32_t tb = ds_pin(tbl_u32 + (uint32_t)((lane & 1) * 64));       // NF4 table copy of this lane
  int rs = rg.s, rph = rg.ph;
  int ks = sb % nks;  // the first tile may start in the middle (stream-K ops)
  for (int u = 0; u < nunits; ++u) {
    mbar_wait(bar0 + 8 * rs, rph);
    const uint32_t st = ring_u32 + (uint32_t)(rs * stage_stride);
    const int kb = ks * GS_KB + kbl;  // K-block of the row
    if (kb < nkb) {
      if (FMT == LP_W_INT8) {
        // K-block = 128 weights; this warp: columns [64 sub0, 64 sub0 + 64) = two k32 steps
        int c[4] = {0, 0, 0, 0};
        const uint32_t xb = xs_u32 + (uint32_t)(bcol * ldx + kb * 128 + sub0 * 64 + 4 * t);
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const uint32_t c0 = (uint32_t)(4 * sub0 + 2 * s2);
          const uint32_t a0 = ds_lds32(st + wrow + (((c0) ^ sw) << 4) + 4 * t), a1 = ds_lds32(st + wrow + 1024 + (((c0) ^ sw) << 4) + 4 * t);
          const uint32_t a2 = ds_lds32(st + wrow + (((c0 + 1) ^ sw) << 4) + 4 * t), a3 = ds_lds32(st + wrow + 1024 + (((c0 + 1) ^ sw) << 4) + 4 * t);
          const uint32_t b0 = ds_lds32(xb + s2 * 32), b1 = ds_lds32(xb + s2 * 32 + 16);
          gs_imma_s8(c, a0, a1, a2, a3, b0, b1);
        }
        const float ag = __uint_as_float(ds_lds32(xsum_u32 + (uint32_t)(kb * 4 + 3) * 4));  // activation scale of this 128-column block
        acc[0] = fmaf(ag, (float)c[0], acc[0]);
        acc[1] = fmaf(ag, (float)c[1], acc[1]);
        acc[2] = fmaf(ag, (float)c[2], acc[2]);
        acc[3] = fmaf(ag, (float)c[3], acc[3]);
      } else {
        // K-block = 256 weights = 128 bytes; this warp: bytes [64 sub0, 64 sub0 + 64) = 8 k16 groups = 2 absmax blocks
        const uint32_t xb = xs_u32 + (uint32_t)(bcol * ldx + kb * 256 + sub0 * 128 + 2 * t) * 2;
        const uint32_t am = st + GS_KB * GS_BLK_BYTES + (uint32_t)(((kbl * 4 + sub0 * 2) * 16 + g) * 4);
        const uint32_t sh = (uint32_t)(8 * t);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          float ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int q = blk * 4 + qq;
            const uint32_t off = ((((uint32_t)(4 * sub0 + (q >> 1))) ^ sw) << 4) + (uint32_t)(8 * (q & 1));
            const uint2 w0 = ds_lds64(st + wrow + off), w1 = ds_lds64(st + wrow + 1024 + off);
            // bytes t (k = 2t, 2t + 1) and t + 4 (k = 2t + 8, 2t + 9) of the group, rows g and g + 8; first weight in the HIGH nibble
            uint32_t areg[2][4];
            const uint32_t by[4] = {(w0.x >> sh) & 0xffu, (w1.x >> sh) & 0xffu, (w0.y >> sh) & 0xffu, (w1.y >> sh) & 0xffu};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const uint32_t le = ds_lds32(tb + ((by[r] >> 2) & 0x3cu)), lo = ds_lds32(tb + ((by[r] << 2) & 0x3cu));
              areg[0][r] = __byte_perm(le, lo, 0x5410);  // bf16 hi terms of (k even, k odd)
              areg[1][r] = __byte_perm(le, lo, 0x7632);  // bf16 lo terms
            }
            const uint32_t b0 = ds_lds32(xb + (uint32_t)(q * 32)), b1 = ds_lds32(xb + (uint32_t)(q * 32 + 16));
            gs_mma(ab, areg[0][0], areg[0][1], areg[0][2], areg[0][3], b0, b1);
            gs_mma(ab, areg[1][0], areg[1][1], areg[1][2], areg[1][3], b0, b1);
          }
          const float m0 = __uint_as_float(ds_lds32(am + (uint32_t)(blk * 64))), m1 = __uint_as_float(ds_lds32(am + (uint32_t)(blk * 64 + 32)));
          acc[0] = fmaf(m0, ab[0], acc[0]);
          acc[1] = fmaf(m0, ab[1], acc[1]);
          acc[2] = fmaf(m1, ab[2], acc[2]);
          acc[3] = fmaf(m1, ab[3], acc[3]);
        }
      }
    }
    __syncwarp();
    if (lane0) mbar_arrive(bar0 + 8 * (nstages + rs));
    if (++rs == nstages) { rs = 0; rph ^= 1; }
    if (++ks == nks || u == nunits - 1) {  // end of the tile, or of this CTA's part of it (as in ds_linear_main)
      ks = 0;
      const int par = gt & 1;
      while (done[par] < (gt >> 1)) {}
      if (t < 2) {
        const uint32_t r = red_u32 + (uint32_t)(((par * GS_CWARPS + warp) * 16 + g) * 4 + 2 * t) * 4;
        ds_sts64f(r, acc[0], acc[1]);
        ds_sts64f(r + 8 * 4 * 4, acc[2], acc[3]);
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      if (par == 0) asm volatile("bar.arrive 2, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
      else asm volatile("bar.arrive 3, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
      ++gt;
    }
  }
  rg.s = rs;
  rg.ph = rph;
}

template <int HS, class WaitDep>
__device__ __forceinline__ void ds_linear(const DsParams& p, const DsOp& o, const DsAttnGeo<HS>& geo, DsRing& rg, uint32_t red_u32,
                                          float* colscale, float* xsum, uint32_t xs_u32, volatile int* done, float* s_stat,
                                          unsigned long long* tr, int& gt, WaitDep wait_dep, float* xraw, float* xstat, bool& have_xraw,
                                          uint32_t tbl_u32) {
  int sb, se;
  ds_stage_range(o, sb, se);
  if (se == sb) {  // CTA-uniform: nothing to compute, but later ops rely on the (cumulative) dependency
    wait_dep();
    have_xraw = false;  // this CTA did not stage the row: a following reuse_x op stages from global memory
    return;
  }
  const int xmode = (o.reuse_x && have_xraw) ? 2 : (o.save_x ? 1 : 0);
  have_xraw = xmode == 1;
  // ---- stage x ----
  const bool small = o.K <= 2 * DS_CTHREADS * 8;
  if (o.x_attn) {
    // x = attention output: merge the P sequence-split partials (m, l, o[hs]) of every head on the fly
    const int nvalid = (geo.nblk + geo.bpp - 1) / geo.bpp;  // splits with at least one key block
    const float* part = p.part;
    const int P = p.P;
    auto load_attn = [&](int k) -> float4 {
      constexpr int LDM = DsAttnGeo<HS>::LDM;
      const int h = k / HS, d = k % HS;
      const float* base = part + (size_t)h * P * LDM;
      float m[DS_MAXP], l[DS_MAXP];
      float4 ov[DS_MAXP];
#pragma unroll
      for (int q = 0; q < DS_MAXP; ++q) {
        if (q < nvalid) {
          const float2 ml = __ldcg(reinterpret_cast<const float2*>(base + q * LDM + HS));
          m[q] = ml.x;
          l[q] = ml.y;
          ov[q] = ds_ldcg4(base + q * LDM + d);
        } else {
          m[q] = -CUDART_INF_F;
          l[q] = 0.f;
          ov[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float mx = m[0];
#pragma unroll
      for (int q = 1; q < DS_MAXP; ++q) mx = fmaxf(mx, m[q]);
      float L = 0.f;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < DS_MAXP; ++q) {  // split order: deterministic
        const float c = exp2f(m[q] - mx);
        L = fmaf(l[q], c, L);
        acc.x = fmaf(ov[q].x, c, acc.x); acc.y = fmaf(ov[q].y, c, acc.y);
        acc.z = fmaf(ov[q].z, c, acc.z); acc.w = fmaf(ov[q].w, c, acc.w);
      }
      const float inv = 1.0f / L;
      return make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    };
    ds_stage_row<2>(o, load_attn, wait_dep, s_stat, xs_u32, xsum, tr);  // H * hs <= 8192 (checked by lp_decode_step_plan)
  } else {
    const float* xg = o.x;
    auto load_plain = [&](int k) -> float4 { return ds_ldcg4(xg + k); };
    if (small) ds_stage_row<2>(o, load_plain, wait_dep, s_stat, xs_u32, xsum, tr, xmode, xraw, xstat);
    else if (o.K <= 4 * DS_CTHREADS * 8) ds_stage_row<4>(o, load_plain, wait_dep, s_stat, xs_u32, xsum, tr);
    else ds_stage_row<6>(o, load_plain, wait_dep, s_stat, xs_u32, xsum, tr);  // e.g. falcon-7b mlp.proj, K = 18176
  }
  gs_bar_consumers();
  if (threadIdx.x == 0) {
    // what the tile epilogue multiplies the accumulator columns with: digit weights (int4) x the scalar that was pulled out of the
    // product (RMSNorm: rstd from the sum of squares the 16 warps left in s_stat).  Written before this warp's first tile arrival.
    float post = 1.0f;
    if (o.norm_kind == LP_NORM_RMS) {
      float ss = 0.f;
#pragma unroll
      for (int w = 0; w < GS_CWARPS; ++w) ss += s_stat[2 * GS_CWARPS + w];
      post = 1.0f / sqrtf(ss / (float)o.K + o.eps);
    }
    const bool i4 = o.fmt == LP_W_INT4 || o.fmt == LP_W_INT8;  // int8 digit planes of the activations
    colscale[0] = post;
    colscale[1] = i4 ? 256.0f * post : post;
    colscale[2] = i4 ? 65536.0f * post : post;
    if (tr) tr[2] = gs_now();
  }
  const uint32_t xsum_u32 = gs_smem_u32(xsum);
  if (o.fmt == LP_W_INT8) ds_linear_main_bnb<LP_W_INT8>(o, rg, red_u32, xsum_u32, xs_u32, tbl_u32, done, sb, se, gt);
  else if (o.fmt == LP_W_NF4) ds_linear_main_bnb<LP_W_NF4>(o, rg, red_u32, xsum_u32, xs_u32, tbl_u32, done, sb, se, gt);
  else if (o.fmt == LP_W_BF16) ds_linear_main<LP_W_BF16, false>(o, rg, red_u32, xsum_u32, xs_u32, done, sb, se, gt);
  else if (p.i4pair && o.aux_bytes == 4) ds_linear_main_i4pair<true>(o, rg, red_u32, xsum_u32, xs_u32, done, sb, se, gt, p.i4pair == 2);
  else if (p.i4pair) ds_linear_main_i4pair<false>(o, rg, red_u32, xsum_u32, xs_u32, done, sb, se, gt, p.i4pair == 2);
  else if (o.aux_bytes == 4) ds_linear_main<LP_W_INT4, true>(o, rg, red_u32, xsum_u32, xs_u32, done, sb, se, gt);
  else ds_linear_main<LP_W_INT4, false>(o, rg, red_u32, xsum_u32, xs_u32, done, sb, se, gt);
}

// Epilogue warp `e` finalises the tiles of parity e: cross-warp reduction of the 16 partial sums, digit / term recombination,
// bias, activation, residual, store.  The consumers never stop for this: they drop their partials and stream on.
__device__ __forceinline__ void ds_epilogue_warp(const DsParams& p, int e, uint32_t red_u32, const float* colscale, volatile int* done,
                                                 const DsOp* s_ops, float* s_loc) {
  const int lane = threadIdx.x & 31;
  const int rr = lane & 15, half = lane >> 4;
  int gt = 0;
  // tensor parallel: epoch counters of the two exchange slots as of the start of this step (the senders publish epoch + use + 1)
  unsigned int tp_epoch[2] = {0u, 0u};
  if (p.tp_state0) tp_epoch[0] = *reinterpret_cast<volatile unsigned int*>(p.tp_state0);
  if (p.tp_state1) tp_epoch[1] = *reinterpret_cast<volatile unsigned int*>(p.tp_state1);
  asm volatile("bar.sync 4, %0;\n" ::"n"(DS_OPEND_THREADS) : "memory");  // record of op 0 is in shared memory
  for (int op = 0; op < p.nops; ++op) {
    const DsOp& o = s_ops[op % DS_NREC];
    const int signal = o.signal;
    const int hsync = o.kind == DS_KIND_LINEAR || o.kind == DS_KIND_SLAB ? -1 : o.hsync;
    if (o.kind == DS_KIND_LINEAR) {
      const int nks = o.nks, fmt = o.fmt, split = o.split, epi = o.epi, streamk = o.streamk, keep_local = o.keep_local;
      const int push = o.tp_size;  // > 0: row-parallel sender of a tensor-parallel exchange (see ds_exchange)
      const unsigned long long* push_bufs = o.tp_bufs;
      const unsigned long long push_off = o.tp_buf_off;
      const float push_tag = __uint_as_float((o.tp_state == p.tp_state1 ? tp_epoch[1] : tp_epoch[0]) + (unsigned)o.tp_use + 1u);
      const float* bias = o.bias;
      const float* residual = o.residual;
      const float* rowscale = reinterpret_cast<const float*>(o.aux2);  // INT8 only
      float* out = o.out;
      int sb, se;
      ds_stage_range(o, sb, se);
      const int loc0 = (sb / nks) * (GS_ROWS / 2);  // keep_local (SwiGLU): first output of this CTA
      for (int s0 = sb; s0 < se; ++gt) {
        const int tile = s0 / nks;
        const bool first = s0 == tile * nks;  // this CTA's part of the tile starts at K = 0: it adds the bias
        s0 = min(se, (tile + 1) * nks);
        if ((gt & 1) != e) continue;
        if (e == 0) asm volatile("bar.sync 2, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
        else asm volatile("bar.sync 3, %0;\n" ::"n"(DS_TILE_BAR_THREADS) : "memory");
        // lane = (row of the tile, half): each half adds the partial sums of 8 warps, then one shuffle
        const uint32_t rb = red_u32 + (uint32_t)(((e * GS_CWARPS + half * (GS_CWARPS / 2)) * 16 + rr) * 4) * 4;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int w = 0; w < GS_CWARPS / 2; ++w) {
          const float4 v = ds_lds128f(rb + (uint32_t)w * 16 * 4 * 4);
          c0 += v.x;
          c1 += v.y;
          c2 += v.z;
        }
        c0 += __shfl_xor_sync(0xffffffffu, c0, 16);
        c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
        c2 += __shfl_xor_sync(0xffffffffu, c2, 16);
        __syncwarp();
        if (lane == 0) done[e] = (gt >> 1) + 1;  // the partial sums are in registers: the buffer may be rewritten
        float y;
        if (fmt == LP_W_INT4 || fmt == LP_W_INT8) {  // digit 0 (smallest) first
          y = c0 * colscale[0];
          y = fmaf(c1, colscale[1], y);
          y = fmaf(c2, colscale[2], y);
          if (fmt == LP_W_INT8) y *= rowscale[tile * GS_ROWS + rr];  // SCB / 127 of the output row
        } else {  // last (smallest) bf16 term first
          y = split == 3 ? c2 : 0.f;
          y += c1;
          y += c0;
          y *= colscale[0];  // 1, or the RMSNorm scale pulled out of the product
        }
        const int row = tile * GS_ROWS + rr;
        if (bias && first) y += bias[row];
        if (streamk) {
          if (half == 0) atomicAdd(out + row, y);  // x += (partial) W . u, in place
        } else if (epi == LP_EPI_SWIGLU) {
          const float other = __shfl_xor_sync(0xffffffffu, y, 1);  // fc_2 row of the pair
          if (half == 0 && (rr & 1) == 0) {
            if (keep_local) s_loc[(row >> 1) - loc0] = silu(y) * other;  // consumed by this CTA's slab op, never leaves the SM
            else out[row >> 1] = silu(y) * other;
          }
        } else if (push > 0) {
          // this rank's partial of 16 rows -> slot [s][rank] of every rank's buffer as {value, epoch} pairs (128 contiguous
          // bytes per peer; 8-byte stores: whoever sees the epoch sees the value)
          if (half == 0) {
#pragma unroll
            for (int r = 0; r < DS_MAX_TP; ++r)
              if (r < push) reinterpret_cast<float2*>(push_bufs[r] + push_off)[row] = make_float2(y, push_tag);
          }
        } else if (half == 0) {
          if (epi == LP_EPI_GELU) y = gelu_erf(y);
          else if (epi == LP_EPI_RESIDUAL) y = __ldcg(residual + row) + y;
          out[row] = y;
        }
      }
    }
    // end of the op: this warp's rows are written.  bar.sync, not arrive: an epilogue warp must not run ahead into the next op's
    // barrier phase.  One lane of epilogue warp 0 then publishes the op: release at gpu scope covers the rows written by ALL warps
    // of this CTA (ordered before by the barrier).  The ~1 us of that fence is off the consumers' path: they are already staging
    // the next op.
    asm volatile("bar.sync 4, %0;\n" ::"n"(DS_OPEND_THREADS) : "memory");
    if (e == 0 && lane == 0) {
      if (signal) ds_red_release(p.counters + op);
      // attention feeding a slab: the P CTAs of a head only wait for each other (per-head counter), not for the grid
      if (hsync >= 0 && (int)blockIdx.x < p.H * p.P) ds_red_release(p.counters + hsync + blockIdx.x / p.P);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention op (consumers)
// Head h, sequence split sp of this CTA: flash-decoding with WARP-private state.  Of every 16 KB key tile (AT keys) warp w owns
// keys w*KPW .. w*KPW+KPW-1: it computes their scores (TPK lanes per key, 16 dims each), keeps its own running (m, l, o[hs]) —
// o spread over the lanes, hs/32 dims each — and never synchronises with the other warps inside the block loop: a warp only
// waits for the ring stage it needs and releases it when done, exactly like the linear ops.  The 16 warp states are merged once
// per (head, split) through shared memory.  The new token's k / v row (RoPE'd here, appended to the cache by the CTA that owns
// its block) is read from shared memory instead of the landed tile.
template <int HS>
__device__ __forceinline__ void ds_attention(const DsParams& p, const DsOp& o, const DsAttnGeo<HS>& geo, DsRing& rg, unsigned char* scratch,
                                             unsigned long long* tr) {
  using G = DsAttnGeo<HS>;
  constexpr int TPK = G::TPK, AT = G::AT, LDM = G::LDM;
  constexpr int KPW = 32 / TPK;   // keys per warp per tile
  constexpr int DPL = HS / 32;    // output dims per lane
  constexpr int LDW = HS + 4;     // warp record: o[HS], m, l
  float* sQ = reinterpret_cast<float*>(scratch);                              // [HS] rotated q * scale * log2(e)
  __nv_bfloat16* sNew = reinterpret_cast<__nv_bfloat16*>(sQ + HS);            // [2][HS] rotated new k, new v
  float* sW = reinterpret_cast<float*>(sNew + 2 * HS);                        // [warps][LDW]
  float* sZero = sW + GS_CWARPS * LDW;                                        // 16 zero bytes
  const uint32_t sNew_u32 = gs_smem_u32(sNew);
  const int ctid = threadIdx.x, warp = ctid >> 5, lane = ctid & 31;
  const int kq = lane / TPK, dc = lane % TPK;
  const int npairs = p.H * p.P;
  const int half = p.n_elem >> 1;
  for (int j = blockIdx.x; j < npairs; j += gridDim.x) {
    const int h = j / p.P, sp = j % p.P;
    const int qpk = p.H / p.G, g = h / qpk;  // KV group of this q head: its K / V tiles are shared by qpk heads (CTAs)
    const int b0 = sp * geo.bpp, b1 = min(geo.nblk, b0 + geo.bpp);
    if (b0 >= b1) continue;  // CTA-uniform; the producer applies the same rule
    const bool patch = geo.slot_blk >= b0 && geo.slot_blk < b1;
    // ---- q (and, in the CTA that owns the new token's block, k_new / v_new): RoPE, scale; cache append ----
    {
      // rows of group g in the raw projection: [q x qpk | k | v] (model.py:210-214); MHA qpk = 1, MQA one group
      const float* grp = o.qkv + (size_t)g * (qpk + 2) * HS;
      const int nrows = patch ? 3 : 1;
      for (int i = ctid; i < nrows * HS; i += DS_CTHREADS) {
        const int r = i / HS, d = i % HS;
        const float* src = grp + (size_t)(r == 0 ? h - g * qpk : qpk + r - 1) * HS;
        float v = __ldcg(src + d);
        if (r <= 1 && d < p.n_elem) {
          const float partner = (d < half) ? -__ldcg(src + d + half) : __ldcg(src + d - half);
          const float c = p.cosT[(size_t)geo.pos * p.n_elem + d], s = p.sinT[(size_t)geo.pos * p.n_elem + d];
          v = __fadd_rn(__fmul_rn(v, c), __fmul_rn(partner, s));  // same op order as apply_rope (model.py:330-336)
        }
        if (i < 4) sZero[i] = 0.f;
        if (r == 0) {
          sQ[d] = v * p.scale_log2;
        } else {
          const __nv_bfloat16 hb = __float2bfloat16_rn(v);
          sNew[(r - 1) * HS + d] = hb;
          // the first q head of the group appends to the cache; the others only need the row in shared memory
          if (h == g * qpk) (r == 1 ? o.kc : o.vc)[((size_t)g * p.max_seq + geo.slot) * HS + d] = hb;
        }
      }
    }
    gs_bar_consumers();
    if (tr && ctid == 0) tr[4] = gs_now();  // q staged
    float q[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      q[i] = sQ[dc * 8 + i];
      q[8 + i] = sQ[(dc + TPK) * 8 + i];
    }
    float m_run = -CUDART_INF_F, l_lane = 0.f, ov[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) ov[i] = 0.f;
    // loop invariants, pinned (see ds_pin): ring geometry, this lane's byte offsets inside a K / V tile
    const uint32_t ring_u32 = ds_pin(rg.ring_u32), bar0 = ds_pin(rg.bar0);
    const int nstages = ds_pin(rg.nstages), stage_stride = ds_pin(rg.stage_stride);
    const uint32_t koff = ds_pin((uint32_t)((warp * KPW + kq) * HS * 2 + dc * 16));  // this lane's key row, first 8-dim chunk
    const uint32_t voff = ds_pin((uint32_t)(warp * KPW * HS * 2 + lane * (DPL * 2)));  // first key row of the warp, this lane's dims
    const uint32_t knew = ds_pin(sNew_u32 + dc * 16), vnew = ds_pin(sNew_u32 + HS * 2 + lane * (DPL * 2));
    const uint32_t vzero = ds_pin(gs_smem_u32(sZero));
    const bool lane0 = ds_pin(lane) == 0;
    const bool dc0 = ds_pin(dc) == 0;
    int rs = rg.s, rph = rg.ph;

    for (int blk = b0; blk < b1; ++blk) {
      const int key0 = blk * AT + warp * KPW;  // first key of this warp in the tile
      const int nvalid = geo.kv_len - key0;    // keys key0 .. key0 + nvalid - 1 exist
      const int pslot = (patch && blk == geo.slot_blk) ? geo.slot - key0 : -1;  // index of the new token among the warp's keys
      // ---- K tile: scores of this warp's keys ----
      mbar_wait(bar0 + 8 * rs, rph);
      if (tr && ctid == 0 && blk == b0) tr[5] = gs_now();  // first K tile present
      float sc;
      {
        const uint32_t rowa = kq == pslot ? knew : ring_u32 + (uint32_t)(rs * stage_stride) + koff;
        const uint4 ka = ds_lds128(rowa), kb = ds_lds128(rowa + TPK * 16);
        // four independent chains
        float s0 = q[0] * bf16lo(ka.x), s1 = q[2] * bf16lo(ka.y), s2 = q[4] * bf16lo(ka.z), s3 = q[6] * bf16lo(ka.w);
        s0 = fmaf(q[1], bf16hi(ka.x), s0); s1 = fmaf(q[3], bf16hi(ka.y), s1);
        s2 = fmaf(q[5], bf16hi(ka.z), s2); s3 = fmaf(q[7], bf16hi(ka.w), s3);
        s0 = fmaf(q[8], bf16lo(kb.x), s0); s1 = fmaf(q[10], bf16lo(kb.y), s1);
        s2 = fmaf(q[12], bf16lo(kb.z), s2); s3 = fmaf(q[14], bf16lo(kb.w), s3);
        s0 = fmaf(q[9], bf16hi(kb.x), s0); s1 = fmaf(q[11], bf16hi(kb.y), s1);
        s2 = fmaf(q[13], bf16hi(kb.z), s2); s3 = fmaf(q[15], bf16hi(kb.w), s3);
        sc = (s0 + s1) + (s2 + s3);
      }
      __syncwarp();
      if (lane0) mbar_arrive(bar0 + 8 * (nstages + rs));
      if (++rs == nstages) { rs = 0; rph ^= 1; }
#pragma unroll
      for (int off = 1; off < TPK; off <<= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);
      if (kq >= nvalid) sc = -CUDART_INF_F;  // rows past the end of the sequence hold stale bytes
      float mx = sc;
#pragma unroll
      for (int off = TPK; off < 32; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      // ---- V tile: online softmax of the warp's keys, P.V ----
      mbar_wait(bar0 + 8 * rs, rph);
      if (mx > -CUDART_INF_F) {  // warp-uniform: at least one valid key
        const uint32_t tv = ring_u32 + (uint32_t)(rs * stage_stride) + voff;
        const float m_new = fmaxf(m_run, mx);
        const float corr = exp2f(m_run - m_new);  // 0 for the first tile (m_run = -inf, m_new finite)
        const float pk = exp2f(sc - m_new);       // 0 for keys past the end
        m_run = m_new;
        l_lane = fmaf(l_lane, corr, dc0 ? pk : 0.f);
#pragma unroll
        for (int i = 0; i < DPL; ++i) ov[i] *= corr;
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
          const float pkk = __shfl_sync(0xffffffffu, pk, kk * TPK);
          // keys past the end read a zero row (their stage bytes are stale: keep NaN bit patterns out of 0 * v)
          const uint32_t rowa = kk == pslot ? vnew : (kk < nvalid ? tv + kk * HS * 2 : vzero);
          if (DPL == 4) {
            const uint2 vv = ds_lds64(rowa);
            ov[0] = fmaf(pkk, bf16lo(vv.x), ov[0]);
            ov[1] = fmaf(pkk, bf16hi(vv.x), ov[1]);
            ov[DPL - 2] = fmaf(pkk, bf16lo(vv.y), ov[DPL - 2]);
            ov[DPL - 1] = fmaf(pkk, bf16hi(vv.y), ov[DPL - 1]);
          } else {
            const uint32_t vv = ds_lds32(rowa);
            ov[0] = fmaf(pkk, bf16lo(vv), ov[0]);
            ov[1] = fmaf(pkk, bf16hi(vv), ov[1]);
          }
        }
      }
      __syncwarp();
      if (lane0) mbar_arrive(bar0 + 8 * (nstages + rs));
      if (++rs == nstages) { rs = 0; rph ^= 1; }
    }
    rg.s = rs;
    rg.ph = rph;
    // ---- merge the 16 warp states, write this split's partial ----
    if (tr && ctid == 0) tr[6] = gs_now();  // block loop done (warp 0)
    {
      const float lw = warp_sum(l_lane);
      float* rec = sW + warp * LDW;
#pragma unroll
      for (int i = 0; i < DPL; ++i) rec[lane * DPL + i] = ov[i];
      if (lane == 0) {
        rec[HS] = m_run;
        rec[HS + 1] = lw;
      }
    }
    gs_bar_consumers();
    {
      // 4 lanes per output dim (HS 128: all 512 threads; HS 64: the first 256), each folds 4 of the 16 warp records
      const int d = ctid >> 2, qd = ctid & 3;
      if (d < HS) {
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < GS_CWARPS; ++w) mx = fmaxf(mx, sW[w * LDW + HS]);
        float acc = 0.f, L = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < GS_CWARPS / 4; ++w4) {  // fixed order: deterministic
          const int w = qd * (GS_CWARPS / 4) + w4;
          const float c = exp2f(sW[w * LDW + HS] - mx);  // 0 for warps that saw no valid key
          acc = fmaf(sW[w * LDW + d], c, acc);
          L = fmaf(sW[w * LDW + HS + 1], c, L);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        L += __shfl_xor_sync(0xffffffffu, L, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        L += __shfl_xor_sync(0xffffffffu, L, 2);
        if (qd == 0) {
          float* dst = p.part + ((size_t)h * p.P + sp) * LDM;
          dst[d] = acc;
          if (d == 0) {
            dst[HS] = mx;
            dst[HS + 1] = L;
          }
        }
      }
    }
    gs_bar_consumers();  // scratch is reused by the next pair / the next op's staging
  }
}


__constant__ float c_ds_nf4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

struct DsSlabLoop {
  uint32_t sc_off, ring_u32, bar0;
  int nstages, stage_stride, nrb, stagger, dbg;
  bool lane0, t0, t0e;  // t0e: t == 0 and g even: the lane that issues the 4-row reduction
  float* out_row;
  const float* bias;  // NULL, or the bias shifted to this lane's first row
  float a_seg[2], zs0[2], zs1[2];
};

// Stage loop of ds_slab for NCHA chunks of the first scale group and NCH chunks in all.  Two stages at a time (the stage counts
// are even): per stage a warp has ONE short dependent chain (wait -> loads -> IMMA -> IMMA -> fold -> shuffle -> red) and all 16
// warps sit on the same stage, so interleaving the chains of two stages (same digit words, two weight tiles) overlaps them.
template <int NCHA, int NCH>
__device__ __forceinline__ void ds_slab_loop(const DsSlabLoop& L, const uint32_t (&woff)[3][2], const uint32_t (&doff)[3][2], int& rs_io,
                                             int& rph_io) {
  const uint32_t ML = 0x0F0F0F0Fu, MH = 0xF0F0F0F0u;
  const uint32_t sc_off = L.sc_off, ring_u32 = L.ring_u32, bar0 = L.bar0;
  const int nstages = L.nstages, stage_stride = L.stage_stride, nrb = L.nrb;
  int rs = rs_io, rph = rph_io;
  int rba = L.stagger;  // the pair (rba, rba + 1): nrb and the pair index are even, the stagger wraps as a whole pair
  auto fold = [&](const int (&cl)[4], const int (&ch)[4], uint32_t stb, int sg, float (&acc)[4]) {
    const uint2 sz = ds_lds64(stb + sc_off + (uint32_t)(sg * 1024));  // packed scale / zero of rows 2g, 2g + 1
    const float s0 = __uint_as_float(sz.x & 0xffff0000u) * L.a_seg[sg], s1 = __uint_as_float(sz.y & 0xffff0000u) * L.a_seg[sg];
    const float z0 = __uint_as_float(sz.x << 16), z1 = __uint_as_float(sz.y << 16);
    // sum (q - z) s x = s * (sum q X - z * sum X), all integers exact in fp32
    acc[0] = fmaf(s0, fmaf((float)ch[0], 0.0625f, fmaf(-z0, L.zs0[sg], (float)cl[0])), acc[0]);
    acc[1] = fmaf(s0, fmaf((float)ch[1], 0.0625f, fmaf(-z0, L.zs1[sg], (float)cl[1])), acc[1]);
    acc[2] = fmaf(s1, fmaf((float)ch[2], 0.0625f, fmaf(-z1, L.zs0[sg], (float)cl[2])), acc[2]);
    acc[3] = fmaf(s1, fmaf((float)ch[3], 0.0625f, fmaf(-z1, L.zs1[sg], (float)cl[3])), acc[3]);
  };
  auto emit = [&](const float (&acc)[4], int rb) {
    // digit recombination: lane t = 0 holds planes 0 and 1, lane t = 1 plane 2 of rows 2g, 2g + 1
    float y0 = L.t0 ? fmaf(acc[1], 256.0f, acc[0]) : acc[0] * 65536.0f;
    float y1 = L.t0 ? fmaf(acc[3], 256.0f, acc[2]) : acc[2] * 65536.0f;
    y0 += __shfl_xor_sync(0xffffffffu, y0, 1);
    y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
    // rows 2g, 2g + 1 of this lane and 2g + 2, 2g + 3 of lane g + 1 (4 lanes up): ONE 16-byte reduction per 4 rows
    const float y2 = __shfl_down_sync(0xffffffffu, y0, 4), y3 = __shfl_down_sync(0xffffffffu, y1, 4);
    if (L.t0e) {
      float* dst = L.out_row + rb * DS_SLAB_ROWS;
      float4 yy = make_float4(y0, y1, y2, y3);
      if (L.bias) {
        const float4 bb = *reinterpret_cast<const float4*>(L.bias + rb * DS_SLAB_ROWS);
        yy.x += bb.x; yy.y += bb.y; yy.z += bb.z; yy.w += bb.w;
      }
      if (L.dbg & 4) *reinterpret_cast<float4*>(dst) = yy;
      else if (!(L.dbg & 1))
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "f"(yy.x), "f"(yy.y), "f"(yy.z), "f"(yy.w) : "memory");
    }
  };
  for (int i = 0; i < nrb; i += 2) {
    int rs2 = rs + 1, rph2 = rph;
    if (rs2 == nstages) { rs2 = 0; rph2 ^= 1; }
    mbar_wait(bar0 + 8 * rs, rph);
    mbar_wait(bar0 + 8 * rs2, rph2);
    const uint32_t sta = ring_u32 + (uint32_t)(rs * stage_stride), stb = ring_u32 + (uint32_t)(rs2 * stage_stride);
    float acca[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
    int cla[4] = {0, 0, 0, 0}, cha[4] = {0, 0, 0, 0}, clb[4] = {0, 0, 0, 0}, chb[4] = {0, 0, 0, 0};
    if (!(L.dbg & 32))
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint2 d0 = ds_lds64(doff[c][0]), d1 = ds_lds64(doff[c][1]);
      const uint2 wa0 = ds_lds64(sta + woff[c][0]), wa1 = ds_lds64(sta + woff[c][1]);
      const uint2 wb0 = ds_lds64(stb + woff[c][0]), wb1 = ds_lds64(stb + woff[c][1]);
      if (!(L.dbg & 8)) {
      gs_imma(cla, wa0.x & ML, wa0.y & ML, wa1.x & ML, wa1.y & ML, d0.x, d1.x);
      gs_imma(clb, wb0.x & ML, wb0.y & ML, wb1.x & ML, wb1.y & ML, d0.x, d1.x);
      gs_imma(cha, wa0.x & MH, wa0.y & MH, wa1.x & MH, wa1.y & MH, d0.y, d1.y);
      gs_imma(chb, wb0.x & MH, wb0.y & MH, wb1.x & MH, wb1.y & MH, d0.y, d1.y);
      } else {
        cla[0] += wa0.x + wa1.y + d0.x; clb[0] += wb0.x + wb1.y + d1.x; cha[0] += wa0.y + wa1.x + d0.y; chb[0] += wb0.y + wb1.x + d1.y;
      }
      if ((c == NCHA - 1 || c == NCH - 1) && !(L.dbg & 16)) {  // compile time: last chunk of a scale group
        fold(cla, cha, sta, c < NCHA ? 0 : 1, acca);
        fold(clb, chb, stb, c < NCHA ? 0 : 1, accb);
        if (c != NCH - 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) cla[q] = cha[q] = clb[q] = chb[q] = 0;
        }
      }
    }
    __syncwarp();
    if (L.lane0) {
      mbar_arrive(bar0 + 8 * (nstages + rs));
      mbar_arrive(bar0 + 8 * (nstages + rs2));
    }
    rs = rs2;
    rph = rph2;
    if (++rs == nstages) { rs = 0; rph ^= 1; }
    if (!(L.dbg & 64)) {
      emit(acca, rba);
      emit(accb, rba + 1);
    } else if (cla[0] + clb[0] + cha[0] + chb[0] == 0x7fffffff) {
      L.out_row[0] = acca[0] + accb[0];
    }
    rba += 2;
    if (rba >= nrb) rba -= nrb;
  }
  rs_io = rs;
  rph_io = rph;
}

// ------------------------------------------------------------------------------------------------ slab op (consumers)
// out += W[:, cols] . v  for the input columns THIS CTA produced itself: its SwiGLU outputs of the preceding up-projection (72 / 80
// columns of mlp.proj) or the attention output of its head (hs = 128 columns of attn.proj, rows split over the head's P CTAs).
// The column -> row pairing of tensor parallelism applied inside one GPU: u / att never travel through L2, and fc -> mlp.proj,
// attention -> attn.proj stop being grid-wide dependencies (5 -> 3 per Llama layer); what is left is ONE reduction of the partial
// rows into the residual stream (red.global.add.v2.f32, staggered over the row blocks so that the CTAs hit different rows).
//   weights: the CTA's slab image (lp_decode_step_slab_build): per 256-row stage [unit][128 row pairs][2 rows x 4 B] — a word holds
//   the 8 int4 columns of one unit, row pair index XOR 4 * (unit & 3) (bank-conflict-free 64-bit loads) — then one KB of packed
//   (bf16 scale << 16 | bf16 zero) words per scale group touched.  One 16-row m16n8k32 tile per warp and stage: MMA row g <-> output
//   row 2g, row g + 8 <-> 2g + 1; k-quad t <-> unit c0 + t, the even columns of a unit go through the low-nibble IMMA, the odd
//   ones through the high-nibble IMMA (w & 0xF0F0F0F0 = 16 q), B columns 0..2 = the three int8 digit planes of v.
//   v: block fixed point per scale-group segment, X = rint(v * 2^22 / max|v_seg|), balanced base-256 digits (ds_stage_row).
template <int HS, class WaitHs>
__device__ __forceinline__ void ds_slab(const DsParams& p, const DsOp& o, const DsSlabMeta& mt, const DsAttnGeo<HS>& geo, DsRing& rg,
                                        unsigned char* xs, float* s_loc, WaitHs wait_hs, unsigned long long* tr) {
  const int nrb = mt.nrb;
  if (nrb == 0) return;  // CTA-uniform: this CTA has no slab (e.g. the CTAs beyond H * P in an attention projection)
  const int ctid = threadIdx.x, warp = ctid >> 5, lane = ctid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nunits = mt.nunits, units_a = mt.units_a, nseg = mt.nseg;
  uint32_t* dig = reinterpret_cast<uint32_t*>(xs);              // [3 planes][DS_SLAB_MAXU units][even word, odd word]
  unsigned* s_am = reinterpret_cast<unsigned*>(xs + 384);       // [2] max|v| per segment (float bits)
  int* s_sum = reinterpret_cast<int*>(xs + 392);                // [2][3] digit sums per segment and plane
  if (o.slab_src == 1) {
    // v = attention output of head h: merge the P sequence-split partials (m, l, o[hs]) once their CTAs have arrived
    wait_hs();
    if (tr && ctid == 0) tr[1] = gs_now();
    constexpr int LDM = DsAttnGeo<HS>::LDM;
    if (ctid < HS) {
      const int h = blockIdx.x / p.P;
      const int nvalid = (geo.nblk + geo.bpp - 1) / geo.bpp;
      const float* base = p.part + (size_t)h * p.P * LDM;
      float m[DS_MAXP], l[DS_MAXP], ov[DS_MAXP], mx = -CUDART_INF_F;
#pragma unroll
      for (int q = 0; q < DS_MAXP; ++q) {
        m[q] = -CUDART_INF_F;
        l[q] = ov[q] = 0.f;
        if (q < nvalid) {
          m[q] = __ldcg(base + q * LDM + HS);
          l[q] = __ldcg(base + q * LDM + HS + 1);
          ov[q] = __ldcg(base + q * LDM + ctid);
        }
        mx = fmaxf(mx, m[q]);
      }
      float L = 0.f, acc = 0.f;
#pragma unroll
      for (int q = 0; q < DS_MAXP; ++q) {  // split order: deterministic
        const float c = exp2f(m[q] - mx);
        L = fmaf(l[q], c, L);
        acc = fmaf(ov[q], c, acc);
      }
      s_loc[ctid] = acc / L;
    }
  }
  if (ctid < 2) s_am[ctid] = 0u;
  if (ctid < 6) s_sum[ctid] = 0;
  if (ctid < 2) reinterpret_cast<uint32_t*>(xs + 416)[ctid] = 0u;  // an all-zero digit word pair for the empty k-quads of a chunk
  gs_bar_consumers();
  const int nval = nunits * 8;
  float v = 0.f;
  const int my_seg = (ctid >> 3) >= units_a ? 1 : 0;
  if (ctid < nval) {
    v = s_loc[ctid];
    atomicMax(&s_am[my_seg], __float_as_uint(fabsf(v)));  // non-negative floats order like their bit patterns
  }
  gs_bar_consumers();
  if (ctid < nval) {
    const float am = __uint_as_float(s_am[my_seg]);
    const float inv = am > 0.f ? 4194304.0f / am : 0.f;
    const uint32_t z = (uint32_t)(__float2int_rn(v * inv) + 0x8080) ^ 0x8080u;  // bytes 0..2 = balanced digits (ds_stage_row)
    const int unit = ctid >> 3, col = ctid & 7;
    unsigned char* db = reinterpret_cast<unsigned char*>(dig);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int b = (int)(signed char)((z >> (8 * d)) & 0xffu);
      db[((d * DS_SLAB_MAXU + unit) * 2 + (col & 1)) * 4 + (col >> 1)] = (unsigned char)b;
      if (b != 0) atomicAdd(&s_sum[my_seg * 3 + d], b);
    }
  }
  gs_bar_consumers();
  if (tr && ctid == 0) tr[2] = gs_now();
  // per-lane constants of the two segments: activation scale, digit sums of this lane's accumulator columns (2t, 2t + 1)
  float a_seg[2], zs0[2], zs1[2];
#pragma unroll
  for (int sg = 0; sg < 2; ++sg) {
    a_seg[sg] = __uint_as_float(s_am[sg]) * (1.0f / 4194304.0f);
    zs0[sg] = t == 0 ? (float)s_sum[sg * 3 + 0] : (float)s_sum[sg * 3 + 2];
    zs1[sg] = t == 0 ? (float)s_sum[sg * 3 + 1] : 0.f;
  }
  // Chunks of 8 units (one low + one high IMMA each): ceil(units_a / 8) for the first scale group, then those of the second; at most
  // 3 for 16 units.  Everything a lane needs per chunk is loop invariant: the byte offsets of its two weight words (k-quads t and
  // t + 4) inside a stage and the addresses of the matching digit words — an empty k-quad reads unit 0's weights against the
  // all-zero digit word, so the stage loop has no lane-divergent branch and no address arithmetic beyond `stage base + offset`.
  const int plane = g < 3 ? g : 2;  // B column n = g: digit plane g; columns 3..7 feed accumulator columns nobody reads
  const uint32_t dig_u32 = gs_smem_u32(dig), zero_u32 = gs_smem_u32(xs + 416);
  const uint32_t pair_off = (uint32_t)(8 * warp + g);  // row pair of this lane inside the 256-row stage
  const int nch_a = (units_a + 7) >> 3, nch = nch_a + (nseg > 1 ? (nunits - units_a + 7) >> 3 : 0);
  uint32_t woff[3][2], doff[3][2];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int sg = c < nch_a ? 0 : 1, ci = c - (sg ? nch_a : 0);
    const int ub = sg ? units_a : 0, ue = sg ? nunits : units_a;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int j = ub + 8 * ci + 4 * hf + t;
      const bool valid = c < nch && j < ue;
      const int jj = valid ? j : 0;
      woff[c][hf] = ds_pin((uint32_t)(jj * 1024) + ((pair_off ^ (uint32_t)(4 * (jj & 3))) << 3));
      doff[c][hf] = ds_pin(valid ? dig_u32 + (uint32_t)((plane * DS_SLAB_MAXU + j) * 8) : zero_u32);
    }
  }
  const uint32_t ring_u32 = ds_pin(rg.ring_u32), bar0 = ds_pin(rg.bar0);
  const int nstages = ds_pin(rg.nstages), stage_stride = ds_pin(rg.stage_stride);
  const uint32_t sc_off = ds_pin((uint32_t)(nunits * 1024 + (16 * warp + 2 * g) * 4));
  const bool lane0 = ds_pin(lane) == 0, t0 = ds_pin(t) == 0;
  const int stagger = (p.dbg & 2) ? 0 : (2 * (o.slab_src == 1 ? (int)blockIdx.x / p.P : (int)blockIdx.x)) % nrb;  // whole stage pairs
  const float* bias = o.bias;
  const bool add_bias = bias && (o.slab_src == 1 ? (int)blockIdx.x / p.P == 0 : blockIdx.x == 0);
  float* out_row = o.out + mt.row0 + 16 * warp + 2 * g;
  int rs = rg.s, rph = rg.ph;
  const uint32_t ML = 0x0F0F0F0Fu, MH = 0xF0F0F0F0u;
  DsSlabLoop lp_;
  lp_.sc_off = sc_off;
  lp_.ring_u32 = ring_u32;
  lp_.bar0 = bar0;
  lp_.nstages = nstages;
  lp_.stage_stride = stage_stride;
  lp_.nrb = nrb;
  lp_.stagger = stagger;
  lp_.lane0 = lane0;
  lp_.t0 = t0;
  lp_.t0e = t0 && (g & 1) == 0;
  lp_.out_row = out_row;
  lp_.bias = add_bias ? bias - (o.out - out_row) : nullptr;  // indexed like out_row
  lp_.dbg = p.dbg;
#pragma unroll
  for (int sg = 0; sg < 2; ++sg) {
    lp_.a_seg[sg] = a_seg[sg];
    lp_.zs0[sg] = zs0[sg];
    lp_.zs1[sg] = zs1[sg];
  }
  // straight-line stage loops for the chunk patterns that occur (first scale group: 1 or 2 chunks, second: 0, 1 or 2)
  const int pat = nch_a * 4 + nch;
  if (pat == 1 * 4 + 1) ds_slab_loop<1, 1>(lp_, woff, doff, rs, rph);
  else if (pat == 2 * 4 + 2) ds_slab_loop<2, 2>(lp_, woff, doff, rs, rph);
  else if (pat == 1 * 4 + 2) ds_slab_loop<1, 2>(lp_, woff, doff, rs, rph);
  else if (pat == 2 * 4 + 3) ds_slab_loop<2, 3>(lp_, woff, doff, rs, rph);
  else ds_slab_loop<1, 3>(lp_, woff, doff, rs, rph);
  rg.s = rs;
  rg.ph = rph;
  if (tr && ctid == 0) {
    tr[5] = gs_now();
    tr[7] = (unsigned long long)((nunits << 8) | nseg);
  }
  gs_bar_consumers();  // xs / s_loc are reused by the next op's staging
}

// ------------------------------------------------------------------------------------------------ tensor-parallel exchange op
// One-shot all-reduce over NVLink peer memory inside the step kernel: PUSH style, flag in the data (the LL protocol of NCCL).
// The row-parallel linear op before it is the sender: its tile epilogue stores every finished 16-row piece of this rank's partial
// straight into slot [s][rank] of EVERY rank's symmetric buffer as 8-byte {value, epoch} pairs (remote st.global.v2 over NVLink:
// an aligned 8-byte store is single-copy atomic, so a reader that sees the epoch sees the value), overlapped with the rest of the
// op.  No fence, no flag, no counter crosses the link.  This op is the receiver: every CTA polls ITS slice of the row
// (n / #CTAs elements x tp ranks, LOCAL memory) until all tags carry the slot's epoch, adds the tp values in rank order —
// bit-identical on every rank — plus the residual and stores the row.  One NVLink traversal on the critical path (the pull protocol
// of round 1 had three: flag, remote-load request, reply; a flag-after-data push still pays a system-scope fence per CTA and a
// flag store per peer: measured 15-22 us per exchange at tp = 8).  `dep` = the sender op keeps the LOCAL ordering: this rank's x
// may only be overwritten when all its CTAs are done reading it.  Two slots alternate; a slot's epoch grows by one per use, so
// stale pairs of the previous use never match.  A rank can be at most one exchange ahead of its slowest peer (it needs that peer's
// pairs to get past an exchange), so a slot is never rewritten while a peer still reads it.  Buffers and state of this protocol
// are disjoint from lp_tp_allreduce_residual (the per-op pull kernel of prefill / batches): the two may alternate freely.
__device__ __forceinline__ void ds_exchange(const DsParams& p, const DsOp& o, int op, unsigned int epoch0) {
  const int tid = threadIdx.x, tp = o.tp_size;
  const unsigned int epoch = epoch0 + (unsigned)o.tp_use + 1u;
  const int n = o.N;
  const float2* local = reinterpret_cast<const float2*>(o.tp_bufs[o.tp_rank] + o.tp_buf_off);  // [tp][n] {value, epoch} pairs
  const int i0 = (int)((long long)n * blockIdx.x / gridDim.x), i1 = (int)((long long)n * (blockIdx.x + 1) / gridDim.x);
  // one thread per (element, rank): 8 neighbouring lanes hold the tp pairs of one element, every thread polls ONE pair (all loads of
  // the slice are in flight at once: the op is a single L2 / NVLink latency, not tp of them), then a fixed shuffle tree adds them —
  // the same tree on every GPU, so the sums are bit-identical across ranks (rank order within the tree: ((0+1)+(2+3))+((4+5)+(6+7)))
  for (int base = i0; base < i1; base += DS_CTHREADS / DS_MAX_TP) {
    const int i = base + (tid >> 3), r = tid & (DS_MAX_TP - 1);
    float val = 0.f;
    if (i < i1 && r < tp) {
      const float2* src = local + (size_t)r * n + i;
      float2 v;
      unsigned it = 0;
      const unsigned long long t0 = gs_now();
      for (;;) {
        asm volatile("ld.relaxed.sys.global.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "l"(src) : "memory");
        if (__float_as_uint(v.y) == epoch) break;
        if (p.timeout_ns) {  // a peer GPU that never delivers: watchdog (ds_report)
          const bool expired = gs_now() - t0 > p.timeout_ns;
          if (expired) ds_report(p.err, DS_ERR_EXCHANGE_TIMEOUT, op, r, __float_as_uint(v.y));
          if (expired || ((++it & 15u) == 0 && ds_ld_relaxed(p.err) != 0u)) break;
        }
      }
      val = v.x;
    }
    val += __shfl_xor_sync(0xffffffffu, val, 1);
    val += __shfl_xor_sync(0xffffffffu, val, 2);
    val += __shfl_xor_sync(0xffffffffu, val, 4);
    if (i < i1 && r == 0) {
      if (o.residual) val += __ldcg(o.residual + i);
      o.out[i] = val;
    }
  }
  // the last exchange of this slot in the step advances the shared epoch counter (every CTA read it at kernel start)
  if (blockIdx.x == 0 && tid == 0 && o.tp_use == o.tp_uses - 1) o.tp_state[0] = epoch0 + (unsigned)o.tp_uses;
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int HS>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_step_kernel(const DsParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  unsigned char* after = smem + (size_t)p.nstages * p.stage_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after);                    // full[nstages], empty[nstages] (<= 15 stages)
  volatile int* done = reinterpret_cast<volatile int*>(after + 240);
  float* red = reinterpret_cast<float*>(after + 256);
  float* colscale = red + DS_RED_FLOATS;
  float* xsum = colscale + 8;
  unsigned char* xs = reinterpret_cast<unsigned char*>(xsum + p.xsum_floats);  // activation columns; attention scratch
  __shared__ __align__(16) float s_stat[4 * GS_CWARPS];
  __shared__ float s_xstat[2];
  __shared__ __align__(128) unsigned char s_ops[DS_NREC * sizeof(DsOp)];  // op records: previous / current / next (see fetch_op)
  __shared__ int s_dep_ready;  // highest op index whose completion on ALL CTAs the watcher thread has observed
  __shared__ int s_hs_ready;   // highest slab op whose head (the P CTAs that split its sequence) the watcher has seen arrive
  __shared__ __align__(16) float s_loc[DS_SLAB_MAXU * 8];  // CTA-local input columns of a slab op (SwiGLU outputs / head output)
  __shared__ DsSlabMeta s_meta[2];
  __shared__ __align__(128) uint32_t s_nf4[32];  // two copies of {bf16 hi term | bf16 lo term << 16} of the 16 NF4 values
  float* xraw = reinterpret_cast<float*>(xs + p.xs_bytes);  // raw activation row kept for a reuse_x op (may be empty)
  const uint32_t red_u32 = gs_smem_u32(red), xs_u32 = gs_smem_u32(xs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DsRing rg;
  rg.bar0 = gs_smem_u32(bars);
  rg.ring_u32 = gs_smem_u32(ring);
  rg.ring = ring;
  rg.nstages = p.nstages;
  rg.stage_stride = p.stage_stride;
  rg.s = 0;
  rg.ph = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(rg.bar0 + 8 * s, 1);
      mbar_init(rg.bar0 + 8 * (p.nstages + s), GS_CWARPS);
    }
    done[0] = done[1] = 0;
    s_dep_ready = -1;
    for (int k = 0; k < 32; ++k) {
      const float v = c_ds_nf4[k & 15];
      const uint32_t hi = gs_bf16_bits(v), lo = gs_bf16_bits(v - __uint_as_float(hi << 16));
      s_nf4[k] = hi | (lo << 16);
    }
    s_hs_ready = -1;
    for (int k = 0; k < 2; ++k) {
      if (p.slab_meta[k]) s_meta[k] = p.slab_meta[k][blockIdx.x];
      else s_meta[k].nrb = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == GS_CWARPS + 1 + DS_EWARPS) {
    if (lane != 0) return;
    {
      // =========================== WATCHER: polls the grid-wide dependencies on behalf of the consumers ==================
      // One thread of a warp of its own (sharing the producer's warp starves both: divergent lanes of a warp do not overlap their
      // stalls) walks the dependency list ahead of the consumers and polls the arrival
      // counter of each new dependency (relaxed polls, one acquire), then publishes it in shared memory: the 512 consumer
      // threads only ever spin on a shared-memory word, nobody leaves the critical path to poll L2, and the poll is already in
      // flight while the consumers finish the previous op.  This is also where the watchdog lives (ds_report).
      pdl_wait();  // the counters are zeroed by the prologue kernel of this step
      int waited = -1;
      bool dead = false;
      for (int op = 0; op < p.nops; ++op) {
        int dep = __ldg(p.deps + op);
        const bool hs = dep <= -2;  // a slab fed by attention waits for the P CTAs of its head (counter -2 - dep + head)
        if (hs && (int)blockIdx.x >= p.H * p.P) continue;  // no head, no slab
        if (!hs && dep <= waited) continue;
        if (!dead) {
          const unsigned* ctr = p.counters + (hs ? -2 - dep + (int)blockIdx.x / p.P : dep);
          const unsigned target = hs ? (unsigned)p.P : gridDim.x;
          unsigned it = 0, seen;
          const unsigned long long t0 = gs_now();
          while ((seen = ds_ld_relaxed(ctr)) < target) {
            if (p.timeout_ns) {  // a timer read per poll is noise next to the poll's own L2 round trip
              const bool expired = gs_now() - t0 > p.timeout_ns;
              if (expired || ((++it & 15u) == 0 && ds_ld_relaxed(p.err) != 0u)) {
                if (expired) ds_report(p.err, DS_ERR_DEP_TIMEOUT, op, dep, seen);
                dead = true;
                break;
              }
            }
          }
          (void)ds_ld_acquire(ctr);
        }
        if (hs) {
          ds_sts_release(&s_hs_ready, op);
        } else {
          waited = dep;
          ds_sts_release(&s_dep_ready, dep);
        }
      }
      return;
    }
  }
  if (warp > GS_CWARPS) {
    pdl_wait();
    ds_epilogue_warp(p, warp - GS_CWARPS - 1, red_u32, colscale, done, reinterpret_cast<const DsOp*>(s_ops), s_loc);
    return;
  }
  if (warp == GS_CWARPS) {
    // =========================== PRODUCER: walks the op table, waits only for free ring slots =========================
    if (lane != 0) return;
    bool have_geo = false;
    DsAttnGeo<HS> geo;
    // L2 look-ahead, used ONLY while this thread is blocked on a full ring.  The producer is always a ring depth ahead of the
    // consumers, so whenever they stop consuming (activation staging, dependency waits, attention start-up, epilogue tails:
    // ~6 us around every op) the ring is already full, this thread cannot issue, and HBM idles for the whole pause whatever the
    // ring depth.  While blocked it therefore walks a second cursor over the coming weight stages (same op order, linear ops
    // only) and prefetches them into L2 (`cp.async.bulk.prefetch.tensor`), at most `l2_ahead` stages beyond the TMA cursor: HBM
    // keeps working during the pause, and the ring then refills from L2.  (Prefetching at a fixed distance all the time — the
    // first version of this knob — only adds requests while HBM is saturated and was slower.)
    int pf_op = -1, pf_s = 0, pf_e = 0;
    int n_issued = 0, n_pf = 0;  // linear stages handed to TMA / passed by the prefetch cursor (n_pf >= n_issued)
    auto pf_step = [&](bool fetch) {
      while (pf_op < p.nops) {
        if (pf_s < pf_e) {
          if (fetch) {
            const DsOp& q = p.ops[pf_op];
            const int tile = pf_s / q.nks, ks = pf_s - tile * q.nks;
            ds_prefetch_stage_l2(&q.map, 0, tile * GS_ROWS, ks * GS_KB);
            if (q.fmt == LP_W_INT4) {
              const int nch = (q.K + 127) / 128;
              const int c_begin = ks * GS_KB * 2, c_end = min(nch, (ks + 1) * GS_KB * 2);
              const int g_begin = c_begin / q.gp128;
              const uint32_t len = (uint32_t)((c_end + q.gp128 - 1) / q.gp128 - g_begin) * 16 * q.aux_bytes;
              ds_prefetch_l2(reinterpret_cast<const char*>(q.aux2) + ((size_t)tile * q.ngroups + g_begin) * 16 * q.aux_bytes, len);
            }
          }
          ++pf_s;
          ++n_pf;
          return;
        }
        do ++pf_op; while (pf_op < p.nops && p.ops[pf_op].kind != DS_KIND_LINEAR);
        if (pf_op < p.nops) ds_stage_range(p.ops[pf_op], pf_s, pf_e);
      }
    };
    const int l2_ahead = p.l2_ahead;
    auto wait_empty = [&]() {
      const uint32_t bar = rg.empty(), par = (uint32_t)(rg.ph ^ 1);
      if (l2_ahead > 0) {
        while (!mbar_test(bar, par)) {
          if (n_pf - n_issued < l2_ahead && pf_op < p.nops) pf_step(true);
        }
      } else {
        mbar_wait(bar, par);
      }
    };
    for (int op = 0; op < p.nops; ++op) {
      const DsOp& o = p.ops[op];
      if (o.kind == DS_KIND_LINEAR) {
        const int nks = o.nks, fmt = o.fmt, gp128 = o.gp128, aux_bytes = o.aux_bytes, ngroups = o.ngroups;
        const int nch128 = (o.K + 127) / 128;
        int sb, se;
        ds_stage_range(o, sb, se);
        const char* aux2 = reinterpret_cast<const char*>(o.aux2);
        // the norm parameters of this op are read once per step (HBM misses of ~2 us on the consumers' critical path
        // otherwise): one CTA pulls them into L2 now — the producer runs a ring depth ahead of the consumers
        if (o.norm_kind >= 0 && (int)blockIdx.x == op % (int)gridDim.x) {
          ds_prefetch_l2(o.nw, (uint32_t)o.K * 4);
          if (o.nb) ds_prefetch_l2(o.nb, (uint32_t)o.K * 4);
        }
        int tile = sb / nks, ks = sb % nks;
        for (int s0 = sb; s0 < se; ++s0) {
          wait_empty();
          if (l2_ahead > 0) {
            if (n_pf == n_issued) pf_step(false);  // the cursor never falls behind the TMA cursor
            ++n_issued;
          }
          const uint32_t dst = rg.ring_u32 + (uint32_t)rg.s * rg.stage_stride;
          uint32_t aux_len = 0;
          int g_begin = 0;
          if (fmt == LP_W_INT4) {
            const int c_begin = ks * GS_KB * 2, c_end = min(nch128, (ks + 1) * GS_KB * 2);
            g_begin = c_begin / gp128;
            aux_len = (uint32_t)((c_end + gp128 - 1) / gp128 - g_begin) * 16 * aux_bytes;
          }
          if (fmt == LP_W_NF4) aux_len = 2048;  // absmax of the stage: [32 blocks of 64 weights][16 rows] floats, tile-major
          mbar_expect_tx(rg.full(), GS_KB * GS_BLK_BYTES + aux_len);
          tma_load_3d(dst, &o.map, 0, tile * GS_ROWS, ks * GS_KB, rg.full());
          if (fmt == LP_W_INT4)
            bulk_g2s(dst + GS_KB * GS_BLK_BYTES, aux2 + ((size_t)tile * ngroups + g_begin) * 16 * aux_bytes, aux_len, rg.full());
          else if (fmt == LP_W_NF4)
            bulk_g2s(dst + GS_KB * GS_BLK_BYTES, aux2 + ((size_t)tile * nks + ks) * 2048, 2048, rg.full());
          rg.advance();
          if (++ks == nks) {
            ks = 0;
            ++tile;
          }
        }
      } else if (o.kind == DS_KIND_SLAB) {
        const DsSlabMeta& mt = s_meta[o.slab_src];
        const int nrb = mt.nrb;
        if (nrb > 0) {
          const unsigned char* img = o.slab_img + mt.off;
          const int stagger = (p.dbg & 2) ? 0 : (2 * (o.slab_src == 1 ? (int)blockIdx.x / p.P : (int)blockIdx.x)) % nrb;
          for (int i = 0; i < nrb; ++i) {
            int rb = i + stagger;
            if (rb >= nrb) rb -= nrb;
            wait_empty();
            mbar_expect_tx(rg.full(), (uint32_t)mt.stage_bytes);
            bulk_g2s(rg.ring_u32 + (uint32_t)rg.s * rg.stage_stride, img + (size_t)rb * mt.stage_bytes, (uint32_t)mt.stage_bytes, rg.full());
            rg.advance();
          }
        }
      } else if (o.kind != DS_KIND_EXCHANGE) {
        if (!have_geo) {
          pdl_wait();  // the position is written by the previous step's sampler
          geo.init(p);
          have_geo = true;
          // this position's RoPE rows are read by every attention op of the step and are cold in the first one: pull them into L2
          if (blockIdx.x == 0 && p.n_elem > 0 && (p.n_elem & 3) == 0) {
            ds_prefetch_l2(p.cosT + (size_t)geo.pos * p.n_elem, (uint32_t)p.n_elem * 4);
            ds_prefetch_l2(p.sinT + (size_t)geo.pos * p.n_elem, (uint32_t)p.n_elem * 4);
          }
        }
        constexpr int AT = DsAttnGeo<HS>::AT;
        const int npairs = p.H * p.P;
        for (int j = blockIdx.x; j < npairs; j += gridDim.x) {
          const int h = j / p.P, sp = j % p.P;
          const int b0 = sp * geo.bpp, b1 = min(geo.nblk, b0 + geo.bpp);
          for (int blk = b0; blk < b1; ++blk) {
            const int rows = min(AT, p.max_seq - blk * AT);
            const uint32_t bytes = (uint32_t)rows * HS * 2;
            const size_t off = ((size_t)(h / (p.H / p.G)) * p.max_seq + (size_t)blk * AT) * HS;
            wait_empty();
            mbar_expect_tx(rg.full(), bytes);
            bulk_g2s(rg.ring_u32 + (uint32_t)rg.s * rg.stage_stride, o.kc + off, bytes, rg.full());
            rg.advance();
            wait_empty();
            mbar_expect_tx(rg.full(), bytes);
            bulk_g2s(rg.ring_u32 + (uint32_t)rg.s * rg.stage_stride, o.vc + off, bytes, rg.full());
            rg.advance();
          }
        }
      }
    }
    return;
  }

  // =============================== CONSUMERS ==========================================================================
  pdl_wait();  // embedding row, zeroed counters, position
  pdl_launch_dependents();
  DsAttnGeo<HS> geo;
  geo.init(p);
  // tensor parallel: epoch counters of the two exchange slots as of the start of this step
  unsigned int tp_epoch[2] = {0u, 0u};
  if (p.tp_state0) tp_epoch[0] = *reinterpret_cast<volatile unsigned int*>(p.tp_state0);
  if (p.tp_state1) tp_epoch[1] = *reinterpret_cast<volatile unsigned int*>(p.tp_state1);
  int waited = -1, gt = 0;
  bool have_xraw = false;  // the previous op left the raw activation row + statistics in shared memory (save_x)
  // Op records live in SHARED memory (an op would otherwise begin with an L2 round trip, ~0.7 us under streaming load, for its
  // own 384-byte record): 24 lanes of warp 1 copy the NEXT op's record with cp.async while the current op runs; the op-end
  // barrier publishes it to the consumers and the epilogue warps.  Three slots: record k+1 replaces record k-2, which nobody can
  // still be reading (everybody has passed the op-end barriers of k-2 and k-1).  The producer and the watcher read the table in
  // global memory: they run ahead of the critical path, and the TMA descriptor must stay in global memory.
  const uint32_t s_ops_u32 = gs_smem_u32(s_ops);
  auto fetch_op = [&](int op) {
    if (warp == 1 && lane < (int)(sizeof(DsOp) / 16)) {
      const char* src = reinterpret_cast<const char*>(p.ops + op) + lane * 16;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s_ops_u32 + (uint32_t)((op % DS_NREC) * sizeof(DsOp) + lane * 16)), "l"(src) : "memory");
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
  };
  auto fetch_wait = [&]() {
    if (warp == 1) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  };
  if (p.nops > 0) fetch_op(0);
  fetch_wait();
  asm volatile("bar.sync 4, %0;\n" ::"n"(DS_OPEND_THREADS) : "memory");
  for (int op = 0; op < p.nops; ++op) {
    const DsOp& o = reinterpret_cast<const DsOp*>(s_ops)[op % DS_NREC];
    if (op + 1 < p.nops) fetch_op(op + 1);
    unsigned long long* tr = p.trace ? p.trace + ((size_t)op * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (tr && threadIdx.x == 0) tr[0] = gs_now();
    const int dep = o.dep;
    // counters are cumulative: a CTA arrives for op d only after all of its earlier ops
    auto wait_dep = [&]() {
      if (dep > waited) {
        while (ds_lds_acquire(&s_dep_ready) < dep) {}
        waited = dep;
      }
      if (tr && threadIdx.x == 0) tr[1] = gs_now();
    };
    if (o.kind == DS_KIND_LINEAR) {
      ds_linear<HS>(p, o, geo, rg, red_u32, colscale, xsum, xs_u32, done, s_stat, tr, gt, wait_dep, xraw, s_xstat, have_xraw,
                    gs_smem_u32(s_nf4));
    } else if (o.kind == DS_KIND_SLAB) {
      auto wait_hs = [&]() {
        while (ds_lds_acquire(&s_hs_ready) < op) {}
      };
      have_xraw = false;
      ds_slab<HS>(p, o, s_meta[o.slab_src], geo, rg, xs, s_loc, wait_hs, tr);
    } else if (o.kind == DS_KIND_EXCHANGE) {
      wait_dep();
      ds_exchange(p, o, op, o.tp_state == p.tp_state1 ? tp_epoch[1] : tp_epoch[0]);
    } else {
      wait_dep();
      ds_attention<HS>(p, o, geo, rg, xs, tr);
    }
    fetch_wait();  // the next op's record has landed (issued at the start of this op)
    // the epilogue warps have written this op's rows; one of their lanes signals the op after this barrier
    asm volatile("bar.sync 4, %0;\n" ::"n"(DS_OPEND_THREADS) : "memory");
    if (tr && threadIdx.x == 0) tr[3] = gs_now();
  }
}

// Load time: slab images of a projection whose input columns stay inside the CTA that produced them (ds_slab).  One CTA of this
// kernel writes one step-kernel CTA's image: per 256-row stage [unit][row pair ^ 4 (unit & 3)][2 rows] words of 8 int4 columns,
// then the packed scale / zero words of the one or two scale groups the CTA's columns touch.
__global__ void slab_build_kernel(const uint32_t* __restrict__ rows32, int row_words, const uint32_t* __restrict__ aux2, int ngroups,
                                  const DsSlabMeta* __restrict__ meta, uint32_t* __restrict__ image) {
  const DsSlabMeta mt = meta[blockIdx.x];
  const int stage_words = mt.stage_bytes / 4;
  uint32_t* img = image + mt.off / 4;
  for (int w = threadIdx.x; w < mt.nrb * stage_words; w += blockDim.x) {
    const int rb = w / stage_words, r = w - rb * stage_words;
    uint32_t v;
    if (r < mt.nunits * DS_SLAB_ROWS) {
      const int unit = r / DS_SLAB_ROWS, pos = r % DS_SLAB_ROWS;
      const int pr = (pos >> 1) ^ (4 * (unit & 3));
      const int row = mt.row0 + rb * DS_SLAB_ROWS + 2 * pr + (pos & 1);
      v = rows32[(size_t)row * row_words + mt.unit0 + unit];
    } else {
      const int q = r - mt.nunits * DS_SLAB_ROWS, sg = q / DS_SLAB_ROWS;
      const int row = mt.row0 + rb * DS_SLAB_ROWS + q % DS_SLAB_ROWS;
      v = aux2[((size_t)(row >> 4) * ngroups + mt.group_a + sg) * 16 + (row & 15)];  // tile-major [N/16][groups][16 rows]
    }
    img[w] = v;
  }
}

// Prologue of a step: x = wte[token] (model.py:99), arrival counters = 0.
__global__ void decode_step_prep_kernel(const void* __restrict__ idx, int idx64, const int* __restrict__ idx_offset,
                                        const void* __restrict__ wte, int wte_dtype, float* __restrict__ x, int E,
                                        unsigned* __restrict__ counters, int nops, float4* __restrict__ zero, size_t zero_n4) {
  pdl_launch_dependents();  // the step kernel prefetches weights while the previous step's sampler is still running
  pdl_wait();
  // outputs of ops that accumulate in place into a buffer of their own (QKV rows split stream-K): cleared here, the step kernel's
  // consumers pass their own griddepcontrol.wait only when this grid has completed
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n4; i += (size_t)gridDim.x * blockDim.x)
    zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int at = idx_offset ? idx_offset[0] : 0;
  const long long tok = idx64 ? reinterpret_cast<const long long*>(idx)[at] : (long long)reinterpret_cast<const int*>(idx)[at];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x)
    x[e] = wte_dtype == LP_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(wte)[(size_t)tok * E + e])
                                : reinterpret_cast<const float*>(wte)[(size_t)tok * E + e];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nops; i += blockDim.x) counters[i] = 0u;
}

// ------------------------------------------------------------------------------------------------ host side
struct DsHostPlan {  // lp_step_handle, opaque to the caller
  uint32_t magic;
  int nops, nstages, stage_stride, xsum_floats, hs, H, G, P, n_elem, max_seq, E, wte_dtype, idx64, grid, i4pair, l2_ahead, xs_bytes, coop, ncounters;
  const DsSlabMeta* slab_meta[2];
  unsigned long long timeout_ns;
  const int* deps;
  unsigned* err;
  float scale_log2;
  size_t smem;
  const DsOp* ops_dev;
  unsigned* counters;
  const int* pos;
  const float* cosT;
  const float* sinT;
  float* part;
  const void* idx;
  const int* idx_offset;
  const void* wte;
  float* x0;
  unsigned int* tp_state0;
  unsigned int* tp_state1;
  void* zero_ptr;
  size_t zero_bytes;
};
static_assert(sizeof(DsHostPlan) <= sizeof(lp_step_handle), "lp_step_handle too small");
constexpr uint32_t DS_MAGIC = 0x4c504453u;

static unsigned long long* g_ds_trace = nullptr;

constexpr size_t DS_SMEM_OPTIN = 224 * 1024;  // 227 KB per CTA minus the static block (statistics, op records, slab vector: < 3 KB)

template <int HS>
static int ds_prepare(int device) {  // per DEVICE: the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  static bool attr_set[64] = {};
  if (device < 0 || device >= 64) return LP_ERR_INVALID_ARG;
  if (!attr_set[device]) {
    LP_CUDA_TRY(cudaFuncSetAttribute(decode_step_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DS_SMEM_OPTIN));
    attr_set[device] = true;
  }
  return LP_OK;
}

// The step kernel spins on counters that OTHER CTAs of the same grid advance, so all CTAs must be resident at the same time.
// (a) lp_decode_step_plan refuses a grid that the occupancy calculator says cannot be co-resident (1 CTA of 608 threads and
// ~220 KB of shared memory per SM: grid <= #SMs); (b) the launch carries cudaLaunchAttributeCooperative, which makes the driver
// guarantee co-residency (or fail the launch) whatever else is running on the device; (c) the watcher's waits are bounded
// (ds_report).  LP_DS_COOP=0 drops (b) (A/B measurements only).
template <int HS>
static int ds_launch(const DsParams& p, const DsHostPlan& h, void* stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h.grid);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = h.smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (h.coop) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  LP_CUDA_TRY(cudaLaunchKernelEx(&cfg, decode_step_kernel<HS>, p));
  count_launch();
  return LP_OK;
}

// Can the device run this grid cooperatively together with the PDL attribute?  Probed once per device at plan time with an empty
// op table (the kernel starts, finds nothing to do and exits).
template <int HS>
static int ds_probe_coop(const DsHostPlan& h, int device) {
  static int mode[64];  // 0: unknown, 1: cooperative launch works, 2: it does not (plain launch + occupancy check + watchdog)
  if (const char* e = getenv("LP_DS_COOP"))
    if (e[0] == '0') return 0;
  if (mode[device] == 0) {
    DsParams p;
    memset(&p, 0, sizeof(p));
    p.pos = h.pos;
    p.nstages = h.nstages;
    p.stage_stride = h.stage_stride;
    p.P = 1;
    p.H = p.G = 1;
    p.max_seq = 1;
    DsHostPlan hp = h;
    hp.coop = 1;
    int rc = ds_launch<HS>(p, hp, nullptr);
    cudaError_t e = rc == LP_OK ? cudaStreamSynchronize(nullptr) : cudaErrorUnknown;
    (void)cudaGetLastError();
    mode[device] = (rc == LP_OK && e == cudaSuccess) ? 1 : 2;
  }
  return mode[device] == 1;
}

}  // namespace lp

extern "C" {

/* op records | arrival counters [n + H per fused attention op] | dependency list [n] | error record [8] */
size_t lp_decode_step_plan_bytes(int n_ops) { return n_ops > 0 ? (size_t)n_ops * (sizeof(lp::DsOp) + 8 + 160) + 256 : 0; }

size_t lp_decode_step_workspace_bytes(int H, int hs) { return (size_t)H * lp::DS_MAXP * (hs + 4) * 4; }

int lp_debug_step_trace(void* device_buf) {
  lp::g_ds_trace = reinterpret_cast<unsigned long long*>(device_buf);
  return LP_OK;
}

int lp_decode_step_plan(const lp_step_op* ops, int n_ops, const lp_step_geom* gm, void* plan_dev, size_t plan_bytes,
                        lp_step_handle* handle) {
  using namespace lp;
  if (!ops || n_ops <= 0 || !gm || !plan_dev || !handle) return LP_ERR_INVALID_ARG;
  if (plan_bytes < lp_decode_step_plan_bytes(n_ops) || (reinterpret_cast<uintptr_t>(plan_dev) & 127)) return LP_ERR_WORKSPACE;
  if (!gm->pos || !gm->idx || !gm->wte || !gm->x0 || !gm->workspace || gm->E <= 0) return LP_ERR_INVALID_ARG;
  if (gm->hs != 64 && gm->hs != 128) return LP_ERR_UNSUPPORTED;
  if (gm->H <= 0 || gm->G <= 0 || gm->H % gm->G) return LP_ERR_INVALID_ARG;
  if (gm->kv_dtype != LP_BF16 || gm->max_seq <= 0 || gm->n_elem < 0 || gm->n_elem > gm->hs || (gm->n_elem & 1)) return LP_ERR_UNSUPPORTED;
  if (gm->n_elem > 0 && (!gm->cos || !gm->sin)) return LP_ERR_INVALID_ARG;
  if (gm->workspace_bytes < lp_decode_step_workspace_bytes(gm->H, gm->hs)) return LP_ERR_WORKSPACE;
  const int grid = num_sms();
  if (gm->H > grid) return LP_ERR_UNSUPPORTED;

  std::vector<DsOp> dev(n_ops);
  const DsSlabMeta* slab_meta[2] = {nullptr, nullptr};
  int n_hsync = 0;
  unsigned int* tp_state[2] = {nullptr, nullptr};
  int tp_uses[2] = {0, 0};
  int stage_stride = GS_KB * GS_BLK_BYTES, xsum_floats = 0;
  // attention scratch inside the activation area: q, new k/v, scores, slice sums (ds_attention)
  size_t xs_bytes = (size_t)gm->hs * 4 + 2 * gm->hs * 2 + (size_t)GS_CWARPS * (gm->hs + 4) * 4 + 16;
  for (int i = 0; i < n_ops; ++i) {
    const lp_step_op& s = ops[i];
    DsOp& d = dev[i];
    memset(&d, 0, sizeof(d));
    d.kind = s.kind;
    d.dep = s.dep;
    d.signal = 0;
    d.hsync = -1;
    if (s.dep >= i || s.dep < -1) return LP_ERR_INVALID_ARG;
    if (s.dep >= 0) dev[s.dep].signal = 1;
    if (s.kind == LP_STEP_SLAB) {
      // out += W[:, this CTA's columns] . (what this CTA produced in the op before): see ds_slab
      if (!s.W || !s.out || !s.slab_image || !s.slab_meta || i == 0 || s.slab_src < 0 || s.slab_src > 1) return LP_ERR_INVALID_ARG;
      const lp_weight& W = *s.W;
      if (W.fmt != LP_W_INT4 || W.group != 128 || !(W.flags & LP_WF_AUX_PACKED) || W.N % DS_SLAB_ROWS) return LP_ERR_UNSUPPORTED;
      if (W.out_bias || W.out_scale) return LP_ERR_UNSUPPORTED;  // adapter-v2 models take the per-op path
      const lp_step_op& prev = ops[i - 1];
      if (s.slab_src == 0) {  // columns = SwiGLU outputs of the preceding up-projection, kept in shared memory
        if (prev.kind != LP_STEP_LINEAR || !prev.keep_local || prev.epilogue != LP_EPI_SWIGLU || !prev.W || prev.W->N != 2 * W.K ||
            s.dep != -1)
          return LP_ERR_INVALID_ARG;
      } else {  // columns = attention output of the CTA's head
        if (prev.kind != LP_STEP_ATTENTION || gm->hs != 128 || W.K != gm->H * gm->hs || s.dep != i - 1) return LP_ERR_INVALID_ARG;
        if (gm->H * std::max(1, std::min(DS_MAXP, grid / gm->H)) > grid) return LP_ERR_UNSUPPORTED;
        dev[i - 1].hsync = n_ops + n_hsync * gm->H;  // H per-head counters behind the per-op ones
        dev[i - 1].signal = 0;                       // nobody waits for the attention op grid-wide
        d.hsync = dev[i - 1].hsync;
        d.dep = -2 - d.hsync;                        // watcher: wait for the P CTAs of this CTA's head
        ++n_hsync;
      }
      if (slab_meta[s.slab_src] && slab_meta[s.slab_src] != s.slab_meta) return LP_ERR_UNSUPPORTED;  // one layout per source kind
      slab_meta[s.slab_src] = reinterpret_cast<const DsSlabMeta*>(s.slab_meta);
      d.kind = DS_KIND_SLAB;
      d.slab_img = reinterpret_cast<const unsigned char*>(s.slab_image);
      d.slab_src = s.slab_src;
      d.out = s.out;
      d.bias = W.bias;
      d.N = W.N;
      d.K = W.K;
      stage_stride = std::max(stage_stride, (DS_SLAB_MAXU + 1) * 1024);
      continue;
    }
    if (s.kind == LP_STEP_EXCHANGE) {
      if (!s.tp_buf_ptrs || !s.tp_pad_ptrs || !s.tp_state || !s.out || s.tp_size < 1 || s.tp_size > DS_MAX_TP || s.tp_rank < 0 ||
          s.tp_rank >= s.tp_size || s.dep < 0)
        return LP_ERR_INVALID_ARG;
      if (gm->E % 4 || s.tp_buf_offset % 16) return LP_ERR_UNSUPPORTED;
      if (i == 0 || ops[i - 1].kind != LP_STEP_LINEAR || ops[i - 1].tp_size != s.tp_size || s.dep != i - 1) return LP_ERR_INVALID_ARG;
      d.kind = DS_KIND_EXCHANGE;
      d.tp_bufs = reinterpret_cast<const unsigned long long*>(s.tp_buf_ptrs);
      d.tp_pads = reinterpret_cast<const unsigned long long*>(s.tp_pad_ptrs);
      d.tp_state = reinterpret_cast<unsigned int*>(s.tp_state);
      d.tp_buf_off = s.tp_buf_offset;
      d.tp_pad_base = s.tp_pad_base;
      d.tp_rank = s.tp_rank;
      d.tp_size = s.tp_size;
      d.residual = s.residual;
      d.out = s.out;
      d.N = gm->E;
      if (!tp_state[0] || tp_state[0] == d.tp_state) {
        tp_state[0] = d.tp_state;
        d.tp_use = tp_uses[0]++;
      } else if (!tp_state[1] || tp_state[1] == d.tp_state) {
        tp_state[1] = d.tp_state;
        d.tp_use = tp_uses[1]++;
      } else {
        return LP_ERR_UNSUPPORTED;  // two alternating slots
      }
      continue;
    }
    if (s.kind == LP_STEP_ATTENTION) {
      if (!s.qkv || !s.k_cache || !s.v_cache) return LP_ERR_INVALID_ARG;
      if ((reinterpret_cast<uintptr_t>(s.k_cache) | reinterpret_cast<uintptr_t>(s.v_cache) | reinterpret_cast<uintptr_t>(s.qkv)) & 15)
        return LP_ERR_UNSUPPORTED;
      d.qkv = s.qkv;
      d.kc = reinterpret_cast<__nv_bfloat16*>(s.k_cache);
      d.vc = reinterpret_cast<__nv_bfloat16*>(s.v_cache);
      continue;
    }
    if (s.kind != LP_STEP_LINEAR || !s.W || !s.out) return LP_ERR_INVALID_ARG;
    const lp_weight& W = *s.W;
    if (!W.w || W.N <= 0 || W.K <= 0) return LP_ERR_INVALID_ARG;
    if (W.fmt != LP_W_BF16 && W.fmt != LP_W_INT4 && W.fmt != LP_W_NF4 && W.fmt != LP_W_INT8) return LP_ERR_UNSUPPORTED;
    if (W.out_bias || W.out_scale) return LP_ERR_UNSUPPORTED;  // adapter-v2 output affine: the per-op kernels carry it
    if (W.N % GS_ROWS || W.K % 16 || W.K > 6 * DS_CTHREADS * 8 || (reinterpret_cast<uintptr_t>(W.w) & 15)) return LP_ERR_UNSUPPORTED;
    if (s.epilogue < LP_EPI_NONE || s.epilogue > LP_EPI_RESIDUAL) return LP_ERR_INVALID_ARG;
    if (s.epilogue == LP_EPI_RESIDUAL && !s.residual) return LP_ERR_INVALID_ARG;
    if (s.x_is_attention ? (W.K != gm->H * gm->hs) : !s.x) return LP_ERR_INVALID_ARG;
    if (s.x_is_attention && W.K > 2 * DS_CTHREADS * 8) return LP_ERR_UNSUPPORTED;
    if (s.norm_kind >= 0 && (s.norm_kind > LP_NORM_RMS || !s.norm_w)) return LP_ERR_INVALID_ARG;
    const int K = W.K, kpad = (K + 255) / 256 * 256;
    size_t row_bytes;
    int aux_stage = 0;
    d.fmt = W.fmt;
    if (W.fmt == LP_W_BF16) {
      if (K % 64) return LP_ERR_UNSUPPORTED;
      row_bytes = (size_t)K * 2;
      d.split = ((size_t)3 * K * 2 > 48 * 1024) ? 2 : 3;  // long rows: 16 mantissa bits per term pair are plenty
      d.ldx = kpad + 8;
      d.gp128 = 1;
      xs_bytes = std::max(xs_bytes, (size_t)d.split * d.ldx * 2);
    } else if (W.fmt == LP_W_NF4) {
      // bitsandbytes NF4: rows of K / 2 bytes, absmax per 64 weights TILE-MAJOR in aux2 ([N/16][K-stages][32 blocks][16 rows]
      // floats, LP_WF_AUX_TILED): one 2 KB bulk copy per stage behind the 16 KB of codes
      if (K % 256 || W.group != 64 || !W.aux2 || !(W.flags & LP_WF_AUX_TILED)) return LP_ERR_UNSUPPORTED;
      row_bytes = (size_t)K / 2;
      d.split = ((size_t)3 * K * 2 > 48 * 1024) ? 2 : 3;
      d.ldx = kpad + 8;
      d.gp128 = 1;
      aux_stage = 2048;
      xs_bytes = std::max(xs_bytes, (size_t)d.split * d.ldx * 2);
    } else if (W.fmt == LP_W_INT8) {
      // row-wise int8: rows of K bytes, row scales (SCB / 127) in aux0, applied by the tile epilogue
      if (K % 128 || !W.aux0) return LP_ERR_UNSUPPORTED;
      row_bytes = (size_t)K;
      d.split = 3;
      d.ldx = kpad + 16;
      d.gp128 = 1;
      xs_bytes = std::max(xs_bytes, (size_t)3 * d.ldx);
      xsum_floats = std::max(xsum_floats, (K + 127) / 128 * 8);
    } else {
      if (W.group <= 0 || W.group % 128 || !W.aux2) return LP_ERR_UNSUPPORTED;
      row_bytes = (size_t)kpad / 2;
      d.split = 3;
      d.ldx = kpad + 16;
      d.gp128 = W.group / 128;
      d.ngroups = (K + W.group - 1) / W.group;
      d.aux_bytes = (W.flags & LP_WF_AUX_PACKED) ? 4 : 8;
      // groups that can intersect one stage (16 chunks of 128 columns); group 128: exactly 16, no straddling
      const int groups_per_stage = d.gp128 == 1 ? GS_KB * 2 : (GS_KB * 2 + d.gp128 - 1) / d.gp128 + 1;
      aux_stage = groups_per_stage * 16 * d.aux_bytes;
      xs_bytes = std::max(xs_bytes, (size_t)3 * d.ldx);
      xsum_floats = std::max(xsum_floats, (K + 127) / 128 * 8);
    }
    if (s.epilogue == LP_EPI_RESIDUAL) {
      // who wrote the residual?  covered by `dep`, a step input, or the same CTA (equal N -> equal tile ranges)
      for (int r = i - 1; r >= 0; --r) {
        if (ops[r].kind == LP_STEP_LINEAR && ops[r].out == s.residual) {
          const bool inplace = s.residual == s.out, r_inplace = ops[r].epilogue == LP_EPI_RESIDUAL && ops[r].residual == ops[r].out;
          const bool same_rows = !inplace && ops[r].epilogue != LP_EPI_SWIGLU && ops[r].W->N == W.N;
          // two in-place residual ops may overlap: both accumulate atomically
          if (r > s.dep && !same_rows && !(inplace && r_inplace)) return LP_ERR_INVALID_ARG;
          break;
        }
      }
    }
    const GsMap& gmap = gs_tensor_map(W, row_bytes);
    if (gmap.rank != 3) return LP_ERR_UNSUPPORTED;
    d.map = gmap.map;
    d.x = s.x;
    d.residual = s.residual;
    d.out = s.out;
    d.bias = W.bias;
    d.nw = s.norm_kind >= 0 ? s.norm_w : nullptr;
    d.nb = s.norm_kind >= 0 ? s.norm_b : nullptr;
    d.aux2 = W.fmt == LP_W_INT8 ? reinterpret_cast<const void*>(W.aux0) : W.aux2;  // INT8: the row scales ride in aux2
    d.eps = s.eps;
    d.norm_kind = s.norm_kind >= 0 ? s.norm_kind : -1;
    d.epi = s.epilogue;
    d.N = W.N;
    d.K = K;
    d.nkb = (int)((row_bytes + 127) / 128);
    d.nks = (d.nkb + GS_KB - 1) / GS_KB;
    d.ntiles = W.N / GS_ROWS;
    d.x_attn = s.x_is_attention ? 1 : 0;
    d.keep_local = s.keep_local ? 1 : 0;
    if (s.tp_size > 0) {
      // row-parallel SENDER of a push exchange: the tile epilogue stores this rank's partial into slot [s][rank] of every rank's
      // symmetric buffer (tp_buf_offset addresses that slot) and the last CTA publishes the epoch (see ds_exchange)
      if (!s.tp_buf_ptrs || !s.tp_pad_ptrs || !s.tp_state || s.tp_size > DS_MAX_TP || s.tp_rank < 0 || s.tp_rank >= s.tp_size ||
          s.epilogue != LP_EPI_NONE || i + 1 >= n_ops || ops[i + 1].kind != LP_STEP_EXCHANGE || ops[i + 1].tp_state != s.tp_state)
        return LP_ERR_INVALID_ARG;
      if (s.tp_buf_offset % 16) return LP_ERR_UNSUPPORTED;
      d.tp_bufs = reinterpret_cast<const unsigned long long*>(s.tp_buf_ptrs);
      d.tp_pads = reinterpret_cast<const unsigned long long*>(s.tp_pad_ptrs);
      d.tp_state = reinterpret_cast<unsigned int*>(s.tp_state);
      d.tp_buf_off = s.tp_buf_offset;
      d.tp_pad_base = s.tp_pad_base;
      d.tp_rank = s.tp_rank;
      d.tp_size = s.tp_size;
      // same use index as the exchange that follows (which takes and advances it)
      if (!tp_state[0] || tp_state[0] == d.tp_state) {
        tp_state[0] = d.tp_state;
        d.tp_use = tp_uses[0];
      } else if (!tp_state[1] || tp_state[1] == d.tp_state) {
        tp_state[1] = d.tp_state;
        d.tp_use = tp_uses[1];
      } else {
        return LP_ERR_UNSUPPORTED;
      }
    }
    if (d.keep_local && (s.epilogue != LP_EPI_SWIGLU || i + 1 >= n_ops || ops[i + 1].kind != LP_STEP_SLAB || ops[i + 1].slab_src != 0 ||
                         (W.N / GS_ROWS + grid - 1) / grid > DS_SLAB_MAXU))
      return LP_ERR_INVALID_ARG;
    // in-place residual (x += W . u): evenly split stages, atomic accumulation (see ds_stage_range)
    d.streamk = (s.epilogue == LP_EPI_RESIDUAL && s.residual == s.out) ? 1 : 0;
    stage_stride = std::max(stage_stride, (GS_KB * GS_BLK_BYTES + aux_stage + 1023) / 1024 * 1024);
  }
  for (int i = 0; i < n_ops; ++i)
    if (dev[i].kind == DS_KIND_EXCHANGE) dev[i].tp_uses = tp_uses[dev[i].tp_state == tp_state[1] ? 1 : 0];
  xs_bytes = (xs_bytes + 15) / 16 * 16;
  // back-to-back LayerNorm ops on the same row (parallel residual: QKV, then FC on norm_2 of the same x): the second stages from
  // the raw row + statistics the first leaves in shared memory (LP_DS_REUSEX=0 disables)
  size_t xraw_bytes = 0;
  {
    const char* e = getenv("LP_DS_REUSEX");
    const bool on = !(e && e[0] == '0');
    for (int i = 1; on && i < n_ops; ++i) {
      DsOp &a = dev[i - 1], &b = dev[i];
      if (a.kind != DS_KIND_LINEAR || b.kind != DS_KIND_LINEAR) continue;
      if (a.norm_kind != LP_NORM_LAYERNORM || b.norm_kind != LP_NORM_LAYERNORM || a.x_attn || b.x_attn) continue;
      if (a.x != b.x || a.K != b.K || a.eps != b.eps || a.out == a.x || a.K > 2 * DS_CTHREADS * 8 || a.reuse_x) continue;
      if (ops[i].dep != ops[i - 1].dep) continue;
      a.save_x = 1;
      b.reuse_x = 1;
      xraw_bytes = std::max(xraw_bytes, (size_t)a.K * 4);
    }
  }
  const size_t tail = 256 + (size_t)DS_RED_FLOATS * 4 + 32 + (size_t)xsum_floats * 4 + xs_bytes + xraw_bytes;
  const size_t budget = DS_SMEM_OPTIN;
  if (tail + 3 * (size_t)stage_stride + 1024 > budget) return LP_ERR_UNSUPPORTED;
  int ns = (int)((budget - tail - 1024) / stage_stride);
  if (ns > 14) ns = 14;
  if (const char* cap = getenv("LP_DS_STAGES")) {  // tuning aid: cap the ring depth
    const int c = atoi(cap);
    if (c >= 3 && c < ns) ns = c;
  }
  // paired int4 main loop (ds_linear_main_i4pair) needs an even stage count; without LP_DS_I4PAIR=1 the one-group-per-warp loop runs
  // (the same transformation of the bf16 loop was measured too: 1591 -> 1586 us on stablelm-3b, 3226 -> 3192 us on falcon-7b —
  // not worth a second summation order next to the per-op path, so bf16 keeps one loop)
  bool any_int4 = false;
  for (int i = 0; i < n_ops; ++i) any_int4 |= dev[i].kind == DS_KIND_LINEAR && dev[i].fmt == LP_W_INT4;
  const char* pair_env = getenv("LP_DS_I4PAIR");  // opt-in: "1" paired loop, "2" paired loop without arithmetic
  const bool i4pair = any_int4 && pair_env && pair_env[0] != '0' && ns >= 5;
  if (i4pair) ns &= ~1;

  DsHostPlan h;
  memset(&h, 0, sizeof(h));
  h.magic = DS_MAGIC;
  h.nops = n_ops;
  h.nstages = ns;
  {
    const char* e = getenv("LP_DS_L2AHEAD");
    h.l2_ahead = e ? atoi(e) : 0;
    if (h.l2_ahead < 0) h.l2_ahead = 0;
    // bound of every cross-CTA / cross-GPU wait of the kernel (watchdog, ds_report); LP_DS_TIMEOUT_MS=0: unbounded (tools that
    // slow the kernel down by orders of magnitude: compute-sanitizer)
    const char* to = getenv("LP_DS_TIMEOUT_MS");
    const long long ms = to ? atoll(to) : 4000;
    h.timeout_ns = ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
    if (const char* ns_env = getenv("LP_DS_TIMEOUT_NS")) h.timeout_ns = (unsigned long long)atoll(ns_env);  // test aid: force expiry
  }
  h.i4pair = i4pair ? (pair_env && pair_env[0] == '2' ? 2 : 1) : 0;  // 2: timing experiment, arithmetic skipped
  h.stage_stride = stage_stride;
  h.xsum_floats = xsum_floats;
  h.xs_bytes = (int)xs_bytes;
  h.hs = gm->hs;
  h.H = gm->H;
  h.G = gm->G;
  h.P = std::max(1, std::min(DS_MAXP, grid / gm->H));
  h.n_elem = gm->n_elem;
  h.max_seq = gm->max_seq;
  h.E = gm->E;
  h.wte_dtype = gm->wte_dtype;
  h.idx64 = gm->idx_is_int64;
  h.grid = grid;
  h.scale_log2 = gm->scale * 1.4426950408889634f;
  h.smem = (size_t)ns * stage_stride + tail + 1024;
  h.ops_dev = reinterpret_cast<const DsOp*>(plan_dev);
  h.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(plan_dev) + (size_t)n_ops * sizeof(DsOp));
  h.ncounters = n_ops + n_hsync * gm->H;
  h.deps = reinterpret_cast<const int*>(h.counters + h.ncounters);
  h.err = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(plan_dev) + ((size_t)n_ops * sizeof(DsOp) + (size_t)(h.ncounters + n_ops) * 4 + 15) / 16 * 16);
  if ((size_t)(reinterpret_cast<char*>(h.err) - reinterpret_cast<char*>(plan_dev)) + 32 > plan_bytes) return LP_ERR_WORKSPACE;
  h.slab_meta[0] = slab_meta[0];
  h.slab_meta[1] = slab_meta[1];
  h.pos = gm->pos;
  h.cosT = gm->cos;
  h.sinT = gm->sin;
  h.part = reinterpret_cast<float*>(gm->workspace);
  h.idx = gm->idx;
  h.idx_offset = gm->idx_offset;
  h.wte = gm->wte;
  h.x0 = gm->x0;
  h.zero_ptr = gm->zero_ptr;
  h.zero_bytes = gm->zero_ptr ? gm->zero_bytes : 0;
  if (h.zero_bytes % 16 || (reinterpret_cast<uintptr_t>(h.zero_ptr) & 15)) return LP_ERR_INVALID_ARG;
  h.tp_state0 = tp_state[0];
  h.tp_state1 = tp_state[1];
  // load-time copy of the op table, the dependency list and a clean error record (synchronous: the staging vectors die at return)
  LP_CUDA_TRY(cudaMemcpy(plan_dev, dev.data(), (size_t)n_ops * sizeof(DsOp), cudaMemcpyHostToDevice));
  {
    std::vector<int> deps(n_ops);
    for (int i = 0; i < n_ops; ++i) deps[i] = dev[i].dep;
    LP_CUDA_TRY(cudaMemcpy(const_cast<int*>(h.deps), deps.data(), (size_t)n_ops * sizeof(int), cudaMemcpyHostToDevice));
    LP_CUDA_TRY(cudaMemset(h.err, 0, 32));
  }
  // co-residency of the whole grid: occupancy check here, cooperative launch attribute where the device accepts it (ds_launch)
  {
    int device = 0, per_sm = 0;
    LP_CUDA_TRY(cudaGetDevice(&device));
    const int rc = gm->hs == 128 ? ds_prepare<128>(device) : ds_prepare<64>(device);
    if (rc != LP_OK) return rc;
    if (gm->hs == 128) LP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<128>, DS_THREADS, h.smem));
    else LP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<64>, DS_THREADS, h.smem));
    if (per_sm < 1 || h.grid > per_sm * num_sms()) return LP_ERR_UNSUPPORTED;
    h.coop = gm->hs == 128 ? ds_probe_coop<128>(h, device) : ds_probe_coop<64>(h, device);
  }
  memset(handle, 0, sizeof(*handle));
  memcpy(handle, &h, sizeof(h));
  return LP_OK;
}

int lp_decode_step(const lp_step_handle* handle, void* stream) {
  using namespace lp;
  if (!handle) return LP_ERR_INVALID_ARG;
  DsHostPlan h;
  memcpy(&h, handle, sizeof(h));
  if (h.magic != DS_MAGIC) return LP_ERR_INVALID_ARG;
  const size_t zero_n4 = h.zero_bytes / 16;
  const int prep_grid = std::max((h.E + 1023) / 1024, (int)std::min<size_t>((zero_n4 + 2047) / 2048, (size_t)num_sms()));
  int rc = launch(decode_step_prep_kernel, dim3(prep_grid), dim3(256), 0, stream, h.idx, h.idx64, h.idx_offset, h.wte, h.wte_dtype, h.x0,
                  h.E, h.counters, h.ncounters, reinterpret_cast<float4*>(h.zero_ptr), zero_n4);
  if (rc != LP_OK) return rc;
  DsParams p;
  p.ops = h.ops_dev;
  p.counters = h.counters;
  p.pos = h.pos;
  p.cosT = h.cosT;
  p.sinT = h.sinT;
  p.part = h.part;
  p.trace = g_ds_trace;
  p.scale_log2 = h.scale_log2;
  p.nops = h.nops;
  p.H = h.H;
  p.G = h.G;
  p.n_elem = h.n_elem;
  p.max_seq = h.max_seq;
  p.P = h.P;
  p.nstages = h.nstages;
  p.stage_stride = h.stage_stride;
  p.xsum_floats = h.xsum_floats;
  p.i4pair = h.i4pair;
  p.l2_ahead = h.l2_ahead;
  p.xs_bytes = h.xs_bytes;
  p.deps = h.deps;
  p.err = h.err;
  p.slab_meta[0] = h.slab_meta[0];
  p.slab_meta[1] = h.slab_meta[1];
  p.ncounters = h.ncounters;
  {
    static const char* dbg_env = getenv("LP_DS_DEBUG");
    p.dbg = dbg_env ? atoi(dbg_env) : 0;
  }
  p.timeout_ns = h.timeout_ns;
  p.tp_state0 = h.tp_state0;
  p.tp_state1 = h.tp_state1;
  return h.hs == 128 ? ds_launch<128>(p, h, stream) : ds_launch<64>(p, h, stream);
}

int lp_decode_step_slab_layout(int src, int N, int K, int fc_tiles_or_heads, int hs, lp_slab_meta* meta, int max_ctas, int* n_ctas,
                               size_t* image_bytes) {
  using namespace lp;
  if (!meta || !n_ctas || !image_bytes || N <= 0 || K <= 0 || src < 0 || src > 1 || fc_tiles_or_heads <= 0) return LP_ERR_INVALID_ARG;
  const int grid = num_sms();
  if (max_ctas < grid) return LP_ERR_WORKSPACE;
  if (N % DS_SLAB_ROWS || K % 128) return LP_ERR_UNSUPPORTED;
  long long off = 0;
  for (int c = 0; c < grid; ++c) {
    DsSlabMeta mt;
    memset(&mt, 0, sizeof(mt));
    if (src == 0) {
      // the CTA's tiles of the up-projection (ds_stage_range, whole tiles): one 16-row SwiGLU tile = 8 outputs = one unit
      const int tiles = fc_tiles_or_heads;
      if (tiles * 8 != K) return LP_ERR_INVALID_ARG;
      const int t0 = (int)((long long)tiles * c / grid), t1 = (int)((long long)tiles * (c + 1) / grid);
      mt.unit0 = t0;
      mt.nunits = t1 - t0;
      mt.row0 = 0;
      mt.nrb = mt.nunits > 0 ? N / DS_SLAB_ROWS : 0;
      if (mt.nrb & 1) return LP_ERR_UNSUPPORTED;  // the consumers take the stages in pairs
    } else {
      // CTA c = (head c / P, sequence split c % P): the head's hs columns, rows split over its P CTAs
      const int H = fc_tiles_or_heads;
      if (hs != 128 || H * hs != K || H > grid) return LP_ERR_UNSUPPORTED;
      const int P = std::max(1, std::min(DS_MAXP, grid / H));
      if ((N / P) % (2 * DS_SLAB_ROWS)) return LP_ERR_UNSUPPORTED;  // the consumers take the stages in pairs
      if (c < H * P) {
        mt.unit0 = (c / P) * (hs / 8);
        mt.nunits = hs / 8;
        mt.row0 = (c % P) * (N / P);
        mt.nrb = N / P / DS_SLAB_ROWS;
      }
    }
    if (mt.nunits > DS_SLAB_MAXU) return LP_ERR_UNSUPPORTED;
    if (mt.nrb > 0) {
      mt.group_a = mt.unit0 / 16;
      const int boundary = (mt.group_a + 1) * 16;  // first unit of the next 128-column scale group
      mt.units_a = std::min(mt.nunits, boundary - mt.unit0);
      mt.nseg = mt.units_a < mt.nunits ? 2 : 1;
      mt.stage_bytes = (mt.nunits + mt.nseg) * 1024;
      mt.off = off;
      off += (long long)mt.nrb * mt.stage_bytes;
    }
    memcpy(&meta[c], &mt, sizeof(mt));
  }
  *n_ctas = grid;
  *image_bytes = (size_t)off;
  return LP_OK;
}

int lp_decode_step_slab_build(const lp_weight* W, const lp_slab_meta* meta_dev, int n_ctas, void* image, void* stream) {
  using namespace lp;
  if (!W || !meta_dev || !image || n_ctas <= 0) return LP_ERR_INVALID_ARG;
  if (W->fmt != LP_W_INT4 || W->group != 128 || !(W->flags & LP_WF_AUX_PACKED) || !W->aux2 || !W->w) return LP_ERR_UNSUPPORTED;
  const int row_words = (int)(lp_int4_row_bytes(W->K) / 4);
  slab_build_kernel<<<n_ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint32_t*>(W->w), row_words, reinterpret_cast<const uint32_t*>(W->aux2), (W->K + 127) / 128,
      reinterpret_cast<const DsSlabMeta*>(meta_dev), reinterpret_cast<uint32_t*>(image));
  LP_CUDA_TRY(cudaGetLastError());
  return LP_OK;
}

int lp_decode_step_status(const lp_step_handle* handle, int32_t info[8]) {
  using namespace lp;
  if (!handle) return LP_ERR_INVALID_ARG;
  DsHostPlan h;
  memcpy(&h, handle, sizeof(h));
  if (h.magic != DS_MAGIC) return LP_ERR_INVALID_ARG;
  unsigned rec[8] = {};
  LP_CUDA_TRY(cudaMemcpy(rec, h.err, sizeof(rec), cudaMemcpyDeviceToHost));  // synchronises with the steps launched so far
  if (info) memcpy(info, rec, sizeof(rec));
  if (rec[0] == 0) return LP_OK;
  LP_CUDA_TRY(cudaMemset(h.err, 0, sizeof(rec)));  // reported once; the plan stays usable
  return LP_ERR_TIMEOUT;
}

int lp_decode_step_cooperative(const lp_step_handle* handle) {
  using namespace lp;
  if (!handle) return LP_ERR_INVALID_ARG;
  DsHostPlan h;
  memcpy(&h, handle, sizeof(h));
  return h.magic == DS_MAGIC ? h.coop : LP_ERR_INVALID_ARG;
}

}  // extern "C"
