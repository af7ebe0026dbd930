/*
 * lp_abi.h — C ABI of liblitparrot_b200.so: hand-written sm_100a CUDA kernels for the Lit-GPT
 * inference hot path (lit_gpt.model.GPT.forward with KV caches + generate/base.py::generate).
 *
 * The reference (griff4692/lit-parrot) is pure Python and has NO FFI: every device operation on
 * this path is an aten / library call issued from lit_gpt/model.py and generate/base.py.  Each entry
 * point below therefore cites the reference *Python* lines whose device work it replaces; the
 * Python-side binding a maintainer adds is the ctypes stub in INTEGRATION.md
 * (lit_parrot_b200/_lib.py is that stub, in full).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors on the Python side);
 *     the library never allocates, frees or synchronises, so every call is CUDA-graph capturable;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - activations are fp32 row-major [rows, dim]; `round_bf16 != 0` rounds results to bf16 precision
 *     at exactly the points where the reference's `bf16-true` mode rounds (values stay in fp32 storage);
 *   - return value: 0 (LP_OK) or a negative lp_status; lp_status_str() names it.  No exceptions,
 *     no global mutable state apart from the one-time attribute setup in lp_init().
 */
#ifndef LP_ABI_H
#define LP_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_ABI_VERSION 6

typedef enum {
  LP_OK = 0,
  LP_ERR_INVALID_ARG = -1,   /* bad enum / null pointer / non-positive size                 */
  LP_ERR_UNSUPPORTED = -2,   /* shape or alignment the kernels do not cover                 */
  LP_ERR_CUDA = -3,          /* a CUDA runtime call failed (see lp_last_cuda_error)          */
  LP_ERR_WORKSPACE = -4,     /* caller-provided workspace too small                          */
  LP_ERR_TIMEOUT = -5        /* a bounded wait inside lp_decode_step expired (lp_decode_step_status) */
} lp_status;

/* storage types of tensors that are not fp32 activations */
typedef enum { LP_F32 = 0, LP_BF16 = 1 } lp_dtype;

/* weight formats of a linear layer, W is logically [N, K] (out_features, in_features) */
typedef enum {
  LP_W_F32 = 0,   /* float  [N,K] row-major (nn.Linear.weight)                                          */
  LP_W_BF16 = 1,  /* bf16   [N,K] row-major (nn.Linear.weight)                                          */
  LP_W_INT4 = 2,  /* GPTQ int4, repacked by lp_repack_gptq_int4: [N, Kp/2] bytes row-major, byte j of a
                     row holds column 2j (low nibble) and 2j+1 (high nibble); Kp = K rounded up to 256 (128-byte rows);
                     scales/zeros fp32 [N, n_groups]; w = (q - zero) * scale   (quantize/gptq.py:243-252) */
  LP_W_NF4 = 3,   /* bitsandbytes NF4: bytes [N*K/2], first element in the HIGH nibble, absmax fp32 per
                     `group` (=blocksize 64) consecutive elements of the flattened weight               */
  LP_W_INT8 = 4   /* int8 [N,K] row-major, scales fp32 [N] = SCB/127 (row-wise absmax quantisation)      */
} lp_wfmt;

typedef enum {
  LP_EPI_NONE = 0,      /* out[M,N]   = y                                                               */
  LP_EPI_GELU = 1,      /* out[M,N]   = gelu_erf(y)                      (model.py:285-286)              */
  LP_EPI_SWIGLU = 2,    /* out[M,N/2] = silu(y[2i]) * y[2i+1]; W rows interleaved fc_1/fc_2 (model.py:298-300) */
  LP_EPI_RESIDUAL = 3   /* out[M,N]   = residual + y                     (model.py:171, 178-179)         */
} lp_epilogue;

typedef enum { LP_NORM_LAYERNORM = 0, LP_NORM_RMS = 1 } lp_norm_kind;

/* lp_weight.flags */
#define LP_WF_AUX_PACKED 1 /* aux2 holds one 32-bit word per (row, group): bf16 scale bits << 16 | bf16 zero-point bits;
                              otherwise a float2 {scale, zero}.  Packed is exact when the scales are bf16-representable
                              (a bf16 checkpoint) and halves the scale/zero traffic: 0.5 + 4/128 bytes per weight.   */

#define LP_WF_AUX_TILED 2  /* NF4: aux2 holds the absmax values TILE-MAJOR [N/16][ceil(K/2048) K-stages][32 blocks][16 rows] fp32 (zero
                              padded), the layout the step kernel fetches with one 2 KB bulk copy per stage                       */

/* One linear layer's weights.  `aux0`/`aux1`: INT4 -> scales/zeros; NF4 -> absmax/unused; INT8 -> row scales. */
typedef struct {
  const void* w;
  const float* aux0;
  const float* aux1;
  const void* aux2;    /* NF4, optional: tile-major absmax (LP_WF_AUX_TILED).  INT4, optional: scale/zero pairs TILE-MAJOR [N/16][n_groups][16 rows] (see flags), the layout
                          the streaming kernel fetches with one bulk copy per stage; NULL -> exact CUDA-core kernel   */
  const float* bias;   /* fp32 [N] or NULL                                                              */
  int32_t fmt;         /* lp_wfmt                                                                       */
  int32_t N, K;
  int32_t group;       /* INT4: columns per scale/zero group (K for per-row); NF4: blocksize            */
  int32_t flags;       /* LP_WF_*                                                                        */
  int32_t reserved;
  const float* out_bias;   /* adapter-v2 (lit_gpt/adapter_v2.py:34-35), fp32 [N] or NULL: y = out_scale * ((x.W^T + bias) + out_bias),  */
  const float* out_scale;  /* evaluated in that order BEFORE the epilogue's activation / residual (bf16 mode: rounded after each step) */
} lp_weight;

int lp_abi_version(void);
const char* lp_status_str(int status);
const char* lp_last_cuda_error(void);

/* Number of kernels this library has launched (or recorded into a graph under stream capture) since load. */
unsigned long long lp_launch_count(void);

/* Programmatic Dependent Launch on/off (default on; env LP_PDL=0 disables).  Debug aid. */
int lp_set_pdl(int enabled);

/* Which kernel family lp_linear uses: 0 = auto (streaming family where it applies, else FMA), 1 = FMA family only
 * (exact fp32 CUDA-core math, every format), 2 = streaming family only (persistent TMA-bulk ring + mma.sync;
 * LP_ERR_UNSUPPORTED where it does not apply).  Test aid. */
int lp_set_linear_path(int path);

/* Debug aid: when `device_buf` (>= 8 * 148 uint64) is non-NULL every streaming-GEMV CTA records globaltimer stamps
 * [start, after griddepcontrol.wait, x staged, first four stages consumed, end]; NULL switches it off. */
int lp_debug_stream_trace(void* device_buf);

/* One-time per-device setup (cudaFuncSetAttribute for large dynamic shared memory).  Idempotent, thread-safe. */
int lp_init(int device);

/* Range check of a step's inputs ON THE DEVICE, for callers whose ids / positions live in device memory they own (the replayed
 * decode graph): the reference raises on a token id outside [0, vocab) (nn.Embedding, model.py:99) and on a position outside the
 * RoPE table (index_select, model.py:88-92).  Out-of-range values are CLAMPED IN PLACE — the kernels that follow never index out
 * of bounds — and *flag |= 1 (token) | 2 (position); `flag` may be mapped pinned host memory, so the host sees it without a copy.
 * Callers that can read their inputs on the host (prefill) validate there and raise like the reference. */
int lp_validate_inputs(void* idx, int idx_is_int64, int n_idx, int vocab, int32_t* pos, int n_pos, int block_size, int32_t* flag,
                       void* stream);

/* The same check fused with staging a step's inputs for a replayed graph: idx_src (int32 / int64 [n_idx]) -> idx_dst int64, pos_src
 * (int32 / int64 [n_pos]) -> pos_dst int32, clamped and flagged like lp_validate_inputs.  One launch per decode step instead of two
 * device copies and a check (GPT.forward with T == 1, model.py:88-99). */
int lp_stage_inputs(const void* idx_src, int idx_is_int64, int n_idx, const void* pos_src, int pos_is_int64, int n_pos, int64_t* idx_dst,
                    int32_t* pos_dst, int vocab, int block_size, int32_t* flag, void* stream);

/* replaces nn.Embedding `self.transformer.wte(idx)` (model.py:99).  idx: int32 or int64 [rows]; if idx_offset is
 * non-NULL the rows are idx[*idx_offset + r] (device-side position, so a captured decode step can be replayed). */
int lp_embed(const void* idx, int idx_is_int64, const int32_t* idx_offset, const void* wte, int wte_dtype, float* out,
             int rows, int E, int round_bf16, void* stream);

/* replaces `config.norm_class` forward: torch.nn.LayerNorm / lit_gpt/rmsnorm.py:17-21 (model.py:109,167,169,179).
 * weight/bias fp32 [E] (bias NULL for RMSNorm).  round_bf16: LayerNorm output rounded; RMSNorm evaluated with
 * the reference's input-dtype roundings (x*x, mean, +eps, rsqrt, x*r, w*xn each rounded to bf16). */
int lp_norm(int kind, const float* x, const float* weight, const float* bias, float eps, float* y, int rows, int E,
            int round_bf16, void* stream);

/* replaces nn.Linear / ColBlockQuantizedLinear.forward (quantize/gptq.py:254-264) / bnb Linear4bit, Linear8bitLt
 * (quantize/bnb.py:18-75) at model.py:111,205,252,285-287,298-301, with the elementwise tail fused:
 *   y = x[M,K] . W[N,K]^T + bias ; out = epilogue(y [, residual]).
 * Weight-streaming kernel for small M (decode, M = batch); M > LP_LINEAR_MAX_M is processed in row chunks. */
int lp_linear(const float* x, int M, const lp_weight* W, int epilogue, const float* residual, float* out,
              int round_bf16, void* stream);

/* lp_norm fused into lp_linear as a prologue (each CTA normalises x while its first weight stages are in flight):
 * out = epilogue(norm(x) . W^T + bias).  Returns LP_ERR_UNSUPPORTED when the fusion is not available for this
 * shape / format / mode (bf16-faithful rounding, M > 8, formats other than bf16 and int4): callers then issue
 * lp_norm followed by lp_linear. */
int lp_norm_linear(int norm_kind, const float* norm_w, const float* norm_b, float eps, const float* x, int M,
                   const lp_weight* W, int epilogue, const float* residual, float* out, int round_bf16, void* stream);

/* ---- prefill (T > 1) projections on tcgen05 tensor cores ----------------------------------------------------------
 * lp_split_bf16: x fp32 [rows, K] -> `nterms` bf16 arrays [nterms][rows, K] with x = t0 + t1 (+ t2) exactly (1 term in
 * bf16-faithful mode), optionally through LayerNorm / RMSNorm first (norm_kind < 0: none).  These terms are the A operand
 * of lp_gemm_bf16_tc, which multiplies each with the same weight tile: fp32-activation accuracy on bf16 tensor cores.
 * lp_gemm_bf16_tc: out = epilogue(x . W^T + bias), W bf16 [N, K] row-major (nn.Linear.weight; other formats are first
 * expanded with lp_dequant_bf16), TMA-fed tcgen05.mma with the accumulator in tensor memory.  The result is written as
 * fp32 [M, Nout] (out_f32) and / or as `out_terms` bf16 split arrays [out_terms][M, Nout] (out_bf16) — the next GEMM's
 * operand.  Nout = N (N / 2 for LP_EPI_SWIGLU).  Requires N % 8 == 0 and K % 8 == 0, else LP_ERR_UNSUPPORTED.
 * replaces nn.Linear.forward on T > 1 tokens (model.py:111, 205, 252, 285-301). */
/* Prefill-sized problems on CTA pairs (tcgen05.mma.cta_group::2, one 256 x 256 tile per 2-CTA cluster: each SM stages its 128
 * rows of X and half of the W tile) instead of single CTAs.  On by default (env LP_GEMM_PAIR=0 turns it off at load). */
int lp_set_gemm_pair(int enabled);
int lp_split_bf16(const float* x, void* out_bf16, int rows, int K, int nterms, int norm_kind, const float* norm_w,
                  const float* norm_b, float eps, int round_bf16, void* stream);
int lp_gemm_bf16_tc(const void* x_terms, int nterms, int M, const void* w_bf16, int N, int K, const float* bias, int epilogue,
                    const float* residual, float* out_f32, void* out_bf16, int out_terms, int round_bf16, void* stream);
/* The same with the adapter-v2 output affine of lp_weight (out_bias / out_scale, fp32 [N] or NULL). */
int lp_gemm_bf16_tc_affine(const void* x_terms, int nterms, int M, const void* w_bf16, int N, int K, const float* bias,
                           const float* out_bias, const float* out_scale, int epilogue, const float* residual, float* out_f32,
                           void* out_bf16, int out_terms, int round_bf16, void* stream);
/* W in any lp_wfmt -> dense bf16 [N, K] (int4 / NF4 rounded to bf16 like the reference's bf16 dequantisation). */
int lp_dequant_bf16(const lp_weight* W, void* out_bf16, void* stream);

/* replaces the q/k regroup + apply_rope + torch.cat + cache index_copy_ of CausalSelfAttention.forward
 * (model.py:208-245, 330-336).  qkv [B*T, (H+2G)*hs] rows group-interleaved [q x q_per_kv, k, v] (model.py:210-214);
 * cos/sin fp32 [block_size, n_elem]; pos int32 [T] (shared by the batch, model.py:88-92).
 * q_out [B*T, H*hs] (head-major, rotated).  k/v cache [B, G, max_seq, hs] in kv_dtype, compact (one head per
 * query group); slot = pos % max_seq — a ring index replaces the reference's physical roll (model.py:238-242). */
int lp_rope_kv_append(const float* qkv, const float* cos, const float* sin, const int32_t* pos, float* q_out,
                      void* k_cache, void* v_cache, int kv_dtype, int B, int T, int H, int G, int hs, int n_elem,
                      int max_seq, int round_bf16, void* stream);

/* replaces `scaled_dot_product_attention(q, k_cache, v_cache, attn_mask)` (model.py:247, 256-275) for queries
 * against the KV cache: split-K over sequence blocks, warp-shuffle online softmax, all q heads of one KV group in
 * one CTA (MHA/GQA/MQA).  Query row r = b*T + t attends to min(pos[t]+1, max_seq) cache slots.
 * out [B*T, H*hs].  workspace: lp_attn_workspace_bytes(). */
size_t lp_attn_workspace_bytes(int B, int T, int H, int hs, int max_seq);
int lp_attn_decode(const float* q, const void* k_cache, const void* v_cache, int kv_dtype, const int32_t* pos,
                   float* out, void* workspace, size_t workspace_bytes, int B, int T, int H, int G, int hs,
                   int max_seq, float scale, int round_bf16, void* stream);

/* ONE launch for a single-token step (T == 1): lp_rope_kv_append + lp_attn_decode + the split merge, on the tensor cores
 * (mma.sync; q and P split into bf16 hi + lo terms, so accuracy is that of fp32 activations over the bf16 cache).
 * qkv [B, (H+2G)*hs] is the raw QKV projection; pos int32 [1]; out [B, H*hs].  The kernel writes the rotated k and v
 * of the new token into slot pos % max_seq of the caches.  Covers bf16 caches with hs 64 / 128; LP_ERR_UNSUPPORTED
 * otherwise (callers then issue the two calls above).  `workspace` (lp_attn_fused_workspace_bytes) must be
 * zero-filled before its first use; every launch leaves its ticket area zeroed again. */
size_t lp_attn_fused_workspace_bytes(int B, int H, int G, int hs, int max_seq);
int lp_attn_decode_fused(const float* qkv, const float* cos, const float* sin, const int32_t* pos, float* out, void* k_cache,
                         void* v_cache, int kv_dtype, void* workspace, size_t workspace_bytes, int B, int H, int G, int hs,
                         int n_elem, int max_seq, float scale, int round_bf16, void* stream);

/* Causal attention of T > 1 consecutive query positions (pos[0] .. pos[0] + T - 1, already appended to the cache by
 * lp_rope_kv_append) against the bf16 KV cache (model.py:247, 256-275 with the mask rows of model.py:91-92).  T >= 1024: tcgen05
 * kernel (csrc/attention_tc.cu: S = Q.K^T and O += P.V as tcgen05.mma with the accumulators in tensor memory, K / V tiles by TMA
 * straight from the cache, V as an MN-major operand); shorter prompts: FlashAttention-2 style mma.sync kernel.  q, out fp32 [B*T, H*hs].  hs 64 / 128, bf16 cache, pos[0] + T <= max_seq (the caller
 * guarantees the positions are consecutive and do not wrap); LP_ERR_UNSUPPORTED otherwise -> lp_attn_decode. */
/* Which kernel lp_attn_prefill uses: 0 = auto, 1 = mma.sync kernel only, 2 = tcgen05 kernel only.  Test aid. */
int lp_set_attn_prefill_path(int path);
int lp_attn_prefill(const float* q, const void* k_cache, const void* v_cache, int kv_dtype, const int32_t* pos, float* out, int B,
                    int T, int H, int G, int hs, int max_seq, float scale, int round_bf16, void* stream);

/* LLaMA-Adapter prefix attention (lit_gpt/adapter.py:234-254): for every query row r = b*T + t and head h
 *   out[r, h, :] += gating[h] * softmax(scale * q_rot[r, h] . ak[g(h)]^T) . av[g(h)]
 * over the aT adaption-prompt keys (no mask), added to the causal attention output `out` [B*T, H*hs] in place.  q is taken
 * from the raw QKV projection (layout of lp_rope_kv_append) and rotated here with the row's RoPE entries (pos int32 [T]); the prefix
 * keys / values ak, av fp32 [G, aT, hs] are the k / v parts of attn.attn(adapter_wte.weight), NOT rotated (adapter.py:238-249:
 * the reference caches them as adapter_kv_cache).  gating fp32 [H] (gating_factor (1, H, 1, 1)).  aT <= 64, hs <= 256. */
int lp_adapter_attn(const float* qkv, const float* cos, const float* sin, const int32_t* pos, const float* ak, const float* av,
                    const float* gating, float* out, int B, int T, int H, int G, int hs, int n_elem, int aT, float scale,
                    int round_bf16, void* stream);

/* load time: merged-LoRA weights (lit_gpt/lora.py:154-164, 338-361): W[N, K] += scaling * (B[N, r] . A[r, K]) restricted to the
 * rows listed in `rows` (int32 [n_rows], NULL = all N rows; LoRAQKVLinear.zero_pad scatters the update to lora_ind),
 * W in fp32 or bf16 (w_dtype: lp_dtype; the sum is rounded once to the stored dtype, like `weight.data += delta`).
 * B_rows fp32 [n_rows, r] row-major, A fp32 [r, K]. */
int lp_lora_merge(void* W, int w_dtype, int N, int K, const float* B_rows, const float* A, int r, const int32_t* rows, int n_rows,
                  float scaling, void* stream);

/* ---- GPTQ quantiser on the device (quantize/gptq.py:267-431, GPTQQuantizer) -------------------------------------------------------
 * lp_gptq_hessian_update — collect_input_stats (gptq.py:349-363): H[K, K] = keep * H + (x_scale * X)^T (x_scale * X), X fp32
 *   [rows, K] the layer's inputs of one calibration batch; callers pass keep = n / (n + b), x_scale = sqrt(2 / (n + b)).  The update
 *   is ONE tcgen05 GEMM with the tokens as the reduction dimension: the scaled X is transposed and split into `terms` (2 or 3)
 *   bf16 terms whose pairwise products (3 or 6, the rest is below 2^-17 / 2^-25 relative) are laid side by side along the
 *   reduction.  K % 8 == 0.  workspace: lp_gptq_hessian_workspace_bytes(rows, K, terms).
 * lp_gptq_find_params — find_params_weight (gptq.py:318-347), per-channel: for every row and every group gi in [group_first,
 *   group_first + n_groups) the min / max of W[row, gi * group : (gi + 1) * group] -> scales / zeros [N, ld] at column gi.
 * lp_gptq_block_sweep — the inner loop of quantize() (gptq.py:400-419) for the block of `count` <= 128 columns starting at i1:
 *   reads W[:, i1 : i1 + count] (not modified) and the upper Cholesky factor Hinv [K, K] of the inverse Hessian, writes the
 *   de-quantised values Q[:, i1 : i1 + count], the scaled errors Err [N, 128] and adds the block's loss per row to loss_rows [N].
 *   fp32 with the reference's operation order: same inputs -> bit-identical codes.
 * lp_gptq_trailing_update — W[:, i1 + count :] -= Err[:, :count] . Hinv[i1 : i1 + count, i1 + count :] (gptq.py:424).
 * The Cholesky factorisations in between (gptq.py:387-391) are cuSOLVER calls issued by the host side. */
size_t lp_gptq_hessian_workspace_bytes(int rows, int K, int terms);
int lp_gptq_hessian_update(float* H, int K, const float* x, int rows, float keep, float x_scale, int terms, void* workspace,
                           size_t workspace_bytes, void* stream);
int lp_gptq_find_params(const float* W, int N, int K, int group_first, int n_groups, int group, int maxq, int sym, float* scales,
                        float* zeros, int ld, void* stream);
int lp_gptq_block_sweep(const float* W, int N, int K, int i1, int count, const float* Hinv, const float* scales, const float* zeros,
                        int ld, int group, int maxq, float* Q, float* Err, float* loss_rows, void* stream);
int lp_gptq_trailing_update(float* W, int N, int K, int i1, int count, const float* Hinv, const float* Err, void* stream);
/* out[i] = silu(a[i]) * b[i] (LLaMAMLP, model.py:298-300) where fc_1 / fc_2 are separate layers (calibration forward of the quantiser). */
int lp_swiglu(const float* a, const float* b, float* out, size_t n, int round_bf16, void* stream);

/* ---- the whole single-token decode step (batch 1) as ONE persistent kernel ------------------------------------------
 * replaces GPT.forward for T == 1 (model.py:63-111 -> Block.forward 158-180 -> CausalSelfAttention.forward 194-254 -> MLP
 * 284-301): the caller describes the step once as an ordered table of ops — LINEAR (lp_norm_linear semantics, M = 1),
 * ATTENTION (lp_attn_decode_fused semantics) and, under tensor parallelism, EXCHANGE (lp_tp_allreduce_residual semantics) — and lp_decode_step then runs it with two launches (embedding-row prologue +
 * step kernel).  Inside the step kernel one TMA ring per SM streams the weights and the old K/V rows of ALL ops back to back
 * (they do not depend on the activations), so HBM stays busy while the consumers wait on the grid-wide arrival counter of an
 * op they depend on (csrc/decode_step.cu).
 *   dep: index of the latest earlier op whose output this op's input (x / qkv / attention partials) is — the op waits until
 *        every CTA has finished op `dep` (and therefore all ops before it); -1 = inputs of the step only.
 *   residual: must be covered by `dep`, be a step input, or be the output of an earlier LINEAR op with the same N (the same
 *        CTA then wrote the rows it reads); checked by lp_decode_step_plan.  residual == out (x += W . u in place) is the
 *        preferred form: such ops are split over the CTAs at 16 KB-stage granularity and accumulate with atomic adds, so two
 *        of them may overlap (parallel-residual blocks); the summation order of their partial sums is not fixed.  An op without a
 *        residual can use the same form on a zeroed buffer of its own (residual == out inside geom.zero_ptr .. + zero_bytes):
 *        every SM then streams the same number of weight bytes whatever the tile count (QKV: 768 tiles on 148 SMs = 6 vs 5.19).
 * Covers fp32-activation mode, bf16 / GPTQ-int4 (tile-major aux2) / bnb NF4 (tile-major absmax) / row-wise int8 weights, MHA / GQA / MQA with H <= #SMs, bf16 KV cache,
 * hs 64 / 128, batch 1; LP_ERR_UNSUPPORTED otherwise (callers then issue the per-op calls above). */
typedef enum { LP_STEP_LINEAR = 0, LP_STEP_ATTENTION = 1, LP_STEP_EXCHANGE = 2, LP_STEP_SLAB = 3 } lp_step_kind;

typedef struct {
  int32_t kind;            /* lp_step_kind */
  int32_t dep;
  /* LINEAR: out = epilogue(norm(x) . W^T + bias [, residual]); x_is_attention: x is the output of the preceding ATTENTION
   * op (its split partials are merged while the row is staged) */
  const lp_weight* W;
  const float* x;
  int32_t x_is_attention;
  int32_t norm_kind;       /* lp_norm_kind, or -1: none */
  const float* norm_w;
  const float* norm_b;
  float eps;
  int32_t epilogue;        /* lp_epilogue */
  const float* residual;
  float* out;
  /* ATTENTION: qkv [ (H+2G)*hs ] raw projection of this token, caches [G, max_seq, hs] bf16 */
  const float* qkv;
  void* k_cache;
  void* v_cache;
  /* EXCHANGE (tensor parallelism): out = residual + sum over the tp ranks r of the [E] partial of rank r — a one-shot all-reduce
   * over NVLink peer memory inside the step kernel, PUSH style with the flag in the data.  Sender = the row-parallel LINEAR op
   * right before it (epilogue NONE, `dep` of the exchange): it carries the same tp_* fields with tp_size > 0, and its tile epilogue
   * stores this rank's partial into every rank's symmetric buffer at tp_buf_offset (= slot base + tp_rank * E * 8 for the sender)
   * as 8-byte {value, epoch} pairs.  Receiver = this op: polls the pairs of its slice in its LOCAL buffer at tp_buf_offset (= slot
   * base; rank r's pairs at + r * E * 8) until they carry the slot's epoch, adds them in rank order, adds `residual`, writes `out`.
   * A slot therefore holds tp * E * 8 bytes.  `tp_state`: two zero-initialised uint32 per slot (epoch, -), owned by the library
   * afterwards.  At most two slots (state pointers), used alternately.  tp_pad_* are unused by this protocol.  Buffers and state
   * are NOT shared with lp_tp_allreduce_residual (the per-op pull kernel): give the two protocols disjoint regions. */
  const void* tp_buf_ptrs;
  const void* tp_pad_ptrs;
  void* tp_state;
  uint64_t tp_buf_offset;
  int32_t tp_pad_base, tp_rank, tp_size, reserved;
  /* SLAB (column -> row pairing inside the GPU): out += W[:, c] . v for the input columns c whose values v THIS CTA produced in
   * the op right before — slab_src 0: the SwiGLU outputs of the preceding LINEAR op (which sets keep_local: its epilogue leaves
   * them in shared memory, `out` of that op is not written; dep = -1); slab_src 1: the attention output of the CTA's head (the
   * preceding ATTENTION op; dep = that op: the P CTAs of a head wait for each other only).  Neither u nor the attention output
   * travels through L2 and neither pair is a grid-wide dependency; the partial rows are added to `out` with vector reductions.
   * W: the projection (GPTQ int4, group 128, packed aux2) for N, K, bias; slab_image / slab_meta from lp_decode_step_slab_build /
   * lp_decode_step_slab_layout. */
  const void* slab_image;
  const void* slab_meta;
  int32_t slab_src, keep_local;
} lp_step_op;

/* What one CTA of the step kernel streams for a SLAB op (filled by lp_decode_step_slab_layout; opaque to callers). */
typedef struct {
  int64_t off;
  int32_t nunits, units_a, nseg, row0, nrb, unit0, group_a, stage_bytes;
} lp_slab_meta;

/* Load time.  lp_decode_step_slab_layout: which input columns every CTA of the step kernel owns for a projection [N, K] fed by
 * src 0 (SwiGLU up-projection of `fc_tiles_or_heads` 16-row tiles, K = 8 * tiles) or src 1 (attention, `fc_tiles_or_heads` heads of
 * size hs); writes *n_ctas records (max_ctas >= #SMs) and the image size.  lp_decode_step_slab_build: gathers W (LP_W_INT4 row
 * major + packed tile-major aux2) into the per-CTA slab images (`meta_dev`: the records copied to the device). */
int lp_decode_step_slab_layout(int src, int N, int K, int fc_tiles_or_heads, int hs, lp_slab_meta* meta, int max_ctas, int* n_ctas,
                               size_t* image_bytes);
int lp_decode_step_slab_build(const lp_weight* W, const lp_slab_meta* meta_dev, int n_ctas, void* image, void* stream);

typedef struct {
  const int32_t* pos;        /* [1] device-side position of the token being decoded                               */
  const void* idx;           /* token ids (int32 / int64); the step embeds idx[idx_offset ? *idx_offset : 0]      */
  const int32_t* idx_offset;
  const void* wte;           /* embedding table [V, E] in wte_dtype (lp_dtype)                                    */
  float* x0;                 /* [E] residual stream: written by the prologue, input of the first op               */
  const float* cos;          /* RoPE tables fp32 [block_size, n_elem]                                             */
  const float* sin;
  void* workspace;           /* lp_decode_step_workspace_bytes(H, hs)                                             */
  size_t workspace_bytes;
  int32_t idx_is_int64, wte_dtype, E, H, G, hs, n_elem, max_seq, kv_dtype;
  float scale;               /* softmax scale, 1/sqrt(hs)                                                         */
  void* zero_ptr;            /* optional: zero_bytes (multiple of 16) cleared by the step's prologue before any op runs —      */
  size_t zero_bytes;         /* outputs of LINEAR ops that accumulate in place into a buffer of their own (see below)      */
} lp_step_geom;

typedef struct { uint64_t opaque[32]; } lp_step_handle;  /* filled by lp_decode_step_plan; plain data, copyable */

size_t lp_decode_step_plan_bytes(int n_ops);
size_t lp_decode_step_workspace_bytes(int H, int hs);
/* Load time (synchronous copy, not capturable): validates the table, builds the TMA descriptors and writes the device-side
 * op table into plan_dev (128-byte aligned, lp_decode_step_plan_bytes(n_ops)). */
int lp_decode_step_plan(const lp_step_op* ops, int n_ops, const lp_step_geom* geom, void* plan_dev, size_t plan_bytes,
                        lp_step_handle* handle);
/* One decode step: 2 launches, graph-capturable, no allocation, no synchronisation.
 * The step kernel is persistent (one CTA per SM) and its CTAs wait for each other through arrival counters, so the whole grid
 * must be co-resident: lp_decode_step_plan checks the occupancy (LP_ERR_UNSUPPORTED if the grid cannot fit) and, where the
 * device accepts it together with programmatic dependent launch (probed once per device), the kernel is launched with
 * cudaLaunchAttributeCooperative, i.e. the driver guarantees co-residency.  Every wait on another CTA or another GPU is bounded
 * (default 4 s, env LP_DS_TIMEOUT_MS, 0 = unbounded): when a bound expires the kernel records {code, op, dep, CTA, value seen}
 * in the plan's sticky error record, stops waiting everywhere and runs to its end — no hang, the results of that step are
 * garbage, and lp_decode_step_status reports it. */
int lp_decode_step(const lp_step_handle* handle, void* stream);
/* Synchronous health check of the steps launched so far with this plan (one 32-byte device-to-host copy): LP_OK, or
 * LP_ERR_TIMEOUT with info = {code (1: dependency counter, 2: tensor-parallel peer flag), op index, dep / peer rank, CTA, value
 * seen, 0, 0, 0}; the record is cleared when it has been reported.  info may be NULL.  generate() calls it once per call. */
int lp_decode_step_status(const lp_step_handle* handle, int32_t info[8]);
/* 1 if steps of this plan are launched cooperatively, 0 if the device refused the attribute combination (plain launch,
 * occupancy check + watchdog only). */
int lp_decode_step_cooperative(const lp_step_handle* handle);
/* Debug aid: device_buf (>= n_ops * 148 * 8 uint64) receives per-(op, CTA) globaltimer stamps [start, dependency met,
 * activations staged, end, x loaded, normalised, max|x| known, -]; NULL switches it off. */
int lp_debug_step_trace(void* device_buf);
/* Debug aid (env LP_GEMM_DEBUG=2 at load): {SM cycles, ns, k-blocks, 0} of the main loops of CTA 0's MMA issuer in the last
 * lp_gemm_bf16_tc launch (synchronises the device). */
int lp_debug_gemm_stats(long long* out4);

/* Tensor-parallel exchange fused with the residual add: out[n] = residual[n] + sum over ranks of partial_r[n], n fp32.
 * `buf_ptrs_dev` / `pad_ptrs_dev`: device arrays of `tp` peer-mapped addresses (symmetric buffer and signal pad of every
 * rank, e.g. torch.distributed._symmetric_memory: buffer_ptrs_dev / signal_pad_ptrs_dev).  This rank's partial must
 * already be at its symmetric buffer + buf_offset_bytes (the producing GEMV writes it there); callers alternate between
 * two slots (buf_offset_bytes, pad_base) so that a slot is not rewritten while a slower peer reads it.  `state`: two
 * zero-initialised uint32 per slot (epoch, ticket), owned by this library afterwards.  No reference counterpart: the
 * reference's multi-GPU path is FSDP (generate/base.py:187-205); the oracle is the single-device model. */
int lp_tp_allreduce_residual(const void* buf_ptrs_dev, const void* pad_ptrs_dev, int rank, int tp, size_t buf_offset_bytes,
                             int pad_base, void* state, int n, const float* residual, float* out, int round_bf16, void* stream);

/* replaces the sampling tail of generate() (generate/base.py:136-153): logits/temperature, top-k threshold
 * (ties with the k-th value survive), softmax, one multinomial draw (exponential race, Philox keyed by
 * (seed, *step)).  top_k == 1 is lowest-index arg-max.  logits fp32 [rows, V] -> token_out int32 [rows].
 * If pos_inout != NULL the kernel advances the device-side position: *pos_inout += 1 (and, when seq_buf != NULL,
 * rows == 1: seq_buf[*pos_inout + 1] = token first); *step += 1 once per launch for any row count (the noise is keyed by
 * (index, *step, row)) — so a captured decode step can be replayed without host round trips (generate/base.py:147-153). */
int lp_sample(const float* logits, int rows, int V, float temperature, int top_k, uint64_t seed, int32_t* step,
              int32_t* token_out, int32_t* seq_buf, int32_t* pos_inout, void* stream);
/* The same on bf16 logits [rows, V] — what GPT.forward returns for a bf16 checkpoint (model.py:111 in the parameter dtype): no
 * conversion pass in front of the sampler. */
int lp_sample_bf16(const void* logits_bf16, int rows, int V, float temperature, int top_k, uint64_t seed, int32_t* step,
                   int32_t* token_out, int32_t* seq_buf, int32_t* pos_inout, void* stream);

/* load-time: reference GPTQ storage (uint8 (N, K/2) with strides (1, N), quantize/gptq.py:216-222) ->
 * LP_W_INT4 layout.  dst holds N * lp_int4_row_bytes(K) bytes. */
size_t lp_int4_row_bytes(int K);
int lp_repack_gptq_int4(const uint8_t* quant_weight_colmajor, uint8_t* dst, int N, int K, void* stream);

#define LP_LINEAR_MAX_M 8

#ifdef __cplusplus
}
#endif
#endif /* LP_ABI_H */
