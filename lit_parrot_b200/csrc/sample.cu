// On-device sampling tail of generate() (generate/base.py:136-153).
//
//   logits / temperature -> keep everything >= the k-th largest value (ties survive, base.py:139-141)
//   -> softmax -> one multinomial draw.
// torch.multinomial on CUDA is an exponential race: argmax(p_i / E_i), E_i ~ Exp(1).  Taking logs, that
// is argmax(l_i - log E_i) over the kept logits — the softmax normalisation cancels, so the kernel never
// forms probabilities.  E_i comes from Philox4x32-10 keyed by (seed, step, i).  top_k == 1 (the
// reference's "greedy") is a plain arg-max with lowest-index tie-break.
// One CTA per row; the k-th value is found by an 8-bit radix select over order-preserving keys.
#include <math_constants.h>

#include "common.cuh"

namespace lp {

constexpr int SAMPLE_THREADS = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// logits as the model hands them out: fp32, or bf16 (the parameter dtype of a bf16 checkpoint: what GPT.forward returns)
__device__ __forceinline__ float lg_f(float v) { return v; }
__device__ __forceinline__ float lg_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename LT>
__global__ void __launch_bounds__(SAMPLE_THREADS)
sample_kernel(const LT* __restrict__ logits, int V, float temperature, int top_k, uint64_t seed, int* __restrict__ step,
              int* __restrict__ token_out, int* __restrict__ seq_buf, int* __restrict__ pos_inout, unsigned int* __restrict__ ticket) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_krem;
  __shared__ Best s_best[SAMPLE_THREADS / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x, tid = threadIdx.x;
  const LT* lg = logits + (size_t)row * V;
  const bool greedy = (top_k == 1);

  uint32_t thr_key = 0;  // keep everything
  if (!greedy && top_k > 0 && top_k < V) {
    // radix select of the top_k-th largest key, 8 bits per pass from the top
    if (tid == 0) {
      s_prefix = 0;
      s_krem = (unsigned)top_k;
    }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += SAMPLE_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = tid; i < V; i += SAMPLE_THREADS) {
        const uint32_t k = float_key(lg_f(lg[i]) / temperature);
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned krem = s_krem, cum = 0;
        int d = 255;
        for (; d > 0; --d) {
          if (cum + hist[d] >= krem) break;
          cum += hist[d];
        }
        s_krem = krem - cum;
        s_prefix = prefix | ((uint32_t)d << shift);
      }
      __syncthreads();
    }
    thr_key = s_prefix;
  }

  const int st = step ? *step : 0;
  Best best = {-CUDART_INF_F, 0x7fffffff};
  for (int i = tid; i < V; i += SAMPLE_THREADS) {
    float l = lg_f(lg[i]) / temperature;
    if (!greedy) {
      if (float_key(l) < thr_key) continue;
      const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)st, (uint32_t)row, 0u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
      l = l - logf(-logf(u));
    }
    best = better(best, Best{l, i});
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    best = better(best, other);
  }
  if ((tid & 31) == 0) s_best[tid >> 5] = best;
  __syncthreads();
  if (tid < 32) {
    best = s_best[tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = better(best, other);
    }
    if (tid == 0) {
      token_out[row] = best.i;
      if (row == 0 && pos_inout) {  // no other CTA of this kernel reads *pos_inout
        const int p = *pos_inout;
        if (seq_buf) seq_buf[p + 1] = best.i;
        *pos_inout = p + 1;
      }
      // the Philox step advances once per launch.  Every row read `st` before it took its ticket, so the LAST row to finish may
      // publish st + 1 without racing a slower row of this launch (a replayed batched step never reuses a noise key).
      if (step && (gridDim.x == 1 || atomicAdd(ticket, 1u) == gridDim.x - 1)) {
        if (gridDim.x > 1) *ticket = 0;
        *step = st + 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Greedy fast path (top_k == 1, the reference's "greedy"): the arg-max of a 50k-65k vocabulary is spread over
// GREEDY_CTAS CTAs per row; partial (value, index) pairs go to a small global scratch and the last CTA of the row (atomic
// ticket, self-resetting) reduces them and performs the same device-side bookkeeping as sample_kernel.  Lowest index
// wins ties, exactly like the single-CTA kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GREEDY_CTAS = 32;
constexpr int GREEDY_THREADS = 256;
constexpr int GREEDY_MAX_ROWS = 256;
// scratch of the arg-max reduction, one set PER DEVICE (lp_init(device) allocates the current device's)
static Best* g_greedy_part_dev[64] = {};          // [GREEDY_MAX_ROWS][GREEDY_CTAS]
static unsigned int* g_greedy_ticket_dev[64] = {};  // [GREEDY_MAX_ROWS]
static int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
static Best*& greedy_part() { return g_greedy_part_dev[cur_dev()]; }
static unsigned int*& greedy_ticket() { return g_greedy_ticket_dev[cur_dev()]; }

int init_sample() {
  if (greedy_part()) return LP_OK;
  LP_CUDA_TRY(cudaMalloc(&greedy_part(), sizeof(Best) * GREEDY_MAX_ROWS * GREEDY_CTAS));
  LP_CUDA_TRY(cudaMalloc(&greedy_ticket(), sizeof(unsigned int) * (GREEDY_MAX_ROWS + 1)));  // + 1: sample_kernel's row ticket
  LP_CUDA_TRY(cudaMemset(greedy_ticket(), 0, sizeof(unsigned int) * (GREEDY_MAX_ROWS + 1)));
  return LP_OK;
}

template <typename LT>
__global__ void __launch_bounds__(GREEDY_THREADS)
greedy_kernel(const LT* __restrict__ logits, int V, float temperature, int* __restrict__ step, int* __restrict__ token_out,
              int* __restrict__ seq_buf, int* __restrict__ pos_inout, Best* __restrict__ part, unsigned int* __restrict__ ticket) {
  __shared__ Best s_best[GREEDY_THREADS / 32];
  __shared__ int s_last;
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.y, tid = threadIdx.x;
  const LT* lg = logits + (size_t)row * V;
  Best best = {-CUDART_INF_F, 0x7fffffff};
  for (int i = blockIdx.x * GREEDY_THREADS + tid; i < V; i += GREEDY_CTAS * GREEDY_THREADS) best = better(best, Best{lg_f(lg[i]) / temperature, i});
  auto block_best = [&](Best b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = better(b, Best{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)});
    __syncthreads();
    if ((tid & 31) == 0) s_best[tid >> 5] = b;
    __syncthreads();
    b = tid < GREEDY_THREADS / 32 ? s_best[tid] : Best{-CUDART_INF_F, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = better(b, Best{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)});
    return b;
  };
  best = block_best(best);
  if (tid == 0) {
    part[row * GREEDY_CTAS + blockIdx.x] = best;
    __threadfence();
    s_last = (atomicAdd(ticket + row, 1u) == GREEDY_CTAS - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Best b = {-CUDART_INF_F, 0x7fffffff};
  if (tid < GREEDY_CTAS) {
    const volatile Best* pp = part + row * GREEDY_CTAS + tid;
    b = Best{pp->v, pp->i};
  }
  b = block_best(b);
  if (tid == 0) {
    ticket[row] = 0;
    token_out[row] = b.i;
    if (row == 0 && pos_inout) {
      const int p = *pos_inout;
      if (seq_buf) seq_buf[p + 1] = b.i;
      *pos_inout = p + 1;
    }
    if (step && row == 0) *step = *step + 1;  // nobody reads the step in the greedy kernel: row 0 advances it for any row count
  }
}

}  // namespace lp

template <typename LT>
static int sample_launch(const LT* logits, int rows, int V, float temperature, int top_k, uint64_t seed, int32_t* step, int32_t* token_out,
                         int32_t* seq_buf, int32_t* pos_inout, void* stream) {
  if (!logits || !token_out || rows <= 0 || V <= 0 || !(temperature > 0.f) || top_k < 0) return LP_ERR_INVALID_ARG;
  if (seq_buf && (rows != 1 || !pos_inout)) return LP_ERR_INVALID_ARG;  // pos_inout alone (any rows): just advance
  if (rows > 1 && step && !lp::greedy_ticket()) return LP_ERR_INVALID_ARG;  // lp_init() allocates the row ticket
  if (top_k == 1 && rows <= lp::GREEDY_MAX_ROWS && lp::greedy_part())
    return lp::launch(lp::greedy_kernel<LT>, dim3(lp::GREEDY_CTAS, rows), dim3(lp::GREEDY_THREADS), 0, stream, logits, V, temperature, step,
                      token_out, seq_buf, pos_inout, lp::greedy_part(), lp::greedy_ticket());
  return lp::launch(lp::sample_kernel<LT>, dim3(rows), dim3(lp::SAMPLE_THREADS), 0, stream, logits, V, temperature, top_k, seed, step,
                    token_out, seq_buf, pos_inout, lp::greedy_ticket() ? lp::greedy_ticket() + lp::GREEDY_MAX_ROWS : nullptr);
}

extern "C" int lp_sample(const float* logits, int rows, int V, float temperature, int top_k, uint64_t seed, int32_t* step,
                         int32_t* token_out, int32_t* seq_buf, int32_t* pos_inout, void* stream) {
  return sample_launch<float>(logits, rows, V, temperature, top_k, seed, step, token_out, seq_buf, pos_inout, stream);
}

extern "C" int lp_sample_bf16(const void* logits_bf16, int rows, int V, float temperature, int top_k, uint64_t seed, int32_t* step,
                              int32_t* token_out, int32_t* seq_buf, int32_t* pos_inout, void* stream) {
  return sample_launch<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(logits_bf16), rows, V, temperature, top_k, seed, step, token_out,
                                      seq_buf, pos_inout, stream);
}
