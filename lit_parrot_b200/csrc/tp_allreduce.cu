// Tensor-parallel exchange step, fused with the residual add: out = residual + sum_r partial_r.
//
//   Llama-2-70b sharded over tp GPUs: every rank's row-/column-split projection produces a partial [rows, E];
//   the reference has no tensor parallelism (its multi-GPU path is FSDP), so the oracle is the single-device model.
//
// One-shot all-reduce over NVLink peer memory, no NCCL call on the path: the preceding GEMV writes this rank's partial
// straight into its slot of a SYMMETRIC buffer (same offset on every rank, peer-mapped by torch's symmetric memory);
// this kernel (a) publishes "slot written" to every peer (st.release.sys of a monotonically increasing epoch into the
// peer's signal pad), (b) waits until all peers have published the same epoch, (c) every CTA then loads its slice of ALL
// ranks' partials through the peer mappings (ld.volatile: point of coherence, no stale L1 line of the re-used slot) and
// adds them in rank order — every rank computes bit-identical sums — plus the residual.  Slots alternate (double
// buffering): a rank can run at most one exchange ahead of its slowest peer, so a slot is never overwritten while a peer
// still reads it.  The epoch lives in device memory and is advanced by the last CTA, so the kernel replays inside a CUDA graph.
#include "common.cuh"

namespace lp {

constexpr int TPA_THREADS = 256;
constexpr int TPA_MAX_TP = 8;

struct TpaParams {
  const unsigned long long* buf_ptrs;  // [tp] symmetric buffers (device array of peer-mapped addresses)
  const unsigned long long* pad_ptrs;  // [tp] signal pads
  unsigned int* state;                 // [2]: epoch of this slot, CTA ticket
  const float* residual;
  float* out;
  size_t buf_off;                      // bytes, this slot inside the symmetric buffer
  int pad_base;                        // first uint32 of this slot's tp flags inside the signal pad
  int rank, tp, n4, round_bf16;        // n4: float4 elements
};

__global__ void __launch_bounds__(TPA_THREADS) tp_allreduce_residual_kernel(const TpaParams p) {
  __shared__ unsigned int s_epoch;
  __shared__ int s_last;
  pdl_wait();  // this rank's partial (previous kernel) is complete and visible device-wide
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  if (tid == 0) s_epoch = p.state[0] + 1;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  if (tid < p.tp) {
    if (blockIdx.x == 0) {  // publish to peer `tid` (and to ourselves)
      __threadfence_system();
      unsigned int* flag = reinterpret_cast<unsigned int*>(p.pad_ptrs[tid]) + p.pad_base + p.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(flag), "r"(epoch) : "memory");
    }
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(p.pad_ptrs[p.rank]) + p.pad_base + tid;
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
    } while ((int)(v - epoch) < 0);
  }
  __syncthreads();
  const float4* src[TPA_MAX_TP];
#pragma unroll
  for (int r = 0; r < TPA_MAX_TP; ++r)
    src[r] = reinterpret_cast<const float4*>(p.buf_ptrs[r < p.tp ? r : 0] + p.buf_off);
  for (int i = blockIdx.x * TPA_THREADS + tid; i < p.n4; i += gridDim.x * TPA_THREADS) {
    float4 v[TPA_MAX_TP];
#pragma unroll
    for (int r = 0; r < TPA_MAX_TP; ++r) {
      if (r < p.tp)
        asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "l"(src[r] + i));
    }
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < TPA_MAX_TP; ++r) {
      if (r < p.tp) {
        s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w;
      }
    }
    if (p.round_bf16) {  // the sum of the shards plays the role of the (bf16) projection output
      s.x = bf16_round(s.x); s.y = bf16_round(s.y); s.z = bf16_round(s.z); s.w = bf16_round(s.w);
    }
    if (p.residual) {
      const float4 rv = reinterpret_cast<const float4*>(p.residual)[i];
      s.x = maybe_round(rv.x + s.x, p.round_bf16); s.y = maybe_round(rv.y + s.y, p.round_bf16);
      s.z = maybe_round(rv.z + s.z, p.round_bf16); s.w = maybe_round(rv.w + s.w, p.round_bf16);
    }
    reinterpret_cast<float4*>(p.out)[i] = s;
  }
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.state + 1, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && tid == 0) {
    p.state[0] = epoch;
    p.state[1] = 0;
  }
}

}  // namespace lp

extern "C" int lp_tp_allreduce_residual(const void* buf_ptrs_dev, const void* pad_ptrs_dev, int rank, int tp, size_t buf_offset_bytes,
                                        int pad_base, void* state, int n, const float* residual, float* out, int round_bf16,
                                        void* stream) {
  if (!buf_ptrs_dev || !pad_ptrs_dev || !state || !out || n <= 0 || tp < 1 || tp > lp::TPA_MAX_TP || rank < 0 || rank >= tp)
    return LP_ERR_INVALID_ARG;
  if (n % 4 || buf_offset_bytes % 16) return LP_ERR_UNSUPPORTED;
  lp::TpaParams p;
  p.buf_ptrs = reinterpret_cast<const unsigned long long*>(buf_ptrs_dev);
  p.pad_ptrs = reinterpret_cast<const unsigned long long*>(pad_ptrs_dev);
  p.state = reinterpret_cast<unsigned int*>(state);
  p.residual = residual;
  p.out = out;
  p.buf_off = buf_offset_bytes;
  p.pad_base = pad_base;
  p.rank = rank;
  p.tp = tp;
  p.n4 = n / 4;
  p.round_bf16 = round_bf16;
  int grid = (p.n4 + lp::TPA_THREADS - 1) / lp::TPA_THREADS;
  if (grid > 2 * lp::num_sms()) grid = 2 * lp::num_sms();
  return lp::launch(lp::tp_allreduce_residual_kernel, dim3(grid), dim3(lp::TPA_THREADS), 0, stream, p);
}
