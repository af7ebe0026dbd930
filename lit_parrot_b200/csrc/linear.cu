// lp_linear: dispatch between the FMA family (exact fp32 CUDA-core math, any format) and the MMA family
// (mma.sync tensor-core skinny GEMM; bf16 / int4 / int8 weights, up to 32 activation rows per pass).
#include <atomic>

#include "common.cuh"

namespace lp {
int linear_fma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16, void* stream);
int linear_mma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16, void* stream);
int linear_mma_max_m();
static std::atomic<int> g_path{0};  // 0 auto, 1 force FMA, 2 force MMA
}  // namespace lp

extern "C" {

int lp_set_linear_path(int path) {
  if (path < 0 || path > 2) return LP_ERR_INVALID_ARG;
  lp::g_path.store(path);
  return LP_OK;
}

int lp_linear(const float* x, int M, const lp_weight* Wp, int epilogue, const float* residual, float* out, int round_bf16,
              void* stream) {
  if (!x || !Wp || !out || M <= 0) return LP_ERR_INVALID_ARG;
  const lp_weight& W = *Wp;
  if (!W.w || W.N <= 0 || W.K <= 0) return LP_ERR_INVALID_ARG;
  if (epilogue < LP_EPI_NONE || epilogue > LP_EPI_RESIDUAL) return LP_ERR_INVALID_ARG;
  if (epilogue == LP_EPI_RESIDUAL && !residual) return LP_ERR_INVALID_ARG;
  if (epilogue == LP_EPI_SWIGLU && (W.N & 1)) return LP_ERR_INVALID_ARG;
  const int path = lp::g_path.load();
  const int out_ld = epilogue == LP_EPI_SWIGLU ? W.N / 2 : W.N;

  if (path != 1) {
    const int step = lp::linear_mma_max_m();
    int rc = LP_OK;
    bool ok = true;
    for (int m0 = 0; m0 < M && ok; m0 += step) {
      const int mc = (M - m0) < step ? (M - m0) : step;
      rc = lp::linear_mma(x + (size_t)m0 * W.K, mc, W, epilogue, residual ? residual + (size_t)m0 * W.N : nullptr,
                          out + (size_t)m0 * out_ld, round_bf16, stream);
      if (rc == LP_ERR_UNSUPPORTED && m0 == 0) ok = false;  // shape not covered: use the FMA family
      else if (rc != LP_OK) return rc;
    }
    if (ok) return LP_OK;
    if (path == 2) return LP_ERR_UNSUPPORTED;
  }
  for (int m0 = 0; m0 < M; m0 += LP_LINEAR_MAX_M) {
    const int mc = (M - m0) < LP_LINEAR_MAX_M ? (M - m0) : LP_LINEAR_MAX_M;
    int rc = lp::linear_fma(x + (size_t)m0 * W.K, mc, W, epilogue, residual ? residual + (size_t)m0 * W.N : nullptr,
                            out + (size_t)m0 * out_ld, round_bf16, stream);
    if (rc != LP_OK) return rc;
  }
  return LP_OK;
}

}  // extern "C"
