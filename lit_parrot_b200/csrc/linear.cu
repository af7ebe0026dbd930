// lp_linear: dispatch between the FMA family (exact fp32 CUDA-core math, any format) and the streaming family
// (persistent TMA-bulk ring + mma.sync; bf16 / int4 weights, decode batches).
#include <atomic>

#include "common.cuh"

namespace lp {
int linear_fma(const float* x, int M, const lp_weight& W, int epi, const float* residual, float* out, int round_bf16, void* stream);
struct NormArgs {
  const float* w;
  const float* b;
  float eps;
  int kind;
};
int linear_stream(const float* x, int M, const lp_weight& W, const NormArgs& nrm, int epi, const float* residual, float* out,
                  int round_bf16, void* stream);
void set_stream_trace(void* buf);
int dequant_bf16(const lp_weight& W, void* out, void* stream);
static std::atomic<int> g_path{0};  // 0 auto, 1 force FMA, 2 force streaming
}  // namespace lp

extern "C" {

int lp_set_linear_path(int path) {
  if (path < 0 || path > 2) return LP_ERR_INVALID_ARG;
  lp::g_path.store(path);
  return LP_OK;
}

int lp_debug_stream_trace(void* device_buf) {
  lp::set_stream_trace(device_buf);
  return LP_OK;
}

int lp_dequant_bf16(const lp_weight* W, void* out_bf16, void* stream) {
  if (!W || !W->w || !out_bf16 || W->N <= 0 || W->K <= 0) return LP_ERR_INVALID_ARG;
  return lp::dequant_bf16(*W, out_bf16, stream);
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
    // the streaming family takes the call when the whole activation block fits its column budget, else the FMA family does
    const lp::NormArgs none = {nullptr, nullptr, 0.f, -1};
    const int rc = lp::linear_stream(x, M, W, none, epilogue, residual, out, round_bf16, stream);
    if (rc != LP_ERR_UNSUPPORTED || path == 2) return rc;
  }
  for (int m0 = 0; m0 < M; m0 += LP_LINEAR_MAX_M) {
    const int mc = (M - m0) < LP_LINEAR_MAX_M ? (M - m0) : LP_LINEAR_MAX_M;
    int rc = lp::linear_fma(x + (size_t)m0 * W.K, mc, W, epilogue, residual ? residual + (size_t)m0 * W.N : nullptr,
                            out + (size_t)m0 * out_ld, round_bf16, stream);
    if (rc != LP_OK) return rc;
  }
  return LP_OK;
}


int lp_norm_linear(int norm_kind, const float* norm_w, const float* norm_b, float eps, const float* x, int M, const lp_weight* Wp,
                   int epilogue, const float* residual, float* out, int round_bf16, void* stream) {
  if (!x || !Wp || !out || !norm_w || M <= 0) return LP_ERR_INVALID_ARG;
  if (norm_kind != LP_NORM_LAYERNORM && norm_kind != LP_NORM_RMS) return LP_ERR_INVALID_ARG;
  const lp_weight& W = *Wp;
  if (!W.w || W.N <= 0 || W.K <= 0) return LP_ERR_INVALID_ARG;
  if (epilogue < LP_EPI_NONE || epilogue > LP_EPI_RESIDUAL) return LP_ERR_INVALID_ARG;
  if (epilogue == LP_EPI_RESIDUAL && !residual) return LP_ERR_INVALID_ARG;
  if (epilogue == LP_EPI_SWIGLU && (W.N & 1)) return LP_ERR_INVALID_ARG;
  if (round_bf16 || lp::g_path.load() == 1) return LP_ERR_UNSUPPORTED;  // bf16-faithful norm roundings live in lp_norm
  lp::NormArgs nrm = {norm_w, norm_b, eps, norm_kind};
  return lp::linear_stream(x, M, W, nrm, epilogue, residual, out, round_bf16, stream);
}

}  // extern "C"
