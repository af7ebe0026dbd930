// MMA family placeholder (filled in next): reports "unsupported" so lp_linear uses the FMA family.
#include "common.cuh"
namespace lp {
int init_linear_mma() { return LP_OK; }
int linear_mma_max_m() { return 32; }
int linear_mma(const float*, int, const lp_weight&, int, const float*, float*, int, void*) { return LP_ERR_UNSUPPORTED; }
}  // namespace lp
