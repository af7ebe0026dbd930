// Library-level plumbing: version, status strings, one-time init, PDL switch.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace lp {

static thread_local char g_err[512] = "";
static std::atomic<int> g_pdl{-1};
static std::atomic<int> g_sms{0};
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  (void)cudaGetLastError();  // clear the sticky-less error so later calls report their own
}

bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* env = getenv("LP_PDL");
    v = (env && env[0] == '0') ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

int num_sms() {
  int v = g_sms.load(std::memory_order_relaxed);
  if (v == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;  // B200
    g_sms.store(v, std::memory_order_relaxed);
  }
  return v;
}

int init_attention();   // attention.cu
int init_sample();      // sample.cu

}  // namespace lp

extern "C" {

int lp_abi_version(void) { return LP_ABI_VERSION; }

const char* lp_status_str(int s) {
  switch (s) {
    case LP_OK: return "LP_OK";
    case LP_ERR_INVALID_ARG: return "LP_ERR_INVALID_ARG";
    case LP_ERR_UNSUPPORTED: return "LP_ERR_UNSUPPORTED";
    case LP_ERR_CUDA: return "LP_ERR_CUDA";
    case LP_ERR_WORKSPACE: return "LP_ERR_WORKSPACE";
    case LP_ERR_TIMEOUT: return "LP_ERR_TIMEOUT";
    default: return "LP_ERR_UNKNOWN";
  }
}

const char* lp_last_cuda_error(void) { return lp::g_err; }

unsigned long long lp_launch_count(void) { return lp::g_launches.load(); }

int lp_set_pdl(int enabled) {
  lp::g_pdl.store(enabled ? 1 : 0);
  return LP_OK;
}

int lp_init(int device) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  LP_CUDA_TRY(cudaSetDevice(device));
  int major = 0, minor = 0;
  LP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  LP_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    snprintf(lp::g_err, sizeof(lp::g_err), "liblitparrot_b200 is built for sm_100a only; device is sm_%d%d", major, minor);
    return LP_ERR_UNSUPPORTED;
  }
  lp::g_sms.store(0);
  (void)lp::num_sms();
  int rc = lp::init_attention();
  if (rc != LP_OK) return rc;
  return lp::init_sample();
}

}  // extern "C"
