// Error plumbing + device query for the C ABI (include/fedfr_b200.h).
#include "common.cuh"
#include <mutex>

namespace pfc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_arch_state[64];   // 0 unknown, 1 ok, -1 bad (per device)
static int g_sm_count[64];

int require_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device available: %s (fedfr_b200 has no CPU fallback)", cudaGetErrorString(e));
    return PFC_E_ARCH;
  }
  if (dev < 0 || dev >= 64) return PFC_E_ARCH;
  if (g_arch_state[dev] == 0) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
      return PFC_E_ARCH;
    }
    g_sm_count[dev] = p.multiProcessorCount;
    g_arch_state[dev] = (p.major == 10) ? 1 : -1;
    if (p.major != 10) set_error("device %d is sm_%d%d; fedfr_b200 kernels are built for sm_100a only", dev, p.major, p.minor);
  }
  return g_arch_state[dev] == 1 ? 0 : PFC_E_ARCH;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (g_arch_state[dev] == 0) require_sm100();
  return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

}  // namespace pfc

extern "C" {

int pfc_version(void) { return 100; }

const char* pfc_last_error(void) { return pfc::g_err; }

int pfc_query_device(int device, int* cc_major, int* cc_minor, int* sm_count) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    pfc::set_error("cudaGetDeviceProperties(%d) failed: %s (fedfr_b200 has no CPU fallback)", device, cudaGetErrorString(e));
    return (int)e;
  }
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (p.major != 10) {
    pfc::set_error("device %d is sm_%d%d; fedfr_b200 kernels are built for sm_100a only", device, p.major, p.minor);
    return PFC_E_ARCH;
  }
  return 0;
}

}  // extern "C"
