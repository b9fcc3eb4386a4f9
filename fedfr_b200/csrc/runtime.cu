// Error plumbing + device query for the C ABI (include/fedfr_b200.h).
#include "common.cuh"
#include <mutex>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>      // header-only: resolves the injection library at run time, nothing to link

namespace pfc {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

// ---- per-phase profiling (off by default) ----
static int g_prof_on = 0;
struct ProfSpan { cudaEvent_t a, b; int phase; };
static ProfSpan g_spans[4096];
static int g_n_spans = 0;
static cudaEvent_t g_open[PH_COUNT];
bool prof_enabled() { return g_prof_on != 0; }
void nvtx_set(int on);
void prof_begin(int phase, cudaStream_t st) {
  if (!g_prof_on || g_n_spans >= 4096) return;
  cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); g_open[phase] = e;
}
void prof_end(int phase, cudaStream_t st) {
  if (!g_prof_on || g_n_spans >= 4096) return;
  cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
  g_spans[g_n_spans++] = ProfSpan{g_open[phase], e, phase};
}

static int g_nvtx = -1;          // -1: read FEDFR_NVTX on first use
static inline bool nvtx_on() {
  if (g_nvtx < 0) { const char* e = getenv("FEDFR_NVTX"); g_nvtx = (e && atoi(e) != 0) ? 1 : 0; }
  return g_nvtx == 1;
}
void nvtx_push(const char* name) { if (nvtx_on()) nvtxRangePushA(name); }
void nvtx_pop() { if (nvtx_on()) nvtxRangePop(); }
void nvtx_set(int on) { g_nvtx = on ? 1 : 0; }

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

long long pfc_launch_count(void) { return pfc::g_launch_count; }

/* profiling: enable(1)/disable(0); collect() synchronises, sums the spans per phase (ms) and their counts, and resets */
int pfc_profile_enable(int on) { pfc::g_prof_on = on; return 0; }
int pfc_profile_collect(float* ms_out /*[5]*/, int* count_out /*[5]*/) {
  for (int i = 0; i < pfc::PH_COUNT; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
  for (int i = 0; i < pfc::g_n_spans; ++i) {
    float ms = 0.f;
    cudaEventSynchronize(pfc::g_spans[i].b);
    cudaEventElapsedTime(&ms, pfc::g_spans[i].a, pfc::g_spans[i].b);
    ms_out[pfc::g_spans[i].phase] += ms;
    count_out[pfc::g_spans[i].phase] += 1;
    cudaEventDestroy(pfc::g_spans[i].a);
    cudaEventDestroy(pfc::g_spans[i].b);
  }
  pfc::g_n_spans = 0;
  return 0;
}

/* NVTX ranges around every compute entry of this library (default: the FEDFR_NVTX environment variable) */
int pfc_set_nvtx(int on) { pfc::nvtx_set(on); return 0; }

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
