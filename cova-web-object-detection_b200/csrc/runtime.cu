// Error reporting + device queries of the C ABI (include/cova_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace cova {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm = 0, g_smem = 0;
static void query() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  cudaDeviceGetAttribute(&g_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}
int sm_count() {
  if (!g_sm) query();
  return g_sm ? g_sm : 148;
}

// Tuning knobs (include/cova_b200.h COVA_KNOB_*): plain ints read at launch time; -1 = kernel default.
static int g_knob[COVA_KNOB_COUNT] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
int knob(int id, int dflt) { return (id >= 0 && id < COVA_KNOB_COUNT && g_knob[id] >= 0) ? g_knob[id] : dflt; }
static unsigned long long* g_dbg = nullptr;
static long long g_dbg_words = 0;
unsigned long long* debug_words(long long need) { return (g_dbg && g_dbg_words >= need) ? g_dbg : nullptr; }

int max_smem_optin() {
  if (!g_smem) query();
  return g_smem ? g_smem : 232448;
}
}  // namespace cova

extern "C" int cova_abi_version(void) { return COVA_ABI_VERSION; }
extern "C" const char* cova_last_error(void) { return cova::g_err; }
extern "C" int cova_device_info(int* sm, int* smem) {
  int dev = 0;
  COVA_CUDA_OK(cudaGetDevice(&dev));
  if (sm) *sm = cova::sm_count();
  if (smem) *smem = cova::max_smem_optin();
  return COVA_OK;
}

extern "C" int cova_set_knob(int id, int value) {
  COVA_REQUIRE(id >= 0 && id < COVA_KNOB_COUNT, "cova_set_knob: unknown knob %d", id);
  cova::g_knob[id] = value;
  return COVA_OK;
}
extern "C" int cova_debug_buffer(void* dev_words, int64_t n_words) {
  cova::g_dbg = (unsigned long long*)dev_words;
  cova::g_dbg_words = dev_words ? n_words : 0;
  return COVA_OK;
}
