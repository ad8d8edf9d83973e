// Library-wide host plumbing of libnfk: error string, launch counter, device queries.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "nfk_common.cuh"

namespace nfk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_scan_order{1};  // Sklansky: matches torch.cumsum on CUDA bit for bit (tools/probe_exact.py)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NFK_ECUDA;
  }
  return NFK_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int scan_order() { return g_scan_order.load(std::memory_order_relaxed); }

}  // namespace nfk

extern "C" {

int nfk_abi_version(void) { return NFK_ABI_VERSION; }
const char* nfk_last_error(void) { return nfk::g_err; }
int64_t nfk_launch_count(void) { return nfk::g_launches.load(std::memory_order_relaxed); }
int nfk_set_scan_order(int order) {
  if (order < 0 || order > 2) {
    nfk::set_error("scan order must be 0, 1 or 2 (got %d)", order);
    return NFK_EINVAL;
  }
  nfk::g_scan_order.store(order, std::memory_order_relaxed);
  return NFK_OK;
}
int nfk_get_scan_order(void) { return nfk::scan_order(); }

}  // extern "C"
