// Error reporting, device check and launch accounting for libdeepxi_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <map>
#include <string>
#include <vector>

namespace dxi {

static thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// No fallback path exists: anything that is not compute capability 10.x is refused.
int check_device() {
  static thread_local int cached_dev = -1, cached_rc = 0;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  cached_dev = dev;
  if (major != 10) {
    set_error("device %d has compute capability %d.x; libdeepxi_b200 is built for sm_100a only", dev, major);
    cached_rc = DXI_E_ARCH;
  } else {
    cached_rc = DXI_OK;
  }
  return cached_rc;
}

// ---- per-kernel timing ------------------------------------------------------------------------------
struct Span { cudaEvent_t a, b; int launches; };
static thread_local bool g_prof_on = false;
static thread_local std::map<std::string, std::vector<Span>>* g_spans = nullptr;

ProfScope::ProfScope(const char* key, cudaStream_t st, int launches) : span_(nullptr), st_(st) {
  if (!g_prof_on) return;
  if (!g_spans) g_spans = new std::map<std::string, std::vector<Span>>();
  Span sp{};
  sp.launches = launches;
  if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
  cudaEventRecord(sp.a, st);
  auto& v = (*g_spans)[key];
  v.push_back(sp);
  span_ = reinterpret_cast<void*>(static_cast<uintptr_t>(v.size()));   // index + 1
  key_ = key;
}
ProfScope::~ProfScope() {
  if (!span_) return;
  auto& v = (*g_spans)[key_];
  cudaEventRecord(v[reinterpret_cast<uintptr_t>(span_) - 1].b, st_);
}

}  // namespace dxi

extern "C" {
void dxi_profile_enable(int on) { dxi::g_prof_on = on != 0; }
int dxi_profile_read(const char* key, double* total_ms, int64_t* launches) {
  double ms = 0.0;
  int64_t n = 0;
  if (dxi::g_spans) {
    auto it = dxi::g_spans->find(key);
    if (it != dxi::g_spans->end()) {
      for (auto& sp : it->second) {
        float t = 0.0f;
        cudaEventSynchronize(sp.b);
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { ms += t; n += sp.launches; }
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
      }
      dxi::g_spans->erase(it);
    }
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return DXI_OK;
}
// Pinned host buffers for the serving loop.  write_combined = 1 (cudaHostAllocWriteCombined): memory the HOST only writes
// sequentially and the device reads (the noisy-speech input batches): not snooped during the DMA, slow for the host to read back.
int dxi_host_alloc(void** ptr, size_t bytes, int write_combined) {
  if (!ptr || bytes == 0) { dxi::set_error("dxi_host_alloc: bad argument"); return DXI_E_INVALID; }
  cudaError_t e = cudaHostAlloc(ptr, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) return dxi::cuda_fail(e, "cudaHostAlloc");
  return DXI_OK;
}
int dxi_host_free(void* ptr) {
  if (!ptr) return DXI_OK;
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) return dxi::cuda_fail(e, "cudaFreeHost");
  return DXI_OK;
}
const char* dxi_last_error(void) { return dxi::g_err; }
int dxi_version(void) { return 100; }
int dxi_device_check(void) { return dxi::check_device(); }
int64_t dxi_launch_count(void) { return dxi::g_launches; }
void dxi_launch_count_reset(void) { dxi::g_launches = 0; }
}
