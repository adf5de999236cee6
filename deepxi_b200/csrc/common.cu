// Error reporting, device check and launch accounting for libdeepxi_b200.so.
#include "common.cuh"
#include <stdarg.h>

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

}  // namespace dxi

extern "C" {
const char* dxi_last_error(void) { return dxi::g_err; }
int dxi_version(void) { return 100; }
int dxi_device_check(void) { return dxi::check_device(); }
int64_t dxi_launch_count(void) { return dxi::g_launches; }
void dxi_launch_count_reset(void) { dxi::g_launches = 0; }
}
